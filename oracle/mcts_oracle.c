/*
 * oracle/mcts_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement, in plain C, of the Monte Carlo hot path of
 * jensbrodersen/mc-trade-sizing: run_all_strategies() and everything it calls
 * (reference: src/trading_models.py:43-168, 170-404, 445-530), including the
 * parts of NumPy's legacy RandomState the path consumes (MT19937 seeding,
 * random_sample, choice(p=...), shuffle), so that a seeded run of this file
 * reproduces a seeded run of the reference bit for bit.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this.  The product (libmcts.so) never does.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every entry
 * point below against fixtures under tests/golden/ that were produced by
 * importing the live reference module in the build container
 * (tests/golden/make_golden.py), covering all four generator modes, the
 * post-shuffle sequences, per-path (profit, drawdown) for all 20 strategies
 * and the final summary rows.  The reference itself ships no tests or golden
 * vectors.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off, no -ffast-math).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define ORACLE_MAX_REGIMES 16
#define ORACLE_NSTRAT 20

typedef struct {
    int length;
    double hit_rate, avg_win, avg_loss;
} oracle_regime;

typedef struct {
    int mode; /* 0 = "No Markov" (3-phase dynamic), 1 = Markov-1, 2 = Markov-2, 3 = regime */
    int num_trades;
    int num_simulations;
    int num_mc_shuffles;
    double hit_rate, avg_win, avg_loss;
    double p_win_after_win, p_win_after_loss;
    double p_win_ww, p_win_wl, p_win_lw, p_win_ll;
    int n_regimes; /* < 0: regimes=None (defaults of trading_models.py:108-112) */
    oracle_regime regimes[ORACLE_MAX_REGIMES];
} oracle_params;

enum { ORACLE_OK = 0, ORACLE_EINVAL = -1, ORACLE_EEMPTY = -2, ORACLE_ENOMEM = -3 };

/* ------------------------------------------------------------------------- */
/* NumPy legacy RandomState: MT19937 (numpy/random/src/mt19937/mt19937.c).    */
/* The reference draws from the global RandomState at                         */
/* trading_models.py:55,65,70,77,79,89,101,119,492.                           */
/* ------------------------------------------------------------------------- */
typedef struct {
    uint32_t key[624];
    int pos;
} mt_state;

static void mt_seed(mt_state *s, uint32_t seed)
{
    /* np.random.seed(int) -> _legacy_seeding -> mt19937_seed (Knuth LCG fill) */
    for (int pos = 0; pos < 624; pos++) {
        s->key[pos] = seed;
        seed = 1812433253u * (seed ^ (seed >> 30)) + (uint32_t)pos + 1u;
    }
    s->pos = 624;
}

static void mt_gen(mt_state *s)
{
    const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MAT = 0x9908b0dfu;
    uint32_t *k = s->key, y;
    int i;
    for (i = 0; i < 624 - 397; i++) {
        y = (k[i] & UPPER) | (k[i + 1] & LOWER);
        k[i] = k[i + 397] ^ (y >> 1) ^ ((y & 1u) ? MAT : 0u);
    }
    for (; i < 623; i++) {
        y = (k[i] & UPPER) | (k[i + 1] & LOWER);
        k[i] = k[i + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? MAT : 0u);
    }
    y = (k[623] & UPPER) | (k[0] & LOWER);
    k[623] = k[396] ^ (y >> 1) ^ ((y & 1u) ? MAT : 0u);
    s->pos = 0;
}

static inline uint32_t mt_u32(mt_state *s)
{
    if (s->pos == 624) mt_gen(s);
    uint32_t y = s->key[s->pos++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

/* random_sample(): 53-bit uniform in [0,1) from two 32-bit outputs */
static inline double mt_double(mt_state *s)
{
    uint32_t a = mt_u32(s) >> 5, b = mt_u32(s) >> 6;
    return (a * 67108864.0 + b) / 9007199254740992.0;
}

/* legacy random_interval(max): masked rejection, 32-bit draws when max fits */
static inline uint64_t mt_interval(mt_state *s, uint64_t max)
{
    if (max == 0) return 0;
    uint64_t mask = max, value;
    mask |= mask >> 1;
    mask |= mask >> 2;
    mask |= mask >> 4;
    mask |= mask >> 8;
    mask |= mask >> 16;
    mask |= mask >> 32;
    if (max <= 0xffffffffull) {
        while ((value = (mt_u32(s) & mask)) > max) {
        }
    } else {
        do {
            uint64_t hi = mt_u32(s), lo = mt_u32(s);
            value = ((hi << 32) | lo) & mask;
        } while (value > max);
    }
    return value;
}

/* RandomState.shuffle on a 1-d ndarray: Fisher-Yates from the top
 * (trading_models.py:492). */
static void mt_shuffle(mt_state *s, double *x, int64_t n)
{
    for (int64_t i = n - 1; i >= 1; i--) {
        int64_t j = (int64_t)mt_interval(s, (uint64_t)i);
        double t = x[i];
        x[i] = x[j];
        x[j] = t;
    }
}

/* RandomState.choice([a0, a1], size=l, p=[p0, p1]) with replacement:
 * cdf = cumsum(p); cdf /= cdf[-1]; idx = searchsorted(cdf, u, side='right').
 * Returns ORACLE_EINVAL where NumPy raises ValueError (negative / non-normalised p). */
static int mt_choice2(mt_state *s, double a0, double a1, double p0, double p1, int64_t l, double *out)
{
    if (!(p0 >= 0.0) || !(p1 >= 0.0)) return ORACLE_EINVAL; /* also catches NaN */
    if (fabs((p0 + p1) - 1.0) > 1.4901161193847656e-08) return ORACLE_EINVAL;
    double c0 = p0, c1 = p0 + p1;
    c0 /= c1;
    c1 /= c1;
    for (int64_t i = 0; i < l; i++) {
        double u = mt_double(s);
        int idx = (c0 <= u) + (c1 <= u); /* side='right': number of cdf entries <= u */
        if (idx > 1) idx = 1;            /* unreachable: c1 == 1.0 > u */
        out[i] = idx == 0 ? a0 : a1;
    }
    return ORACLE_OK;
}

/* ------------------------------------------------------------------------- */
/* Generators (trading_models.py:43-128). Each returns the produced length.   */
/* ------------------------------------------------------------------------- */

/* simulate_trades_dynamic, :43-66 */
static int64_t gen_dynamic(mt_state *s, const oracle_params *p, double *out, int *err)
{
    int n = p->num_trades;
    int len[3] = {(int)(n * 0.2), (int)(n * 0.2), n - (int)(n * 0.4)};
    double hr[3] = {fmin(p->hit_rate + 0.2, 1.0), fmax(p->hit_rate - 0.3, 0.05), p->hit_rate};
    double aw[3] = {p->avg_win * 1.1, p->avg_win * 0.9, p->avg_win};
    double al[3] = {p->avg_loss * 0.9, p->avg_loss * 1.1, p->avg_loss};
    int64_t pos = 0;
    int left = n;
    for (int k = 0; k < 3; k++) {
        int l = len[k] < left ? len[k] : left;
        if (l <= 0) continue;
        if ((*err = mt_choice2(s, aw[k], -al[k], hr[k], 1 - hr[k], l, out + pos))) return 0;
        pos += l;
        left -= l;
        if (left <= 0) break;
    }
    if (left > 0) { /* :64-65 tail fill at base parameters */
        if ((*err = mt_choice2(s, p->avg_win, -p->avg_loss, p->hit_rate, 1 - p->hit_rate, left, out + pos)))
            return 0;
        pos += left;
    }
    return pos;
}

/* simulate_trades_markov, :68-85 */
static int64_t gen_markov1(mt_state *s, const oracle_params *p, double *out)
{
    int last_win = mt_double(s) < p->hit_rate;
    int64_t pos = 0;
    out[pos++] = last_win ? p->avg_win : -p->avg_loss;
    for (int t = 1; t < p->num_trades; t++) {
        int win = last_win ? (mt_double(s) < p->p_win_after_win) : (mt_double(s) < p->p_win_after_loss);
        out[pos++] = win ? p->avg_win : -p->avg_loss;
        last_win = win;
    }
    return pos;
}

/* simulate_trades_markov2, :87-104 (always emits the first two trades) */
static int64_t gen_markov2(mt_state *s, const oracle_params *p, double *out)
{
    int older = mt_double(s) < p->hit_rate;
    int newer = mt_double(s) < p->hit_rate;
    int64_t pos = 0;
    out[pos++] = older ? p->avg_win : -p->avg_loss;
    out[pos++] = newer ? p->avg_win : -p->avg_loss;
    for (int t = 2; t < p->num_trades; t++) {
        double pw;
        if (older && newer) pw = p->p_win_ww;
        else if (older && !newer) pw = p->p_win_wl;
        else if (!older && newer) pw = p->p_win_lw;
        else pw = p->p_win_ll;
        int win = mt_double(s) < pw;
        out[pos++] = win ? p->avg_win : -p->avg_loss;
        older = newer;
        newer = win;
    }
    return pos;
}

/* simulate_trades_regime_switch, :106-128 (no tail fill; ignores hit_rate/avg_*) */
static int64_t gen_regime(mt_state *s, const oracle_params *p, double *out, int *err)
{
    int n = p->num_trades;
    oracle_regime def[3] = {{(int)(n * 0.3), 0.9, 200, 100}, {(int)(n * 0.2), 0.5, 100, 100}, {n - (int)(n * 0.5), 0.2, 100, 200}};
    const oracle_regime *rg = p->n_regimes < 0 ? def : p->regimes;
    int nr = p->n_regimes < 0 ? 3 : p->n_regimes;
    int64_t pos = 0;
    int left = n;
    for (int k = 0; k < nr; k++) {
        int l = rg[k].length < left ? rg[k].length : left;
        if (l <= 0) continue;
        if ((*err = mt_choice2(s, rg[k].avg_win, -rg[k].avg_loss, rg[k].hit_rate, 1 - rg[k].hit_rate, l, out + pos)))
            return 0;
        pos += l;
        left -= l;
        if (left <= 0) break;
    }
    return pos;
}

static int64_t seq_capacity(const oracle_params *p)
{
    int64_t n = p->num_trades > 0 ? p->num_trades : 0;
    return n + 2;
}

/* ------------------------------------------------------------------------- */
/* NumPy reductions used by the path.                                         */
/* ------------------------------------------------------------------------- */

/* DOUBLE_pairwise_sum (numpy/_core/src/umath/loops_utils.h.src), contiguous input */
static double np_pairwise(const double *a, int64_t n)
{
    if (n < 8) {
        double res = 0.;
        for (int64_t i = 0; i < n; i++) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8];
        int64_t i;
        for (int j = 0; j < 8; j++) r[j] = a[j];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return np_pairwise(a, n2) + np_pairwise(a + n2, n - n2);
    }
}

/* np.sum over a contiguous float64 array: add.reduce starts from the identity 0.0 */
static double np_sum(const double *a, int64_t n) { return 0.0 + np_pairwise(a, n); }

/* calculate_drawdown(np.cumsum(contrib)) and np.sum(contrib): :130-133,:136-137,:167-168 */
static void profit_and_drawdown(const double *contrib, int64_t n, double *profit, double *dd)
{
    double equity = 0, peak = 0, worst = 0;
    for (int64_t i = 0; i < n; i++) {
        equity = i == 0 ? contrib[0] : equity + contrib[i]; /* np.cumsum: strictly sequential */
        peak = i == 0 ? equity : (equity > peak ? equity : peak); /* np.maximum.accumulate */
        double d = equity - peak;
        worst = i == 0 ? d : (d < worst ? d : worst); /* np.min */
    }
    *profit = np_sum(contrib, n);
    *dd = worst;
}

/* ------------------------------------------------------------------------- */
/* The 20 position-sizing closures, restated literally (make_condition_func,  */
/* trading_models.py:170-404) on the state dict of strategy_dynamic (:144).   */
/* ------------------------------------------------------------------------- */
typedef struct {
    int win_streak, loss_streak;
    int paused; /* state['mode'] == 'pause' */
    double last_result, last2_result;
} strat_state;

static int condition(int id, double result, int size, strat_state *st)
{
    switch (id) {
    case 1: size = 1; break;
    case 2: size = result > 0 ? 2 : 1; break;
    case 3: case 4: case 5: { /* :181-214, thresholds 2,3,4 */
        int thr = id - 1;
        if (result > 0) {
            st->win_streak += 1;
            size = st->win_streak >= thr ? 1 : 2;
        } else {
            size = 1;
            st->win_streak = 0;
        }
        break;
    }
    case 6: size = result <= 0 ? 2 : 1; break;
    case 7: case 8: { /* :221-236, thresholds 2,3 */
        int thr = id - 5;
        if (result > 0) {
            size = 1;
            st->loss_streak = 0;
        } else {
            st->loss_streak += 1;
            size = st->loss_streak >= thr ? 2 : 1;
        }
        break;
    }
    case 9: /* :239-249 */
        if (!st->paused) {
            if (result > 0) st->paused = 1;
            size = 1;
        } else {
            if (result <= 0) {
                st->paused = 0;
                size = 1;
            } else
                size = 0;
        }
        break;
    case 10: case 11: case 12: { /* :252-303, thresholds 2,3,4 */
        int thr = id - 8;
        if (!st->paused) {
            if (result > 0) {
                st->win_streak += 1;
                if (st->win_streak >= thr) st->paused = 1;
            } else
                st->win_streak = 0;
            size = 1;
        } else {
            if (result <= 0) {
                st->paused = 0;
                st->win_streak = 0;
                size = 1;
            } else
                size = 0;
        }
        break;
    }
    case 13: /* :306-314 (size carries over when neither threshold fires) */
        if (result > 0) {
            st->win_streak += 1;
            if (st->win_streak >= 2) size = 2;
        } else {
            st->loss_streak += 1;
            if (st->loss_streak >= 2) size = 1;
        }
        break;
    case 14: case 17: /* :317-321, :364-368 */
        size = (result > 0 && st->last2_result <= 0) ? 2 : 1;
        break;
    case 15: /* :324-342 */
        if (!st->paused) {
            if (result > 0) {
                st->win_streak += 1;
                if (st->win_streak >= 2) {
                    st->paused = 1;
                    size = 2;
                } else
                    size = 1;
            } else {
                st->win_streak = 0;
                size = 1;
            }
        } else {
            if (result <= 0) {
                st->paused = 0;
                st->win_streak = 0;
                size = 1;
            } else
                size = 0;
        }
        break;
    case 16: /* :345-361 */
        if (!st->paused) {
            if (result > 0) {
                st->paused = 1;
                size = 1;
            } else {
                st->loss_streak += 1;
                size = st->loss_streak >= 2 ? 2 : 1;
            }
        } else {
            if (result <= 0) {
                st->paused = 0;
                size = 1;
            } else
                size = 0;
        }
        break;
    case 18: /* :371-378 (size carries over below the threshold) */
        if (result > 0) {
            st->win_streak += 1;
            if (st->win_streak >= 3) size = 3;
        } else {
            size = 1;
            st->win_streak = 0;
        }
        break;
    case 19: /* :381-387 */
        if (st->win_streak >= 2) size = 2;
        else if (st->loss_streak >= 2) size = 3;
        else size = 1;
        break;
    case 20: /* :390-398 (the second `result > 0` arm is dead code) */
        if (result > 0) size = 1;
        else if (st->loss_streak >= 2) size = 3;
        else size = 1;
        break;
    default: size = 1;
    }
    return size;
}

/* strategy_static (:135-139) for id 1, strategy_dynamic (:141-168) otherwise.
 * `contrib` is scratch of length n. */
static void eval_strategy(int id, const double *results, int64_t n, double *contrib, double *profit, double *dd)
{
    if (id == 1) {
        profit_and_drawdown(results, n, profit, dd);
        return;
    }
    int size = 1;
    strat_state st = {0, 0, 0, 0.0, 0.0};
    for (int64_t i = 0; i < n; i++) {
        double trade_result = results[i] * (double)size;
        contrib[i] = st.paused ? 0.0 : trade_result; /* :148-151, integer 0 -> +0.0 */
        if (results[i] > 0) { /* :154-159 */
            st.win_streak += 1;
            st.loss_streak = 0;
        } else {
            st.loss_streak += 1;
            st.win_streak = 0;
        }
        st.last2_result = st.last_result; /* :162-163 */
        st.last_result = results[i];
        size = condition(id, results[i], size, &st); /* :165 */
    }
    profit_and_drawdown(contrib, n, profit, dd);
}

/* ------------------------------------------------------------------------- */
/* Aggregates (:509-522), id order (the ranking of :524-528 is done by the    */
/* caller in Python, it is a stable list sort).                               */
/* summary row = avg_profit, avg_drawdown, ratio, min_profit, max_profit,     */
/*               min_dd, max_dd, avg_per_trade, ratio_max_dd                  */
/* ------------------------------------------------------------------------- */
static void summarise(const double *profit, const double *dd, int64_t P, int num_trades, double *row)
{
    double avg_p = np_sum(profit, P) / (double)P;
    double avg_d = np_sum(dd, P) / (double)P;
    double min_p = profit[0], max_p = profit[0], min_d = dd[0], max_d = dd[0];
    for (int64_t i = 1; i < P; i++) {
        if (profit[i] < min_p) min_p = profit[i];
        if (profit[i] > max_p) max_p = profit[i];
        if (dd[i] < min_d) min_d = dd[i];
        if (dd[i] > max_d) max_d = dd[i];
    }
    row[0] = avg_p;
    row[1] = avg_d;
    row[2] = avg_d != 0 ? avg_p / fabs(avg_d) : INFINITY;
    row[3] = min_p;
    row[4] = max_p;
    row[5] = min_d;
    row[6] = max_d;
    row[7] = avg_p / (double)num_trades;
    row[8] = max_d != 0 ? avg_p / fabs(max_d) : INFINITY;
}

/* ------------------------------------------------------------------------- */
/* Entry points                                                               */
/* ------------------------------------------------------------------------- */

/* Length of the trade sequence the selected generator returns (it is not
 * always num_trades: Markov-2 emits >= 2, Markov-1 >= 1, regime can be short). */
int64_t oracle_sequence_length(const oracle_params *p)
{
    int n = p->num_trades;
    if (p->mode == 3) {
        oracle_regime def[3] = {{(int)(n * 0.3), 0, 0, 0}, {(int)(n * 0.2), 0, 0, 0}, {n - (int)(n * 0.5), 0, 0, 0}};
        const oracle_regime *rg = p->n_regimes < 0 ? def : p->regimes;
        int nr = p->n_regimes < 0 ? 3 : p->n_regimes;
        int64_t pos = 0;
        int left = n;
        for (int k = 0; k < nr; k++) {
            int l = rg[k].length < left ? rg[k].length : left;
            if (l <= 0) continue;
            pos += l;
            left -= l;
            if (left <= 0) break;
        }
        return pos;
    }
    if (p->mode == 2) return n > 2 ? n : 2;
    if (p->mode == 1) return n > 1 ? n : 1;
    return n > 0 ? n : 0;
}

/*
 * run_all_strategies(...) after np.random.seed(seed)  (trading_models.py:445-522).
 *   summary : [20][9] rows in strategy-id order
 *   profit  : optional [20][P] per-path profits, P = num_simulations*num_mc_shuffles
 *   dd      : optional [20][P] per-path drawdowns
 *   seqs    : optional [P][L] post-shuffle sequences, L = oracle_sequence_length()
 */
int oracle_run_all(const oracle_params *p, uint32_t seed, double *summary, double *profit, double *dd, double *seqs)
{
    if (p->num_simulations <= 0 || p->num_mc_shuffles <= 0) return ORACLE_EEMPTY;
    if (p->mode == 3 && p->n_regimes > ORACLE_MAX_REGIMES) return ORACLE_EINVAL;
    int64_t L = oracle_sequence_length(p);
    if (L <= 0) return ORACLE_EEMPTY; /* np.min of an empty array raises */
    int64_t P = (int64_t)p->num_simulations * p->num_mc_shuffles;
    int own = profit == NULL || dd == NULL;
    double *pp = profit, *dp = dd;
    if (own) {
        pp = (double *)malloc(sizeof(double) * ORACLE_NSTRAT * P);
        dp = (double *)malloc(sizeof(double) * ORACLE_NSTRAT * P);
    }
    double *base = (double *)malloc(sizeof(double) * seq_capacity(p));
    double *contrib = (double *)malloc(sizeof(double) * seq_capacity(p));
    if (!pp || !dp || !base || !contrib) return ORACLE_ENOMEM;
    mt_state rng;
    mt_seed(&rng, seed);
    int err = 0;
    int64_t path = 0;
    for (int sim = 0; sim < p->num_simulations && !err; sim++) {
        int64_t n;
        if (p->mode == 3) n = gen_regime(&rng, p, base, &err);
        else if (p->mode == 2) n = gen_markov2(&rng, p, base);
        else if (p->mode == 1) n = gen_markov1(&rng, p, base);
        else n = gen_dynamic(&rng, p, base, &err);
        if (err) break;
        for (int sh = 0; sh < p->num_mc_shuffles; sh++, path++) {
            mt_shuffle(&rng, base, n); /* :492, in place and cumulative */
            if (seqs) memcpy(seqs + path * L, base, sizeof(double) * n);
            for (int id = 1; id <= ORACLE_NSTRAT; id++)
                eval_strategy(id, base, n, contrib, &pp[(int64_t)(id - 1) * P + path], &dp[(int64_t)(id - 1) * P + path]);
        }
    }
    if (!err && summary)
        for (int id = 0; id < ORACLE_NSTRAT; id++)
            summarise(pp + (int64_t)id * P, dp + (int64_t)id * P, P, p->num_trades, summary + 9 * id);
    free(base);
    free(contrib);
    if (own) {
        free(pp);
        free(dp);
    }
    return err;
}

/*
 * Replay: evaluate caller-supplied sequences values[P][T] (row-major) with all
 * 20 strategies.  profit/dd are [20][P]; summary is optional [20][9] and uses
 * `num_trades` for avg_per_trade exactly as :519 does.
 */
int oracle_eval_paths(const double *values, int64_t P, int64_t T, int num_trades, double *profit, double *dd, double *summary)
{
    if (P <= 0 || T <= 0) return ORACLE_EEMPTY;
    double *contrib = (double *)malloc(sizeof(double) * T);
    if (!contrib) return ORACLE_ENOMEM;
    for (int64_t path = 0; path < P; path++)
        for (int id = 1; id <= ORACLE_NSTRAT; id++)
            eval_strategy(id, values + path * T, T, contrib, &profit[(int64_t)(id - 1) * P + path], &dd[(int64_t)(id - 1) * P + path]);
    free(contrib);
    if (summary)
        for (int id = 0; id < ORACLE_NSTRAT; id++)
            summarise(profit + (int64_t)id * P, dd + (int64_t)id * P, P, num_trades, summary + 9 * id);
    return ORACLE_OK;
}

/*
 * Equity curves (the np.cumsum temporaries of :136 / :167): curves[20][P][T], sequential fp64 prefix
 * sums of each strategy's contributions.  Checker for the curve-export extension.
 */
int oracle_equity_curves(const double *values, int64_t P, int64_t T, double *curves)
{
    if (P <= 0 || T <= 0) return ORACLE_EEMPTY;
    double *contrib = (double *)malloc(sizeof(double) * T);
    if (!contrib) return ORACLE_ENOMEM;
    for (int64_t path = 0; path < P; path++) {
        const double *results = values + path * T;
        for (int id = 1; id <= ORACLE_NSTRAT; id++) {
            double *out = curves + ((int64_t)(id - 1) * P + path) * T;
            if (id == 1) {
                for (int64_t i = 0; i < T; i++) contrib[i] = results[i];
            } else {
                int size = 1;
                strat_state st = {0, 0, 0, 0.0, 0.0};
                for (int64_t i = 0; i < T; i++) {
                    double trade_result = results[i] * (double)size;
                    contrib[i] = st.paused ? 0.0 : trade_result;
                    if (results[i] > 0) {
                        st.win_streak += 1;
                        st.loss_streak = 0;
                    } else {
                        st.loss_streak += 1;
                        st.win_streak = 0;
                    }
                    st.last2_result = st.last_result;
                    st.last_result = results[i];
                    size = condition(id, results[i], size, &st);
                }
            }
            double equity = 0;
            for (int64_t i = 0; i < T; i++) {
                equity = i == 0 ? contrib[0] : equity + contrib[i];
                out[i] = equity;
            }
        }
    }
    free(contrib);
    return ORACLE_OK;
}

/* Aggregate rows from per-path arrays only (np.mean/np.min/np.max restatement). */
int oracle_summarise(const double *profit, const double *dd, int64_t P, int num_trades, double *summary)
{
    if (P <= 0) return ORACLE_EEMPTY;
    for (int id = 0; id < ORACLE_NSTRAT; id++)
        summarise(profit + (int64_t)id * P, dd + (int64_t)id * P, P, num_trades, summary + 9 * id);
    return ORACLE_OK;
}

/* np.sum restatement exposed for the host-side tree planner tests. */
double oracle_np_sum(const double *a, int64_t n) { return np_sum(a, n); }

/* Raw RNG taps, so tests can pin the MT19937 restatement against NumPy itself. */
void oracle_rng_taps(uint32_t seed, int n_doubles, double *doubles, int shuffle_n, double *shuffled)
{
    mt_state rng;
    mt_seed(&rng, seed);
    for (int i = 0; i < n_doubles; i++) doubles[i] = mt_double(&rng);
    for (int i = 0; i < shuffle_n; i++) shuffled[i] = (double)i;
    mt_shuffle(&rng, shuffled, shuffle_n);
}

/* ------------------------------------------------------------------------- */
/* CPU baseline runner: n_threads independent run_all_strategies() calls      */
/* (distinct seeds), wall-clock seconds returned.  Paths are independent, so  */
/* this is the fair all-cores figure for the reference's algorithm.           */
/* ------------------------------------------------------------------------- */
typedef struct {
    const oracle_params *p;
    uint32_t seed;
    int rc;
    double sink;
} bench_job;

static void *bench_worker(void *arg)
{
    bench_job *j = (bench_job *)arg;
    double summary[ORACLE_NSTRAT * 9];
    j->rc = oracle_run_all(j->p, j->seed, summary, NULL, NULL, NULL);
    j->sink = summary[0];
    return NULL;
}

double oracle_bench(const oracle_params *p, uint32_t seed0, int n_threads, double *sink)
{
    if (n_threads < 1) n_threads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
    bench_job *jobs = (bench_job *)malloc(sizeof(bench_job) * n_threads);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int i = 0; i < n_threads; i++) {
        jobs[i].p = p;
        jobs[i].seed = seed0 + (uint32_t)i;
        jobs[i].rc = 0;
        pthread_create(&th[i], NULL, bench_worker, &jobs[i]);
    }
    double acc = 0;
    int rc = 0;
    for (int i = 0; i < n_threads; i++) {
        pthread_join(th[i], NULL);
        acc += jobs[i].sink;
        rc |= jobs[i].rc;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (sink) *sink = acc;
    free(th);
    free(jobs);
    if (rc) return -1.0;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
