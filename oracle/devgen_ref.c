/*
 * oracle/devgen_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Independent plain-C restatement of the DEVICE generator's published specification
 * (DESIGN.md section 3.3), written from the spec and not from the CUDA sources, so that the GPU's
 * on-device RNG path can be checked bit for bit on the CPU:
 *
 *   - Philox4x32-10 (Salmon et al. 2011; Random123 known answers are checked by the tests);
 *   - base multiset of simulation `sim`: the reference generators of src/trading_models.py:43-128
 *     (three-phase dynamic incl. clipping and tail fill :44-65, Markov-1 :68-85, Markov-2 :87-104,
 *     regime switch :106-128) driven by uniforms u_i = k_i / 2^53, k_i = h_i * 2^37 + rest_i with
 *     h_i = 16-bit half (i%8) of Philox(key = seed, ctr = (i/8, 0, sim, 0)) (word (i%8)/2; upper half for even i, lower for odd i) and
 *     rest_i = top 37 bits of y0:y1, y = Philox(ctr = (i, 1, sim, 0)); the acceptance tests are the reference's own
 *     (`rand() < p`, choice(): `u < cdf[0]` after cdf normalisation) in floating point;
 *   - shuffle of path (sim, k) (np.random.shuffle of :492).  The run's VALUE LIST is the distinct values in order of first
 *     appearance in the generator's (win, -loss) list, leaving out outcomes the acceptance test can never produce (a segment
 *     whose normalised win probability is >= 1 has no loss value, one with <= 0 no win value).  With K values in the list:
 *       K <= 6: sequential draws from the URN of remaining counts.  Trades are filled in blocks of 8: the n-th draw (n = 0..7)
 *               of block b goes to trade 8b + 7 - n (a partial last block of m trades: 8b + m - 1 - n); it takes the first
 *               value r of the list with idx < C_r, where C_r = remaining count of values 0..r and
 *               idx = floor(x * remaining / 2^32), x = word (n%4) of Philox(ctr = (2b + n/4, k, sim, 1));
 *       K  > 6: Fisher-Yates from the top: the array starts as the values in list order, each repeated by its count;
 *               for i = L-1 .. 1:  j = floor(x_i * (i+1) / 2^32), swap(a_i, a_j),  x_i = word (i%4) of Philox(ctr = (i/4, k, sim, 1)).
 *
 * It shares the parameter struct of mcts_oracle.c and produces the post-shuffle sequences as doubles;
 * feeding them to oracle_eval_paths() gives the expected per-path results of mcts_simulate().
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define DG_MAX_REGIMES 16

typedef struct {
    int length;
    double hit_rate, avg_win, avg_loss;
} dg_regime;

typedef struct { /* identical layout to oracle_params */
    int mode;
    int num_trades;
    int num_simulations;
    int num_mc_shuffles;
    double hit_rate, avg_win, avg_loss;
    double p_win_after_win, p_win_after_loss;
    double p_win_ww, p_win_wl, p_win_lw, p_win_ll;
    int n_regimes;
    dg_regime regimes[DG_MAX_REGIMES];
} dg_params;

static void philox(uint32_t k0, uint32_t k1, const uint32_t ctr[4], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    for (int round = 0; round < 10; round++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void devgen_philox(uint32_t k0, uint32_t k1, const uint32_t *ctr, uint32_t *out) { philox(k0, k1, ctr, out); }

typedef struct {
    uint64_t seed;
    uint32_t sim;
    uint32_t i;
    uint32_t buf[4];
} base_stream;

static double next_uniform(base_stream *s)
{
    /* u_i = (h * 2^37 + rest) / 2^53:
     *   h    = 16-bit half (i%8) of Philox(ctr = (i/8, 0, sim, 0)) -- word (i%8)/2, upper half for even i, lower half for odd i;
     *   rest = top 37 bits of the 64-bit number y0:y1, y = Philox(ctr = (i, 1, sim, 0)) */
    uint32_t ctr_a[4] = {s->i >> 3, 0u, s->sim, 0u}, ctr_b[4] = {s->i, 1u, s->sim, 0u}, xa[4], xb[4];
    philox((uint32_t)s->seed, (uint32_t)(s->seed >> 32), ctr_a, xa);
    philox((uint32_t)s->seed, (uint32_t)(s->seed >> 32), ctr_b, xb);
    uint32_t w = xa[(s->i & 7u) >> 1];
    uint64_t h = (s->i & 1u) ? (w & 0xffffu) : (w >> 16);
    uint64_t rest = (((uint64_t)xb[0] << 32) | (uint64_t)xb[1]) >> 27;
    s->i++;
    return (double)((h << 37) | rest) / 9007199254740992.0;
}

/* choice([a0, a1], p=[p0, 1-p0]) on the base stream; returns -1 where NumPy raises */
static int choice2(base_stream *s, double a0, double a1, double p0, int l, double *out)
{
    double p1 = 1 - p0;
    if (!(p0 >= 0.0) || !(p1 >= 0.0)) return -1;
    if (fabs((p0 + p1) - 1.0) > 1.4901161193847656e-08) return -1;
    double c0 = p0 / (p0 + p1);
    for (int i = 0; i < l; i++) out[i] = next_uniform(s) < c0 ? a0 : a1;
    return 0;
}

/* the base sequence of one simulation, in generation order; returns its length or -1 */
static int base_sequence(const dg_params *p, uint64_t seed, uint32_t sim, double *out)
{
    base_stream s = {seed, sim, 0, {0, 0, 0, 0}};
    int n = p->num_trades, pos = 0;
    if (p->mode == 0) {
        int len[3] = {(int)(n * 0.2), (int)(n * 0.2), n - (int)(n * 0.4)};
        double hr[3] = {fmin(p->hit_rate + 0.2, 1.0), fmax(p->hit_rate - 0.3, 0.05), p->hit_rate};
        double aw[3] = {p->avg_win * 1.1, p->avg_win * 0.9, p->avg_win};
        double al[3] = {p->avg_loss * 0.9, p->avg_loss * 1.1, p->avg_loss};
        int left = n;
        for (int k = 0; k < 3; k++) {
            int l = len[k] < left ? len[k] : left;
            if (l <= 0) continue;
            if (choice2(&s, aw[k], -al[k], hr[k], l, out + pos)) return -1;
            pos += l;
            left -= l;
            if (left <= 0) break;
        }
        if (left > 0) {
            if (choice2(&s, p->avg_win, -p->avg_loss, p->hit_rate, left, out + pos)) return -1;
            pos += left;
        }
    } else if (p->mode == 3) {
        dg_regime def[3] = {{(int)(n * 0.3), 0.9, 200, 100}, {(int)(n * 0.2), 0.5, 100, 100}, {n - (int)(n * 0.5), 0.2, 100, 200}};
        const dg_regime *rg = p->n_regimes < 0 ? def : p->regimes;
        int nr = p->n_regimes < 0 ? 3 : p->n_regimes, left = n;
        for (int k = 0; k < nr; k++) {
            int l = rg[k].length < left ? rg[k].length : left;
            if (l <= 0) continue;
            if (choice2(&s, rg[k].avg_win, -rg[k].avg_loss, rg[k].hit_rate, l, out + pos)) return -1;
            pos += l;
            left -= l;
            if (left <= 0) break;
        }
    } else if (p->mode == 1) {
        int last = next_uniform(&s) < p->hit_rate;
        out[pos++] = last ? p->avg_win : -p->avg_loss;
        for (int t = 1; t < n; t++) {
            last = next_uniform(&s) < (last ? p->p_win_after_win : p->p_win_after_loss);
            out[pos++] = last ? p->avg_win : -p->avg_loss;
        }
    } else {
        int older = next_uniform(&s) < p->hit_rate;
        int newer = next_uniform(&s) < p->hit_rate;
        out[pos++] = older ? p->avg_win : -p->avg_loss;
        out[pos++] = newer ? p->avg_win : -p->avg_loss;
        for (int t = 2; t < n; t++) {
            double pw = older ? (newer ? p->p_win_ww : p->p_win_wl) : (newer ? p->p_win_lw : p->p_win_ll);
            int w = next_uniform(&s) < pw;
            out[pos++] = w ? p->avg_win : -p->avg_loss;
            older = newer;
            newer = w;
        }
    }
    return pos;
}

/* normalised win probability of choice([w, -l], p=[p0, 1-p0]) as choice2() uses it */
static double norm_p(double p0) { return p0 / (p0 + (1 - p0)); }

/* the distinct values of a run in order of first appearance in the generator's (win, -loss) list; outcomes that the acceptance
 * test `u < c0`, u in [0, 1), can never produce are left out (c0 >= 1: never a loss; c0 <= 0: never a win) */
static int value_list(const dg_params *p, double *vals)
{
    int n = p->num_trades, nv = 0;
    double cand[2 * DG_MAX_REGIMES + 8];
    int nc = 0;
#define ADD_SEGMENT(w, l, p0)                   \
    do {                                        \
        double c0__ = norm_p(p0);               \
        if (c0__ > 0.0) cand[nc++] = (w);       \
        if (!(c0__ >= 1.0)) cand[nc++] = -(l);  \
    } while (0)
    if (p->mode == 0) {
        int len[3] = {(int)(n * 0.2), (int)(n * 0.2), n - (int)(n * 0.4)};
        double hr[3] = {fmin(p->hit_rate + 0.2, 1.0), fmax(p->hit_rate - 0.3, 0.05), p->hit_rate};
        double aw[3] = {p->avg_win * 1.1, p->avg_win * 0.9, p->avg_win};
        double al[3] = {p->avg_loss * 0.9, p->avg_loss * 1.1, p->avg_loss};
        int left = n;
        for (int k = 0; k < 3; k++) {
            int l = len[k] < left ? len[k] : left;
            if (l <= 0) continue;
            ADD_SEGMENT(aw[k], al[k], hr[k]);
            left -= l;
            if (left <= 0) break;
        }
        if (left > 0) ADD_SEGMENT(p->avg_win, p->avg_loss, p->hit_rate);
    } else if (p->mode == 3) {
        dg_regime def[3] = {{(int)(n * 0.3), 0.9, 200, 100}, {(int)(n * 0.2), 0.5, 100, 100}, {n - (int)(n * 0.5), 0.2, 100, 200}};
        const dg_regime *rg = p->n_regimes < 0 ? def : p->regimes;
        int nr = p->n_regimes < 0 ? 3 : p->n_regimes, left = n;
        for (int k = 0; k < nr; k++) {
            int l = rg[k].length < left ? rg[k].length : left;
            if (l <= 0) continue;
            ADD_SEGMENT(rg[k].avg_win, rg[k].avg_loss, rg[k].hit_rate);
            left -= l;
            if (left <= 0) break;
        }
    } else {
        cand[nc++] = p->avg_win;
        cand[nc++] = -p->avg_loss;
    }
#undef ADD_SEGMENT
    for (int i = 0; i < nc; i++) {
        int seen = 0;
        for (int j = 0; j < nv; j++) seen |= memcmp(&vals[j], &cand[i], 8) == 0;
        if (!seen) vals[nv++] = cand[i];
    }
    return nv;
}

/*
 * Post-shuffle sequences of paths [first_path, first_path + n) as doubles [n][L].
 * Returns L (>= 0) or -1 for invalid probabilities.
 */
int64_t devgen_sequences(const dg_params *p, uint64_t seed, int64_t first_path, int64_t n, double *out)
{
    int cap = (p->num_trades > 0 ? p->num_trades : 0) + 2;
    double *base = (double *)malloc(sizeof(double) * cap);
    double *arr = (double *)malloc(sizeof(double) * cap);
    double vals[2 * DG_MAX_REGIMES + 8];
    int nv = value_list(p, vals);
    int64_t L = -1;
    for (int64_t q = 0; q < n; q++) {
        int64_t path = first_path + q;
        uint32_t sim = (uint32_t)(path / p->num_mc_shuffles), k = (uint32_t)(path % p->num_mc_shuffles);
        int len = base_sequence(p, seed, sim, base);
        if (len < 0) {
            L = -1;
            break;
        }
        L = len;
        /* counts per distinct value, array in value order */
        int pos = 0, first_count = 0;
        for (int v = 0; v < nv; v++) {
            int c = 0;
            for (int t = 0; t < len; t++) c += memcmp(&base[t], &vals[v], 8) == 0;
            if (v == 0) first_count = c;
            for (int t = 0; t < c; t++) arr[pos++] = vals[v];
        }
        uint32_t x[4];
        (void)first_count;
        if (nv <= 6) { /* urn over the value list */
            uint32_t cnt[6] = {0, 0, 0, 0, 0, 0}, rem = (uint32_t)len;
            for (int v = 0; v < nv; v++)
                for (int t = 0; t < len; t++) cnt[v] += memcmp(&base[t], &vals[v], 8) == 0;
            for (int b = 0; 8 * b < len; b++) {
                int m = len - 8 * b < 8 ? len - 8 * b : 8;
                for (int nd = 0; nd < m; nd++) {
                    if ((nd & 3) == 0) {
                        uint32_t ctr[4] = {2u * (uint32_t)b + (uint32_t)(nd >> 2), k, sim, 1u};
                        philox((uint32_t)seed, (uint32_t)(seed >> 32), ctr, x);
                    }
                    uint32_t idx = (uint32_t)(((uint64_t)x[nd & 3] * rem) >> 32), cum = 0;
                    int r = nv - 1;
                    for (int v = 0; v < nv; v++) {
                        cum += cnt[v];
                        if (idx < cum) {
                            r = v;
                            break;
                        }
                    }
                    cnt[r]--;
                    rem--;
                    out[q * len + 8 * b + m - 1 - nd] = vals[r];
                }
            }
        } else { /* Fisher-Yates from the top */
            for (int i = len - 1; i >= 1; i--) {
                uint32_t ctr[4] = {(uint32_t)i >> 2, k, sim, 1u};
                philox((uint32_t)seed, (uint32_t)(seed >> 32), ctr, x);
                uint32_t j = (uint32_t)(((uint64_t)x[i & 3] * (uint32_t)(i + 1)) >> 32);
                double t = arr[i];
                arr[i] = arr[j];
                arr[j] = t;
            }
            memcpy(out + q * len, arr, sizeof(double) * len);
        }
    }
    free(base);
    free(arr);
    return L;
}
