// planner.hpp -- host-side lowering of run_all_strategies' arguments to what the kernels consume.
// Pure C++ (no CUDA): also compiled into the CPU unit-test harness.
#pragma once
#include <math.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/mcts.h"
#include "spec.hpp"

namespace mcts {

// ---- generator lowering ---------------------------------------------------------------------

// win <=> u < p for u = k / 2^53, k integer  <=>  k < ceil(p * 2^53)
inline uint64_t threshold53(double p)
{
    if (!(p > 0.0)) return 0;  // also NaN: `u < nan` is False
    if (p >= 1.0) return 1ull << 53;
    return (uint64_t)ceil(ldexp(p, 53));
}

// np.random.choice([a0, a1], p=[p0, 1-p0]) acceptance rule; false where NumPy raises ValueError.
inline bool choice_threshold(double p0, uint64_t *thr, std::string *err)
{
    const double p1 = 1 - p0;
    if (!(p0 >= 0.0) || !(p1 >= 0.0)) {
        *err = "probabilities are not non-negative (np.random.choice would raise ValueError)";
        return false;
    }
    if (fabs((p0 + p1) - 1.0) > 1.4901161193847656e-08) {
        *err = "probabilities do not sum to 1 (np.random.choice would raise ValueError)";
        return false;
    }
    double c1 = p0 + p1;
    double c0 = p0 / c1;  // cdf /= cdf[-1]
    *thr = threshold53(c0);
    return true;
}

inline uint64_t bits_of(double v)
{
    uint64_t b;
    memcpy(&b, &v, 8);
    return b;
}

struct SpecBuilder {
    GenSpec g;
    std::string err;
    int status = MCTS_OK;

    int category(double v)
    {
        for (int i = 0; i < g.K; i++)
            if (bits_of(g.val[i]) == bits_of(v)) return i;
        if (g.K >= kMaxCat) return -1;
        g.val[g.K] = v;
        return g.K++;
    }

    bool add_segment(int len, double p, double win, double loss)
    {
        if (len <= 0) return true;
        Segment s;
        memset(&s, 0, sizeof s);
        s.len = len;
        if (!choice_threshold(p, &s.thr, &err)) {
            status = MCTS_EINVAL;
            return false;
        }
        // an outcome the acceptance test can never produce (p >= 1: no loss; p <= 0: no win) gets no category, so that it does not
        // count towards K (which picks the generator and the urn width); its count is zero in every simulation anyway
        const uint64_t kOne53 = (uint64_t)1 << 53;
        s.cat_win = s.thr == 0 ? -1 : category(win);
        s.cat_loss = s.thr >= kOne53 ? -1 : category(-loss);
        if ((s.thr != 0 && s.cat_win < 0) || (s.thr < kOne53 && s.cat_loss < 0) || g.n_seg >= kMaxSeg) {
            status = MCTS_EUNSUPPORTED;
            err = "more than 32 distinct trade values in one run";
            return false;
        }
        g.seg[g.n_seg++] = s;
        g.L += len;
        return true;
    }
};

// Lower the generator arguments. Mirrors the control flow of :43-128 exactly (segment clipping,
// tail fill, Markov-2's two unconditional first trades).
inline int build_spec(const mcts_params &p, GenSpec *out, std::string *err)
{
    SpecBuilder b;
    memset(&b.g, 0, sizeof b.g);
    GenSpec &g = b.g;
    const int n = p.num_trades;
    if (p.mode == MCTS_MODE_DYNAMIC || p.mode == MCTS_MODE_REGIME) {
        g.kind = GEN_SEGMENTS;
        int left = n;
        if (p.mode == MCTS_MODE_DYNAMIC) {
            const int len[3] = {(int)(n * 0.2), (int)(n * 0.2), n - (int)(n * 0.4)};
            const double hr[3] = {fmin(p.hit_rate + 0.2, 1.0), fmax(p.hit_rate - 0.3, 0.05), p.hit_rate};
            const double aw[3] = {p.avg_win * 1.1, p.avg_win * 0.9, p.avg_win};
            const double al[3] = {p.avg_loss * 0.9, p.avg_loss * 1.1, p.avg_loss};
            for (int k = 0; k < 3; k++) {
                int l = len[k] < left ? len[k] : left;
                if (l <= 0) continue;
                if (!b.add_segment(l, hr[k], aw[k], al[k])) goto fail;
                left -= l;
                if (left <= 0) break;
            }
            if (left > 0 && !b.add_segment(left, p.hit_rate, p.avg_win, p.avg_loss)) goto fail;  // :64-65
        } else {
            mcts_regime def[3] = {{(int)(n * 0.3), 0.9, 200, 100}, {(int)(n * 0.2), 0.5, 100, 100}, {n - (int)(n * 0.5), 0.2, 100, 200}};
            const mcts_regime *rg = p.n_regimes < 0 ? def : p.regimes;
            const int nr = p.n_regimes < 0 ? 3 : p.n_regimes;
            if (nr > MCTS_MAX_REGIMES) {
                b.status = MCTS_EUNSUPPORTED;
                b.err = "more than 16 regimes";
                goto fail;
            }
            for (int k = 0; k < nr; k++) {
                int l = rg[k].length < left ? rg[k].length : left;
                if (l <= 0) continue;
                if (!b.add_segment(l, rg[k].hit_rate, rg[k].avg_win, rg[k].avg_loss)) goto fail;
                left -= l;
                if (left <= 0) break;
            }
        }
    } else {
        g.kind = p.mode == MCTS_MODE_MARKOV1 ? GEN_MARKOV1 : GEN_MARKOV2;
        g.L = g.kind == GEN_MARKOV1 ? (n > 1 ? n : 1) : (n > 2 ? n : 2);
        g.thr_init = threshold53(p.hit_rate);
        g.thr1[0] = threshold53(p.p_win_after_loss);
        g.thr1[1] = threshold53(p.p_win_after_win);
        g.thr2[0] = threshold53(p.p_win_ll);
        g.thr2[1] = threshold53(p.p_win_lw);
        g.thr2[2] = threshold53(p.p_win_wl);
        g.thr2[3] = threshold53(p.p_win_ww);
        g.cat_win = b.category(p.avg_win);
        g.cat_loss = b.category(-p.avg_loss);
    }
    if (g.L <= 0 || p.num_simulations <= 0 || p.num_mc_shuffles <= 0) {
        b.status = MCTS_EEMPTY;
        b.err = "zero paths or an empty trade sequence (np.min of an empty list raises in the reference)";
        goto fail;
    }
    {
        // 4-bit codes: bit 0 = (value > 0), bits 1..3 = index among the categories of that sign class
        int n_win = 0, n_loss = 0;
        double vmax = 0;
        g.all_int = 1;
        for (int k = 0; k < g.K; k++) {  // can every value get a 4-bit code (bit 0 = value > 0)?
            const int idx = g.val[k] > 0 ? n_win++ : n_loss++;
            if (idx >= 8 || k >= kMaxCodes) g.wide = 1;
        }
        n_win = n_loss = 0;
        for (int k = 0; k < g.K; k++) {
            const double v = g.val[k];
            const int win = v > 0;
            const int idx = win ? n_win++ : n_loss++;
            // >= 2^-969: every equity difference is then a multiple of 2^-1021, i.e. never a non-zero subnormal, which lets the
            // kernels test "new peak" on the high word alone (eval_core.cuh trade_step) and share the low word of v and 2v
            if (v != 0.0 && !(fabs(v) >= 2.004168360008973e-292 && fabs(v) <= 8.98846567431158e+307)) {
                *err = "trade values must be 0 or between 2^-969 and 2^1023 in magnitude in simulate mode (use replay for such inputs)";
                return MCTS_EUNSUPPORTED;
            }
            if (g.wide) {
                g.all_int = 0;
                continue;
            }
            const int code = (idx << 1) | win;
            if (k < 8) g.rank_codes |= (uint32_t)code << (4 * k);
            else g.rank_codes_hi |= (uint32_t)code << (4 * (k - 8));
            g.tab[code].v1 = v;
            g.tab[code].v2 = v * 2.0;
            g.tab[code].v3 = v * 3.0;
            if (!(fabs(v) <= 1e9) || v != floor(v) || (v == 0.0 && signbit(v))) g.all_int = 0;
            else g.ival[code] = (int32_t)v;
            if (fabs(v) > vmax) vmax = fabs(v);
        }
        if (g.all_int && 6.0 * vmax * (double)g.L >= 2147483648.0) g.all_int = 0;
        // equity and profit sums are bounded by 3 * max|v| * L: keep them finite, so that no NaN can arise (inf - inf in
        // the drawdown; the reference would report NaN there, the ordered comparisons of the kernels would drop it)
        if (vmax > 1.7976931348623157e308 / (3.0 * (double)g.L)) {
            *err = "trade values too large: 3 * max|value| * trades must stay below DBL_MAX in simulate mode";
            return MCTS_EUNSUPPORTED;
        }
    }
    *out = g;
    return MCTS_OK;
fail:
    *err = b.err;
    return b.status;
}

// ---- NumPy pairwise tree (numpy/_core/src/umath/loops_utils.h.src) ---------------

inline int64_t np_split(int64_t n)
{
    int64_t n2 = n / 2;
    return n2 - n2 % 8;
}

// Leaves of the tree over the L trades of one path, in order, with the number of internal
// nodes that complete right after each leaf (post-order merges).
inline void sched_visit(int64_t n, std::vector<int64_t> &leaf_len, std::vector<int> &merges, int depth, int *max_depth)
{
    if (depth > *max_depth) *max_depth = depth;
    if (n <= 128) {
        leaf_len.push_back(n);
        merges.push_back(0);
        return;
    }
    const int64_t n2 = np_split(n);
    sched_visit(n2, leaf_len, merges, depth + 1, max_depth);
    sched_visit(n - n2, leaf_len, merges, depth + 1, max_depth);
    merges.back() += 1;
}

inline int build_leaf_sched(int64_t L, LeafSched *s, std::string *err)
{
    std::vector<int64_t> len;
    std::vector<int> mg;
    int depth = 0;
    sched_visit(L, len, mg, 1, &depth);
    if ((int)len.size() > kMaxLeaves || depth > kMaxStackDepth) {
        *err = "num_trades too large for this build (pairwise tree over trades deeper than 12 levels)";
        return MCTS_EUNSUPPORTED;
    }
    memset(s, 0, sizeof *s);
    s->n_leaves = (int)len.size();
    s->n_oct = (int)(L / 8);
    s->tail = (int)(L % 8);
    s->max_depth = depth;
    int64_t pos = 0;
    for (size_t i = 0; i < len.size(); i++) {
        pos += len[i];
        // every leaf but the last is a whole number of 8-trade blocks (np_split returns multiples of 8)
        if (i + 1 < len.size() && (len[i] % 8 != 0 || len[i] < 8)) {
            *err = "internal: non-final pairwise leaf is not a multiple of 8";
            return MCTS_EUNSUPPORTED;
        }
        s->end_oct[i] = (int32_t)(pos / 8);
        s->merges[i] = mg[i];
    }
    return MCTS_OK;
}

// ---- chunk plan: nodes of the pairwise tree over PATHS (np.mean at :513-514) -----------------

inline int64_t clamp_chunk(int64_t chunk_paths)
{
    if (chunk_paths <= 0) chunk_paths = (int64_t)1 << 22;
    if (chunk_paths < 128) chunk_paths = 128;
    return chunk_paths;
}

inline void plan_visit(int64_t start, int64_t n, int64_t chunk, std::vector<int64_t> *first, std::vector<int64_t> *count)
{
    if (n <= chunk) {
        first->push_back(start);
        count->push_back(n);
        return;
    }
    const int64_t n2 = np_split(n);
    plan_visit(start, n2, chunk, first, count);
    plan_visit(start + n2, n - n2, chunk, first, count);
}

inline double combine_visit(int64_t n, int64_t chunk, const double *partials, int64_t *idx)
{
    if (n <= chunk) return partials[(*idx)++];
    const int64_t n2 = np_split(n);
    const double a = combine_visit(n2, chunk, partials, idx);
    const double b = combine_visit(n - n2, chunk, partials, idx);
    return a + b;
}

// Depth of the perfect-binary top of the tree used by the device reduction of one chunk:
// all nodes above depth d are internal (> 128 elements), nodes at depth d hold ~kReduceNode..2*kReduceNode-1.
#ifndef MCTS_REDUCE_NODE
#define MCTS_REDUCE_NODE 1024
#endif
constexpr int64_t kReduceNode = MCTS_REDUCE_NODE;
inline int reduce_depth(int64_t n)
{
    int d = 0;
    while ((n >> (d + 1)) >= kReduceNode && d < 24) d++;
    return d;
}

}  // namespace mcts
