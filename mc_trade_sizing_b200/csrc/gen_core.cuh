// gen_core.cuh -- counter-based trade-outcome generation (host/device, same code).
//
// Replaces NumPy's global MT19937 stream of the reference (RNG call sites
// src/trading_models.py:55,65,70,77,79,89,101,119,492) by Philox4x32-10 keyed with the user seed
// and countered by (simulation, shuffle, trade), so any path can be regenerated anywhere:
//
//   base multiset of simulation `sim` (generators :43-128): uniform #i is the 53-bit integer k_i = h_i * 2^37 + rest_i with
//       h_i = 16-bit half (i%8) of Philox(ctr = (i/8, 0, sim, 0)) and rest_i = the top 37 bits of y0:y1, y = Philox(ctr = (i, 1, sim, 0));
//       win <=> k_i < thr, thr = ceil(p * 2^53) -- the same law as `np.random.rand() < p` / choice(p=[p, 1-p]).  (The device
//       computes rest_i only when h_i ties with the top 16 bits of thr.)
//       Only the per-category COUNTS survive, because the reference shuffles every sequence before
//       using it (:491-492) and a shuffled sequence is a uniform permutation of its multiset.
//   path (sim, k) (the np.random.shuffle of :492): a uniform permutation of the simulation's multiset, generated one of two ways
//       by the number K of distinct values that can occur in the run:
//       K <= 6   sequential draws from the URN of remaining counts (gen_path_codes_urn): the n-th draw of 8-trade block b takes the
//                first value r with idx < C_r (C = cumulative remaining counts in value order), idx = mulhi32(x, remaining),
//                x = word (n%4) of Philox(ctr = (2b + n/4, k, sim, 1)), and fills trade 8b + 7 - n (a partial last block of m trades:
//                8b + m - 1 - n).  No array, no read-modify-write: K - 1 compares / predicated decrements per trade.
//       K  > 6   Fisher-Yates over the packed code column in shared memory (gen_path_codes_fy), exactly the loop of
//                RandomState.shuffle:  for i = L-1 .. 1:  j = mulhi32(x_i, i + 1);  swap(a_i, a_j),  x_i = word (i%4) of
//                Philox(ctr = (i/4, k, sim, 1)).
//       Both yield a uniform permutation of the multiset (sequential sampling without replacement, assigned to the positions in
//       any fixed order, is exchangeable); the multiply-shift index has bias <= remaining / 2^32 (< 2.4e-7 for 1000 trades).
#pragma once
#include <stdint.h>

#include "spec.hpp"

namespace mcts {

MCTS_HD uint32_t mulhi32(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

// Philox4x32-10 (Salmon et al., SC'11); constants as in Random123 / cuRAND.
MCTS_HD void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t (&out)[4])
{
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = mulhi32(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
        k0 += W0;
        k1 += W1;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

constexpr uint32_t kStreamBase = 0u, kStreamShuffle = 1u;

// Sequential reader of the 53-bit uniforms of one simulation's base stream.
//
// Is the i-th 53-bit uniform integer k_i below thr (win <=> u < p, thr = ceil(p * 2^53))?
//     k_i = h_i * 2^37 + rest_i
//     h_i    = 16-bit half (i % 8) of Philox(ctr = (i / 8, 0, sim, 0)): word (i % 8) / 2, its upper half when i is even, lower when odd
//     rest_i = the top 37 bits of the 64-bit number y0:y1, y = Philox(ctr = (i, 1, sim, 0))
// The 16 bits of h_i decide unless they tie with thr's top 16 bits (probability 2^-16), so the second block is generated only
// then: EIGHT Bernoulli draws per Philox block (round 1: two, round 2 first version: four), the same law -- k_i is uniform on
// [0, 2^53) whatever part of it is looked at first.
struct BaseStream {
    uint32_t k0, k1, sim;
    uint32_t i;
    uint32_t x[4];
    MCTS_HD void init(uint64_t seed, uint32_t sim_)
    {
        k0 = (uint32_t)seed;
        k1 = (uint32_t)(seed >> 32);
        sim = sim_;
        i = 0;
    }
    MCTS_HD void load_block() { philox4x32_10(k0, k1, i >> 3, 0u, sim, kStreamBase, x); }
    // the rare second look: draw `idx` tied with the top 16 bits of thr
    MCTS_HD bool tie_break(uint32_t idx, uint64_t thr) const
    {
        uint32_t y[4];
        philox4x32_10(k0, k1, idx, 1u, sim, kStreamBase, y);
        const uint64_t rest = (((uint64_t)y[0] << 32) | (uint64_t)y[1]) >> 27;
        return rest < (thr & ((1ull << 37) - 1ull));
    }
    // draw (i - i % 8) + J of the block in x, J a compile-time index; thr_hi = thr() >> 37 (thr() is evaluated on a tie only)
    template <int J, class THR>
    MCTS_HD bool lt_at(uint32_t thr_hi, THR &&thr) const
    {
        const uint32_t h = (J & 1) ? (x[J >> 1] & 0xffffu) : (x[J >> 1] >> 16);
        bool lt = h < thr_hi;
        if (h == thr_hi) lt = tie_break((i & ~7u) + (uint32_t)J, thr());
        return lt;
    }
    MCTS_HD bool next_lt(uint64_t thr)
    {
        if ((i & 7u) == 0u) load_block();
        const uint32_t w = (i & 4u) ? ((i & 2u) ? x[3] : x[2]) : ((i & 2u) ? x[1] : x[0]);
        const uint32_t h = (i & 1u) ? (w & 0xffffu) : (w >> 16);
        const uint32_t thr_hi = (uint32_t)(thr >> 37);
        bool lt = h < thr_hi;
        if (h == thr_hi) lt = tie_break(i, thr);
        i++;
        return lt;
    }
    // the next eight draws against one threshold (i % 8 == 0): how many are below it
    MCTS_HD uint32_t count_lt8(uint64_t thr)
    {
        load_block();
        const uint32_t thr_hi = (uint32_t)(thr >> 37);
        uint32_t wins = 0;
        wins += lt_at<0>(thr_hi, [&] { return thr; }) ? 1u : 0u;
        wins += lt_at<1>(thr_hi, [&] { return thr; }) ? 1u : 0u;
        wins += lt_at<2>(thr_hi, [&] { return thr; }) ? 1u : 0u;
        wins += lt_at<3>(thr_hi, [&] { return thr; }) ? 1u : 0u;
        wins += lt_at<4>(thr_hi, [&] { return thr; }) ? 1u : 0u;
        wins += lt_at<5>(thr_hi, [&] { return thr; }) ? 1u : 0u;
        wins += lt_at<6>(thr_hi, [&] { return thr; }) ? 1u : 0u;
        wins += lt_at<7>(thr_hi, [&] { return thr; }) ? 1u : 0u;
        i += 8;
        return wins;
    }
};

// Per-category counts of the base sequence of one simulation (what :477-490 would have produced).  Whole Philox blocks of
// eight draws take the unrolled forms (no word selection by index); the draws before the first block boundary and after
// the last one go through next_lt.
MCTS_HD void gen_base_counts(const GenSpec &g, uint64_t seed, uint32_t sim, uint32_t (&counts)[kMaxCat])
{
#pragma unroll
    for (int k = 0; k < kMaxCat; k++) counts[k] = 0;
    BaseStream rs;
    rs.init(seed, sim);
    if (g.kind == GEN_SEGMENTS) {
        for (int s = 0; s < g.n_seg; s++) {
            const Segment sg = g.seg[s];
            uint32_t wins = 0;
            int t = 0;
            for (; t < sg.len && (rs.i & 7u) != 0u; t++) wins += rs.next_lt(sg.thr) ? 1u : 0u;
            for (; t + 8 <= sg.len; t += 8) wins += rs.count_lt8(sg.thr);
            for (; t < sg.len; t++) wins += rs.next_lt(sg.thr) ? 1u : 0u;
#pragma unroll
            for (int k = 0; k < kMaxCat; k++) {
                counts[k] += (k == sg.cat_win ? wins : 0u) + (k == sg.cat_loss ? (uint32_t)sg.len - wins : 0u);
            }
        }
    } else {
        uint32_t wins = 0;
        if (g.kind == GEN_MARKOV1) {
            uint32_t last = rs.next_lt(g.thr_init) ? 1u : 0u;  // :70
            wins = last;
            const uint64_t t_l = g.thr1[0], t_w = g.thr1[1];
            const uint32_t h_l = (uint32_t)(t_l >> 37), h_w = (uint32_t)(t_w >> 37);
            auto step = [&](bool lt) {
                last = lt ? 1u : 0u;
                wins += last;
            };
            int t = 1;
            for (; t < g.L && (rs.i & 7u) != 0u; t++) step(rs.next_lt(last ? t_w : t_l));  // :74-84
            for (; t + 8 <= g.L; t += 8) {
                rs.load_block();
                step(rs.lt_at<0>(last ? h_w : h_l, [&] { return last ? t_w : t_l; }));
                step(rs.lt_at<1>(last ? h_w : h_l, [&] { return last ? t_w : t_l; }));
                step(rs.lt_at<2>(last ? h_w : h_l, [&] { return last ? t_w : t_l; }));
                step(rs.lt_at<3>(last ? h_w : h_l, [&] { return last ? t_w : t_l; }));
                step(rs.lt_at<4>(last ? h_w : h_l, [&] { return last ? t_w : t_l; }));
                step(rs.lt_at<5>(last ? h_w : h_l, [&] { return last ? t_w : t_l; }));
                step(rs.lt_at<6>(last ? h_w : h_l, [&] { return last ? t_w : t_l; }));
                step(rs.lt_at<7>(last ? h_w : h_l, [&] { return last ? t_w : t_l; }));
                rs.i += 8;
            }
            for (; t < g.L; t++) step(rs.next_lt(last ? t_w : t_l));
        } else {
            uint32_t older = rs.next_lt(g.thr_init) ? 1u : 0u;  // :89, two unconditional trades
            uint32_t newer = rs.next_lt(g.thr_init) ? 1u : 0u;
            wins = older + newer;
            const uint64_t t_ll = g.thr2[0], t_lw = g.thr2[1], t_wl = g.thr2[2], t_ww = g.thr2[3];
            const uint32_t h_ll = (uint32_t)(t_ll >> 37), h_lw = (uint32_t)(t_lw >> 37), h_wl = (uint32_t)(t_wl >> 37), h_ww = (uint32_t)(t_ww >> 37);
            auto thr_of = [&]() { return older ? (newer ? t_ww : t_wl) : (newer ? t_lw : t_ll); };
            auto hi_of = [&]() { return older ? (newer ? h_ww : h_wl) : (newer ? h_lw : h_ll); };
            auto step = [&](bool lt) {
                const uint32_t w = lt ? 1u : 0u;
                wins += w;
                older = newer;
                newer = w;
            };
            int t = 2;
            for (; t < g.L && (rs.i & 7u) != 0u; t++) step(rs.next_lt(thr_of()));  // :92-103
            for (; t + 8 <= g.L; t += 8) {
                rs.load_block();
                step(rs.lt_at<0>(hi_of(), thr_of));
                step(rs.lt_at<1>(hi_of(), thr_of));
                step(rs.lt_at<2>(hi_of(), thr_of));
                step(rs.lt_at<3>(hi_of(), thr_of));
                step(rs.lt_at<4>(hi_of(), thr_of));
                step(rs.lt_at<5>(hi_of(), thr_of));
                step(rs.lt_at<6>(hi_of(), thr_of));
                step(rs.lt_at<7>(hi_of(), thr_of));
                rs.i += 8;
            }
            for (; t < g.L; t++) step(rs.next_lt(thr_of()));
        }
#pragma unroll
        for (int k = 0; k < kMaxCat; k++) {
            counts[k] += (k == g.cat_win ? wins : 0u) + (k == g.cat_loss ? (uint32_t)g.L - wins : 0u);
        }
    }
}

// The urn for a handful of values (K <= kMaxUrnValues): no array, no read-modify-write -- every draw is one multiply-high,
// K - 1 compares against the cumulative remaining counts, K - 1 predicated decrements and a select chain for the code.  The n-th
// draw of an 8-trade block fills trade 7 - n of the block (n = 0 .. 7; a partial last block of m trades fills m - 1 - n), which
// lets the word be assembled as word * 16 + code; draw n uses word n % 4 of Philox(ctr = (2 * block + n / 4, k, sim, 1)).
// Measured on the B200 against Fisher-Yates for the five values of the bench workload: k_sim_paths 19.72 -> 19.22 ms (six values:
// 19.41 ms; from seven values on Fisher-Yates is cheaper).
constexpr int kMaxUrnValues = 6;

// The kernels are instantiated for urns of 2, 5 and 6 values; a run with 3 or 4 values uses the 5-value urn with empty categories
// appended (an empty category at the end has the cumulative count `remaining` like the last real one, and the smallest matching
// category wins).  0 = Fisher-Yates.
MCTS_HD constexpr int urn_width(int K) { return K <= 2 ? 2 : K <= 5 ? 5 : K <= kMaxUrnValues ? 6 : 0; }

template <int K>
MCTS_HD void gen_path_codes_urn(const GenSpec &g, uint64_t seed, uint32_t sim, uint32_t k, const uint32_t *counts, uint32_t *codes, int stride)
{
    static_assert(K >= 2 && K <= 8, "urn generator: 2 .. 8 values");
    uint32_t C[K - 1], code[K];
    uint32_t acc = 0;
#pragma unroll
    for (int j = 0; j < K; j++) {
        code[j] = (g.rank_codes >> (4 * j)) & 15u;
        acc += counts[j];
        if (j < K - 1) C[j] = acc;
    }
    uint32_t rem = acc;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    auto draw = [&](uint32_t x) -> uint32_t {
        const uint32_t idx = mulhi32(x, rem);
        uint32_t r = code[K - 1];
#pragma unroll
        for (int j = K - 2; j >= 0; j--) {  // the smallest category whose cumulative count exceeds idx
#if defined(__CUDA_ARCH__)
            // one compare, one PREDICATED in-place decrement, one select (written as C -= lt, ptxas emits an unconditional
            // decrement into a new register plus a predicated move back: four instructions per category instead of three)
            asm("{ .reg .pred p;\n\tsetp.lt.u32 p, %2, %0;\n\t@p add.u32 %0, %0, -1;\n\tselp.u32 %1, %3, %1, p; }"
                : "+r"(C[j]), "+r"(r) : "r"(idx), "r"(code[j]));
#else
            const bool lt = idx < C[j];
            r = lt ? code[j] : r;
            C[j] -= lt ? 1u : 0u;
#endif
        }
        rem -= 1u;
        return r;
    };
    const int n_oct = g.L >> 3, tail = g.L & 7;
    for (int oct = 0; oct < n_oct; oct++) {
        uint32_t xa[4], xb[4];
        philox4x32_10(k0, k1, 2u * oct, k, sim, kStreamShuffle, xa);
        philox4x32_10(k0, k1, 2u * oct + 1u, k, sim, kStreamShuffle, xb);
        uint32_t word = 0;
#pragma unroll
        for (int n = 0; n < 8; n++) word = word * 16u + draw(n < 4 ? xa[n] : xb[n - 4]);
        codes[(size_t)oct * stride] = word;
    }
    if (tail) {
        uint32_t xa[4], xb[4];
        philox4x32_10(k0, k1, 2u * n_oct, k, sim, kStreamShuffle, xa);
        philox4x32_10(k0, k1, 2u * n_oct + 1u, k, sim, kStreamShuffle, xb);
        uint32_t word = 0;
#pragma unroll
        for (int n = 0; n < 8; n++)
            if (n < tail) word = word * 16u + draw(n < 4 ? xa[n] : xb[n - 4]);
        codes[(size_t)n_oct * stride] = word;
    }
}

// Fisher-Yates over the packed codes (any number of categories).  The code column is first filled in
// base order (runs of each category, in category order), then shuffled from the top exactly like
// RandomState.shuffle (:492):  for i = L-1 .. 1:  j = mulhi32(x_i, i + 1);  swap(a_i, a_j),
// x_i = word (i%4) of Philox(ctr = (i/4, k, sim, 1)).  The word holding position i lives in a register
// while i sweeps through it; a_j is read-modify-written in (thread-private) shared memory.
MCTS_HD void gen_path_codes_fy(const GenSpec &g, uint64_t seed, uint32_t sim, uint32_t k, const uint32_t *counts, uint32_t *codes, int stride)
{
    const uint32_t L = (uint32_t)g.L;
    const int n_words = (int)((L + 7u) >> 3);
    const uint32_t rank_codes_lo = g.rank_codes, rank_codes_hi = g.rank_codes_hi;
    {
        int cat = 0;
        uint32_t end = counts[0];
        uint32_t code = rank_codes_lo & 15u;
        for (int w = 0; w < n_words; w++) {
            const uint32_t t0 = 8u * (uint32_t)w;
            uint32_t word;
            if (t0 + 8u <= end) {
                word = code * 0x11111111u;
            } else {
                word = 0;
                for (uint32_t i = 0; i < 8u; i++) {
                    const uint32_t t = t0 + i;
                    while (t >= end && cat < g.K - 1) {
                        cat++;
                        end += counts[cat];
                        code = ((cat < 8 ? rank_codes_lo >> (4 * cat) : rank_codes_hi >> (4 * (cat - 8)))) & 15u;
                    }
                    if (t < L) word |= code << (4u * i);
                }
            }
            codes[(size_t)w * stride] = word;
        }
    }
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    // one swap: position i = 8*w + P (word `cur` in a register), partner j anywhere below
    auto swap_step = [&](uint32_t &cur, uint32_t w, int P, uint32_t x) {
        const uint32_t i1 = 8u * w + (uint32_t)P + 1u;
        const uint32_t j = mulhi32(x, i1);
        const uint32_t wj = j >> 3, sh = (j << 2) & 28u;
        uint32_t *slot = codes + (size_t)wj * stride;
        const uint32_t tj = *slot;
        const bool alias = wj == w;  // then the live copy is `cur`, the stored one is stale
        const uint32_t srcw = alias ? cur : tj;
        const uint32_t d = ((srcw >> sh) ^ (cur >> (4 * P))) & 15u;  // a_i ^ a_j
        const uint32_t dsh = d << sh;
        cur = cur ^ (d << (4 * P)) ^ (alias ? dsh : 0u);
        *slot = tj ^ dsh;  // for an aliased word this is overwritten when `cur` is stored
    };
    int w = n_words - 1;
    if ((L & 7u) != 0u) {  // partial top word: positions >= L do not exist
        uint32_t cur = codes[(size_t)w * stride];
        uint32_t xa[4], xb[4];
        philox4x32_10(k0, k1, 2u * w, k, sim, kStreamShuffle, xa);
        philox4x32_10(k0, k1, 2u * w + 1u, k, sim, kStreamShuffle, xb);
        for (int p = (int)(L & 7u) - 1; p >= 0; p--) {
            uint32_t x = xa[0];
#pragma unroll
            for (int q = 1; q < 8; q++) x = (p == q) ? (q < 4 ? xa[q] : xb[q - 4]) : x;
            // dynamic nibble position: same swap with run-time shifts
            const uint32_t i1 = 8u * (uint32_t)w + (uint32_t)p + 1u;
            const uint32_t j = mulhi32(x, i1);
            const uint32_t wj = j >> 3, sh = (j << 2) & 28u, shp = 4u * (uint32_t)p;
            uint32_t *slot = codes + (size_t)wj * stride;
            const uint32_t tj = *slot;
            const bool alias = wj == (uint32_t)w;
            const uint32_t srcw = alias ? cur : tj;
            const uint32_t d = ((srcw >> sh) ^ (cur >> shp)) & 15u;
            const uint32_t dsh = d << sh;
            cur = cur ^ (d << shp) ^ (alias ? dsh : 0u);
            *slot = tj ^ dsh;
        }
        codes[(size_t)w * stride] = cur;
        w--;
    }
    for (; w >= 0; w--) {
        uint32_t cur = codes[(size_t)w * stride];
        uint32_t xa[4], xb[4];
        philox4x32_10(k0, k1, 2u * w, k, sim, kStreamShuffle, xa);
        philox4x32_10(k0, k1, 2u * w + 1u, k, sim, kStreamShuffle, xb);
        swap_step(cur, (uint32_t)w, 7, xb[3]);
        swap_step(cur, (uint32_t)w, 6, xb[2]);
        swap_step(cur, (uint32_t)w, 5, xb[1]);
        swap_step(cur, (uint32_t)w, 4, xb[0]);
        swap_step(cur, (uint32_t)w, 3, xa[3]);
        swap_step(cur, (uint32_t)w, 2, xa[2]);
        swap_step(cur, (uint32_t)w, 1, xa[1]);
        swap_step(cur, (uint32_t)w, 0, xa[0]);
        codes[(size_t)w * stride] = cur;
    }
}

// The same Fisher-Yates on plain fp64 values in global memory (runs with more distinct values than
// 4-bit codes can name): vals[t * stride], base order = runs of each value in category order.
MCTS_HD void gen_path_values_fy(const GenSpec &g, uint64_t seed, uint32_t sim, uint32_t k, const uint32_t *counts, double *vals, size_t stride)
{
    const uint32_t L = (uint32_t)g.L;
    uint32_t t = 0;
    for (int cat = 0; cat < g.K; cat++) {
        const double v = g.val[cat];
        for (uint32_t c = counts[cat]; c > 0 && t < L; c--, t++) vals[(size_t)t * stride] = v;
    }
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t x[4] = {0, 0, 0, 0};
    uint32_t have = 0xffffffffu;  // Philox block currently in x
    for (uint32_t i = L - 1; i >= 1; i--) {
        if ((i >> 2) != have) {
            have = i >> 2;
            philox4x32_10(k0, k1, have, k, sim, kStreamShuffle, x);
        }
        const uint32_t xi = (i & 3u) == 0u ? x[0] : (i & 3u) == 1u ? x[1] : (i & 3u) == 2u ? x[2] : x[3];
        const uint32_t j = mulhi32(xi, i + 1u);
        const double a = vals[(size_t)i * stride], b = vals[(size_t)j * stride];
        vals[(size_t)i * stride] = b;
        vals[(size_t)j * stride] = a;
    }
}

}  // namespace mcts
