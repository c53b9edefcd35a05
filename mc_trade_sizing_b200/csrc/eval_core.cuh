// eval_core.cuh -- the per-path evaluator: position-sizing rules, equity / peak / drawdown and the
// NumPy-ordered profit sum, written once for device code and (for CPU unit tests of the very same
// code, never as a fallback) for the host.
//
// Reference semantics (src/trading_models.py):
//   strategy_dynamic  :141-168   contribution c_t = value_t * size (0 while paused), host streak
//                                update, closure returns the NEXT size
//   make_condition_func :170-404 the 20 closures.  Because the host loop and many closures both
//                                bump the same streak counters, every one of them reduces to a
//                                function of the previous TWO outcomes only (verified
//                                against the literal closures by tests/test_eval_core_cpu.py and the
//                                GPU replay parity tests).
//   np.cumsum :167 / :136        equity: strictly sequential fp64 adds
//   np.sum    :168 / :137        profit: pairwise tree, 8 interleaved accumulators per <=128 leaf
//   calculate_drawdown :130-133  min_t (E_t - max_{j<=t} E_j), peak starts at E_0
//
// No multiply-add contraction can occur here: the categorical path only adds (products v*1, v*2,
// v*3 are tabulated on the host), the replay path uses __dmul_rn.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <type_traits>
#include <utility>

#include "spec.hpp"

namespace mcts {

// ---- the 11 distinct behaviours of the 20 strategy ids -------------------------
enum Rule : int {
    R_C1 = 0,  // ids 1, 3          always size 1
    R_AM,      // ids 2, 13         2 after a win
    R_F1,      // ids 4, 5, 14, 17  2 after the first win of a run only
    R_MG,      // ids 6, 7          2 after a loss
    R_MG2,     // id 8              2 after >= 2 consecutive losses
    R_P1,      // ids 9, 10, 15     pause after a win until a loss has passed
    R_P2,      // ids 11, 12        pause after 2 consecutive wins until a loss has passed
    R_S16,     // id 16             win -> pause; loss while trading -> 2; loss ending a pause -> 1
    R_S18,     // id 18             3 after >= 2 consecutive wins
    R_S19,     // id 19             2 after >= 2 wins, 3 after >= 2 losses
    R_S20      // id 20             3 after >= 2 consecutive losses
};

MCTS_HD constexpr int rule_of_id(int id)
{
    constexpr int t[21] = {-1, R_C1, R_AM, R_C1, R_F1, R_F1, R_MG, R_MG, R_MG2, R_P1, R_P1,
                           R_P2, R_P2, R_AM, R_F1, R_P1, R_S16, R_F1, R_S18, R_S19, R_S20};
    return t[id];
}

template <int... IDS>
struct IdList {
    static constexpr int N = sizeof...(IDS);
    MCTS_HD static constexpr int id(int i)
    {
        constexpr int a[N] = {IDS...};
        return a[i];
    }
};

// all 20 state machines, five per pass.  Ids of one behaviour class SHARE a pass: their
// contribution -- the select among {0, v, 2v, fl(3v)} and its predicate test -- is computed once per trade
// and feeds every id of the class, while each id keeps its own equity / peak / drawdown / profit chain
// (20 chains in all).  The chains of one class would be provably identical, so the device code starts every
// chain from a value read from constant memory (c_chain_start, -0.0 at run time) that the compiler cannot see
// through: nothing is merged, all 20 state machines are carried to the end.  Which classes share a pass and
// the order the ids are written in decides how well ptxas allocates the register pairs of the 8-trade loop
// body: the lists below come from a search (tools/pass_search) scored by the SASS instruction count of the loop
// bodies, followed by timing the ~140 shortest candidates on the B200 (tools/autotune): the fastest measured
// (19.17 ms per 2.1 M paths; SASS 1878 instructions per 8 trades x 20 ids, the shortest found is 1853 at 19.21 ms).
#ifndef MCTS_FP0
#define MCTS_FP0 4, 5, 14, 20, 17
#define MCTS_FP1 9, 10, 7, 6, 15
#define MCTS_FP2 13, 8, 16, 19, 2
#define MCTS_FP3 3, 12, 1, 11, 18
#endif
// the replay kernel (caller-supplied values, kernels.cuh k_replay_paths) evaluates the same ids with a different value source;
// its loop bodies are allocated differently by ptxas, so it has pass lists of its own from the same search
#ifndef MCTS_RP0
#define MCTS_RP0 MCTS_FP0
#define MCTS_RP1 MCTS_FP1
#define MCTS_RP2 MCTS_FP2
#define MCTS_RP3 MCTS_FP3
#endif
// curve export (HBM-bound, its sink keeps one more double per id): the lists that compile without spills there
#ifndef MCTS_CP0
#define MCTS_CP0 9, 10, 6, 15, 7
#define MCTS_CP1 14, 4, 5, 8, 17
#define MCTS_CP2 3, 18, 1, 12, 11
#define MCTS_CP3 2, 16, 19, 20, 13
#endif
using CurvePass0 = IdList<MCTS_CP0>;
using CurvePass1 = IdList<MCTS_CP1>;
using CurvePass2 = IdList<MCTS_CP2>;
using CurvePass3 = IdList<MCTS_CP3>;
using ReplayPass0 = IdList<MCTS_RP0>;
using ReplayPass1 = IdList<MCTS_RP1>;
using ReplayPass2 = IdList<MCTS_RP2>;
using ReplayPass3 = IdList<MCTS_RP3>;
using FullPass0 = IdList<MCTS_FP0>;
using FullPass1 = IdList<MCTS_FP1>;
using FullPass2 = IdList<MCTS_FP2>;
using FullPass3 = IdList<MCTS_FP3>;
// one representative per behaviour class (results fanned out by the caller); composition and order from
// the same search (re-run in round 2 for the structure-of-arrays value table: SASS 1246 -> 1135 instructions per 8 trades)
#ifndef MCTS_DP0
#define MCTS_DP0 8, 16, 9
#define MCTS_DP1 6, 1, 2
#define MCTS_DP2 4, 18, 20, 11, 19
#endif
using DedupPass0 = IdList<MCTS_DP0>;
using DedupPass1 = IdList<MCTS_DP1>;
using DedupPass2 = IdList<MCTS_DP2>;

// the pass lists are macros (searched, see above): check them where they are defined
template <class P>
constexpr int count_id(int id)
{
    int n = 0;
    for (int i = 0; i < P::N; i++) n += P::id(i) == id ? 1 : 0;
    return n;
}
template <class P>
constexpr int count_rule(int rule)
{
    int n = 0;
    for (int i = 0; i < P::N; i++) n += rule_of_id(P::id(i)) == rule ? 1 : 0;
    return n;
}
template <class P>
constexpr bool rules_distinct()
{
    for (int r = R_C1; r <= R_S20; r++)
        if (count_rule<P>(r) > 1) return false;
    return true;
}
constexpr bool full_passes_cover_all_ids()
{
    for (int id = 1; id <= 20; id++) {
        if (count_id<FullPass0>(id) + count_id<FullPass1>(id) + count_id<FullPass2>(id) + count_id<FullPass3>(id) != 1) return false;
        if (count_id<ReplayPass0>(id) + count_id<ReplayPass1>(id) + count_id<ReplayPass2>(id) + count_id<ReplayPass3>(id) != 1) return false;
        if (count_id<CurvePass0>(id) + count_id<CurvePass1>(id) + count_id<CurvePass2>(id) + count_id<CurvePass3>(id) != 1) return false;
    }
    return true;
}
constexpr bool dedup_passes_cover_all_rules()
{
    for (int r = R_C1; r <= R_S20; r++)
        if (count_rule<DedupPass0>(r) + count_rule<DedupPass1>(r) + count_rule<DedupPass2>(r) != 1) return false;
    return true;
}
static_assert(full_passes_cover_all_ids(), "every strategy id must appear in exactly one full pass");
static_assert(rules_distinct<DedupPass0>() && rules_distinct<DedupPass1>() && rules_distinct<DedupPass2>(), "dedup passes hold one id per behaviour class");
static_assert(dedup_passes_cover_all_rules(), "every behaviour class must appear in exactly one dedup pass");

template <int N, class F, int... Is>
MCTS_HD void static_for_impl(F &&f, std::integer_sequence<int, Is...>)
{
    (f(std::integral_constant<int, Is>{}), ...);
}
template <int N, class F>
MCTS_HD void static_for(F &&f)
{
    static_for_impl<N>(static_cast<F &&>(f), std::make_integer_sequence<int, N>{});
}

// Outcome-history masks of one 8-trade block.  Bit 4*i of a mask is the predicate for trade i:
//   pw  previous trade won          pl  previous trade lost
//   pww previous two both won       pll previous two both lost
//   fw  previous won, the one before did not (first win of a run; also true at the 2nd trade)
//   s16 previous lost and the one before did not win
// "lost" means "existed and value <= 0": before the first trade nothing won and nothing lost.
struct Masks {
    uint32_t pw, pl, pww, pll, fw, s16;
};

struct History {
    uint32_t w, l;  // win / loss bits of the previous block (nibble LSBs)
};

// Mask bits sit at position 4*i + 1 (never bit 0: a test of bit 0 costs ptxas an extra compare).
MCTS_HD Masks make_masks(uint32_t W, uint32_t Lb, History &h)
{
    Masks m;
    const uint32_t pw = (W << 5) | (h.w >> 27), ppw = (W << 9) | (h.w >> 23);
    const uint32_t pl = (Lb << 5) | (h.l >> 27), ppl = (Lb << 9) | (h.l >> 23);
    m.pw = pw;
    m.pl = pl;
    m.pww = pw & ppw;
    m.pll = pl & ppl;
    m.fw = pw & ~ppw;
    m.s16 = pl & ~ppw;
    h.w = W;
    h.l = Lb;
    return m;
}
MCTS_HD constexpr uint32_t mask_bit(int i) { return 2u << (4 * i); }

// From the second block on every trade has two real predecessors, so "lost" is simply "did not win":
// one history word, pl = ~pw, s16 = pll, and the pl-rules reuse the pw predicate negated.
MCTS_HD Masks make_masks_binary(uint32_t W, History &h)
{
    Masks m;
    const uint32_t pw = (W << 5) | (h.w >> 27), ppw = (W << 9) | (h.w >> 23);
    m.pw = pw;
    m.pl = ~pw;
    m.pww = pw & ppw;
    m.pll = ~(pw | ppw);
    m.fw = pw & ~ppw;
    m.s16 = m.pll;
    h.w = W;
    h.l = W ^ 0x11111111u;
    return m;
}

// the predicates of one trade (unused ones are dead code)
struct Preds {
    bool pw, pl, pww, pll, fw, s16;
};
template <bool BINARY>
MCTS_HD Preds trade_preds(const Masks &m, uint32_t bit)
{
    Preds p;
    p.pw = (m.pw & bit) != 0;
    p.pww = (m.pww & bit) != 0;
    p.pll = (m.pll & bit) != 0;
    p.fw = (m.fw & bit) != 0;
    p.pl = BINARY ? !p.pw : (m.pl & bit) != 0;
    p.s16 = BINARY ? p.pll : (m.s16 & bit) != 0;
    return p;
}

// p ? v2 : v1 where v2 == 2*v1.  SHARED_LO: the caller guarantees that v1 is zero or a normal number
// whose double is finite (checked on the host for the categorical path), so v2 and v1 have the same
// low word and only the high word needs a select -- one ALU instruction instead of two.
template <bool SHARED_LO>
MCTS_HD double select_double_or_single(bool p, double v2, double v1)
{
#if defined(__CUDA_ARCH__)
    if constexpr (SHARED_LO) return __hiloint2double(p ? __double2hiint(v2) : __double2hiint(v1), __double2loint(v1));
#endif
    return p ? v2 : v1;
}

// contribution of the current trade for one behaviour: a select among {0, v, 2v, fl(3v)}
template <int RULE, bool SHARED_LO>
MCTS_HD double contribution(const Preds &p, double v1, double v2, double v3)
{
    if constexpr (RULE == R_C1) return v1;
    if constexpr (RULE == R_AM) return select_double_or_single<SHARED_LO>(p.pw, v2, v1);
    if constexpr (RULE == R_F1) return select_double_or_single<SHARED_LO>(p.fw, v2, v1);
    if constexpr (RULE == R_MG) return select_double_or_single<SHARED_LO>(p.pl, v2, v1);
    if constexpr (RULE == R_MG2) return select_double_or_single<SHARED_LO>(p.pll, v2, v1);
    if constexpr (RULE == R_P1) return p.pw ? 0.0 : v1;
    if constexpr (RULE == R_P2) return p.pww ? 0.0 : v1;
    if constexpr (RULE == R_S16) return p.pw ? 0.0 : select_double_or_single<SHARED_LO>(p.s16, v2, v1);
    if constexpr (RULE == R_S18) return p.pww ? v3 : v1;
    if constexpr (RULE == R_S19) return p.pll ? v3 : select_double_or_single<SHARED_LO>(p.pww, v2, v1);
    if constexpr (RULE == R_S20) return p.pll ? v3 : v1;
    return v1;
}

MCTS_HD double neg_zero() { return -0.0; }
MCTS_HD double neg_inf()
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(0xfff0000000000000ll);
#else
    return -__builtin_inf();
#endif
}

// Running state of NS strategies on one path.
template <int NS>
struct Acc {
    double E[NS];     // equity (np.cumsum, sequential)
    double peak[NS];  // np.maximum.accumulate
    double dd[NS];    // np.min(equity - peak)
    double r[NS][8];  // the 8 interleaved accumulators of the current pairwise leaf
    History h;
};

#if defined(__CUDACC__)
// Start value (-0.0) of every id's chains, read from constant memory so that the compiler cannot prove two
// chains of one behaviour class equal and merge them (cf. c_int_zero of the int32 evaluator).
static __constant__ double c_chain_start[kNumStrategies + 1] = {-0.0, -0.0, -0.0, -0.0, -0.0, -0.0, -0.0, -0.0, -0.0, -0.0, -0.0,
                                                                -0.0, -0.0, -0.0, -0.0, -0.0, -0.0, -0.0, -0.0, -0.0, -0.0};
#endif

// -0.0 for chain `id`: a literal, or (OPAQUE, device only) the constant-memory copy
template <bool OPAQUE>
MCTS_HD double chain_start(int id)
{
#if defined(__CUDA_ARCH__)
    if constexpr (OPAQUE) return c_chain_start[id];
#endif
    (void)id;
    return neg_zero();
}

template <class IDS, bool OPAQUE>
MCTS_HD void acc_reset_leaf(Acc<IDS::N> &a)
{
#pragma unroll
    for (int s = 0; s < IDS::N; s++) {
        const double z = chain_start<OPAQUE>(IDS::id(s));  // leaf starts with r[j] = a[j]
#pragma unroll
        for (int j = 0; j < 8; j++) a.r[s][j] = z;
    }
}

template <class IDS, bool OPAQUE>
MCTS_HD void acc_init(Acc<IDS::N> &a)
{
#pragma unroll
    for (int s = 0; s < IDS::N; s++) {
        a.E[s] = chain_start<OPAQUE>(IDS::id(s));  // -0.0 + c == c bit for bit, so E_0 = c_0 needs no special case
        a.peak[s] = neg_inf();  // first trade: u = +inf > 0 -> peak_0 = E_0, d_0 = +0.0
        a.dd[s] = 0.0;
    }
    acc_reset_leaf<IDS, OPAQUE>(a);
    a.h.w = 0;
    a.h.l = 0;
}

// one trade of one strategy:  E += c;  r_j += c;  u = E - peak_old;  peak = max;  dd = min(dd, u)
// (u <= 0 is exactly the reference's E_t - peak_t; when u > 0 the reference records +0.0, which
// never lowers dd because dd <= 0 from the first trade on)
#if defined(__CUDACC__)
// 64-bit conditional moves as predicated IMADs (FMA pipe) instead of FSEL pairs (ALU pipe, the
// saturated port).  The multiplier 1 and addend 0 come from constant memory so that ptxas cannot
// fold the multiply-add into a move (which it would turn back into selects).
static __constant__ uint32_t c_sel_one = 1u, c_sel_zero = 0u;
template <bool INT_SIGN>
__device__ __forceinline__ void psel_gt0(double &dst, double u, double src)
{
    if constexpr (INT_SIGN) {  // u > 0 as a signed test of the high word (see trade_step): ALU pipe instead of FP64
        asm("{ .reg .pred p; .reg .b32 lo, hi, slo, shi, ulo, uhi;\n\t"
            "mov.b64 {ulo,uhi}, %1;\n\t"
            "setp.gt.s32 p, uhi, 0;\n\t"
            "mov.b64 {lo,hi}, %0;\n\tmov.b64 {slo,shi}, %2;\n\t"
            "@p mad.lo.u32 lo, slo, %3, %4;\n\t@p mad.lo.u32 hi, shi, %3, %4;\n\t"
            "mov.b64 %0, {lo,hi}; }"
            : "+d"(dst) : "d"(u), "d"(src), "r"(c_sel_one), "r"(c_sel_zero));
    } else {
        asm("{ .reg .pred p; .reg .b32 lo, hi, slo, shi;\n\t"
            "setp.gt.f64 p, %1, 0d0000000000000000;\n\t"
            "mov.b64 {lo,hi}, %0;\n\tmov.b64 {slo,shi}, %2;\n\t"
            "@p mad.lo.u32 lo, slo, %3, %4;\n\t@p mad.lo.u32 hi, shi, %3, %4;\n\t"
            "mov.b64 %0, {lo,hi}; }"
            : "+d"(dst) : "d"(u), "d"(src), "r"(c_sel_one), "r"(c_sel_zero));
    }
}
__device__ __forceinline__ void psel_lt(double &dst, double u)
{
    asm("{ .reg .pred p; .reg .b32 lo, hi, slo, shi;\n\t"
        "setp.lt.f64 p, %1, %0;\n\t"
        "mov.b64 {lo,hi}, %0;\n\tmov.b64 {slo,shi}, %1;\n\t"
        "@p mad.lo.u32 lo, slo, %2, %3;\n\t@p mad.lo.u32 hi, shi, %2, %3;\n\t"
        "mov.b64 %0, {lo,hi}; }"
        : "+d"(dst) : "d"(u), "r"(c_sel_one), "r"(c_sel_zero));
}
#endif

// SEL: bit 0 = peak, bit 1 = drawdown updated by predicated IMAD moves (device only); 0 = plain selects.
// INT_SIGN: decide "u > 0" (new equity peak) by a signed integer test of u's HIGH WORD (ISETP, ALU pipe) instead of an
// fp64 compare (DSETP, FP64 pipe -- the busiest pipe of the kernel).  The two agree for every double except positive
// subnormals (high word 0, low word != 0) and NaN.  The categorical kernels may use it: build_spec admits only trade
// values that are 0 or at least 2^-969 in magnitude, so every contribution, equity and equity difference is an integer
// multiple of 2^-1021 -- never a non-zero subnormal -- and it bounds the sums so that no NaN can arise.  The replay
// kernels (arbitrary caller values) keep the fp64 compare.
// MEASURED SLOWER on the B200 and therefore OFF: k_sim_paths 20.02 ms with the integer test against 19.76 ms with DSETP
// (each with its own searched pass lists, same SASS length) -- the FP64 pipe at 58 % is not what holds the kernel back,
// the extra ALU-pipe instruction is.  Kept as a compile-time switch with its tests, not shipped.
#ifndef MCTS_INT_SIGN
#define MCTS_INT_SIGN 0
#endif
MCTS_HD bool gt_zero_hi(double u)
{
#if defined(__CUDA_ARCH__)
    return __double2hiint(u) > 0;
#else
    return u > 0.0;
#endif
}

// MCTS_STEP_ORDER / MCTS_NEST: source-order experiments (the arithmetic is the same; ptxas' schedule and register allocation are
// not).  0 / 0 is what ships; the others were timed with tools/autotune and were not faster.
#ifndef MCTS_STEP_ORDER
#define MCTS_STEP_ORDER 0
#endif
#ifndef MCTS_NEST
#define MCTS_NEST 0
#endif
template <int SEL = 0, bool INT_SIGN = false>
MCTS_HD void trade_step(double &E, double &peak, double &dd, double &rj, double c)
{
#if MCTS_STEP_ORDER == 1
    rj = rj + c;
    E = E + c;
#else
    E = E + c;
    rj = rj + c;
#endif
    const double u = E - peak;
#if defined(__CUDA_ARCH__)
    // peak before drawdown: the other order costs ptxas ~2 more register moves per step
#if MCTS_STEP_ORDER == 2
    if constexpr ((SEL & 2) != 0) psel_lt(dd, u); else dd = (u < dd) ? u : dd;
    if constexpr ((SEL & 1) != 0) psel_gt0<INT_SIGN>(peak, u, E);
    else peak = (INT_SIGN ? gt_zero_hi(u) : (u > 0.0)) ? E : peak;
#else
    if constexpr ((SEL & 1) != 0) psel_gt0<INT_SIGN>(peak, u, E);
    else peak = (INT_SIGN ? gt_zero_hi(u) : (u > 0.0)) ? E : peak;
    if constexpr ((SEL & 2) != 0) psel_lt(dd, u); else dd = (u < dd) ? u : dd;
#endif
#else
    peak = (u > 0.0) ? E : peak;
    dd = (u < dd) ? u : dd;
#endif
}

// A full block of 8 trades.  SRC provides, per trade i, the three tabulated contributions.
// SINK receives the equity of every (strategy slot, trade) when curves are exported.
struct NoSink {
    template <int S>
    MCTS_HD void put(int, double)
    {
    }
};

template <class IDS, bool BINARY, int SEL, class SRC, class SINK>
MCTS_HD void eval_block8(Acc<IDS::N> &a, uint32_t W, SRC &src, SINK &sink, int t0)
{
    const Masks m = BINARY ? make_masks_binary(W, a.h) : make_masks(W, W ^ 0x11111111u, a.h);
#if MCTS_NEST == 1
    if constexpr (std::is_same<SINK, NoSink>::value) {  // strategy-major source order (experiment)
        double v1[8], v2[8], v3[8];
        static_for<8>([&](auto I) {
            constexpr int i = decltype(I)::value;
            src.template get<i>(v1[i], v2[i], v3[i]);
        });
        static_for<IDS::N>([&](auto S) {
            constexpr int s = decltype(S)::value;
            constexpr int rule = rule_of_id(IDS::id(s));
            static_for<8>([&](auto I) {
                constexpr int i = decltype(I)::value;
                const Preds p = trade_preds<BINARY>(m, mask_bit(i));
                const double c = contribution<rule, SRC::kSharedLo>(p, v1[i], v2[i], v3[i]);
                trade_step<(i < (SEL >> 2) ? 0 : (SEL & 3)), SRC::kIntSign>(a.E[s], a.peak[s], a.dd[s], a.r[s][i], c);
            });
        });
        return;
    }
#endif
    static_for<8>([&](auto I) {
        constexpr int i = decltype(I)::value;
        double v1, v2, v3;
        src.template get<i>(v1, v2, v3);
        const Preds p = trade_preds<BINARY>(m, mask_bit(i));
        static_for<IDS::N>([&](auto S) {
            constexpr int s = decltype(S)::value;
            constexpr int rule = rule_of_id(IDS::id(s));
            const double c = contribution<rule, SRC::kSharedLo>(p, v1, v2, v3);
            // ptxas turns the predicated moves of the first trades of a loop body into copy + select (the state
            // crosses the back edge); plain selects cost the same there and need no copies
            // (SEL >> 2 = how many leading trades of the block use plain selects)
            trade_step<(i < (SEL >> 2) ? 0 : (SEL & 3)), SRC::kIntSign>(a.E[s], a.peak[s], a.dd[s], a.r[s][i], c);
            sink.template put<s>(t0 + i, a.E[s]);
        });
    });
}

// The < 8 trades after the last full block: added one by one to the leaf result (the n % 8 loop
// of the pairwise leaf, or the whole n < 8 case).  W/Lb bits beyond `count` are ignored.
template <class IDS, int SEL, class SRC, class SINK>
MCTS_HD void eval_tail(Acc<IDS::N> &a, double (&res)[IDS::N], uint32_t W, uint32_t Lb, int count, SRC &src, SINK &sink, int t0)
{
    const Masks m = make_masks(W, Lb, a.h);
    for (int i = 0; i < count; i++) {
        double v1, v2, v3;
        src.get_dyn(i, v1, v2, v3);
        const Preds p = trade_preds<false>(m, mask_bit(i));
        static_for<IDS::N>([&](auto S) {
            constexpr int s = decltype(S)::value;
            constexpr int rule = rule_of_id(IDS::id(s));
            const double c = contribution<rule, SRC::kSharedLo>(p, v1, v2, v3);
            trade_step<(SEL & 3), SRC::kIntSign>(a.E[s], a.peak[s], a.dd[s], res[s], c);
            sink.template put<s>(t0 + i, a.E[s]);
        });
    }
}

// Evaluate the strategies of one pass over a whole path.
//   SRC: uint32_t load(int oct)       -> win bits of block `oct` (and latches its values)
//        uint32_t load_tail(int oct, int count, uint32_t *loss_bits)
//        get<i>() / get_dyn(i)
//   OPAQUE: every id's chains start from constant memory (ids of one class in one pass stay separate chains)
//   returns true when an equity or a profit sum left the finite range (non-finite input values, or sums beyond
//   DBL_MAX): the reference's drawdown is NaN then (np.min propagates it) while the ordered comparisons used here
//   drop NaNs, so such paths are outside the bit-exactness contract and the replay entry point rejects them.
template <class IDS, int SEL = 0, bool OPAQUE = false, class SRC, class SINK>
MCTS_HD bool eval_path(SRC &src, SINK &sink, const LeafSched *__restrict__ sch, double (&profit)[IDS::N], double (&drawdown)[IDS::N])
{
    constexpr int NS = IDS::N;
    Acc<NS> a;
    acc_init<IDS, OPAQUE>(a);
    double stk[NS][kMaxStackDepth];
    double res[NS];
    int sp = 0;
    int oct = 0, leaf_start = 0;
    const int n_leaves = sch->n_leaves;
    if (sch->n_oct > 0) {  // first block: three-valued history ("nothing before the first trade")
        const uint32_t W = src.load(0);
        eval_block8<IDS, false, SEL>(a, W, src, sink, 0);
        oct = 1;
    }
    for (int leaf = 0; leaf < n_leaves; leaf++) {
        const int end = sch->end_oct[leaf];
        const bool has_blocks = leaf_start < end;
        leaf_start = end;
        for (; oct < end; oct++) {
            const uint32_t W = src.load(oct);
            eval_block8<IDS, true, SEL>(a, W, src, sink, oct * 8);
        }
#pragma unroll
        for (int s = 0; s < NS; s++) {
            // leaf result: ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)); a leaf without a full block starts from 0.
            res[s] = has_blocks ? ((a.r[s][0] + a.r[s][1]) + (a.r[s][2] + a.r[s][3])) + ((a.r[s][4] + a.r[s][5]) + (a.r[s][6] + a.r[s][7])) : 0.0;
        }
        acc_reset_leaf<IDS, OPAQUE>(a);
        if (leaf == n_leaves - 1 && sch->tail > 0) {
            uint32_t Lb;
            const uint32_t W = src.load_tail(oct, sch->tail, &Lb);
            eval_tail<IDS, SEL>(a, res, W, Lb, sch->tail, src, sink, oct * 8);
        }
        int merges = sch->merges[leaf];
#pragma unroll
        for (int s = 0; s < NS; s++) stk[s][sp] = res[s];
        sp++;
        for (; merges > 0; merges--) {
            sp--;
#pragma unroll
            for (int s = 0; s < NS; s++) stk[s][sp - 1] = stk[s][sp - 1] + stk[s][sp];
        }
    }
    bool nonfinite = false;
#pragma unroll
    for (int s = 0; s < NS; s++) {
        profit[s] = 0.0 + stk[s][0];  // np.add.reduce starts from the identity 0.0
        drawdown[s] = a.dd[s];
        // x - x is 0 for finite x and NaN for +-inf / NaN; a non-finite equity never becomes finite again
        nonfinite = nonfinite || (a.E[s] - a.E[s] != 0.0) || (profit[s] - profit[s] != 0.0);
    }
    return nonfinite;
}

// ---- exact int32 evaluator -----------------------------------------------------------------------
// When every trade value is an integer and 6 * max|v| * L < 2^31 (checked on the host), equity, peak
// and drawdown are exact integers below 2^30 in magnitude, every summation order gives the same
// value (profit == final equity) and the fp64 results of the reference are reproduced exactly by
// converting the int32 results at the end.  Same masks, same rules.
//
// One pass evaluates ALL strategies of the launch (20 ids = 60 registers of state), so the code word,
// the masks, the table lookup and the predicate tests of a trade are paid once per trade, not once per
// five strategies.  The contribution is never selected: the equity takes one unconditional and up to
// two predicated integer multiply-adds (E += v*1, on the FMA pipe: `one` comes from constant memory so
// that ptxas keeps the IMAD), which leaves only the running minimum and maximum (VIADDMNMX + VIMNMX)
// on the ALU pipe: ~2.3 ALU and ~1.7 FMA instructions per trade-step.
struct IntTab {
    int32_t v1, v2;  // v, 2v
};

#if defined(__CUDACC__)
static __constant__ int32_t c_int_one = 1;
// Twenty zeros the compiler cannot see through: start values of the equities when all 20 ids are to be
// evaluated for real (ids of one behaviour class would otherwise be merged into one chain).
static __constant__ int32_t c_int_zero[kNumStrategies] = {};
#endif

MCTS_HD void iadd(int32_t &E, int32_t v)
{
#if defined(__CUDA_ARCH__)
    asm("mad.lo.s32 %0, %1, %2, %0;" : "+r"(E) : "r"(v), "r"(c_int_one));
#else
    E += v;
#endif
}
// E += v when bit `bit` of the history mask `mask` is set (clear, if NEG).  The bit test lives inside the
// asm so that it stays ONE LOP3-to-predicate per (trade, mask) -- ptxas merges the identical tests of the
// strategies that share it; handed over as a bool it became shift + and + compare per use.
template <bool NEG>
MCTS_HD void iadd_if(int32_t &E, int32_t v, uint32_t mask, uint32_t bit)
{
#if defined(__CUDA_ARCH__)
    if constexpr (NEG)
        asm("{ .reg .pred q; .reg .b32 t;\n\tand.b32 t, %3, %4;\n\tsetp.eq.u32 q, t, 0;\n\t@q mad.lo.s32 %0, %1, %2, %0; }"
            : "+r"(E) : "r"(v), "r"(c_int_one), "r"(mask), "r"(bit));
    else
        asm("{ .reg .pred q; .reg .b32 t;\n\tand.b32 t, %3, %4;\n\tsetp.ne.u32 q, t, 0;\n\t@q mad.lo.s32 %0, %1, %2, %0; }"
            : "+r"(E) : "r"(v), "r"(c_int_one), "r"(mask), "r"(bit));
#else
    if (((mask & bit) != 0) != NEG) E += v;
#endif
}

// E += size * v, size in {0, 1, 2, 3} by the behaviour's rule (cf. contribution<RULE>).  BINARY: from the
// second block on "lost" is "did not win" (pl = !pw) and s16 = pll (make_masks_binary).
template <int RULE, bool BINARY>
MCTS_HD void add_contribution_i(int32_t &E, const Masks &m, uint32_t bit, int32_t v1, int32_t v2)
{
    if constexpr (RULE == R_C1) iadd(E, v1);
    if constexpr (RULE == R_AM) { iadd(E, v1); iadd_if<false>(E, v1, m.pw, bit); }
    if constexpr (RULE == R_F1) { iadd(E, v1); iadd_if<false>(E, v1, m.fw, bit); }
    if constexpr (RULE == R_MG) {
        iadd(E, v1);
        if constexpr (BINARY) iadd_if<true>(E, v1, m.pw, bit); else iadd_if<false>(E, v1, m.pl, bit);
    }
    if constexpr (RULE == R_MG2) { iadd(E, v1); iadd_if<false>(E, v1, m.pll, bit); }
    if constexpr (RULE == R_P1) iadd_if<true>(E, v1, m.pw, bit);
    if constexpr (RULE == R_P2) iadd_if<true>(E, v1, m.pww, bit);
    if constexpr (RULE == R_S16) {  // s16 implies !pw
        iadd_if<true>(E, v1, m.pw, bit);
        iadd_if<false>(E, v1, BINARY ? m.pll : m.s16, bit);
    }
    if constexpr (RULE == R_S18) { iadd(E, v1); iadd_if<false>(E, v2, m.pww, bit); }
    if constexpr (RULE == R_S19) { iadd(E, v1); iadd_if<false>(E, v1, m.pww, bit); iadd_if<false>(E, v2, m.pll, bit); }  // pww, pll exclusive
    if constexpr (RULE == R_S20) { iadd(E, v1); iadd_if<false>(E, v2, m.pll, bit); }
}

MCTS_HD int32_t imin(int32_t a, int32_t b) { return a < b ? a : b; }
MCTS_HD int32_t imax(int32_t a, int32_t b) { return a > b ? a : b; }

using FullAll = IdList<1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20>;
using DedupAll = IdList<1, 2, 4, 6, 8, 9, 11, 16, 18, 19, 20>;

template <class IDS, bool OPAQUE_START = false>
MCTS_HD void eval_path_int(const uint32_t *codes, int stride, const IntTab *tab, int L, int32_t (&profit)[IDS::N], int32_t (&drawdown)[IDS::N])
{
    constexpr int NS = IDS::N;
    int32_t E[NS], peak[NS], dd[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) {
        E[s] = 0;
#if defined(__CUDA_ARCH__)
        if constexpr (OPAQUE_START) E[s] = c_int_zero[s];
#endif
        peak[s] = -(1 << 30);  // below any reachable equity; E_0 - peak cannot overflow
        dd[s] = 0;
    }
    History h;
    h.w = 0;
    h.l = 0;
    const int n_oct = L >> 3, tail = L & 7;
    auto block = [&](int oct, auto B) {
        constexpr bool BINARY = decltype(B)::value;
        const uint32_t word = codes[(size_t)oct * stride];
        const uint32_t W = word & 0x11111111u;
        const Masks m = BINARY ? make_masks_binary(W, h) : make_masks(W, W ^ 0x11111111u, h);
        static_for<8>([&](auto I) {
            constexpr int i = decltype(I)::value;
            const IntTab t = tab[(word >> (4 * i)) & 15u];
            static_for<NS>([&](auto S) {
                constexpr int s = decltype(S)::value;
                constexpr int rule = rule_of_id(IDS::id(s));
                add_contribution_i<rule, BINARY>(E[s], m, mask_bit(i), t.v1, t.v2);
                dd[s] = imin(dd[s], E[s] - peak[s]);
                peak[s] = imax(peak[s], E[s]);
            });
        });
    };
    if (n_oct > 0) block(0, std::false_type{});
    for (int oct = 1; oct < n_oct; oct++) block(oct, std::true_type{});
    if (tail) {
        const uint32_t word = codes[(size_t)n_oct * stride];
        const uint32_t valid = 0x11111111u & ((1u << (4 * tail)) - 1u);
        const uint32_t W = word & valid;
        const Masks m = make_masks(W, W ^ valid, h);
        for (int i = 0; i < tail; i++) {
            const IntTab t = tab[(word >> (4 * i)) & 15u];
            static_for<NS>([&](auto S) {
                constexpr int s = decltype(S)::value;
                constexpr int rule = rule_of_id(IDS::id(s));
                add_contribution_i<rule, false>(E[s], m, mask_bit(i), t.v1, t.v2);
                dd[s] = imin(dd[s], E[s] - peak[s]);
                peak[s] = imax(peak[s], E[s]);
            });
        }
    }
#pragma unroll
    for (int s = 0; s < NS; s++) {
        profit[s] = E[s];
        drawdown[s] = dd[s];
    }
}

// ---- value sources ----------------------------------------------------------------------------

// The value table as the kernels keep it in shared memory: three arrays with 8-byte stride, so the 16
// entries of each span exactly the 32 banks once -- lanes reading different codes never conflict, lanes
// reading the same code broadcast.  (The array-of-structs Tab of spec.hpp has a 32-byte stride: codes k and
// k + 4 alias, 1.1e9 bank conflicts per launch in the round-1 profile.)
struct TabSoA {
    double v1[16], v2[16], v3[16];
};
MCTS_HD void tab_get(const Tab *tab, uint32_t code, double &v1, double &v2, double &v3)
{
    const Tab &t = tab[code];
    v1 = t.v1;
    v2 = t.v2;
    v3 = t.v3;
}
MCTS_HD void tab_get(const TabSoA *tab, uint32_t code, double &v1, double &v2, double &v3)
{
    v1 = tab->v1[code];
    v2 = tab->v2[code];
    v3 = tab->v3[code];
}

// categorical: 8 trades per 32-bit word, 4-bit code each; code -> {v, 2v, fl(3v)} table
template <class TAB>
struct CodeSourceT {
    static constexpr bool kSharedLo = true;  // build_spec rejects subnormal / overflowing values
    static constexpr bool kIntSign = MCTS_INT_SIGN != 0;  // ... and values below 2^-969: see trade_step
    const uint32_t *codes;  // word of block `oct` at codes[oct * stride]
    int stride;
    const TAB *tab;
    uint32_t word;
    MCTS_HD uint32_t load(int oct)
    {
        word = codes[(size_t)oct * stride];
        return word & 0x11111111u;
    }
    MCTS_HD uint32_t load_tail(int oct, int count, uint32_t *loss_bits)
    {
        word = codes[(size_t)oct * stride];
        const uint32_t valid = count >= 8 ? 0x11111111u : (0x11111111u & ((1u << (4 * count)) - 1u));
        const uint32_t W = word & valid;
        *loss_bits = W ^ valid;
        return W;
    }
    template <int I>
    MCTS_HD void get(double &v1, double &v2, double &v3) const
    {
        tab_get(tab, (word >> (4 * I)) & 15u, v1, v2, v3);
    }
    MCTS_HD void get_dyn(int i, double &v1, double &v2, double &v3) const
    {
        tab_get(tab, (word >> (4 * i)) & 15u, v1, v2, v3);
    }
};
using CodeSource = CodeSourceT<Tab>;        // host harness: GenSpec::tab as it is
using CodeSourceSoA = CodeSourceT<TabSoA>;  // kernels: shared-memory copy

MCTS_HD double mul_rn(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;  // host harness is compiled with -ffp-contract=off
#endif
}

// replay: arbitrary fp64 trade results of one path, contiguous in memory
struct ValueSource {
    static constexpr bool kSharedLo = false;  // arbitrary caller values
    static constexpr bool kIntSign = false;
    const double *row;
    double v[8];
    MCTS_HD uint32_t load(int oct)
    {
        uint32_t W = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            v[i] = row[oct * 8 + i];
            W |= (v[i] > 0.0 ? 1u : 0u) << (4 * i);  // win <=> value > 0 (:154)
        }
        return W;
    }
    MCTS_HD uint32_t load_tail(int oct, int count, uint32_t *loss_bits)
    {
        uint32_t W = 0, V = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            v[i] = i < count ? row[oct * 8 + i] : 0.0;
            W |= ((i < count && v[i] > 0.0) ? 1u : 0u) << (4 * i);
            V |= (i < count ? 1u : 0u) << (4 * i);
        }
        *loss_bits = W ^ V;
        return W;
    }
    template <int I>
    MCTS_HD void get(double &v1, double &v2, double &v3) const
    {
        v1 = v[I];
        v2 = mul_rn(v[I], 2.0);
        v3 = mul_rn(v[I], 3.0);
    }
    MCTS_HD void get_dyn(int i, double &v1, double &v2, double &v3) const
    {
        double x = v[0];
#pragma unroll
        for (int j = 1; j < 8; j++) x = (i == j) ? v[j] : x;
        v1 = x;
        v2 = mul_rn(x, 2.0);
        v3 = mul_rn(x, 3.0);
    }
};

}  // namespace mcts
