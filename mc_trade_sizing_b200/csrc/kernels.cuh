// kernels.cuh -- the sm_100a kernels of the path-simulation core.
//
//   k_sim_paths       THE hot kernel, one thread per path (= simulation x shuffle, :491-507):
//                     shuffle the simulation's outcome multiset (Fisher-Yates; urn for 2 values) into a
//                     thread-private code column (4 bits per trade; shared memory, or global memory for
//                     paths longer than ~3,500 trades), then evaluate the 20 strategies in passes of
//                     five with equity / peak / drawdown / profit accumulators in fp64 registers;
//                     per-path results go to HBM once (320 B per path = 0.016 B per trade-step).
//                     MODE: all 20 ids, the 11 distinct behaviours (dedup), or all 20 + equity curves
//   k_sim_paths_int   the same with the exact int32 evaluator (all-integer runs)
//   k_replay_paths    the fp64 evaluator on caller-supplied sequences (replay mode): rows staged through shared memory
//                     by cp.async in transposed tiles, all passes in one launch
//   k_sequences       audit hook: the generated sequences as doubles
// The non-template kernels (k_gen_base, k_reduce_nodes/top, k_shuffle_values, k_peak_*) live in aux_kernels.cuh,
// which only mcts.cu includes, so they are compiled once and not into every paths_*.cu translation unit.
#pragma once
#include <cuda_runtime.h>

#include "../../include/mcts.h"
#include "eval_core.cuh"
#include "gen_core.cuh"
#include "pairwise.cuh"
#include "spec.hpp"

namespace mcts {

constexpr int kBlock = 64;         // threads per block of the replay / audit kernels
constexpr int kBlocksPerSM = 6;    // 384 threads/SM -> up to 168 registers per thread
constexpr int kThreadsPerSM = 384;
constexpr int kTabWords = 16 * sizeof(Tab) / 4;

enum PathMode : int { PM_FULL = 0, PM_DEDUP = 1, PM_CURVES = 2 };

// How peak / drawdown are updated in each kernel flavour (eval_core.cuh trade_step<SEL>): predicated
// IMAD moves on the FMA pipe (3) where that measured faster on the B200, select pairs (0) elsewhere;
// + 4 * (number of leading trades of every 8-trade block that keep the select pairs).
#ifndef MCTS_SEL_FULL
#define MCTS_SEL_FULL 7    // predicated moves (+4.4 %: the select version saturates the ALU pipe), first trade of a block: selects
                           // (k_sim_paths per 2.1 M paths, each with its own searched pass orders: 0 -> 21.72 ms, 1 -> 21.09, 2 -> 21.49, 3 -> 21.68)
#endif
#ifndef MCTS_SEL_DEDUP
#define MCTS_SEL_DEDUP 7   // +10 % with the searched pass composition (10 % slower with the first one)
#endif
#ifndef MCTS_SEL_CURVES
#define MCTS_SEL_CURVES 0  // HBM-bound
#endif
#ifndef MCTS_SEL_REPLAY
#define MCTS_SEL_REPLAY 0
#endif
__host__ __device__ constexpr int sel_of_mode(int mode) { return mode == PM_FULL ? MCTS_SEL_FULL : mode == PM_DEDUP ? MCTS_SEL_DEDUP : MCTS_SEL_CURVES; }

struct PathArgs {
    GenSpec g;
    uint64_t seed;
    int64_t path0;  // global index of the first path of this launch
    int64_t n;      // paths in this launch
    int64_t sim0;   // first simulation in `counts`
    int32_t n_sh;   // num_mc_shuffles
    int32_t pad;
    const uint32_t *counts;   // [sims][kMaxCat]
    const LeafSched *sched;
    double *out;              // [40][ld]: row id-1 = profit, row 20+id-1 = drawdown
    int64_t ld;
    double *curves;           // PM_CURVES: [20][ceil(L/2)][n][2]
    uint32_t *scratch;        // GLOBAL code store (paths too long for shared memory): [word][scratch_stride]
    int32_t scratch_stride;   // threads in the launch, rounded up to the block size
    int32_t pad1;
    int64_t out_off;          // column offset of this launch inside `out` (sub-launches of one chunk)
    uint32_t *flags;          // replay: bit 0 is set when a path's equity / profit left the finite range
};

// equity-curve sink: two consecutive trades per 128-bit store, paths contiguous
template <class IDS>
struct CurveSink {
    double2 *base;    // curves as double2 [20][half][n]
    int64_t n, local;
    int32_t half, L;
    double hold[IDS::N];
    template <int S>
    __device__ __forceinline__ void put(int t, double E)
    {
        if ((t & 1) == 0) {
            hold[S] = E;
            if (t != L - 1) return;
            E = 0.0;  // odd L: pad the last pair
            t++;
        }
        constexpr int id = IDS::id(S);
        double2 v = make_double2(hold[S], E);
        __stcs(base + ((int64_t)(id - 1) * half + (t >> 1)) * n + local, v);
    }
};

template <class IDS, int MODE, int SEL, class SRC>
__device__ __forceinline__ void run_pass(SRC &src, const PathArgs &a, int64_t local)
{
    double profit[IDS::N], dd[IDS::N];
    if constexpr (MODE == PM_CURVES) {
        CurveSink<IDS> sink;
        sink.base = reinterpret_cast<double2 *>(a.curves);
        sink.n = a.n;
        sink.local = local;
        sink.L = a.g.L;
        sink.half = (a.g.L + 1) >> 1;
        eval_path<IDS, SEL, false>(src, sink, a.sched, profit, dd);
    } else {
        NoSink sink;
        eval_path<IDS, SEL, MODE == PM_FULL>(src, sink, a.sched, profit, dd);  // PM_FULL: 20 separate chains
    }
    if (a.out == nullptr) return;
    static_for<IDS::N>([&](auto S) {
        constexpr int s = decltype(S)::value;
        constexpr int id = IDS::id(s);
        if constexpr (MODE == PM_DEDUP) {
#pragma unroll
            for (int other = 1; other <= kNumStrategies; other++) {
                if (rule_of_id(other) == rule_of_id(id)) {
                    a.out[(int64_t)(other - 1) * a.ld + a.out_off + local] = profit[s];
                    a.out[(int64_t)(kNumStrategies + other - 1) * a.ld + a.out_off + local] = dd[s];
                }
            }
        } else {
            a.out[(int64_t)(id - 1) * a.ld + a.out_off + local] = profit[s];
            a.out[(int64_t)(kNumStrategies + id - 1) * a.ld + a.out_off + local] = dd[s];
        }
    });
}

// replay flavour of run_pass: caller-supplied values (finite check), stores only for `store` threads
template <class IDS, int MODE, class SRC>
__device__ __forceinline__ void run_pass_replay(SRC &src, const PathArgs &a, int64_t local, bool store)
{
    double profit[IDS::N], dd[IDS::N];
    NoSink sink;
    const bool nonfinite = eval_path<IDS, MCTS_SEL_REPLAY, MODE == PM_FULL>(src, sink, a.sched, profit, dd);
    if (!store) return;
    if (nonfinite && a.flags != nullptr) atomicOr(a.flags, 1u);
    if (a.out == nullptr) return;
    static_for<IDS::N>([&](auto S) {
        constexpr int s = decltype(S)::value;
        constexpr int id = IDS::id(s);
        if constexpr (MODE == PM_DEDUP) {
#pragma unroll
            for (int other = 1; other <= kNumStrategies; other++) {
                if (rule_of_id(other) == rule_of_id(id)) {
                    a.out[(int64_t)(other - 1) * a.ld + a.out_off + local] = profit[s];
                    a.out[(int64_t)(kNumStrategies + other - 1) * a.ld + a.out_off + local] = dd[s];
                }
            }
        } else {
            a.out[(int64_t)(id - 1) * a.ld + a.out_off + local] = profit[s];
            a.out[(int64_t)(kNumStrategies + id - 1) * a.ld + a.out_off + local] = dd[s];
        }
    });
}

// One pass of k_sim_paths as a function of its own (MCTS_NOINLINE_PASSES): ptxas then allocates the registers of every
// 8-trade loop body independently of the other passes, which makes the pass search of tools/pass_search (one pass compiled
// alone) predict the code of the whole kernel.  Four calls per path; the source is passed by value.
#ifndef MCTS_NOINLINE_PASSES
#define MCTS_NOINLINE_PASSES 0
#endif
template <class IDS, int MODE, int SEL, class SRC>
__device__ __noinline__ void run_pass_call(SRC src, const PathArgs &a, int64_t local)
{
    run_pass<IDS, MODE, SEL>(src, a, local);
}
template <class IDS, int MODE, int SEL, class SRC>
__device__ __forceinline__ void run_pass_k(SRC &src, const PathArgs &a, int64_t local)
{
    if constexpr (MCTS_NOINLINE_PASSES != 0 && MODE != PM_CURVES) run_pass_call<IDS, MODE, SEL>(src, a, local);
    else run_pass<IDS, MODE, SEL>(src, a, local);
}

// Phase 0 of the path kernels: shuffle the multiset of this thread's simulation into its code column.
template <int K>
__device__ __forceinline__ void generate_codes(const PathArgs &a, int64_t local, uint32_t *column, int stride)
{
    const int64_t path = a.path0 + local;
    const int64_t sim = path / a.n_sh;
    const uint32_t k = (uint32_t)(path - sim * a.n_sh);
    const uint32_t *cnt = a.counts + (sim - a.sim0) * kMaxCat;
    if constexpr (K <= kMaxUrnValues) {  // a handful of values: the predicated urn, no array access at all (K = urn_width(g.K))
        uint32_t c[K];
#pragma unroll
        for (int j = 0; j < K; j++) c[j] = __ldg(cnt + j);   // counts beyond the run's own values are zero
        gen_path_codes_urn<K>(a.g, a.seed, (uint32_t)sim, k, c, column, stride);
    } else {  // up to 16 values: Fisher-Yates on the packed column
        gen_path_codes_fy(a.g, a.seed, (uint32_t)sim, k, cnt, column, stride);
    }
}

// BLOCK threads per block, 384 threads per SM.  The threads of a block do not exchange data; the
// barriers between phases only keep the warps of a block inside the same ~8 KB loop body so that the
// instruction caches see few bodies at a time: the kernel is ~100 KB of unrolled code, and with
// 64-thread blocks and no barriers 47 % of the issue slots were lost to instruction fetch
// (profiles/r01_ncu_sim_paths_block64_nobarrier.txt).  128 threads (one warp per scheduler, three
// blocks per SM) measured best; 384 (one block per SM) is 3 % slower.
// GLOBAL = true keeps the code columns in global memory ([word][thread], coalesced per word) for
// paths too long for shared memory; same code otherwise, much slower generation (scattered RMW).
template <int K, int MODE, int BLOCK, bool GLOBAL = false>
__global__ void __launch_bounds__(BLOCK, kThreadsPerSM / BLOCK) k_sim_paths(const __grid_constant__ PathArgs a)
{
    extern __shared__ __align__(16) uint32_t smem[];
    TabSoA *tab = reinterpret_cast<TabSoA *>(smem);
    const int tid = threadIdx.x;
    uint32_t *codes = GLOBAL ? a.scratch + (size_t)blockIdx.x * BLOCK : smem + kTabWords;
    const int cstride = GLOBAL ? a.scratch_stride : BLOCK;
    if (tid < 16) {
        tab->v1[tid] = a.g.tab[tid].v1;
        tab->v2[tid] = a.g.tab[tid].v2;
        tab->v3[tid] = a.g.tab[tid].v3;
    }
    __syncthreads();
    const int64_t local = (int64_t)blockIdx.x * BLOCK + tid;
    const bool active = local < a.n;
    // the code column of a thread is private: the barriers below are for phase alignment only
    if (active) generate_codes<K>(a, local, codes + tid, cstride);
    CodeSourceSoA src;
    src.codes = codes + tid;
    src.stride = cstride;
    src.tab = tab;
    src.word = 0;
    if constexpr (MODE == PM_DEDUP) {
        __syncthreads();
        if (active) run_pass_k<DedupPass0, MODE, sel_of_mode(MODE)>(src, a, local);
        __syncthreads();
        if (active) run_pass_k<DedupPass1, MODE, sel_of_mode(MODE)>(src, a, local);
        __syncthreads();
        if (active) run_pass_k<DedupPass2, MODE, sel_of_mode(MODE)>(src, a, local);
    } else if constexpr (MODE == PM_CURVES) {
        __syncthreads();
        if (active) run_pass_k<CurvePass0, MODE, sel_of_mode(MODE)>(src, a, local);
        __syncthreads();
        if (active) run_pass_k<CurvePass1, MODE, sel_of_mode(MODE)>(src, a, local);
        __syncthreads();
        if (active) run_pass_k<CurvePass2, MODE, sel_of_mode(MODE)>(src, a, local);
        __syncthreads();
        if (active) run_pass_k<CurvePass3, MODE, sel_of_mode(MODE)>(src, a, local);
    } else {
        __syncthreads();
        if (active) run_pass_k<FullPass0, MODE, sel_of_mode(MODE)>(src, a, local);
        __syncthreads();
        if (active) run_pass_k<FullPass1, MODE, sel_of_mode(MODE)>(src, a, local);
        __syncthreads();
        if (active) run_pass_k<FullPass2, MODE, sel_of_mode(MODE)>(src, a, local);
        __syncthreads();
        if (active) run_pass_k<FullPass3, MODE, sel_of_mode(MODE)>(src, a, local);
    }
}

// ---- exact int32 variant of the hot kernel (all-integer runs: Markov defaults, README regimes) -----
template <class IDS, int MODE>
__device__ __forceinline__ void run_pass_int(const uint32_t *codes, int stride, const IntTab *tab, const PathArgs &a, int64_t local)
{
    int32_t profit[IDS::N], dd[IDS::N];
    eval_path_int<IDS, MODE != PM_DEDUP>(codes, stride, tab, a.g.L, profit, dd);
    static_for<IDS::N>([&](auto S) {
        constexpr int s = decltype(S)::value;
        constexpr int id = IDS::id(s);
        if constexpr (MODE == PM_DEDUP) {
#pragma unroll
            for (int other = 1; other <= kNumStrategies; other++) {
                if (rule_of_id(other) == rule_of_id(id)) {
                    a.out[(int64_t)(other - 1) * a.ld + a.out_off + local] = (double)profit[s];
                    a.out[(int64_t)(kNumStrategies + other - 1) * a.ld + a.out_off + local] = (double)dd[s];
                }
            }
        } else {
            a.out[(int64_t)(id - 1) * a.ld + a.out_off + local] = (double)profit[s];
            a.out[(int64_t)(kNumStrategies + id - 1) * a.ld + a.out_off + local] = (double)dd[s];
        }
    });
}

template <int K, int MODE, int BLOCK>
__global__ void __launch_bounds__(BLOCK, kThreadsPerSM / BLOCK) k_sim_paths_int(const __grid_constant__ PathArgs a)
{
    extern __shared__ __align__(16) uint32_t smem[];
    IntTab *tab = reinterpret_cast<IntTab *>(smem);
    uint32_t *codes = smem + kTabWords;
    const int tid = threadIdx.x;
    if (tid < 16) {
        const int32_t v = a.g.ival[tid];
        IntTab t;
        t.v1 = v;
        t.v2 = 2 * v;
        tab[tid] = t;
    }
    __syncthreads();
    const int64_t local = (int64_t)blockIdx.x * BLOCK + tid;
    const bool active = local < a.n;
    if (active) generate_codes<K>(a, local, codes + tid, BLOCK);
    __syncthreads();  // phase alignment only (the code column is thread-private)
    if (!active) return;
    if constexpr (MODE == PM_DEDUP) run_pass_int<DedupAll, MODE>(codes + tid, BLOCK, tab, a, local);
    else run_pass_int<FullAll, MODE>(codes + tid, BLOCK, tab, a, local);
}

// ---- replay: the fp64 evaluator on caller-supplied sequences -------------------------------------------------
// values [n][T] row-major fp64 in device memory.  One thread owns one path, like k_sim_paths, but a thread reading its
// own row touches 32 different 128-byte lines per warp load (the round-1 kernel: 1.19e12 trade-steps/s, bound by L1
// tag look-ups).  Here a block of 128 paths streams its rows through shared memory in tiles of kReplayTile trades:
//   * every warp copies whole row segments -- 32 lanes x 8 bytes = 256 contiguous bytes of ONE path per instruction --
//     with cp.async (LDGSTS: global -> shared without a register round trip), transposing on the way into
//     tile[trade][path] with an odd pitch, so that neither the asynchronous writes (lane stride = pitch) nor the
//     evaluator's reads (lane stride = 1) conflict on shared-memory banks;
//   * two tiles: the copy of tile k+1 is in flight while tile k is evaluated (one barrier per tile);
//   * all passes of a mode run inside ONE launch (they re-stream the block's rows, which the previous pass left in L2
//     when the chunk is small enough), so a chunk is one launch instead of four.
#ifndef MCTS_REPLAY_WARP_TILES
#define MCTS_REPLAY_WARP_TILES 1
#endif
#ifndef MCTS_REPLAY_TILE
#define MCTS_REPLAY_TILE 32
#endif
#ifndef MCTS_REPLAY_STAGES
#define MCTS_REPLAY_STAGES 2
#endif
constexpr int kReplayBlock = 128;
constexpr int kReplayTile = MCTS_REPLAY_TILE;      // trades per tile: 8, 16 or 32 (one lane copies one trade of one row)
constexpr int kReplayStages = MCTS_REPLAY_STAGES;  // tiles per owner group: stages - 1 copies are in flight while one is evaluated
static_assert((kReplayTile == 8 || kReplayTile == 16 || kReplayTile == 32) && kReplayStages >= 2 && kReplayStages <= 8, "replay tile geometry");
#if MCTS_REPLAY_WARP_TILES
// every warp owns the tiles of its own 32 paths: no block barrier anywhere, the four warps of a block drift apart and
// hide each other's copy latency
constexpr int kReplayRows = 32;
#else
constexpr int kReplayRows = kReplayBlock;          // block-wide tiles, one __syncthreads per tile
#endif
constexpr int kReplayPitch = kReplayRows + 1;      // doubles per tile row: odd -> conflict-free transposing writes
constexpr size_t kReplaySmem = (kReplayBlock / kReplayRows) * kReplayStages * (size_t)kReplayTile * kReplayPitch * sizeof(double);   // 66 KB per block, 3 blocks per SM

struct TileSource {
    static constexpr bool kSharedLo = false;  // arbitrary caller values
    static constexpr bool kIntSign = false;
    const double *rows;   // first row of this block
    int64_t T;
    int n_rows;           // valid rows of this block (<= kReplayBlock)
    int n_tiles;
    double *tiles;        // [kReplayStages][kReplayTile][kReplayPitch] in shared memory
    int tid;              // the row this thread EVALUATES (clamped to the last valid row of the block)
    int copy_tid;         // threadIdx.x: this thread's place in the cooperative copy
    double v[8];

    __device__ __forceinline__ void issue(int tile)
    {
        if (tile < n_tiles) {
            // one copy instruction of a warp moves 32 / kReplayTile row segments of kReplayTile consecutive trades each
            constexpr int kRowsPerCopy = 32 / kReplayTile;
            const int lane = copy_tid & 31, tl = lane % kReplayTile, sub = lane / kReplayTile;
            const int64_t t = (int64_t)tile * kReplayTile + tl;
            double *dst = tiles + (size_t)(tile % kReplayStages) * kReplayTile * kReplayPitch + (size_t)tl * kReplayPitch;
            if (t < T) {
#if MCTS_REPLAY_WARP_TILES
                for (int r = sub; r < n_rows; r += kRowsPerCopy) {   // `rows`, `tiles` and `n_rows` are this WARP's
#else
                for (int r = (copy_tid >> 5) * kRowsPerCopy + sub; r < n_rows; r += (kReplayBlock / 32) * kRowsPerCopy) {
#endif
                    const unsigned saddr = (unsigned)__cvta_generic_to_shared(dst + r);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(saddr), "l"(rows + (int64_t)r * T + t) : "memory");
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    __device__ __forceinline__ void sync_owners()
    {
#if MCTS_REPLAY_WARP_TILES
        __syncwarp();
#else
        __syncthreads();
#endif
    }
    // start of a pass: nobody may still be reading the tiles of the previous pass
    __device__ __forceinline__ void begin_pass()
    {
        sync_owners();
#pragma unroll
        for (int k = 0; k < kReplayStages - 1; k++) issue(k);
    }
    // make tile `tile` readable and start the copy of the next one into the other buffer
    __device__ __forceinline__ void acquire(int tile)
    {
        asm volatile("cp.async.wait_group %0;" ::"n"(kReplayStages - 2) : "memory");  // this thread's share of `tile` has landed ...
        sync_owners();                                        // ... and every owner's; every owner is done with tile - 1
        issue(tile + kReplayStages - 1);
    }
    __device__ __forceinline__ const double *slot(int oct) const
    {
        const int tile = oct / (kReplayTile / 8), o = oct % (kReplayTile / 8);
        return tiles + (size_t)(tile % kReplayStages) * kReplayTile * kReplayPitch + (size_t)(o * 8) * kReplayPitch + tid;
    }
    __device__ __forceinline__ uint32_t load(int oct)
    {
        if (oct % (kReplayTile / 8) == 0) acquire(oct / (kReplayTile / 8));
        const double *p = slot(oct);
        uint32_t W = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            v[i] = p[(size_t)i * kReplayPitch];
            W |= (v[i] > 0.0 ? 1u : 0u) << (4 * i);  // win <=> value > 0 (:154)
        }
        return W;
    }
    __device__ __forceinline__ uint32_t load_tail(int oct, int count, uint32_t *loss_bits)
    {
        if (oct % (kReplayTile / 8) == 0) acquire(oct / (kReplayTile / 8));
        const double *p = slot(oct);
        uint32_t W = 0, V = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            v[i] = i < count ? p[(size_t)i * kReplayPitch] : 0.0;
            W |= ((i < count && v[i] > 0.0) ? 1u : 0u) << (4 * i);
            V |= (i < count ? 1u : 0u) << (4 * i);
        }
        *loss_bits = W ^ V;
        return W;
    }
    template <int I>
    __device__ __forceinline__ void get(double &v1, double &v2, double &v3) const
    {
        v1 = v[I];
        v2 = mul_rn(v[I], 2.0);
        v3 = mul_rn(v[I], 3.0);
    }
    __device__ __forceinline__ void get_dyn(int i, double &v1, double &v2, double &v3) const
    {
        double x = v[0];
#pragma unroll
        for (int j = 1; j < 8; j++) x = (i == j) ? v[j] : x;
        v1 = x;
        v2 = mul_rn(x, 2.0);
        v3 = mul_rn(x, 3.0);
    }
};

// Every thread of a tile's owner group (the warp, or the block with MCTS_REPLAY_WARP_TILES 0) runs every pass -- the copy is
// cooperative: threads beyond the last path of the launch evaluate the group's last valid row again and store nothing; a
// warp without any valid row leaves at once.
template <int MODE>
__global__ void __launch_bounds__(kReplayBlock, 3) k_replay_paths(const __grid_constant__ PathArgs a, const double *__restrict__ values, int64_t T)
{
    extern __shared__ __align__(16) double replay_tiles[];
    const int group = (int)threadIdx.x / kReplayRows, member = (int)threadIdx.x % kReplayRows;
    const int64_t first = (int64_t)blockIdx.x * kReplayBlock + (int64_t)group * kReplayRows;
    const int64_t left = a.n - first;
    if (left <= 0) return;  // whole owner group beyond the launch (never the case for block-wide tiles)
    TileSource src;
    src.rows = values + first * T;  // `values` already points at this launch's first path
    src.T = T;
    src.n_rows = (int)(left < kReplayRows ? left : kReplayRows);
    src.n_tiles = (int)((T + kReplayTile - 1) / kReplayTile);
    src.tiles = replay_tiles + (size_t)group * kReplayStages * kReplayTile * kReplayPitch;
    src.tid = member < src.n_rows ? member : src.n_rows - 1;
    src.copy_tid = member;
    const bool active = member < src.n_rows;
    const int64_t local = first + src.tid;
    auto pass = [&](auto IDS) {
        using L = decltype(IDS);
        src.begin_pass();
        // inactive threads still walk the pass (barriers); their stores are suppressed by a null output pointer
        run_pass_replay<L, MODE>(src, a, local, active);
    };
    if constexpr (MODE == PM_DEDUP) {
        pass(DedupPass0{});
        pass(DedupPass1{});
        pass(DedupPass2{});
    } else {
        pass(ReplayPass0{});
        pass(ReplayPass1{});
        pass(ReplayPass2{});
        pass(ReplayPass3{});
    }
}

template <int K, bool GLOBAL = false>
__global__ void __launch_bounds__(kBlock) k_sequences(const __grid_constant__ PathArgs a, double *__restrict__ values)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const int tid = threadIdx.x;
    uint32_t *codes = GLOBAL ? a.scratch + (size_t)blockIdx.x * kBlock : smem;
    const int cstride = GLOBAL ? a.scratch_stride : kBlock;
    const int64_t local = (int64_t)blockIdx.x * kBlock + tid;
    if (local >= a.n) return;
    generate_codes<K>(a, local, codes + tid, cstride);
    for (int t = 0; t < a.g.L; t++) {
        const uint32_t code = (codes[(size_t)(t >> 3) * cstride + tid] >> (4 * (t & 7))) & 15u;
        values[local * a.g.L + t] = a.g.tab[code].v1;
    }
}

}  // namespace mcts
