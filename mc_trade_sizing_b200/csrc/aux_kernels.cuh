// aux_kernels.cuh -- the non-template kernels of libmcts, included by mcts.cu ONLY (one translation unit):
//
//   k_gen_base          one thread per simulation: per-value counts of the base sequence (generators
//                       src/trading_models.py:43-128 on Philox uniforms)
//   k_shuffle_values    wide runs: Fisher-Yates on fp64 values in global memory
//   k_reduce_nodes/top  np.mean-ordered (pairwise) sums + min/max over the paths of a chunk (:513-518)
//   k_peak_dadd/int/step  issue-rate microbenchmarks (roofline denominators)
#pragma once
#include "kernels.cuh"

namespace mcts {

static __global__ void __launch_bounds__(128) k_gen_base(const __grid_constant__ GenSpec g, uint64_t seed, int64_t sim0, int64_t n_sims, uint32_t *__restrict__ counts)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_sims) return;
    uint32_t c[kMaxCat];
    gen_base_counts(g, seed, (uint32_t)(sim0 + i), c);
    uint4 *dst = reinterpret_cast<uint4 *>(counts + i * kMaxCat);
#pragma unroll
    for (int q = 0; q < kMaxCat / 4; q++) dst[q] = make_uint4(c[4 * q], c[4 * q + 1], c[4 * q + 2], c[4 * q + 3]);
}

// wide runs (more values than 4-bit codes): one thread per path shuffles its fp64 sequence in global
// memory, values[local][t]; k_replay_paths then evaluates it.  Rare fallback, not tuned.
static __global__ void __launch_bounds__(128) k_shuffle_values(const __grid_constant__ PathArgs a, double *__restrict__ values)
{
    const int64_t local = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (local >= a.n) return;
    const int64_t path = a.path0 + local;
    const int64_t sim = path / a.n_sh;
    const uint32_t k = (uint32_t)(path - sim * a.n_sh);
    gen_path_values_fy(a.g, a.seed, (uint32_t)sim, k, a.counts + (sim - a.sim0) * kMaxCat, values + local * a.g.L, 1);
}

// ---- exact-order reduction over the paths of one chunk --------------------------------------

// grid (ceil(2^d / 8), 40), 64-thread blocks: partial[arr][j] = raw pairwise sum of node j at depth d, one
// 8-lane group per node.  The eight lanes ARE the eight interleaved accumulators of NumPy's pairwise leaf
// (lane j adds a[j], a[8+j], a[16+j], ...), so every load instruction of a group reads 64 contiguous bytes
// (whole sectors, four groups per warp) and the leaf result ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) is a
// three-step butterfly (fp addition commutes, the association is the tree's).  All lanes of a group walk
// the node's sub-tree (a few leaves) in lock step.  64-thread blocks with <= 64 registers fit beside three
// resident k_sim_paths blocks (4096 registers are left per SM), so the reduction of one chunk really
// overlaps the path kernel of the next.
__device__ __forceinline__ double group8_leaf_sum(const double *__restrict__ a, int64_t n, int lane, double &mn, double &mx)
{
    const unsigned gmask = 0xffu << ((threadIdx.x & 31) & ~7);
    double res;
    int64_t i = 0;
    if (n >= 8) {
        double r = a[lane];
        mn = r < mn ? r : mn;
        mx = r > mx ? r : mx;
        for (i = 8; i < n - (n % 8); i += 8) {
            const double v = a[i + lane];
            r = r + v;
            mn = v < mn ? v : mn;
            mx = v > mx ? v : mx;
        }
        r = r + __shfl_xor_sync(gmask, r, 1);
        r = r + __shfl_xor_sync(gmask, r, 2);
        res = r + __shfl_xor_sync(gmask, r, 4);
    } else {
        res = 0.;
    }
    for (; i < n; i++) {  // n % 8 tail (only the last leaf of an array has one), or the whole n < 8 case
        const double v = a[i];
        res = res + v;
        mn = v < mn ? v : mn;
        mx = v > mx ? v : mx;
    }
    return res;
}

static __global__ void __launch_bounds__(64) k_reduce_nodes(const double *__restrict__ out, int64_t ld, int64_t n, int d, double *__restrict__ part, int64_t pstride)
{
    const int lane = threadIdx.x & 7;
    const int64_t j = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 3);
    const int arr = blockIdx.y;
    if (j >= ((int64_t)1 << d)) return;  // whole groups leave together
    int64_t start, len;
    np_node_at_depth(n, d, j, &start, &len);
    const double *a = out + (int64_t)arr * ld + start;
    double mn = a[0], mx = a[0];
    // post-order walk of the node's sub-tree (np_pairwise_raw with a cooperative leaf); depth <= 8 here
    int64_t nn[12];
    double left[12];
    int8_t st[12];
    int sp = 1;
    int64_t pos = 0;
    nn[0] = len;
    st[0] = 0;
    double v = 0.0;
    bool done = false;
    while (!done) {
        const int top = sp - 1;
        if (st[top] == 0) {
            if (nn[top] <= 128) {
                v = group8_leaf_sum(a + pos, nn[top], lane, mn, mx);
                pos += nn[top];
                sp--;
            } else {
                st[top] = 1;
                nn[sp] = np_split_hd(nn[top]);
                st[sp] = 0;
                sp++;
                continue;
            }
        }
        while (true) {
            if (sp == 0) {
                done = true;
                break;
            }
            const int par = sp - 1;
            if (st[par] == 1) {
                left[par] = v;
                st[par] = 2;
                nn[sp] = nn[par] - np_split_hd(nn[par]);
                st[sp] = 0;
                sp++;
                break;
            }
            v = left[par] + v;
            sp--;
        }
    }
    const unsigned gmask = 0xffu << ((threadIdx.x & 31) & ~7);
#pragma unroll
    for (int s = 1; s < 8; s <<= 1) {
        const double omn = __shfl_xor_sync(gmask, mn, s), omx = __shfl_xor_sync(gmask, mx, s);
        mn = omn < mn ? omn : mn;
        mx = omx > mx ? omx : mx;
    }
    if (lane != 0) return;
    part[(int64_t)arr * pstride + j] = v;
    part[(int64_t)(40 + arr) * pstride + j] = mn;
    part[(int64_t)(80 + arr) * pstride + j] = mx;
}

// grid 40 (one block per array): fold the 2^d node sums pairwise (the perfect-binary top of the
// tree), level by level, ping-ponging between `part` row `arr` and scratch row 120 + arr; min / max
// by block reduction.  The additions are exactly those of the sequential fold, only run in parallel.
// 64 threads per block so that it, too, fits beside resident k_sim_paths blocks.
static __global__ void __launch_bounds__(64) k_reduce_top(double *__restrict__ part, int64_t pstride, int d, mcts_chunk_result *__restrict__ rec, int64_t first_path, int64_t n)
{
    const int arr = blockIdx.x, tid = threadIdx.x;
    double *p = part + (int64_t)arr * pstride, *q = part + (int64_t)(120 + arr) * pstride;
    const int64_t w0 = (int64_t)1 << d;
    for (int64_t w = w0; w > 1; w >>= 1) {
        for (int64_t i = tid; i < (w >> 1); i += 64) q[i] = p[2 * i] + p[2 * i + 1];
        __syncthreads();
        double *t = p;
        p = q;
        q = t;
    }
    const double *pmn = part + (int64_t)(40 + arr) * pstride, *pmx = part + (int64_t)(80 + arr) * pstride;
    double mn = pmn[0], mx = pmx[0];
    for (int64_t i = tid; i < w0; i += 64) {
        mn = pmn[i] < mn ? pmn[i] : mn;
        mx = pmx[i] > mx ? pmx[i] : mx;
    }
    __shared__ double smn[64], smx[64];
    smn[tid] = mn;
    smx[tid] = mx;
    __syncthreads();
    for (int s = 32; s > 0; s >>= 1) {
        if (tid < s) {
            smn[tid] = smn[tid + s] < smn[tid] ? smn[tid + s] : smn[tid];
            smx[tid] = smx[tid + s] > smx[tid] ? smx[tid + s] : smx[tid];
        }
        __syncthreads();
    }
    if (tid != 0) return;
    if (arr == 0) {
        rec->first_path = first_path;
        rec->n_paths = n;
    }
    if (arr < 20) {
        rec->sum_profit[arr] = p[0];
        rec->min_profit[arr] = smn[0];
        rec->max_profit[arr] = smx[0];
    } else {
        rec->sum_dd[arr - 20] = p[0];
        rec->min_dd[arr - 20] = smn[0];
        rec->max_dd[arr - 20] = smx[0];
    }
}

// ---- roofline denominators --------------------------------------------------------------------

__device__ __forceinline__ uint64_t globaltimer_ns()
{
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// 8 independent DADD chains per thread
static __global__ void __launch_bounds__(256) k_peak_dadd(double *out, int iters, double x, unsigned long long *clk)
{
    double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const long long c0 = clock64();
    const uint64_t t0 = globaltimer_ns();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            a0 = __dadd_rn(a0, x); a1 = __dadd_rn(a1, x); a2 = __dadd_rn(a2, x); a3 = __dadd_rn(a3, x);
            a4 = __dadd_rn(a4, x); a5 = __dadd_rn(a5, x); a6 = __dadd_rn(a6, x); a7 = __dadd_rn(a7, x);
        }
    }
    const long long c1 = clock64();
    const uint64_t t1 = globaltimer_ns();
    const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 1.2345e-300) out[0] = s;
    if (threadIdx.x == 0) {
        clk[2 * blockIdx.x] = (unsigned long long)(c1 - c0);
        clk[2 * blockIdx.x + 1] = (unsigned long long)(t1 - t0);
    }
}

// 8 independent integer chains per thread, two ALU instructions (LOP3 + IADD3) per link
static __global__ void __launch_bounds__(256) k_peak_int(unsigned *out, int iters, unsigned x, unsigned y)
{
    unsigned a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            a0 = (a0 ^ x) + y; a1 = (a1 ^ x) + y; a2 = (a2 ^ x) + y; a3 = (a3 ^ x) + y;
            a4 = (a4 ^ x) + y; a5 = (a5 ^ x) + y; a6 = (a6 ^ x) + y; a7 = (a7 ^ x) + y;
        }
    }
    const unsigned s = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
    if (s == 0x12345u) out[0] = s;
}

// Register-only replica of the evaluator's trade step (3 DADD + 2 DSETP + 5 FSEL + 1 predicate test per
// strategy-step; five strategies per thread, 128-thread blocks, three blocks per SM like k_sim_paths):
// the rate this instruction mix can reach on the chip with no generator, no shared memory and no
// stores.  Measured: ~2.3e12 strategy-steps/s at any occupancy from 8 to 64 warps per SM -- DSETP and
// the ALU selects do not overlap (tools/mix_bench.cu).
static __global__ void __launch_bounds__(128, 3) k_peak_step(double *out, int steps8, double v1, double v2, unsigned seed)
{
    constexpr int NS = 5;
    double E[NS], peak[NS], dd[NS], r[NS][8];
#pragma unroll
    for (int s = 0; s < NS; s++) {
        E[s] = -0.0;
        peak[s] = -1e300;
        dd[s] = 0.0;
#pragma unroll
        for (int j = 0; j < 8; j++) r[s][j] = -0.0;
    }
    unsigned w = seed * 2654435761u + threadIdx.x * 40503u + blockIdx.x;
    for (int it = 0; it < steps8; it++) {
        w = w * 1664525u + 1013904223u;
#pragma unroll
        for (int i = 0; i < 8; i++) {
#pragma unroll
            for (int s = 0; s < NS; s++) {
                const bool p = (w >> (4 * i + (s & 3))) & 1u;
                const double c = __hiloint2double(p ? __double2hiint(v2) : __double2hiint(v1), __double2loint(v1));
                trade_step(E[s], peak[s], dd[s], r[s][i], (s & 1) ? c : -c);
            }
        }
    }
    double acc = 0;
#pragma unroll
    for (int s = 0; s < NS; s++) {
        acc += E[s] + peak[s] + dd[s];
#pragma unroll
        for (int j = 0; j < 8; j++) acc += r[s][j];
    }
    if (steps8 < 0) out[0] = acc;
}

// Register-only loop with the INSTRUCTION MIX of the shipped evaluator's trade step -- per step 5 fp64 adds (FP64 pipe), 4 integer
// multiply-adds (FMA-heavy pipe, where the predicated moves run) and 4 logic operations (ALU pipe), every instruction on its own
// dependency chain, interleaved -- at the hot kernel's occupancy (128-thread blocks, three per SM).  What the chip's schedulers can
// issue for this mix with no generator, no shared memory, no data dependence between neighbours: the practical ceiling of the issue
// roofline (tools/pipe_mix_bench.cu sweeps other mixes and occupancies; profiles/r02_pipe_mix_bench.txt).
static __global__ void __launch_bounds__(128, 3) k_peak_mix(double *out, int steps, double x, unsigned y, unsigned z)
{
    double d[5];
    unsigned f[4], a[4];
#pragma unroll
    for (int i = 0; i < 5; i++) d[i] = threadIdx.x + i;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        f[i] = threadIdx.x * 3 + i;
        a[i] = threadIdx.x * 5 + i;
    }
    for (int it = 0; it < steps; it += 4) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int i = 0; i < 5; i++) {
                asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(x));
                if (i < 4) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(f[i]) : "r"(y));
                if (i < 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x1e;" : "+r"(a[i]) : "r"(y), "r"(z));
            }
        }
    }
    double s = 0;
    unsigned t = 0;
#pragma unroll
    for (int i = 0; i < 5; i++) s += d[i];
#pragma unroll
    for (int i = 0; i < 4; i++) t ^= f[i] ^ a[i];
    if (s == 1.2345e-300 && t == 0x12345u) out[0] = s;
}

}  // namespace mcts
