// Stand-alone timing harness of the hot kernel for ONE candidate set of pass lists (-include cand.h defines MCTS_FP0..3):
// base generation + k_sim_paths<16, PM_FULL, 128> on the bench workload (C5 parameters, 2,097,140 paths), best of 4 launches.
// Used by tools/autotune/run.py to choose the pass lists by MEASURED time instead of SASS length.
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false -I mc_trade_sizing_b200/csrc -include cand.h -o cand tools/autotune/harness.cu
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "aux_kernels.cuh"
#include "kernels.cuh"
#include "planner.hpp"

using namespace mcts;

#ifndef HARNESS_BLOCK
#define HARNESS_BLOCK 128   // threads per block (384 per SM in all)
#endif
#ifndef HARNESS_K
#define HARNESS_K 16   // 16: Fisher-Yates generator; 5: five-value urn (the C5 parameters have five distinct values)
#endif

#define CK(x)                                                                  \
    do {                                                                       \
        cudaError_t e = (x);                                                   \
        if (e != cudaSuccess) {                                                \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            return 1;                                                          \
        }                                                                      \
    } while (0)

int main(int argc, char **argv)
{
    mcts_params p;
    memset(&p, 0, sizeof p);
    p.mode = MCTS_MODE_DYNAMIC;
    p.num_trades = 1000;
    p.num_simulations = 104857;
    p.num_mc_shuffles = 20;
    p.hit_rate = 0.82;
    p.avg_win = 186.0;
    p.avg_loss = 219.0;
    p.n_regimes = -1;
    GenSpec g;
    std::string err;
    if (build_spec(p, &g, &err)) { printf("spec: %s\n", err.c_str()); return 1; }
    LeafSched sch;
    if (build_leaf_sched(g.L, &sch, &err)) { printf("sched: %s\n", err.c_str()); return 1; }
    const int64_t n = (int64_t)p.num_simulations * p.num_mc_shuffles, ld = (n + 7) / 8 * 8;
    uint32_t *d_counts;
    LeafSched *d_sched;
    double *d_out;
    CK(cudaMalloc(&d_counts, sizeof(uint32_t) * kMaxCat * p.num_simulations));
    CK(cudaMalloc(&d_sched, sizeof sch));
    CK(cudaMalloc(&d_out, sizeof(double) * 40 * ld));
    CK(cudaMemcpy(d_sched, &sch, sizeof sch, cudaMemcpyHostToDevice));
    const uint64_t seed = 12345;
    k_gen_base<<<(p.num_simulations + 127) / 128, 128>>>(g, seed, 0, p.num_simulations, d_counts);
    CK(cudaGetLastError());
    PathArgs a;
    memset(&a, 0, sizeof a);
    a.g = g;
    a.seed = seed;
    a.n = n;
    a.n_sh = p.num_mc_shuffles;
    a.counts = d_counts;
    a.sched = d_sched;
    a.out = d_out;
    a.ld = ld;
    const size_t smem = (size_t)kTabWords * 4 + (size_t)((g.L + 7) / 8) * HARNESS_BLOCK * 4;
    CK(cudaFuncSetAttribute(k_sim_paths<HARNESS_K, PM_FULL, HARNESS_BLOCK, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        a.seed = seed + rep;
        CK(cudaEventRecord(e0));
        k_sim_paths<HARNESS_K, PM_FULL, HARNESS_BLOCK, false><<<(unsigned)((n + HARNESS_BLOCK - 1) / HARNESS_BLOCK), HARNESS_BLOCK, smem>>>(a);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    std::vector<double> h(8);
    CK(cudaMemcpy(h.data(), d_out, sizeof(double) * 8, cudaMemcpyDeviceToHost));
    printf("%s %.4f ms  %.4e trade-steps/s  out[0..1] = %.3f %.3f\n", argc > 1 ? argv[1] : "cand", best, n * 1000.0 * 20.0 / (best * 1e-3), h[0], h[1]);
    return 0;
}
