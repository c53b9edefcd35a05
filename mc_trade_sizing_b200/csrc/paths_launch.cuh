// paths_launch.cuh -- explicit instantiations of k_sim_paths for ONE PathMode; included by the
// paths_*.cu translation units so that the big unrolled kernels compile in parallel.
#pragma once
#include "kernels.cuh"

namespace mcts {

inline size_t path_smem_bytes(int L, int block) { return (size_t)kTabWords * 4 + (size_t)((L + 7) / 8) * block * 4; }

template <int K, int MODE, int BLOCK>
static cudaError_t launch_kb(const PathArgs &a, cudaStream_t stream)
{
    const size_t smem = path_smem_bytes(a.g.L, BLOCK);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_sim_paths<K, MODE, BLOCK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
    }
    const unsigned grid = (unsigned)((a.n + BLOCK - 1) / BLOCK);
    k_sim_paths<K, MODE, BLOCK><<<grid, BLOCK, smem, stream>>>(a);
    return cudaGetLastError();
}

template <int K, int MODE>
static cudaError_t launch_k(const PathArgs &a, int block, cudaStream_t stream)
{
    switch (block) {
        case 128: return launch_kb<K, MODE, 128>(a, stream);
        case 64: return launch_kb<K, MODE, 64>(a, stream);
        case 32: return launch_kb<K, MODE, 32>(a, stream);
#ifdef MCTS_BLOCK192   // experiment: two 192-thread blocks per SM (mcts_options.block_threads = 192)
        case 192: return launch_kb<K, MODE, 192>(a, stream);
#endif
    }
    return cudaErrorInvalidValue;
}

template <int K, int MODE>
static cudaError_t launch_global(const PathArgs &a, cudaStream_t stream)
{
    const unsigned grid = (unsigned)((a.n + 127) / 128);
    k_sim_paths<K, MODE, 128, true><<<grid, 128, (size_t)kTabWords * 4, stream>>>(a);
    return cudaGetLastError();
}

// block == 0: code columns in global memory (a.scratch)
template <int MODE>
static cudaError_t launch_mode(const PathArgs &a, int block, cudaStream_t stream)
{
    const int uw = urn_width(a.g.K);  // 2 / 5 / 6: register urn of that many values; 0: Fisher-Yates on the code column
    if (block == 0) return uw == 2 ? launch_global<2, MODE>(a, stream) : uw == 5 ? launch_global<5, MODE>(a, stream) : uw == 6 ? launch_global<6, MODE>(a, stream) : launch_global<16, MODE>(a, stream);
    switch (uw) {
        case 2: return launch_k<2, MODE>(a, block, stream);
        case 5: return launch_k<5, MODE>(a, block, stream);
        case 6: return launch_k<6, MODE>(a, block, stream);
    }
    return launch_k<16, MODE>(a, block, stream);
}

}  // namespace mcts

namespace mcts {
template <int K, int MODE, int BLOCK>
static cudaError_t launch_int_kb(const PathArgs &a, cudaStream_t stream)
{
    const size_t smem = path_smem_bytes(a.g.L, BLOCK);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_sim_paths_int<K, MODE, BLOCK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
    }
    const unsigned grid = (unsigned)((a.n + BLOCK - 1) / BLOCK);
    k_sim_paths_int<K, MODE, BLOCK><<<grid, BLOCK, smem, stream>>>(a);
    return cudaGetLastError();
}
template <int K, int MODE>
static cudaError_t launch_int_k(const PathArgs &a, int block, cudaStream_t stream)
{
    switch (block) {
        case 128: return launch_int_kb<K, MODE, 128>(a, stream);
        case 64: return launch_int_kb<K, MODE, 64>(a, stream);
        case 32: return launch_int_kb<K, MODE, 32>(a, stream);
    }
    return cudaErrorInvalidValue;
}
template <int MODE>
static cudaError_t launch_int_mode(const PathArgs &a, int block, cudaStream_t stream)
{
    switch (urn_width(a.g.K)) {
        case 2: return launch_int_k<2, MODE>(a, block, stream);
        case 5: return launch_int_k<5, MODE>(a, block, stream);
        case 6: return launch_int_k<6, MODE>(a, block, stream);
    }
    return launch_int_k<16, MODE>(a, block, stream);
}
}  // namespace mcts

#define MCTS_DEFINE_INT_LAUNCHER(NAME, MODE) \
    cudaError_t NAME(const mcts::PathArgs &a, int block, cudaStream_t stream) { return mcts::launch_int_mode<MODE>(a, block, stream); }

#define MCTS_DEFINE_PATH_LAUNCHER(NAME, MODE) \
    cudaError_t NAME(const mcts::PathArgs &a, int block, cudaStream_t stream) { return mcts::launch_mode<MODE>(a, block, stream); }
