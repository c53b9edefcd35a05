// One evaluation pass of the hot kernel alone (Fisher-Yates generator + the ids of MCTS_ONE_PASS with separate
// chains), so that a candidate pass compiles in ~2 s.  See search_shared.py.
#include "kernels.cuh"
namespace mcts {
using OnePass = IdList<MCTS_ONE_PASS>;
__global__ void __launch_bounds__(128, 3) k_one_pass(const __grid_constant__ PathArgs a)
{
    extern __shared__ __align__(16) uint32_t smem[];
    TabSoA *tab = reinterpret_cast<TabSoA *>(smem);
    const int tid = threadIdx.x;
    uint32_t *codes = smem + kTabWords;
    if (tid < 16) {
        tab->v1[tid] = a.g.tab[tid].v1;
        tab->v2[tid] = a.g.tab[tid].v2;
        tab->v3[tid] = a.g.tab[tid].v3;
    }
    __syncthreads();
    const int64_t local = (int64_t)blockIdx.x * 128 + tid;
    const bool active = local < a.n;
    if (active) generate_codes<16>(a, local, codes + tid, 128);
    CodeSourceSoA src;
    src.codes = codes + tid;
    src.stride = 128;
    src.tab = tab;
    src.word = 0;
    __syncthreads();
    if (active) run_pass_k<OnePass, PM_FULL, MCTS_SEL_FULL>(src, a, local);
}
}  // namespace mcts
void launch_one(const mcts::PathArgs &a) { mcts::k_one_pass<<<1, 128>>>(a); }
