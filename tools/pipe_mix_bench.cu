// pipe_mix_bench.cu -- what issue rate can an instruction MIX like the evaluator's reach on a B200 scheduler?
// Register-only loops of independent chains: per iteration ND fp64 adds (FP64 pipe), NF integer multiply-adds (FMA-heavy
// pipe) and NA logic ops (ALU pipe), written interleaved or clustered, run with W warps per scheduler.  Prints cycles per
// iteration per scheduler and the resulting IPC.  (is 0.78 IPC the mix's ceiling at 3 warps?)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/pipe_mix_bench tools/pipe_mix_bench.cu && tools/pipe_mix_bench
#include <cstdio>
#include <cuda_runtime.h>

template <int ND, int NF, int NA, bool CLUSTER>
__global__ void __launch_bounds__(128) k_mix(double *out, int iters, double x, unsigned y, unsigned z)
{
    double d[8];
    unsigned f[8], a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { d[i] = threadIdx.x + i; f[i] = threadIdx.x * 3 + i; a[i] = threadIdx.x * 5 + i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (CLUSTER) {
#pragma unroll
                for (int i = 0; i < ND; i++) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(x));
#pragma unroll
                for (int i = 0; i < NF; i++) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(f[i]) : "r"(y));
#pragma unroll
                for (int i = 0; i < NA; i++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x1e;" : "+r"(a[i]) : "r"(y), "r"(z));
            } else {
                constexpr int N = ND > NF ? (ND > NA ? ND : NA) : (NF > NA ? NF : NA);
#pragma unroll
                for (int i = 0; i < N; i++) {
                    if (i < ND) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(x));
                    if (i < NF) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(f[i]) : "r"(y));
                    if (i < NA) asm volatile("lop3.b32 %0, %0, %1, %2, 0x1e;" : "+r"(a[i]) : "r"(y), "r"(z));
                }
            }
        }
    }
    double s = 0;
    unsigned t = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { s += d[i]; t ^= f[i] ^ a[i]; }
    if (s == 1.2345e-300 && t == 0x12345u) out[0] = s;
}

template <int ND, int NF, int NA, bool CLUSTER>
void run(const char *name, int blocks_per_sm)
{
    int dev = 0, sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    double *out;
    cudaMalloc(&out, 64);
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms = 0;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        k_mix<ND, NF, NA, CLUSTER><<<sms * blocks_per_sm, 128>>>(out, iters, 1.0000001, 0x9e3779b9u, 0x7f4a7c15u);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    const double cycles = ms * 1e-3 * khz * 1e3;
    const double per_iter = cycles / (iters * 4.0 * blocks_per_sm);   // one warp of every block per scheduler
    const int n = ND + NF + NA;
    printf("%-34s %d warps/scheduler: %6.2f cycles per %2d-instruction group = IPC %.3f  (FP64 share %.0f%%)\n", name, blocks_per_sm, per_iter, n, n / per_iter,
           100.0 * ND / n);
    cudaFree(out);
}

int main()
{
    for (int w : {1, 2, 3, 4}) {
        run<5, 4, 3, false>("5 D + 4 F + 3 A interleaved", w);
        run<5, 4, 3, true>("5 D + 4 F + 3 A clustered", w);
        run<4, 4, 4, false>("4 D + 4 F + 4 A interleaved", w);
        run<5, 4, 5, false>("5 D + 4 F + 5 A interleaved", w);
        run<5, 1, 0, false>("5 D + 1 F", w);
        run<1, 4, 4, false>("1 D + 4 F + 4 A", w);
        run<5, 4, 4, false>("5 D + 4 F + 4 A interleaved (13.x-instruction step of the kernel)", w);
        run<6, 3, 3, false>("6 D + 3 F + 3 A interleaved", w);
        run<3, 4, 5, false>("3 D + 4 F + 5 A interleaved", w);
    }
    return 0;
}
