// mix_bench2.cu -- where can the two 64-bit conditional moves of a trade step (peak, drawdown) execute?
// Register-only replica of the evaluator's step like mix_bench.cu, with the moves written as
//   V=0   select pairs (what nvcc emits: FSEL x2 each, ALU pipe)                      <- the old kernel
//   V=5/6/7  predicated add.f64 (ptxas lowers it to DADD + FSEL x2: more work, slower)
//   V=8   sign mask + LOP3 (peak) ; V=9 integer sign test of u (peak)
//   V=10-14  unpredicated integer arithmetic on a 0/1 mask, m*a + (1-m)*b (5 IMAD per move: slower)
//   V=15/16/17  predicated IMAD with opaque multiplier 1 / addend 0 (both / peak only / drawdown only)  <- the shipped kernel
//   V=18/19  predicated mad.wide (ptxas splits it into IMAD.WIDE + IADD3 + IMAD.X + SEL)
// Results on the B200 are in profiles/r01_mix_bench2.txt; the replica itself is dispatch-bound (it carries ~3.5 register
// moves per step that the real kernel does not have), so the decisive measurement was the real kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/mix_bench2 tools/mix_bench2.cu && tools/mix_bench2
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void pmov_gt0(double &dst, double u, double src, double nz) {
    asm("{ .reg .pred p; setp.gt.f64 p, %1, 0d0000000000000000; @p add.rn.f64 %0, %2, %3; }" : "+d"(dst) : "d"(u), "d"(src), "d"(nz));
}
__device__ __forceinline__ void pmov_lt(double &dst, double u, double nz) {
    asm("{ .reg .pred p; setp.lt.f64 p, %1, %0; @p add.rn.f64 %0, %1, %2; }" : "+d"(dst) : "d"(u), "d"(nz));
}

__device__ __forceinline__ unsigned opaque_sign(double u) {
    unsigned m1; asm("shr.u32 %0, %1, 31;" : "=r"(m1) : "r"(__double2hiint(u))); return m1;
}
__device__ __forceinline__ unsigned madlo(unsigned a, unsigned b, unsigned c) {
    unsigned d; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}
__device__ __forceinline__ double sel_fma(unsigned m1, double keep, double take) {
    // m1 == 1 -> keep, 0 -> take; all on the FMA-heavy pipe
    const unsigned n1 = madlo(m1, 0xffffffffu, 1u);
    const unsigned kh = __double2hiint(keep), kl = __double2loint(keep);
    const unsigned th = __double2hiint(take), tl = __double2loint(take);
    return __hiloint2double(madlo(m1, kh, madlo(n1, th, 0u)), madlo(m1, kl, madlo(n1, tl, 0u)));
}
__device__ __forceinline__ void pmov64_gt0(double &dst, double u, double src, unsigned one, unsigned zero) {
    asm("{ .reg .pred p; .reg .b32 lo, hi, slo, shi;\n\t"
        "setp.gt.f64 p, %1, 0d0000000000000000;\n\t"
        "mov.b64 {lo,hi}, %0;\n\tmov.b64 {slo,shi}, %2;\n\t"
        "@p mad.lo.u32 lo, slo, %3, %4;\n\t@p mad.lo.u32 hi, shi, %3, %4;\n\t"
        "mov.b64 %0, {lo,hi}; }"
        : "+d"(dst) : "d"(u), "d"(src), "r"(one), "r"(zero));
}
__device__ __forceinline__ void pmov64_lt(double &dst, double u, unsigned one, unsigned zero) {
    asm("{ .reg .pred p; .reg .b32 lo, hi, slo, shi;\n\t"
        "setp.lt.f64 p, %1, %0;\n\t"
        "mov.b64 {lo,hi}, %0;\n\tmov.b64 {slo,shi}, %1;\n\t"
        "@p mad.lo.u32 lo, slo, %2, %3;\n\t@p mad.lo.u32 hi, shi, %2, %3;\n\t"
        "mov.b64 %0, {lo,hi}; }"
        : "+d"(dst) : "d"(u), "r"(one), "r"(zero));
}
__device__ __forceinline__ void wmov64_gt0(double &dst, double u, double src, unsigned zero) {
    asm("{ .reg .pred p; .reg .b64 d, s;\n\t"
        "setp.gt.f64 p, %1, 0d0000000000000000;\n\t"
        "mov.b64 d, %0;\n\tmov.b64 s, %2;\n\t"
        "@p mad.wide.u32 d, %3, %3, s;\n\t"
        "mov.b64 %0, d; }"
        : "+d"(dst) : "d"(u), "d"(src), "r"(zero));
}
__device__ __forceinline__ void wmov64_lt(double &dst, double u, unsigned zero) {
    asm("{ .reg .pred p; .reg .b64 d, s;\n\t"
        "setp.lt.f64 p, %1, %0;\n\t"
        "mov.b64 d, %0;\n\tmov.b64 s, %1;\n\t"
        "@p mad.wide.u32 d, %2, %2, s;\n\t"
        "mov.b64 %0, d; }"
        : "+d"(dst) : "d"(u), "r"(zero));
}
template <int NS, int V>
__global__ void __launch_bounds__(128) k_mix(double *out, int steps8, double v1, double v2, unsigned seed, double nz, unsigned one, unsigned zero)
{
    double E[NS], peak[NS], dd[NS], r[NS][8];
#pragma unroll
    for (int s = 0; s < NS; s++) {
        E[s] = -0.0; peak[s] = -1e300; dd[s] = 0.0;
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
                const double c = __hiloint2double(p ? __double2hiint(v2) : __double2hiint(v1), __double2loint(v1)) ;
                E[s] = E[s] + (s & 1 ? c : -c);
                r[s][i] = r[s][i] + c;
                const double u = E[s] - peak[s];
                if (V == 0) {
                    peak[s] = (u > 0.0) ? E[s] : peak[s];
                    dd[s] = (u < dd[s]) ? u : dd[s];
                } else if (V == 5) {
                    peak[s] = (u > 0.0) ? E[s] : peak[s];
                    pmov_lt(dd[s], u, nz);
                } else if (V == 6) {
                    pmov_gt0(peak[s], u, E[s], nz);
                    dd[s] = (u < dd[s]) ? u : dd[s];
                } else if (V == 7) {
                    pmov_gt0(peak[s], u, E[s], nz);
                    pmov_lt(dd[s], u, nz);
                } else if (V == 8) {
                    const int m = __double2hiint(u) >> 31;  // all ones when u < 0
                    const int eh = __double2hiint(E[s]), el = __double2loint(E[s]);
                    const int ph = __double2hiint(peak[s]), pl = __double2loint(peak[s]);
                    peak[s] = __hiloint2double((ph & m) | (eh & ~m), (pl & m) | (el & ~m));
                    pmov_lt(dd[s], u, nz);
                } else if (V == 15) {
                    pmov64_gt0(peak[s], u, E[s], one, zero);
                    pmov64_lt(dd[s], u, one, zero);
                } else if (V == 16) {
                    pmov64_gt0(peak[s], u, E[s], one, zero);
                    dd[s] = (u < dd[s]) ? u : dd[s];
                } else if (V == 17) {
                    peak[s] = (u > 0.0) ? E[s] : peak[s];
                    pmov64_lt(dd[s], u, one, zero);
                } else if (V == 18) {
                    wmov64_gt0(peak[s], u, E[s], zero);
                    wmov64_lt(dd[s], u, zero);
                } else if (V == 19) {
                    wmov64_gt0(peak[s], u, E[s], zero);
                    dd[s] = (u < dd[s]) ? u : dd[s];
                } else if (V == 13) {
                    peak[s] = sel_fma(opaque_sign(u), peak[s], E[s]);
                    dd[s] = (u < dd[s]) ? u : dd[s];
                } else if (V == 14) {
                    if (s & 1) peak[s] = (u > 0.0) ? E[s] : peak[s];
                    else peak[s] = sel_fma(opaque_sign(u), peak[s], E[s]);
                    dd[s] = (u < dd[s]) ? u : dd[s];
                } else if (V == 10) {  // peak select on the FMA pipe: 0/1 multipliers
                    const unsigned m1 = (unsigned)__double2hiint(u) >> 31;  // 1 when u < 0: keep peak
                    const unsigned n1 = m1 * 0xffffffffu + 1u;
                    const unsigned eh = __double2hiint(E[s]), el = __double2loint(E[s]);
                    const unsigned ph = __double2hiint(peak[s]), pl = __double2loint(peak[s]);
                    peak[s] = __hiloint2double(m1 * ph + n1 * eh, m1 * pl + n1 * el);
                    dd[s] = (u < dd[s]) ? u : dd[s];
                } else if (V == 11) {  // hi word by LOP3 mask, lo word on the FMA pipe
                    const int m = __double2hiint(u) >> 31;
                    const unsigned m1 = (unsigned)m * 0xffffffffu, n1 = (unsigned)m + 1u;
                    const unsigned eh = __double2hiint(E[s]), el = __double2loint(E[s]);
                    const unsigned ph = __double2hiint(peak[s]), pl = __double2loint(peak[s]);
                    peak[s] = __hiloint2double((ph & m) | (eh & ~m), m1 * pl + n1 * el);
                    dd[s] = (u < dd[s]) ? u : dd[s];
                } else if (V == 12) {  // V10 for even slots, V0 for odd
                    if (s & 1) {
                        peak[s] = (u > 0.0) ? E[s] : peak[s];
                    } else {
                        const unsigned m1 = (unsigned)__double2hiint(u) >> 31;
                        const unsigned n1 = m1 * 0xffffffffu + 1u;
                        const unsigned eh = __double2hiint(E[s]), el = __double2loint(E[s]);
                        const unsigned ph = __double2hiint(peak[s]), pl = __double2loint(peak[s]);
                        peak[s] = __hiloint2double(m1 * ph + n1 * eh, m1 * pl + n1 * el);
                    }
                    dd[s] = (u < dd[s]) ? u : dd[s];
                } else if (V == 9) {
                    peak[s] = (__double2hiint(u) > 0) ? E[s] : peak[s];
                    pmov_lt(dd[s], u, nz);
                }
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
    if (acc == 1.2345e-300) out[0] = acc;
}

template <int NS, int V>
void run(int blocks_per_sm, int sms)
{
    double *d;
    cudaMalloc(&d, 64);
    const int steps8 = 2000;
    const size_t smem = (size_t)(227 * 1024 / blocks_per_sm) - 1024;
    cudaFuncSetAttribute(k_mix<NS, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_mix<NS, V>, 128, smem);
    const int grid = sms * occ * 4;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float ms = 0;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(a);
        k_mix<NS, V><<<grid, 128, smem>>>(d, steps8, 186.0, 372.0, rep, -0.0, 1u, 0u);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b);
    }
    const double steps = (double)grid * 128 * steps8 * 8 * NS;
    const double rate = steps / (ms * 1e-3);
    const double cyc = 32.0 * (sms * 4 * 1.965e9) / rate;
    printf("V=%d NS=%d warps/SM=%2d: %.3e strategy-steps/s, %.2f cycles per warp-step\n", V, NS, occ * 4, rate, cyc);
    cudaFree(d);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    run<5, 0>(3, sms);
    run<5, 5>(3, sms);
    run<5, 6>(3, sms);
    run<5, 7>(3, sms);
    run<5, 8>(3, sms);
    run<5, 9>(3, sms);
    run<5, 10>(3, sms);
    run<5, 13>(3, sms);
    run<5, 15>(3, sms);
    run<5, 16>(3, sms);
    run<5, 17>(3, sms);
    run<5, 18>(3, sms);
    run<5, 15>(2, sms);
    run<5, 15>(4, sms);
    run<5, 0>(4, sms);
    return 0;
}
