// dsetp_bench.cu -- issue rate of DSETP (fp64 compare -> predicate) vs DADD on sm_100a.
// The DSETP loop is written in PTX so that the compares accumulate into predicates directly
// (setp.gt.or.f64 p, x, y, p -> DSETP.GT.OR) and no ALU instruction consumes them.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) k_dsetp(int *out, int iters, double y0)
{
    double x0 = threadIdx.x * 0.37, x1 = x0 + 11.5, x2 = x0 + 23.0, x3 = x0 + 34.5, x4 = x0 + 46.0, x5 = x0 + 57.5, x6 = x0 + 69.0, x7 = x0 + 80.5;
    int res;
    asm volatile(
        "{\n\t"
        ".reg .pred p0, p1, p2, p3, q;\n\t"
        ".reg .f64 y;\n\t"
        ".reg .s32 i, r;\n\t"
        "mov.f64 y, %10;\n\t"
        "mov.s32 i, %9;\n\t"
        "setp.ne.s32 p0, i, i;\n\t"
        "setp.ne.s32 p1, i, i;\n\t"
        "setp.ne.s32 p2, i, i;\n\t"
        "setp.ne.s32 p3, i, i;\n\t"
        "LOOP:\n\t"
        "add.rn.f64 y, y, 0d3FF0000000000000;\n\t"
        "setp.gt.or.f64 p0, %1, y, p0;\n\t"
        "setp.gt.or.f64 p1, %2, y, p1;\n\t"
        "setp.gt.or.f64 p2, %3, y, p2;\n\t"
        "setp.gt.or.f64 p3, %4, y, p3;\n\t"
        "setp.lt.or.f64 p0, %5, y, p0;\n\t"
        "setp.lt.or.f64 p1, %6, y, p1;\n\t"
        "setp.lt.or.f64 p2, %7, y, p2;\n\t"
        "setp.lt.or.f64 p3, %8, y, p3;\n\t"
        "setp.gt.or.f64 p0, %2, y, p0;\n\t"
        "setp.gt.or.f64 p1, %3, y, p1;\n\t"
        "setp.gt.or.f64 p2, %4, y, p2;\n\t"
        "setp.gt.or.f64 p3, %5, y, p3;\n\t"
        "setp.lt.or.f64 p0, %6, y, p0;\n\t"
        "setp.lt.or.f64 p1, %7, y, p1;\n\t"
        "setp.lt.or.f64 p2, %8, y, p2;\n\t"
        "setp.lt.or.f64 p3, %1, y, p3;\n\t"
        "sub.s32 i, i, 1;\n\t"
        "setp.gt.s32 q, i, 0;\n\t"
        "@q bra LOOP;\n\t"
        "selp.s32 %0, 1, 0, p0;\n\t"
        "selp.s32 r, 2, 0, p1;\n\t"
        "add.s32 %0, %0, r;\n\t"
        "selp.s32 r, 4, 0, p2;\n\t"
        "add.s32 %0, %0, r;\n\t"
        "selp.s32 r, 8, 0, p3;\n\t"
        "add.s32 %0, %0, r;\n\t"
        "}\n"
        : "=r"(res)
        : "d"(x0), "d"(x1), "d"(x2), "d"(x3), "d"(x4), "d"(x5), "d"(x6), "d"(x7), "r"(iters), "d"(y0));
    if (iters < 0) out[0] = res;
}
__global__ void __launch_bounds__(256) k_dadd(double *out, int iters, double y0)
{
    double x[8];
#pragma unroll
    for (int j = 0; j < 8; j++) x[j] = threadIdx.x * 0.37 + j * 11.5;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 2; u++) {
#pragma unroll
            for (int j = 0; j < 8; j++) x[j] = x[j] + y0;
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s += x[j];
    if (iters < 0) out[0] = s;
}
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    void *d; cudaMalloc(&d, 64);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int blocks = p.multiProcessorCount * 8, iters = 8192;
    float ms;
    const double n = (double)blocks * 256 * iters * 16;
    const double slots = p.multiProcessorCount * 4 * 32 * 1.965e9;
    for (int rep = 0; rep < 3; rep++) { cudaEventRecord(a); k_dsetp<<<blocks, 256>>>((int *)d, iters, 1.0); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b); }
    printf("DSETP: %.3e thread-inst/s -> %.2f cycles per warp-instruction per SMSP\n", n / (ms * 1e-3), slots / (n / (ms * 1e-3)));
    for (int rep = 0; rep < 3; rep++) { cudaEventRecord(a); k_dadd<<<blocks, 256>>>((double *)d, iters, 1.0000001); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b); }
    printf("DADD : %.3e thread-inst/s -> %.2f cycles per warp-instruction per SMSP\n", n / (ms * 1e-3), slots / (n / (ms * 1e-3)));
    return 0;
}
