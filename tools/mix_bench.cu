// mix_bench.cu -- what issue rate can sm_100a sustain on the evaluator's instruction mix?
// Register-only replica of trade_step (3 DADD + 2 DSETP + 6 FSEL per strategy-step, select driven by a
// bit stream), NS strategies per thread, W warps per SM.  Prints trade-steps/s and the implied IPC.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o mix_bench tools/mix_bench.cu && ./mix_bench
#include <cstdio>
#include <cuda_runtime.h>

template <int NS, int V>
__global__ void __launch_bounds__(128) k_mix(double *out, int steps8, double v1, double v2, unsigned seed)
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
                if (V == 0) {  // the evaluator's step
                    peak[s] = (u > 0.0) ? E[s] : peak[s];
                    dd[s] = (u < dd[s]) ? u : dd[s];
                } else if (V == 1) {  // no drawdown min: -1 DSETP, -2 FSEL
                    peak[s] = (u > 0.0) ? E[s] : peak[s];
                } else if (V == 2) {  // adds only (+ the contribution select): 3 DADD + 1 FSEL
                    dd[s] = u;
                } else if (V == 3) {  // sign test of u on the ALU pipe instead of DSETP
                    peak[s] = (__double2hiint(u) > 0) ? E[s] : peak[s];
                    dd[s] = (u < dd[s]) ? u : dd[s];
                } else if (V == 4) {  // selects only: no FP64 compare at all (predicates from the bit stream)
                    const bool q = (w >> (4 * i + ((s + 1) & 3))) & 1u;
                    peak[s] = p ? E[s] : peak[s];
                    dd[s] = q ? u : dd[s];
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
    // limit residency with dynamic shared memory: 227 KB / blocks_per_sm each
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
        k_mix<NS, V><<<grid, 128, smem>>>(d, steps8, 186.0, 372.0, rep);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b);
    }
    const double steps = (double)grid * 128 * steps8 * 8 * NS;
    const double rate = steps / (ms * 1e-3);
    // 11 issue slots per strategy-step; IPC per SMSP at 1.965 GHz
    const double ipc = rate * 11.0 / 32.0 / (sms * 4 * 1.965e9);
    printf("V=%d NS=%d warps/SM=%2d (occupancy %d blocks): %.3e strategy-steps/s, IPC(11 instr/step) = %.2f\n", V, NS, occ * 4, occ, rate, ipc);
    cudaFree(d);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    for (int b : {1, 2, 3}) run<5, 0>(b, sms);
    for (int b : {3, 8}) run<2, 0>(b, sms);
    run<5, 1>(3, sms); run<2, 1>(8, sms);
    run<5, 2>(3, sms); run<2, 2>(8, sms);
    run<5, 3>(3, sms); run<2, 3>(8, sms);
    run<5, 4>(3, sms); run<2, 4>(8, sms);
    return 0;
}
