// How do DMMA streams and DEPENDENT DFMA chains share the fp64 pipe of one SM sub-partition (B200)?
// Per sub-partition: WM warps issue DMMAs on 13 independent accumulators (the transform's consumer), WF warps issue DFMAs
// on CH independent chains (the Box-Muller producer: CH = 1..2 dependent chains per warp in k_xform_ws).  Fixed work per
// warp: `rounds` x (182 DMMA) or `rounds` x (598 DFMA); prints cycles per round per sub-partition for each mix.
// Ideal if the pipe time simply adds: WM * 182 * 16 + WF * 598 * 2 cycles per round.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/fp64_mix tools/fp64_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CH>
__global__ void __launch_bounds__(512) k(double* out, int rounds, int wm, int wf, double a, double b) {
    const int warp = threadIdx.x >> 5;
    const int slot = warp >> 2;  // warps w, w+4, w+8, ... share a sub-partition: slot = index inside the sub-partition
    double s = 0;
    if (slot < wm) {
        double c[13][2];
#pragma unroll
        for (int i = 0; i < 13; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
        for (int r = 0; r < rounds; ++r) {
#pragma unroll
            for (int j = 0; j < 14; ++j)
#pragma unroll
                for (int i = 0; i < 13; ++i)
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                 : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
#pragma unroll
        for (int i = 0; i < 13; ++i) s += c[i][0] + c[i][1];
    } else if (slot < wm + wf) {
        double c[CH];
#pragma unroll
        for (int i = 0; i < CH; ++i) c[i] = threadIdx.x * 1e-3 + i;
        for (int r = 0; r < rounds; ++r) {
            for (int j = 0; j < 598 / CH; ++j)
#pragma unroll
                for (int i = 0; i < CH; ++i) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(c[i]) : "d"(a), "d"(b));
        }
#pragma unroll
        for (int i = 0; i < CH; ++i) s += c[i];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH>
static void run(int sms, int wm, int wf, double* out) {
    const int rounds = 400;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<CH><<<sms, 512>>>(out, 4, wm, wf, 0.999, 1e-3);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        k<CH><<<sms, 512>>>(out, rounds, wm, wf, 0.999, 1e-3);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double cyc = best * 1e-3 * khz * 1e3 / rounds;
    printf("{\"dmma_warps\": %d, \"dfma_warps\": %d, \"chains\": %d, \"cycles_per_round\": %.0f, \"ideal\": %d}\n", wm, wf, CH, cyc,
           wm * 182 * 16 + wf * (598 / CH) * CH * 2);
}

int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out; cudaMalloc(&out, sizeof(double) * sms * 512);
    run<1>(sms, 2, 0, out); run<1>(sms, 1, 0, out);
    run<1>(sms, 0, 2, out); run<2>(sms, 0, 2, out); run<4>(sms, 0, 2, out); run<8>(sms, 0, 2, out);
    run<1>(sms, 2, 2, out); run<2>(sms, 2, 2, out); run<4>(sms, 2, 2, out); run<8>(sms, 2, 2, out); run<13>(sms, 2, 2, out);
    run<2>(sms, 2, 1, out); run<2>(sms, 1, 1, out); run<2>(sms, 3, 1, out); run<4>(sms, 3, 1, out);
    return 0;
}
