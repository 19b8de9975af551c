// Do DMMA (mma.sync f64) and DFMA share the fp64 pipe on B200?  Even warps issue DMMA chains, odd warps DFMA chains.
// Prints the time of DMMA-only, DFMA-only and mixed launches (same per-warp work): if mixed ~= max the pipes are
// separate, if mixed ~= sum they are one resource.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) k(double* out, int iters, int mode, double a, double b) {
    const int warp = threadIdx.x >> 5;
    const bool do_mma = (mode == 0) || (mode == 2 && (warp & 1) == 0);
    const bool do_fma = (mode == 1) || (mode == 2 && (warp & 1) == 1);
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x + i; c[i][1] = i; }
    if (do_mma) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
    }
    if (do_fma) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 4; ++r)   // 8 DFMA warp-instr * 2 clk = same pipe time as one DMMA (16 clk)
#pragma unroll
                for (int i = 0; i < 8; ++i) { c[i][0] = fma(c[i][0], a, b); c[i][1] = fma(c[i][1], a, b); }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out; cudaMalloc(&out, sizeof(double) * sms * 4 * 256);
    const int iters = 2000;
    const char* names[3] = {"dmma_all_warps", "dfma_all_warps", "mixed_even_dmma_odd_dfma"};
    printf("{");
    for (int mode = 0; mode < 3; ++mode) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<<<sms * 4, 256>>>(out, 10, mode, 0.999, 1e-3);
        cudaDeviceSynchronize();
        float best = 1e30f;
        for (int r = 0; r < 3; ++r) {
            cudaEventRecord(e0);
            k<<<sms * 4, 256>>>(out, iters, mode, 0.999, 1e-3);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        printf("\"%s_ms\": %.3f%s", names[mode], best, mode < 2 ? ", " : "");
    }
    printf("}\n");
    return 0;
}
