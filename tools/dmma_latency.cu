// DMMA issue/latency probe (sm_100a): throughput of mma.sync.m8n8k4.f64 as a function of the number of warps per
// SM sub-partition and of independent accumulator chains per warp.  One CTA per SM (grid = #SMs).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/dmma_latency tools/dmma_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void k(double* out, int iters, double a, double b) {
    double c[CHAINS][2];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS>
void run(int sms, int threads, double* out) {
    const int iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<CHAINS><<<sms, threads>>>(out, 64, 0.999, 1e-3);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        k<CHAINS><<<sms, threads>>>(out, iters, 0.999, 1e-3);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    int prop_clk = 0; cudaDeviceGetAttribute(&prop_clk, cudaDevAttrClockRate, 0);
    double clk = best * 1e-3 * prop_clk * 1e3;              // SM cycles at the nominal max clock
    double per_smsp = (double)(threads / 32) / 4.0 * iters * CHAINS;  // DMMAs issued per sub-partition
    printf("{\"warps_per_smsp\": %.2f, \"chains\": %d, \"cycles_per_dmma_per_smsp\": %.2f, \"tflops\": %.2f}\n",
           threads / 128.0, CHAINS, clk / per_smsp, (double)sms * (threads / 32) * iters * CHAINS * 512.0 / (best * 1e-3) * 1e-12);
}

int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out; cudaMalloc(&out, sizeof(double) * sms * 1024);
    for (int threads : {128, 256, 512, 1024}) {
        run<1>(sms, threads, out); run<2>(sms, threads, out); run<4>(sms, threads, out);
        run<8>(sms, threads, out); run<16>(sms, threads, out);
    }
    return 0;
}
