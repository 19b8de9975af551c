// Sustained fp64 tensor (DMMA m8n8k4) throughput on B200 with *varying* operand data, and the SM clock the GPU
// actually holds while doing it (clock64 ticks / event time).  The burst probe tools/fp64_peak.cu runs ~3 ms with
// constant operands; a CBREE sweep runs DMMAs on real data for tens of milliseconds back to back, which is the
// regime the power cap governs (MEASURED_PEAKS.json shows the same burst / sustained split for bf16).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_sustained tools/fp64_sustained.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <bool DATA>
__global__ void __launch_bounds__(256) k_dmma(double* out, long long* ticks, int iters, double a0, double b0) {
    double c[8][2], a[8], b[8];
    unsigned s = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        c[i][0] = threadIdx.x;
        c[i][1] = i;
        s = s * 1664525u + 1013904223u;
        a[i] = DATA ? ((double)(int)s) * 4.6566e-10 * 1.000000123 : a0;
        s = s * 1664525u + 1013904223u;
        b[i] = DATA ? ((double)(int)s) * 4.6566e-10 * 0.999999871 : b0;
    }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a[(i + it) & 7]), "d"(b[(i + 3 * it) & 7]));
    }
    long long t1 = clock64();
    double sum = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) sum += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
    if (threadIdx.x == 0 && blockIdx.x == 0) ticks[0] = t1 - t0;
}

int main() {
    int sms = 0;
    CHECK(cudaSetDevice(0));
    CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int grid = sms * 8, iters = 20000;
    double* out;
    long long* ticks;
    CHECK(cudaMalloc(&out, sizeof(double) * grid * 256));
    CHECK(cudaMalloc(&ticks, sizeof(long long)));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const double flop_per_launch = (double)grid * 8 /*warps*/ * iters * 8 * 256.0 * 2;
    double res[3], mhz[3];
    for (int mode = 0; mode < 3; ++mode) {  // 0: burst const, 1: burst data, 2: sustained data (3 s back to back)
        const bool data = mode > 0;
        const int launches = mode == 2 ? 1000 : 1;
        if (data) k_dmma<true><<<grid, 256>>>(out, ticks, iters / 8, 0.999, 1e-3);
        else k_dmma<false><<<grid, 256>>>(out, ticks, iters / 8, 0.999, 1e-3);
        cudaDeviceSynchronize();
        float best = 1e30f, total = 0;
        const int reps = mode == 2 ? 1 : 5;
        for (int r = 0; r < reps; ++r) {
            cudaEventRecord(e0);
            for (int l = 0; l < launches; ++l) {
                if (data) k_dmma<true><<<grid, 256>>>(out, ticks, iters, 0.999, 1e-3);
                else k_dmma<false><<<grid, 256>>>(out, ticks, iters, 0.999, 1e-3);
            }
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
            total = ms;
        }
        (void)total;
        long long tk = 0;
        cudaMemcpy(&tk, ticks, sizeof(tk), cudaMemcpyDeviceToHost);
        res[mode] = flop_per_launch * launches / (best * 1e-3) * 1e-12;
        // block 0 runs 1/8 .. of the launch; its tick count over its own share of the time gives the SM clock:
        // all blocks are resident at once (8 per SM), so block 0 spans the whole launch
        mhz[mode] = (double)tk / (best * 1e-3 / launches) * 1e-6;
    }
    CHECK(cudaGetLastError());
    printf("{\"sms\": %d, \"dmma_burst_const_tflops\": %.2f, \"dmma_burst_data_tflops\": %.2f, "
           "\"dmma_sustained_data_tflops\": %.2f, \"sm_mhz_burst_const\": %.0f, \"sm_mhz_burst_data\": %.0f, "
           "\"sm_mhz_sustained_data\": %.0f, \"sustained_seconds\": %.2f}\n",
           sms, res[0], res[1], res[2], mhz[0], mhz[1], mhz[2], flop_per_launch * 1000 / (res[2] * 1e12));
    return 0;
}
