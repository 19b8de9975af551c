// fp64 roofline probe for B200 (sm_100a): dependent-free DFMA chains and DMMA (mma.sync f64) shapes.
// SURVEY.md 8d asks for a measured fp64 peak (MEASURED_PEAKS.json only has HBM and bf16).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_peak tools/fp64_peak.cu
// Prints one JSON object.
#include <cstdio>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m8n8k4: 256 FMA per warp instruction
__global__ void __launch_bounds__(256) k_dmma884(double* out, int iters, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m16n8k8: 1024 FMA per warp instruction (sm_90+ shape)
__global__ void __launch_bounds__(256) k_dmma1688(double* out, int iters, double a, double b) {
    double c[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m16n8k16: 2048 FMA per warp instruction
__global__ void __launch_bounds__(256) k_dmma16816(double* out, int iters, double a, double b) {
    double c[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename K>
static double time_kernel(K kern, int grid, int iters, double* out) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    kern<<<grid, 256>>>(out, iters / 8, 0.999, 1e-3);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        kern<<<grid, 256>>>(out, iters, 0.999, 1e-3);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best * 1e-3;
}

int main() {
    int dev = 0, sms = 0;
    CHECK(cudaSetDevice(dev));
    CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int grid = sms * 8;
    double* out;
    CHECK(cudaMalloc(&out, sizeof(double) * grid * 256));
    const int iters = 20000;
    double threads = (double)grid * 256, warps = threads / 32;
    double t_fma = time_kernel(k_dfma, grid, iters, out);
    double t_884 = time_kernel(k_dmma884, grid, iters, out);
    double t_1688 = time_kernel(k_dmma1688, grid, iters, out);
    double t_16816 = time_kernel(k_dmma16816, grid, iters, out);
    CHECK(cudaGetLastError());
    double f_fma = threads * iters * 16 * 2 / t_fma;
    double f_884 = warps * iters * 8 * 256.0 * 2 / t_884;
    double f_1688 = warps * iters * 4 * 1024.0 * 2 / t_1688;
    double f_16816 = warps * iters * 4 * 2048.0 * 2 / t_16816;
    printf("{\"sms\": %d, \"dfma_tflops\": %.2f, \"dmma_m8n8k4_tflops\": %.2f, \"dmma_m16n8k8_tflops\": %.2f, "
           "\"dmma_m16n8k16_tflops\": %.2f}\n", sms, f_fma * 1e-12, f_884 * 1e-12, f_1688 * 1e-12, f_16816 * 1e-12);
    return 0;
}
