// How fast can the fp64 tensor pipe be fed on B200?  DMMA m8n8k4 with operands (0) held in registers, (1) streamed
// from shared memory as the Gram kernel does (4 x ld.shared.v2.f64 per 16 DMMAs), (2) one ld.shared.f64 per DMMA as
// the transform kernels do, (3) like 1 but with twice the DMMAs per loaded fragment (register-blocked 4x2 tiles).
// One persistent CTA per SM, W warps.  Prints cycles per DMMA per SM sub-partition (16.0 = pipe peak).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_feed tools/dmma_feed.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double2 lds128(unsigned a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ double lds64(unsigned a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}

template <int MODE, bool RANDOM>
__global__ void __launch_bounds__(512, 1) k_feed(double* out, int iters) {
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) {
        if (RANDOM) {  // full-entropy mantissas, values in (-2, 2): what a centred ensemble tile looks like
            unsigned long long h = (i + 1) * 0x9E3779B97F4A7C15ull + blockIdx.x * 0xD1B54A32D192ED03ull;
            h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
            sm[i] = __longlong_as_double((h & 0x800fffffffffffffull) | 0x3ff0000000000000ull);
        } else {
            sm[i] = 1e-3 * (i % 97) - 0.04;
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned base = (unsigned)__cvta_generic_to_shared(sm) + warp * 4096 + lane * 16;
    double acc[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i][0] = acc[i][1] = 0.0;
    double ra[8], rb[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { ra[i] = 1e-3 * (lane + i); rb[i] = 2e-3 * (lane - i); }
    for (int it = 0; it < iters; ++it) {
        const unsigned o = base + (it & 7) * 512;
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) dmma(acc[i][0], acc[i][1], ra[i & 7], rb[(i * 3) & 7]);
        } else if (MODE == 1) {
            const double2 a0 = lds128(o), a1 = lds128(o + 8192), b0 = lds128(o + 16384), b1 = lds128(o + 24576);
            dmma(acc[0][0], acc[0][1], a0.x, b0.x); dmma(acc[1][0], acc[1][1], a0.x, b1.x);
            dmma(acc[2][0], acc[2][1], a1.x, b0.x); dmma(acc[3][0], acc[3][1], a1.x, b1.x);
            dmma(acc[4][0], acc[4][1], a0.x, b0.y); dmma(acc[5][0], acc[5][1], a0.x, b1.y);
            dmma(acc[6][0], acc[6][1], a1.x, b0.y); dmma(acc[7][0], acc[7][1], a1.x, b1.y);
            dmma(acc[0][0], acc[0][1], a0.y, b0.y); dmma(acc[1][0], acc[1][1], a0.y, b1.y);
            dmma(acc[2][0], acc[2][1], a1.y, b0.y); dmma(acc[3][0], acc[3][1], a1.y, b1.y);
            dmma(acc[4][0], acc[4][1], a0.y, b0.x); dmma(acc[5][0], acc[5][1], a0.y, b1.x);
            dmma(acc[6][0], acc[6][1], a1.y, b0.x); dmma(acc[7][0], acc[7][1], a1.y, b1.x);
        } else if (MODE == 2) {
            const double b0 = ra[it & 7], b1 = rb[it & 7];
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma(acc[i][0], acc[i][1], lds64(o + i * 256 - lane * 8), b0);
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma(acc[i][0], acc[i][1], lds64(o + 8192 + i * 256 - lane * 8), b1);
        } else if (MODE == 4) {  // unweighted Gram ratio: 4 x lds128 per 8 DMMAs
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const unsigned oo = o + h * 256;
                const double2 a0 = lds128(oo), a1 = lds128(oo + 8192), b0 = lds128(oo + 16384), b1 = lds128(oo + 24576);
                dmma(acc[0][0], acc[0][1], a0.x, b0.x); dmma(acc[1][0], acc[1][1], a0.x, b1.x);
                dmma(acc[2][0], acc[2][1], a1.x, b0.x); dmma(acc[3][0], acc[3][1], a1.x, b1.x);
                dmma(acc[0][0], acc[0][1], a0.y, b0.y); dmma(acc[1][0], acc[1][1], a0.y, b1.y);
                dmma(acc[2][0], acc[2][1], a1.y, b0.y); dmma(acc[3][0], acc[3][1], a1.y, b1.y);
            }
        } else if (MODE == 5) {  // weighted Gram: 5 x lds128 + 4 DMUL per 16 DMMAs
            const double2 a0 = lds128(o), a1 = lds128(o + 8192), b0 = lds128(o + 16384), b1 = lds128(o + 24576);
            const double2 w = lds128(o + 32768);
            const double2 a0w = make_double2(a0.x * w.x, a0.y * w.y), a1w = make_double2(a1.x * w.x, a1.y * w.y);
            dmma(acc[0][0], acc[0][1], a0.x, b0.x); dmma(acc[1][0], acc[1][1], a0.x, b1.x);
            dmma(acc[2][0], acc[2][1], a1.x, b0.x); dmma(acc[3][0], acc[3][1], a1.x, b1.x);
            dmma(acc[4][0], acc[4][1], a0w.x, b0.x); dmma(acc[5][0], acc[5][1], a0w.x, b1.x);
            dmma(acc[6][0], acc[6][1], a1w.x, b0.x); dmma(acc[7][0], acc[7][1], a1w.x, b1.x);
            dmma(acc[0][0], acc[0][1], a0.y, b0.y); dmma(acc[1][0], acc[1][1], a0.y, b1.y);
            dmma(acc[2][0], acc[2][1], a1.y, b0.y); dmma(acc[3][0], acc[3][1], a1.y, b1.y);
            dmma(acc[4][0], acc[4][1], a0w.y, b0.y); dmma(acc[5][0], acc[5][1], a0w.y, b1.y);
            dmma(acc[6][0], acc[6][1], a1w.y, b0.y); dmma(acc[7][0], acc[7][1], a1w.y, b1.y);
        } else if (MODE == 6) {  // diagonal macro tile: 2 x lds128 per 6 DMMAs, twice, + 4 singles (ratio of the real plan)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const unsigned oo = o + h * 256;
                const double2 a0 = lds128(oo), a1 = lds128(oo + 8192);
                dmma(acc[0][0], acc[0][1], a0.x, a0.x); dmma(acc[1][0], acc[1][1], a1.x, a0.x); dmma(acc[2][0], acc[2][1], a1.x, a1.x);
                dmma(acc[0][0], acc[0][1], a0.y, a0.y); dmma(acc[1][0], acc[1][1], a1.y, a0.y); dmma(acc[2][0], acc[2][1], a1.y, a1.y);
            }
            const double2 c0 = lds128(o + 16384), c1 = lds128(o + 24576);
            dmma(acc[3][0], acc[3][1], c0.x, c1.x); dmma(acc[4][0], acc[4][1], c0.x, c1.y);
            dmma(acc[3][0], acc[3][1], c0.y, c1.y); dmma(acc[4][0], acc[4][1], c0.y, c1.x);
        } else {
            const double2 a0 = lds128(o), a1 = lds128(o + 8192);
            const double2 b0 = lds128(o + 16384), b1 = lds128(o + 24576), b2 = lds128(o + 32768), b3 = lds128(o + 40960);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const double x0 = h ? a0.y : a0.x, x1 = h ? a1.y : a1.x;
                dmma(acc[0][0], acc[0][1], x0, h ? b0.y : b0.x); dmma(acc[1][0], acc[1][1], x0, h ? b1.y : b1.x);
                dmma(acc[2][0], acc[2][1], x0, h ? b2.y : b2.x); dmma(acc[3][0], acc[3][1], x0, h ? b3.y : b3.x);
                dmma(acc[4][0], acc[4][1], x1, h ? b0.y : b0.x); dmma(acc[5][0], acc[5][1], x1, h ? b1.y : b1.x);
                dmma(acc[6][0], acc[6][1], x1, h ? b2.y : b2.x); dmma(acc[7][0], acc[7][1], x1, h ? b3.y : b3.x);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i][0] + acc[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, bool RANDOM = false>
static void run(int sms, int warps, double* out, const char* what) {
    const int iters = 20000, per_it = 16;
    cudaFuncSetAttribute(k_feed<MODE, RANDOM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_feed<MODE, RANDOM><<<sms, warps * 32, 131072>>>(out, 100);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        k_feed<MODE, RANDOM><<<sms, warps * 32, 131072>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double dmma_per_smsp = (double)iters * per_it * warps / 4.0;
    const double cyc = best * 1e-3 * 1.965e9 / dmma_per_smsp;
    printf("{\"mode\": \"%s\", \"warps\": %d, \"ms\": %.3f, \"cycles_per_dmma_per_smsp\": %.2f, \"tflops\": %.2f}\n", what, warps,
           best, cyc, sms * 4 * 512.0 / cyc * 1.965e9 * 1e-12);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out;
    cudaMalloc(&out, sizeof(double) * sms * 512);
    for (int w : {8, 12, 16}) {
        if (w == 8) { run<0>(sms, 8, out, "registers"); run<1>(sms, 8, out, "lds128 x4 per 16"); run<2>(sms, 8, out, "lds64 per dmma"); run<3>(sms, 8, out, "lds128 x6 per 16"); }
        if (w == 12) { run<1, true>(sms, 12, out, "lds128 x4 per 16, random data"); run<3, true>(sms, 12, out, "lds128 x6 per 16, random data"); run<2, true>(sms, 12, out, "lds64 per dmma, random data"); run<4, true>(sms, 12, out, "lds128 x8 per 16 (unweighted Gram)"); run<5, true>(sms, 12, out, "lds128 x5 + 4 DMUL per 16 (weighted Gram)"); run<6, true>(sms, 12, out, "diagonal macro mix"); run<0>(sms, 12, out, "registers"); run<1>(sms, 12, out, "lds128 x4 per 16"); run<2>(sms, 12, out, "lds64 per dmma"); run<3>(sms, 12, out, "lds128 x6 per 16"); }
        if (w == 16) { run<0>(sms, 16, out, "registers"); run<1>(sms, 16, out, "lds128 x4 per 16"); run<2>(sms, 16, out, "lds64 per dmma"); run<3>(sms, 16, out, "lds128 x6 per 16"); }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
