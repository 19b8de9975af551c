// Shared device helpers for the CBREE engine (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/ree_b200.h"

#define REE_THREADS 256
#define REE_LOG_2PI 1.8378770664093454835606594728112

// ---------------------------------------------------------------------------------------
// error plumbing (ree_api.cu owns the storage)
// ---------------------------------------------------------------------------------------
int ree_set_error(int code, const char* what);
int ree_check_launch(const char* what);
int ree_grid_ctas();

static inline cudaStream_t ree_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---------------------------------------------------------------------------------------
// log target density, solver.py:321-363.  Same operation order as the NumPy expression;
// explicit _rn intrinsics so that nvcc cannot contract a*b+c into an FMA (the algebraic
// form cancels catastrophically for sigma*g >> 1).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double ree_log_ndtr(double x) {
    // log Phi(x) the way scipy.special.log_ndtr evaluates it (behind norm.logcdf, solver.py:358)
    double t = x * 0.70710678118654752440;
    if (x < -1.0) return log(erfcx(-t) * 0.5) - t * t;
    return log1p(-erfc(t) * 0.5);
}

__device__ __forceinline__ double ree_logtgt(double g, double e, double sigma, int kind) {
    const double neg_log2 = -0.69314718055994530942;  // -log(2)
    switch (kind) {
        case REE_TGT_TANH: {  // -log(2) + log1p(tanh(-sigma*g)) - e
            double a = __dmul_rn(-sigma, g);
            return __dsub_rn(__dadd_rn(neg_log2, log1p(tanh(a))), e);
        }
        case REE_TGT_ARCTAN: {  // -log(2) + log1p(2/pi*arctan(-sigma*g)) - e
            double a = atan(__dmul_rn(-sigma, g));
            double b = __dmul_rn(0.63661977236758134308 /* 2/pi */, a);
            return __dsub_rn(__dadd_rn(neg_log2, log1p(b)), e);
        }
        case REE_TGT_ERF: {  // norm.logcdf(-sigma*g) - e
            return __dsub_rn(ree_log_ndtr(__dmul_rn(-sigma, g)), e);
        }
        case REE_TGT_RELU: {  // -sigma*max(0,g)**3 - e
            double m = fmax(0.0, g);
            double m3 = __dmul_rn(__dmul_rn(m, m), m);
            return __dsub_rn(__dmul_rn(-sigma, m3), e);
        }
        case REE_TGT_SIGMOID: {  // -log1p(exp(sigma*g)) - e
            return __dsub_rn(-log1p(exp(__dmul_rn(sigma, g))), e);
        }
        default: {  // algebraic: -log(2) + log1p(-sigma*g/sqrt(1+sigma**2*g**2)) - e
            double num = __dmul_rn(-sigma, g);
            double s2g2 = __dmul_rn(__dmul_rn(sigma, sigma), __dmul_rn(g, g));
            double den = __dsqrt_rn(__dadd_rn(1.0, s2g2));
            double q = __ddiv_rn(num, den);
            return __dsub_rn(__dadd_rn(neg_log2, log1p(q)), e);
        }
    }
}

// ---------------------------------------------------------------------------------------
// shifted exponential sums: the triple (M, S1, S2) represents sum exp(a_i) = exp(M)*S1 and
// sum exp(2 a_i) = exp(2M)*S2; merging rescales to the larger shift (online softmax).
// ---------------------------------------------------------------------------------------
struct ExpSums {
    double m, s1, s2, cnt;
};

__device__ __forceinline__ ExpSums expsums_empty() { return ExpSums{-CUDART_INF, 0.0, 0.0, 0.0}; }

__device__ __forceinline__ void expsums_add(ExpSums& a, double v, double mult) {
    // add one finite log-value v whose exp is multiplied by `mult` (0/1 indicator or 1)
    if (v > a.m) {
        double f = exp(a.m - v);  // exp(-inf) = 0 on the first element
        a.s1 *= f;
        a.s2 *= f * f;
        a.m = v;
    }
    double w = exp(v - a.m) * mult;
    a.s1 += w;
    a.s2 += w * w;
    a.cnt += 1.0;
}

__device__ __forceinline__ ExpSums expsums_merge(const ExpSums& a, const ExpSums& b) {
    ExpSums r;
    r.m = fmax(a.m, b.m);
    if (r.m == -CUDART_INF) {
        r.s1 = 0.0;
        r.s2 = 0.0;
    } else {
        double fa = exp(a.m - r.m), fb = exp(b.m - r.m);
        r.s1 = a.s1 * fa + b.s1 * fb;
        r.s2 = a.s2 * fa * fa + b.s2 * fb * fb;
    }
    r.cnt = a.cnt + b.cnt;
    return r;
}

__device__ __forceinline__ ExpSums expsums_shfl_xor(const ExpSums& a, int lane_mask) {
    ExpSums r;
    r.m = __shfl_xor_sync(0xffffffffu, a.m, lane_mask);
    r.s1 = __shfl_xor_sync(0xffffffffu, a.s1, lane_mask);
    r.s2 = __shfl_xor_sync(0xffffffffu, a.s2, lane_mask);
    r.cnt = __shfl_xor_sync(0xffffffffu, a.cnt, lane_mask);
    return r;
}

// block-wide merge (fixed tree => deterministic); result valid in thread 0
__device__ __forceinline__ ExpSums expsums_block_reduce(ExpSums v, ExpSums* smem /* >= 32 */) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = expsums_merge(v, expsums_shfl_xor(v, o));
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < nw ? smem[lane] : expsums_empty();
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = expsums_merge(v, expsums_shfl_xor(v, o));
    }
    __syncthreads();
    return v;
}

// plain fp64 sums -------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double block_sum(double v, double* smem /* >= 32 */) {
    v = warp_sum(v);
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < nw ? smem[lane] : 0.0;
        v = warp_sum(v);
    }
    __syncthreads();
    return v;
}

// ---------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011) + Box-Muller in fp64.
// counter = (particle lo, particle hi, dim pair, draw[31:0]), key = seed.  One call yields two N(0,1).
// Pair p of dimension j: p = (j/8)*4 + (j%4), slot = (j/4)%2  (dims j and j+4 of an 8-block
// share one call; matches the per-thread ownership of the MMA fragments in ree_step.cu).
// ---------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c[0];
        uint64_t p1 = (uint64_t)M1 * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += W0;
        k1 += W1;
    }
}

// two standard normals for (particle, pair, draw)
__device__ __forceinline__ void philox_normal_pair(uint64_t seed, uint64_t particle, uint32_t pair, uint64_t draw,
                                                   double& z0, double& z1) {
    uint32_t c[4] = {(uint32_t)particle, (uint32_t)(particle >> 32), pair, (uint32_t)draw};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    // u1 in (0,1]: 53 bits ; u2 in [0,1): 53 bits
    uint64_t a = ((uint64_t)c[0] << 32) | c[1];
    uint64_t b = ((uint64_t)c[2] << 32) | c[3];
    double u1 = ((double)(a >> 11) + 1.0) * 1.1102230246251565e-16;  // 2^-53
    double u2 = (double)(b >> 11) * 1.1102230246251565e-16;
    double r = sqrt(-2.0 * log(u1));
    double s, co;
    sincospi(2.0 * u2, &s, &co);
    z0 = r * co;
    z1 = r * s;
}

__host__ __device__ __forceinline__ uint32_t philox_pair_of_dim(int j) { return (uint32_t)((j >> 3) * 4 + (j & 3)); }
__host__ __device__ __forceinline__ int philox_slot_of_dim(int j) { return (j >> 2) & 1; }
