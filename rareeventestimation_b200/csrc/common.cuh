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

// ---------------------------------------------------------------------------------------
// Branch-free fp64 Box-Muller.  The uniforms are 53-bit integers, so the arguments of log / sincospi are
// known to be normal numbers in a fixed range: no special cases, no slow paths -- straight-line code that the
// scheduler can interleave with DMMA issue.  Accuracy (checked against long double on 2e6 inputs): log <= 3.2 ulp,
// sqrt <= 1.5 ulp, sin/cos <= 1.8e-16 absolute.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double bm_neg2log(uint64_t k53) {
    // -2*ln(u), u = (k53+1)*2^-53 in (0,1]
    const double x = (double)(long long)(k53 + 1);  // exact, in [1, 2^53]
    const long long bits = __double_as_longlong(x);
    int e = (int)(bits >> 52) - (1023 + 53);
    double m = __longlong_as_double((bits & 0x000fffffffffffffLL) | 0x3ff0000000000000LL);  // [1,2)
    const bool big = m > 1.4142135623730951;
    m = big ? m * 0.5 : m;
    e = big ? e + 1 : e;
    const double f = m - 1.0;   // exact
    const double y = 2.0 + f;   // [1.7, 2.42]
    double r = (double)__frcp_rn((float)y);
    r = r * fma(-y, r, 2.0);
    r = r * fma(-y, r, 2.0);
    double s = f * r;
    s = fma(r, fma(-s, y, f), s);  // s = f / y
    const double s2 = s * s;
    double p = 2.0 / 21.0;
    p = fma(p, s2, 2.0 / 19.0);
    p = fma(p, s2, 2.0 / 17.0);
    p = fma(p, s2, 2.0 / 15.0);
    p = fma(p, s2, 2.0 / 13.0);
    p = fma(p, s2, 2.0 / 11.0);
    p = fma(p, s2, 2.0 / 9.0);
    p = fma(p, s2, 2.0 / 7.0);
    p = fma(p, s2, 2.0 / 5.0);
    p = fma(p, s2, 2.0 / 3.0);
    p = p * s2;
    const double logm = fma(s, p, s + s);  // 2 atanh(s)
    const double ed = (double)e;
    const double ln_u = fma(ed, 0.6931471803691238, fma(ed, 1.9082149292705877e-10, logm));
    return -2.0 * ln_u;
}

__device__ __forceinline__ double bm_sqrt(double w) {
    // sqrt(w), w in [0, 74]
    const float wf = fmaxf((float)w, 1e-30f);
    const double rs = (double)rsqrtf(wf);
    double r = w * rs;
    const double h = 0.5 * rs;
    r = fma(h, fma(-r, r, w), r);
    r = fma(h, fma(-r, r, w), r);
    return r;
}

__device__ __forceinline__ void bm_sincos2pi(uint64_t b53, double& sn, double& cs) {
    // sin / cos of 2*pi*u, u = b53*2^-53: octant from the top 3 bits, Taylor polynomials on [0, pi/4]
    const int o = (int)(b53 >> 50);
    const long long r50 = (long long)(b53 & ((1ULL << 50) - 1));
    const long long xi = (o & 1) ? ((1LL << 50) - r50) : r50;
    const double y = (double)xi * (0.7853981633974483 * 8.8817841970012523e-16);  // * pi/4 * 2^-50
    const double y2 = y * y;
    double ps = -1.0 / 1307674368000.0;          // -1/15!
    ps = fma(ps, y2, 1.0 / 6227020800.0);        // 1/13!
    ps = fma(ps, y2, -1.0 / 39916800.0);         // -1/11!
    ps = fma(ps, y2, 1.0 / 362880.0);            // 1/9!
    ps = fma(ps, y2, -1.0 / 5040.0);             // -1/7!
    ps = fma(ps, y2, 1.0 / 120.0);               // 1/5!
    ps = fma(ps, y2, -1.0 / 6.0);                // -1/3!
    ps = ps * y2;
    const double S = fma(y, ps, y);
    double pc = 1.0 / 20922789888000.0;          // 1/16!
    pc = fma(pc, y2, -1.0 / 87178291200.0);      // -1/14!
    pc = fma(pc, y2, 1.0 / 479001600.0);         // 1/12!
    pc = fma(pc, y2, -1.0 / 3628800.0);          // -1/10!
    pc = fma(pc, y2, 1.0 / 40320.0);             // 1/8!
    pc = fma(pc, y2, -1.0 / 720.0);              // -1/6!
    pc = fma(pc, y2, 1.0 / 24.0);                // 1/4!
    pc = fma(pc, y2, -0.5);
    const double C = fma(pc, y2, 1.0);
    const int q = o & 3;
    const bool keep = (q == 0) || (q == 3);
    double sv = keep ? S : C;
    double cv = keep ? C : S;
    cv = (q >= 2) ? -cv : cv;
    const bool neg = o >= 4;
    sn = neg ? -sv : sv;
    cs = neg ? -cv : cv;
}

// two standard normals for (particle, pair, draw)
__device__ __forceinline__ void philox_normal_pair(uint64_t seed, uint64_t particle, uint32_t pair, uint64_t draw,
                                                   double& z0, double& z1) {
    uint32_t c[4] = {(uint32_t)particle, (uint32_t)(particle >> 32), pair, (uint32_t)draw};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    // u1 = (a53+1)*2^-53 in (0,1] ; u2 = b53*2^-53 in [0,1)
    const uint64_t a53 = (((uint64_t)c[0] << 32) | c[1]) >> 11;
    const uint64_t b53 = (((uint64_t)c[2] << 32) | c[3]) >> 11;
    const double r = bm_sqrt(bm_neg2log(a53));
    double sn, cs;
    bm_sincos2pi(b53, sn, cs);
    z0 = r * cs;
    z1 = r * sn;
}

__host__ __device__ __forceinline__ uint32_t philox_pair_of_dim(int j) { return (uint32_t)((j >> 3) * 4 + (j & 3)); }
__host__ __device__ __forceinline__ int philox_slot_of_dim(int j) { return (j >> 2) & 1; }
