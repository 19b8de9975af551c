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

// Opt a kernel in to more than 48 KB of dynamic shared memory ONCE per device (the attribute is per function and
// device, not per launch).  Use inside the launcher of one kernel / one template instantiation: the flag is a
// function-local static.
#define REE_OPT_IN_SMEM(kern, bytes)                                                                     \
    do {                                                                                                 \
        static bool ree_done_[64] = {false};                                                             \
        int ree_dev_ = 0;                                                                                \
        cudaGetDevice(&ree_dev_);                                                                        \
        if (ree_dev_ >= 0 && ree_dev_ < 64 && !ree_done_[ree_dev_]) {                                    \
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes));      \
            ree_done_[ree_dev_] = true;                                                                  \
        }                                                                                                \
    } while (0)
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
        double fa = (a.m == r.m) ? 1.0 : exp(a.m - r.m), fb = (b.m == r.m) ? 1.0 : exp(b.m - r.m);  // exp(0) = 1: same bits
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

// ---------------------------------------------------------------------------------------
// Shifted exponential sums at streaming speed.  The generic expsums_add costs two libdevice exp calls and a divergent
// branch per element (127 instructions per element in k_reduce_scaled: issue-bound at 1.4 TB/s, profiles/
// r1_small_d_kernels_ncu.txt).  Here a thread raises its running shift once per group of four elements (rare after the first
// few groups) and evaluates exp(v - shift) for arguments <= 0 with a 64-entry table of 2^(j/64) in shared memory and a
// degree-5 polynomial: exp(x) = 2^e * 2^(j/64) * exp(r), n = rint(x * 64/ln2) = 64 e + j, |r| <= ln2/128, relative
// error <= 1.2 ulp (r^6/720 < 4e-17), 10 fp64 instructions.  Arguments below -708 give 0 (the reference's exp
// underflows to subnormals there: relative weight < 1e-307).
// ---------------------------------------------------------------------------------------
#include "exp_table.cuh"

__device__ __forceinline__ uint32_t load_exp_table(double* T) {  // returns the 32-bit shared address of the table
    for (int j = threadIdx.x; j < 64; j += blockDim.x) T[j] = EXP2_64[j];
    __syncthreads();
    return (uint32_t)__cvta_generic_to_shared(T);
}

// branch-free; x <= 0, -inf or NaN (NaN / -inf / x < -708 -> 0)
__device__ __forceinline__ double exp_nonpos(double x, uint32_t T) {
    const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52: the low word of x*c + MAGIC is rint(x*c)
    const double t = fma(x, 92.33248261689366, MAGIC);
    const int n = __double2loint(t);
    const double nd = t - MAGIC;
    double r = fma(nd, -0x1.62e42ff000000p-7, x);  // ln2/64 = hi + lo, nd*hi exact
    r = fma(nd, 0x1.718432a1b0e26p-41, r);
    double p = fma(r, 1.0 / 120.0, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    double tj;
    asm("ld.shared.f64 %0, [%1];" : "=d"(tj) : "r"(T + (uint32_t)((n & 63) << 3)));
    const double v = tj * p;
    const double scaled = __hiloint2double(__double2hiint(v) + ((n >> 6) << 20), __double2loint(v));
    return x >= -708.0 ? scaled : 0.0;
}

struct ExpAcc {  // per-thread running (shift, sum, sum of squares, count)
    double m, s1, s2;
    unsigned int cnt;
};
__device__ __forceinline__ ExpAcc expacc_empty() { return ExpAcc{-CUDART_INF, 0.0, 0.0, 0u}; }

__device__ __forceinline__ void expacc_raise(ExpAcc& a, double m_new, uint32_t T) {
    if (m_new > a.m) {  // exp(-inf) = 0 on the first group
        const double f = exp_nonpos(a.m - m_new, T);
        a.s1 *= f;
        a.s2 *= f * f;
        a.m = m_new;
    }
}

__device__ __forceinline__ void expacc_add4(ExpAcc& a, const double (&v)[4], uint32_t T) {
    bool ok[4];
    double m4 = -CUDART_INF;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        ok[k] = isfinite(v[k]);
        m4 = (ok[k] && v[k] > m4) ? v[k] : m4;  // compare + select (fmax carries NaN handling: ~10 instructions)
    }
    expacc_raise(a, m4, T);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double w = exp_nonpos(ok[k] ? v[k] - a.m : -CUDART_INF, T);  // select the argument: no branch
        a.s1 += w;
        a.s2 = fma(w, w, a.s2);
        a.cnt += ok[k] ? 1u : 0u;
    }
}

__device__ __forceinline__ void expacc_add1(ExpAcc& a, double v, uint32_t T) {
    if (!isfinite(v)) return;
    expacc_raise(a, v, T);
    const double w = exp_nonpos(v - a.m, T);
    a.s1 += w;
    a.s2 = fma(w, w, a.s2);
    a.cnt += 1u;
}

// one value whose exponential is multiplied by `mult` (an indicator): IS statistics of sweep 2
__device__ __forceinline__ void expacc_add1_mult(ExpAcc& a, double v, double mult, uint32_t T) {
    if (!isfinite(v)) return;
    expacc_raise(a, v, T);
    const double w = exp_nonpos(v - a.m, T) * mult;
    a.s1 += w;
    a.s2 = fma(w, w, a.s2);
    a.cnt += 1u;
}

__device__ __forceinline__ ExpSums expacc_finish(const ExpAcc& a) { return ExpSums{a.m, a.s1, a.s2, (double)a.cnt}; }

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

// Block-wide merge of per-thread ExpAcc: one max reduction, one rescale per thread (exp_nonpos), three plain sums --
// instead of ten pairwise merges with two libdevice exp each (which cost as many instructions as a thread's whole
// stream at ~8 quads per thread).  Fixed trees => deterministic.  Result valid in thread 0.
__device__ __forceinline__ ExpSums expacc_block_reduce(const ExpAcc& a, uint32_t T, double* smem /* >= 32 */) {
    double m = a.m;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double q = __shfl_xor_sync(0xffffffffu, m, o);
        m = q > m ? q : m;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if (lane == 0) smem[warp] = m;
    __syncthreads();
    double mb = -CUDART_INF;
    for (int w = 0; w < nw; ++w) mb = smem[w] > mb ? smem[w] : mb;  // every thread: the block's shift
    __syncthreads();
    const double f = exp_nonpos(a.m - mb, T);  // a.m = -inf (empty thread) or mb = -inf (empty block): 0 / NaN -> 0
    const double s1 = block_sum(a.s1 * f, smem);
    const double s2 = block_sum(a.s2 * f * f, smem);
    const double cnt = block_sum((double)a.cnt, smem);
    return ExpSums{mb, s1, s2, cnt};
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
#ifdef __CUDA_ARCH__
        // high / low halves asked for separately: they still end up in one IMAD.WIDE each, but the 64-bit product form
        // dragged extra integer adds along (250 -> 180 SASS instructions for the four blocks of a particle at d <= 8,
        // profiles/r2_xform_variants.txt).  Same bits.
        uint32_t n0 = __umulhi(M1, c[2]) ^ c[1] ^ k0;
        uint32_t n1 = M1 * c[2];
        uint32_t n2 = __umulhi(M0, c[0]) ^ c[3] ^ k1;
        uint32_t n3 = M0 * c[0];
#else
        uint64_t p0 = (uint64_t)M0 * c[0];
        uint64_t p1 = (uint64_t)M1 * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
#endif
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += W0;
        k1 += W1;
    }
}

// ---------------------------------------------------------------------------------------
// Table-driven, branch-free fp64 Box-Muller.
// DMMA and DFMA share one fp64 pipe on B200 (tools/fp64_pipe_share.cu), and under a saturating DMMA stream every
// DEPENDENT fp64 instruction of another warp costs ~28 cycles (tools/fp64_mix.cu): what counts is the number of fp64
// instructions on the critical path of a pair.  27 fp64 instructions per pair, no int<->fp conversions:
//   radius^2  w = -2 ln u1: exponent / mantissa of the 53-bit integer by shifts (no I2F), table of (1/m_i, -2 ln m_i)
//             on 128ths of [1, 2], table of k * 2 ln 2, degree-5 polynomial of -2 log1p(r), |r| <= 2^-8     (7)
//   radius    MUFU.RSQ64H seed + one cubic correction                                                          (5)
//   angle     theta from the integer remainder by the 2^52 magic-number trick inside one FMA, Taylor polynomials on
//             |theta| <= pi/128, rotation of the table entry already scaled by the radius                      (15)
// The uniforms are 53-bit integers, so all arguments are normal numbers in fixed ranges: no special cases, no slow
// paths.  Accuracy against long double: |dz| <= 4e-15 absolute (tests/test_gpu_parity.py compares with NumPy).
// ---------------------------------------------------------------------------------------
#include "bm_tables.cuh"

#define BM_TWO_LN2 1.3862943611198906
#define BM_SMEM_DOUBLES (BM_TABLE_DOUBLES + 54 + 2)  // staged tables: log (x -2) | trig | k * 2 ln 2, k = 0..53 | pad

// copy the tables into shared memory in the form the pair generator reads them (all threads of the CTA call it)
__device__ __forceinline__ void bm_stage_tables(double* dst, int tid, int nthreads) {
    for (int idx = tid; idx < BM_TABLE_DOUBLES; idx += nthreads) {
        const double v = BM_TABLE[idx];
        dst[idx] = (idx < 258 && (idx & 1)) ? -2.0 * v : v;  // T_i -> -2 T_i (exact)
    }
    for (int k = tid; k < 56; k += nthreads) dst[BM_TABLE_DOUBLES + k] = k < 54 ? (double)k * BM_TWO_LN2 : 0.0;
}

struct BmTableGlobal {  // tables read through the read-only path (modular kernels); same values as the staged form
    __device__ __forceinline__ double2 log_entry(int i) const {
        return make_double2(__ldg(BM_TABLE + 2 * i), -2.0 * __ldg(BM_TABLE + 2 * i + 1));
    }
    __device__ __forceinline__ double2 trig_entry(int j) const {
        return make_double2(__ldg(BM_TABLE + 258 + 2 * j), __ldg(BM_TABLE + 258 + 2 * j + 1));
    }
    __device__ __forceinline__ double e_entry(int k) const { return (double)k * BM_TWO_LN2; }
};

struct BmTableShared {  // tables staged in shared memory by bm_stage_tables; addr = 32-bit shared address
    uint32_t addr;
    __device__ __forceinline__ double2 log_entry(int i) const {
        double2 v;
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr + 16u * (uint32_t)i));
        return v;
    }
    __device__ __forceinline__ double2 trig_entry(int j) const {
        double2 v;
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr + 2064u + 16u * (uint32_t)j));
        return v;
    }
    __device__ __forceinline__ double e_entry(int k) const {
        double v;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr + 4112u + 8u * (uint32_t)k));
        return v;
    }
};

template <class Tab>
__device__ __forceinline__ double bm_neg2log(uint64_t k53, const Tab& tab) {
    // w = -2 ln(u), u = x * 2^-53, x = min(k53 + 1, 2^53 - 1) in [1, 2^53):  x = 2^e * m, m in [1, 2),
    // i = round((m - 1) * 128), m = m_i (1 + r) with r = m / m_i - 1 (one fma, |r| <= 2^-8):
    //   w = (53 - e) 2 ln 2 - 2 ln m_i - 2 log1p(r);  for i >= BM_LOG_I0 the table holds ln(m_i / 2) and e counts one more
    //   (keeps w = -2 log1p(r) exact in the last octave, where the other two terms would cancel).
    const uint64_t x = k53 + (k53 != 0x1fffffffffffffULL ? 1ULL : 0ULL);
    const int lz = __clzll((long long)x);             // 63 - e
    const uint64_t mant = ((x << lz) >> 11) & 0x000fffffffffffffULL;
    const int i = (int)(((mant >> 44) + 1) >> 1);
    const int ke = lz - 10 - (i >= BM_LOG_I0 ? 1 : 0);  // 53 - e (- 1)
    const double m = __longlong_as_double((long long)(mant | 0x3ff0000000000000ULL));
    const double2 t = tab.log_entry(i);                 // (1 / m_i, -2 ln m_i [+ 2 ln 2])
    const double base = tab.e_entry(ke) + t.y;
    const double r = fma(m, t.x, -1.0);
    double q = fma(r, -0.4, 0.5);                       // -2 log1p(r) = r (-2 + r (1 + r (-2/3 + r (1/2 - 2/5 r))))
    q = fma(q, r, -2.0 / 3.0);
    q = fma(q, r, 1.0);
    q = fma(q, r, -2.0);
    return fma(q, r, base);
}

__device__ __forceinline__ double bm_sqrt(double w) {
    // sqrt(w), w in [2.2e-16, 74]: hardware seed of 1/sqrt(w) (~2^-21) + one cubic correction (-> ~2^-60)
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(w));
    const double t = w * y0;
    const double e = fma(-t, y0, 1.0);
    return fma(t * e, fma(e, 0.375, 0.5), t);
}

// (r cos alpha, r sin alpha), alpha = 2 pi b53 2^-53: j = nearest 128th of the circle, theta = alpha - 2 pi j / 128 in
// [-pi/128, pi/128], angle addition with short Taylor polynomials on the table entry scaled by r
template <class Tab>
__device__ __forceinline__ void bm_rotate(uint64_t b53, double r, const Tab& tab, double& zc, double& zs) {
    const long long j = (long long)((b53 + (1ULL << 45)) >> 46);
    const long long rem = (long long)b53 - (j << 46);   // |rem| <= 2^45
    // theta = rem * 2 pi 2^-53 in ONE fma: (2^52 + 2^51 + rem) as a double is exact, the subtraction happens inside
    const double dm = __longlong_as_double(0x4338000000000000LL + rem);
    const double C = 6.283185307179586476925286766559 * 1.1102230246251565e-16;  // 2 pi 2^-53
    const double th = fma(dm, C, -6755399441055744.0 * C);
    const double t2 = th * th;
    double ps = fma(t2, -1.0 / 5040.0, 1.0 / 120.0);
    ps = fma(ps, t2, -1.0 / 6.0);
    ps = ps * t2;
    const double sn = fma(th, ps, th);                  // sin(theta)
    double pc = fma(t2, -1.0 / 720.0, 1.0 / 24.0);
    pc = fma(pc, t2, -0.5);
    pc = pc * t2;                                       // cos(theta) - 1
    const double2 sc = tab.trig_entry((int)(j & 127));  // (sin, cos) of 2 pi j / 128
    const double a = r * sc.x, b = r * sc.y;
    zs = fma(b, sn, fma(a, pc, a));
    zc = fma(-a, sn, fma(b, pc, b));
}

// two standard normals for (particle, pair, draw)
template <class Tab>
__device__ __forceinline__ void philox_normal_pair(uint64_t seed, uint64_t particle, uint32_t pair, uint64_t draw,
                                                   const Tab& tab, double& z0, double& z1) {
    uint32_t c[4] = {(uint32_t)particle, (uint32_t)(particle >> 32), pair, (uint32_t)draw};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    // u1 = (a53+1)*2^-53 in (0,1] ; u2 = b53*2^-53 in [0,1)
    const uint64_t a53 = (((uint64_t)c[0] << 32) | c[1]) >> 11;
    const uint64_t b53 = (((uint64_t)c[2] << 32) | c[3]) >> 11;
    bm_rotate(b53, bm_sqrt(bm_neg2log(a53, tab)), tab, z0, z1);
}

__device__ __forceinline__ void philox_normal_pair(uint64_t seed, uint64_t particle, uint32_t pair, uint64_t draw,
                                                   double& z0, double& z1) {
    philox_normal_pair(seed, particle, pair, draw, BmTableGlobal{}, z0, z1);
}

__host__ __device__ __forceinline__ uint32_t philox_pair_of_dim(int j) { return (uint32_t)((j >> 3) * 4 + (j & 3)); }
__host__ __device__ __forceinline__ int philox_slot_of_dim(int j) { return (j >> 2) & 1; }
