// N-vector kernels of the CBREE engine: log-target, ESS / CVAR / failure-count reductions.
// HBM-bound streaming kernels (8-16 B per particle); fp64 warp-shuffle + block reductions with a
// fixed-order two-stage combine (deterministic, no floating-point atomics).
#include "common.cuh"

// workspace layout for the small reductions: [0,8) doubles reserved for the ticket counter,
// partial 4-tuples start at double index 8.
#define REE_WS_HEADER 8

// ---------------------------------------------------------------------------------------
// last-block finalisation of per-CTA ExpSums partials (fixed order => run-to-run identical)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void expsums_grid_finalize(ExpSums blk, double* ws, double* out4, ExpSums* sh,
                                                      bool* is_last) {
    double* part = ws + REE_WS_HEADER;
    unsigned int* ticket = reinterpret_cast<unsigned int*>(ws);
    if (threadIdx.x == 0) {
        double* p = part + 4 * (size_t)blockIdx.x;
        p[0] = blk.m; p[1] = blk.s1; p[2] = blk.s2; p[3] = blk.cnt;
        __threadfence();
        unsigned int t = atomicInc(ticket, gridDim.x - 1);  // wraps to 0: self-resetting
        *is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!*is_last) return;
    __threadfence();
    ExpSums v = expsums_empty();
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
        const volatile double* p = part + 4 * (size_t)b;
        ExpSums o{p[0], p[1], p[2], p[3]};
        v = expsums_merge(v, o);
    }
    v = expsums_block_reduce(v, sh);
    if (threadIdx.x == 0) {
        out4[0] = v.m; out4[1] = v.s1; out4[2] = v.s2; out4[3] = v.cnt;
    }
}

// ---------------------------------------------------------------------------------------
// Streaming access pattern of the N-vector kernels.  One 8-byte load per thread and iteration keeps ~0.6 MB in flight
// on the whole GPU (0.9 TB/s by Little's law); here every thread owns groups of 4 consecutive elements, fetched as two
// 16-byte loads, and requests its next group before it works on the current one; with 8 resident CTAs per SM that is
// ~10 MB in flight.  `body(i, v0)` / `body2(i, v0, v1)` see every element exactly once (order differs from the scalar
// loop only across threads; the per-thread order is ascending).
// ---------------------------------------------------------------------------------------
struct Quad {
    double v[4];
};
__device__ __forceinline__ Quad ldg_quad(const double* p) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(p)), b = __ldg(reinterpret_cast<const double2*>(p) + 1);
    return Quad{{a.x, a.y, b.x, b.y}};
}
__device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---------------------------------------------------------------------------------------
// Shifted exponential sums at streaming speed.  The generic expsums_add costs two libdevice exp calls and a divergent
// branch per element (127 instructions per element in k_reduce_scaled: issue-bound at 1.4 TB/s, profiles/
// r1_nvec_small_ncu.txt).  Here a thread raises its running shift once per group of four elements (rare after the first
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

__device__ __forceinline__ ExpSums expacc_finish(const ExpAcc& a) { return ExpSums{a.m, a.s1, a.s2, (double)a.cnt}; }

// quad streams: `quad(i0, q...)` sees four consecutive elements starting at i0 (a multiple of 4), `one(i, ...)` the tail
template <class QuadFn, class OneFn>
__device__ __forceinline__ void stream1q(const double* __restrict__ a, int64_t n, QuadFn quad, OneFn one) {
    const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t nq = aligned16(a) ? n / 4 : 0;
    int64_t q = gt;
    Quad cur{};
    if (q < nq) cur = ldg_quad(a + 4 * q);
    for (; q < nq; q += stride) {
        const Quad now = cur;
        if (q + stride < nq) cur = ldg_quad(a + 4 * (q + stride));
        quad(4 * q, now);
    }
    for (int64_t i = 4 * nq + gt; i < n; i += stride) one(i, a[i]);
}

template <class QuadFn, class OneFn>
__device__ __forceinline__ void stream2q(const double* __restrict__ a, const double* __restrict__ b, int64_t n,
                                         QuadFn quad, OneFn one) {
    const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t nq = (aligned16(a) && aligned16(b)) ? n / 4 : 0;
    int64_t q = gt;
    Quad ca{}, cb{};
    if (q < nq) {
        ca = ldg_quad(a + 4 * q);
        cb = ldg_quad(b + 4 * q);
    }
    for (; q < nq; q += stride) {
        const Quad na = ca, nb = cb;
        if (q + stride < nq) {
            ca = ldg_quad(a + 4 * (q + stride));
            cb = ldg_quad(b + 4 * (q + stride));
        }
        quad(4 * q, na, nb);
    }
    for (int64_t i = 4 * nq + gt; i < n; i += stride) one(i, a[i], b[i]);
}

template <class Body>
__device__ __forceinline__ void stream1(const double* __restrict__ a, int64_t n, Body body) {
    const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t nq = aligned16(a) ? n / 4 : 0;
    int64_t q = gt;
    Quad cur{};
    if (q < nq) cur = ldg_quad(a + 4 * q);
    for (; q < nq; q += stride) {
        const Quad now = cur;
        if (q + stride < nq) cur = ldg_quad(a + 4 * (q + stride));
#pragma unroll
        for (int k = 0; k < 4; ++k) body(4 * q + k, now.v[k]);
    }
    for (int64_t i = 4 * nq + gt; i < n; i += stride) body(i, a[i]);
}

template <class Body>
__device__ __forceinline__ void stream2(const double* __restrict__ a, const double* __restrict__ b, int64_t n, Body body) {
    const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t nq = (aligned16(a) && aligned16(b)) ? n / 4 : 0;
    int64_t q = gt;
    Quad ca{}, cb{};
    if (q < nq) {
        ca = ldg_quad(a + 4 * q);
        cb = ldg_quad(b + 4 * q);
    }
    for (; q < nq; q += stride) {
        const Quad na = ca, nb = cb;
        if (q + stride < nq) {
            ca = ldg_quad(a + 4 * (q + stride));
            cb = ldg_quad(b + 4 * (q + stride));
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) body(4 * q + k, na.v[k], nb.v[k]);
    }
    for (int64_t i = 4 * nq + gt; i < n; i += stride) body(i, a[i], b[i]);
}

// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(REE_THREADS) k_logtgt(const double* __restrict__ g, const double* __restrict__ e,
                                                        int64_t n, double sigma, int kind, double* __restrict__ ell) {
    if (e) stream2(g, e, n, [&](int64_t i, double gi, double ei) { ell[i] = ree_logtgt(gi, ei, sigma, kind); });
    else stream1(g, n, [&](int64_t i, double gi) { ell[i] = ree_logtgt(gi, 0.0, sigma, kind); });
}

// a_i = beta * logtgt(g_i, e_i; sigma)   (my_softmax(log_tgt_evals * beta), solver.py:522, 595)
__global__ void __launch_bounds__(REE_THREADS) k_reduce_ess(const double* __restrict__ g, const double* __restrict__ e,
                                                            int64_t n, double sigma, double beta, int kind,
                                                            double* __restrict__ a_out, double* ws, double* out4) {
    __shared__ ExpSums sh[32];
    __shared__ bool is_last;
    __shared__ double Tsh[64];
    const uint32_t T = load_exp_table(Tsh);
    ExpAcc acc = expacc_empty();
    const bool vec_out = a_out && aligned16(a_out);
    auto quad = [&](int64_t i0, const Quad& gq, const Quad& eq) {
        double a[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) a[k] = __dmul_rn(ree_logtgt(gq.v[k], eq.v[k], sigma, kind), beta);
        if (vec_out) {
            reinterpret_cast<double2*>(a_out + i0)[0] = make_double2(a[0], a[1]);
            reinterpret_cast<double2*>(a_out + i0)[1] = make_double2(a[2], a[3]);
        } else if (a_out) {
#pragma unroll
            for (int k = 0; k < 4; ++k) a_out[i0 + k] = a[k];
        }
        expacc_add4(acc, a, T);
    };
    auto one = [&](int64_t i, double gi, double ei) {
        const double a = __dmul_rn(ree_logtgt(gi, ei, sigma, kind), beta);
        if (a_out) a_out[i] = a;
        expacc_add1(acc, a, T);
    };
    if (e) stream2q(g, e, n, quad, one);
    else stream1q(g, n, [&](int64_t i0, const Quad& gq) { quad(i0, gq, Quad{}); }, [&](int64_t i, double gi) { one(i, gi, 0.0); });
    ExpSums blk = expsums_block_reduce(expacc_finish(acc), sh);
    expsums_grid_finalize(blk, ws, out4, sh, &is_last);
}

// a_i = scale * ell_i for a precomputed log-target vector (the beta root find re-evaluates ESS(beta) ~20 times
// for a fixed sigma: one exp per particle instead of the whole log-target)
__global__ void __launch_bounds__(REE_THREADS) k_reduce_scaled(const double* __restrict__ ell, int64_t n, double scale,
                                                               double* ws, double* out4) {
    __shared__ ExpSums sh[32];
    __shared__ bool is_last;
    __shared__ double Tsh[64];
    const uint32_t T = load_exp_table(Tsh);
    ExpAcc acc = expacc_empty();
    stream1q(ell, n,
             [&](int64_t, const Quad& lq) {
                 double a[4];
#pragma unroll
                 for (int k = 0; k < 4; ++k) a[k] = __dmul_rn(lq.v[k], scale);
                 expacc_add4(acc, a, T);
             },
             [&](int64_t, double li) { expacc_add1(acc, __dmul_rn(li, scale), T); });
    ExpSums blk = expsums_block_reduce(expacc_finish(acc), sh);
    expsums_grid_finalize(blk, ws, out4, sh, &is_last);
}

// algebraic target, sigma_new >= sigma_old: exp(a_i) = I(sigma_new; g_i) / I(sigma_old; g_i) with the smoothed
// indicator I = (1 - sigma g / sqrt(1 + sigma^2 g^2)) / 2 -- the same rounded 1 - q the reference feeds to log1p, so
// the ratio equals exp(l_new - l_old) to an ulp, without log / exp.  Entries whose l_new or l_old is -inf (1 - q
// rounds to 0) are non-finite in the reference and masked there; they are skipped here.  Ratios are <= 2: no shift.
__global__ void __launch_bounds__(REE_THREADS) k_reduce_cvar_alg(const double* __restrict__ g, int64_t n,
                                                                 double sigma_new, double sigma_old, double* ws,
                                                                 double* out4) {
    __shared__ ExpSums sh[32];
    __shared__ bool is_last;
    double s1 = 0.0, s2 = 0.0, cnt = 0.0;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const double sn2 = __dmul_rn(sigma_new, sigma_new), so2 = __dmul_rn(sigma_old, sigma_old);
    (void)stride;
    stream1(g, n, [&](int64_t, double gi) {
        const double g2 = __dmul_rn(gi, gi);
        const double qn = __ddiv_rn(__dmul_rn(-sigma_new, gi), __dsqrt_rn(__dadd_rn(1.0, __dmul_rn(sn2, g2))));
        const double qo = __ddiv_rn(__dmul_rn(-sigma_old, gi), __dsqrt_rn(__dadd_rn(1.0, __dmul_rn(so2, g2))));
        const double in = __dadd_rn(1.0, qn), io = __dadd_rn(1.0, qo);  // 1 + q as log1p forms it
        if (in > 0.0 && io > 0.0 && isfinite(in) && isfinite(io)) {
            const double v = in / io;
            s1 += v;
            s2 = fma(v, v, s2);
            cnt += 1.0;
        }
    });
    ExpSums acc{cnt > 0.0 ? 0.0 : -CUDART_INF, s1, s2, cnt};
    ExpSums blk = expsums_block_reduce(acc, sh);
    expsums_grid_finalize(blk, ws, out4, sh, &is_last);
}

// a_i = logtgt(g_i,0;sigma_new) - logtgt(g_i,0;sigma_old)   (solver.py:538-545)
__global__ void __launch_bounds__(REE_THREADS) k_reduce_cvar(const double* __restrict__ g, int64_t n,
                                                             double sigma_new, double sigma_old, int kind,
                                                             double* ws, double* out4) {
    __shared__ ExpSums sh[32];
    __shared__ bool is_last;
    __shared__ double Tsh[64];
    const uint32_t T = load_exp_table(Tsh);
    ExpAcc acc = expacc_empty();
    auto val = [&](double gi) {
        return __dsub_rn(ree_logtgt(gi, 0.0, sigma_new, kind), ree_logtgt(gi, 0.0, sigma_old, kind));
    };
    stream1q(g, n,
             [&](int64_t, const Quad& gq) {
                 double a[4];
#pragma unroll
                 for (int k = 0; k < 4; ++k) a[k] = val(gq.v[k]);
                 expacc_add4(acc, a, T);
             },
             [&](int64_t, double gi) { expacc_add1(acc, val(gi), T); });
    ExpSums blk = expsums_block_reduce(expacc_finish(acc), sh);
    expsums_grid_finalize(blk, ws, out4, sh, &is_last);
}

__global__ void __launch_bounds__(REE_THREADS) k_count_failed(const double* __restrict__ g, int64_t n, double* ws,
                                                              double* out1) {
    __shared__ double sh[32];
    __shared__ bool is_last;
    double cnt = 0.0;
    stream1(g, n, [&](int64_t, double gi) { cnt += (gi <= 0.0) ? 1.0 : 0.0; });
    cnt = block_sum(cnt, sh);
    double* part = ws + REE_WS_HEADER;
    unsigned int* ticket = reinterpret_cast<unsigned int*>(ws);
    if (threadIdx.x == 0) {
        part[blockIdx.x] = cnt;
        __threadfence();
        unsigned int t = atomicInc(ticket, gridDim.x - 1);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double v = 0.0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) v += ((const volatile double*)part)[b];
    v = block_sum(v, sh);  // integer-valued doubles: exact in any order
    if (threadIdx.x == 0) out1[0] = v;
}

// ---------------------------------------------------------------------------------------
static int nvec_grid(int64_t n) {
    int64_t want = (n + 4 * REE_THREADS - 1) / (4 * REE_THREADS);  // 4 elements per thread and pass
    int cap = 4 * ree_grid_ctas();  // 8 resident CTAs per SM; the partials region of the workspace holds 4*G tuples
    return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

extern "C" int ree_logtgt(const double* g, const double* e, int64_t n, double sigma, int tgt_kind, double* ell,
                          void* stream) {
    if (!g || !ell || n < 0) return ree_set_error(REE_E_BADARG, "ree_logtgt: bad argument");
    if (n == 0) return 0;
    k_logtgt<<<nvec_grid(n), REE_THREADS, 0, ree_stream(stream)>>>(g, e, n, sigma, tgt_kind, ell);
    return ree_check_launch("ree_logtgt");
}

extern "C" int ree_reduce_ess(const double* g, const double* e, int64_t n, double sigma, double beta, int tgt_kind,
                              double* a_out, double* workspace, double* out4, void* stream) {
    if (!g || !workspace || !out4 || n < 0) return ree_set_error(REE_E_BADARG, "ree_reduce_ess: bad argument");
    k_reduce_ess<<<nvec_grid(n), REE_THREADS, 0, ree_stream(stream)>>>(g, e, n, sigma, beta, tgt_kind, a_out, workspace,
                                                                      out4);
    return ree_check_launch("ree_reduce_ess");
}

extern "C" int ree_reduce_cvar(const double* g, int64_t n, double sigma_new, double sigma_old, int tgt_kind,
                               double* workspace, double* out4, void* stream) {
    if (!g || !workspace || !out4 || n < 0) return ree_set_error(REE_E_BADARG, "ree_reduce_cvar: bad argument");
    if (tgt_kind == REE_TGT_ALGEBRAIC && sigma_new >= sigma_old && sigma_old >= 0.0) {
        k_reduce_cvar_alg<<<nvec_grid(n), REE_THREADS, 0, ree_stream(stream)>>>(g, n, sigma_new, sigma_old, workspace,
                                                                                out4);
        return ree_check_launch("ree_reduce_cvar(algebraic)");
    }
    k_reduce_cvar<<<nvec_grid(n), REE_THREADS, 0, ree_stream(stream)>>>(g, n, sigma_new, sigma_old, tgt_kind, workspace,
                                                                        out4);
    return ree_check_launch("ree_reduce_cvar");
}

extern "C" int ree_count_failed(const double* g, int64_t n, double* workspace, double* out1, void* stream) {
    if (!g || !workspace || !out1 || n < 0) return ree_set_error(REE_E_BADARG, "ree_count_failed: bad argument");
    k_count_failed<<<nvec_grid(n), REE_THREADS, 0, ree_stream(stream)>>>(g, n, workspace, out1);
    return ree_check_launch("ree_count_failed");
}

extern "C" int ree_reduce_scaled(const double* ell, int64_t n, double scale, double* workspace, double* out4,
                                 void* stream) {
    if (!ell || !workspace || !out4 || n < 0) return ree_set_error(REE_E_BADARG, "ree_reduce_scaled: bad argument");
    k_reduce_scaled<<<nvec_grid(n), REE_THREADS, 0, ree_stream(stream)>>>(ell, n, scale, workspace, out4);
    return ree_check_launch("ree_reduce_scaled");
}
