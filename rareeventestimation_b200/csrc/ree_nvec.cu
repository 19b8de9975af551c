// N-vector kernels of the CBREE engine: log-target, ESS / CVAR / failure-count reductions.
// HBM-bound streaming kernels (8-16 B per particle); fp64 warp-shuffle + block reductions with a
// fixed-order two-stage combine (deterministic, no floating-point atomics).
#include "nvec_stream.cuh"

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
    __shared__ double Tsh[64], red[32];
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
    ExpSums blk = expacc_block_reduce(acc, T, red);
    expsums_grid_finalize(blk, ws, out4, sh, &is_last);
}

// a_i = scale * ell_i for a precomputed log-target vector (the beta root find re-evaluates ESS(beta) ~20 times
// for a fixed sigma: one exp per particle instead of the whole log-target)
__global__ void __launch_bounds__(REE_THREADS) k_reduce_scaled(const double* __restrict__ ell, int64_t n, double scale,
                                                               double* ws, double* out4) {
    __shared__ ExpSums sh[32];
    __shared__ bool is_last;
    __shared__ double Tsh[64], red[32];
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
    ExpSums blk = expacc_block_reduce(acc, T, red);
    expsums_grid_finalize(blk, ws, out4, sh, &is_last);
}

// Same sums with the shift known in advance: a = scale * ell is monotone in ell, so max a = scale * (max ell or min ell)
// exactly (rounding is monotone).  No running shift, no divergent block: ~28 instructions per element.
__global__ void __launch_bounds__(REE_THREADS) k_reduce_scaled_known(const double* __restrict__ ell, int64_t n,
                                                                     double scale, const double* __restrict__ range2,
                                                                     double* ws, double* out4) {
    __shared__ ExpSums sh[32];
    __shared__ bool is_last;
    __shared__ double Tsh[64], red[32];
    const uint32_t T = load_exp_table(Tsh);
    const double M = __dmul_rn(scale >= 0.0 ? range2[0] : range2[1], scale);
    double s1 = 0.0, s2 = 0.0;
    unsigned int cnt = 0;
    auto one = [&](double li) {
        const double a = __dmul_rn(li, scale);
        const bool ok = isfinite(a);
        const double w = exp_nonpos(ok ? a - M : -CUDART_INF, T);
        s1 += w;
        s2 = fma(w, w, s2);
        cnt += ok ? 1u : 0u;
    };
    stream1q(ell, n,
             [&](int64_t, const Quad& lq) {
#pragma unroll
                 for (int k = 0; k < 4; ++k) one(lq.v[k]);
             },
             [&](int64_t, double li) { one(li); });
    s1 = block_sum(s1, red);  // every thread used the same shift: plain sums
    s2 = block_sum(s2, red);
    const double cn = block_sum((double)cnt, red);
    ExpSums blk{cn > 0.0 ? M : -CUDART_INF, s1, s2, cn};
    expsums_grid_finalize(blk, ws, out4, sh, &is_last);
}

// out2 = [max, min] over the finite entries of v (-inf / +inf if there are none); two-stage, order-independent
__global__ void __launch_bounds__(REE_THREADS) k_vector_range(const double* __restrict__ v, int64_t n, double* ws,
                                                              double* out2) {
    __shared__ double shx[32], shn[32];
    __shared__ bool is_last;
    double mx = -CUDART_INF, mn = CUDART_INF;
    stream1(v, n, [&](int64_t, double x) {
        if (isfinite(x)) {
            mx = x > mx ? x : mx;
            mn = x < mn ? x : mn;
        }
    });
    auto block_minmax = [&]() {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double a = __shfl_xor_sync(0xffffffffu, mx, o), b = __shfl_xor_sync(0xffffffffu, mn, o);
            mx = a > mx ? a : mx;
            mn = b < mn ? b : mn;
        }
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
        if (lane == 0) {
            shx[warp] = mx;
            shn[warp] = mn;
        }
        __syncthreads();
        if (warp == 0) {
            mx = lane < nw ? shx[lane] : -CUDART_INF;
            mn = lane < nw ? shn[lane] : CUDART_INF;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double a = __shfl_xor_sync(0xffffffffu, mx, o), b = __shfl_xor_sync(0xffffffffu, mn, o);
                mx = a > mx ? a : mx;
                mn = b < mn ? b : mn;
            }
        }
        __syncthreads();
    };
    block_minmax();
    double* part = ws + REE_WS_HEADER;
    unsigned int* ticket = reinterpret_cast<unsigned int*>(ws);
    if (threadIdx.x == 0) {
        part[2 * blockIdx.x] = mx;
        part[2 * blockIdx.x + 1] = mn;
        __threadfence();
        unsigned int t = atomicInc(ticket, gridDim.x - 1);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    mx = -CUDART_INF;
    mn = CUDART_INF;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
        const double a = ((const volatile double*)part)[2 * b], c = ((const volatile double*)part)[2 * b + 1];
        mx = a > mx ? a : mx;
        mn = c < mn ? c : mn;
    }
    block_minmax();
    if (threadIdx.x == 0) {
        out2[0] = mx;
        out2[1] = mn;
    }
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
    __shared__ double Tsh[64], red[32];
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
    ExpSums blk = expacc_block_reduce(acc, T, red);
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

extern "C" int ree_reduce_scaled(const double* ell, int64_t n, double scale, const double* range2, double* workspace,
                                 double* out4, void* stream) {
    if (!ell || !workspace || !out4 || n < 0) return ree_set_error(REE_E_BADARG, "ree_reduce_scaled: bad argument");
    if (range2)
        k_reduce_scaled_known<<<nvec_grid(n), REE_THREADS, 0, ree_stream(stream)>>>(ell, n, scale, range2, workspace, out4);
    else
        k_reduce_scaled<<<nvec_grid(n), REE_THREADS, 0, ree_stream(stream)>>>(ell, n, scale, workspace, out4);
    return ree_check_launch("ree_reduce_scaled");
}

extern "C" int ree_vector_range(const double* v, int64_t n, double* workspace, double* out2, void* stream) {
    if (!v || !workspace || !out2 || n < 0) return ree_set_error(REE_E_BADARG, "ree_vector_range: bad argument");
    k_vector_range<<<nvec_grid(n), REE_THREADS, 0, ree_stream(stream)>>>(v, n, workspace, out2);
    return ree_check_launch("ree_vector_range");
}
