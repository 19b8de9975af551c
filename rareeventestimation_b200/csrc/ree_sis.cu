// Secondary paths on the same primitives: sequential importance sampling with adaptive conditional sampling
// (SIS-aCS, reference: sis/SIS_aCS.py:64-306) -- tempering weights, multinomial seed resampling (scan + search +
// column gather), pCN proposals and the accept/reject step.  Streaming kernels over the SoA population.
#include "common.cuh"

#define REE_WS_HEADER 8

// shifted-sum 4-tuple with shift 0: [0, sum v, sum v^2, count]  (values are O(1): ratios of normal cdfs / indicators)
__device__ __forceinline__ void plain_grid_finalize(double s1, double s2, double cnt, double* ws, double* out4) {
    __shared__ double sh[3][32];
    __shared__ bool is_last;
    s1 = block_sum(s1, sh[0]);
    s2 = block_sum(s2, sh[1]);
    cnt = block_sum(cnt, sh[2]);
    double* part = ws + REE_WS_HEADER;
    unsigned int* ticket = reinterpret_cast<unsigned int*>(ws);
    if (threadIdx.x == 0) {
        double* p = part + 4 * (size_t)blockIdx.x;
        p[0] = 0.0; p[1] = s1; p[2] = s2; p[3] = cnt;
        __threadfence();
        unsigned int t = atomicInc(ticket, gridDim.x - 1);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double a = 0.0, b = 0.0, c = 0.0;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x) {
        const volatile double* p = part + 4 * (size_t)i;
        a += p[1]; b += p[2]; c += p[3];
    }
    a = block_sum(a, sh[0]);
    b = block_sum(b, sh[1]);
    c = block_sum(c, sh[2]);
    if (threadIdx.x == 0) {
        out4[0] = 0.0; out4[1] = a; out4[2] = b; out4[3] = c;
    }
}

// tempering weight  w = Phi(-g/sigma_new) / Phi(-g/sigma_old)   (sigma_old <= 0: first level, denominator 1)
// SIS_aCS.py:151-183
__device__ __forceinline__ double sis_weight(double g, double sigma_new, double sigma_old) {
    double num = normcdf(-g / sigma_new);
    return sigma_old > 0.0 ? num / normcdf(-g / sigma_old) : num;
}

// mode 0: v = tempering weight ; mode 1: v = 1[g < 0] / Phi(-g/sigma_new)  (SIS_aCS.py:282-300, strict <)
__global__ void __launch_bounds__(REE_THREADS) k_sis_reduce(const double* __restrict__ g, int64_t n, double sigma_new,
                                                            double sigma_old, int mode, double* __restrict__ w_out,
                                                            double* ws, double* out4) {
    double s1 = 0.0, s2 = 0.0, cnt = 0.0;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double gi = g[i];
        double v = (mode == 0) ? sis_weight(gi, sigma_new, sigma_old)
                               : ((gi < 0.0) ? 1.0 / normcdf(-gi / sigma_new) : 0.0);
        if (w_out) w_out[i] = v;
        s1 += v;
        s2 = fma(v, v, s2);
        cnt += 1.0;
    }
    plain_grid_finalize(s1, s2, cnt, ws, out4);
}

// ---------------------------------------------------------------------------------------
// inclusive prefix sum (three passes: per-CTA sums, scan of the CTA sums, local scan + offset)
// ---------------------------------------------------------------------------------------
#define SCAN_T 256
#define SCAN_ITEMS 8  // elements per thread => 2048 per CTA

__global__ void __launch_bounds__(SCAN_T) k_scan_block_sums(const double* __restrict__ w, int64_t n,
                                                            double* __restrict__ bsum) {
    __shared__ double sh[32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_T * SCAN_ITEMS;
    double s = 0.0;
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        int64_t i = base + (int64_t)threadIdx.x * SCAN_ITEMS + k;
        if (i < n) s += w[i];
    }
    s = block_sum(s, sh);
    if (threadIdx.x == 0) bsum[blockIdx.x] = s;
}

__global__ void __launch_bounds__(1024) k_scan_of_sums(double* __restrict__ bsum, int nb) {
    // exclusive scan of the CTA sums, sequential chunks per thread + shared scan (nb <= a few 1e5)
    __shared__ double sh[1024];
    const int per = (nb + 1023) / 1024;
    const int lo = threadIdx.x * per, hi = min(nb, lo + per);
    double s = 0.0;
    for (int i = lo; i < hi; ++i) s += bsum[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double run = 0.0;
        for (int t = 0; t < 1024; ++t) {
            double v = sh[t];
            sh[t] = run;
            run += v;
        }
    }
    __syncthreads();
    double run = sh[threadIdx.x];
    for (int i = lo; i < hi; ++i) {
        double v = bsum[i];
        bsum[i] = run;
        run += v;
    }
}

__global__ void __launch_bounds__(SCAN_T) k_scan_apply(const double* __restrict__ w, int64_t n,
                                                       const double* __restrict__ boff, double* __restrict__ cdf) {
    __shared__ double sh[SCAN_T];
    const int64_t base = (int64_t)blockIdx.x * SCAN_T * SCAN_ITEMS;
    double v[SCAN_ITEMS], s = 0.0;
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        int64_t i = base + (int64_t)threadIdx.x * SCAN_ITEMS + k;
        v[k] = (i < n) ? w[i] : 0.0;
        s += v[k];
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    // Hillis-Steele inclusive scan of the per-thread sums
    for (int o = 1; o < SCAN_T; o <<= 1) {
        double t = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0.0;
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    double run = boff[blockIdx.x] + sh[threadIdx.x] - s;
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        int64_t i = base + (int64_t)threadIdx.x * SCAN_ITEMS + k;
        run += v[k];
        if (i < n) cdf[i] = run;
    }
}

// multinomial seed selection: idx_k = first i with cdf[i] > u_k * cdf[n-1]   (np.random.choice with p, SIS_aCS.py:194)
__global__ void __launch_bounds__(REE_THREADS) k_sis_search(const double* __restrict__ cdf, int64_t n, int64_t nchain,
                                                            uint64_t seed, uint64_t draw, int64_t first,
                                                            int64_t* __restrict__ idx) {
    const double total = cdf[n - 1];
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nchain; k += stride) {
        uint32_t c[4] = {(uint32_t)(first + k), (uint32_t)((uint64_t)(first + k) >> 32), 0xFFFFFFF0u, (uint32_t)draw};
        philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        const double u = (double)((((uint64_t)c[0] << 32) | c[1]) >> 11) * 1.1102230246251565e-16;
        const double target = u * total;
        int64_t lo = 0, hi = n - 1;
        while (lo < hi) {
            int64_t mid = (lo + hi) >> 1;
            if (cdf[mid] > target) hi = mid; else lo = mid + 1;
        }
        idx[k] = lo;
    }
}

// seeds: out[j][k] = x[j][idx_k], g_out[k] = g[idx_k]
__global__ void __launch_bounds__(REE_THREADS) k_gather_columns(const double* __restrict__ x, int d, int64_t ld_src,
                                                                const int64_t* __restrict__ idx, int64_t nchain,
                                                                double* __restrict__ out, int64_t ld_out,
                                                                const double* __restrict__ g, double* __restrict__ g_out) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nchain; k += stride) {
        const int64_t s = idx[k];
        for (int j = 0; j < d; ++j) out[(int64_t)j * ld_out + k] = x[(int64_t)j * ld_src + s];
        if (g_out) g_out[k] = g[s];
    }
}

// pCN proposal  cand = rho * cur + sigma_f * z   (np.random.normal(loc=rho*u0, scale=sigma_f), SIS_aCS.py:232)
__global__ void __launch_bounds__(REE_THREADS) k_acs_propose(const double* __restrict__ cur, int64_t nchain, int d,
                                                             int64_t ld, double rho, double sigma_f, uint64_t seed,
                                                             uint64_t draw, int64_t first, double* __restrict__ cand) {
    const int npairs = ((d + 7) / 8) * 4;
    const int pair = blockIdx.y;
    if (pair >= npairs) return;
    const int j0 = (pair >> 2) * 8 + (pair & 3), j1 = j0 + 4;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nchain; k += stride) {
        double z0, z1;
        philox_normal_pair(seed, (uint64_t)(first + k), (uint32_t)pair, draw, z0, z1);
        if (j0 < d) cand[(int64_t)j0 * ld + k] = fma(sigma_f, z0, rho * cur[(int64_t)j0 * ld + k]);
        if (j1 < d) cand[(int64_t)j1 * ld + k] = fma(sigma_f, z1, rho * cur[(int64_t)j1 * ld + k]);
    }
}

// accept / reject (SIS_aCS.py:238-254): alpha = min(1, Phi(-g_c/s)/Phi(-g_0/s)); the chain state moves to the
// candidate with probability alpha; the (possibly repeated) state is appended to the new population at column
// out_col0 + k.  stats[0] += sum alpha (per-CTA partials -> fixed-order finalisation).
__global__ void __launch_bounds__(REE_THREADS) k_acs_accept(double* __restrict__ cur, double* __restrict__ g_cur,
                                                            const double* __restrict__ cand,
                                                            const double* __restrict__ g_cand, int64_t nchain, int d,
                                                            int64_t ld, double sigma, uint64_t seed, uint64_t draw,
                                                            int64_t first, double* __restrict__ out_x, int64_t ld_out,
                                                            int64_t out_col0, double* __restrict__ out_g, double* ws,
                                                            double* out4) {
    double s_alpha = 0.0, n_acc = 0.0, cnt = 0.0;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nchain; k += stride) {
        const double g0 = g_cur[k], gc = g_cand[k];
        const double alpha = fmin(1.0, normcdf(-gc / sigma) / normcdf(-g0 / sigma));
        uint32_t c[4] = {(uint32_t)(first + k), (uint32_t)((uint64_t)(first + k) >> 32), 0xFFFFFFF1u, (uint32_t)draw};
        philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        const double u = (double)((((uint64_t)c[0] << 32) | c[1]) >> 11) * 1.1102230246251565e-16;
        const bool acc = u <= alpha;  // also false for NaN alpha
        const double gn = acc ? gc : g0;
        g_cur[k] = gn;
        out_g[out_col0 + k] = gn;
        for (int j = 0; j < d; ++j) {
            const double v = acc ? cand[(int64_t)j * ld + k] : cur[(int64_t)j * ld + k];
            if (acc) cur[(int64_t)j * ld + k] = v;
            out_x[(int64_t)j * ld_out + out_col0 + k] = v;
        }
        s_alpha += isfinite(alpha) ? alpha : 0.0;
        n_acc += acc ? 1.0 : 0.0;
        cnt += 1.0;
    }
    plain_grid_finalize(s_alpha, n_acc, cnt, ws, out4);
}

// ---------------------------------------------------------------------------------------
static int grid_for(int64_t n) {
    int64_t want = (n + REE_THREADS - 1) / REE_THREADS;
    int cap = ree_grid_ctas();
    return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

extern "C" int ree_sis_reduce(const double* g, int64_t n, double sigma_new, double sigma_old, int mode, double* w_out,
                              double* workspace, double* out4, void* stream) {
    if (!g || !workspace || !out4 || n < 1 || !(sigma_new > 0.0))
        return ree_set_error(REE_E_BADARG, "ree_sis_reduce: bad argument");
    k_sis_reduce<<<grid_for(n), REE_THREADS, 0, ree_stream(stream)>>>(g, n, sigma_new, sigma_old, mode, w_out, workspace,
                                                                      out4);
    return ree_check_launch("ree_sis_reduce");
}

extern "C" int64_t ree_scan_scratch_doubles(int64_t n) { return (n + SCAN_T * SCAN_ITEMS - 1) / (SCAN_T * SCAN_ITEMS) + 8; }

extern "C" int ree_scan_inclusive(const double* w, int64_t n, double* cdf, double* scratch, void* stream) {
    if (!w || !cdf || !scratch || n < 1) return ree_set_error(REE_E_BADARG, "ree_scan_inclusive: bad argument");
    cudaStream_t st = ree_stream(stream);
    const int nb = (int)((n + SCAN_T * SCAN_ITEMS - 1) / (SCAN_T * SCAN_ITEMS));
    k_scan_block_sums<<<nb, SCAN_T, 0, st>>>(w, n, scratch);
    int rc = ree_check_launch("k_scan_block_sums");
    if (rc) return rc;
    k_scan_of_sums<<<1, 1024, 0, st>>>(scratch, nb);
    rc = ree_check_launch("k_scan_of_sums");
    if (rc) return rc;
    k_scan_apply<<<nb, SCAN_T, 0, st>>>(w, n, scratch, cdf);
    return ree_check_launch("k_scan_apply");
}

extern "C" int ree_sis_resample(const double* cdf, int64_t n, int64_t nchain, uint64_t seed, uint64_t draw,
                                int64_t first, int64_t* idx, void* stream) {
    if (!cdf || !idx || n < 1 || nchain < 1) return ree_set_error(REE_E_BADARG, "ree_sis_resample: bad argument");
    k_sis_search<<<grid_for(nchain), REE_THREADS, 0, ree_stream(stream)>>>(cdf, n, nchain, seed, draw, first, idx);
    return ree_check_launch("ree_sis_resample");
}

extern "C" int ree_gather_columns(const double* x, int d, int64_t ld_src, const int64_t* idx, int64_t nchain, double* out,
                                  int64_t ld_out, const double* g, double* g_out, void* stream) {
    if (!x || !idx || !out || d < 1 || nchain < 1) return ree_set_error(REE_E_BADARG, "ree_gather_columns: bad argument");
    k_gather_columns<<<grid_for(nchain), REE_THREADS, 0, ree_stream(stream)>>>(x, d, ld_src, idx, nchain, out, ld_out, g,
                                                                               g_out);
    return ree_check_launch("ree_gather_columns");
}

extern "C" int ree_acs_propose(const double* cur, int64_t nchain, int d, int64_t ld, double rho, double sigma_f,
                               uint64_t seed, uint64_t draw, int64_t first, double* cand, void* stream) {
    if (!cur || !cand || d < 1 || nchain < 1) return ree_set_error(REE_E_BADARG, "ree_acs_propose: bad argument");
    dim3 grid((unsigned)grid_for(nchain), (unsigned)(((d + 7) / 8) * 4));
    k_acs_propose<<<grid, REE_THREADS, 0, ree_stream(stream)>>>(cur, nchain, d, ld, rho, sigma_f, seed, draw, first, cand);
    return ree_check_launch("ree_acs_propose");
}

extern "C" int ree_acs_accept(double* cur, double* g_cur, const double* cand, const double* g_cand, int64_t nchain, int d,
                              int64_t ld, double sigma, uint64_t seed, uint64_t draw, int64_t first, double* out_x,
                              int64_t ld_out, int64_t out_col0, double* out_g, double* workspace, double* out4,
                              void* stream) {
    if (!cur || !g_cur || !cand || !g_cand || !out_x || !out_g || !workspace || !out4 || d < 1 || nchain < 1)
        return ree_set_error(REE_E_BADARG, "ree_acs_accept: bad argument");
    k_acs_accept<<<grid_for(nchain), REE_THREADS, 0, ree_stream(stream)>>>(cur, g_cur, cand, g_cand, nchain, d, ld, sigma,
                                                                           seed, draw, first, out_x, ld_out, out_col0,
                                                                           out_g, workspace, out4);
    return ree_check_launch("ree_acs_accept");
}
