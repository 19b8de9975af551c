// The two ensemble sweeps of a CBREE iteration:
//   sweep 1  ree_cbree_step : X' = t X + m_noise + F z, g', e', unweighted shifted sums
//   sweep 2  ree_cbree_maha : Mahalanobis / log q, IS statistics, weights, weighted shifted sums
// plus ree_ensemble_sums (moments of an existing ensemble).
//
// v0 kernels: thread-per-particle contraction with the noise / centred tile staged in shared
// memory, and a shared-memory tiled Gram with register accumulators.  All cross-CTA combines are
// fixed-order two-stage reductions (deterministic).
#include "common.cuh"

int ree_lsf_launch(const ree_lsf* lsf, const double* x, int64_t n, int64_t ld, double* g, cudaStream_t st);
bool ree_mma_supported(int d);
size_t ree_mma_partial_doubles(int d);
int ree_mma_step(const double* x, double* x_new, int64_t n, int d, int64_t ld, double t, const double* m_noise,
                 const double* F, int tri, const double* z, uint64_t seed, uint64_t draw, int64_t first, int lsf_mode,
                 const double* coef, double offset, const double* shift, double* g_new, double* e_new, double* ws_big,
                 double* ws_small, double* sums, cudaStream_t st);
int ree_mma_maha(const double* x, int64_t n, int d, int64_t ld, const double* g, const double* e, const double* mean,
                 const double* Linv, const double* logdet, double sigma, double beta, int tgt_kind, const double* wmax,
                 double* logq, double* w, double* ws_big, double* ws_small, double* sums, double* is4, cudaStream_t st);
int ree_mma_sums(const double* x, int64_t n, int d, int64_t ld, const double* g, const double* shift, double* ws_big,
                 double* ws_small, double* sums, cudaStream_t st);

bool ree_small_supported(int d);
int ree_small_step(const double* x, double* x_new, int64_t n, int d, int64_t ld, double t, const double* m_noise,
                   const double* F, int tri, const double* z, uint64_t seed, uint64_t draw, int64_t first, int lsf_kind,
                   const double* coef, double offset, const double* shift, double* g_new, double* e_new, double* ws_big,
                   double* sums, cudaStream_t st);
int ree_small_post(const double* x, int64_t n, int d, int64_t ld, const double* g, const double* e, const double* mean,
                   const double* Linv, const double* logdet, double sigma, double beta, int tgt_kind, const double* wmax,
                   double* logq, double* w, double* ws_big, double* ws_small, double* sums, double* is4,
                   cudaStream_t st);
bool ree_split_supported(int d);
int ree_split_step(const double* x, double* x_new, int64_t n, int d, int64_t ld, double t, const double* m_noise,
                   const double* F, int tri, const double* z, uint64_t seed, uint64_t draw, int64_t first, int lsf_mode,
                   const double* coef, double offset, double* g_new, double* e_new, cudaStream_t st);
int ree_split_maha(const double* x, int64_t n, int d, int64_t ld, const double* g, const double* e, const double* mean,
                   const double* Linv, const double* logdet, double* logq, double* ws_small, double* is4,
                   cudaStream_t st);
int ree_split_moments(const double* x, int64_t n, int d, int64_t ld, const double* g, const double* a_vec,
                      const double* wmax, const double* shift, double* ws_big, double* ws_small, double* sums_u,
                      double* sums_w, cudaStream_t st);

#define TP_WIDE 128  // particles per CTA tile in the contraction kernels (d <= 200: the tile fits 200 KB)
#define TP_NARROW 64 // ... for 200 < d <= 256
#define JB 4         // output dims per register block

// workspace layout (doubles): [0,8) ticket header | [8, 8+4*G) small partials | [WS_BIG, ...) big partials
#define REE_WS_HEADER 8
static inline size_t ws_big_offset() { return REE_WS_HEADER + 4 * (size_t)ree_grid_ctas() * 4; }

// ---------------------------------------------------------------------------------------
// contraction kernel.  MODE 0: Euler-Maruyama transform, MODE 1: Mahalanobis + IS + weights.
//   out_j = sum_k A[j][k] * tile[k][p]         (A = F or Linv, row-major d x d)
// tile[k][p] = z_k (noise) or x_k - mean_k.
// ---------------------------------------------------------------------------------------
struct StepArgs {
    const double* x;      // input ensemble (SoA)
    double* x_new;        // MODE 0 output
    int64_t n, ld;
    int d;
    double t;
    const double* m_noise;
    const double* A;      // F (MODE 0) / Linv (MODE 1)
    const double* z;      // injected noise or NULL
    uint64_t seed, draw;
    int64_t first;
    double* e_new;        // MODE 0: energy of x'
    // MODE 1
    const double* g;
    const double* e;
    const double* mean;
    const double* logdet;
    const double* wmax;
    double sigma, beta;
    int tgt_kind;
    double* logq;
    double* w;
    double* ws;           // workspace (small partials + ticket)
    double* is4;
};

template <int MODE, int TP>
__global__ void __launch_bounds__(TP) k_contract(StepArgs a) {
    extern __shared__ double tile[];  // [d][TP]
    __shared__ ExpSums sh[32];
    __shared__ bool is_last;
    const int d = a.d, tid = threadIdx.x;
    ExpSums is_acc = expsums_empty();
    for (int64_t base = (int64_t)blockIdx.x * TP; base < a.n; base += (int64_t)gridDim.x * TP) {
        const int64_t i = base + tid;
        const bool live = i < a.n;
        __syncthreads();
        if (MODE == 0) {
            if (a.z) {
                for (int k = 0; k < d; ++k) tile[k * TP + tid] = live ? a.z[(int64_t)k * a.ld + i] : 0.0;
            } else {
                const int npairs = ((d + 7) / 8) * 4;
                for (int p = 0; p < npairs; ++p) {
                    double z0, z1;
                    philox_normal_pair(a.seed, (uint64_t)(a.first + i), (uint32_t)p, a.draw, z0, z1);
                    int j0 = (p >> 2) * 8 + (p & 3), j1 = j0 + 4;
                    if (j0 < d) tile[j0 * TP + tid] = z0;
                    if (j1 < d) tile[j1 * TP + tid] = z1;
                }
            }
        } else {
            for (int k = 0; k < d; ++k)
                tile[k * TP + tid] = live ? a.x[(int64_t)k * a.ld + i] - __ldg(a.mean + k) : 0.0;
        }
        __syncthreads();
        double q = 0.0;  // MODE 0: |x'|^2 ; MODE 1: |Linv y|^2
        for (int j0 = 0; j0 < d; j0 += JB) {
            double acc[JB];
#pragma unroll
            for (int r = 0; r < JB; ++r) acc[r] = 0.0;
            const int kmax = (MODE == 1) ? min(d, j0 + JB) : d;  // Linv is lower triangular
            for (int k = 0; k < kmax; ++k) {
                double v = tile[k * TP + tid];
#pragma unroll
                for (int r = 0; r < JB; ++r) {
                    int j = j0 + r;
                    double c = (j < d) ? __ldg(a.A + (size_t)j * d + k) : 0.0;
                    acc[r] = fma(c, v, acc[r]);
                }
            }
#pragma unroll
            for (int r = 0; r < JB; ++r) {
                int j = j0 + r;
                if (j < d) {
                    if (MODE == 0) {
                        // t*X + (m_noise + Z F^T)   -- grouping of solver.py:667
                        double xn = 0.0;
                        if (live) {
                            xn = __dadd_rn(__dmul_rn(a.t, a.x[(int64_t)j * a.ld + i]),
                                           __dadd_rn(__ldg(a.m_noise + j), acc[r]));
                            a.x_new[(int64_t)j * a.ld + i] = xn;
                        }
                        q = fma(xn, xn, q);
                    } else {
                        q = fma(acc[r], acc[r], q);
                    }
                }
            }
        }
        if (MODE == 0) {
            if (live && a.e_new) a.e_new[i] = 0.5 * ((double)d * REE_LOG_2PI + q);
        } else if (live) {
            // logpdf of N(mean, cov):  -0.5*(d log 2pi + log det + maha)     (scipy _logpdf)
            double lq = -0.5 * ((double)d * REE_LOG_2PI + a.logdet[0] + q);
            if (a.logq) a.logq[i] = lq;
            double gi = a.g[i], ei = a.e[i];
            double r = __dsub_rn(-ei, lq);  // -e_fun - log_pdf  (solver.py:684)
            if (isfinite(r)) expsums_add(is_acc, r, gi <= 0.0 ? 1.0 : 0.0);
            if (a.w) {
                double al = __dmul_rn(ree_logtgt(gi, ei, a.sigma, a.tgt_kind), a.beta);
                a.w[i] = isfinite(al) ? exp(al - a.wmax[0]) : 0.0;
            }
        }
    }
    if (MODE == 1) {
        ExpSums blk = expsums_block_reduce(is_acc, sh);
        double* part = a.ws + REE_WS_HEADER;
        unsigned int* ticket = reinterpret_cast<unsigned int*>(a.ws);
        if (tid == 0) {
            double* p = part + 4 * (size_t)blockIdx.x;
            p[0] = blk.m; p[1] = blk.s1; p[2] = blk.s2; p[3] = blk.cnt;
            __threadfence();
            unsigned int t = atomicInc(ticket, gridDim.x - 1);
            is_last = (t == gridDim.x - 1);
        }
        __syncthreads();
        if (!is_last) return;
        __threadfence();
        ExpSums v = expsums_empty();
        for (unsigned int b = tid; b < gridDim.x; b += blockDim.x) {
            const volatile double* p = part + 4 * (size_t)b;
            ExpSums o{p[0], p[1], p[2], p[3]};
            v = expsums_merge(v, o);
        }
        v = expsums_block_reduce(v, sh);
        if (tid == 0) {
            a.is4[0] = v.m; a.is4[1] = v.s1; a.is4[2] = v.s2; a.is4[3] = v.cnt;
        }
    }
}

// ---------------------------------------------------------------------------------------
// Gram kernel: per CTA partial of  S1_j = sum_p w_p y_pj ,  S2_jk = sum_p w_p y_pj y_pk ,
// y = x - shift, plus two scalars (unweighted: #{g<=0}, n ; weighted: sum w, sum w^2).
// Each thread owns MAXO entries of S2 in registers for the whole sweep.
// ---------------------------------------------------------------------------------------
#define GP 32  // particles per shared-memory tile
#define GT 256

// blockIdx.y selects a slab of GT*MAXO entries of S2 (d > 160: d*d entries do not fit one CTA's registers); slab 0
// also owns S1 and the scalars.  Every slab stages the same tiles.
template <int MAXO>
__global__ void __launch_bounds__(GT) k_gram(const double* __restrict__ x, int64_t n, int d, int64_t ld,
                                             const double* __restrict__ shift, const double* __restrict__ w,
                                             const double* __restrict__ g, double* __restrict__ partials) {
    extern __shared__ double ys[];  // [d][GP+1] then wsm[GP]
    const int lds = GP + 1;
    double* wsm = ys + (size_t)d * lds;
    __shared__ double red[32];
    const int tid = threadIdx.x;
    double acc[MAXO];
    unsigned int ojk[MAXO];  // (row offset of j) | (row offset of k) << 16
#pragma unroll
    for (int r = 0; r < MAXO; ++r) {
        acc[r] = 0.0;
        int o = (int)blockIdx.y * GT * MAXO + tid + r * GT;
        int j = o / d, k = o - j * d;
        bool valid = o < d * d;
        ojk[r] = valid ? ((unsigned)(j * lds) | ((unsigned)(k * lds) << 16)) : 0u;
    }
    double s1 = 0.0;            // thread tid < d owns S1[tid]
    double sa = 0.0, sb = 0.0;  // scalars
    const int64_t ntiles = (n + GP - 1) / GP;
    for (int64_t tix = blockIdx.x; tix < ntiles; tix += gridDim.x) {
        const int64_t base = tix * GP;
        __syncthreads();
        for (int idx = tid; idx < d * GP; idx += GT) {
            int j = idx / GP, p = idx - j * GP;
            int64_t i = base + p;
            ys[j * lds + p] = (i < n) ? x[(int64_t)j * ld + i] - (shift ? __ldg(shift + j) : 0.0) : 0.0;
        }
        if (tid < GP) {
            int64_t i = base + tid;
            double wi = (i < n) ? (w ? w[i] : 1.0) : 0.0;
            wsm[tid] = wi;
            if (blockIdx.y != 0) {  // the scalars belong to slab 0
            } else if (w) {
                sa += wi;
                sb += wi * wi;
            } else if (i < n) {
                sa += (g && g[i] <= 0.0) ? 1.0 : 0.0;
                sb += 1.0;
            }
        }
        __syncthreads();
        for (int p = 0; p < GP; ++p) {
            const double wp = wsm[p];
#pragma unroll
            for (int r = 0; r < MAXO; ++r)
                acc[r] = fma(wp, ys[(ojk[r] & 0xffffu) + p] * ys[(ojk[r] >> 16) + p], acc[r]);
        }
        if (blockIdx.y == 0 && tid < d) {  // (d <= GT: one row per thread)
            for (int p = 0; p < GP; ++p) s1 = fma(wsm[p], ys[tid * lds + p], s1);
        }
    }
    const size_t len = (size_t)d + (size_t)d * d + 2;
    double* out = partials + len * blockIdx.x;
    if (blockIdx.y == 0 && tid < d) out[tid] = s1;
#pragma unroll
    for (int r = 0; r < MAXO; ++r) {
        int o = (int)blockIdx.y * GT * MAXO + tid + r * GT;
        if (o < d * d) out[d + o] = acc[r];
    }
    sa = block_sum(sa, red);
    sb = block_sum(sb, red);
    if (tid == 0 && blockIdx.y == 0) {
        out[d + d * d] = sa;
        out[d + d * d + 1] = sb;
    }
}

// second stage: sums[o] = sum over CTAs (fixed order)
__global__ void __launch_bounds__(256) k_reduce_partials(const double* __restrict__ partials, int nparts, size_t len,
                                                         double* __restrict__ sums) {
    size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= len) return;
    double s = 0.0;
    for (int b = 0; b < nparts; ++b) s += partials[len * b + o];
    sums[o] = s;
}

static int launch_gram(const double* x, int64_t n, int d, int64_t ld, const double* shift, const double* w,
                       const double* g, double* workspace, double* sums, cudaStream_t st) {
    const size_t len = (size_t)d + (size_t)d * d + 2;
    double* partials = workspace + ws_big_offset();
    int64_t ntiles = (n + GP - 1) / GP;
    int grid = ree_grid_ctas();
    if (ntiles < grid) grid = (int)(ntiles < 1 ? 1 : ntiles);
    size_t smem = ((size_t)d * (GP + 1) + GP) * sizeof(double);
    int maxo = (d * d + GT - 1) / GT;
    if (d > GT) return ree_set_error(REE_E_UNSUPPORTED, "gram: d > 256 not supported");
#define LAUNCH_GRAM(M)                                                                                  \
    do {                                                                                                \
        const int slabs = (maxo + M - 1) / M;                                                           \
        if (grid * slabs > 2 * ree_grid_ctas()) grid = 2 * ree_grid_ctas() / slabs;                     \
        REE_OPT_IN_SMEM(k_gram<M>, 72 * 1024); /* d <= 256: (33 d + 32) doubles */                     \
        k_gram<M><<<dim3(grid, slabs), GT, smem, st>>>(x, n, d, ld, shift, w, g, partials);             \
    } while (0)
    if (maxo <= 1) LAUNCH_GRAM(1);
    else if (maxo <= 4) LAUNCH_GRAM(4);
    else if (maxo <= 16) LAUNCH_GRAM(16);
    else if (maxo <= 40) LAUNCH_GRAM(40);
    else if (maxo <= 64) LAUNCH_GRAM(64);
    else if (maxo <= 100) LAUNCH_GRAM(100);
    else LAUNCH_GRAM(64);  // 160 < d <= 256: up to four slabs of 64 * 256 entries
#undef LAUNCH_GRAM
    int rc = ree_check_launch("k_gram");
    if (rc) return rc;
    k_reduce_partials<<<(unsigned)((len + 255) / 256), 256, 0, st>>>(partials, grid, len, sums);
    return ree_check_launch("k_reduce_partials");
}

template <int MODE, int TP>
static int launch_contract_tp(const StepArgs& a, cudaStream_t st) {
    const size_t smem = (size_t)a.d * TP * sizeof(double);
    if (smem > 200 * 1024) return ree_set_error(REE_E_UNSUPPORTED, "modular contraction: d too large (d <= 256)");
    int64_t want = (a.n + TP - 1) / TP;
    const int cap = ree_grid_ctas();
    const int grid = (int)(want < 1 ? 1 : (want > cap ? cap : want));
    static bool attr_done[64] = {false};  // the opt-in is per function and device, not per launch
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_done[dev]) {
        cudaFuncSetAttribute(k_contract<MODE, TP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_done[dev] = true;
    }
    k_contract<MODE, TP><<<grid, TP, smem, st>>>(a);
    return ree_check_launch(MODE == 0 ? "ree_cbree_step(transform)" : "ree_cbree_maha(contract)");
}

template <int MODE>
static int launch_contract(const StepArgs& a, cudaStream_t st) {
    if ((size_t)a.d * TP_WIDE * sizeof(double) <= 200 * 1024) return launch_contract_tp<MODE, TP_WIDE>(a, st);
    return launch_contract_tp<MODE, TP_NARROW>(a, st);
}

// ---------------------------------------------------------------------------------------
extern "C" int64_t ree_sums_len(int d) { return (int64_t)d + (int64_t)d * d + 2; }

extern "C" int64_t ree_workspace_bytes(int d) {
    size_t len = (size_t)d + (size_t)d * d + 2;
    size_t big = len * (size_t)ree_grid_ctas();
    size_t mma = ree_mma_partial_doubles(d);
    return (int64_t)((ws_big_offset() + (big > mma ? big : mma)) * sizeof(double));
}

extern "C" int ree_ensemble_sums(const double* x, int64_t n, int d, int64_t ld, const double* g, const double* shift,
                                 double* workspace, double* sums, void* stream) {
    if (!x || !workspace || !sums || d < 1 || n < 1) return ree_set_error(REE_E_BADARG, "ree_ensemble_sums: bad argument");
    if (ree_small_supported(d))
        return ree_small_post(x, n, d, ld, g, nullptr, shift, nullptr, nullptr, 0.0, 0.0, 0, nullptr, nullptr, nullptr,
                              workspace + ws_big_offset(), workspace + REE_WS_HEADER, sums, nullptr, ree_stream(stream));
    if (ree_mma_supported(d))
        return ree_mma_sums(x, n, d, ld, g, shift, workspace + ws_big_offset(), workspace + REE_WS_HEADER, sums,
                            ree_stream(stream));
    return launch_gram(x, n, d, ld, shift, nullptr, g, workspace, sums, ree_stream(stream));
}

extern "C" int ree_cbree_moments(const double* x, int64_t n, int d, int64_t ld, const double* g, const double* a,
                                 const double* wmax, const double* shift, double* workspace, double* sums_u,
                                 double* sums_w, void* stream) {
    if (!x || !workspace || !sums_u || d < 1 || n < 1 || ld < n || (sums_w && (!a || !wmax)))
        return ree_set_error(REE_E_BADARG, "ree_cbree_moments: bad argument");
    if (!ree_split_supported(d))
        return ree_set_error(REE_E_UNSUPPORTED, "ree_cbree_moments: only where ree_fused_path(d) == 2");
    return ree_split_moments(x, n, d, ld, g, a, wmax, shift, workspace + ws_big_offset(), workspace + REE_WS_HEADER,
                             sums_u, sums_w, ree_stream(stream));
}

extern "C" int ree_cbree_step(const double* x, double* x_new, int64_t n, int d, int64_t ld, double t,
                              const double* m_noise, const double* F, int noise_mode, const double* z, uint64_t seed,
                              uint64_t draw, int64_t first_particle, const ree_lsf* lsf, const double* shift,
                              double* g_new, double* e_new, double* workspace, double* sums, void* stream) {
    if (!x || !x_new || !m_noise || !F || !workspace || d < 1 || n < 1 || ld < n)
        return ree_set_error(REE_E_BADARG, "ree_cbree_step: bad argument");
    if (!sums && !ree_split_supported(d))
        return ree_set_error(REE_E_BADARG, "ree_cbree_step: sums may be NULL only where ree_fused_path(d) == 2");
    if (noise_mode == REE_NOISE_INJECT && !z) return ree_set_error(REE_E_BADARG, "ree_cbree_step: INJECT needs z");
    cudaStream_t st = ree_stream(stream);
    const bool fused_lsf = lsf && lsf->kind == REE_LSF_AFFINE;
    if (lsf && lsf->kind != REE_LSF_EXTERNAL && !g_new) return ree_set_error(REE_E_BADARG, "ree_cbree_step: g_new is NULL");
    if (!sums && noise_mode == REE_NOISE_INJECT && d > 103) {
        // parity-mode (dense) factor beyond the DMMA transform's shared-memory budget: modular transform, no Gram
        StepArgs a{};
        a.x = x; a.x_new = x_new; a.n = n; a.ld = ld; a.d = d; a.t = t; a.m_noise = m_noise; a.A = F; a.z = z;
        a.seed = seed; a.draw = draw; a.first = first_particle; a.e_new = e_new;
        int rc = launch_contract<0>(a, st);
        if (rc) return rc;
        if (lsf && lsf->kind != REE_LSF_EXTERNAL) return ree_lsf_launch(lsf, x_new, n, ld, g_new, st);
        return 0;
    }
    if (!sums) {  // split path: transform only
        const int tri = (noise_mode == REE_NOISE_PHILOX) ? 1 : 0;
        int lsf_mode = fused_lsf ? (lsf->coef ? 1 : 2) : 0;
        int rc = ree_split_step(x, x_new, n, d, ld, t, m_noise, F, tri, noise_mode == REE_NOISE_INJECT ? z : nullptr, seed,
                                draw, first_particle, lsf_mode, fused_lsf ? lsf->coef : nullptr,
                                fused_lsf ? lsf->offset : 0.0, g_new, e_new, st);
        if (rc) return rc;
        if (lsf && lsf->kind != REE_LSF_EXTERNAL && !fused_lsf) return ree_lsf_launch(lsf, x_new, n, ld, g_new, st);
        return 0;
    }
    if (ree_small_supported(d)) {  // thread-per-particle sweep; every closed-form LSF is evaluated in registers
        int kind = 0;
        if (lsf && lsf->kind >= REE_LSF_AFFINE && lsf->kind <= REE_LSF_OSCILLATOR) {
            kind = (lsf->kind == REE_LSF_AFFINE && !lsf->coef) ? 100 /* -sum(x)/sqrt(d) + offset */ : lsf->kind;
            if ((lsf->kind == REE_LSF_OSCILLATOR && d != 6) || (lsf->kind == REE_LSF_CONVEX && d < 2)) kind = 0;
        }
        int rc = ree_small_step(x, x_new, n, d, ld, t, m_noise, F, noise_mode == REE_NOISE_PHILOX ? 1 : 0,
                                noise_mode == REE_NOISE_INJECT ? z : nullptr, seed, draw, first_particle, kind,
                                lsf ? lsf->coef : nullptr, lsf ? lsf->offset : 0.0, shift, g_new, e_new,
                                workspace + ws_big_offset(), sums, st);
        if (rc) return rc;
        if (lsf && lsf->kind != REE_LSF_EXTERNAL && !kind) {
            rc = ree_lsf_launch(lsf, x_new, n, ld, g_new, st);
            if (rc) return rc;
            return ree_count_failed(g_new, n, workspace, sums + d + (size_t)d * d, stream);
        }
        return 0;
    }
    if (ree_mma_supported(d)) {
        // a Cholesky factor is lower triangular; the parity-mode factor U*sqrt(S) is dense
        const int tri = (noise_mode == REE_NOISE_PHILOX) ? 1 : 0;
        int lsf_mode = fused_lsf ? (lsf->coef ? 1 : 2) : 0;
        int rc = ree_mma_step(x, x_new, n, d, ld, t, m_noise, F, tri, noise_mode == REE_NOISE_INJECT ? z : nullptr, seed,
                              draw, first_particle, lsf_mode, fused_lsf ? lsf->coef : nullptr,
                              fused_lsf ? lsf->offset : 0.0, shift, g_new, e_new, workspace + ws_big_offset(),
                              workspace + REE_WS_HEADER, sums, st);
        if (rc) return rc;
        if (lsf && lsf->kind != REE_LSF_EXTERNAL && !fused_lsf) {
            rc = ree_lsf_launch(lsf, x_new, n, ld, g_new, st);
            if (rc) return rc;
            return ree_count_failed(g_new, n, workspace, sums + d + (size_t)d * d, stream);
        }
        return 0;
    }
    StepArgs a{};
    a.x = x; a.x_new = x_new; a.n = n; a.ld = ld; a.d = d; a.t = t; a.m_noise = m_noise; a.A = F;
    a.z = (noise_mode == REE_NOISE_INJECT) ? z : nullptr;
    a.seed = seed; a.draw = draw; a.first = first_particle; a.e_new = e_new;
    int rc = launch_contract<0>(a, st);
    if (rc) return rc;
    const double* g_for_count = nullptr;
    if (lsf && lsf->kind != REE_LSF_EXTERNAL) {
        if (!g_new) return ree_set_error(REE_E_BADARG, "ree_cbree_step: g_new is NULL");
        rc = ree_lsf_launch(lsf, x_new, n, ld, g_new, st);
        if (rc) return rc;
        g_for_count = g_new;
    }
    return launch_gram(x_new, n, d, ld, shift, nullptr, g_for_count, workspace, sums, st);
}

extern "C" int ree_cbree_maha(const double* x, int64_t n, int d, int64_t ld, const double* g, const double* e,
                              const double* mean, const double* Linv, const double* logdet, double sigma, double beta,
                              int tgt_kind, const double* wmax, double* logq, double* w, double* workspace,
                              double* sums, double* is4, void* stream) {
    if (!x || !g || !e || !mean || !Linv || !logdet || !workspace || !is4 || d < 1 || n < 1)
        return ree_set_error(REE_E_BADARG, "ree_cbree_maha: bad argument");
    cudaStream_t st = ree_stream(stream);
    if (!sums) {
        if (!ree_split_supported(d))
            return ree_set_error(REE_E_BADARG, "ree_cbree_maha: sums may be NULL only where ree_fused_path(d) == 2");
        return ree_split_maha(x, n, d, ld, g, e, mean, Linv, logdet, logq, workspace + REE_WS_HEADER, is4, st);
    }
    if (!wmax) return ree_set_error(REE_E_BADARG, "ree_cbree_maha: wmax is NULL");
    if (ree_small_supported(d))
        return ree_small_post(x, n, d, ld, g, e, mean, Linv, logdet, sigma, beta, tgt_kind, wmax, logq, w,
                              workspace + ws_big_offset(), workspace + REE_WS_HEADER, sums, is4, st);
    if (ree_mma_supported(d))
        return ree_mma_maha(x, n, d, ld, g, e, mean, Linv, logdet, sigma, beta, tgt_kind, wmax, logq, w,
                            workspace + ws_big_offset(), workspace + REE_WS_HEADER, sums, is4, st);
    if (!w) return ree_set_error(REE_E_BADARG, "ree_cbree_maha: this dimension needs the weight scratch vector w");
    StepArgs a{};
    a.x = x; a.n = n; a.ld = ld; a.d = d; a.A = Linv; a.g = g; a.e = e; a.mean = mean; a.logdet = logdet;
    a.wmax = wmax; a.sigma = sigma; a.beta = beta; a.tgt_kind = tgt_kind; a.logq = logq; a.w = w;
    a.ws = workspace; a.is4 = is4;
    int rc = launch_contract<1>(a, st);
    if (rc) return rc;
    return launch_gram(x, n, d, ld, mean, w, nullptr, workspace, sums, st);
}
