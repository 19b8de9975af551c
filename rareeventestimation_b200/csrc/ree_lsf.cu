// Limit-state functions, energy, layout transposes and the Philox sample generator.
// Thread-per-particle kernels on the SoA ensemble (coalesced 8-byte loads along the particle axis).
#include "common.cuh"
#include "lsf_point.cuh"


// ---------------------------------------------------------------------------------------
// toy LSFs (problem/toy_problems.py).  Operation order mirrors the NumPy expressions; _rn
// intrinsics keep nvcc from contracting mul+add into FMA (NumPy does not fuse).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(REE_THREADS) k_lsf_simple(ree_lsf lsf, const double* __restrict__ x, int64_t n,
                                                            int64_t ld, double* __restrict__ g) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int d = lsf.d;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double out;
        switch (lsf.kind) {
            case REE_LSF_AFFINE: {  // g = offset + sum_j coef_j x_j
                double s = 0.0;
                for (int j = 0; j < d; ++j) s = __dadd_rn(s, __dmul_rn(__ldg(lsf.coef + j), x[(int64_t)j * ld + i]));
                out = __dadd_rn(s, lsf.offset);
                break;
            }
            case REE_LSF_LINEAR_SUM_INTERNAL: {  // -sum(x)/sqrt(d) + beta, toy_problems.py:36-37
                double s = 0.0;
                for (int j = 0; j < d; ++j) s = __dadd_rn(s, x[(int64_t)j * ld + i]);
                out = __dadd_rn(__ddiv_rn(-s, __dsqrt_rn((double)d)), lsf.offset);
                break;
            }
            case REE_LSF_CONVEX: {  // 0.1*(x0-x1)**2 - 1/sqrt(2)*sum(x) + 2.5, toy_problems.py:12-13
                double x0 = x[i], x1 = x[ld + i];
                double df = __dsub_rn(x0, x1);
                double a = __dmul_rn(0.1, __dmul_rn(df, df));
                double s = __dadd_rn(x0, x1);
                for (int j = 2; j < d; ++j) s = __dadd_rn(s, x[(int64_t)j * ld + i]);
                double b = __dmul_rn(__ddiv_rn(1.0, __dsqrt_rn(2.0)), s);
                out = __dadd_rn(__dsub_rn(a, b), 2.5);
                break;
            }
            case REE_LSF_FUJITA_RACKWITZ: {  // -sum(log(Phi(-x))) - c, toy_problems.py:57-58
                double s = 0.0;
                for (int j = 0; j < d; ++j) s = __dadd_rn(s, log(normcdf(-x[(int64_t)j * ld + i])));
                out = __dsub_rn(-s, lsf.offset);
                break;
            }
            case REE_LSF_OSCILLATOR: {
                out = lsf_oscillator(x[i], x[ld + i], x[2 * ld + i], x[3 * ld + i], x[4 * ld + i], x[5 * ld + i]);
                break;
            }
            default:
                out = CUDART_NAN;
        }
        g[i] = out;
    }
}

// ---------------------------------------------------------------------------------------
// 1-D diffusion PDE LSF (problem/diffusion.py:131-171, 232-242), v0: thread per particle.
// theta tile staged in shared memory ([k][tid]), the mode table row is a warp-uniform
// (broadcast) read-only load.  Per element e: kl_q = exp(mu + sigma * <T[3e+q,:], theta>),
// Ke = sum_q wGL_q kl_q / (2h); the tridiagonal system is eliminated on the fly (Thomas forward
// sweep; only u(1) = rhs'_last / diag'_last is needed).
// ---------------------------------------------------------------------------------------
#define REE_PDE_TPB 128

__global__ void __launch_bounds__(REE_PDE_TPB) k_lsf_diffusion(ree_lsf lsf, const double* __restrict__ x, int64_t n,
                                                               int64_t ld, double* __restrict__ g) {
    extern __shared__ double th[];  // [d][REE_PDE_TPB]
    const int d = lsf.d, ne = lsf.n_ele;
    const double h = 1.0 / (double)ne;
    const double inv2h = 1.0 / h / 2.0;  // 1/h/2 as in diffusion.py:146
    const double w0 = 5.0 / 9.0, w1 = 8.0 / 9.0;
    for (int64_t base = (int64_t)blockIdx.x * REE_PDE_TPB; base < n; base += (int64_t)gridDim.x * REE_PDE_TPB) {
        int64_t i = base + threadIdx.x;
        bool live = i < n;
        __syncthreads();
        for (int k = 0; k < d; ++k) th[k * REE_PDE_TPB + threadIdx.x] = live ? x[(int64_t)k * ld + i] : 0.0;
        __syncthreads();
        double ke_prev = 0.0, dp = 0.0, rp = 0.0;
        for (int e = 0; e <= ne; ++e) {
            double ke = 0.0;
            if (e < ne) {
                const double* row = lsf.coef + (size_t)(3 * e) * d;
                double a0 = 0.0, a1 = 0.0, a2 = 0.0;
                for (int k = 0; k < d; ++k) {
                    double t = th[k * REE_PDE_TPB + threadIdx.x];
                    a0 = fma(__ldg(row + k), t, a0);
                    a1 = fma(__ldg(row + d + k), t, a1);
                    a2 = fma(__ldg(row + 2 * d + k), t, a2);
                }
                double k0 = exp(lsf.p0 + lsf.p1 * a0) * inv2h;
                double k1 = exp(lsf.p0 + lsf.p1 * a1) * inv2h;
                double k2 = exp(lsf.p0 + lsf.p1 * a2) * inv2h;
                ke = k0 * w0 + k1 * w1 + k2 * w0;
            }
            // unknown row r = e-1 (vertex e): diag = Ke[e-1] + Ke[e] (Ke[ne] := 0), off to row r-1 = -Ke[e-1]
            if (e >= 1) {
                int r = e - 1;
                double diag = ke_prev + ke;
                double rhs = (r == ne - 1) ? 0.5 * h : h;
                if (r == 0) {
                    dp = diag;
                    rp = rhs;
                } else {
                    double off = -ke_prev;
                    double wgt = off / dp;
                    dp = diag - wgt * off;
                    rp = rhs - wgt * rp;
                }
            }
            ke_prev = ke;
        }
        if (live) g[i] = lsf.offset - rp / dp;
    }
}

// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(REE_THREADS) k_energy(const double* __restrict__ x, int64_t n, int d, int64_t ld,
                                                        double* __restrict__ e) {
    // 0.5*(d*log(2 pi) + |x|^2)  == -multivariate_normal.logpdf(x, 0, I)   utilities.py:70-73
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const double c = (double)d * REE_LOG_2PI;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double s = 0.0;
        for (int j = 0; j < d; ++j) {
            double v = x[(int64_t)j * ld + i];
            s = fma(v, v, s);
        }
        e[i] = 0.5 * (c + s);
    }
}

// ---------------------------------------------------------------------------------------
// layout: (N,d) C-order <-> SoA [d][ld] through a 32x32 shared-memory transpose
// ---------------------------------------------------------------------------------------
__global__ void k_aos_to_soa(const double* __restrict__ aos, double* __restrict__ soa, int64_t n, int d, int64_t ld) {
    __shared__ double tile[32][33];
    int64_t i0 = (int64_t)blockIdx.x * 32;
    int j0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int64_t i = i0 + r;
        int j = j0 + threadIdx.x;
        if (i < n && j < d) tile[r][threadIdx.x] = aos[i * d + j];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int j = j0 + r;
        int64_t i = i0 + threadIdx.x;
        if (i < n && j < d) soa[(int64_t)j * ld + i] = tile[threadIdx.x][r];
    }
}

__global__ void k_soa_to_aos(const double* __restrict__ soa, double* __restrict__ aos, int64_t n, int d, int64_t ld) {
    __shared__ double tile[32][33];
    int64_t i0 = (int64_t)blockIdx.x * 32;
    int j0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int j = j0 + r;
        int64_t i = i0 + threadIdx.x;
        if (i < n && j < d) tile[r][threadIdx.x] = soa[(int64_t)j * ld + i];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int64_t i = i0 + r;
        int j = j0 + threadIdx.x;
        if (i < n && j < d) aos[i * d + j] = tile[threadIdx.x][r];
    }
}

// Z[j][i] ~ N(0,1): one thread per (particle, dim pair); writes both slots of the pair.
__global__ void __launch_bounds__(REE_THREADS) k_philox_normal(double* __restrict__ z, int64_t n, int d, int64_t ld,
                                                               uint64_t seed, uint64_t draw, int64_t first) {
    int npairs = ((d + 7) / 8) * 4;
    int pair = blockIdx.y;
    if (pair >= npairs) return;
    int j0 = (pair >> 2) * 8 + (pair & 3), j1 = j0 + 4;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double z0, z1;
        philox_normal_pair(seed, (uint64_t)(first + i), (uint32_t)pair, draw, z0, z1);
        if (j0 < d) z[(int64_t)j0 * ld + i] = z0;
        if (j1 < d) z[(int64_t)j1 * ld + i] = z1;
    }
}

// ---------------------------------------------------------------------------------------
static int lsf_grid(int64_t n, int tpb) {
    int64_t want = (n + tpb - 1) / tpb;
    int cap = ree_grid_ctas() * 4;
    return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

bool ree_pde_mma_supported(const ree_lsf* l);
int ree_pde_mma_launch(const ree_lsf* l, const double* x, int64_t n, int64_t ld, double* g, cudaStream_t st);

static bool coef_is_linear_sum(const ree_lsf* lsf) { return lsf->kind == REE_LSF_AFFINE && lsf->coef == nullptr; }

int ree_lsf_launch(const ree_lsf* lsf, const double* x, int64_t n, int64_t ld, double* g, cudaStream_t st) {
    if (!lsf || !x || !g) return ree_set_error(REE_E_BADARG, "ree_lsf_eval: null pointer");
    if (n == 0) return 0;
    ree_lsf l = *lsf;
    switch (l.kind) {
        case REE_LSF_AFFINE:
            if (coef_is_linear_sum(lsf)) l.kind = REE_LSF_LINEAR_SUM_INTERNAL;
            break;
        case REE_LSF_CONVEX:
            if (l.d < 2) return ree_set_error(REE_E_BADARG, "convex LSF needs d >= 2");
            break;
        case REE_LSF_FUJITA_RACKWITZ:
            break;
        case REE_LSF_OSCILLATOR:
            if (l.d != 6) return ree_set_error(REE_E_BADARG, "oscillator LSF needs d == 6");
            break;
        case REE_LSF_DIFFUSION: {
            if (!l.coef || l.n_ele < 1) return ree_set_error(REE_E_BADARG, "diffusion LSF needs a mode table");
            if (ree_pde_mma_supported(&l)) return ree_pde_mma_launch(&l, x, n, ld, g, st);
            size_t smem = (size_t)l.d * REE_PDE_TPB * sizeof(double);
            if (smem > 200 * 1024) return ree_set_error(REE_E_UNSUPPORTED, "diffusion LSF: d too large for shared memory");
            REE_OPT_IN_SMEM(k_lsf_diffusion, 200 * 1024);
            int grid = lsf_grid(n, REE_PDE_TPB);
            k_lsf_diffusion<<<grid, REE_PDE_TPB, smem, st>>>(l, x, n, ld, g);
            return ree_check_launch("ree_lsf_eval(diffusion)");
        }
        default:
            return ree_set_error(REE_E_UNSUPPORTED, "ree_lsf_eval: this LSF kind is evaluated by the caller");
    }
    k_lsf_simple<<<lsf_grid(n, REE_THREADS), REE_THREADS, 0, st>>>(l, x, n, ld, g);
    return ree_check_launch("ree_lsf_eval");
}

extern "C" int ree_lsf_eval(const ree_lsf* lsf, const double* x, int64_t n, int64_t ld, double* g, void* stream) {
    return ree_lsf_launch(lsf, x, n, ld, g, ree_stream(stream));
}

extern "C" int ree_energy(const double* x, int64_t n, int d, int64_t ld, double* e, void* stream) {
    if (!x || !e || d < 1) return ree_set_error(REE_E_BADARG, "ree_energy: bad argument");
    if (n == 0) return 0;
    k_energy<<<lsf_grid(n, REE_THREADS), REE_THREADS, 0, ree_stream(stream)>>>(x, n, d, ld, e);
    return ree_check_launch("ree_energy");
}

extern "C" int ree_aos_to_soa(const double* aos, double* soa, int64_t n, int d, int64_t ld, void* stream) {
    if (!aos || !soa || d < 1 || ld < n) return ree_set_error(REE_E_BADARG, "ree_aos_to_soa: bad argument");
    if (n == 0) return 0;
    dim3 grid((unsigned)((n + 31) / 32), (unsigned)((d + 31) / 32)), block(32, 8);
    k_aos_to_soa<<<grid, block, 0, ree_stream(stream)>>>(aos, soa, n, d, ld);
    return ree_check_launch("ree_aos_to_soa");
}

extern "C" int ree_soa_to_aos(const double* soa, double* aos, int64_t n, int d, int64_t ld, void* stream) {
    if (!aos || !soa || d < 1 || ld < n) return ree_set_error(REE_E_BADARG, "ree_soa_to_aos: bad argument");
    if (n == 0) return 0;
    dim3 grid((unsigned)((n + 31) / 32), (unsigned)((d + 31) / 32)), block(32, 8);
    k_soa_to_aos<<<grid, block, 0, ree_stream(stream)>>>(soa, aos, n, d, ld);
    return ree_check_launch("ree_soa_to_aos");
}

extern "C" int ree_upload_ensemble(const double* host_aos, double* soa, int64_t n, int d, int64_t ld, double* staging,
                                   int64_t staging_elems, void* stream) {
    if (!host_aos || !soa || !staging || d < 1 || staging_elems < d)
        return ree_set_error(REE_E_BADARG, "ree_upload_ensemble: bad argument");
    cudaStream_t st = ree_stream(stream);
    int64_t rows = staging_elems / d;
    for (int64_t i0 = 0; i0 < n; i0 += rows) {
        int64_t m = (n - i0 < rows) ? (n - i0) : rows;
        cudaError_t err = cudaMemcpyAsync(staging, host_aos + i0 * d, (size_t)m * d * sizeof(double),
                                          cudaMemcpyHostToDevice, st);
        if (err != cudaSuccess) return ree_set_error((int)err, cudaGetErrorString(err));
        int rc = ree_aos_to_soa(staging, soa + i0, m, d, ld, stream);
        if (rc) return rc;
        // staging is reused by the next chunk; the copy and the transpose are stream ordered
    }
    cudaError_t err = cudaStreamSynchronize(st);
    return err == cudaSuccess ? 0 : ree_set_error((int)err, cudaGetErrorString(err));
}

extern "C" int ree_download_ensemble(const double* soa, double* host_aos, int64_t n, int d, int64_t ld,
                                     double* staging, int64_t staging_elems, void* stream) {
    if (!host_aos || !soa || !staging || d < 1 || staging_elems < d)
        return ree_set_error(REE_E_BADARG, "ree_download_ensemble: bad argument");
    cudaStream_t st = ree_stream(stream);
    int64_t rows = staging_elems / d;
    for (int64_t i0 = 0; i0 < n; i0 += rows) {
        int64_t m = (n - i0 < rows) ? (n - i0) : rows;
        int rc = ree_soa_to_aos(soa + i0, staging, m, d, ld, stream);
        if (rc) return rc;
        cudaError_t err = cudaMemcpyAsync(host_aos + i0 * d, staging, (size_t)m * d * sizeof(double),
                                          cudaMemcpyDeviceToHost, st);
        if (err != cudaSuccess) return ree_set_error((int)err, cudaGetErrorString(err));
    }
    cudaError_t err = cudaStreamSynchronize(st);
    return err == cudaSuccess ? 0 : ree_set_error((int)err, cudaGetErrorString(err));
}

extern "C" int ree_philox_normal(double* z, int64_t n, int d, int64_t ld, uint64_t seed, uint64_t draw,
                                 int64_t first_particle, void* stream) {
    if (!z || d < 1 || ld < n) return ree_set_error(REE_E_BADARG, "ree_philox_normal: bad argument");
    if (n == 0) return 0;
    int npairs = ((d + 7) / 8) * 4;
    int64_t want = (n + REE_THREADS - 1) / REE_THREADS;
    int cap = ree_grid_ctas() * 2;
    dim3 grid((unsigned)(want > cap ? cap : want), (unsigned)npairs);
    k_philox_normal<<<grid, REE_THREADS, 0, ree_stream(stream)>>>(z, n, d, ld, seed, draw, first_particle);
    return ree_check_launch("ree_philox_normal");
}
