// von Mises-Fisher-Nakagami model with one component: the resampling branch of CBREE
// (solver.py:654-663, 828-836 -> mixturemodel.py:153-239 -> era/EMvMFNM.py).
//
// With a single component the soft EM of EMvMFNM.py degenerates to one M-step over the whole (weighted) sample
// (the responsibilities stay 1, EMvMFNM.py:139-145), so the fit is three reductions over the ensemble:
//     S = sum w x/|x| (d),  sum w |x|^2,  sum w |x|^4,  sum w            EMvMFNM.py:169-199
// Sampling (EMvMFNM.py:274-378): radius sqrt(Gamma(m, omega/m)), cosine w to the mean direction by Wood's rejection
// sampler (Beta((d-1)/2, (d-1)/2) proposal), uniform direction in the orthogonal complement, Householder reflection
// e_d -> mu.  log-density (mixturemodel.py:196-215): vMF(x/|x|) + Nakagami(|x|) - (d-1) log|x|.
//
// All kernels are thread-per-particle on the SoA layout (a warp reads 256 contiguous bytes per dimension); fixed-order
// two-stage reductions.  HBM-bound: 8d bytes per particle for the fit (twice: radii, then the d row dots), 8d written
// by the sampler, 8d read by the density.
#include "common.cuh"

#define VM_T 256

// ---------------------------------------------------------------------------------------
// fit: radii pass and row dots
// ---------------------------------------------------------------------------------------
// rinv_i = w_i / |x_i| (w = 1 without weights); per-CTA partials [sum w r^2, sum w r^4, sum w]
__global__ void __launch_bounds__(VM_T) k_vmfn_radii(const double* __restrict__ x, int64_t n, int d, int64_t ld,
                                                     const double* __restrict__ w, double* __restrict__ rinv,
                                                     double* __restrict__ part3) {
    __shared__ double sh[32];
    double s2 = 0.0, s4 = 0.0, sw = 0.0;
    const int64_t stride = (int64_t)gridDim.x * VM_T;
    for (int64_t i = (int64_t)blockIdx.x * VM_T + threadIdx.x; i < n; i += stride) {
        double r2 = 0.0;
        for (int j = 0; j < d; ++j) {
            const double v = x[(int64_t)j * ld + i];
            r2 = fma(v, v, r2);
        }
        const double wi = w ? w[i] : 1.0;
        rinv[i] = wi / sqrt(r2);
        s2 = fma(wi, r2, s2);
        s4 = fma(wi, r2 * r2, s4);
        sw += wi;
    }
    s2 = block_sum(s2, sh);
    s4 = block_sum(s4, sh);
    sw = block_sum(sw, sh);
    if (threadIdx.x == 0) {
        part3[3 * blockIdx.x + 0] = s2;
        part3[3 * blockIdx.x + 1] = s4;
        part3[3 * blockIdx.x + 2] = sw;
    }
}

// part[j][chunk] = sum over the chunk's particles of x[j][i] * rinv[i]   (grid: chunks x d)
__global__ void __launch_bounds__(VM_T) k_vmfn_rowdots(const double* __restrict__ x, int64_t n, int64_t ld,
                                                       const double* __restrict__ rinv, double* __restrict__ part) {
    __shared__ double sh[32];
    const double* row = x + (int64_t)blockIdx.y * ld;
    double s = 0.0;
    const int64_t stride = (int64_t)gridDim.x * VM_T;
    for (int64_t i = (int64_t)blockIdx.x * VM_T + threadIdx.x; i < n; i += stride) s = fma(row[i], rinv[i], s);
    s = block_sum(s, sh);
    if (threadIdx.x == 0) part[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = s;
}

// sums = [S (d) | sum w r^2 | sum w r^4 | sum w]
__global__ void __launch_bounds__(VM_T) k_vmfn_finish(const double* __restrict__ part, const double* __restrict__ part3,
                                                      int chunks, int ctas, int d, double* __restrict__ sums) {
    for (int j = threadIdx.x; j < d + 3; j += VM_T) {
        double s = 0.0;
        if (j < d)
            for (int c = 0; c < chunks; ++c) s += part[(size_t)j * chunks + c];
        else
            for (int c = 0; c < ctas; ++c) s += part3[3 * c + (j - d)];
        sums[j] = s;
    }
}

// ---------------------------------------------------------------------------------------
// Philox helpers of the sampler: independent sub-streams by the "pair" word of the counter
// ---------------------------------------------------------------------------------------
#define VM_SLOT_SCALAR 0x40000000u  // scalar variates (gamma / beta / uniform attempts) live above the direction's pairs

struct VmRng {
    uint64_t seed, particle, draw;
    uint32_t slot;
    __device__ __forceinline__ void uniforms(double& u0, double& u1) {  // two uniforms in (0, 1)
        uint32_t c[4] = {(uint32_t)particle, (uint32_t)(particle >> 32), slot++, (uint32_t)draw};
        philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        const uint64_t a = (((uint64_t)c[0] << 32) | c[1]) >> 11, b = (((uint64_t)c[2] << 32) | c[3]) >> 11;
        u0 = ((double)a + 0.5) * 1.1102230246251565e-16;
        u1 = ((double)b + 0.5) * 1.1102230246251565e-16;
    }
    __device__ __forceinline__ void normals(double& z0, double& z1) {
        philox_normal_pair(seed, particle, slot++, draw, z0, z1);
    }
    // Gamma(shape, 1), Marsaglia & Tsang (2000); shape < 1 through Gamma(shape + 1) * U^(1/shape)
    __device__ double gamma(double shape) {
        const double a = shape < 1.0 ? shape + 1.0 : shape;
        const double dd = a - 1.0 / 3.0, cc = 1.0 / sqrt(9.0 * dd);
        double out = dd;
        for (int it = 0; it < 64; ++it) {
            double z0, z1, u0, u1;
            normals(z0, z1);
            uniforms(u0, u1);
            const double t = 1.0 + cc * z0;
            if (t <= 0.0) continue;
            const double v = t * t * t;
            if (log(u0) < 0.5 * z0 * z0 + dd - dd * v + dd * log(v)) {
                out = dd * v;
                if (shape < 1.0) out *= pow(u1, 1.0 / shape);
                break;
            }
        }
        return out;
    }
};

struct VmSampleArgs {
    double* x;  // output ensemble, SoA
    int64_t n, ld;
    int d;
    const double* hv;  // Householder vector (d), device; reflection I - hb hv hv^T maps e_d to mu (EMvMFNM.py:361-365)
    double hb;
    double kappa, m, omega;
    uint64_t seed, draw;
    int64_t first;
    // injected variates (parity mode): radius r[n], cosine w[n], raw direction normals vraw[(d-1)][ld]
    const double* r_in;
    const double* w_in;
    const double* v_in;
};

__global__ void __launch_bounds__(VM_T) k_vmfn_sample(const VmSampleArgs a) {
    extern __shared__ double hv_s[];
    for (int j = threadIdx.x; j < a.d; j += VM_T) hv_s[j] = a.hv[j];
    __syncthreads();
    const int d = a.d, dm1 = d - 1;
    // constants of Wood's sampler (EMvMFNM.py:327-333)
    const double t1 = sqrt(4.0 * a.kappa * a.kappa + (double)dm1 * dm1);
    const double b = (-2.0 * a.kappa + t1) / dm1;
    const double x0 = (1.0 - b) / (1.0 + b);
    const double mh = 0.5 * dm1;
    const double c = a.kappa * x0 + dm1 * log(1.0 - x0 * x0);
    const int64_t stride = (int64_t)gridDim.x * VM_T;
    for (int64_t i = (int64_t)blockIdx.x * VM_T + threadIdx.x; i < a.n; i += stride) {
        double r, w;
        VmRng rng{a.seed, (uint64_t)(a.first + i), a.draw, VM_SLOT_SCALAR};
        if (a.r_in) {
            r = a.r_in[i];
            w = a.w_in[i];
        } else {
            r = sqrt(rng.gamma(a.m) * a.omega / a.m);  // sqrt(Gamma(m, scale = omega/m)), EMvMFNM.py:280
            w = 1.0;
            for (int it = 0; it < 256; ++it) {  // EMvMFNM.py:335-346
                const double g1 = rng.gamma(mh), g2 = rng.gamma(mh);
                const double z = g1 / (g1 + g2);  // Beta(mh, mh)
                double u, u_unused;
                rng.uniforms(u, u_unused);
                w = (1.0 - (1.0 + b) * z) / (1.0 - (1.0 - b) * z);
                const double t = a.kappa * w + dm1 * log(1.0 - x0 * w) - c;
                if (t >= log(u)) break;
            }
        }
        // direction in the orthogonal complement: pass 1 accumulates |v|^2 and hv . v, pass 2 rescales and reflects
        double s2 = 0.0, sv = 0.0;
        const int npairs = (dm1 + 1) >> 1;
        for (int p = 0; p < npairs; ++p) {
            double z0, z1;
            if (a.v_in) {
                z0 = a.v_in[(int64_t)(2 * p) * a.ld + i];
                z1 = (2 * p + 1 < dm1) ? a.v_in[(int64_t)(2 * p + 1) * a.ld + i] : 0.0;
            } else {
                // generated once: the raw normals are parked in the output rows and rescaled in place by pass 2
                philox_normal_pair(a.seed, (uint64_t)(a.first + i), (uint32_t)p, a.draw, z0, z1);
                if (2 * p + 1 >= dm1) z1 = 0.0;
                a.x[(int64_t)(2 * p) * a.ld + i] = z0;
                if (2 * p + 1 < dm1) a.x[(int64_t)(2 * p + 1) * a.ld + i] = z1;
            }
            s2 = fma(z0, z0, fma(z1, z1, s2));
            sv = fma(hv_s[2 * p], z0, sv);
            if (2 * p + 1 < dm1) sv = fma(hv_s[2 * p + 1], z1, sv);
        }
        const double scale = sqrt(fmax(1.0 - w * w, 0.0)) / sqrt(s2);  // X[:d-1] = sqrt(1 - w^2) v/|v|
        const double dot = fma(scale, sv, hv_s[dm1] * w);                // hv . (unit vector before the reflection)
        const double f = a.hb * dot;
        for (int p = 0; p < npairs; ++p) {
            double z0, z1;
            if (a.v_in) {
                z0 = a.v_in[(int64_t)(2 * p) * a.ld + i];
                z1 = (2 * p + 1 < dm1) ? a.v_in[(int64_t)(2 * p + 1) * a.ld + i] : 0.0;
            } else {
                z0 = a.x[(int64_t)(2 * p) * a.ld + i];
                z1 = (2 * p + 1 < dm1) ? a.x[(int64_t)(2 * p + 1) * a.ld + i] : 0.0;
            }
            a.x[(int64_t)(2 * p) * a.ld + i] = r * (scale * z0 - f * hv_s[2 * p]);
            if (2 * p + 1 < dm1) a.x[(int64_t)(2 * p + 1) * a.ld + i] = r * (scale * z1 - f * hv_s[2 * p + 1]);
        }
        a.x[(int64_t)dm1 * a.ld + i] = r * (w - f * hv_s[dm1]);
    }
}

// ---------------------------------------------------------------------------------------
// log-density + IS statistics
// ---------------------------------------------------------------------------------------
// logq_i = kappa mu.x/r + c0 - m r^2/omega + (2m - d) log r,   c0 = vMF constant + Nakagami constant (host)
// FUSE: 0 g and e are inputs; 1 / 2: the kernel also evaluates the energy and the affine limit state of the points it
// reads anyway (1: g = offset + coef . x, 2: g = offset - sum(x)/sqrt(d); same operation order as k_lsf_simple / k_energy)
// and writes them to g / e -- the redraw of the resampling branch needs all three for the same fresh ensemble.
template <int FUSE>
__global__ void __launch_bounds__(VM_T) k_vmfn_logpdf(const double* __restrict__ x, int64_t n, int d, int64_t ld,
                                                      const double* __restrict__ mu, double kappa, double m,
                                                      double omega, double c0, double* __restrict__ g,
                                                      double* __restrict__ e, double* __restrict__ logq,
                                                      double* __restrict__ small, const double* __restrict__ coef,
                                                      double offset) {
    extern __shared__ double mu_s[];  // mu (d) | coef (d, FUSE == 1)
    __shared__ ExpSums sh[32];
    for (int j = threadIdx.x; j < d; j += VM_T) {
        mu_s[j] = mu[j];
        if (FUSE == 1) mu_s[d + j] = coef[j];
    }
    __syncthreads();
    const double e_const = (double)d * REE_LOG_2PI;
    ExpSums acc = expsums_empty();
    const int64_t stride = (int64_t)gridDim.x * VM_T;
    for (int64_t i = (int64_t)blockIdx.x * VM_T + threadIdx.x; i < n; i += stride) {
        double r2 = 0.0, dot = 0.0, gs = 0.0;
        for (int j = 0; j < d; ++j) {
            const double v = x[(int64_t)j * ld + i];
            r2 = fma(v, v, r2);
            dot = fma(mu_s[j], v, dot);
            if (FUSE == 1) gs = __dadd_rn(gs, __dmul_rn(mu_s[d + j], v));
            if (FUSE == 2) gs = __dadd_rn(gs, v);
        }
        const double r = sqrt(r2);
        const double lq = kappa * (dot / r) + c0 - m * (r2 / omega) + (2.0 * m - (double)d) * log(r);
        if (logq) logq[i] = lq;
        double gi, ei;
        if (FUSE) {
            gi = (FUSE == 1) ? __dadd_rn(gs, offset) : __dadd_rn(__ddiv_rn(-gs, __dsqrt_rn((double)d)), offset);
            ei = 0.5 * (e_const + r2);
            g[i] = gi;
            e[i] = ei;
        } else if (small) {
            gi = g[i];
            ei = e[i];
        }
        if (small) {
            const double rr = __dsub_rn(-ei, lq);  // -e_fun - log_pdf  (solver.py:684)
            if (isfinite(rr)) expsums_add(acc, rr, gi <= 0.0 ? 1.0 : 0.0);
        }
    }
    if (small) {
        ExpSums blk = expsums_block_reduce(acc, sh);
        if (threadIdx.x == 0) {
            double* p = small + 4 * (size_t)blockIdx.x;
            p[0] = blk.m; p[1] = blk.s1; p[2] = blk.s2; p[3] = blk.cnt;
        }
    }
}

int ree_mma_merge_small(const double* small, int n, double* out4, cudaStream_t st);  // ree_mma.cu

static int vm_grid(int64_t n) {
    const int64_t want = (n + VM_T - 1) / VM_T;
    const int cap = 4 * ree_grid_ctas();
    return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

// ---------------------------------------------------------------------------------------
extern "C" int ree_vmfn_fit_sums(const double* x, int64_t n, int d, int64_t ld, const double* w, double* rinv,
                                 double* workspace, double* sums, void* stream) {
    if (!x || !rinv || !workspace || !sums || d < 1 || n < 1 || ld < n)
        return ree_set_error(REE_E_BADARG, "ree_vmfn_fit_sums: bad argument");
    cudaStream_t st = ree_stream(stream);
    const int ctas = vm_grid(n);
    const int chunks = ctas < 64 ? ctas : 64;
    // workspace layout (ree_workspace_bytes): 8 doubles of ticket header | 16 G doubles of small partials | big partials
    double* part3 = workspace + 8;                                // 3 * ctas <= 12 G
    double* part = workspace + 8 + 16 * (size_t)ree_grid_ctas();  // 64 d doubles <= (d + d^2 + 2) G
    k_vmfn_radii<<<ctas, VM_T, 0, st>>>(x, n, d, ld, w, rinv, part3);
    int rc = ree_check_launch("k_vmfn_radii");
    if (rc) return rc;
    k_vmfn_rowdots<<<dim3((unsigned)chunks, (unsigned)d), VM_T, 0, st>>>(x, n, ld, rinv, part);
    rc = ree_check_launch("k_vmfn_rowdots");
    if (rc) return rc;
    k_vmfn_finish<<<1, VM_T, 0, st>>>(part, part3, chunks, ctas, d, sums);
    return ree_check_launch("k_vmfn_finish");
}

extern "C" int ree_vmfn_sample(double* x, int64_t n, int d, int64_t ld, const double* householder_v, double householder_b,
                               double kappa, double m, double omega, uint64_t seed, uint64_t draw,
                               int64_t first_particle, const double* r_in, const double* w_in, const double* v_in,
                               void* stream) {
    if (!x || !householder_v || d < 2 || n < 0 || ld < n || !(m > 0.0) || !(omega > 0.0) || kappa < 0.0)
        return ree_set_error(REE_E_BADARG, "ree_vmfn_sample: bad argument");
    if ((r_in != nullptr) != (w_in != nullptr) || (r_in != nullptr) != (v_in != nullptr))
        return ree_set_error(REE_E_BADARG, "ree_vmfn_sample: injected variates need r, w and v together");
    if (n == 0) return 0;
    VmSampleArgs a{};
    a.x = x; a.n = n; a.ld = ld; a.d = d; a.hv = householder_v; a.hb = householder_b; a.kappa = kappa; a.m = m;
    a.omega = omega; a.seed = seed; a.draw = draw; a.first = first_particle; a.r_in = r_in; a.w_in = w_in; a.v_in = v_in;
    k_vmfn_sample<<<vm_grid(n), VM_T, (size_t)d * sizeof(double), ree_stream(stream)>>>(a);
    return ree_check_launch("k_vmfn_sample");
}

extern "C" int ree_vmfn_logpdf(const double* x, int64_t n, int d, int64_t ld, const double* mu, double kappa, double m,
                               double omega, double c0, double* g, double* e, double* logq, double* workspace,
                               double* is4, const ree_lsf* fuse_lsf, void* stream) {
    if (!x || !mu || d < 1 || n < 1 || ld < n || (is4 && (!g || !e || !workspace)) || (!is4 && !logq && !fuse_lsf))
        return ree_set_error(REE_E_BADARG, "ree_vmfn_logpdf: bad argument");
    int fuse = 0;
    if (fuse_lsf) {
        if (fuse_lsf->kind == REE_LSF_AFFINE && fuse_lsf->d == d && g && e) fuse = fuse_lsf->coef ? 1 : 2;
        else return ree_set_error(REE_E_UNSUPPORTED, "ree_vmfn_logpdf: only the affine limit state can be fused");
    }
    cudaStream_t st = ree_stream(stream);
    const int ctas = vm_grid(n);
    double* small = is4 ? workspace + 8 : nullptr;
    const size_t smem = (size_t)(fuse == 1 ? 2 * d : d) * sizeof(double);
    const double* coef = fuse ? fuse_lsf->coef : nullptr;
    const double off = fuse ? fuse_lsf->offset : 0.0;
    if (fuse == 1) k_vmfn_logpdf<1><<<ctas, VM_T, smem, st>>>(x, n, d, ld, mu, kappa, m, omega, c0, g, e, logq, small, coef, off);
    else if (fuse == 2) k_vmfn_logpdf<2><<<ctas, VM_T, smem, st>>>(x, n, d, ld, mu, kappa, m, omega, c0, g, e, logq, small, coef, off);
    else k_vmfn_logpdf<0><<<ctas, VM_T, smem, st>>>(x, n, d, ld, mu, kappa, m, omega, c0, g, e, logq, small, coef, off);
    int rc = ree_check_launch("k_vmfn_logpdf");
    if (rc || !is4) return rc;
    return ree_mma_merge_small(small, ctas, is4, st);
}
