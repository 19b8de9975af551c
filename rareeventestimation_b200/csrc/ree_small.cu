// Thread-per-particle sweeps for small dimensions (d <= 8): the HBM-bound regime of the CBREE iteration (oscillator
// d = 6 at N = 1e8, the README example d = 4, the reference's own d = 2 test problems).
//
// At these sizes a particle is 16-64 bytes per array and the d x d contractions are a few dozen FMAs, so the DMMA tile
// machinery of ree_mma.cu (8-particle warp tiles, shared-memory Gram) only adds overhead: its sweep 2 ran at 3 % of
// the HBM roofline.  Here every thread owns whole particles: coalesced 8-byte SoA loads (a warp reads 256 contiguous
// bytes per dimension), the next particle's rows are requested before the current one is processed, the point stays
// in registers through transform -> limit state -> energy -> moments, and the d + d(d+1)/2 + 2 running sums live in
// registers for the whole sweep.  Cross-thread combination: warp shuffles, one shared-memory stage per CTA, a fixed-order
// second kernel over the CTA partials (deterministic: no floating-point atomics).
//
//   k_small_step   X' = t X + m + F z (z: Philox in registers, or injected), g' = lsf(X') for every closed-form kind,
//                  e', shifted sums of X', #{g' <= 0}                     solver.py:665-671, 677-682
//   k_small_post   u = Linv (x - mean), log q, IS statistics, weights, weighted shifted sums; or (plain mode) the
//                  unweighted shifted sums of an existing ensemble          solver.py:629-636, 672-674, 683-686
#include "lsf_point.cuh"
#include "mma_common.cuh"  // second-stage merge of the per-CTA ExpSums

#define SM_T 256

struct SmallArgs {
    const double* x;
    int64_t n, ld;
    // step
    double* x_new;
    double t;
    const double* m_noise;
    const double* F;  // row-major D x D
    int tri;          // F is lower triangular
    const double* z;
    uint64_t seed, draw;
    int64_t first;
    int lsf_kind;  // closed-form kind evaluated in the sweep (REE_LSF_* / LINEAR_SUM_INTERNAL), 0: none
    const double* coef;
    double offset;
    const double* shift;
    double* g_new;
    double* e_new;
    // post
    const double* g;
    const double* e;
    const double* mean;
    const double* Linv;
    const double* logdet;
    const double* wmax;  // NULL: plain unweighted sums
    double sigma, beta;
    int tgt_kind;
    double* logq;
    double* w_out;
    // outputs
    double* partials;  // [cta][NS]
    double* small;     // [cta][4] ExpSums (post, weighted)
};

template <int D>
struct SmallDims {
    static constexpr int NT = D * (D + 1) / 2;  // lower triangle
    static constexpr int NS = D + NT + 2;       // S1 | S2 (lower) | a | b
};

// block reduction of the NS per-thread sums into partials[blockIdx.x][NS] (fixed order)
template <int NS>
__device__ __forceinline__ void small_block_sums(double (&acc)[NS], double* stage /* [SM_T/32][NS] */, double* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        double v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) stage[warp * NS + k] = v;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < NS; k += SM_T) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < SM_T / 32; ++w) v += stage[w * NS + k];
        out[k] = v;
    }
}

// ---------------------------------------------------------------------------------------
// Moment sums on the tensor pipe (D <= 6).  The d + d(d+1)/2 + 2 per-thread accumulators of the scalar version are 58
// registers: with the point, its prefetch and the Philox state the kernel spilled (STL = 21 % of the stall samples,
// profiles/r1_small_d_kernels_ncu.txt) at two CTAs per SM.  Instead every lane puts its row vector
//     (y_0 .. y_{D-1}, c, f, 0..)        c = 1 (or sqrt(w)), f = 1[g <= 0] (scaled alike)
// into a shared 8 x 32 tile of its warp and the warp accumulates the 8 x 8 Gram of its 32 particles with 8 DMMAs
// (m8n8k4: the A and the B fragment of a Gram are the same value, tile[lane/4][4k + lane%4]): S2 = G[0:D,0:D],
// S1 = G[D,0:D], n (or sum w) = G[D,D], #{g<=0} = G[D+1,D].  Two accumulator registers per lane.
// ---------------------------------------------------------------------------------------
#ifndef SG_CTAS
#define SG_CTAS 2
#endif
#define SG_S 36  // row stride of the tile in doubles: rows 8*S/2 banks apart -> conflict-free 64-bit reads per half warp

__device__ __forceinline__ void small_gram_accumulate(double* tile, const double (&row)[8], double& c0, double& c1) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int r = 0; r < 8; ++r) tile[r * SG_S + lane] = row[r];
    __syncwarp();
    const double* src = tile + (lane >> 2) * SG_S + (lane & 3);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const double f = src[4 * k];
        dmma(c0, c1, f, f);
    }
    __syncwarp();
}

// per-CTA partial: the 8 x 8 Grams of the CTA's warps summed in warp order -> out[64]
__device__ __forceinline__ void small_gram_store(double c0, double c1, double* stage /* [SM_T/32][64] */, double* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    stage[warp * 64 + (lane >> 2) * 8 + 2 * (lane & 3)] = c0;
    stage[warp * 64 + (lane >> 2) * 8 + 2 * (lane & 3) + 1] = c1;
    __syncthreads();
    if (threadIdx.x < 64) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < SM_T / 32; ++w) v += stage[w * 64 + threadIdx.x];
        out[threadIdx.x] = v;
    }
}

template <int D>
__global__ void __launch_bounds__(SM_T, SG_CTAS) k_small_step_g(const SmallArgs a) {
    static_assert(D <= 6, "two spare rows of the 8 x 8 Gram carry the count and the failure indicator");
    __shared__ double sF[D * D], sm[D], ss[D], sc[D];
    __shared__ double tiles[(SM_T / 32) * 8 * SG_S];
    __shared__ double stage[(SM_T / 32) * 64];
    __shared__ double bm[BM_SMEM_DOUBLES];
    for (int k = threadIdx.x; k < D * D; k += SM_T) sF[k] = a.F[k];
    if (threadIdx.x < D) {
        sm[threadIdx.x] = a.m_noise[threadIdx.x];
        ss[threadIdx.x] = a.shift ? a.shift[threadIdx.x] : 0.0;
        sc[threadIdx.x] = (a.lsf_kind == REE_LSF_AFFINE && a.coef) ? a.coef[threadIdx.x] : 0.0;
    }
    bm_stage_tables(bm, threadIdx.x, SM_T);
    __syncthreads();
    BmTableShared tab;
    tab.addr = smem_u32(bm);
    double* tile = tiles + (threadIdx.x >> 5) * 8 * SG_S;
    double c0 = 0.0, c1 = 0.0;
    const double e_const = (double)D * REE_LOG_2PI;
    const int64_t stride = (int64_t)gridDim.x * SM_T;
    const int lane = threadIdx.x & 31;
    int64_t base = (int64_t)blockIdx.x * SM_T + (threadIdx.x - lane);  // the warp's first particle: warp-uniform loop
    double xc[D], zc[D];
#pragma unroll
    for (int j = 0; j < D; ++j) xc[j] = zc[j] = 0.0;
    if (base + lane < a.n) {
#pragma unroll
        for (int j = 0; j < D; ++j) xc[j] = a.x[(int64_t)j * a.ld + base + lane];
        if (a.z) {
#pragma unroll
            for (int j = 0; j < D; ++j) zc[j] = a.z[(int64_t)j * a.ld + base + lane];
        }
    }
    for (; base < a.n; base += stride) {
        const int64_t i = base + lane;
        const bool live = i < a.n;
        double x[D], z[D];
#pragma unroll
        for (int j = 0; j < D; ++j) {
            x[j] = xc[j];
            z[j] = zc[j];
        }
        const int64_t in = i + stride;  // request the next particle's rows before working on this one
        if (in < a.n) {
#pragma unroll
            for (int j = 0; j < D; ++j) xc[j] = a.x[(int64_t)j * a.ld + in];
            if (a.z) {
#pragma unroll
                for (int j = 0; j < D; ++j) zc[j] = a.z[(int64_t)j * a.ld + in];
            }
        }
        if (!a.z) {  // Philox: pair p serves dimensions p and p + 4 (same counters as ree_philox_normal)
#pragma unroll
            for (int p = 0; p < (D < 4 ? D : 4); ++p) {
                double z0, z1;
                philox_normal_pair(a.seed, (uint64_t)(a.first + i), (uint32_t)p, a.draw, tab, z0, z1);
                z[p] = z0;
                if (p + 4 < D) z[p + 4 < D ? p + 4 : 0] = z1;
            }
        }
        double row[8];
        double xn[D];
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) {
            double c = 0.0;
            if (a.tri) {
#pragma unroll
                for (int k = 0; k <= j; ++k) c = fma(sF[j * D + k], z[k], c);
            } else {
#pragma unroll
                for (int k = 0; k < D; ++k) c = fma(sF[j * D + k], z[k], c);
            }
            // t*X + (m_noise + Z F^T): grouping of solver.py:667
            xn[j] = __dadd_rn(__dmul_rn(a.t, x[j]), __dadd_rn(sm[j], c));
            if (live) a.x_new[(int64_t)j * a.ld + i] = xn[j];
            s = fma(xn[j], xn[j], s);
            row[j] = live ? xn[j] - ss[j] : 0.0;
        }
        double fail = 0.0;
        if (live) {
            if (a.e_new) a.e_new[i] = 0.5 * (e_const + s);
            if (a.lsf_kind) {
                const double gv = lsf_point<D>(a.lsf_kind, sc, a.offset, xn);
                a.g_new[i] = gv;
                fail = (gv <= 0.0) ? 1.0 : 0.0;
            }
        }
#pragma unroll
        for (int r = D; r < 8; ++r) row[r] = 0.0;
        row[D] = live ? 1.0 : 0.0;
        row[D + 1] = fail;
        small_gram_accumulate(tile, row, c0, c1);
    }
    small_gram_store(c0, c1, stage, a.partials + (size_t)blockIdx.x * 64);
}

template <int D>
__global__ void __launch_bounds__(SM_T, SG_CTAS) k_small_post_g(const SmallArgs a) {
    static_assert(D <= 6, "two spare rows of the 8 x 8 Gram carry the count and the failure indicator");
    __shared__ double sL[D * D], sm[D];
    __shared__ double tiles[(SM_T / 32) * 8 * SG_S];
    __shared__ double stage[(SM_T / 32) * 64];
    __shared__ ExpSums sh[32];
    __shared__ double Tsh[64];
    const bool weighted = a.wmax != nullptr;
    for (int k = threadIdx.x; k < D * D; k += SM_T) sL[k] = weighted ? a.Linv[k] : 0.0;
    if (threadIdx.x < D) sm[threadIdx.x] = a.mean ? a.mean[threadIdx.x] : 0.0;
    const uint32_t T = load_exp_table(Tsh);  // includes the CTA barrier
    double* tile = tiles + (threadIdx.x >> 5) * 8 * SG_S;
    double c0 = 0.0, c1 = 0.0;
    ExpAcc is_acc = expacc_empty();
    const double lq_const = weighted ? (double)D * REE_LOG_2PI + a.logdet[0] : 0.0;
    const double wmax = weighted ? a.wmax[0] : 0.0;
    const int64_t stride = (int64_t)gridDim.x * SM_T;
    const int lane = threadIdx.x & 31;
    int64_t base = (int64_t)blockIdx.x * SM_T + (threadIdx.x - lane);
    double xc[D], gc = 1.0, ec = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) xc[j] = 0.0;
    if (base + lane < a.n) {
#pragma unroll
        for (int j = 0; j < D; ++j) xc[j] = a.x[(int64_t)j * a.ld + base + lane];
        if (a.g) gc = a.g[base + lane];
        if (weighted) ec = a.e[base + lane];
    }
    for (; base < a.n; base += stride) {
        const int64_t i = base + lane;
        const bool live = i < a.n;
        double y[D];
#pragma unroll
        for (int j = 0; j < D; ++j) y[j] = xc[j] - sm[j];
        const double gi = gc, ei = ec;
        const int64_t in = i + stride;
        if (in < a.n) {
#pragma unroll
            for (int j = 0; j < D; ++j) xc[j] = a.x[(int64_t)j * a.ld + in];
            if (a.g) gc = a.g[in];
            if (weighted) ec = a.e[in];
        }
        double sw = live ? 1.0 : 0.0;  // sqrt of the particle's weight
        if (weighted && live) {
            double q = 0.0;
#pragma unroll
            for (int j = 0; j < D; ++j) {
                double u = 0.0;
#pragma unroll
                for (int k = 0; k <= j; ++k) u = fma(sL[j * D + k], y[k], u);
                q = fma(u, u, q);
            }
            const double lq = -0.5 * (lq_const + q);
            if (a.logq) a.logq[i] = lq;
            const double r = __dsub_rn(-ei, lq);  // -e_fun - log_pdf  (solver.py:684)
            expacc_add1_mult(is_acc, r, gi <= 0.0 ? 1.0 : 0.0, T);
            const double al = __dmul_rn(ree_logtgt(gi, ei, a.sigma, a.tgt_kind), a.beta);
            const double wv = exp_nonpos(isfinite(al) ? al - wmax : -CUDART_INF, T);
            if (a.w_out) a.w_out[i] = wv;
            sw = sqrt(wv);
        }
        double row[8];
#pragma unroll
        for (int j = 0; j < D; ++j) row[j] = live ? sw * y[j] : 0.0;
#pragma unroll
        for (int r = D; r < 8; ++r) row[r] = 0.0;
        row[D] = sw;
        row[D + 1] = (!weighted && live && gi <= 0.0) ? 1.0 : 0.0;
        small_gram_accumulate(tile, row, c0, c1);
    }
    small_gram_store(c0, c1, stage, a.partials + (size_t)blockIdx.x * 64);
    if (weighted) {
        ExpSums blk = expsums_block_reduce(expacc_finish(is_acc), sh);
        if (threadIdx.x == 0) {
            double* p = a.small + 4 * (size_t)blockIdx.x;
            p[0] = blk.m; p[1] = blk.s1; p[2] = blk.s2; p[3] = blk.cnt;
        }
    }
}

// second stage of the Gram variants: fixed-order sum of the per-CTA 8 x 8 Grams, expanded to the public layout
//   sums = [S1 (d) | S2 (d x d) | a | b]: step / plain: a = #{g<=0} = G[D+1][D], b = n = G[D][D]; weighted: a = sum w, b = 0
__global__ void __launch_bounds__(512) k_small_finish_g(const double* __restrict__ partials, int nparts, int d,
                                                        int weighted, double* __restrict__ sums) {
    __shared__ double G[64], part[8][64];
    const int e = threadIdx.x & 63, grp = threadIdx.x >> 6;  // 8 groups of CTA partials, summed in group order
    double s = 0.0;
    for (int p = grp; p < nparts; p += 8) s += partials[(size_t)p * 64 + e];
    part[grp][e] = s;
    __syncthreads();
    if (threadIdx.x >= 64) return;
    s = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += part[q][e];
    G[e] = s;
    __syncwarp();
    asm volatile("bar.sync 1, 64;" ::: "memory");  // the 64 surviving threads (two warps)
    const int r = threadIdx.x >> 3, c = threadIdx.x & 7;
    if (r < d && c < d) sums[d + r * d + c] = G[(r >= c ? r : c) * 8 + (r >= c ? c : r)];  // the lower triangle's bits
    if (r == d && c < d) sums[c] = G[d * 8 + c];
    if (threadIdx.x == 0) {
        sums[d + d * d] = weighted ? G[d * 8 + d] : G[(d + 1) * 8 + d];
        sums[d + d * d + 1] = weighted ? 0.0 : G[d * 8 + d];
    }
}

// ---------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(SM_T, 2) k_small_step(const SmallArgs a) {
    constexpr int NT = SmallDims<D>::NT, NS = SmallDims<D>::NS;
    __shared__ double sF[D * D], sm[D], ss[D], sc[D];
    __shared__ double stage[(SM_T / 32) * NS];
    for (int k = threadIdx.x; k < D * D; k += SM_T) sF[k] = a.F[k];
    if (threadIdx.x < D) {
        sm[threadIdx.x] = a.m_noise[threadIdx.x];
        ss[threadIdx.x] = a.shift ? a.shift[threadIdx.x] : 0.0;
        sc[threadIdx.x] = (a.lsf_kind == REE_LSF_AFFINE && a.coef) ? a.coef[threadIdx.x] : 0.0;
    }
    __syncthreads();
    double acc[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k) acc[k] = 0.0;
    const double e_const = (double)D * REE_LOG_2PI;
    const int64_t stride = (int64_t)gridDim.x * SM_T;
    int64_t i = (int64_t)blockIdx.x * SM_T + threadIdx.x;
    double xc[D], zc[D];
    if (i < a.n) {
#pragma unroll
        for (int j = 0; j < D; ++j) xc[j] = a.x[(int64_t)j * a.ld + i];
        if (a.z) {
#pragma unroll
            for (int j = 0; j < D; ++j) zc[j] = a.z[(int64_t)j * a.ld + i];
        }
    }
    for (; i < a.n; i += stride) {
        double x[D], z[D];
#pragma unroll
        for (int j = 0; j < D; ++j) {
            x[j] = xc[j];
            z[j] = zc[j];
        }
        const int64_t in = i + stride;  // request the next particle's rows before working on this one
        if (in < a.n) {
#pragma unroll
            for (int j = 0; j < D; ++j) xc[j] = a.x[(int64_t)j * a.ld + in];
            if (a.z) {
#pragma unroll
                for (int j = 0; j < D; ++j) zc[j] = a.z[(int64_t)j * a.ld + in];
            }
        }
        if (!a.z) {  // Philox: pair p serves dimensions p and p + 4 (same counters as ree_philox_normal)
#pragma unroll
            for (int p = 0; p < (D < 4 ? D : 4); ++p) {
                double z0, z1;
                philox_normal_pair(a.seed, (uint64_t)(a.first + i), (uint32_t)p, a.draw, z0, z1);
                z[p] = z0;
                if (p + 4 < D) z[p + 4 < D ? p + 4 : 0] = z1;
            }
        }
        double xn[D];
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) {
            double c = 0.0;
            if (a.tri) {
#pragma unroll
                for (int k = 0; k <= j; ++k) c = fma(sF[j * D + k], z[k], c);
            } else {
#pragma unroll
                for (int k = 0; k < D; ++k) c = fma(sF[j * D + k], z[k], c);
            }
            // t*X + (m_noise + Z F^T): grouping of solver.py:667
            xn[j] = __dadd_rn(__dmul_rn(a.t, x[j]), __dadd_rn(sm[j], c));
            a.x_new[(int64_t)j * a.ld + i] = xn[j];
            s = fma(xn[j], xn[j], s);
        }
        if (a.e_new) a.e_new[i] = 0.5 * (e_const + s);
        if (a.lsf_kind) {
            const double gv = lsf_point<D>(a.lsf_kind, sc, a.offset, xn);
            a.g_new[i] = gv;
            acc[D + NT] += (gv <= 0.0) ? 1.0 : 0.0;
        }
        acc[D + NT + 1] += 1.0;
        double y[D];
#pragma unroll
        for (int j = 0; j < D; ++j) y[j] = xn[j] - ss[j];
#pragma unroll
        for (int j = 0; j < D; ++j) {
            acc[j] += y[j];
#pragma unroll
            for (int k = 0; k <= j; ++k) acc[D + j * (j + 1) / 2 + k] = fma(y[j], y[k], acc[D + j * (j + 1) / 2 + k]);
        }
    }
    small_block_sums<NS>(acc, stage, a.partials + (size_t)blockIdx.x * NS);
}

// ---------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(SM_T, 2) k_small_post(const SmallArgs a) {
    constexpr int NT = SmallDims<D>::NT, NS = SmallDims<D>::NS;
    __shared__ double sL[D * D], sm[D];
    __shared__ double stage[(SM_T / 32) * NS];
    __shared__ ExpSums sh[32];
    const bool weighted = a.wmax != nullptr;
    for (int k = threadIdx.x; k < D * D; k += SM_T) sL[k] = weighted ? a.Linv[k] : 0.0;
    if (threadIdx.x < D) sm[threadIdx.x] = a.mean ? a.mean[threadIdx.x] : 0.0;
    __syncthreads();
    double acc[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k) acc[k] = 0.0;
    ExpSums is_acc = expsums_empty();
    const double lq_const = weighted ? (double)D * REE_LOG_2PI + a.logdet[0] : 0.0;
    const double wmax = weighted ? a.wmax[0] : 0.0;
    const int64_t stride = (int64_t)gridDim.x * SM_T;
    int64_t i = (int64_t)blockIdx.x * SM_T + threadIdx.x;
    double xc[D], gc = 1.0, ec = 0.0;
    if (i < a.n) {
#pragma unroll
        for (int j = 0; j < D; ++j) xc[j] = a.x[(int64_t)j * a.ld + i];
        if (a.g) gc = a.g[i];
        if (weighted) ec = a.e[i];
    }
    for (; i < a.n; i += stride) {
        double y[D];
#pragma unroll
        for (int j = 0; j < D; ++j) y[j] = xc[j] - sm[j];
        const double gi = gc, ei = ec;
        const int64_t in = i + stride;
        if (in < a.n) {
#pragma unroll
            for (int j = 0; j < D; ++j) xc[j] = a.x[(int64_t)j * a.ld + in];
            if (a.g) gc = a.g[in];
            if (weighted) ec = a.e[in];
        }
        double wv = 1.0;
        if (weighted) {
            double q = 0.0;
#pragma unroll
            for (int j = 0; j < D; ++j) {
                double u = 0.0;
#pragma unroll
                for (int k = 0; k <= j; ++k) u = fma(sL[j * D + k], y[k], u);
                q = fma(u, u, q);
            }
            const double lq = -0.5 * (lq_const + q);
            if (a.logq) a.logq[i] = lq;
            const double r = __dsub_rn(-ei, lq);  // -e_fun - log_pdf  (solver.py:684)
            if (isfinite(r)) expsums_add(is_acc, r, gi <= 0.0 ? 1.0 : 0.0);
            const double al = __dmul_rn(ree_logtgt(gi, ei, a.sigma, a.tgt_kind), a.beta);
            wv = isfinite(al) ? exp(al - wmax) : 0.0;
            if (a.w_out) a.w_out[i] = wv;
            acc[D + NT] += wv;
        } else {
            acc[D + NT] += (gi <= 0.0) ? 1.0 : 0.0;
            acc[D + NT + 1] += 1.0;
        }
#pragma unroll
        for (int j = 0; j < D; ++j) {
            const double wy = weighted ? wv * y[j] : y[j];
            acc[j] += wy;
#pragma unroll
            for (int k = 0; k <= j; ++k) acc[D + j * (j + 1) / 2 + k] = fma(wy, y[k], acc[D + j * (j + 1) / 2 + k]);
        }
    }
    small_block_sums<NS>(acc, stage, a.partials + (size_t)blockIdx.x * NS);
    if (weighted) {
        ExpSums blk = expsums_block_reduce(is_acc, sh);
        if (threadIdx.x == 0) {
            double* p = a.small + 4 * (size_t)blockIdx.x;
            p[0] = blk.m; p[1] = blk.s1; p[2] = blk.s2; p[3] = blk.cnt;
        }
    }
}

// second stage: fixed-order sum over the CTA partials, expanded to the public layout
//   sums = [S1 (d) | S2 (d x d, symmetric) | a | b]
__global__ void __launch_bounds__(64) k_small_finish(const double* __restrict__ partials, int nparts, int d,
                                                     double* __restrict__ sums) {
    const int NT = d * (d + 1) / 2, NS = d + NT + 2;
    const int k = threadIdx.x;
    if (k >= NS) return;
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += partials[(size_t)p * NS + k];
    if (k < d) {
        sums[k] = s;
    } else if (k < d + NT) {
        int t = k - d, j = 0;
        while ((j + 1) * (j + 2) / 2 <= t) ++j;
        const int c = t - j * (j + 1) / 2;
        sums[d + j * d + c] = s;
        sums[d + c * d + j] = s;
    } else {
        sums[d + d * d + (k - d - NT)] = s;
    }
}

// ---------------------------------------------------------------------------------------
bool ree_small_supported(int d) {
    static int forced = -1;
    if (forced < 0) {  // REE_NO_SMALL=1: test hook that routes small d through the DMMA sweeps (both are CUDA paths)
        const char* e = getenv("REE_NO_SMALL");
        forced = (e && e[0] == '1') ? 1 : 0;
    }
    return !forced && d >= 1 && d <= 8;
}

static bool small_gram_path(int d) {
    static int off = -1;
    if (off < 0) {  // REE_SMALL_SCALAR=1: test / tuning hook that keeps the per-thread accumulators for every d <= 8
        const char* e = getenv("REE_SMALL_SCALAR");
        off = (e && e[0] == '1') ? 1 : 0;
    }
    return !off && d <= 6;
}

static int small_grid(int64_t n, int d) {
    const int64_t want = (n + SM_T - 1) / SM_T;
    const int cap = small_gram_path(d) ? SG_CTAS * (ree_grid_ctas() / 2) : ree_grid_ctas();
    return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

template <int D>
static int launch_small(int which, const SmallArgs& a, int grid, cudaStream_t st) {
    if constexpr (D <= 6) {
        if (small_gram_path(D)) {
            if (which == 0) k_small_step_g<D><<<grid, SM_T, 0, st>>>(a);
            else k_small_post_g<D><<<grid, SM_T, 0, st>>>(a);
            return ree_check_launch(which == 0 ? "k_small_step_g" : "k_small_post_g");
        }
    }
    if (which == 0) k_small_step<D><<<grid, SM_T, 0, st>>>(a);
    else k_small_post<D><<<grid, SM_T, 0, st>>>(a);
    return ree_check_launch(which == 0 ? "k_small_step" : "k_small_post");
}

static int dispatch_small(int which, int d, const SmallArgs& a, int grid, cudaStream_t st) {
    switch (d) {
        case 1: return launch_small<1>(which, a, grid, st);
        case 2: return launch_small<2>(which, a, grid, st);
        case 3: return launch_small<3>(which, a, grid, st);
        case 4: return launch_small<4>(which, a, grid, st);
        case 5: return launch_small<5>(which, a, grid, st);
        case 6: return launch_small<6>(which, a, grid, st);
        case 7: return launch_small<7>(which, a, grid, st);
        case 8: return launch_small<8>(which, a, grid, st);
    }
    return ree_set_error(REE_E_UNSUPPORTED, "small-d path: d > 8");
}

static int finish_small(const double* partials, int grid, int d, int weighted, double* sums, cudaStream_t st) {
    if (small_gram_path(d)) {
        k_small_finish_g<<<1, 512, 0, st>>>(partials, grid, d, weighted, sums);
        return ree_check_launch("k_small_finish_g");
    }
    k_small_finish<<<1, 64, 0, st>>>(partials, grid, d, sums);
    return ree_check_launch("k_small_finish");
}

// lsf_kind: closed-form kind evaluated inside the sweep, 0 if the caller evaluates g itself
int ree_small_step(const double* x, double* x_new, int64_t n, int d, int64_t ld, double t, const double* m_noise,
                   const double* F, int tri, const double* z, uint64_t seed, uint64_t draw, int64_t first, int lsf_kind,
                   const double* coef, double offset, const double* shift, double* g_new, double* e_new, double* ws_big,
                   double* sums, cudaStream_t st) {
    SmallArgs a{};
    a.x = x; a.n = n; a.ld = ld; a.x_new = x_new; a.t = t; a.m_noise = m_noise; a.F = F; a.tri = tri; a.z = z;
    a.seed = seed; a.draw = draw; a.first = first; a.lsf_kind = lsf_kind; a.coef = coef; a.offset = offset;
    a.shift = shift; a.g_new = g_new; a.e_new = e_new; a.partials = ws_big;
    const int grid = small_grid(n, d);
    int rc = dispatch_small(0, d, a, grid, st);
    if (rc) return rc;
    return finish_small(ws_big, grid, d, 0, sums, st);
}

// weighted (wmax != NULL): sums shifted by `mean`, tail [sum w, 0], is4 = IS statistics.
// plain (wmax == NULL): sums of x - mean (mean = shift, may be NULL), tail [#{g<=0}, n].
int ree_small_post(const double* x, int64_t n, int d, int64_t ld, const double* g, const double* e, const double* mean,
                   const double* Linv, const double* logdet, double sigma, double beta, int tgt_kind, const double* wmax,
                   double* logq, double* w, double* ws_big, double* ws_small, double* sums, double* is4,
                   cudaStream_t st) {
    SmallArgs a{};
    a.x = x; a.n = n; a.ld = ld; a.g = g; a.e = e; a.mean = mean; a.Linv = Linv; a.logdet = logdet; a.sigma = sigma;
    a.beta = beta; a.tgt_kind = tgt_kind; a.wmax = wmax; a.logq = logq; a.w_out = w; a.partials = ws_big;
    a.small = ws_small;
    const int grid = small_grid(n, d);
    int rc = dispatch_small(1, d, a, grid, st);
    if (rc) return rc;
    rc = finish_small(ws_big, grid, d, wmax != nullptr, sums, st);
    if (rc || !wmax) return rc;
    return ree_mma_merge_small(ws_small, grid, is4, st);
}
