// Small dense algebra on the d x d consensus statistics: moments from raw sums, Cholesky, inverse
// of the factor.  One CTA per call; the matrix lives in shared memory for d <= REE_MAX_D_SMEM and in a per-device
// global scratch (L1 / L2 resident, same code: CTA barriers order global accesses inside a CTA too) up to REE_MAX_D.
#include "common.cuh"

#define REE_MAX_D_SMEM 160
#define REE_MAX_D 256
#define REE_LA_THREADS 256

// in-place lower Cholesky of A (row-major, leading dimension lda) in shared memory.
// psd = true: a pivot <= tiny * max_diag gives a zero column (semi-definite factor, no failure);
// psd = false: the first non-positive pivot is reported in *info (1-based) and the sweep stops.
//
// Blocked right-looking sweep, panels of CH_NB columns: warp 0 factors a panel with warp-level synchronisation only
// (lane = row), then all warps apply the panel to the trailing matrix (warp = row, lane = column).  Every element still
// receives its updates one column at a time in increasing column order, so the factor has the same bits as the
// column-by-column sweep it replaces -- with 2 CTA barriers per panel instead of 3 per column, and the logarithms of
// the pivots (log det) taken off the critical path.  pivlog: d doubles of shared scratch.
#define CH_NB 8

__device__ void chol_smem(double* A, int d, int lda, bool psd, double* s_stat /* [4] smem */, double* pivlog) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    if (tid == 0) {
        double mx = 0.0;
        for (int j = 0; j < d; ++j) mx = fmax(mx, A[j * lda + j]);
        s_stat[0] = 0.0;          // logdet accumulator
        s_stat[1] = 0.0;          // info
        s_stat[2] = CUDART_INF;   // min pivot
        s_stat[3] = mx;           // max diag
    }
    for (int j = tid; j < d; j += nt) pivlog[j] = 0.0;  // accepted pivots (0: zero column / not reached)
    __syncthreads();
    const double mx = s_stat[3];
    for (int kb = 0; kb < d; kb += CH_NB) {
        const int ke = min(kb + CH_NB, d);
        if (warp == 0) {  // ---- panel: columns kb .. ke-1, rows kb .. d-1 ----
            for (int k = kb; k < ke; ++k) {
                if (lane == 0) {
                    const double piv = A[k * lda + k];
                    s_stat[2] = fmin(s_stat[2], piv);
                    if (!(piv > 0.0) || !isfinite(piv)) {
                        if (psd && isfinite(piv)) {
                            A[k * lda + k] = 0.0;
                        } else if (s_stat[1] == 0.0) {
                            s_stat[1] = (double)(k + 1);
                        }
                    } else if (psd && piv <= 1e-14 * mx) {
                        A[k * lda + k] = 0.0;
                    } else {
                        pivlog[k] = piv;
                        A[k * lda + k] = sqrt(piv);
                    }
                }
                __syncwarp();
                if (s_stat[1] != 0.0) break;
                const double lkk = A[k * lda + k];
                const double inv = lkk > 0.0 ? 1.0 / lkk : 0.0;
                for (int j = k + 1 + lane; j < d; j += 32) A[j * lda + k] *= inv;
                __syncwarp();
                // the rest of the panel: A[j][l] -= A[j][k] * A[l][k], k < l < ke, l <= j
                for (int j = k + 1 + lane; j < d; j += 32) {
                    const double ajk = A[j * lda + k];
                    const int le = min(ke, j + 1);
                    for (int l = k + 1; l < le; ++l) A[j * lda + l] = fma(-ajk, A[l * lda + k], A[j * lda + l]);
                }
                __syncwarp();
            }
        }
        __syncthreads();
        if (s_stat[1] != 0.0) break;
        // ---- trailing matrix: rows j >= ke, columns ke .. j; the panel's columns one after the other ----
        for (int j = ke + warp; j < d; j += nw) {
            double ajk[CH_NB];
#pragma unroll
            for (int c = 0; c < CH_NB; ++c) ajk[c] = (kb + c < ke) ? A[j * lda + kb + c] : 0.0;
            for (int l = ke + lane; l <= j; l += 32) {
                double v = A[j * lda + l];
#pragma unroll
                for (int c = 0; c < CH_NB; ++c)
                    if (kb + c < ke) v = fma(-ajk[c], A[l * lda + kb + c], v);
                A[j * lda + l] = v;
            }
        }
        __syncthreads();
    }
    // log det = sum of log(pivot) in column order (same summation order as before)
    for (int j = tid; j < d; j += nt) pivlog[j] = pivlog[j] > 0.0 ? log(pivlog[j]) : 0.0;
    __syncthreads();
    if (tid == 0) {
        double acc = 0.0;
        for (int j = 0; j < d; ++j)
            if (pivlog[j] != 0.0) acc += pivlog[j];
        s_stat[0] = acc;
    }
    __syncthreads();
}

__device__ __forceinline__ double block_max_la(double v, double* smem /* >= 32 */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    double m = 0.0;
    for (int w = 0; w < nw; ++w) m = fmax(m, smem[w]);
    __syncthreads();
    return m;
}

// mean / cov from shifted raw sums, then chol(cov) and its inverse.
// sums = [S1 (d) | S2 (d*d) | a | b]; unweighted: n = b ; weighted: W = a (sum of weights).
__global__ void __launch_bounds__(REE_LA_THREADS) k_moments_chol(const double* __restrict__ sums,
                                                                 const double* __restrict__ shift, int d, int ddof,
                                                                 int weighted, double* __restrict__ mean,
                                                                 double* __restrict__ cov, double* __restrict__ L,
                                                                 double* __restrict__ Linv, double* __restrict__ stats4,
                                                                 double* gscratch) {
    extern __shared__ double sm[];
    const int lda = d + 1;
    double* A = gscratch ? gscratch : sm;            // d x lda
    double* ybar = gscratch ? sm : A + d * lda;     // d
    double* s_stat = ybar + d;      // 4
    double* pivlog = s_stat + 4;    // d
    __shared__ double red[32];
    const int tid = threadIdx.x, nt = blockDim.x;
    const double tot = weighted ? sums[d + d * d] : sums[d + d * d + 1];
    for (int j = tid; j < d; j += nt) {
        double yb = sums[j] / tot;
        ybar[j] = yb;
        mean[j] = (shift ? shift[j] : 0.0) + yb;
    }
    __syncthreads();
    // np.cov: fact = W - ddof*sum(w^2)/W ; for the unweighted case w=1: fact = n - ddof.  The
    // reference only uses (aweights, ddof=0) and (no weights, ddof=1)   solver.py:634-636, 671.
    const double fact = tot - (double)ddof;
    for (int idx = tid; idx < d * d; idx += nt) {
        int j = idx / d, k = idx - j * d;
        double c = (sums[d + idx] - tot * ybar[j] * ybar[k]) / fact;
        A[j * lda + k] = c;
        cov[idx] = c;
    }
    __syncthreads();
    if (!L) return;
    if (Linv)  // stays zero if the factorisation fails (callers check stats4[1])
        for (int idx = tid; idx < d * d; idx += nt) Linv[idx] = 0.0;
    chol_smem(A, d, lda, false, s_stat, pivlog);
    for (int idx = tid; idx < d * d; idx += nt) {
        int j = idx / d, k = idx - j * d;
        L[idx] = (k <= j) ? A[j * lda + k] : 0.0;
    }
    // ||cov||_inf (largest absolute row sum, from the global copy written above): an upper bound of lambda_max
    {
        double rs = 0.0;
        for (int j = tid; j < d; j += nt) {
            double acc = 0.0;
            for (int k = 0; k < d; ++k) acc += fabs(cov[j * d + k]);
            rs = fmax(rs, acc);
        }
        __syncthreads();
        rs = block_max_la(rs, red);
        if (tid == 0) {
            stats4[0] = s_stat[0];
            stats4[1] = s_stat[1];
            stats4[2] = s_stat[2];
            stats4[3] = s_stat[3];
            stats4[4] = 0.0;
            stats4[5] = rs;
            stats4[6] = stats4[7] = 0.0;
        }
    }
    if (!Linv || s_stat[1] != 0.0) return;
    // inverse of the lower factor, one column per thread: solve L x = e_c by forward substitution (same arithmetic as
    // before: bit-identical).  Column c of the inverse lives in shared memory next to the factor: the strict upper
    // triangle plus the padding column of A has exactly d(d+1)/2 free slots, x_j goes to A[c][j+1] (row c: contiguous
    // for the thread, rows d+1 apart across threads: conflict-free) -- no global-memory round trip per term.
    __syncthreads();
    double fro = 0.0;  // |Linv|_F^2 = trace(cov^-1) >= 1 / lambda_min
    for (int c = tid; c < d; c += nt) {
        double* x = A + c * lda + 1;  // x[j], j >= c
        for (int j = c; j < d; ++j) {
            double s = (j == c) ? 1.0 : 0.0;
            const double* Lj = A + j * lda;
            for (int k = c; k < j; ++k) s = fma(-Lj[k], x[k], s);
            const double v = s / Lj[j];
            x[j] = v;
            fro = fma(v, v, fro);
        }
    }
    __syncthreads();
    fro = block_sum(fro, red);
    if (tid == 0) stats4[4] = fro;
    for (int idx = tid; idx < d * d; idx += nt) {
        int j = idx / d, c = idx - j * d;
        if (c <= j) Linv[idx] = A[c * lda + j + 1];
    }
}

__global__ void __launch_bounds__(REE_LA_THREADS) k_chol_scaled(const double* __restrict__ C, int d, double scale,
                                                                double* __restrict__ L, double* __restrict__ stats4,
                                                                double* gscratch) {
    extern __shared__ double sm[];
    const int lda = d + 1;
    double* A = gscratch ? gscratch : sm;
    double* s_stat = gscratch ? sm : A + d * lda;
    double* pivlog = s_stat + 4;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int idx = tid; idx < d * d; idx += nt) {
        int j = idx / d, k = idx - j * d;
        // symmetrise from the lower triangle so that the factor does not depend on upper-triangle noise
        double v = (k <= j) ? C[idx] : C[k * d + j];
        A[j * lda + k] = scale * v;
    }
    __syncthreads();
    chol_smem(A, d, lda, true, s_stat, pivlog);
    for (int idx = tid; idx < d * d; idx += nt) {
        int j = idx / d, k = idx - j * d;
        L[idx] = (k <= j) ? A[j * lda + k] : 0.0;
    }
    if (tid == 0 && stats4) {
        stats4[0] = s_stat[0];
        stats4[1] = s_stat[1];
        stats4[2] = s_stat[2];
        stats4[3] = s_stat[3];
        stats4[4] = stats4[5] = stats4[6] = stats4[7] = 0.0;
    }
}

// d x (d+1) scratch of the current device for REE_MAX_D_SMEM < d <= REE_MAX_D (allocated once; calls on one device are
// stream-ordered by the engine: one solve at a time per device)
static double* la_scratch() {
    static double* buf[64] = {nullptr};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (!buf[dev] && cudaMalloc((void**)&buf[dev], sizeof(double) * REE_MAX_D * (REE_MAX_D + 1)) != cudaSuccess) {
        buf[dev] = nullptr;
        ree_set_error((int)cudaErrorMemoryAllocation, "linalg scratch: cudaMalloc failed");
    }
    return buf[dev];
}

static void la_attrs() {  // once per device: opt in to the large dynamic shared-memory size
    static bool done[64] = {false};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || done[dev]) return;
    const int mx = (int)(((size_t)REE_MAX_D_SMEM * (REE_MAX_D_SMEM + 1) + 2 * REE_MAX_D + 4) * sizeof(double));
    cudaFuncSetAttribute(k_moments_chol, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    cudaFuncSetAttribute(k_chol_scaled, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    done[dev] = true;
}

extern "C" int ree_max_dim(void) { return REE_MAX_D; }

extern "C" int ree_moments_chol(const double* sums, const double* shift, int d, int ddof, int weighted, double* mean,
                                double* cov, double* L, double* Linv, double* stats4, void* stream) {
    if (!sums || !mean || !cov || d < 1 || (L && !stats4))
        return ree_set_error(REE_E_BADARG, "ree_moments_chol: bad argument");
    if (d > REE_MAX_D) return ree_set_error(REE_E_UNSUPPORTED, "ree_moments_chol: d > 256 not supported");
    double* gs = nullptr;
    if (d > REE_MAX_D_SMEM && !(gs = la_scratch())) return (int)cudaErrorMemoryAllocation;
    size_t smem = ((gs ? 0 : (size_t)d * (d + 1)) + 2 * d + 4) * sizeof(double);
    la_attrs();
    k_moments_chol<<<1, REE_LA_THREADS, smem, ree_stream(stream)>>>(sums, shift, d, ddof, weighted, mean, cov, L, Linv,
                                                                    stats4, gs);
    return ree_check_launch("ree_moments_chol");
}

extern "C" int ree_chol_scaled(const double* C, int d, double scale, double* L, double* stats4, void* stream) {
    if (!C || !L || d < 1) return ree_set_error(REE_E_BADARG, "ree_chol_scaled: bad argument");
    if (d > REE_MAX_D) return ree_set_error(REE_E_UNSUPPORTED, "ree_chol_scaled: d > 256 not supported");
    double* gs = nullptr;
    if (d > REE_MAX_D_SMEM && !(gs = la_scratch())) return (int)cudaErrorMemoryAllocation;
    size_t smem = ((gs ? 0 : (size_t)d * (d + 1)) + d + 4) * sizeof(double);
    la_attrs();
    k_chol_scaled<<<1, REE_LA_THREADS, smem, ree_stream(stream)>>>(C, d, scale, L, stats4, gs);
    return ree_check_launch("ree_chol_scaled");
}
