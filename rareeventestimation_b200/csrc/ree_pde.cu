// 1-D lognormal-field diffusion PDE limit state (problem/diffusion.py:131-171, 232-242) on the fp64 tensor pipe.
//
//   a_q   = <T[q, :], theta>                q = 0 .. 3 n_ele - 1     (mode table T: 1536 x 150, diffusion.py:222-229)
//   Ke_e  = (1/2h) sum_{r<3} wGL_r exp(mu + sigma a_{3e+r})         (diffusion.py:141-151)
//   u(1)  = sum_e h (n_ele - e - 1/2) / Ke_e ,   g = u_max - u(1)
//
// The last line replaces the 512 x 512 tridiagonal solve: the load is constant, so the flux telescopes and the boundary
// value is a plain sum over the elements (SURVEY 7.2 "PDE LSF"; equal to the reference's SuperLU result to 7e-13 on
// the golden vectors, the same distance as a Thomas sweep).  No sequential recurrence is left, and the field
// contraction -- 4.6e5 flop per particle, 99 % of the work -- is a (1536 x 152) x (152 x N) product on DMMA:
//
//   * M = quadrature points, N = particles, K = modes.  One warp owns 8 particles; its 38 B fragments (theta, 152 x 8)
//     stay in registers for the whole particle tile.
//   * T is re-packed once per call into fragment order with its rows permuted so that lane (lr, lc) of the three
//     M-tiles of block b receives the three Gauss points of element e = 8 b + lr: exp, the Gauss-Legendre fold and
//     the element's term of the sum happen in the accumulator registers, no shuffles.
//   * The 64 packed blocks (29 KB each) stream L2 -> shared memory through the TMA engine (cp.async.bulk) in a ring of
//     stages; completion and reuse are tracked by mbarriers and a per-stage counter -- no CTA barrier in the loop.
#include "mma_common.cuh"

template <int KK>
struct PdeDims {
    static constexpr int BLOCK_DOUBLES = 3 * KK * 32;  // one packed block: [mt][kk][lane]
    static constexpr uint32_t BLOCK_BYTES = BLOCK_DOUBLES * 8;
};

// Tp[b][mt][kk][lane] = T[3 (8 b + lr) + mt][4 kk + lc]   (lane = 4 lr + lc; zero beyond d)
__global__ void __launch_bounds__(256) k_pde_pack(const double* __restrict__ T, int d, int nb, int kk_n,
                                                  double* __restrict__ Tp) {
    const int64_t total = (int64_t)nb * 3 * kk_n * 32;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int lane = (int)(idx & 31);
        int64_t r = idx >> 5;
        const int kk = (int)(r % kk_n);
        r /= kk_n;
        const int mt = (int)(r % 3), b = (int)(r / 3);
        const int lr = lane >> 2, lc = lane & 3;
        const int q = 3 * (8 * b + lr) + mt, k = 4 * kk + lc;
        Tp[idx] = (k < d) ? T[(size_t)q * d + k] : 0.0;
    }
}

template <int KK, int NW, int S>
__global__ void __launch_bounds__(NW * 32, 1) k_pde_mma(const double* __restrict__ Tp, int nb, const double* __restrict__ x,
                                                        int64_t n, int64_t ld, int d, double mu, double sigma, double u_max,
                                                        double* __restrict__ g) {
    constexpr int BD = PdeDims<KK>::BLOCK_DOUBLES;
    constexpr uint32_t BB = PdeDims<KK>::BLOCK_BYTES;
    extern __shared__ __align__(128) double sm[];  // S stages of one packed block
    __shared__ __align__(8) unsigned long long full_bar[S];
    __shared__ unsigned int done_cnt[S];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, lr = lane >> 2, lc = lane & 3;
    const uint32_t sm_addr = smem_u32(sm), bar0 = smem_u32(&full_bar[0]);
    const int ne = 8 * nb;
    const double h = 1.0 / (double)ne;
    const double inv2h = 1.0 / h / 2.0;  // 1/h/2 as in diffusion.py:146
    const double w0 = 5.0 / 9.0, w1 = 8.0 / 9.0;

    const int64_t nwt = (n + 7) >> 3;
    // rounds of this CTA: in round r warp w owns particle tile (blockIdx.x + r gridDim.x) NW + w
    int64_t rounds = 0;
    for (int64_t t0 = (int64_t)blockIdx.x * NW; t0 < nwt; t0 += (int64_t)gridDim.x * NW) ++rounds;
    const int64_t total = rounds * nb;  // packed blocks this CTA consumes (the T stream restarts every round)
    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(bar0 + 8 * s, 1);
            done_cnt[s] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto load_block = [&](int64_t j) {  // one thread: block j of the stream -> stage j % S
        const int s = (int)(j % S);
        fence_proxy_async();
        mbar_expect_tx(bar0 + 8 * s, BB);
        bulk_g2s(sm_addr + (uint32_t)s * BB, Tp + (size_t)(j % nb) * BD, BB, bar0 + 8 * s);
    };
    if (tid == 0)
        for (int64_t j = 0; j < S && j < total; ++j) load_block(j);

    int64_t j = 0;
    for (int64_t r = 0; r < rounds; ++r) {
        const int64_t wt = ((int64_t)blockIdx.x + r * gridDim.x) * NW + warp;
        const bool live = wt < nwt;
        const int64_t i0 = wt * 8;
        // theta of the warp's 8 particles as B fragments: lane holds theta[4 kk + lc][i0 + lr]
        double bf[KK];
        {
            const int64_t ip = i0 + lr;
            const bool pv = live && ip < n;
#pragma unroll
            for (int kk = 0; kk < KK; ++kk) {
                const int k = 4 * kk + lc;
                bf[kk] = (pv && k < d) ? __ldg(x + (int64_t)k * ld + ip) : 0.0;
            }
        }
        double u0 = 0.0, u1 = 0.0;  // partial sums of u(1) for the lane's two particles (elements e = lr mod 8)
#pragma unroll 1
        for (int b = 0; b < nb; ++b, ++j) {
            const int s = (int)(j % S);
            mbar_wait(bar0 + 8 * s, (uint32_t)((j / S) & 1));
            if (live) {
                const uint32_t a_lane = sm_addr + (uint32_t)s * BB + (uint32_t)lane * 8u;
                double c[3][2];
#pragma unroll
                for (int mt = 0; mt < 3; ++mt) c[mt][0] = c[mt][1] = 0.0;
#pragma unroll
                for (int kk = 0; kk < KK; ++kk) {
#pragma unroll
                    for (int mt = 0; mt < 3; ++mt)
                        dmma(c[mt][0], c[mt][1], lds64(a_lane + (uint32_t)((mt * KK + kk) * 256)), bf[kk]);
                }
                // element e = 8 b + lr: Gauss-Legendre fold of exp(mu + sigma a) and the element's flux term
                const double coef = h * ((double)(ne - (8 * b + lr)) - 0.5);
                {
                    const double k0 = exp(mu + sigma * c[0][0]) * inv2h, k1 = exp(mu + sigma * c[1][0]) * inv2h,
                                 k2 = exp(mu + sigma * c[2][0]) * inv2h;
                    u0 += coef / (k0 * w0 + k1 * w1 + k2 * w0);
                }
                {
                    const double k0 = exp(mu + sigma * c[0][1]) * inv2h, k1 = exp(mu + sigma * c[1][1]) * inv2h,
                                 k2 = exp(mu + sigma * c[2][1]) * inv2h;
                    u1 += coef / (k0 * w0 + k1 * w1 + k2 * w0);
                }
            }
            // stage s is consumed: the last warp to get here refills it with block j + S of the stream
            __syncwarp();
            unsigned int last = 0;
            if (lane == 0) {
                __threadfence_block();
                last = (atomicAdd(&done_cnt[s], 1u) == NW - 1) ? 1u : 0u;
                if (last) {
                    done_cnt[s] = 0;
                    if (j + S < total) load_block(j + S);
                }
            }
        }
        if (live) {
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
                u0 += __shfl_xor_sync(0xffffffffu, u0, o);
                u1 += __shfl_xor_sync(0xffffffffu, u1, o);
            }
            if (lr == 0) {
                const int64_t ip = i0 + 2 * lc;
                if (ip < n) g[ip] = u_max - u0;
                if (ip + 1 < n) g[ip + 1] = u_max - u1;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// host side: one packed-table buffer per device, re-packed on every call (1.9 MB, a few microseconds)
// ---------------------------------------------------------------------------------------
static double* g_pack[64] = {nullptr};
static size_t g_pack_doubles[64] = {0};

bool ree_pde_mma_supported(const ree_lsf* l) {
    static int forced = -1;
    if (forced < 0) {  // REE_PDE_SIMT=1: test hook, keeps the thread-per-particle Thomas kernel (both are CUDA paths)
        const char* e = getenv("REE_PDE_SIMT");
        forced = (e && e[0] == '1') ? 1 : 0;
    }
    return !forced && l->d >= 1 && l->d <= 152 && l->n_ele >= 8 && l->n_ele % 8 == 0;
}

int ree_pde_mma_launch(const ree_lsf* l, const double* x, int64_t n, int64_t ld, double* g, cudaStream_t st) {
    constexpr int KK = 38, NW = 12, S = 4;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return ree_set_error(REE_E_BADARG, "bad device");
    const int nb = l->n_ele / 8;
    const size_t need = (size_t)nb * PdeDims<KK>::BLOCK_DOUBLES;
    if (g_pack_doubles[dev] < need) {
        if (g_pack[dev]) cudaFree(g_pack[dev]);
        g_pack[dev] = nullptr;
        g_pack_doubles[dev] = 0;
        cudaError_t err = cudaMalloc(&g_pack[dev], need * sizeof(double));
        if (err != cudaSuccess) return ree_set_error((int)err, "diffusion LSF: cannot allocate the packed mode table");
        g_pack_doubles[dev] = need;
    }
    k_pde_pack<<<148, 256, 0, st>>>(l->coef, l->d, nb, KK, g_pack[dev]);
    int rc = ree_check_launch("k_pde_pack");
    if (rc) return rc;
    const int sms = ree_grid_ctas() / 2;
    const int64_t nct = ((n + 7) / 8 + NW - 1) / NW;
    const int grid = (int)(nct < sms ? nct : sms);
    constexpr size_t smem = (size_t)S * PdeDims<KK>::BLOCK_BYTES;
    static_assert(smem <= 227 * 1024, "shared memory budget");
    auto kern = k_pde_mma<KK, NW, S>;
    REE_OPT_IN_SMEM(kern, smem);
    kern<<<grid, NW * 32, smem, st>>>(g_pack[dev], nb, x, n, ld, l->d, l->p0, l->p1, l->offset, g);
    return ree_check_launch("k_pde_mma");
}
