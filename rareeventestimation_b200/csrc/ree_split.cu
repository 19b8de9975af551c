// Split DMMA path of a CBREE iteration for the fp64-bound dimensions (56 <= d <= 103): three homogeneous kernels
// instead of the two fused sweeps of ree_mma.cu.
//
//   k_xform<1>   X' = t X + m 1^T + F Z^T, g', e'        (transform only: no tile buffer, no CTA barrier; every warp
//                                                          owns 8 particles and runs on its own; Z from Philox)
//   k_gram2      both Grams of X' in one pass: sum y y^T, sum y, n  and  sum w y y^T, sum w y, sum w  (y = x' - shift,
//                w = exp(a - wmax) from the a = beta*l vector of ree_reduce_ess): pure LDS + DMMA stream
//   k_xform<2>   u = Linv (x' - mean), q = |u|^2 -> log q, IS statistics   (transform only)
//
// Why split: in the fused sweeps all warps of a CTA walk through the same phases (Philox-heavy transform, epilogue,
// barrier, Gram) in lock step, so the fp64 pipe idles whenever the phase is not DMMA-dense (49 % / 65 % tensor-active,
// profiles/r1_a1_*).  Here the integer-heavy Philox work of one warp overlaps the DMMAs of the 15 others, and the
// Gram kernel feeds the pipe from 16 independent accumulator chains per warp.  Price: X' is read three times instead
// of twice (32 GB instead of 24 GB of HBM traffic per iteration at N=1e7, d=100 -- 4.9 ms of a >= 12.5 ms DMMA bound).
#include "mma_common.cuh"


struct XfArgs {
    const double* x;  // MODE 1: X (input) ; MODE 2: X'
    int64_t n, ld;
    int d;
    const double* A;  // F (MODE 1) or Linv (MODE 2), row-major d x d
    // MODE 1
    double* x_new;
    double t;
    const double* m_noise;
    const double* z;
    uint64_t seed, draw;
    int64_t first;
    int lsf_mode;  // 0: not fused, 1: affine coef, 2: -sum(x)/sqrt(d) + offset
    const double* coef;
    double offset;
    double* g_new;
    double* e_new;
    // MODE 2
    const double* g;
    const double* e;
    const double* mean;
    const double* logdet;
    double* logq;
    double* small;  // per-CTA ExpSums
};

struct GenStage {  // rows 8q+lc and 8q+4+lc of the warp's staging block at particle column lr, minus the row mean
    uint32_t addr;       // &stg[lc*8 + lr]
    uint32_t mean_addr;  // &mean[lc]
    bool valid;
    __device__ __forceinline__ void operator()(int q, double& b0, double& b1) const {
        const double x0 = lds64(addr + (uint32_t)(8 * q * 64)), x1 = lds64(addr + (uint32_t)((8 * q + 4) * 64));
        const double m0 = lds64(mean_addr + 64 * q), m1 = lds64(mean_addr + 64 * q + 32);
        b0 = valid ? x0 - m0 : 0.0;  // rows >= d pair with zero columns of Linv (staging rows >= d stay zero)
        b1 = valid ? x1 - m1 : 0.0;
    }
};

// ---------------------------------------------------------------------------------------
// transform kernel: one warp = 8 particles x all dimensions, persistent over warp tiles
// ---------------------------------------------------------------------------------------
// PGB: Philox + Box-Muller chains generated together (instruction-level parallelism inside a warp)
template <int MODE, int NTB, int NW, bool EXACT, bool TRI, int PGB, bool TMAX>
__global__ void __launch_bounds__(NW * 32, 1) k_xform(const XfArgs a, const __grid_constant__ CUtensorMap tmap) {
    constexpr int MT = NW * 32;
    constexpr int ROWS = 8 * NTB;
    constexpr int STG = ROWS * 8;  // doubles of one warp's staging block: [row][8 particles]
    extern __shared__ __align__(1024) double sm[];
    const int d = a.d, ntd = (d + 7) >> 3;
    double* Lf = sm;
    double* stg = Lf + lf_doubles<NTB, TRI>();
    double* vm = stg + NW * STG;  // m_noise (MODE 1) / mean (MODE 2)
    double* vc = vm + 8 * NTB;    // LSF coefficients (MODE 1)
    double* bm = vc + 8 * NTB;    // Box-Muller tables (MODE 1, Philox noise)
    __shared__ ExpSums sh[32];
    __shared__ __align__(8) unsigned long long xbars[NW];  // "the X tile of warp w has landed" (TMAX)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, lr = lane >> 2, lc = lane & 3;
    if (TMAX && tid < NW) mbar_init(smem_u32(&xbars[0]) + 8u * (uint32_t)tid, 1);
    if (TMAX) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");

    pack_matrix<NTB, TRI, MT>(Lf, a.A, d);
    for (int j = tid; j < 8 * NTB; j += MT) {
        if (MODE == 1) {
            vm[j] = (j < d) ? a.m_noise[j] : 0.0;
            vc[j] = (j < d) ? (a.lsf_mode == 1 ? a.coef[j] : 1.0) : 0.0;
        } else {
            vm[j] = (j < d) ? a.mean[j] : 0.0;
        }
    }
    for (int idx = tid; idx < NW * STG; idx += MT) stg[idx] = 0.0;
    if (MODE == 1 && !a.z)
        bm_stage_tables(bm, tid, MT);
    __syncthreads();

    const uint32_t lf_lane = smem_u32(Lf) + lane * 8;
    const uint32_t stg_w = smem_u32(stg + warp * STG);
    const uint32_t own = stg_w + (uint32_t)(lr * 64 + lc * 16);  // rows 8mt+lr, columns 2lc, 2lc+1
    const uint32_t vm_addr = smem_u32(vm);
    const double e_const = (double)d * REE_LOG_2PI;
    const double lq_const = (MODE == 2) ? e_const + a.logdet[0] : 0.0;
    ExpSums is_acc = expsums_empty();
    const int64_t nwt = (a.n + 7) >> 3;
    const int64_t stride = (int64_t)gridDim.x * NW;

    const uint32_t xbar = smem_u32(&xbars[warp]);
    auto prefetch = [&](int64_t wt) {
        if constexpr (TMAX) {  // one TMA box copy {8 particles, all rows} per warp tile (rows >= d arrive as zeros)
            __syncwarp();
            if (lane == 0) {
                fence_proxy_async();
                mbar_expect_tx(xbar, (uint32_t)(STG * 8));
                tma_load_2d(stg_w, &tmap, (int)(wt * 8), 0, xbar);
            }
        } else {  // every lane fetches the 16-byte chunks (row 8mt+lr, columns 2lc..2lc+1)
            const double* src = a.x + wt * 8 + 2 * lc;  // wt*8 + 7 < ld (ld is a multiple of 8, ld >= n)
#pragma unroll
            for (int mt = 0; mt < NTB; ++mt) {
                const int row = 8 * mt + lr;
                if (row < d) cp_async16(own + (uint32_t)(mt * 512), src + (int64_t)row * a.ld);
            }
        }
    };
    auto landed = [&](uint32_t parity) {
        if constexpr (TMAX) mbar_wait(xbar, parity);
        else cp_async_wait_all();
    };
    int64_t wt = (int64_t)blockIdx.x * NW + warp;
    if (wt < nwt) prefetch(wt);

    uint32_t xk = 0;
    for (; wt < nwt; wt += stride, ++xk) {
        const int64_t i0 = wt * 8;
        const int64_t ip = i0 + 2 * lc;  // first of this lane's two particle columns
        double c[NTB][2];
#pragma unroll
        for (int mt = 0; mt < NTB; ++mt) c[mt][0] = c[mt][1] = 0.0;

        if (MODE == 1) {
            // ---- F z on the tensor pipe: c[mt] = rows 8mt.., particles i0.. ------------------------------
            if (a.z) {
                GenGlobal gen;
                const int64_t ipb = i0 + lr;
                gen.base = (ipb < a.n) ? a.z + (int64_t)lc * a.ld + ipb : nullptr;
                gen.ld = a.ld; gen.d = d; gen.lc = lc;
                transform_tiles<NTB, TRI, EXACT, 3>(lf_lane, ntd, gen, c);
            } else {
                GenPhilox gen;
                gen.seed = a.seed; gen.draw = a.draw; gen.lc = (uint32_t)lc; gen.tab.addr = smem_u32(bm);
                gen.particle = (uint64_t)(a.first + i0 + lr);
                transform_tiles<NTB, TRI, EXACT, PGB>(lf_lane, ntd, gen, c);
            }
            landed(xk & 1u);  // (cp.async: own elements only, no warp barrier needed)
            double se0 = 0.0, se1 = 0.0, sg0 = 0.0, sg1 = 0.0;
#pragma unroll
            for (int mt = 0; mt < NTB; ++mt) {
                if (EXACT || mt < ntd) {
                    const int row = 8 * mt + lr;
                    if (row < d) {
                        const double2 xv = lds128(own + (uint32_t)(mt * 512));
                        const double mrow = vm[row], crow = vc[row];
                        // t*X + (m_noise + Z F^T): grouping of solver.py:667
                        double2 xn;
                        xn.x = __dadd_rn(__dmul_rn(a.t, xv.x), __dadd_rn(mrow, c[mt][0]));
                        xn.y = __dadd_rn(__dmul_rn(a.t, xv.y), __dadd_rn(mrow, c[mt][1]));
                        *reinterpret_cast<double2*>(a.x_new + (int64_t)row * a.ld + ip) = xn;
                        se0 = fma(xn.x, xn.x, se0);
                        se1 = fma(xn.y, xn.y, se1);
                        sg0 = fma(crow, xn.x, sg0);
                        sg1 = fma(crow, xn.y, sg1);
                    }
                }
            }
            if (wt + stride < nwt) prefetch(wt + stride);  // consumed after the next transform
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
                se0 += __shfl_xor_sync(0xffffffffu, se0, o);
                se1 += __shfl_xor_sync(0xffffffffu, se1, o);
                sg0 += __shfl_xor_sync(0xffffffffu, sg0, o);
                sg1 += __shfl_xor_sync(0xffffffffu, sg1, o);
            }
            if (lr == 0) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (ip + h < a.n) {
                        const double s = h ? se1 : se0, gsum = h ? sg1 : sg0;
                        if (a.e_new) a.e_new[ip + h] = 0.5 * (e_const + s);
                        if (a.lsf_mode)
                            a.g_new[ip + h] = (a.lsf_mode == 1)
                                                  ? __dadd_rn(gsum, a.offset)
                                                  : __dadd_rn(__ddiv_rn(-gsum, __dsqrt_rn((double)d)), a.offset);
                    }
                }
            }
        } else {
            landed(xk & 1u);
            __syncwarp();  // the warp's block has landed (cp.async: every lane waited for its own copies)
            GenStage gen;
            gen.addr = stg_w + (uint32_t)((lc * 8 + lr) * 8);
            gen.mean_addr = vm_addr + lc * 8;
            gen.valid = (i0 + lr) < a.n;
            transform_tiles<NTB, TRI, EXACT, 2>(lf_lane, ntd, gen, c);
            __syncwarp();  // all B-fragment reads precede the refill of the block
            if (wt + stride < nwt) prefetch(wt + stride);
            // ---- q = |u|^2, log q, IS statistics (solver.py:672-674, 683-686, 843-846) -------------------
            double q0 = 0.0, q1 = 0.0;
#pragma unroll
            for (int mt = 0; mt < NTB; ++mt) {
                q0 = fma(c[mt][0], c[mt][0], q0);
                q1 = fma(c[mt][1], c[mt][1], q1);
            }
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
                q0 += __shfl_xor_sync(0xffffffffu, q0, o);
                q1 += __shfl_xor_sync(0xffffffffu, q1, o);
            }
            // one lane per particle: particle p = 2*lc' + h lives in the lanes with lc == lc'
            const int srcl = (lane >> 1) & 3;
            const double qa = __shfl_sync(0xffffffffu, q0, srcl), qb = __shfl_sync(0xffffffffu, q1, srcl);
            const int64_t i = i0 + lane;
            if (lane < 8 && i < a.n) {
                const double qq = (lane & 1) ? qb : qa;
                const double lq = -0.5 * (lq_const + qq);
                if (a.logq) a.logq[i] = lq;
                const double gi = a.g[i], ei = a.e[i];
                const double r = __dsub_rn(-ei, lq);  // -e_fun - log_pdf  (solver.py:684)
                if (isfinite(r)) expsums_add(is_acc, r, gi <= 0.0 ? 1.0 : 0.0);
            }
        }
    }
    if (MODE == 2) {
        ExpSums blk = expsums_block_reduce(is_acc, sh);
        if (tid == 0) {
            double* p = a.small + 4 * (size_t)blockIdx.x;
            p[0] = blk.m; p[1] = blk.s1; p[2] = blk.s2; p[3] = blk.cnt;
        }
    }
}

// ---------------------------------------------------------------------------------------
// warp-specialised transform (Philox noise): NC consumer warps run the LDS + DMMA stream of k_xform<1>, NC producer
// warps run Philox + Box-Muller and hand finished B fragments over through shared memory.
//
// Why: a warp issues in order.  In k_xform<1> every warp interleaves ~2500 integer / fp64 instructions of noise
// generation with 182 DMMAs per tile; whenever its next instruction is a DMMA and the tensor pipe is busy, the ~14
// independent Philox instructions behind it wait as well (math_pipe_throttle is the top stall, pipe 54 % busy, issue
// 41 %: profiles/r1_split_kernels_ncu.txt).  With the two instruction streams in different warps the scheduler always
// has a producer warp to issue while the consumers queue on the pipe.
//
// Hand-over: per consumer two noise buffers of NTB k-pairs x 32 lanes x (b0, b1); the producer lane with the same lane
// id computes exactly the consumer lane's fragment values (Philox counter = (particle lr, pair 4q + lc)), so both sides
// use conflict-free 16-byte accesses at lane * 16.  full/empty mbarriers per buffer, one elected arrival per warp.
// ---------------------------------------------------------------------------------------
struct GenNoise {  // B fragments of k-pair q from the consumer's noise buffer
    uint32_t addr;  // buffer + lane * 16
    __device__ __forceinline__ void operator()(int q, double& b0, double& b1) const {
        const double2 v = lds128(addr + (uint32_t)(q * 512));
        b0 = v.x;
        b1 = v.y;
    }
};

// NPC producer warps per consumer share one tile (k-pairs q = h, h + NPC, ...): a producer's fp64 instructions queue
// behind the 16-cycle DMMAs of the shared pipe, so one producer per consumer is latency-bound at ~2.4 consumer tile
// times.  With NPC > 1 the CTA has more than 512 threads; setmaxnreg moves registers from the producers (RP) to the
// consumers (RC), whose 2 * NTB accumulators + fragment prefetch do not fit the launch-time allocation.
// TMAX: a consumer's X tile (8 particles x all rows) arrives by ONE TMA box copy issued by its lane 0 instead of 13
// cp.async per lane with their 64-bit address arithmetic
template <int NTB, int NC, int NPC, bool EXACT, int PGB, int RC, int RP, bool TMAX>
__global__ void __launch_bounds__((1 + NPC) * NC * 32, 1) k_xform_ws(const XfArgs a, const __grid_constant__ CUtensorMap tmap) {
    constexpr int MT = (1 + NPC) * NC * 32;
    constexpr int ROWS = 8 * NTB;
    constexpr int STG = ROWS * 8;   // doubles of one consumer's X staging block: [row][8 particles]
    constexpr int NZ = NTB * 64;    // doubles of one noise buffer: [k-pair][lane][2]
    extern __shared__ __align__(1024) double sm[];
    const int d = a.d, ntd = (d + 7) >> 3;
    double* Lf = sm;
    double* nz = Lf + lf_doubles<NTB, true>();
    double* stg = nz + NC * 2 * NZ;
    double* vm = stg + NC * STG;
    double* vc = vm + 8 * NTB;
    double* bm = vc + 8 * NTB;
    __shared__ __align__(8) unsigned long long bars[NC][2][2];  // [consumer][buffer][0: full, 1: empty]
    __shared__ __align__(8) unsigned long long xbars[NC];       // "the X tile of consumer c has landed" (TMAX)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, lr = lane >> 2, lc = lane & 3;

    pack_matrix<NTB, true, MT>(Lf, a.A, d);
    for (int j = tid; j < 8 * NTB; j += MT) {
        vm[j] = (j < d) ? a.m_noise[j] : 0.0;
        vc[j] = (j < d) ? (a.lsf_mode == 1 ? a.coef[j] : 1.0) : 0.0;
    }
    for (int idx = tid; idx < NC * STG; idx += MT) stg[idx] = 0.0;
    bm_stage_tables(bm, tid, MT);
    if (tid < NC * 4) mbar_init(smem_u32(&bars[0][0][0]) + 8u * (uint32_t)tid, (tid & 1) ? 1 : NPC);
    if (tid < NC) mbar_init(smem_u32(&xbars[0]) + 8u * (uint32_t)tid, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const int64_t nwt = (a.n + 7) >> 3;
    const int64_t stride = (int64_t)gridDim.x * NC;
    const int c = warp % NC, h = warp / NC - 1;  // consumer c; producer h = 0 .. NPC-1 of consumer c
    const uint32_t bar_c = smem_u32(&bars[c][0][0]);           // + 16 b (+ 8 for empty)
    const uint32_t nz_c = smem_u32(nz + c * 2 * NZ) + lane * 16;  // + b * NZ * 8

    if (warp >= NC) {
        // ---------------- producer: Philox + Box-Muller for consumer c ----------------
        if constexpr (RP > 0) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(RP));
        BmTableShared tab;
        tab.addr = smem_u32(bm);
        int k = 0;
        for (int64_t wt = (int64_t)blockIdx.x * NC + c; wt < nwt; wt += stride, ++k) {
            const uint32_t b = (uint32_t)k & 1u, ph = ((uint32_t)k >> 1) & 1u;
            const uint32_t buf = nz_c + b * (uint32_t)(NZ * 8);
            mbar_wait(bar_c + 16u * b + 8u, ph ^ 1u);  // buffer b is free (passes at once the first two times)
            const uint64_t particle = (uint64_t)(a.first + wt * 8 + lr);
            constexpr int NQ = (NTB + NPC - 1) / NPC;  // k-pairs q = h + NPC * i, i < NQ
#pragma unroll
            for (int i0 = 0; i0 < NQ; i0 += PGB) {
                double z0[PGB], z1[PGB];
#pragma unroll
                for (int j = 0; j < PGB; ++j) {
                    const int q = h + NPC * (i0 + j);
                    if (i0 + j < NQ && q < (EXACT ? NTB : ntd))
                        philox_normal_pair(a.seed, particle, (uint32_t)(4 * q) + (uint32_t)lc, a.draw, tab, z0[j], z1[j]);
                }
#pragma unroll
                for (int j = 0; j < PGB; ++j) {
                    const int q = h + NPC * (i0 + j);
                    if (i0 + j < NQ && q < (EXACT ? NTB : ntd))
                        sts128(buf + (uint32_t)(q * 512), make_double2(z0[j], z1[j]));
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_c + 16u * b);
        }
        return;
    }

    // ---------------- consumer: F z on the tensor pipe, update, g', e' ----------------
    // Tile k's accumulators START as t x + m_noise (one DFMA per element while the warp waits for its noise), the DMMAs
    // add F z on top; the epilogue stores the rows and accumulates |x'|^2 and coef . x' in two interleaved chains each
    // (under the other warps' DMMA streams a dependent fp64 instruction costs ~28 cycles, tools/fp64_mix.cu).
    if constexpr (RC > 0) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(RC));
    const uint32_t lf_lane = smem_u32(Lf) + lane * 8;
    const uint32_t stg_w = smem_u32(stg + warp * STG);
    const uint32_t own = stg_w + (uint32_t)(lr * 64 + lc * 16);  // rows 8mt+lr, columns 2lc, 2lc+1
    const uint32_t vm_lane = smem_u32(vm) + (uint32_t)lr * 8u, vc_lane = smem_u32(vc) + (uint32_t)lr * 8u;
    const double e_const = (double)d * REE_LOG_2PI;
    const double tt = a.t;
    const uint32_t xbar = smem_u32(&xbars[warp]);
    auto prefetch = [&](int64_t wt) {
        if constexpr (TMAX) {
            // the box {8 particles, all rows} of the ensemble's tensor map lands as [row][8] in the staging block; rows
            // >= d arrive as zeros.  The proxy fence orders this warp's reads of the previous tile before the refill.
            __syncwarp();
            if (lane == 0) {
                fence_proxy_async();
                mbar_expect_tx(xbar, (uint32_t)(STG * 8));
                tma_load_2d(stg_w, &tmap, (int)(wt * 8), 0, xbar);
            }
        } else {
            const double* src = a.x + wt * 8 + 2 * lc;
#pragma unroll
            for (int mt = 0; mt < NTB; ++mt) {
                const int row = 8 * mt + lr;
                if (row < d) cp_async16(own + (uint32_t)(mt * 512), src + (int64_t)row * a.ld);
            }
        }
    };
    int64_t wt = (int64_t)blockIdx.x * NC + warp;
    if (wt < nwt) prefetch(wt);
    int k = 0;
    for (; wt < nwt; wt += stride, ++k) {
        const int64_t ip = wt * 8 + 2 * lc;
        const uint32_t b = (uint32_t)k & 1u, ph = ((uint32_t)k >> 1) & 1u;
        double c2[NTB][2];
        if constexpr (TMAX) mbar_wait(xbar, (uint32_t)k & 1u);
        else cp_async_wait_all();  // own elements only: no warp barrier needed
#pragma unroll
        for (int mt = 0; mt < NTB; ++mt) {
            const int row = 8 * mt + lr;
            c2[mt][0] = c2[mt][1] = 0.0;
            if ((EXACT || mt < ntd) && row < d) {
                const double2 xv = lds128(own + (uint32_t)(mt * 512));
                const double mrow = lds64(vm_lane + (uint32_t)(mt * 64));
                c2[mt][0] = fma(tt, xv.x, mrow);
                c2[mt][1] = fma(tt, xv.y, mrow);
            }
        }
        if (wt + stride < nwt) prefetch(wt + stride);  // (this lane's reads of the block precede its refill)
        mbar_wait(bar_c + 16u * b, ph);
        const uint32_t nzb = nz_c + b * (uint32_t)(NZ * 8);
        double2 bq = lds128(nzb);
#pragma unroll
        for (int q = 0; q < NTB; ++q) {
            if (EXACT || q < ntd) {
                double2 bn = bq;
                if (q + 1 < NTB && (EXACT || q + 1 < ntd)) bn = lds128(nzb + (uint32_t)((q + 1) * 512));
#pragma unroll
                for (int mt = q; mt < NTB; ++mt)
                    if (EXACT || mt < ntd)
                        dmma(c2[mt][0], c2[mt][1], lds64(lf_lane + lf_tile<NTB, true>(mt, 2 * q) * 256), bq.x);
#pragma unroll
                for (int mt = q; mt < NTB; ++mt)
                    if (EXACT || mt < ntd)
                        dmma(c2[mt][0], c2[mt][1], lds64(lf_lane + lf_tile<NTB, true>(mt, 2 * q + 1) * 256), bq.y);
                bq = bn;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_c + 16u * b + 8u);  // every fragment of the buffer has been read
        double se[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, sg[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
        for (int mt = 0; mt < NTB; ++mt) {
            const int row = 8 * mt + lr;
            if ((EXACT || mt < ntd) && row < d) {
                const double crow = lds64(vc_lane + (uint32_t)(mt * 64));
                const double x0 = c2[mt][0], x1 = c2[mt][1];
                *reinterpret_cast<double2*>(a.x_new + (int64_t)row * a.ld + ip) = make_double2(x0, x1);
                se[mt & 1][0] = fma(x0, x0, se[mt & 1][0]);
                se[mt & 1][1] = fma(x1, x1, se[mt & 1][1]);
                sg[mt & 1][0] = fma(crow, x0, sg[mt & 1][0]);
                sg[mt & 1][1] = fma(crow, x1, sg[mt & 1][1]);
            }
        }
        double se0 = se[0][0] + se[1][0], se1 = se[0][1] + se[1][1];
        double sg0 = sg[0][0] + sg[1][0], sg1 = sg[0][1] + sg[1][1];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
            se0 += __shfl_xor_sync(0xffffffffu, se0, o);
            se1 += __shfl_xor_sync(0xffffffffu, se1, o);
            sg0 += __shfl_xor_sync(0xffffffffu, sg0, o);
            sg1 += __shfl_xor_sync(0xffffffffu, sg1, o);
        }
        if (lr == 0) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (ip + h < a.n) {
                    const double s = h ? se1 : se0, gsum = h ? sg1 : sg0;
                    if (a.e_new) a.e_new[ip + h] = 0.5 * (e_const + s);
                    if (a.lsf_mode)
                        a.g_new[ip + h] = (a.lsf_mode == 1)
                                              ? __dadd_rn(gsum, a.offset)
                                              : __dadd_rn(__ddiv_rn(-gsum, __dsqrt_rn((double)d)), a.offset);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// Gram kernel: unweighted and weighted shifted Grams of an ensemble in one pass
// ---------------------------------------------------------------------------------------
struct GramArgs {
    const double* x;
    int64_t n, ld;
    int d;
    const double* shift;  // d doubles or NULL (0)
    const double* g;      // failure count (or NULL)
    const double* a;      // log-weights beta*l (WEIGHTED)
    const double* wmax;   // device scalar (WEIGHTED)
    double* part_u;       // [(cta*KS + grp)][T][64]
    double* part_w;
    double* small;        // per-CTA failure count
    GramPlan plan;
};

struct GramRegs2 {
    uint32_t offR, offC;  // byte offsets (inside a tile buffer) of this lane's element of the first row / col tile
    int flags;            // 15: off-diagonal 2x2, 13: diagonal 2x2, 3: one row x two cols, 1: single tile, 0: none
};

// one k-step (4 particles) of one 8x8 output tile.  A k-pair is issued as all first k-steps of the macro tile, then
// all second ones: the two DMMAs of one accumulator are 3-7 independent DMMAs apart instead of back to back.
#define DX(acc, a, b) dmma(acc[0], acc[1], (a).x, (b).x)
#define DY(acc, a, b) dmma(acc[0], acc[1], (a).y, (b).y)

// Gram of UPG k-pairs of a shared tile into the warp's macro tiles; macro-outer / k-inner so that every branch on the
// macro type is taken once per tile and the k loop is straight-line code with immediate offsets.
template <int MAXM, int S, int UPG, bool WEIGHTED>
__device__ __forceinline__ void gram2_accumulate(uint32_t buf, uint32_t wts, const GramRegs2 (&gr)[MAXM],
                                                 double (&au)[MAXM][4][2], double (&aw)[WEIGHTED ? MAXM : 1][4][2]) {
    constexpr uint32_t R1 = 8 * S * 8;  // second tile of a macro tile: 8 rows further down
#pragma unroll
    for (int m = 0; m < MAXM; ++m) {
        const int f = gr[m].flags;
        if (f == 0) continue;
        const uint32_t ra = buf + gr[m].offR, ca = buf + gr[m].offC;
        if (f == 15) {  // off-diagonal 2x2
#pragma unroll
            for (int u = 0; u < UPG; ++u) {
                const double2 a0 = lds128(ra + u * 64), a1 = lds128(ra + u * 64 + R1);
                const double2 b0 = lds128(ca + u * 64), b1 = lds128(ca + u * 64 + R1);
                double2 a0w, a1w;
                if (WEIGHTED) {
                    const double2 w = lds128(wts + u * 64);
                    a0w = make_double2(a0.x * w.x, a0.y * w.y);
                    a1w = make_double2(a1.x * w.x, a1.y * w.y);
                }
                DX(au[m][0], a0, b0); DX(au[m][1], a0, b1); DX(au[m][2], a1, b0); DX(au[m][3], a1, b1);
                if (WEIGHTED) { DX(aw[WEIGHTED ? m : 0][0], a0w, b0); DX(aw[WEIGHTED ? m : 0][1], a0w, b1); DX(aw[WEIGHTED ? m : 0][2], a1w, b0); DX(aw[WEIGHTED ? m : 0][3], a1w, b1); }
                DY(au[m][0], a0, b0); DY(au[m][1], a0, b1); DY(au[m][2], a1, b0); DY(au[m][3], a1, b1);
                if (WEIGHTED) { DY(aw[WEIGHTED ? m : 0][0], a0w, b0); DY(aw[WEIGHTED ? m : 0][1], a0w, b1); DY(aw[WEIGHTED ? m : 0][2], a1w, b0); DY(aw[WEIGHTED ? m : 0][3], a1w, b1); }
            }
        } else if (f == 13) {  // diagonal 2x2: tiles (0,0), (1,0), (1,1); B fragments are the A fragments
#pragma unroll
            for (int u = 0; u < UPG; ++u) {
                const double2 a0 = lds128(ra + u * 64), a1 = lds128(ra + u * 64 + R1);
                double2 a0w, a1w;
                if (WEIGHTED) {
                    const double2 w = lds128(wts + u * 64);
                    a0w = make_double2(a0.x * w.x, a0.y * w.y);
                    a1w = make_double2(a1.x * w.x, a1.y * w.y);
                }
                DX(au[m][0], a0, a0); DX(au[m][2], a1, a0); DX(au[m][3], a1, a1);
                if (WEIGHTED) { DX(aw[WEIGHTED ? m : 0][0], a0w, a0); DX(aw[WEIGHTED ? m : 0][2], a1w, a0); DX(aw[WEIGHTED ? m : 0][3], a1w, a1); }
                DY(au[m][0], a0, a0); DY(au[m][2], a1, a0); DY(au[m][3], a1, a1);
                if (WEIGHTED) { DY(aw[WEIGHTED ? m : 0][0], a0w, a0); DY(aw[WEIGHTED ? m : 0][2], a1w, a0); DY(aw[WEIGHTED ? m : 0][3], a1w, a1); }
            }
        } else if (f == 3) {  // one row tile, two column tiles
#pragma unroll
            for (int u = 0; u < UPG; ++u) {
                const double2 a0 = lds128(ra + u * 64);
                const double2 b0 = lds128(ca + u * 64), b1 = lds128(ca + u * 64 + R1);
                double2 a0w;
                if (WEIGHTED) {
                    const double2 w = lds128(wts + u * 64);
                    a0w = make_double2(a0.x * w.x, a0.y * w.y);
                }
                DX(au[m][0], a0, b0); DX(au[m][1], a0, b1);
                if (WEIGHTED) { DX(aw[WEIGHTED ? m : 0][0], a0w, b0); DX(aw[WEIGHTED ? m : 0][1], a0w, b1); }
                DY(au[m][0], a0, b0); DY(au[m][1], a0, b1);
                if (WEIGHTED) { DY(aw[WEIGHTED ? m : 0][0], a0w, b0); DY(aw[WEIGHTED ? m : 0][1], a0w, b1); }
            }
        } else {  // f == 1: single tile (two k-pairs per round so that the accumulator chain has company)
#pragma unroll
            for (int u = 0; u < UPG; ++u) {
                const double2 a0 = lds128(ra + u * 64);
                const double2 b0 = lds128(ca + u * 64);
                double2 a0w;
                if (WEIGHTED) {
                    const double2 w = lds128(wts + u * 64);
                    a0w = make_double2(a0.x * w.x, a0.y * w.y);
                }
                DX(au[m][0], a0, b0);
                if (WEIGHTED) DX(aw[WEIGHTED ? m : 0][0], a0w, b0);
                DY(au[m][0], a0, b0);
                if (WEIGHTED) DY(aw[WEIGHTED ? m : 0][0], a0w, b0);
            }
        }
    }
}
#undef DX
#undef DY

template <int MAXM>
__device__ __forceinline__ void gram2_store(double* __restrict__ part, const GramPlan& plan, int wi, int nmac, int nta,
                                            int lr, int lc, const double (&acc)[MAXM][4][2]) {
#pragma unroll
    for (int m = 0; m < MAXM; ++m) {
        if (m < nmac) {
            const int code = plan.mac[wi][m];
            const int mr = code & 0xff, mc = code >> 8;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int mt = 2 * mr + (q >> 1), nt = 2 * mc + (q & 1);
                if (mt < nta && nt <= mt)
                    *reinterpret_cast<double2*>(part + (size_t)tri_index(mt, nt) * 64 + lr * 8 + 2 * lc) =
                        make_double2(acc[m][q][0], acc[m][q][1]);
            }
        }
    }
}

// WM: 0 unweighted Gram only; 1 unweighted + weighted Gram from one tile (two accumulator sets); 2 weighted Gram only,
// with the tile rewritten as y sqrt(w) (one accumulator set: the d <= 151 bucket has no registers for two)
template <int NTB, int NW, int MAXM, int KS, bool EXACT, int WM>
__global__ void __launch_bounds__(NW * 32, 1) k_gram2(const GramArgs a, const __grid_constant__ CUtensorMap tmap) {
    constexpr bool WEIGHTED = WM != 0;
    constexpr int MT = NW * 32;
    constexpr int P = NW * 8;   // particles per CTA tile
    constexpr int S = P + 8;    // row stride of the shared tile (8 mod 16 doubles)
    constexpr int ROWS = 8 * NTB;
    constexpr uint32_t BUF_BYTES = ROWS * S * 8;
    constexpr int WG = NW / KS, UPG = P / 8 / KS;
    extern __shared__ __align__(1024) double sm[];
    const int d = a.d, ntd = (d + 7) >> 3, nta = (d + 8) >> 3;
    double* Ys = sm;                 // two tile buffers (TMA destinations: 128-byte aligned)
    double* wt = Ys + 2 * ROWS * S;  // two weight vectors
    double* vs = wt + 2 * P;         // shift
    __shared__ double red[32];
    __shared__ __align__(8) unsigned long long full_bar[2];   // "the rows of tile buffer b have landed" (bulk copies)
    __shared__ __align__(8) unsigned long long ready_bar[2];  // "every warp has prepared its part of buffer b"
    __shared__ unsigned int done_cnt[2];                      // warps that finished the Gram of buffer b
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, lr = lane >> 2, lc = lane & 3;
    const int grp = warp / WG, wi = warp % WG;

    for (int j = tid; j < 8 * NTB; j += MT) vs[j] = (j < d && a.shift) ? a.shift[j] : 0.0;
    for (int idx = tid; idx < 2 * ROWS * S + 2 * P; idx += MT) Ys[idx] = 0.0;
    const uint32_t ys_addr = smem_u32(Ys), wt_addr = smem_u32(wt);
    const uint32_t bar0 = smem_u32(&full_bar[0]), rdy0 = smem_u32(&ready_bar[0]);
    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        mbar_init(rdy0, NW);
        mbar_init(rdy0 + 8, NW);
        done_cnt[0] = done_cnt[1] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    GramRegs2 gr[MAXM];
    const int nmac = a.plan.nmac[wi];
#pragma unroll
    for (int m = 0; m < MAXM; ++m) {
        const bool live = m < nmac;
        const int code = live ? a.plan.mac[wi][m] : 0;
        const int mr = code & 0xff, mc = code >> 8;
        // the k-split group's first k-pair is folded into the offsets
        gr[m].offR = (uint32_t)(((16 * mr + lr) * S + 2 * lc) * 8 + grp * UPG * 64);
        gr[m].offC = (uint32_t)(((16 * mc + lr) * S + 2 * lc) * 8 + grp * UPG * 64);
        const bool r1 = 2 * mr + 1 < nta, c1 = 2 * mc + 1 < nta, diag = mr == mc;
        gr[m].flags = live ? (1 | ((c1 && !diag) ? 2 : 0) | (r1 ? 4 : 0) | ((r1 && c1) ? 8 : 0)) : 0;
    }
    double au[MAXM][4][2], aw[WM == 1 ? MAXM : 1][4][2];
#pragma unroll
    for (int m = 0; m < MAXM; ++m)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            au[m][q][0] = au[m][q][1] = 0.0;
            if (WM == 1 || m == 0) aw[WM == 1 ? m : 0][q][0] = aw[WM == 1 ? m : 0][q][1] = 0.0;
        }
    __syncthreads();

    const double wmax = WEIGHTED ? a.wmax[0] : 0.0;
    double n_fail = 0.0;
    const int pw = warp * 8;
    const int64_t ntiles = (a.n + P - 1) / P;
    const uint32_t own_off = (uint32_t)((lr * S + pw + 2 * lc) * 8);
    const uint32_t vs_addr = smem_u32(vs);

    // Tile load = ONE TMA instruction (box S columns x ROWS rows of the ensemble's tensor map; rows >= d and columns
    // >= ld arrive as zeros), issued by lane 0 of the last warp to finish reading the buffer.  The proxy fence orders the
    // generic-proxy reads of the previous tile before the async-proxy writes.
    auto load_tile = [&](int64_t i0, uint32_t buf, uint32_t bar) {
        if (lane == 0) {
            fence_proxy_async();
            mbar_expect_tx(bar, BUF_BYTES);
            tma_load_2d(buf, &tmap, (int)i0, 0, bar);
        }
    };
    // per-particle scalars of the next tile travel in registers (lanes 0..7 of every warp own one particle each)
    double a_nxt = 0.0, g_nxt = 1.0;
    auto fetch_scalars = [&](int64_t i0) {
        const int64_t i = i0 + pw + lane;
        const bool v = lane < 8 && i < a.n;
        a_nxt = (WEIGHTED && v) ? a.a[i] : 0.0;
        g_nxt = (v && a.g) ? a.g[i] : 1.0;
    };
    // Turn the landed tile into Gram operands in place: y = x - shift, zero padding, augmented row of ones; weights.
    // Every lane rewrites a disjoint set of elements (rows 8mt+lr, its warp's columns 2lc, 2lc+1).
    auto prepare = [&](int64_t i0, uint32_t buf, uint32_t wbuf, uint32_t bar, uint32_t rdy, uint32_t phase) {
        mbar_wait(bar, phase);
        const int64_t ip = i0 + pw + 2 * lc;
        const bool v0 = ip < a.n, v1 = ip + 1 < a.n;
        // weights of the warp's 8 columns (lane l < 8 owns particle pw + l), failure count
        double wv = 0.0;
        if (lane < 8) {
            const bool v = i0 + pw + lane < a.n;
            if (WEIGHTED) wv = (v && isfinite(a_nxt)) ? exp(a_nxt - wmax) : 0.0;
            if (WM == 1) asm volatile("st.shared.f64 [%0], %1;" ::"r"(wbuf + (uint32_t)(pw + lane) * 8u), "d"(wv) : "memory");
            if (WM != 2) n_fail += (v && g_nxt <= 0.0) ? 1.0 : 0.0;
        }
        // column factors of the rewrite: 1 (0 for padding), or sqrt(w) when the tile itself carries the weights
        double f0 = v0 ? 1.0 : 0.0, f1 = v1 ? 1.0 : 0.0;
        if (WM == 2) {
            const double sw = sqrt(wv);
            f0 = __shfl_sync(0xffffffffu, sw, 2 * lc);
            f1 = __shfl_sync(0xffffffffu, sw, 2 * lc + 1);
        }
        if (WM != 2 && i0 + P <= a.n) {  // full tile (all but the last one): no column selects
#pragma unroll
            for (int mt = 0; mt < NTB; ++mt) {
                if (EXACT ? (mt < NTB - 1) : (mt < ntd - 1)) {  // all 8 rows of the tile are dimensions
                    const uint32_t sa = buf + own_off + (uint32_t)(8 * mt * S * 8);
                    const double2 xv = lds128(sa);
                    const double srow = lds64(vs_addr + (uint32_t)((8 * mt + lr) * 8));
                    sts128(sa, make_double2(xv.x - srow, xv.y - srow));
                }
            }
        } else {
#pragma unroll
            for (int mt = 0; mt < NTB; ++mt) {
                if (EXACT ? (mt < NTB - 1) : (mt < ntd - 1)) {
                    const uint32_t sa = buf + own_off + (uint32_t)(8 * mt * S * 8);
                    const double2 xv = lds128(sa);
                    const double srow = lds64(vs_addr + (uint32_t)((8 * mt + lr) * 8));
                    // select, not multiply, for padding: padded lanes may hold anything
                    sts128(sa, make_double2(f0 != 0.0 ? (xv.x - srow) * f0 : 0.0, f1 != 0.0 ? (xv.y - srow) * f1 : 0.0));
                }
            }
        }
        {  // last dimension tile: rows >= d are zero, row d (if it falls into this tile) is the row of ones / sqrt(w)
            const int mt = ntd - 1, row = 8 * mt + lr;
            const uint32_t sa = buf + own_off + (uint32_t)(8 * mt * S * 8);
            const double2 xv = lds128(sa);
            const double srow = lds64(vs_addr + (uint32_t)(row * 8));
            double2 y;
            y.x = (row < d && f0 != 0.0) ? (xv.x - srow) * f0 : 0.0;
            y.y = (row < d && f1 != 0.0) ? (xv.y - srow) * f1 : 0.0;
            if (row == d) {
                y.x = f0;
                y.y = f1;
            }
            sts128(sa, y);
        }
        if (nta > ntd && lr == 0)  // d % 8 == 0: the augmented row lives in a tile of its own
            sts128(buf + (uint32_t)((d * S + pw + 2 * lc) * 8), make_double2(f0, f1));
        __syncwarp();
        if (lane == 0) mbar_arrive(rdy);  // release: this warp's part of the buffer is final
    };
    // Pipeline without CTA barriers: tile k lives in buffer k & 1.
    //   full[b]   bulk copies of the tile in buffer b have landed         (armed by the loading warp)
    //   ready[b]  all warps have prepared their part of buffer b          (count NW)
    //   done[b]   warps that finished the Gram of buffer b; the last one loads tile k+2 into it
    // A warp may run ahead of the slowest one by up to a tile.  The Gram of tile k runs in NCH chunks of k-pairs;
    // the warps of one SM sub-partition (same warp % 4) prepare tile k+1 after different chunks, so that the
    // sub-partition's tensor pipe always has DMMAs from the others.
    constexpr int NCH = (UPG % 4 == 0) ? 4 : 1, CH = UPG / NCH;
    const int slot = (NCH == 1) ? 0 : min(warp >> 2, NCH - 2);
    const int64_t t0 = blockIdx.x, tstep = gridDim.x;
    if (t0 < ntiles && warp == 0) load_tile(t0 * P, ys_addr, bar0);
    if (t0 + tstep < ntiles && warp == 1 % NW) load_tile((t0 + tstep) * P, ys_addr + BUF_BYTES, bar0 + 8);
    if (t0 < ntiles) {
        fetch_scalars(t0 * P);
        prepare(t0 * P, ys_addr, wt_addr, bar0, rdy0, 0);
    }
    int k = 0;
    for (int64_t tix = t0; tix < ntiles; tix += tstep, ++k) {
        const int b = k & 1;
        const uint32_t ph = (uint32_t)(k >> 1) & 1u, nph = (uint32_t)((k + 1) >> 1) & 1u;
        const uint32_t buf = ys_addr + (b ? BUF_BYTES : 0u), nbuf = ys_addr + (b ? 0u : BUF_BYTES);
        const uint32_t wbuf = wt_addr + (b ? (uint32_t)(P * 8) : 0u), nwbuf = wt_addr + (b ? 0u : (uint32_t)(P * 8));
        const bool more = tix + tstep < ntiles;
        const int64_t i0n = (tix + tstep) * P;
        if (more) fetch_scalars(i0n);
        mbar_wait(rdy0 + (b ? 8u : 0u), ph);  // acquire: every warp's part of tile k is in place
#pragma unroll 1
        for (int c = 0; c < NCH; ++c) {
            gram2_accumulate<MAXM, S, CH, WM == 1>(buf + (uint32_t)(c * CH * 64),
                                                    wbuf + (uint32_t)(((grp * UPG + c * CH) * 8 + 2 * lc) * 8), gr, au, aw);
            if (more && c == slot) prepare(i0n, nbuf, nwbuf, bar0 + (b ? 0u : 8u), rdy0 + (b ? 0u : 8u), nph);
        }
        // done with buffer b: the last warp to get here refills it with tile k+2
        __syncwarp();
        unsigned int last = 0;
        if (lane == 0) {
            __threadfence_block();
            last = (atomicAdd(&done_cnt[b], 1u) == NW - 1) ? 1u : 0u;
            if (last) done_cnt[b] = 0;
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last && tix + 2 * tstep < ntiles) load_tile((tix + 2 * tstep) * P, buf, bar0 + (b ? 8u : 0u));
    }
    const int T = nta * (nta + 1) / 2;
    const size_t poff = ((size_t)blockIdx.x * KS + grp) * T * 64;
    gram2_store<MAXM>((WM == 2 ? a.part_w : a.part_u) + poff, a.plan, wi, nmac, nta, lr, lc, au);
    if constexpr (WM == 1) gram2_store<MAXM>(a.part_w + poff, a.plan, wi, nmac, nta, lr, lc, aw);
    n_fail = block_sum(n_fail, red);
    if (tid == 0) a.small[blockIdx.x] = n_fail;
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
struct SplitCfg {
    int ntb;
    int xf_nw;                 // warps of the transform kernels
    int g_nw, g_maxm, g_ks;    // Gram kernel
};

static int split_min_d() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("REE_SPLIT_MIN_D");  // tuning / test hook: dimensions below run the fused sweeps
        v = e ? atoi(e) : 56;
        if (v < 32) v = 32;
    }
    return v;
}

static int xf_variant() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("REE_XF_VARIANT");
        v = e ? atoi(e) : 0;
    }
    return v;
}

static bool pick_split(int d, SplitCfg& c) {
    if (d < split_min_d()) return false;
    if (d <= 55) c = {7, 16, 8, 3, 2};
    else if (d <= 103) c = {13, 16, 12, 3, 1};
    else if (d <= 151) c = {19, 8, 8, 7, 1};  // PDE problem (d = 150): Lf 97 KB + 8 staging blocks, 152-row Gram tiles
    else return false;
    return true;
}

bool ree_split_supported(int d) {
    SplitCfg c;
    return pick_split(d, c);
}

template <int MODE, int NTB, int NW, bool EXACT, bool TRI, int PGB = 2>
static int launch_xf(const XfArgs& a, int grid, cudaStream_t st) {
    constexpr size_t smem = ((size_t)lf_doubles<NTB, TRI>() + (size_t)NW * 8 * NTB * 8 + 2 * 8 * NTB +
                             (MODE == 1 ? BM_SMEM_DOUBLES : 0)) * sizeof(double);
    static_assert(smem <= 227 * 1024, "shared memory budget");
    static int use_tma = -1;
    if (use_tma < 0) {
        const char* e = getenv("REE_XF_TMA");  // tuning / test hook: 0 = per-lane cp.async staging of the X tiles
        use_tma = e ? atoi(e) : 1;
    }
    if (use_tma) {
        auto kern = k_xform<MODE, NTB, NW, EXACT, TRI, PGB, true>;
        REE_OPT_IN_SMEM(kern, smem);
        CUtensorMap tmap;
        int rc = make_ensemble_tmap(&tmap, a.x, a.ld, a.d, 8, 8 * NTB);
        if (rc) return rc;
        kern<<<grid, NW * 32, smem, st>>>(a, tmap);
    } else {
        auto kern = k_xform<MODE, NTB, NW, EXACT, TRI, PGB, false>;
        REE_OPT_IN_SMEM(kern, smem);
        CUtensorMap tmap;
        memset(&tmap, 0, sizeof(tmap));
        kern<<<grid, NW * 32, smem, st>>>(a, tmap);
    }
    return ree_check_launch(MODE == 1 ? "k_xform<1>" : "k_xform<2>");
}

static int xf_ws() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("REE_XF_WS");  // tuning / test hook: 0 runs the unspecialised k_xform<1> for Philox noise
        v = e ? atoi(e) : 1;
    }
    return v;
}

template <int NTB, int NC, int NPC, bool EXACT, int PGB, int RC, int RP, bool TMAX>
static int launch_xf_ws(const XfArgs& a, int grid, cudaStream_t st) {
    constexpr size_t smem = ((size_t)lf_doubles<NTB, true>() + (size_t)NC * 3 * NTB * 64 + 2 * 8 * NTB +
                             BM_SMEM_DOUBLES) * sizeof(double);
    static_assert(smem <= 227 * 1024, "shared memory budget");
    static_assert(RC == 0 || (NC * RC + NC * NPC * RP) * 32 <= 65536 / ((1 + NPC) * NC * 32) / 8 * 8 * (1 + NPC) * NC * 32,
                  "register pool");
    auto kern = k_xform_ws<NTB, NC, NPC, EXACT, PGB, RC, RP, TMAX>;
    REE_OPT_IN_SMEM(kern, smem);
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    if (TMAX) {
        int rc = make_ensemble_tmap(&tmap, a.x, a.ld, a.d, 8, 8 * NTB);
        if (rc) return rc;
    }
    kern<<<grid, (1 + NPC) * NC * 32, smem, st>>>(a, tmap);
    return ree_check_launch("k_xform_ws");
}

template <int NTB, int NC>
static int launch_xf_ws_variants(const XfArgs& a, cudaStream_t st) {
    const int sms = ree_grid_ctas() / 2;  // persistent: one CTA per SM
    const int64_t want = ((a.n + 7) / 8 + NC - 1) / NC;
    const int grid = (int)(want < sms ? want : sms);
    const int v = xf_variant();
    if ((a.d + 7) / 8 == NTB) {
        // tuning hooks: 1 = four Philox chains per producer batch, 2 = two producers per consumer
        // with setmaxnreg 120 / 56 (measured slower: the extra DFMA streams take pipe slots from the consumers)
        // tuning hooks (profiles/r2_xform_variants.txt): 1 = four Box-Muller chains per producer batch, 2 = two producers
        // per consumer with setmaxnreg 120 / 56 -- both measured slower than the default
        // 3 = per-lane cp.async staging of X instead of the TMA box copy (7.35 against 7.17 ms at N = 1e7, d = 100)
        if (v == 1) return launch_xf_ws<NTB, NC, 1, true, 4, 0, 0, true>(a, grid, st);
        if (v == 2) return launch_xf_ws<NTB, NC, 2, true, 2, 120, 56, true>(a, grid, st);
        if (v == 3) return launch_xf_ws<NTB, NC, 1, true, 2, 0, 0, false>(a, grid, st);
        return launch_xf_ws<NTB, NC, 1, true, 2, 0, 0, true>(a, grid, st);
    }
    return launch_xf_ws<NTB, NC, 1, false, 2, 0, 0, true>(a, grid, st);
}

template <int MODE, int NTB, int NW, bool TRI>
static int launch_xf_exact(const XfArgs& a, int grid, cudaStream_t st) {
    if ((a.d + 7) / 8 == NTB) return launch_xf<MODE, NTB, NW, true, TRI>(a, grid, st);
    return launch_xf<MODE, NTB, NW, false, TRI>(a, grid, st);
}

template <int MODE, bool TRI>
static int launch_xf_bucket(const SplitCfg& c, const XfArgs& a, cudaStream_t st) {
    const int sms = ree_grid_ctas() / 2;  // persistent: one CTA per SM
    const int64_t nwt = (a.n + 7) / 8;
    if constexpr (MODE == 1 && TRI) {
        if (!a.z && xf_ws()) {  // Philox noise: producer / consumer warps
            if (c.ntb == 7) return launch_xf_ws_variants<7, 8>(a, st);
            if (c.ntb == 13) return launch_xf_ws_variants<13, 8>(a, st);
        }
    }
    if (c.ntb == 7) {
        const int64_t want = (nwt + 15) / 16;
        return launch_xf_exact<MODE, 7, 16, TRI>(a, (int)(want < sms ? want : sms), st);
    }
    if (c.ntb == 19) {
        if constexpr (TRI) {
            const int64_t want = (nwt + 7) / 8;
            return launch_xf_exact<MODE, 19, 8, TRI>(a, (int)(want < sms ? want : sms), st);
        } else {
            return ree_set_error(REE_E_UNSUPPORTED, "split path: a dense factor at d > 103 uses the modular kernel");
        }
    }
    if constexpr (TRI) {
        const int64_t want = (nwt + 15) / 16;
        return launch_xf_exact<MODE, 13, 16, TRI>(a, (int)(want < sms ? want : sms), st);
    } else {  // dense parity-mode factor: 86 KB of fragments, fewer warps
        const int64_t want = (nwt + 11) / 12;
        return launch_xf_exact<MODE, 13, 12, TRI>(a, (int)(want < sms ? want : sms), st);
    }
}

template <int NTB, int NW, int MAXM, int KS, bool EXACT, int WM>
static int launch_g2(const GramArgs& a, int grid, cudaStream_t st) {
    constexpr int P = NW * 8, S = P + 8, ROWS = 8 * NTB;
    constexpr size_t smem = ((size_t)2 * ROWS * S + 2 * P + 8 * NTB) * sizeof(double);
    static_assert(smem <= 227 * 1024, "shared memory budget");
    CUtensorMap tmap;
    int rc = make_ensemble_tmap(&tmap, a.x, a.ld, a.d, S, ROWS);
    if (rc) return rc;
    auto kern = k_gram2<NTB, NW, MAXM, KS, EXACT, WM>;
    REE_OPT_IN_SMEM(kern, smem);
    kern<<<grid, NW * 32, smem, st>>>(a, tmap);
    return ree_check_launch("k_gram2");
}

template <int NTB, int NW, int MAXM, int KS, int WM>
static int launch_g2_variants(const GramArgs& a, int grid, cudaStream_t st) {
    if ((a.d + 7) / 8 == NTB) return launch_g2<NTB, NW, MAXM, KS, true, WM>(a, grid, st);
    return launch_g2<NTB, NW, MAXM, KS, false, WM>(a, grid, st);
}

// ---- entry points used by ree_step.cu -----------------------------------------------------------
int ree_split_step(const double* x, double* x_new, int64_t n, int d, int64_t ld, double t, const double* m_noise,
                   const double* F, int tri, const double* z, uint64_t seed, uint64_t draw, int64_t first, int lsf_mode,
                   const double* coef, double offset, double* g_new, double* e_new, cudaStream_t st) {
    SplitCfg c;
    if (!pick_split(d, c)) return ree_set_error(REE_E_UNSUPPORTED, "split path: unsupported dimension");
    XfArgs a{};
    a.x = x; a.n = n; a.ld = ld; a.d = d; a.A = F; a.x_new = x_new; a.t = t; a.m_noise = m_noise; a.z = z;
    a.seed = seed; a.draw = draw; a.first = first; a.lsf_mode = lsf_mode; a.coef = coef; a.offset = offset;
    a.g_new = g_new; a.e_new = e_new;
    return tri ? launch_xf_bucket<1, true>(c, a, st) : launch_xf_bucket<1, false>(c, a, st);
}

int ree_split_maha(const double* x, int64_t n, int d, int64_t ld, const double* g, const double* e, const double* mean,
                   const double* Linv, const double* logdet, double* logq, double* ws_small, double* is4,
                   cudaStream_t st) {
    SplitCfg c;
    if (!pick_split(d, c)) return ree_set_error(REE_E_UNSUPPORTED, "split path: unsupported dimension");
    XfArgs a{};
    a.x = x; a.n = n; a.ld = ld; a.d = d; a.A = Linv; a.g = g; a.e = e; a.mean = mean; a.logdet = logdet;
    a.logq = logq; a.small = ws_small;
    const int sms = ree_grid_ctas() / 2;
    const int64_t want = ((n + 7) / 8 + c.xf_nw - 1) / c.xf_nw;
    const int grid = (int)(want < sms ? want : sms);
    int rc = launch_xf_bucket<2, true>(c, a, st);
    if (rc) return rc;
    return ree_mma_merge_small(ws_small, grid, is4, st);
}

// sums_u = [sum y | sum y y^T | #{g <= 0} | n], sums_w = [sum w y | sum w y y^T | sum w | 0]   (y = x - shift)
int ree_split_moments(const double* x, int64_t n, int d, int64_t ld, const double* g, const double* a_vec,
                      const double* wmax, const double* shift, double* ws_big, double* ws_small, double* sums_u,
                      double* sums_w, cudaStream_t st) {
    SplitCfg c;
    if (!pick_split(d, c)) return ree_set_error(REE_E_UNSUPPORTED, "split path: unsupported dimension");
    const bool weighted = sums_w != nullptr;
    const int nta = (d + 8) / 8, T = nta * (nta + 1) / 2, P = c.g_nw * 8;
    const int sms = ree_grid_ctas() / 2;
    const int64_t ntiles = (n + P - 1) / P;
    const int grid = (int)(ntiles < sms ? ntiles : sms);
    GramArgs a{};
    a.x = x; a.n = n; a.ld = ld; a.d = d; a.shift = shift; a.g = g; a.a = a_vec; a.wmax = wmax;
    a.part_u = ws_big;
    a.part_w = ws_big + (size_t)grid * c.g_ks * T * 64;
    a.small = ws_small;
    make_plan(a.plan, nta, c.g_nw / c.g_ks, c.g_maxm);
    int rc;
    if (c.ntb == 19) {  // no registers for two accumulator sets: unweighted pass first
        rc = launch_g2_variants<19, 8, 7, 1, 0>(a, grid, st);
    } else if (c.ntb == 7) {
        rc = weighted ? launch_g2_variants<7, 8, 3, 2, 1>(a, grid, st) : launch_g2_variants<7, 8, 3, 2, 0>(a, grid, st);
    } else {
        rc = weighted ? launch_g2_variants<13, 12, 3, 1, 1>(a, grid, st) : launch_g2_variants<13, 12, 3, 1, 0>(a, grid, st);
    }
    if (rc) return rc;
    rc = ree_mma_unpack(a.part_u, grid * c.g_ks, d, 0, sums_u, st);
    if (rc) return rc;
    rc = ree_mma_sum_small(ws_small, grid, sums_u + d + (size_t)d * d, st);
    if (rc || !weighted) return rc;
    if (c.ntb == 19) {  // ... then the pass over the sqrt(w)-scaled tile
        rc = launch_g2_variants<19, 8, 7, 1, 2>(a, grid, st);
        if (rc) return rc;
    }
    return ree_mma_unpack(a.part_w, grid * c.g_ks, d, 1, sums_w, st);
}
