// sigma / beta adaptation of one CBREE iteration in ONE persistent kernel (solver.py:505-612).
//
// The reference runs scipy's bounded Brent search over sigma (objective: (CVAR of the incremental weights - cvar_tgt)^2,
// solver.py:538-556) and then MINPACK's hybrd over beta (objective: ESS(beta) - ess_tgt N, solver.py:591-605).  Every
// objective evaluation is a reduction over all particles; driven from the host that was ~25 evaluations per iteration,
// each one a launch + a cross-rank merge kernel + a readback + ~25 us of interpreter time -- 86 launches per iteration
// and a host-paced 15-30 % of the step (round-1 profile).  Here one cooperative kernel (one launch, one readback) does
//   A. the Brent search: every CTA sweeps its slice of g for the trial sigma, the last CTA to arrive merges the partials
//      in fixed order, exchanges the 4-tuple with the other ranks through the NVLink mailboxes (peer.cuh) and publishes
//      the global tuple; thread 0 of EVERY CTA then advances the same scalar state machine (scalar_opt.cuh: SciPy's
//      arithmetic, deterministic => all CTAs and all ranks agree on the next trial point without a broadcast);
//   B. ell = log target(g, e; sigma_new) into a scratch vector + its range (the softmax shift of every beta follows);
//   C. the hybrd root find over beta on ell;
// and writes sigma, beta, ESS and the search diagnostics.  g / ell are re-read ~25 times: 80 MB at N = 1e7, i.e. largely
// from L2 (126 MB).  This translation unit is compiled with -fmad=false: the scalar searches must not be contracted.
#include <stdio.h>
#include <stdlib.h>

#include "nvec_stream.cuh"
#include "peer.cuh"
#include "scalar_opt.cuh"

#define REE_WS_HEADER 8
#define OPT_THREADS 1024

struct AdaptArgs {
    const double* g;
    const double* e;
    int64_t n;
    double* ell;
    double* ws;
    double* out;
    ree_adapt p;
    PeerArgs peer;  // world == 1: a mailbox of this device alone (the publication path is the same)
};

enum { MERGE_PLAIN = 0, MERGE_EXPS = 1, MERGE_RANGE = 2 };

struct OptShared {
    ExpSums es[32];
    double red[3 * 32];
    double tab[64];
    double xch[REE_PEER_MAX][4];
    double x, y;
    int more;
    int last;
};

__device__ __forceinline__ unsigned int atom_add_acq_rel_u32(unsigned int* p, unsigned int v) {
    unsigned int old;
    asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_f64(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_f64(double* p, double v) {
    asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

__device__ __forceinline__ double block_max(double v, double* smem) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double q = __shfl_xor_sync(0xffffffffu, v, o);
        v = q > v ? q : v;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    double m = -CUDART_INF;
    for (int w = 0; w < nw; ++w) m = smem[w] > m ? smem[w] : m;
    __syncthreads();
    return m;
}

// three plain sums in one pass (two CTA barriers instead of six); results valid in thread 0; fixed tree
__device__ __forceinline__ void block_sum3(double& a, double& b, double& c, double* smem /* >= 96 */) {
    a = warp_sum(a);
    b = warp_sum(b);
    c = warp_sum(c);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if (lane == 0) {
        smem[warp] = a;
        smem[32 + warp] = b;
        smem[64 + warp] = c;
    }
    __syncthreads();
    if (warp == 0) {
        a = warp_sum(lane < nw ? smem[lane] : 0.0);
        b = warp_sum(lane < nw ? smem[32 + lane] : 0.0);
        c = warp_sum(lane < nw ? smem[64 + lane] : 0.0);
    }
    __syncthreads();
}

// Grid-wide (and cross-rank) combination of per-CTA partial tuples -- the epilogue of every objective reduction.
//   1. thread 0 of every CTA stores its partial and arrives (one acq_rel atomic, no fences);
//   2. the last CTA to arrive merges the partials in fixed order and PUBLISHES the rank's tuple with a sequence number
//      into slot [seq & 1][rank] of every rank's mailbox (NVLink stores; with one rank: this device's own mailbox);
//   3. warp 0 of EVERY CTA polls its own rank's mailbox until all ranks' slots carry `seq` and merges them in rank order.
// One hop from the last arrival to every CTA of every rank; identical bits everywhere (fixed orders), so every CTA
// advances the scalar search identically.  `k` counts the combinations of this launch (identical in all CTAs); a rank
// publishes k + 1 only after all its CTAs have read combination k, which keeps the two slot parities safe (ree_peer.cu).
//   ws header (unsigned words): [0] ticket of the one-shot reduction kernels, [2] arrivals, [6] exits
__device__ __forceinline__ ExpSums grid_combine(const AdaptArgs& A, const ExpSums& blk, int mode, int& k,
                                                unsigned long long seq_base, OptShared& sh) {
    const unsigned int G = gridDim.x;
    double* part = A.ws + REE_WS_HEADER + (size_t)(k & 1) * 4 * G;
    unsigned int* ctl = reinterpret_cast<unsigned int*>(A.ws);
    const unsigned long long seq = seq_base + (unsigned long long)k + 1ull;
    const int par = (int)(seq & 1ull);
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        double* p = part + 4 * (size_t)blockIdx.x;
        reinterpret_cast<double2*>(p)[0] = make_double2(blk.m, blk.s1);
        reinterpret_cast<double2*>(p)[1] = make_double2(blk.s2, blk.cnt);
        const unsigned int old = atom_add_acq_rel_u32(ctl + 2, 1u);
        sh.last = (old == (unsigned int)(k + 1) * G - 1u) ? 1 : 0;
    }
    __syncthreads();
    if (sh.last) {
        ExpSums v;
        if (mode == MERGE_EXPS) {
            v = expsums_empty();
            for (unsigned int b = threadIdx.x; b < G; b += blockDim.x) {
                const double* p = part + 4 * (size_t)b;
                v = expsums_merge(v, ExpSums{ld_relaxed_f64(p), ld_relaxed_f64(p + 1), ld_relaxed_f64(p + 2),
                                             ld_relaxed_f64(p + 3)});
            }
            v = expsums_block_reduce(v, sh.es);
        } else if (mode == MERGE_PLAIN) {  // all non-empty partials carry the same shift: plain sums
            double m = -CUDART_INF, s1 = 0.0, s2 = 0.0, c = 0.0;
            for (unsigned int b = threadIdx.x; b < G; b += blockDim.x) {
                const double* p = part + 4 * (size_t)b;
                const double pm = ld_relaxed_f64(p);
                m = pm > m ? pm : m;
                s1 += ld_relaxed_f64(p + 1);
                s2 += ld_relaxed_f64(p + 2);
                c += ld_relaxed_f64(p + 3);
            }
            v.m = block_max(m, sh.red);
            block_sum3(s1, s2, c, sh.red);
            v.s1 = s1; v.s2 = s2; v.cnt = c;
        } else {  // [max, min]
            double hi = -CUDART_INF, nlo = -CUDART_INF;
            for (unsigned int b = threadIdx.x; b < G; b += blockDim.x) {
                const double* p = part + 4 * (size_t)b;
                const double a = ld_relaxed_f64(p), c = -ld_relaxed_f64(p + 1);
                hi = a > hi ? a : hi;
                nlo = c > nlo ? c : nlo;
            }
            v.m = block_max(hi, sh.red);
            v.s1 = -block_max(nlo, sh.red);
            v.s2 = 0.0;
            v.cnt = 0.0;
        }
        if (threadIdx.x < 32) {  // publish: lane r stores into rank r's mailbox
            const double v0 = __shfl_sync(0xffffffffu, v.m, 0), v1 = __shfl_sync(0xffffffffu, v.s1, 0);
            const double v2 = __shfl_sync(0xffffffffu, v.s2, 0), v3 = __shfl_sync(0xffffffffu, v.cnt, 0);
            if (lane < A.peer.world) {
                PeerSlot* dst = &A.peer.box[lane]->slot[par][A.peer.rank];
                st_relaxed_f64(&dst->v[0], v0);
                st_relaxed_f64(&dst->v[1], v1);
                st_relaxed_f64(&dst->v[2], v2);
                st_relaxed_f64(&dst->v[3], v3);
                st_release_sys_u64(&dst->seq, seq);
            }
        }
    }
    ExpSums out = expsums_empty();
    if (threadIdx.x < 32) {
        if (lane < A.peer.world) {
            const PeerSlot* src = &A.peer.box[A.peer.rank]->slot[par][lane];
            unsigned long long spins = 0;
            while (ld_acquire_sys_u64(&src->seq) != seq) {
                if (A.peer.world > 1 && ++spins > REE_PEER_SPIN_LIMIT) {  // a peer died or the call order diverged
                    *A.peer.error = 1;
                    break;
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) sh.xch[lane][q] = ld_relaxed_f64(&src->v[q]);
        }
        __syncwarp();
        if (lane == 0) {
            if (mode == MERGE_RANGE) {  // the range of ell stays rank-local (the exchange only keeps the ranks in step)
                const int r = A.peer.rank;
                out = ExpSums{sh.xch[r][0], sh.xch[r][1], 0.0, 0.0};
            } else {
                out = ExpSums{sh.xch[0][0], sh.xch[0][1], sh.xch[0][2], sh.xch[0][3]};
                for (int q = 1; q < A.peer.world; ++q)
                    out = expsums_merge(out, ExpSums{sh.xch[q][0], sh.xch[q][1], sh.xch[q][2], sh.xch[q][3]});
            }
        }
        __syncwarp();
    }
    k += 1;
    return out;
}

// ---------------------------------------------------------------------------------------------------------------------
// objectives
// ---------------------------------------------------------------------------------------------------------------------
// 1/sqrt(s) and 1/a to ~2^-58 from the MUFU seeds (RSQ64H / RCP64H, ~2^-21) and one cubic correction: 6 / 4 fp64
// instructions instead of the ~25 each of IEEE sqrt + divide.  The objective sums tolerate it: every term is accurate to
// an ulp of the LARGEST term (the smoothed indicator is <= 2), which is what its rounded form in the reference is too.
__device__ __forceinline__ double rsqrt_fast(double s) {
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(s));
    const double t = s * y0;
    const double e = fma(-t, y0, 1.0);
    return fma(y0 * e, fma(e, 0.375, 0.5), y0);
}
__device__ __forceinline__ double rcp_fast(double a) {
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(a));
    const double e = fma(-a, r0, 1.0);
    return fma(r0, fma(e, e, e), r0);
}

// algebraic target, sigma_new >= sigma_old >= 0: ratio of the smoothed indicators (see k_reduce_cvar_alg, ree_nvec.cu)
__device__ __forceinline__ ExpSums cvar_alg_partial(const AdaptArgs& A, double sigma_new, double sigma_old, OptShared& sh) {
    double s1 = 0.0, s2 = 0.0;
    unsigned int cnt = 0;
    const double sn2 = sigma_new * sigma_new, so2 = sigma_old * sigma_old;
    auto one = [&](double gi) {
        const double g2 = gi * gi;
        const double sn = fma(sn2, g2, 1.0), so = fma(so2, g2, 1.0);
        double in, io;
        if (sn < 1.0e15) {  // (so <= sn)
            in = fma(-sigma_new * gi, rsqrt_fast(sn), 1.0);
            io = fma(-sigma_old * gi, rsqrt_fast(so), 1.0);
        } else {  // |sigma g| > 3e7 (or g not finite): the reference's own rounding decides whether 1 + q is 0
            const double qn = __ddiv_rn(__dmul_rn(-sigma_new, gi), __dsqrt_rn(__dadd_rn(1.0, __dmul_rn(sn2, g2))));
            const double qo = __ddiv_rn(__dmul_rn(-sigma_old, gi), __dsqrt_rn(__dadd_rn(1.0, __dmul_rn(so2, g2))));
            in = __dadd_rn(1.0, qn);
            io = __dadd_rn(1.0, qo);
        }
        const bool ok = in > 0.0 && io > 0.0 && in <= 2.0 && io <= 2.0;
        const double v = ok ? in * rcp_fast(ok ? io : 1.0) : 0.0;
        s1 += v;
        s2 = fma(v, v, s2);
        cnt += ok ? 1u : 0u;
    };
    stream1q(A.g, A.n,
             [&](int64_t, const Quad& q) {
#pragma unroll
                 for (int k = 0; k < 4; ++k) one(q.v[k]);
             },
             [&](int64_t, double gi) { one(gi); });
    double cn = (double)cnt;
    block_sum3(s1, s2, cn, sh.red);
    return ExpSums{cn > 0.0 ? 0.0 : -CUDART_INF, s1, s2, cn};
}

// any target: a_i = logtgt(g_i, 0; sigma_new) - logtgt(g_i, 0; sigma_old), shifted exponential sums (k_reduce_cvar)
__device__ __forceinline__ ExpSums cvar_generic_partial(const AdaptArgs& A, double sigma_new, double sigma_old, uint32_t T,
                                                        OptShared& sh) {
    ExpAcc acc = expacc_empty();
    const int kind = A.p.tgt_kind;
    auto val = [&](double gi) {
        return __dsub_rn(ree_logtgt(gi, 0.0, sigma_new, kind), ree_logtgt(gi, 0.0, sigma_old, kind));
    };
    stream1q(A.g, A.n,
             [&](int64_t, const Quad& gq) {
                 double a[4];
#pragma unroll
                 for (int k = 0; k < 4; ++k) a[k] = val(gq.v[k]);
                 expacc_add4(acc, a, T);
             },
             [&](int64_t, double gi) { expacc_add1(acc, val(gi), T); });
    return expacc_block_reduce(acc, T, sh.red);
}

// [M, sum exp(beta ell - M), sum exp(2 (beta ell - M)), #finite] with the shift M = beta * (max or min of ell on this
// rank) known in advance (k_reduce_scaled_known)
__device__ __forceinline__ ExpSums ess_partial(const AdaptArgs& A, double beta, double hi, double lo, uint32_t T,
                                               OptShared& sh) {
    const double M = __dmul_rn(beta >= 0.0 ? hi : lo, beta);
    double s1 = 0.0, s2 = 0.0;
    unsigned int cnt = 0;
    auto one = [&](double li) {
        const double a = __dmul_rn(li, beta);
        const bool ok = isfinite(a);
        const double w = exp_nonpos(ok ? a - M : -CUDART_INF, T);
        s1 += w;
        s2 = fma(w, w, s2);
        cnt += ok ? 1u : 0u;
    };
    stream1q(A.ell, A.n,
             [&](int64_t, const Quad& lq) {
#pragma unroll
                 for (int k = 0; k < 4; ++k) one(lq.v[k]);
             },
             [&](int64_t, double li) { one(li); });
    double cn = (double)cnt;
    block_sum3(s1, s2, cn, sh.red);
    return ExpSums{cn > 0.0 ? M : -CUDART_INF, s1, s2, cn};
}

// ell = logtgt(g, e; sigma) (same arithmetic as k_logtgt) and this CTA's [max, min] over the finite entries
__device__ __forceinline__ ExpSums ell_partial(const AdaptArgs& A, double sigma, OptShared& sh) {
    double hi = -CUDART_INF, nlo = -CUDART_INF;
    const int kind = A.p.tgt_kind;
    double* __restrict__ ell = A.ell;
    const bool vec = aligned16(ell);
    auto track = [&](double v) {
        if (isfinite(v)) {
            hi = v > hi ? v : hi;
            nlo = -v > nlo ? -v : nlo;
        }
    };
    auto quad = [&](int64_t i0, const Quad& gq, const Quad& eq) {
        double a[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            a[k] = ree_logtgt(gq.v[k], eq.v[k], sigma, kind);
            track(a[k]);
        }
        if (vec) {
            reinterpret_cast<double2*>(ell + i0)[0] = make_double2(a[0], a[1]);
            reinterpret_cast<double2*>(ell + i0)[1] = make_double2(a[2], a[3]);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) ell[i0 + k] = a[k];
        }
    };
    auto one = [&](int64_t i, double gi, double ei) {
        const double a = ree_logtgt(gi, ei, sigma, kind);
        ell[i] = a;
        track(a);
    };
    if (A.e) stream2q(A.g, A.e, A.n, quad, one);
    else stream1q(A.g, A.n, [&](int64_t i0, const Quad& gq) { quad(i0, gq, Quad{}); }, [&](int64_t i, double gi) { one(i, gi, 0.0); });
    hi = block_max(hi, sh.red);
    nlo = block_max(nlo, sh.red);
    return ExpSums{hi, -nlo, 0.0, 0.0};
}

// my_log_cvar's conventions on the shifted sums (utilities.py:76-102, engine.cvar_from)
__device__ __forceinline__ double cvar_of(const ExpSums& t) {
    const double mean = t.s1 / t.cnt;
    if (mean == 0.0) return CUDART_INF;
    const double var = t.s2 / t.cnt - mean * mean;
    const double out = sqrt(var > 0.0 ? var : 0.0) / mean;
    return out != out ? CUDART_INF : out;
}

// out[16]: 0 sigma, 1 beta, 2 ESS, 3 sigma status (0 found, 1 maxiter, 2 NaN, -1 empty reduction, -2 not searched),
//          4 beta status (hybrd info 1..5, 0 improper input, -1 empty reduction, -2 not searched), 5 sigma evaluations,
//          6 beta evaluations, 7 grid combinations, 8..11 the ESS tuple [M, S1, S2, cnt] at (sigma, beta),
//          12 hybrd's root before the positivity test, 13 objective there, 14 / 15 range of ell
template <bool ALG>
__global__ void __launch_bounds__(OPT_THREADS, 1) k_adapt(const AdaptArgs A) {
    __shared__ OptShared sh;
    __shared__ ree_opt::BrentBounded br;  // the searches' state: touched by thread 0 only, kept out of the registers
    __shared__ ree_opt::Hybr1 hy;
    const uint32_t T = load_exp_table(sh.tab);
    int k = 0;
    unsigned long long seq_base = 0;  // combinations this mailbox has seen before this launch (warp 0 needs it)
    if (threadIdx.x < 32) {
        if (threadIdx.x == 0) seq_base = A.peer.box[A.peer.rank]->seq;
        seq_base = __shfl_sync(0xffffffffu, seq_base, 0);
    }
    const bool lead = threadIdx.x == 0;

    // ---- A. sigma ------------------------------------------------------------------------------------------------
    double sigma = A.p.sigma;
    int sig_status = -2, sig_nfev = 0;
    if (A.p.do_sigma) {
        double x = A.p.sigma;
        if (lead) sh.x = br.start(A.p.sigma, A.p.sigma_hi, A.p.xatol, A.p.maxiter);
        __syncthreads();
        x = sh.x;
        __syncthreads();
        for (;;) {
            const ExpSums blk = ALG ? cvar_alg_partial(A, x, A.p.sigma, sh) : cvar_generic_partial(A, x, A.p.sigma, T, sh);
            const ExpSums t = grid_combine(A, blk, ALG ? MERGE_PLAIN : MERGE_EXPS, k, seq_base, sh);
            if (lead) {
                if (!(t.cnt > 0.0)) {
                    sh.more = -1;
                } else {
                    const double c = cvar_of(t) - A.p.cvar_tgt;
                    sh.more = br.tell(c * c) ? 1 : 0;
                    sh.x = br.x;
                }
            }
            __syncthreads();  // (the next write to sh.more / sh.x comes after the barriers of the next reduction)
            const int more = sh.more;
            x = sh.x;
            if (more <= 0) {
                __syncthreads();
                if (lead) {
                    sig_status = more < 0 ? -1 : br.flag;
                    sig_nfev = br.num;
                    sh.x = more < 0 ? A.p.sigma : br.result();
                }
                break;
            }
        }
        __syncthreads();
        sigma = sh.x;
        __syncthreads();
    }

    // ---- B. log-target vector for the new sigma and its range ------------------------------------------------------
    double hi, lo;
    {
        const ExpSums blk = ell_partial(A, sigma, sh);
        const ExpSums t = grid_combine(A, blk, MERGE_RANGE, k, seq_base, sh);
        if (lead) {
            sh.x = t.m;
            sh.y = t.s1;
        }
        __syncthreads();
        hi = sh.x;
        lo = sh.y;
        __syncthreads();
    }

    // ---- C. beta -----------------------------------------------------------------------------------------------------
    double beta = A.p.beta, root = A.p.beta, froot = 0.0;
    int beta_status = -2, beta_nfev = 0;
    __shared__ ExpSums t_acc, t_first;  // tuples at the temperature that is finally kept / at the start (thread 0)
    if (A.p.do_beta) {
        if (lead) {
            sh.x = hy.start(A.p.beta, A.p.xtol, A.p.maxfev, A.p.factor, A.p.epsfcn);
            sh.more = hy.done() ? 0 : 1;
        }
        __syncthreads();
        double x = sh.x;
        bool more = sh.more != 0;
        __syncthreads();
        int nev = 0;
        while (more) {
            const ExpSums blk = ess_partial(A, x, hi, lo, T, sh);
            const ExpSums t = grid_combine(A, blk, MERGE_PLAIN, k, seq_base, sh);
            if (lead) {
                if (nev == 0) t_first = t_acc = t;
                if (!(t.cnt > 0.0)) {
                    sh.more = -1;
                } else {
                    const int phase = hy.phase;
                    const double ess = t.s1 * t.s1 / t.s2;
                    const bool m = hy.tell(ess - A.p.ess_tgt_n);
                    if (phase == ree_opt::Hybr1::P_STEP && hy.x == x) t_acc = t;  // the trial point was accepted
                    sh.more = m ? 1 : 0;
                    sh.x = hy.x_eval;
                }
                nev += 1;
            }
            __syncthreads();
            const int mm = sh.more;
            x = sh.x;
            if (mm <= 0) {
                __syncthreads();
                if (lead) {
                    beta_status = mm < 0 ? -1 : hy.info;
                    beta_nfev = nev;
                    root = hy.result();
                    froot = hy.fresult();
                    if (mm < 0 || !(root > 0.0)) {  // solver.py:602-605: a non-positive root is rejected
                        beta = A.p.beta;
                        t_acc = t_first;
                    } else {
                        beta = root;
                    }
                }
                more = false;
            }
        }
    } else {
        const ExpSums blk = ess_partial(A, beta, hi, lo, T, sh);
        const ExpSums t = grid_combine(A, blk, MERGE_PLAIN, k, seq_base, sh);
        if (lead) t_acc = t;
    }

    if (lead && blockIdx.x == 0) {
        double* o = A.out;
        o[0] = sigma;
        o[1] = beta;
        o[2] = t_acc.s1 * t_acc.s1 / t_acc.s2;
        o[3] = (double)sig_status;
        o[4] = (double)beta_status;
        o[5] = (double)sig_nfev;
        o[6] = (double)beta_nfev;
        o[7] = (double)k;
        o[8] = t_acc.m; o[9] = t_acc.s1; o[10] = t_acc.s2; o[11] = t_acc.cnt;
        o[12] = root;
        o[13] = froot;
        o[14] = hi;
        o[15] = lo;
    }
    if (lead) {  // the last CTA to leave rearms the control words for the next launch
        unsigned int* ctl = reinterpret_cast<unsigned int*>(A.ws);
        __threadfence();
        const unsigned int old = atomicAdd(ctl + 6, 1u);
        if (old == gridDim.x - 1) {
            ctl[2] = 0u;
            ctl[6] = 0u;
            A.peer.box[A.peer.rank]->seq = seq_base + (unsigned long long)k;
            __threadfence();
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
static int g_opt_ctas[64][2] = {{0}};
static PeerArgs g_local_box[64];  // single-rank publication mailbox of each device (world = 1)
static bool g_local_ready[64] = {false};

static int opt_max_ctas(int alg) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
    if (g_opt_ctas[dev][alg] == 0) {
        int sms = 0, per = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaError_t err = alg ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_adapt<true>, OPT_THREADS, 0)
                              : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_adapt<false>, OPT_THREADS, 0);
        if (err != cudaSuccess || per < 1 || sms < 1) return 0;
        int ctas = sms;  // one CTA per SM: the fewest arrivals per combination at full thread occupancy
        if (const char* env = getenv("REE_ADAPT_CTAS")) {  // experiments (tools/adapt_bench.py)
            const int v = atoi(env);
            if (v >= 1 && v <= sms * per) ctas = v;
        }
        const int cap = 2 * ree_grid_ctas();  // the small-partials region of the workspace holds 16 G doubles
        g_opt_ctas[dev][alg] = ctas > cap ? cap : ctas;
    }
    return g_opt_ctas[dev][alg];
}

static const PeerArgs* local_mailbox() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (!g_local_ready[dev]) {
        PeerArgs a{};
        cudaError_t err = cudaMalloc((void**)&a.box[0], sizeof(PeerBox));  // (no staging areas: single rank)
        if (err == cudaSuccess) err = cudaMemset(a.box[0], 0, sizeof(PeerBox));
        if (err == cudaSuccess) err = cudaMalloc((void**)&a.error, sizeof(int));
        if (err == cudaSuccess) err = cudaMemset(a.error, 0, sizeof(int));
        if (err == cudaSuccess) err = cudaDeviceSynchronize();
        if (err != cudaSuccess) {
            ree_set_error((int)err, cudaGetErrorString(err));
            return nullptr;
        }
        a.rank = 0;
        a.world = 1;
        g_local_box[dev] = a;
        g_local_ready[dev] = true;
    }
    return &g_local_box[dev];
}

extern "C" int ree_cbree_adapt(const double* g, const double* e, int64_t n, const ree_adapt* p, double* ell,
                               double* workspace, double* out16, void* stream) {
    if (!g || !p || !ell || !workspace || !out16 || n < 0)
        return ree_set_error(REE_E_BADARG, "ree_cbree_adapt: bad argument");
    if (p->do_sigma && !(p->sigma_hi >= p->sigma))
        return ree_set_error(REE_E_BADARG, "ree_cbree_adapt: the lower bound exceeds the upper bound");
    const int alg = (p->tgt_kind == REE_TGT_ALGEBRAIC && p->sigma >= 0.0) ? 1 : 0;
    const int cap = opt_max_ctas(alg);
    if (cap < 1) return ree_set_error(REE_E_UNSUPPORTED, "ree_cbree_adapt: the persistent kernel does not fit this device");
    int64_t want = (n + 4 * OPT_THREADS - 1) / (4 * OPT_THREADS);
    const int ctas = (int)(want < 1 ? 1 : (want > cap ? cap : want));
    AdaptArgs a{};
    a.g = g; a.e = e; a.n = n; a.ell = ell; a.ws = workspace; a.out = out16; a.p = *p;
    const PeerArgs* peer = p->use_peers ? ree_peer_args_current() : local_mailbox();
    if (!peer)
        return ree_set_error(REE_E_BADARG, p->use_peers ? "ree_cbree_adapt: use_peers without connected peer mailboxes"
                                                         : "ree_cbree_adapt: no mailbox");
    a.peer = *peer;
    void* args[] = {&a};
    cudaError_t err = alg ? cudaLaunchCooperativeKernel((void*)k_adapt<true>, dim3(ctas), dim3(OPT_THREADS), args, 0,
                                                        ree_stream(stream))
                          : cudaLaunchCooperativeKernel((void*)k_adapt<false>, dim3(ctas), dim3(OPT_THREADS), args, 0,
                                                        ree_stream(stream));
    if (err != cudaSuccess) {
        cudaGetLastError();
        return ree_set_error((int)err, cudaGetErrorString(err));
    }
    return ree_check_launch("ree_cbree_adapt");
}

// ---------------------------------------------------------------------------------------------------------------------
// Host builds of the scalar searches (the same code the kernel runs): tests/test_scalar_opt.py checks them against
// scipy.optimize.minimize_scalar(method="Bounded") and scipy.optimize.root(method="hybr") bit for bit.
// ---------------------------------------------------------------------------------------------------------------------
extern "C" int ree_scalar_brent(ree_scalar_fn f, void* user, double lo, double hi, double xatol, int maxiter, double* x_out,
                                double* f_out, int* nfev, int* flag) {
    if (!f || !x_out) return ree_set_error(REE_E_BADARG, "ree_scalar_brent: bad argument");
    ree_opt::BrentBounded br;
    double x = br.start(lo, hi, xatol, maxiter);
    while (br.tell(f(x, user))) x = br.x;
    *x_out = br.result();
    if (f_out) *f_out = br.fx;
    if (nfev) *nfev = br.num;
    if (flag) *flag = br.flag;
    return 0;
}

extern "C" int ree_scalar_hybr1(ree_scalar_fn f, void* user, double x0, double xtol, int maxfev, double factor,
                                double epsfcn, double* x_out, double* f_out, int* nfev, int* info) {
    if (!f || !x_out) return ree_set_error(REE_E_BADARG, "ree_scalar_hybr1: bad argument");
    ree_opt::Hybr1 hy;
    double x = hy.start(x0, xtol, maxfev, factor, epsfcn);
    while (!hy.done() && hy.tell(f(x, user))) x = hy.point();
    *x_out = hy.result();
    if (f_out) *f_out = hy.fresult();
    if (nfev) *nfev = hy.nfev;
    if (info) *info = hy.info;
    return 0;
}
