/*
 * ree_b200.h -- C ABI of the B200-native CBREE engine (libree_b200.so).
 *
 * The reference (AlthausKonstantin/RareEventEstimation) is pure Python/NumPy and has
 * no FFI layer; every entry point below therefore cites the
 * reference NumPy/SciPy call site it replaces (paths relative to
 * src/rareeventestimation/).  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - plain pointers and sizes only; all `double*` / `const double*` arguments are
 *     DEVICE pointers unless the name starts with `host_`.
 *   - ensembles are SoA: element (particle i, dimension j) lives at X[j*ld + i],
 *     ld >= n, ld a multiple of 8 (64-byte rows).
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Calls are
 *     asynchronous on that stream; nothing synchronises unless stated.
 *   - return value: 0 = ok, >0 = cudaError_t, <0 = argument error (REE_E_*).
 *     ree_last_error() returns a thread-local message for the last failure.
 *   - results of reductions are deterministic (fixed-order two-stage reductions, no
 *     floating-point atomics): same inputs + same launch => same bits.
 */
#ifndef REE_B200_H
#define REE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define REE_ABI_VERSION 9

/* argument errors */
#define REE_E_BADARG (-1)
#define REE_E_UNSUPPORTED (-2)
#define REE_E_WORKSPACE (-3)

/* smoothed-indicator kinds, solver.py:321-363 */
enum ree_tgt_kind {
    REE_TGT_ALGEBRAIC = 0,
    REE_TGT_TANH = 1,
    REE_TGT_ARCTAN = 2,
    REE_TGT_ERF = 3,
    REE_TGT_RELU = 4,
    REE_TGT_SIGMOID = 5
};

/* limit-state functions */
enum ree_lsf_kind {
    REE_LSF_EXTERNAL = 0,        /* evaluated by the caller (arbitrary Python callable) */
    REE_LSF_AFFINE = 1,          /* g = offset + sum_j coef[j]*x_j ; coef == NULL: g = -sum(x)/sqrt(d) + offset,
                                    the reference's linear problem, problem/toy_problems.py:35-48, README.md:31-49 */
    REE_LSF_CONVEX = 2,          /* problem/toy_problems.py:12-13 (d = 2) */
    REE_LSF_FUJITA_RACKWITZ = 3, /* g = -sum_j log Phi(-x_j) - offset ; toy_problems.py:57-58 */
    REE_LSF_OSCILLATOR = 4,      /* problem/toy_problems.py:72-80 (d = 6) */
    REE_LSF_DIFFUSION = 5        /* problem/diffusion.py:131-171, 232-242 (d = n_modes) */
};

/* Limit-state descriptor.  `coef` (device) is kind specific:
 *   AFFINE     d doubles
 *   DIFFUSION  mode table, row-major [n_quad = 3*n_ele][d]  (diffusion.py:222-229 `function_values`)
 *   others     unused (NULL)                                                                       */
typedef struct ree_lsf {
    int32_t kind;
    int32_t d;
    const double* coef;
    double offset;   /* AFFINE: beta ; FUJITA_RACKWITZ: c ; DIFFUSION: u_max */
    double p0;       /* DIFFUSION: mu  of log-field (diffusion.py:203) */
    double p1;       /* DIFFUSION: sigma of log-field (diffusion.py:202) */
    int32_t n_ele;   /* DIFFUSION: number of finite elements (512) */
    int32_t reserved;
} ree_lsf;

/* noise source of the Euler-Maruyama step */
enum ree_noise_mode {
    REE_NOISE_INJECT = 0, /* Z given as an SoA device array (parity with Generator.multivariate_normal, solver.py:667) */
    REE_NOISE_PHILOX = 1  /* Philox4x32-10 + Box-Muller, counter = (global particle, dim pair, draw) */
};

/* ---- library ----------------------------------------------------------------------- */
int ree_abi_version(void);
const char* ree_last_error(void);
/* kernels launched by this library in this process so far (bench.py reports the delta as gpu_launches) */
int64_t ree_launch_count(void);
/* Which kernels serve dimension d:
 *   1  fused DMMA sweeps (ree_cbree_step and ree_cbree_maha each produce their Gram)      -- HBM-bound d
 *   2  split DMMA path: ree_cbree_step(sums = NULL) -> ree_reduce_ess(a_out) -> ree_cbree_moments ->
 *      ree_moments_chol -> ree_cbree_maha(sums = NULL)                                    -- fp64-bound d (56..103)
 *   0  modular kernels (need the optional `w` scratch vector of ree_cbree_maha)           -- d > 103
 * Path 2 dimensions are also served by the path 1 calling sequence.                                    */
int ree_fused_path(int d);
/* Largest supported dimension (256): beyond it the d x d factorisations and the modular Gram have no kernel.  The host
 * side raises before any work (the reference itself accepts any d; toy_problems.py:95-101 goes up to d = 200). */
int ree_max_dim(void);
/* number of CTAs every reduction kernel uses on the current device (partials per launch) */
int ree_num_ctas(void);
/* bytes of scratch a call with dimension d may need (partials of the d x d sums).  The workspace
 * must be zero-filled once after allocation (it holds a self-resetting ticket counter).          */
int64_t ree_workspace_bytes(int d);
/* Synchronising readback of n <= 64 doubles of device memory, stream-ordered behind the work already queued on
 * `stream`: a one-warp kernel copies them into mapped pinned host memory, the caller spins on a sequence number
 * (a few microseconds instead of a cudaMemcpy round trip).  Replaces the implicit host arrays of the reference's
 * scalar objectives (solver.py:538-556, 591-601: every objective evaluation returns a Python float).              */
int ree_fetch_small(const double* dev_src, int n, double* host_out, void* stream);
/* Synchronising copy of an n-vector of device memory into pageable host memory (double-buffered through pinned
 * staging): what reading CBREECache.lsf_evals / e_fun_evals or a row of Solution.lsf_eval_hist costs
 * (solver.py:63-116, solution.py:9-50 hold these as host arrays).                                                  */
int ree_download_vector(const double* dev_src, double* host_dst, int64_t n, void* stream);

/* ---- layout (host (N,d) C-order <-> device SoA); replaces nothing, it is the boundary -- */
int ree_aos_to_soa(const double* aos, double* soa, int64_t n, int d, int64_t ld, void* stream);
int ree_soa_to_aos(const double* soa, double* aos, int64_t n, int d, int64_t ld, void* stream);
/* HOST-buffer convenience: chunked cudaMemcpyAsync + transpose through `staging`
 * (device, staging_elems doubles).  Synchronises the stream before returning.          */
int ree_upload_ensemble(const double* host_aos, double* soa, int64_t n, int d, int64_t ld,
                        double* staging, int64_t staging_elems, void* stream);
int ree_download_ensemble(const double* soa, double* host_aos, int64_t n, int d, int64_t ld,
                          double* staging, int64_t staging_elems, void* stream);

/* ---- noise / initial sample --------------------------------------------------------- */
/* Z[j*ld+i] ~ N(0,1), counter (first_particle+i, j, draw).  With draw = REE_DRAW_SAMPLE this is
 * the device replacement of NormalProblem.sample_gen (problem/problem.py:115-118).      */
int ree_philox_normal(double* z, int64_t n, int d, int64_t ld, uint64_t seed, uint64_t draw,
                      int64_t first_particle, void* stream);

/* ---- limit-state and energy evaluation ----------------------------------------------- */
/* g[i] = lsf(x_i)            replaces prob.lsf(x), solver.py:376-378 */
int ree_lsf_eval(const ree_lsf* lsf, const double* x, int64_t n, int64_t ld, double* g, void* stream);
/* e[i] = 0.5*|x_i|^2 + d/2 log(2 pi)   replaces NormalProblem.e_fun, problem/problem.py:112-113 */
int ree_energy(const double* x, int64_t n, int d, int64_t ld, double* e, void* stream);

/* ---- N-vector kernels ------------------------------------------------------------------ */
/* ell[i] = log target(g_i, e_i; sigma)    solver.py:321-363 (same operation order, no FMA contraction) */
int ree_logtgt(const double* g, const double* e, int64_t n, double sigma, int tgt_kind, double* ell,
               void* stream);
/* One pass over (g,e): a = beta*logtgt(g,e;sigma), masked to finite a.
 *   out[0] = M = max a, out[1] = sum exp(a-M), out[2] = sum exp(2(a-M)), out[3] = #finite
 * ESS = out[1]^2/out[2].   replaces my_softmax + ESS, utilities.py:105-142, solver.py:523, 591-597.
 * `e` may be NULL (treated as 0).  `a_out` (optional, n doubles) receives a_i = beta*logtgt_i, the log-weights
 * ree_cbree_moments turns into w_i = exp(a_i - M).                                                */
int ree_reduce_ess(const double* g, const double* e, int64_t n, double sigma, double beta, int tgt_kind,
                   double* a_out, double* workspace, double* out4, void* stream);
/* One pass over g: a = logtgt(g,0;sigma_new) - logtgt(g,0;sigma_old), masked to finite a; same
 * four outputs.  cvar = sqrt(out[2]/nf - (out[1]/nf)^2)/(out[1]/nf).
 * replaces the sigma objective, solver.py:538-545 + my_log_cvar utilities.py:76-102.        */
int ree_reduce_cvar(const double* g, int64_t n, double sigma_new, double sigma_old, int tgt_kind,
                    double* workspace, double* out4, void* stream);
/* Same four outputs for a = scale * ell_i with a precomputed log-target vector (ree_logtgt): the beta root find
 * (solver.py:591-601) evaluates ESS(beta) many times for one sigma.
 * range2 (optional, DEVICE [max ell, min ell] over the finite entries, from ree_vector_range): the shift
 * max(scale * ell) is then known in advance and the kernel runs without a running maximum.                      */
int ree_reduce_scaled(const double* ell, int64_t n, double scale, const double* range2, double* workspace, double* out4,
                      void* stream);
/* out2 = [max, min] over the finite entries of v (-inf, +inf if none).                                          */
int ree_vector_range(const double* v, int64_t n, double* workspace, double* out2, void* stream);
/* Merge `nparts` 4-tuples [M, S1, S2, cnt] (e.g. all_gather'ed from the ranks) into one, shifted to the global
 * max; fixed order, so every rank computes identical bits.                                                    */
int ree_merge_expsums(const double* parts, int nparts, double* out4, void* stream);
/* out[0] = #{g_i <= 0} (as double)       replaces sum(lsf_evals <= 0), solver.py:534, 570, 709-712 */
int ree_count_failed(const double* g, int64_t n, double* workspace, double* out1, void* stream);

/* ---- sigma / beta adaptation in one persistent kernel ------------------------------------------------------------- */
/* Replaces CBREE.__update_beta_and_sigma (solver.py:505-523): the bounded Brent search over sigma
 * (scipy.optimize.minimize_scalar(method="Bounded"), solver.py:538-556; objective (cvar - cvar_tgt)^2 of the
 * incremental weights, my_log_cvar utilities.py:76-102) and the MINPACK hybrd root find over beta
 * (scipy.optimize.root, solver.py:591-605; objective ESS(beta) - ess_tgt N) run inside ONE cooperative kernel: the same
 * scalar algorithms in the same arithmetic, every objective evaluation a grid-wide reduction over g / the log-target
 * vector with the cross-rank merge of the 4-tuple (NVLink mailboxes, ree_peer_*) as the epilogue of the reduction.     */
typedef struct ree_adapt {
    int32_t tgt_kind;    /* enum ree_tgt_kind */
    int32_t do_sigma;    /* 1: search sigma in [sigma, sigma_hi]; 0: keep sigma (closed-form / constant updates are the caller's) */
    int32_t do_beta;     /* 1: root find starting at beta; 0: only the ESS of (sigma, beta) */
    int32_t maxiter;     /* Brent: 500 (SciPy default) */
    int32_t maxfev;      /* hybrd: 200 (n + 1) = 400 */
    int32_t use_peers;   /* 1: merge every tuple over the connected peer mailboxes (all ranks launch the same call) */
    double sigma;        /* current sigma = lower bound of the search = sigma_old of the incremental weights */
    double sigma_hi;     /* sigma + lip_sigma log(1 / t_step), solver.py:549-552 */
    double cvar_tgt;
    double xatol;        /* Brent: 1e-5 */
    double beta;         /* starting temperature */
    double ess_tgt_n;    /* ess_tgt * N (global particle count) */
    double xtol;         /* hybrd: 1.49012e-8 */
    double factor;       /* hybrd: 100 */
    double epsfcn;       /* hybrd: machine epsilon */
} ree_adapt;
/* g, e (e may be NULL): this rank's limit-state / energy values; ell: n doubles of scratch, on return the log-target
 * vector for the new sigma; workspace as for the reductions.
 * out16 (DEVICE): [0] sigma, [1] beta (a non-positive root keeps the old beta, solver.py:602-605), [2] ESS(sigma, beta),
 *   [3] sigma status (0 found, 1 maxiter, 2 NaN, -1 a reduction had no finite entry = the exception path of
 *   my_log_cvar, -2 not searched), [4] beta status (hybrd info 1..5, -1 empty reduction, -2 not searched),
 *   [5] / [6] objective evaluations of the two searches, [7] grid-wide combinations, [8..11] the ESS tuple
 *   [M, S1, S2, cnt], [12] / [13] hybrd's x and f(x), [14] / [15] max / min of ell on this rank.                      */
int ree_cbree_adapt(const double* g, const double* e, int64_t n, const ree_adapt* p, double* ell, double* workspace,
                    double* out16, void* stream);
/* The scalar searches of ree_cbree_adapt as plain host routines over a C callback (same source, host build): SciPy's
 * `_minimize_scalar_bounded` and MINPACK hybrd for one unknown.  flag / info as SciPy reports them.                   */
typedef double (*ree_scalar_fn)(double x, void* user);
int ree_scalar_brent(ree_scalar_fn f, void* user, double lo, double hi, double xatol, int maxiter, double* x_out,
                     double* f_out, int* nfev, int* flag);
int ree_scalar_hybr1(ree_scalar_fn f, void* user, double x0, double xtol, int maxfev, double factor, double epsfcn,
                     double* x_out, double* f_out, int* nfev, int* info);

/* ---- the CBREE step (sweep 1) ------------------------------------------------------------ */
/* X'[.,i] = t*X[.,i] + m_noise + F z_i ; g' = lsf(X') (unless EXTERNAL) ; e' = energy(X') ;
 * sums[0..d)      = sum_i (x'_i - shift)          sums[d..d+d*d) = sum_i (x'_i-shift)(x'_i-shift)^T
 * sums[d+d*d]     = #{g'_i <= 0}                  sums[d+d*d+1]  = n
 * replaces the Euler step and the unweighted moments, solver.py:665-671, 677-682.
 *   F        d x d row-major device matrix (U*sqrt(S) in parity mode, Cholesky factor otherwise)
 *   z        SoA noise block (INJECT) or NULL (PHILOX: seed, draw, first_particle)
 *   shift    d doubles (device) subtracted before accumulating (conditioning of the one-pass Gram)  */
int ree_cbree_step(const double* x, double* x_new, int64_t n, int d, int64_t ld, double t,
                   const double* m_noise, const double* F, int noise_mode, const double* z,
                   uint64_t seed, uint64_t draw, int64_t first_particle, const ree_lsf* lsf,
                   const double* shift, double* g_new, double* e_new, double* workspace, double* sums,
                   void* stream);
/* (sums may be NULL on path 2 dimensions: transform only, the moments come from ree_cbree_moments.) */
/* number of doubles in `sums` of ree_cbree_step / ree_cbree_maha / ree_ensemble_sums: d + d*d + 2 */
int64_t ree_sums_len(int d);
/* The same unweighted sums for an existing ensemble (initial sample, or after a callback replaced it);
 * g may be NULL (count = 0).  replaces average/cov in solver.py:790-791, 840-841.                 */
int ree_ensemble_sums(const double* x, int64_t n, int d, int64_t ld, const double* g, const double* shift,
                      double* workspace, double* sums, void* stream);

/* Path 2: both Grams of an ensemble in one pass over X.  With y = x - shift (shift: d device doubles or NULL) and
 * w_i = exp(a_i - wmax[0]) (0 where a_i is not finite):
 *   sums_u = [sum y | sum y y^T | #{g_i <= 0} | n]          (np.average / np.cov, solver.py:670-671)
 *   sums_w = [sum w y | sum w y y^T | sum w | 0]             (weighted mean / cov, solver.py:629-636)
 * a: log-weights from ree_reduce_ess(a_out); wmax: DEVICE scalar, out4[0] of the same call; g may be NULL (count 0).
 * sums_w == NULL computes the unweighted sums only (a, wmax unused).                                  */
int ree_cbree_moments(const double* x, int64_t n, int d, int64_t ld, const double* g, const double* a,
                      const double* wmax, const double* shift, double* workspace, double* sums_u, double* sums_w,
                      void* stream);

/* ---- small dense algebra (single CTA) ---------------------------------------------------- */
/* From raw shifted sums (after any cross-rank allreduce) compute
 *   mean[d], cov[d*d] (ddof given), L = chol(cov) (lower, row-major), Linv = L^-1, and EIGHT statistics:
 *   stats[0] = log det cov, stats[1] = info (0 ok, k>0: pivot k not positive), stats[2] = min pivot,
 *   stats[3] = max diag, stats[4] = |Linv|_F^2 = trace(cov^-1) (0 without Linv), stats[5] = |cov|_inf, stats[6..7] = 0.
 *   (min pivot >= lambda_min >= 1 / stats[4] and max diag <= lambda_max <= stats[5]: the host decides SciPy's
 *    singularity test lambda_min <= 1e6 eps lambda_max from them and only needs eigenvalues inside the gap.)
 * replaces average/cov (solver.py:670-671) + the factorisation inside multivariate_normal.logpdf (672-674).
 * d <= 160: the matrix lives in shared memory; 160 < d <= ree_max_dim(): in a per-device global scratch. */
int ree_moments_chol(const double* sums, const double* shift, int d, int ddof, int weighted, double* mean,
                     double* cov, double* L, double* Linv, double* stats8, void* stream);
/* L = chol(scale * C) for the noise covariance (solver.py:666-667), stats[0..3] as above, stats[4..7] = 0. */
int ree_chol_scaled(const double* C, int d, double scale, double* L, double* stats8, void* stream);

/* ---- IS density + weighted moments (sweep 2) ---------------------------------------------- */
/* For every particle: q = |Linv (x - mean)|^2, logq = -(d log 2pi + logdet + q)/2   (solver.py:672-674)
 *   r = -e - logq ; IS statistics with an online max shift over finite r:
 *     is4[0] = R = max r, is4[1] = sum exp(r-R) 1[g<=0], is4[2] = sum (exp(r-R) 1[g<=0])^2, is4[3] = #finite r
 *   (cvar_is_weights, solver.py:683-686; importance_sampling, utilities.py:43-45)
 * and, with w_i = exp(beta*logtgt(g_i,e_i;sigma) - wmax[0]) (0 where non-finite):
 *     sums as in ree_cbree_step but weighted and shifted by `mean`; sums[d+d*d] = sum w
 *   (weighted mean/cov, solver.py:629-636).  wmax is a DEVICE scalar (out[0] of ree_reduce_ess).
 * logq and w (n doubles each, optional) receive log q_i and the unnormalised weights w_i.
 * Path 2: sums == NULL skips the weights and the weighted sums (wmax may then be NULL as well).   */
int ree_cbree_maha(const double* x, int64_t n, int d, int64_t ld, const double* g, const double* e,
                   const double* mean, const double* Linv, const double* logdet, double sigma, double beta,
                   int tgt_kind, const double* wmax, double* logq, double* w, double* workspace, double* sums,
                   double* is4, void* stream);

/* ---- multi-GPU: 4-tuple exchange over NVLink peer memory (one box, one process per GPU) ------------------------ */
/* No counterpart in the single-process reference.  Replaces all_gather + ree_merge_expsums for the 32-byte results of
 * ree_reduce_ess / _scaled / _cvar and the IS statistics: every rank stores its tuple into every rank's mailbox
 * (CUDA IPC mapping), waits for the others' and merges in rank order -- one kernel, identical bits on all ranks.
 * Setup: ree_peer_create (returns this rank's 64-byte IPC handle), exchange the handles by any transport,
 * ree_peer_connect(handles[world][64]).  All ranks must call ree_peer_combine_expsums in the same order.           */
int ree_peer_create(int rank, int world, unsigned char* handle_out64);
int ree_peer_connect(const unsigned char* handles);
/* The same for ONE process that drives several devices with one host thread each (CBREE(devices=N)): after every thread
 * has called ree_peer_create on its device (host barrier), each calls this with the device ordinal of every rank; the
 * mailboxes are reached through cudaDeviceEnablePeerAccess instead of CUDA IPC. */
int ree_peer_connect_local(const int* devices);
/* In-place sum over all ranks of a small fp64 device vector (n <= ree_peer_allreduce_max()): one kernel that stages the
 * local vector behind this rank's mailbox, uses one mailbox exchange as the barrier and adds all ranks' staging areas in
 * rank order over NVLink (identical bits on every rank).  Replaces the NCCL all_reduce of the packed d x d sums.  Stream
 * ordered; every rank must call it in the same order as the other peer calls. */
int ree_peer_allreduce_sum(double* buf, int64_t n, void* stream);
int64_t ree_peer_allreduce_max(void);
int ree_peer_ready(void);
int ree_peer_combine_expsums(double* out4, void* stream);
int ree_peer_check(void);

/* ---- resampling branch: one-component von Mises-Fisher-Nakagami model ---------------------------------------- */
/* (CBREE(resample=True, mixture_model="vMFNM"), solver.py:654-663, 828-836; VMFNMixture mixturemodel.py:153-239;
 *  EMvMFNM / vMFNM_sample / logvMFpdf / lognakagamipdf, era/EMvMFNM.py)
 * Fit: with one component the EM reduces to one M-step (EMvMFNM.py:169-199), i.e. the weighted sums
 *   sums = [S = sum w x/|x| (d) | sum w |x|^2 | sum w |x|^4 | sum w]        (w == NULL: all ones)
 * rinv (n doubles, scratch) receives w_i/|x_i|.  The caller turns the sums into mu, kappa, m, omega.            */
int ree_vmfn_fit_sums(const double* x, int64_t n, int d, int64_t ld, const double* w, double* rinv, double* workspace,
                      double* sums, void* stream);
/* Sample n points (EMvMFNM.py:274-285, 324-378): |x| = sqrt(Gamma(m, omega/m)), cosine to the mean direction by
 * Wood's rejection sampler, uniform direction in the complement, Householder reflection I - b v v^T (house(mu),
 * EMvMFNM.py:401-429; v is a DEVICE vector).  r_in / w_in / v_in all NULL: Philox variates (counter = global particle
 * index, `draw`); otherwise injected variates (parity mode): radii r_in[n], accepted cosines w_in[n] and the raw
 * normals of the complement direction v_in[(d-1)][ld] in the ensemble's SoA layout.                              */
int ree_vmfn_sample(double* x, int64_t n, int d, int64_t ld, const double* householder_v, double householder_b,
                    double kappa, double m, double omega, uint64_t seed, uint64_t draw, int64_t first_particle,
                    const double* r_in, const double* w_in, const double* v_in, void* stream);
/* logq_i = kappa mu.x_i/|x_i| + c0 - m |x_i|^2/omega + (2m - d) log|x_i|  with c0 = the vMF and Nakagami constants
 * (mixturemodel.py:196-215).  is4 != NULL additionally returns the IS statistics of r = -e - logq with the indicator
 * g <= 0, exactly as ree_cbree_maha does (solver.py:683-686, 843-846).  mu is a DEVICE vector.
 * fuse_lsf != NULL (affine limit state of dimension d): g and e are OUTPUTS -- the kernel evaluates the limit state and
 * the energy of the points it reads anyway (same arithmetic as ree_lsf_eval / ree_energy), as the redraw of the
 * resampling branch needs g, e and log q of the same fresh ensemble (solver.py:663, 676-686).                     */
int ree_vmfn_logpdf(const double* x, int64_t n, int d, int64_t ld, const double* mu, double kappa, double m, double omega,
                    double c0, double* g, double* e, double* logq, double* workspace, double* is4,
                    const ree_lsf* fuse_lsf, void* stream);

/* ---- secondary path: SIS with adaptive conditional sampling (sis/SIS_aCS.py:64-306) ------------------------ */
/* mode 0: v_i = Phi(-g_i/sigma_new) / Phi(-g_i/sigma_old)  (sigma_old <= 0: denominator 1)   SIS_aCS.py:151-183
 * mode 1: v_i = 1[g_i < 0] / Phi(-g_i/sigma_new)                                            SIS_aCS.py:282-300
 * out4 = [0, sum v, sum v^2, n]; w_out (optional) receives v.                                               */
int ree_sis_reduce(const double* g, int64_t n, double sigma_new, double sigma_old, int mode, double* w_out,
                   double* workspace, double* out4, void* stream);
/* inclusive prefix sum cdf[i] = w[0] + ... + w[i]; scratch: ree_scan_scratch_doubles(n) doubles */
int64_t ree_scan_scratch_doubles(int64_t n);
int ree_scan_inclusive(const double* w, int64_t n, double* cdf, double* scratch, void* stream);
/* multinomial choice of nchain seeds with probabilities w/sum(w): idx[k] = first i with cdf[i] > u_k cdf[n-1],
 * u_k from Philox (first + k).  replaces np.random.choice(N, nchain, True, p), SIS_aCS.py:194               */
int ree_sis_resample(const double* cdf, int64_t n, int64_t nchain, uint64_t seed, uint64_t draw, int64_t first,
                     int64_t* idx, void* stream);
/* out[j][k] = x[j][idx[k]], g_out[k] = g[idx[k]]   (seed rows, SIS_aCS.py:195-196) */
int ree_gather_columns(const double* x, int d, int64_t ld_src, const int64_t* idx, int64_t nchain, double* out,
                       int64_t ld_out, const double* g, double* g_out, void* stream);
/* cand = rho*cur + sigma_f*z, z ~ N(0,I)   (SIS_aCS.py:232) */
int ree_acs_propose(const double* cur, int64_t nchain, int d, int64_t ld, double rho, double sigma_f, uint64_t seed,
                    uint64_t draw, int64_t first, double* cand, void* stream);
/* accept with probability min(1, Phi(-g_cand/sigma)/Phi(-g_cur/sigma)); the chain state (cur, g_cur) is updated in
 * place and appended to the new population at columns out_col0 + k.  out4 = [0, sum alpha, #accepted, nchain]
 * (SIS_aCS.py:238-254) */
int ree_acs_accept(double* cur, double* g_cur, const double* cand, const double* g_cand, int64_t nchain, int d,
                   int64_t ld, double sigma, uint64_t seed, uint64_t draw, int64_t first, double* out_x, int64_t ld_out,
                   int64_t out_col0, double* out_g, double* workspace, double* out4, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* REE_B200_H */
