// Device helpers shared by the DMMA kernels (ree_mma.cu: fused sweeps, ree_split.cu: split transform / Gram kernels).
#pragma once
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#include <stdlib.h>

#include "common.cuh"

#define MAXM_MAX 8  // macro tiles (2x2 tiles of 8x8) per warp, upper bound of the plan table

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

// which macro tiles (mr >= mc) of the lower triangle each warp of a k-split group accumulates
struct GramPlan {
    int nmac[16];
    unsigned short mac[16][MAXM_MAX];  // mr | mc << 8
};

// ---------------------------------------------------------------------------------------
// shared-memory access with explicit 32-bit shared addresses (immediate offsets, no generic->shared conversion in
// the inner loops) and cp.async staging
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double lds64(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ double2 lds128(uint32_t a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, double2 v) {
    asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(a), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// mbarrier + bulk-copy (TMA engine, non-tensor form) staging: one thread arms the barrier with the byte count, a
// few threads issue whole-row copies, consumers spin on the phase parity.  No per-lane LDGSTS instructions and no
// register staging; completion is tracked by the hardware transaction count.
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// one-instruction 2-D tile load through the TMA engine: box (c0.., c1..) of the tensor map -> dense box in shared memory
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tmap, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}

// fragment-ordered copy of a d x d matrix: tile (mt, kk) (rows 8mt.., columns 4kk..) -> 32 consecutive doubles,
// lane l holds A[8mt + l/4][4kk + l%4].  TRI keeps only the tiles of the lower triangle (kk <= 2mt+1).
template <int NTB, bool TRI>
__host__ __device__ constexpr int lf_tile(int mt, int kk) {
    return TRI ? mt * (mt + 1) + kk : mt * 2 * NTB + kk;
}
template <int NTB, bool TRI>
__host__ __device__ constexpr int lf_doubles() {
    return (TRI ? NTB * (NTB + 1) : 2 * NTB * NTB) * 32;
}

template <int NTB, bool TRI, int MT>
__device__ __forceinline__ void pack_matrix(double* Lf, const double* __restrict__ A, int d) {
    for (int mt = 0; mt < NTB; ++mt) {
        const int nk = TRI ? 2 * mt + 2 : 2 * NTB;
        for (int idx = threadIdx.x; idx < nk * 32; idx += MT) {
            const int kk = idx >> 5, l = idx & 31;
            const int r = 8 * mt + (l >> 2), c = 4 * kk + (l & 3);
            Lf[lf_tile<NTB, TRI>(mt, kk) * 32 + l] = (A && r < d && c < d) ? A[(size_t)r * d + c] : 0.0;
        }
    }
}

// C[mt] += A(mt, 2q..2q+1) * B(q) for all k-pairs q.  B fragments come from `gen` (Philox / shared tile / global) in
// batches of GB k-pairs: the GB generator chains of batch k+1 are independent of each other and of the DMMAs of batch
// k, so their instruction streams interleave (one Philox + Box-Muller chain alone is ~500 cycles of dependent
// latency, longer when its fp64 instructions queue behind DMMAs on the shared pipe).
// Within a k-pair all tiles take their first k-step before any takes its second: consecutive DMMAs on one accumulator
// are never issued back to back (a dependent DMMA waits out the full pipe latency).
// Fully unrolled: with TRI the zero tiles mt < q vanish at compile time.
template <int NTB, bool TRI, bool EXACT, int GB, class Gen>
__device__ __forceinline__ void transform_tiles(uint32_t lf_lane_addr, int ntd, Gen& gen, double (&c)[NTB][2]) {
    constexpr int NB = (NTB + GB - 1) / GB;
    double pb0[2][GB], pb1[2][GB];  // B fragments of the current and the next batch
#pragma unroll
    for (int j = 0; j < GB; ++j)
        if (j < NTB && (EXACT || j < ntd)) gen(j, pb0[0][j], pb1[0][j]);
#pragma unroll
    for (int k = 0; k < NB; ++k) {
        if (k + 1 < NB) {
#pragma unroll
            for (int j = 0; j < GB; ++j) {
                const int q = (k + 1) * GB + j;
                if (q < NTB && (EXACT || q < ntd)) gen(q, pb0[(k + 1) & 1][j], pb1[(k + 1) & 1][j]);
            }
        }
#pragma unroll
        for (int j = 0; j < GB; ++j) {
            const int q = k * GB + j;
            if (q < NTB && (EXACT || q < ntd)) {  // EXACT (ntd == NTB): no guards, one straight-line block
                const double b0 = pb0[k & 1][j], b1 = pb1[k & 1][j];
#pragma unroll
                for (int mt = (TRI ? q : 0); mt < NTB; ++mt)
                    if (EXACT || mt < ntd)
                        dmma(c[mt][0], c[mt][1], lds64(lf_lane_addr + lf_tile<NTB, TRI>(mt, 2 * q) * 256), b0);
#pragma unroll
                for (int mt = (TRI ? q : 0); mt < NTB; ++mt)
                    if (EXACT || mt < ntd)
                        dmma(c[mt][0], c[mt][1], lds64(lf_lane_addr + lf_tile<NTB, TRI>(mt, 2 * q + 1) * 256), b1);
            }
        }
    }
}

// packed tile index of (mt, nt), mt >= nt
__host__ __device__ __forceinline__ int tri_index(int mt, int nt) { return mt * (mt + 1) / 2 + nt; }

// B-fragment sources ------------------------------------------------------------------------
struct GenPhilox {  // z ~ N(0,1): counter (global particle, dim pair 4q+lc, draw)
    uint64_t seed, draw, particle;
    uint32_t lc;
    BmTableShared tab;
    __device__ __forceinline__ void operator()(int q, double& b0, double& b1) const {
        philox_normal_pair(seed, particle, (uint32_t)(4 * q) + lc, draw, tab, b0, b1);
    }
};

struct GenGlobal {  // rows 8q+lc and 8q+4+lc of an SoA array at this lane's particle (injected noise)
    const double* base;  // &z[lc*ld + particle] (NULL if the particle is out of range)
    int64_t ld;
    int d, lc;
    __device__ __forceinline__ void operator()(int q, double& b0, double& b1) const {
        const int j0 = 8 * q + lc, j1 = j0 + 4;
        b0 = (base && j0 < d) ? __ldg(base + (int64_t)(8 * q) * ld) : 0.0;
        b1 = (base && j1 < d) ? __ldg(base + (int64_t)(8 * q + 4) * ld) : 0.0;
    }
};


// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
static inline void make_plan(GramPlan& plan, int nta, int warps_per_group, int maxm) {
    // longest-processing-time assignment of the lower-triangle macro tiles to the warps of a group
    int MG = (nta + 1) / 2;
    struct Item { int mr, mc, w; };
    Item items[128];
    int n = 0;
    for (int mr = 0; mr < MG; ++mr)
        for (int mc = 0; mc <= mr; ++mc) {
            int w = 0;
            for (int q = 0; q < 4; ++q) {
                int mt = 2 * mr + (q >> 1), nt = 2 * mc + (q & 1);
                if (mt < nta && nt <= mt) ++w;
            }
            items[n++] = {mr, mc, w};
        }
    for (int i = 0; i < n; ++i)  // sort by weight, descending (stable insertion sort; n <= 55)
        for (int j = i; j > 0 && items[j].w > items[j - 1].w; --j) {
            Item t = items[j]; items[j] = items[j - 1]; items[j - 1] = t;
        }
    int load[16] = {0};
    for (int w = 0; w < 16; ++w) plan.nmac[w] = 0;
    for (int i = 0; i < n; ++i) {
        int best = -1;
        for (int w = 0; w < warps_per_group; ++w)
            if (plan.nmac[w] < maxm && (best < 0 || load[w] < load[best])) best = w;
        plan.mac[best][plan.nmac[best]++] = (unsigned short)(items[i].mr | (items[i].mc << 8));
        load[best] += items[i].w;
    }
}


// Tensor map of an SoA ensemble (fp64, dims {ld, d}, row pitch ld*8 bytes) with a {box_cols, box_rows} box; rows
// beyond d and columns beyond ld are zero-filled by the TMA engine.  Returns 0 on success.
static inline int make_ensemble_tmap(CUtensorMap* map, const double* x, int64_t ld, int d, int box_cols, int box_rows) {
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p)
            return ree_set_error(REE_E_UNSUPPORTED, "cuTensorMapEncodeTiled is not available in this driver");
        fn = (encode_fn)p;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)d};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(x), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return ree_set_error(REE_E_BADARG, "cuTensorMapEncodeTiled failed (ensemble tensor map)");
    return 0;
}

// second-stage reductions (defined in ree_mma.cu)
int ree_mma_unpack(const double* partials, int nparts, int d, int weighted, double* sums, cudaStream_t st);
int ree_mma_sum_small(const double* small, int n, double* out, cudaStream_t st);
int ree_mma_merge_small(const double* small, int n, double* out4, cudaStream_t st);
