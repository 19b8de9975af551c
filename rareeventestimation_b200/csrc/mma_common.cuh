// Device helpers shared by the DMMA kernels (ree_mma.cu: fused sweeps, ree_split.cu: split transform / Gram kernels).
#pragma once
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

// C[mt] += A(mt, 2q..2q+1) * B(q) for all k-pairs q.  B fragments come from `gen` (Philox / shared tile / global),
// generated LA k-pairs ahead so that their instruction stream interleaves with the DMMAs of the current pair.
// Fully unrolled: with TRI the zero tiles mt < q vanish at compile time.
template <int NTB, bool TRI, bool EXACT, int LA, class Gen>
__device__ __forceinline__ void transform_tiles(uint32_t lf_lane_addr, int ntd, Gen& gen, double (&c)[NTB][2]) {
    double pb0[LA], pb1[LA];  // B fragments of the next LA k-pairs (in flight)
#pragma unroll
    for (int l = 0; l < LA; ++l) gen(l, pb0[l], pb1[l]);
#pragma unroll
    for (int q = 0; q < NTB; ++q) {
        if (EXACT || q < ntd) {  // EXACT (ntd == NTB): no guards, one straight-line block
            const double b0 = pb0[q % LA], b1 = pb1[q % LA];
            if (q + LA < NTB) gen(q + LA, pb0[q % LA], pb1[q % LA]);
#pragma unroll
            for (int mt = (TRI ? q : 0); mt < NTB; ++mt) {
                if (EXACT || mt < ntd) {
                    const double a0 = lds64(lf_lane_addr + lf_tile<NTB, TRI>(mt, 2 * q) * 256);
                    const double a1 = lds64(lf_lane_addr + lf_tile<NTB, TRI>(mt, 2 * q + 1) * 256);
                    dmma(c[mt][0], c[mt][1], a0, b0);
                    dmma(c[mt][0], c[mt][1], a1, b1);
                }
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


// second-stage reductions (defined in ree_mma.cu)
int ree_mma_unpack(const double* partials, int nparts, int d, int weighted, double* sums, cudaStream_t st);
int ree_mma_sum_small(const double* small, int n, double* out, cudaStream_t st);
int ree_mma_merge_small(const double* small, int n, double* out4, cudaStream_t st);
