// Streaming access patterns and grid finalisation shared by the N-vector kernels (ree_nvec.cu, ree_opt.cu).
#pragma once
#include "common.cuh"

// ---------------------------------------------------------------------------------------
// Streaming access pattern of the N-vector kernels.  One 8-byte load per thread and iteration keeps ~0.6 MB in flight
// on the whole GPU (0.9 TB/s by Little's law); here every thread owns groups of 4 consecutive elements, fetched as two
// 16-byte loads, and requests its next group before it works on the current one; with 8 resident CTAs per SM that is
// ~10 MB in flight.  `body(i, v0)` / `body2(i, v0, v1)` see every element exactly once (order differs from the scalar
// loop only across threads; the per-thread order is ascending).
// ---------------------------------------------------------------------------------------
struct Quad {
    double v[4];
};
__device__ __forceinline__ Quad ldg_quad(const double* p) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(p)), b = __ldg(reinterpret_cast<const double2*>(p) + 1);
    return Quad{{a.x, a.y, b.x, b.y}};
}
__device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// quad streams: `quad(i0, q...)` sees four consecutive elements starting at i0 (a multiple of 4), `one(i, ...)` the tail
template <class QuadFn, class OneFn>
__device__ __forceinline__ void stream1q(const double* __restrict__ a, int64_t n, QuadFn quad, OneFn one) {
    const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t nq = aligned16(a) ? n / 4 : 0;
    int64_t q = gt;
    Quad cur{};
    if (q < nq) cur = ldg_quad(a + 4 * q);
    for (; q < nq; q += stride) {
        const Quad now = cur;
        if (q + stride < nq) cur = ldg_quad(a + 4 * (q + stride));
        quad(4 * q, now);
    }
    for (int64_t i = 4 * nq + gt; i < n; i += stride) one(i, a[i]);
}

template <class QuadFn, class OneFn>
__device__ __forceinline__ void stream2q(const double* __restrict__ a, const double* __restrict__ b, int64_t n,
                                         QuadFn quad, OneFn one) {
    const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t nq = (aligned16(a) && aligned16(b)) ? n / 4 : 0;
    int64_t q = gt;
    Quad ca{}, cb{};
    if (q < nq) {
        ca = ldg_quad(a + 4 * q);
        cb = ldg_quad(b + 4 * q);
    }
    for (; q < nq; q += stride) {
        const Quad na = ca, nb = cb;
        if (q + stride < nq) {
            ca = ldg_quad(a + 4 * (q + stride));
            cb = ldg_quad(b + 4 * (q + stride));
        }
        quad(4 * q, na, nb);
    }
    for (int64_t i = 4 * nq + gt; i < n; i += stride) one(i, a[i], b[i]);
}

template <class Body>
__device__ __forceinline__ void stream1(const double* __restrict__ a, int64_t n, Body body) {
    const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t nq = aligned16(a) ? n / 4 : 0;
    int64_t q = gt;
    Quad cur{};
    if (q < nq) cur = ldg_quad(a + 4 * q);
    for (; q < nq; q += stride) {
        const Quad now = cur;
        if (q + stride < nq) cur = ldg_quad(a + 4 * (q + stride));
#pragma unroll
        for (int k = 0; k < 4; ++k) body(4 * q + k, now.v[k]);
    }
    for (int64_t i = 4 * nq + gt; i < n; i += stride) body(i, a[i]);
}

template <class Body>
__device__ __forceinline__ void stream2(const double* __restrict__ a, const double* __restrict__ b, int64_t n, Body body) {
    const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t nq = (aligned16(a) && aligned16(b)) ? n / 4 : 0;
    int64_t q = gt;
    Quad ca{}, cb{};
    if (q < nq) {
        ca = ldg_quad(a + 4 * q);
        cb = ldg_quad(b + 4 * q);
    }
    for (; q < nq; q += stride) {
        const Quad na = ca, nb = cb;
        if (q + stride < nq) {
            ca = ldg_quad(a + 4 * (q + stride));
            cb = ldg_quad(b + 4 * (q + stride));
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) body(4 * q + k, na.v[k], nb.v[k]);
    }
    for (int64_t i = 4 * nq + gt; i < n; i += stride) body(i, a[i], b[i]);
}
