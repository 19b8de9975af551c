// Point-wise limit-state functions on register values (shared by the streaming LSF kernel and the small-d sweeps).
// Operation order mirrors the NumPy expressions of problem/toy_problems.py; _rn intrinsics keep nvcc from contracting
// mul+add into FMA (NumPy does not fuse).
#pragma once
#include "common.cuh"

#define REE_LSF_LINEAR_SUM_INTERNAL 100

__device__ __forceinline__ double lsf_oscillator(double x0, double x1, double x2, double x3, double x4, double x5) {
    // toy_problems.py:72-80
    double m = __dadd_rn(1.0, __dmul_rn(0.05, x0));
    double c1 = __dadd_rn(1.0, __dmul_rn(0.1, x1));
    double c2 = __dadd_rn(0.1, __dmul_rn(0.01, x2));
    double r = __dadd_rn(0.5, __dmul_rn(0.05, x3));
    double f1 = __dadd_rn(0.3, __dmul_rn(0.2, x4));
    double t1 = __dadd_rn(1.0, __dmul_rn(0.2, x5));
    double w2 = __ddiv_rn(__dadd_rn(c1, c2), m);  // omega^2 before the complex sqrt
    double amp;
    if (w2 >= 0.0) {
        double om = __dsqrt_rn(w2);
        double om2 = __dmul_rn(om, om);  // omega**2 as NumPy forms it
        double a = __ddiv_rn(__dmul_rn(2.0, f1), __dmul_rn(m, om2));
        amp = fabs(__dmul_rn(a, sin(__ddiv_rn(__dmul_rn(om, t1), 2.0))));
    } else {  // omega = i*b: omega^2 = -b^2, sin(i y) = i sinh(y)
        double b = __dsqrt_rn(-w2);
        double om2 = -__dmul_rn(b, b);
        double a = __ddiv_rn(__dmul_rn(2.0, f1), __dmul_rn(m, om2));
        amp = fabs(__dmul_rn(a, sinh(__ddiv_rn(__dmul_rn(b, t1), 2.0))));
    }
    return __dsub_rn(__dmul_rn(3.0, r), amp);
}


// g(x) for the closed-form kinds with the point in registers; `coef` may live in shared or global memory.
template <int D>
__device__ __forceinline__ double lsf_point(int kind, const double* coef, double offset, const double (&x)[D]) {
    switch (kind) {
        case REE_LSF_AFFINE: {  // g = offset + sum_j coef_j x_j
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < D; ++j) s = __dadd_rn(s, __dmul_rn(coef[j], x[j]));
            return __dadd_rn(s, offset);
        }
        case REE_LSF_LINEAR_SUM_INTERNAL: {  // -sum(x)/sqrt(d) + beta, toy_problems.py:36-37
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < D; ++j) s = __dadd_rn(s, x[j]);
            return __dadd_rn(__ddiv_rn(-s, __dsqrt_rn((double)D)), offset);
        }
        case REE_LSF_CONVEX: {  // 0.1*(x0-x1)**2 - 1/sqrt(2)*sum(x) + 2.5, toy_problems.py:12-13
            const double x0 = x[0], x1 = x[D > 1 ? 1 : 0];
            const double df = __dsub_rn(x0, x1);
            const double a = __dmul_rn(0.1, __dmul_rn(df, df));
            double s = __dadd_rn(x0, x1);
#pragma unroll
            for (int j = 2; j < D; ++j) s = __dadd_rn(s, x[j]);
            const double b = __dmul_rn(__ddiv_rn(1.0, __dsqrt_rn(2.0)), s);
            return __dadd_rn(__dsub_rn(a, b), 2.5);
        }
        case REE_LSF_FUJITA_RACKWITZ: {  // -sum(log(Phi(-x))) - c, toy_problems.py:57-58
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < D; ++j) s = __dadd_rn(s, log(normcdf(-x[j])));
            return __dsub_rn(-s, offset);
        }
        case REE_LSF_OSCILLATOR:
            return lsf_oscillator(x[0], x[D > 1 ? 1 : 0], x[D > 2 ? 2 : 0], x[D > 3 ? 3 : 0], x[D > 4 ? 4 : 0],
                                  x[D > 5 ? 5 : 0]);
        default:
            return CUDART_NAN;
    }
}
