// Scalar optimisers of the sigma / beta adaptation as ask/tell state machines (host + device).
//
// The reference adapts sigma with scipy.optimize.minimize_scalar(method="Bounded") (solver.py:548, SciPy's
// `_minimize_scalar_bounded`: Brent's fminbound, xatol 1e-5, maxiter 500) and beta with scipy.optimize.root (solver.py:601,
// MINPACK hybrd with n = 1, xtol 1.49012e-8, maxfev 400, factor 100, epsfcn = eps, mode 1).  Both are sequential scalar
// algorithms whose every objective evaluation is an N-sized reduction; driven from Python each evaluation costs a kernel
// launch, a readback and ~25 us of interpreter time.  Here the same algorithms -- same operation order, no FMA contraction
// (the translation unit is compiled with -fmad=false), IEEE double -- are coroutines: `start()` / `tell(f)` hand out the
// next trial point, so a persistent kernel (ree_opt.cu) can run the whole search next to the reductions.  The host build
// of the same code is exported as ree_scalar_brent / ree_scalar_hybr1 and checked against SciPy bit for bit on CPU
// (tests/test_scalar_opt.py).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define REE_HD __host__ __device__ __forceinline__
#else
#define REE_HD inline
#endif

namespace ree_opt {

REE_HD double nan_max(double a, double b) {  // numpy.maximum: NaN wins
    return (a != a) ? a : ((b != b) ? b : (a > b ? a : b));
}
REE_HD double sign_or_one(double v) {  // numpy.sign(v) + (v == 0)
    return v > 0.0 ? 1.0 : (v < 0.0 ? -1.0 : (v == 0.0 ? 1.0 : v));
}

// ------------------------------------------------------------------------------------------------------------------
// SciPy `_minimize_scalar_bounded`
// ------------------------------------------------------------------------------------------------------------------
struct BrentBounded {
    double a, b, fulc, nfc, xf, rat, e, x, fx, fu, ffulc, fnfc, xm, tol1, tol2, xatol;
    int num, maxfun, flag, started;

    // returns the first trial point
    REE_HD double start(double x1, double x2, double xatol_, int maxiter) {
        const double golden_mean = 0.5 * (3.0 - sqrt(5.0));
        xatol = xatol_;
        maxfun = maxiter;
        flag = 0;
        a = x1;
        b = x2;
        fulc = a + golden_mean * (b - a);
        nfc = fulc;
        xf = fulc;
        rat = 0.0;
        e = 0.0;
        x = xf;
        num = 0;
        started = 0;
        fx = fu = ffulc = fnfc = 0.0;
        return x;
    }

    // the loop head of the `while`: true if another evaluation (at `x`) is needed
    REE_HD bool propose() {
        const double golden_mean = 0.5 * (3.0 - sqrt(5.0));
        if (!(fabs(xf - xm) > (tol2 - 0.5 * (b - a)))) return false;
        int golden = 1;
        if (fabs(e) > tol1) {  // parabolic fit
            golden = 0;
            double r = (xf - nfc) * (fx - ffulc);
            double q = (xf - fulc) * (fx - fnfc);
            double p = (xf - fulc) * q - (xf - nfc) * r;
            q = 2.0 * (q - r);
            if (q > 0.0) p = -p;
            q = fabs(q);
            r = e;
            e = rat;
            if ((fabs(p) < fabs(0.5 * q * r)) && (p > q * (a - xf)) && (p < q * (b - xf))) {
                rat = (p + 0.0) / q;
                x = xf + rat;
                if (((x - a) < tol2) || ((b - x) < tol2)) {
                    const double si = sign_or_one(xm - xf);
                    rat = tol1 * si;
                }
            } else {
                golden = 1;
            }
        }
        if (golden) {
            if (xf >= xm) e = a - xf;
            else e = b - xf;
            rat = golden_mean * e;
        }
        const double si = sign_or_one(rat);
        x = xf + si * nan_max(fabs(rat), tol1);
        return true;
    }

    // objective value at the point handed out last; true if `x` holds a new trial point
    REE_HD bool tell(double f) {
        const double sqrt_eps = sqrt(2.2e-16);
        if (!started) {
            started = 1;
            fx = f;
            num = 1;
            fu = INFINITY;
            ffulc = fnfc = fx;
        } else {
            fu = f;
            num += 1;
            if (fu <= fx) {
                if (x >= xf) a = xf;
                else b = xf;
                fulc = nfc; ffulc = fnfc;
                nfc = xf; fnfc = fx;
                xf = x; fx = fu;
            } else {
                if (x < xf) a = x;
                else b = x;
                if ((fu <= fnfc) || (nfc == xf)) {
                    fulc = nfc; ffulc = fnfc;
                    nfc = x; fnfc = fu;
                } else if ((fu <= ffulc) || (fulc == xf) || (fulc == nfc)) {
                    fulc = x; ffulc = fu;
                }
            }
        }
        xm = 0.5 * (a + b);
        tol1 = sqrt_eps * fabs(xf) + xatol / 3.0;
        tol2 = 2.0 * tol1;
        if (started && num > 1 && num >= maxfun) {
            flag = 1;
            return finish();
        }
        if (propose()) return true;
        return finish();
    }

    REE_HD bool finish() {
        if (xf != xf || fx != fx || fu != fu) flag = 2;
        return false;
    }
    REE_HD double result() const { return xf; }
};

// ------------------------------------------------------------------------------------------------------------------
// MINPACK hybrd specialised to n = 1 (ml = mu = 0, mode 1, nprint 0), as scipy.optimize.root(method="hybr") calls it.
// With one unknown: enorm(v) = |v|; qrfac leaves rdiag = -J and the Householder vector 2 (J != 0), qform gives Q = -1
// (Q = 1 if J == 0), r1mpyq has no rotations to apply, r1updt adds the spike to the single entry of R.
// ------------------------------------------------------------------------------------------------------------------
struct Hybr1 {
    enum { P_INIT = 0, P_JAC = 1, P_STEP = 2, P_DONE = 3 };
    double x, fvec, fnorm, xtol, factor, epsfcn;
    double diag, delta, xnorm, qtf, r, q;    // q: the 1x1 orthogonal factor (+-1)
    double h, wa1, wa2, wa3, pnorm, x_eval;
    int nfev, maxfev, iter, ncsuc, ncfail, nslow1, nslow2, info, phase, sing;
    bool jeval;

    REE_HD double start(double x0, double xtol_, int maxfev_, double factor_, double epsfcn_) {
        x = x0;
        xtol = xtol_;
        maxfev = maxfev_;
        factor = factor_;
        epsfcn = epsfcn_;
        info = 0;
        nfev = 0;
        phase = P_INIT;
        x_eval = x;
        if (xtol < 0.0 || maxfev <= 0 || !(factor > 0.0)) phase = P_DONE;  // improper input: info = 0
        return x_eval;
    }
    REE_HD bool done() const { return phase == P_DONE; }
    REE_HD double point() const { return x_eval; }
    REE_HD double result() const { return x; }
    REE_HD double fresult() const { return fvec; }

    REE_HD bool begin_outer() {  // fdjac1: forward difference
        const double epsmch = 2.220446049250313e-16;
        jeval = true;
        const double eps = sqrt(epsfcn > epsmch ? epsfcn : epsmch);
        h = eps * fabs(x);
        if (h == 0.0) h = eps;
        x_eval = x + h;
        phase = P_JAC;
        return true;
    }

    REE_HD void dogleg() {  // -> wa1 = step (before the sign flip)
        const double epsmch = 2.220446049250313e-16;
        double temp = r;
        if (temp == 0.0) {
            temp = fabs(r);
            temp = epsmch * temp;
            if (temp == 0.0) temp = epsmch;
        }
        double xs = (qtf - 0.0) / temp;  // Gauss-Newton direction
        double w1 = 0.0;
        double w2 = diag * xs;
        const double qnorm = fabs(w2);
        if (qnorm <= delta) {
            wa1 = xs;
            return;
        }
        // scaled gradient direction
        w1 = w1 + r * qtf;
        w1 = w1 / diag;
        const double gnorm = fabs(w1);
        double sgnorm = 0.0;
        double alpha = delta / qnorm;
        if (gnorm != 0.0) {
            w1 = (w1 / gnorm) / diag;
            w2 = 0.0 + r * w1;
            temp = fabs(w2);
            sgnorm = (gnorm / temp) / temp;
            alpha = 0.0;
            if (sgnorm < delta) {
                const double bnorm = fabs(qtf);
                temp = (bnorm / gnorm) * (bnorm / qnorm) * (sgnorm / delta);
                const double dq = delta / qnorm, sd = sgnorm / delta;
                const double t1 = temp - dq;
                temp = temp - dq * (sd * sd) + sqrt(t1 * t1 + (1.0 - dq * dq) * (1.0 - sd * sd));
                alpha = (dq * (1.0 - sd * sd)) / temp;
            }
        }
        temp = (1.0 - alpha) * (sgnorm < delta ? sgnorm : delta);
        wa1 = temp * w1 + alpha * xs;
    }

    REE_HD bool inner_propose() {
        dogleg();
        wa1 = -wa1;
        wa2 = x + wa1;
        wa3 = diag * wa1;
        pnorm = fabs(wa3);
        if (iter == 1) delta = delta < pnorm ? delta : pnorm;
        x_eval = wa2;
        phase = P_STEP;
        return true;
    }

    // value of the function at point(); true while another evaluation is requested
    REE_HD bool tell(double f) {
        const double epsmch = 2.220446049250313e-16;
        const double p1 = 0.1, p5 = 0.5, p001 = 1.0e-3, p0001 = 1.0e-4;
        if (phase == P_INIT) {
            fvec = f;
            nfev = 1;
            fnorm = fabs(fvec);
            iter = 1;
            ncsuc = ncfail = nslow1 = nslow2 = 0;
            return begin_outer();
        }
        if (phase == P_JAC) {
            double fjac = (f - fvec) / h;
            nfev += 1;
            // qrfac
            const double acnorm = fabs(fjac);
            double ajnorm = fabs(fjac);
            double hh = fjac;  // Householder entry left in fjac
            if (ajnorm != 0.0) {
                if (fjac < 0.0) ajnorm = -ajnorm;
                hh = fjac / ajnorm;
                hh = hh + 1.0;
            }
            const double rdiag = -ajnorm;
            if (iter == 1) {
                diag = acnorm;
                if (acnorm == 0.0) diag = 1.0;
                xnorm = fabs(diag * x);
                delta = factor * xnorm;
                if (delta == 0.0) delta = factor;
            }
            // (q transpose) * fvec
            double w4 = fvec;
            if (hh != 0.0) {
                const double sum = hh * w4;
                const double temp = -sum / hh;
                w4 = w4 + hh * temp;
            }
            qtf = w4;
            r = rdiag;
            sing = (rdiag == 0.0);
            // qform
            q = 1.0;
            if (hh != 0.0) {
                const double sum = 1.0 * hh;
                const double temp = sum / hh;
                q = 1.0 - temp * hh;
            }
            diag = diag > acnorm ? diag : acnorm;
            return inner_propose();
        }
        if (phase == P_STEP) {
            const double w4 = f;
            nfev += 1;
            const double fnorm1 = fabs(w4);
            double actred = -1.0;
            if (fnorm1 < fnorm) {
                const double t = fnorm1 / fnorm;
                actred = 1.0 - t * t;
            }
            wa3 = qtf + (0.0 + r * wa1);
            double temp = fabs(wa3);
            double prered = 0.0;
            if (temp < fnorm) {
                const double t = temp / fnorm;
                prered = 1.0 - t * t;
            }
            double ratio = 0.0;
            if (prered > 0.0) ratio = actred / prered;
            if (ratio < p1) {
                ncsuc = 0;
                ncfail += 1;
                delta = p5 * delta;
            } else {
                ncfail = 0;
                ncsuc += 1;
                if (ratio >= p5 || ncsuc > 1) {
                    const double t = pnorm / p5;
                    delta = delta > t ? delta : t;
                }
                if (fabs(ratio - 1.0) <= p1) delta = pnorm / p5;
            }
            if (ratio >= p0001) {  // successful iteration
                x = wa2;
                wa2 = diag * x;
                fvec = w4;
                xnorm = fabs(wa2);
                fnorm = fnorm1;
                iter += 1;
            }
            nslow1 += 1;
            if (actred >= p001) nslow1 = 0;
            if (jeval) nslow2 += 1;
            if (actred >= p1) nslow2 = 0;
            if (delta <= xtol * xnorm || fnorm == 0.0) info = 1;
            if (info != 0) {
                phase = P_DONE;
                return false;
            }
            if (nfev >= maxfev) info = 2;
            {
                const double t = p1 * delta;
                if (p1 * (t > pnorm ? t : pnorm) <= epsmch * xnorm) info = 3;
            }
            if (nslow2 == 5) info = 4;
            if (nslow1 == 10) info = 5;
            if (info != 0) {
                phase = P_DONE;
                return false;
            }
            if (ncfail == 2) return begin_outer();
            // rank one modification of the jacobian (Broyden)
            const double sum = q * w4;
            wa2 = (sum - wa3) / pnorm;
            wa1 = diag * ((diag * wa1) / pnorm);
            if (ratio >= p0001) qtf = sum;
            r = r + wa2 * wa1;  // r1updt
            sing = (r == 0.0);
            jeval = false;
            return inner_propose();
        }
        return false;
    }
};

}  // namespace ree_opt
