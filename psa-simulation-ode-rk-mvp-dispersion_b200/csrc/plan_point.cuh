// plan_point.cuh -- per-scan-point frequency plan, validity and Delta-beta on the device.
// Shared by the stand-alone table kernel (frontend.cu) and the fused sweep kernel (yaman4.cu).
//
// Restates, for one point (paths relative to the reference checkout):
//   frequency_plan.plan_from_wavelengths       frequency_plan.py:291-327
//   frequency_plan.enforce_energy_conservation frequency_plan.py:112-131
//   frequency_plan.infer_symmetry_from_omegas  frequency_plan.py:215-255 (+ SymmetricPlan :134-199)
//   phase_matching.compute_phase_mismatch      phase_matching.py:150-215
//   dispersion.beta_taylor                     dispersion.py:233-279
//   dispersion.delta_beta_from_omegas          dispersion.py:282-318
//   dispersion.delta_beta_symmetric            dispersion.py:321-372
// Explicitly rounded operations (__dmul_rn / __dadd_rn / IEEE division) in the reference's order,
// no FMA contraction: Delta-beta is a difference of nearly equal terms.  Integer powers x**n (libm
// pow in the reference) are evaluated in double-double and rounded once.
#pragma once
#include "fpa_common.cuh"

namespace fpa {

struct PlanParams {
    int64_t       n1, n3;
    const double* lambda1;
    const double* lambda2;
    const double* lambda3;
    int           lambda2_stride;
    int           method, max_order, n_even;
    int           even_orders[FPA_MAX_TAYLOR_ORDER];
    double        beta[FPA_MAX_TAYLOR_ORDER + 1];
    double        omega_ref, atol, rtol, provided;
    double*       omega;
    double*       dbeta;
    double*       dbeta_masked;  // optional: dbeta with 0 at invalid points (integrator input)
    int32_t*      valid;
};

static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
static __device__ __forceinline__ double sub(double a, double b) { return __dadd_rn(a, -b); }

// x**n, n >= 0, as the correctly rounded value of the exact power (double-double accumulate).
static __device__ double pow_int(double x, int n) {
    if (n == 0) return 1.0;
    double hi = x, lo = 0.0;
    for (int i = 1; i < n; ++i) {
        // (hi + lo) * x  ->  (ph, pl)
        const double ph = mul(hi, x);
        const double pe = fma(hi, x, -ph);       // exact error of hi*x
        const double pl = fma(lo, x, pe);        // + lo*x
        const double s  = add(ph, pl);           // renormalise
        lo = sub(pl, sub(s, ph));
        hi = s;
    }
    return hi;  // hi = RN(hi + lo)
}

static __device__ __forceinline__ double factorial_d(int n) {
    double f = 1.0;
    for (int i = 2; i <= n; ++i) f *= (double)i;  // exact for n <= 18
    return f;
}

// numpy.isclose(a, b, rtol, atol) for finite inputs: |a-b| <= atol + rtol*|b|
static __device__ __forceinline__ bool isclose(double a, double b, double atol, double rtol) {
    return fabs(sub(a, b)) <= add(atol, mul(rtol, fabs(b)));
}

static __device__ __forceinline__ bool pos_finite(double v) { return v > 0.0 && !nonfinite(v); }

// beta(omega) Taylor sum, zero coefficients skipped, term = ((bn * dw**n) / n!)  (dispersion.py:271-275)
static __device__ double beta_taylor_dev(double w, const PlanParams& p, const double* beta) {
    const double dw = sub(w, p.omega_ref);
    double out = 0.0;
    for (int n = 0; n <= p.max_order && n <= FPA_MAX_TAYLOR_ORDER; ++n) {
        const double bn = beta[n];
        if (bn == 0.0) continue;
        out = add(out, mul(bn, pow_int(dw, n)) / factorial_d(n));
    }
    return out;
}

// omega[4] of one scan point and whether the reference would accept the plan
// (frequency_plan.plan_from_wavelengths, frequency_plan.py:291-327).
static __device__ bool plan_omegas(double lam1, double lam2, double lam3, double (&w)[4]) {
    // omega = 2*pi*c / lambda, evaluated as (2pi*c)/lambda   (frequency_plan.py:89-92)
    const double two_pi_c = mul(mul(2.0, 3.141592653589793), 299792458.0);
    bool ok = pos_finite(lam1) && pos_finite(lam2) && pos_finite(lam3);
    w[0] = two_pi_c / lam1;
    w[1] = two_pi_c / lam2;
    w[2] = two_pi_c / lam3;
    w[3] = sub(add(w[0], w[1]), w[2]);  // frequency_plan.py:314
    ok = ok && pos_finite(w[0]) && pos_finite(w[1]) && pos_finite(w[2]) && pos_finite(w[3]);
    // plan_from_wavelengths checks with its default tolerances (atol 0, rtol 1e-12)
    return ok && isclose(add(w[0], w[1]), add(w[2], w[3]), 0.0, 1e-12);
}

// Delta-beta of one accepted plan with the coefficient table `beta` / constant `provided`
// (phase_matching.compute_phase_mismatch, phase_matching.py:150-215).  Returns NaN and clears
// `ok` where the reference raises.
static __device__ double plan_dbeta(const PlanParams& p, const double* beta, double provided,
                                    const double (&w)[4], bool& ok) {
    const double w1 = w[0], w2 = w[1], w3 = w[2], w4 = w[3];
    double db = qnan();
    if (!ok) return db;
    if (p.method == FPA_PM_PROVIDED) {
        db = provided;
    } else if (p.method == FPA_PM_GENERAL_TAYLOR) {
        if (isclose(add(w1, w2), add(w3, w4), p.atol, p.rtol)) {  // dispersion.py:304-310
            const double b1 = beta_taylor_dev(w1, p, beta), b2 = beta_taylor_dev(w2, p, beta);
            const double b3 = beta_taylor_dev(w3, p, beta), b4 = beta_taylor_dev(w4, p, beta);
            db = sub(add(b3, b4), add(b1, b2));  // dispersion.py:318
        } else {
            ok = false;
        }
    } else {  // SYMMETRIC_EVEN
        ok = isclose(add(w1, w2), add(w3, w4), p.atol, p.rtol);  // frequency_plan.py:238-240
        const double oc = mul(0.5, add(w1, w2));
        const double od = mul(0.5, sub(w1, w2));
        const double Om = sub(w3, oc);
        ok = ok && pos_finite(oc) && fabs(od) < oc;              // frequency_plan.py:149-159
        const double s1 = add(oc, od), s2 = sub(oc, od), s3 = add(oc, Om), s4 = sub(oc, Om);
        ok = ok && s1 > 0.0 && s2 > 0.0 && s3 > 0.0 && s4 > 0.0;  // :189-195
        ok = ok && isclose(add(s1, s2), add(s3, s4), 0.0, 1e-12);  // :196
        ok = ok && isclose(s4, w4, p.atol, p.rtol);                // :249-253
        if (ok) {
            double out = 0.0;
            for (int e = 0; e < p.n_even; ++e) {
                const int    n  = p.even_orders[e];
                const double bn = (n <= FPA_MAX_TAYLOR_ORDER) ? beta[n] : 0.0;
                if (bn == 0.0) continue;
                // ((bn * (Om**n - od**n)) * 2.0) / n!      (dispersion.py:370)
                const double diff = sub(pow_int(Om, n), pow_int(od, n));
                out = add(out, mul(mul(bn, diff), 2.0) / factorial_d(n));
            }
            db = out;
        }
    }
    if (ok && nonfinite(db)) ok = false;  // float() of a non-finite dbeta is rejected downstream
    return ok ? db : qnan();
}

}  // namespace fpa
