// frontend.cu -- sweep front-end on the device: frequency plan -> validity -> Delta-beta table,
// and the gain metric epilogue of the sweeps.
//
// Vectorised stand-in for the per-point Python of the reference (paths relative to its checkout):
//   frequency_plan.plan_from_wavelengths       frequency_plan.py:291-327
//   frequency_plan.enforce_energy_conservation frequency_plan.py:112-131
//   frequency_plan.infer_symmetry_from_omegas  frequency_plan.py:215-255 (+ SymmetricPlan :134-199)
//   phase_matching.compute_phase_mismatch      phase_matching.py:150-215
//   dispersion.beta_taylor                     dispersion.py:233-279
//   dispersion.delta_beta_from_omegas          dispersion.py:282-318
//   dispersion.delta_beta_symmetric            dispersion.py:321-372
//   sweep metric                               scan_mismtach.py:723-734
//
// Everything here uses explicitly rounded operations (__dmul_rn / __dadd_rn / IEEE division) in the
// reference's operation order -- no FMA contraction -- because Delta-beta is a difference of nearly
// equal terms.  Integer powers x**n (libm pow in the reference) are evaluated in double-double and
// rounded once, which reproduces a correctly rounded pow.
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

__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dadd_rn(a, -b); }

// x**n, n >= 0, as the correctly rounded value of the exact power (double-double accumulate).
__device__ double pow_int(double x, int n) {
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

__device__ __forceinline__ double factorial_d(int n) {
    double f = 1.0;
    for (int i = 2; i <= n; ++i) f *= (double)i;  // exact for n <= 18
    return f;
}

// numpy.isclose(a, b, rtol, atol) for finite inputs: |a-b| <= atol + rtol*|b|
__device__ __forceinline__ bool isclose(double a, double b, double atol, double rtol) {
    return fabs(sub(a, b)) <= add(atol, mul(rtol, fabs(b)));
}

__device__ __forceinline__ bool pos_finite(double v) { return v > 0.0 && !nonfinite(v); }

// beta(omega) Taylor sum, zero coefficients skipped, term = ((bn * dw**n) / n!)  (dispersion.py:271-275)
__device__ double beta_taylor_dev(double w, const PlanParams& p) {
    const double dw = sub(w, p.omega_ref);
    double out = 0.0;
    for (int n = 0; n <= p.max_order && n <= FPA_MAX_TAYLOR_ORDER; ++n) {
        const double bn = p.beta[n];
        if (bn == 0.0) continue;
        out = add(out, mul(bn, pow_int(dw, n)) / factorial_d(n));
    }
    return out;
}

__global__ void plan_dbeta_kernel(const PlanParams p) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t B = p.n1 * p.n3;
    if (b >= B) return;
    const int64_t i1 = b / p.n3, i3 = b - i1 * p.n3;

    const double lam1 = p.lambda1[i1];
    const double lam2 = p.lambda2[i1 * p.lambda2_stride];
    const double lam3 = p.lambda3[i3];

    // omega = 2*pi*c / lambda, evaluated as (2pi*c)/lambda   (frequency_plan.py:89-92)
    const double two_pi_c = mul(mul(2.0, 3.141592653589793), 299792458.0);
    bool ok = pos_finite(lam1) && pos_finite(lam2) && pos_finite(lam3);
    const double w1 = two_pi_c / lam1, w2 = two_pi_c / lam2, w3 = two_pi_c / lam3;
    const double w4 = sub(add(w1, w2), w3);  // frequency_plan.py:314
    ok = ok && pos_finite(w1) && pos_finite(w2) && pos_finite(w3) && pos_finite(w4);
    // plan_from_wavelengths checks with its default tolerances (atol 0, rtol 1e-12)
    ok = ok && isclose(add(w1, w2), add(w3, w4), 0.0, 1e-12);

    double db = qnan();
    if (ok) {
        if (p.method == FPA_PM_PROVIDED) {
            db = p.provided;
        } else if (p.method == FPA_PM_GENERAL_TAYLOR) {
            if (isclose(add(w1, w2), add(w3, w4), p.atol, p.rtol)) {  // dispersion.py:304-310
                const double b1 = beta_taylor_dev(w1, p), b2 = beta_taylor_dev(w2, p);
                const double b3 = beta_taylor_dev(w3, p), b4 = beta_taylor_dev(w4, p);
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
                    const double bn = (n <= FPA_MAX_TAYLOR_ORDER) ? p.beta[n] : 0.0;
                    if (bn == 0.0) continue;
                    // ((bn * (Om**n - od**n)) * 2.0) / n!      (dispersion.py:370)
                    const double diff = sub(pow_int(Om, n), pow_int(od, n));
                    out = add(out, mul(mul(bn, diff), 2.0) / factorial_d(n));
                }
                db = out;
            }
        }
        if (ok && nonfinite(db)) ok = false;  // float() of a non-finite dbeta is rejected downstream
        if (!ok) db = qnan();
    }

    if (p.dbeta) p.dbeta[b] = db;
    if (p.dbeta_masked) p.dbeta_masked[b] = ok ? db : 0.0;
    if (p.valid) p.valid[b] = ok ? 1 : 0;
    if (p.omega) {
        double2* o = reinterpret_cast<double2*>(p.omega + b * 4);
        o[0] = make_double2(w1, w2);
        o[1] = make_double2(w3, w4);
    }
}

int plan_launch(const fpa_plan_desc* d, double* dbeta_masked, cudaStream_t st) {
    FPA_REQUIRE(d != nullptr, "plan descriptor is NULL");
    FPA_REQUIRE(d->n1 >= 0 && d->n3 >= 0, "grid sizes must be >= 0");
    FPA_REQUIRE(d->lambda1 && d->lambda2 && d->lambda3, "wavelength axes must be set");
    FPA_REQUIRE(d->dbeta || dbeta_masked, "a dbeta output must be set");
    FPA_REQUIRE(d->method >= 0 && d->method <= 2, "unknown phase-matching method %d", d->method);
    FPA_REQUIRE(d->max_order >= 0, "max_order must be >= 0");
    FPA_REQUIRE(d->n_even >= 0 && d->n_even <= FPA_MAX_TAYLOR_ORDER, "too many even orders");
    FPA_REQUIRE((d->lambda2_stride | 1) == 1, "lambda2_stride must be 0 or 1");
    if (d->method == FPA_PM_GENERAL_TAYLOR && d->max_order > FPA_MAX_TAYLOR_ORDER) {
        set_error("max_order %d exceeds the supported Taylor order %d", d->max_order, FPA_MAX_TAYLOR_ORDER);
        return FPA_ERR_UNSUPPORTED;
    }
    for (int e = 0; e < d->n_even; ++e) {
        if (d->method == FPA_PM_SYMMETRIC_EVEN && d->even_orders[e] > FPA_MAX_TAYLOR_ORDER) {
            set_error("even order %d exceeds the supported Taylor order %d", d->even_orders[e],
                      FPA_MAX_TAYLOR_ORDER);
            return FPA_ERR_UNSUPPORTED;
        }
    }
    const int64_t B = d->n1 * d->n3;
    if (B == 0) return FPA_OK;

    PlanParams p;
    p.n1 = d->n1;
    p.n3 = d->n3;
    p.lambda1 = d->lambda1;
    p.lambda2 = d->lambda2;
    p.lambda3 = d->lambda3;
    p.lambda2_stride = (int)d->lambda2_stride;
    p.method = d->method;
    p.max_order = d->max_order;
    p.n_even = d->n_even;
    for (int i = 0; i < FPA_MAX_TAYLOR_ORDER; ++i) p.even_orders[i] = d->even_orders[i];
    for (int i = 0; i <= FPA_MAX_TAYLOR_ORDER; ++i) p.beta[i] = d->beta[i];
    p.omega_ref = d->omega_ref;
    p.atol = d->atol;
    p.rtol = d->rtol;
    p.provided = d->provided;
    p.omega = d->omega;
    p.dbeta = d->dbeta;
    p.dbeta_masked = dbeta_masked;
    p.valid = d->valid;

    const int threads = 256;
    plan_dbeta_kernel<<<(unsigned)((B + threads - 1) / threads), threads, 0, st>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "plan_dbeta_kernel launch");
    return FPA_OK;
}

// ------------------------------------------------------------------ sweep glue kernels
// Small per-sweep constants go to device memory through kernel arguments (fully asynchronous, no
// staging copy): consts = [gamma, alpha, A0(8 doubles)].
__global__ void sweep_consts_kernel(double gamma, double alpha, double a0, double a1, double a2,
                                    double a3, double a4, double a5, double a6, double a7,
                                    double* consts) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    consts[0] = gamma;
    consts[1] = alpha;
    consts[2] = a0; consts[3] = a1; consts[4] = a2; consts[5] = a3;
    consts[6] = a4; consts[7] = a5; consts[8] = a6; consts[9] = a7;
}

// gain = Pmax_signal / p_in[2]; NaN for invalid, failed, non-finite or <= 0 (scan_mismtach.py:723-738)
__global__ void sweep_gain_kernel(int64_t B, const double* Pmax, const int32_t* valid,
                                  const int32_t* status, double p_signal, int check_nan,
                                  double* gain_lin) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double g = qnan();
    if (valid[b] && !(check_nan && status[b] != FPA_POINT_OK)) {
        const double P3 = Pmax[b * 4 + 2];
        if (!nonfinite(P3)) {
            const double q = P3 / p_signal;
            if (!nonfinite(q) && q > 0.0) g = q;
        }
    }
    gain_lin[b] = g;
}

int sweep_consts_launch(double gamma, double alpha, const double* A0, double* consts, cudaStream_t st) {
    sweep_consts_kernel<<<1, 32, 0, st>>>(gamma, alpha, A0[0], A0[1], A0[2], A0[3], A0[4], A0[5], A0[6],
                                          A0[7], consts);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "sweep_consts_kernel launch");
    return FPA_OK;
}

int sweep_gain_launch(int64_t B, const double* Pmax, const int32_t* valid, const int32_t* status,
                      double p_signal, int check_nan, double* gain_lin, cudaStream_t st) {
    if (B == 0) return FPA_OK;
    const int threads = 256;
    sweep_gain_kernel<<<(unsigned)((B + threads - 1) / threads), threads, 0, st>>>(
        B, Pmax, valid, status, p_signal, check_nan, gain_lin);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "sweep_gain_kernel launch");
    return FPA_OK;
}

}  // namespace fpa
