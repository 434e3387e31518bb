// yaman4.cu -- fused fixed-step RK4 x 4-wave Yaman/Agrawal FWM right-hand side.
//
// One thread integrates one scan point for ALL z-steps with its state in registers:
// the four RK4 stages, the RHS, the per-step finite check, the max-over-saved-samples
// metric and the trace write-out are a single kernel (one launch per batch).
//
// What it stands in for in the reference (paths relative to the reference checkout):
//   integrators.rk4_step              integrators.py:25-61   (stage formulas :54-59)
//   integrators.integrate_fixed_step  integrators.py:68-142  (grid, saving, NaN guard)
//   integrators.integrate_interval    integrators.py:150-204 (linspace grid)
//   yaman_model.rhs_yaman_simplified  yaman_model.py:10-52   (+ _linear_loss_terms :123-132,
//                                     _kerr_terms :135-156, _fwm_terms :159-186)
//
// Why thread-per-point and not warp-per-point for N = 4: the FWM sum has four terms;
// a warp per point would leave 28 lanes idle.  The bound is the FP64 FMA pipe, so the
// kernel is written to minimise DFMA-pipe instructions per step:
//   RHS      64 FP64 instructions (8 powers, 3 adds, 1+4 Kerr factors, 8 pair products,
//            8 phased pairs, 32 fused assemble ops)
//   RK4      56 FMAs per step (3 stage states + 4 accumulations of 8 components)
//   phase    2 complex rotations per step (8 instr.) with an exact sincos re-sync every
//            FPA_RESYNC steps, instead of the reference's 8 complex exp() per step
//   => ~325 FP64 instructions for 568 algorithmic flops per point.step.
#include "fpa_common.cuh"

namespace fpa {

constexpr int kResync = 32;  // steps between exact sincos re-synchronisations of the phase

struct Yaman4Params {
    int64_t       n_points;
    const double* dbeta;
    const double* gamma;
    const double* alpha;
    const double* A0;
    const double* z_grid;
    double*       A_trace;
    double*       A_end;
    double*       Pmax;
    int32_t*      status;
    double        z0, z_max;
    int           gamma_stride, alpha_stride, A0_stride;
    int           n_steps, save_every;
    int64_t       n_saved;
};

// dA/dz for one point.  y = (x1,y1,x2,y2,x3,y3,x4,y4); (pr,pi) = 2*gamma*exp(i*dbeta*z);
// nha = -alpha/2.  64 FP64 instructions, negations ride on the FMA source modifiers.
__device__ __forceinline__ void rhs4(const double (&y)[8], double pr, double pi, double gamma,
                                     double nha, double (&k)[8], double& Ssum) {
    const double x1 = y[0], y1 = y[1], x2 = y[2], y2 = y[3];
    const double x3 = y[4], y3 = y[5], x4 = y[6], y4 = y[7];

    // powers (yaman_model.py:144-147)
    const double P1 = fma(y1, y1, x1 * x1);
    const double P2 = fma(y2, y2, x2 * x2);
    const double P3 = fma(y3, y3, x3 * x3);
    const double P4 = fma(y4, y4, x4 * x4);
    const double S  = (P1 + P2) + (P3 + P4);
    Ssum = S;
    // gamma*(P_j + 2*sum_{k!=j} P_k) == gamma*(2S - P_j)   (yaman_model.py:148-151)
    const double c2 = (gamma + gamma) * S;
    const double G1 = fma(-gamma, P1, c2);
    const double G2 = fma(-gamma, P2, c2);
    const double G3 = fma(-gamma, P3, c2);
    const double G4 = fma(-gamma, P4, c2);

    // pair products shared by two waves each (yaman_model.py:177-181)
    const double Ur = fma(-y3, y4, x3 * x4), Ui = fma(x3, y4, y3 * x4);  // A3*A4
    const double Vr = fma(-y1, y2, x1 * x2), Vi = fma(x1, y2, y1 * x2);  // A1*A2
    // W = 2g e^{+i th} U ,  Z = 2g e^{-i th} V
    const double Wr = fma(-pi, Ui, pr * Ur), Wi = fma(pr, Ui, pi * Ur);
    const double Zr = fma(pi, Vi, pr * Vr),  Zi = fma(pr, Vi, -(pi * Vr));

    // dA_j = nha*A_j + i*G_j*A_j + i*conj(A_m)*{W|Z}
    //   conj(A_m)*W = (xm Wr + ym Wi) + i (xm Wi - ym Wr)
    k[0] = fma(nha, x1, fma(-G1, y1, fma(y2, Wr, -(x2 * Wi))));
    k[1] = fma(nha, y1, fma(G1, x1, fma(x2, Wr, y2 * Wi)));
    k[2] = fma(nha, x2, fma(-G2, y2, fma(y1, Wr, -(x1 * Wi))));
    k[3] = fma(nha, y2, fma(G2, x2, fma(x1, Wr, y1 * Wi)));
    k[4] = fma(nha, x3, fma(-G3, y3, fma(y4, Zr, -(x4 * Zi))));
    k[5] = fma(nha, y3, fma(G3, x3, fma(x4, Zr, y4 * Zi)));
    k[6] = fma(nha, x4, fma(-G4, y4, fma(y3, Zr, -(x3 * Zi))));
    k[7] = fma(nha, y4, fma(G4, x4, fma(x3, Zr, y3 * Zi)));
}

enum PhaseMode { kRecurrence = 0, kExactUniform = 1, kExplicitGrid = 2 };

template <bool TRACE, bool PMAX, bool CHECK, int PHASE>
__global__ void __launch_bounds__(128, 4) yaman4_rk4_kernel(const Yaman4Params p) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.n_points) return;

    const double dbeta = p.dbeta[b];
    const double gamma = p.gamma[b * p.gamma_stride];
    const double nha   = -0.5 * p.alpha[b * p.alpha_stride];
    const double g2    = gamma + gamma;

    double y[8];
    {
        const double2* a0 = reinterpret_cast<const double2*>(p.A0) + b * p.A0_stride * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double2 v = a0[j];
            y[2 * j] = v.x;
            y[2 * j + 1] = v.y;
        }
    }

    if (nonfinite(dbeta)) {
        // exp(i*dbeta*z) is NaN from the first stage on: every later sample is NaN and the first
        // step is the bad one (integrators.py:132-135).  Invalid scan points of a sweep arrive
        // here with dbeta = NaN and cost nothing.
        const double qn = qnan();
        if (TRACE) {
            double* t = p.A_trace + b * p.n_saved * 8;
#pragma unroll
            for (int j = 0; j < 4; ++j) store_c128(t + 2 * j, y[2 * j], y[2 * j + 1]);
            for (int64_t s = 1; s < p.n_saved; ++s)
#pragma unroll
                for (int j = 0; j < 4; ++j) store_c128(t + s * 8 + 2 * j, qn, qn);
        }
        if (p.A_end)
#pragma unroll
            for (int j = 0; j < 4; ++j) store_c128(p.A_end + b * 8 + 2 * j, qn, qn);
        if (PMAX) {
            double2* o = reinterpret_cast<double2*>(p.Pmax + b * 4);
            o[0] = make_double2(qn, qn);
            o[1] = make_double2(qn, qn);
        }
        if (p.status) p.status[b] = CHECK ? 0 : FPA_POINT_OK;
        return;
    }

    double* tr = nullptr;
    if (TRACE) {
        tr = p.A_trace + b * p.n_saved * 8;
#pragma unroll
        for (int j = 0; j < 4; ++j) store_c128(tr + 2 * j, y[2 * j], y[2 * j + 1]);
        tr += 8;
    }
    double pm[4];
    if (PMAX) {
#pragma unroll
        for (int j = 0; j < 4; ++j) pm[j] = fma(y[2 * j + 1], y[2 * j + 1], y[2 * j] * y[2 * j]);
    }

    const int    n_steps = p.n_steps;
    const double z0 = p.z0, z_max = p.z_max;
    // numpy.linspace: step = (stop-start)/div, z_i = i*step + start, z_n = stop
    const double step = (z_max - z0) / (double)n_steps;

    // rotation by half a nominal step
    double rr, ri;
    if (PHASE == kRecurrence) sincos(dbeta * (0.5 * step), &ri, &rr);

    double  zi = (PHASE == kExplicitGrid) ? p.z_grid[0] : z0;
    double  di = 0.0;
    double  pr = g2, pi = 0.0;  // 2*gamma*exp(i*dbeta*z_i)
    int     save_ctr = p.save_every;
    int32_t bad = FPA_POINT_OK;

    for (int i = 0; i < n_steps; ++i) {
        double zn;
        if (PHASE == kExplicitGrid) {
            zn = p.z_grid[i + 1];
        } else {
            di += 1.0;
            zn = (i + 1 == n_steps) ? z_max : __dadd_rn(__dmul_rn(di, step), z0);
        }
        const double h  = zn - zi;  // integrators.py:128
        const double hh = 0.5 * h;
        const double h6 = h * (1.0 / 6.0);  // integrators.py:59 (h/6: a 1-ulp difference, no FP64 divide in the loop)
        const double h3 = h6 + h6;

        double phr, phi_, p1r, p1i;  // phase at z+h/2 and z+h
        if (PHASE == kRecurrence) {
            if ((i & (kResync - 1)) == 0) {
                double s, c;
                sincos(dbeta * zi, &s, &c);
                pr = g2 * c;
                pi = g2 * s;
            }
            phr  = fma(-pi, ri, pr * rr);
            phi_ = fma(pr, ri, pi * rr);
            p1r  = fma(-phi_, ri, phr * rr);
            p1i  = fma(phr, ri, phi_ * rr);
        } else {
            double s, c;
            sincos(dbeta * zi, &s, &c);
            pr = g2 * c;
            pi = g2 * s;
            sincos(dbeta * (zi + hh), &s, &c);  // integrators.py:55-56
            phr  = g2 * c;
            phi_ = g2 * s;
            sincos(dbeta * (zi + h), &s, &c);  // integrators.py:57
            p1r = g2 * c;
            p1i = g2 * s;
        }

        double k[8], ys[8], yn[8], S, Sx;
        rhs4(y, pr, pi, gamma, nha, k, S);
        if (CHECK) {
            // S = sum |A|^2 of the state that step i-1 produced: non-finite S is the only way
            // a component can be non-finite, so the exact per-component test runs only then.
            if (nonfinite(S) && bad == FPA_POINT_OK && i > 0) {
                bool nf = false;
#pragma unroll
                for (int j = 0; j < 8; ++j) nf |= nonfinite(y[j]);
                if (nf) bad = i - 1;
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            yn[j] = fma(h6, k[j], y[j]);
            ys[j] = fma(hh, k[j], y[j]);
        }
        rhs4(ys, phr, phi_, gamma, nha, k, Sx);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            yn[j] = fma(h3, k[j], yn[j]);
            ys[j] = fma(hh, k[j], y[j]);
        }
        rhs4(ys, phr, phi_, gamma, nha, k, Sx);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            yn[j] = fma(h3, k[j], yn[j]);
            ys[j] = fma(h, k[j], y[j]);
        }
        rhs4(ys, p1r, p1i, gamma, nha, k, Sx);
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = fma(h6, k[j], yn[j]);

        pr = p1r;
        pi = p1i;
        zi = zn;

        if (--save_ctr == 0) {  // (i+1) % save_every == 0, integrators.py:137-140
            save_ctr = p.save_every;
            if (TRACE) {
#pragma unroll
                for (int j = 0; j < 4; ++j) store_c128(tr + 2 * j, y[2 * j], y[2 * j + 1]);
                tr += 8;
            }
            if (PMAX) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const double P = fma(y[2 * j + 1], y[2 * j + 1], y[2 * j] * y[2 * j]);
                    // numpy.max semantics: NaN is sticky
                    pm[j] = (P != P || pm[j] != pm[j]) ? qnan() : fmax(pm[j], P);
                }
            }
        }
    }

    if (CHECK && bad == FPA_POINT_OK) {
        bool nf = false;
#pragma unroll
        for (int j = 0; j < 8; ++j) nf |= nonfinite(y[j]);
        if (nf) bad = n_steps - 1;
    }
    if (p.status) p.status[b] = bad;
    if (p.A_end) {
#pragma unroll
        for (int j = 0; j < 4; ++j) store_c128(p.A_end + b * 8 + 2 * j, y[2 * j], y[2 * j + 1]);
    }
    if (PMAX) {
        double2* o = reinterpret_cast<double2*>(p.Pmax + b * 4);
        o[0] = make_double2(pm[0], pm[1]);
        o[1] = make_double2(pm[2], pm[3]);
    }
}

// RHS-only kernel (direct calls of yaman_model.rhs_yaman_simplified, yaman_model.py:10-52):
// the phase is evaluated exactly at the caller's z.
__global__ void yaman4_rhs_kernel(int64_t B, const double* z, const double* A, const double* gamma,
                                  const double* alpha, const double* dbeta, double* dA) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double y[8], k[8], S, s, c;
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] = A[b * 8 + j];
    const double g = gamma[b];
    sincos(dbeta[b] * z[b], &s, &c);
    rhs4(y, (g + g) * c, (g + g) * s, g, -0.5 * alpha[b], k, S);
#pragma unroll
    for (int j = 0; j < 8; ++j) dA[b * 8 + j] = k[j];
}

template <bool TRACE, bool PMAX, bool CHECK>
static cudaError_t launch_phase(const Yaman4Params& p, int phase, cudaStream_t st) {
    const int  threads = 128;
    const long blocks  = (long)((p.n_points + threads - 1) / threads);
    switch (phase) {
        case kRecurrence:
            yaman4_rk4_kernel<TRACE, PMAX, CHECK, kRecurrence><<<blocks, threads, 0, st>>>(p);
            break;
        case kExactUniform:
            yaman4_rk4_kernel<TRACE, PMAX, CHECK, kExactUniform><<<blocks, threads, 0, st>>>(p);
            break;
        default:
            yaman4_rk4_kernel<TRACE, PMAX, CHECK, kExplicitGrid><<<blocks, threads, 0, st>>>(p);
            break;
    }
    return cudaGetLastError();
}

int yaman4_launch(const fpa_yaman4_desc* d, cudaStream_t st) {
    FPA_REQUIRE(d != nullptr, "descriptor is NULL");
    FPA_REQUIRE(d->n_points >= 0, "n_points must be >= 0");
    FPA_REQUIRE(d->n_steps >= 1 && d->n_steps < 2147483647LL, "n_steps must be in [1, 2^31)");
    FPA_REQUIRE(d->save_every >= 1, "save_every must be a positive integer");
    FPA_REQUIRE(d->dbeta && d->gamma && d->alpha && d->A0, "dbeta/gamma/alpha/A0 must be set");
    FPA_REQUIRE((d->gamma_stride | 1) == 1 && (d->alpha_stride | 1) == 1 && (d->A0_stride | 1) == 1,
                "strides must be 0 (broadcast) or 1 (per point)");
    const bool trace = (d->flags & FPA_OUT_TRACE) != 0;
    const bool pmax  = (d->flags & FPA_OUT_PMAX) != 0;
    const bool endo  = (d->flags & FPA_OUT_END) != 0;
    FPA_REQUIRE(!trace || d->A_trace, "FPA_OUT_TRACE needs A_trace");
    FPA_REQUIRE(!pmax || d->Pmax, "FPA_OUT_PMAX needs Pmax");
    FPA_REQUIRE(!endo || d->A_end, "FPA_OUT_END needs A_end");
    if (d->n_points == 0) return FPA_OK;

    Yaman4Params p;
    p.n_points     = d->n_points;
    p.dbeta        = d->dbeta;
    p.gamma        = d->gamma;
    p.alpha        = d->alpha;
    p.A0           = d->A0;
    p.z_grid       = d->z_grid;
    p.A_trace      = trace ? d->A_trace : nullptr;
    p.A_end        = endo ? d->A_end : nullptr;
    p.Pmax         = pmax ? d->Pmax : nullptr;
    p.status       = d->status;
    p.z0           = d->z0;
    p.z_max        = d->z_max;
    p.gamma_stride = (int)d->gamma_stride;
    p.alpha_stride = (int)d->alpha_stride;
    p.A0_stride    = (int)d->A0_stride;
    p.n_steps      = (int)d->n_steps;
    // save_every > n_steps never fires; clamp so the countdown fits an int
    p.save_every   = (int)(d->save_every > d->n_steps ? d->n_steps + 1 : d->save_every);
    p.n_saved      = fpa_n_saved(d->n_steps, d->save_every);

    const int phase = d->z_grid ? kExplicitGrid
                                : ((d->flags & FPA_PHASE_EXACT) ? kExactUniform : kRecurrence);
    const bool check = (d->flags & FPA_CHECK_NAN) != 0;
    cudaError_t e;
#define FPA_DISPATCH(T, M)                                             \
    (check ? launch_phase<T, M, true>(p, phase, st) : launch_phase<T, M, false>(p, phase, st))
    if (trace && pmax)  e = FPA_DISPATCH(true, true);
    else if (trace)     e = FPA_DISPATCH(true, false);
    else if (pmax)      e = FPA_DISPATCH(false, true);
    else                e = FPA_DISPATCH(false, false);
#undef FPA_DISPATCH
    if (e != cudaSuccess) return cuda_fail(e, "yaman4_rk4_kernel launch");
    return FPA_OK;
}

int yaman4_rhs_launch(int64_t B, const double* z, const double* A, const double* gamma,
                      const double* alpha, const double* dbeta, double* dA, cudaStream_t st) {
    if (B == 0) return FPA_OK;
    const int threads = 128;
    yaman4_rhs_kernel<<<(unsigned)((B + threads - 1) / threads), threads, 0, st>>>(B, z, A, gamma, alpha,
                                                                                 dbeta, dA);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "yaman4_rhs_kernel launch");
    return FPA_OK;
}

}  // namespace fpa
