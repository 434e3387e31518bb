// nwave_comb.cu -- N-wave RK4 for plans on an integer frequency grid, in convolution form
// (NOT in the reference; SURVEY App. C "uniform-comb shortcut").
//
// On a grid w_j = w_0 + g_j*dw the matching condition w_k + w_l - w_m = w_n is g_k + g_l - g_m = g_n,
// so with At_j = A_j exp(i*beta_j*z) placed at grid slot g_j - g_min (empty slots hold 0):
//
//     sum over ordered (k,l) and all m of At_k At_l conj(At_m)  =  sum_m conj(At_m) * C_{n+m},
//     C_s = sum_{k+l=s} At_k At_l                                   (auto-convolution)
//
// which is O(M^2) per RHS (M = grid span) instead of O(N^3) table entries, and contains the SPM/XPM
// terms (m = k or m = l) with exactly the weights of the analytic Kerr factor (2*sum P - P_n), so
//
//     dA_n/dz = -(alpha/2) A_n + i*gamma * conj(E_n) * sum_m conj(At_m) C_{n+m} .
//
// It is the same ODE as nwave.cu integrates from the enumerated triplet table (tests compare both
// with the oracle); N = 64 needs 2*64^2 complex MACs per RHS instead of 84 320 table entries.
//
// Mapping: one CTA per scan point, everything in shared memory for all z-steps; thread s owns C_s
// (the k-sum uses the k<->l symmetry: half the terms, doubled), thread j owns wave j's correlation
// sum and its RK4 update.  Neighbouring threads read neighbouring shared-memory words (C_{n+m},
// At_{s-k}) and the common operand is a broadcast.
#include "fpa_common.cuh"

namespace fpa {

struct CombParams {
    int64_t        n_points;
    int            n_waves, span;  // N lines on M = span grid slots
    int            beta_stride, gamma_stride, alpha_stride, A0_stride;
    const double*  beta;
    const double*  gamma;
    const double*  alpha;
    const double*  A0;
    const int32_t* slot;  // [N] grid slot of each wave (g_j - g_min), device memory
    double         z0, z_max;
    int            n_steps, save_every;
    int64_t        n_saved;
    double*        A_trace;
    double*        A_end;
    double*        Pmax;
    int32_t*       status;
    int            check;
};

struct CombSmem {
    double2 *y, *ys, *yn, *E, *At, *C;
    double*  beta;
    int*     slot;
    int*     flag;
};

__device__ __forceinline__ CombSmem comb_carve(double* base, int N, int M) {
    CombSmem s;
    s.y    = reinterpret_cast<double2*>(base);
    s.ys   = s.y + N;
    s.yn   = s.ys + N;
    s.E    = s.yn + N;
    s.At   = s.E + N;
    s.C    = s.At + M;
    s.beta = reinterpret_cast<double*>(s.C + (2 * M - 1));
    s.slot = reinterpret_cast<int*>(s.beta + N);
    s.flag = s.slot + N;
    return s;
}

static size_t comb_smem_bytes(int N, int M) {
    return sizeof(double2) * (size_t)(4 * N + M + 2 * M - 1) + sizeof(double) * (size_t)N +
           sizeof(int) * (size_t)(N + 2);
}

// One RHS evaluation at z on the stage state s.ys, then for every wave the RK4 bookkeeping:
//   yn += wa*k ; ys = y + wb*k            (or y = yn + wa*k when `last`)
__device__ void comb_stage(const CombSmem& s, const CombParams& p, double z, double gamma, double nha,
                           double wa, double wb, bool last, bool first_of_run_check, int& nonfinite_seen) {
    const int N = p.n_waves, M = p.span, tid = threadIdx.x, nt = blockDim.x;

    for (int j = tid; j < N; j += nt) {
        double sn, cs;
        sincos(s.beta[j] * z, &sn, &cs);
        const double2 a = s.ys[j];
        s.E[j] = make_double2(cs, sn);
        s.At[s.slot[j]] = make_double2(fma(-a.y, sn, a.x * cs), fma(a.x, sn, a.y * cs));
        if (first_of_run_check && (nonfinite(a.x) || nonfinite(a.y))) nonfinite_seen = 1;
    }
    __syncthreads();

    // C_s = sum_{k+l=s} At_k At_l = 2*sum_{k<l} + [s even] At_{s/2}^2
    for (int sidx = tid; sidx < 2 * M - 1; sidx += nt) {
        const int klo = sidx - (M - 1) > 0 ? sidx - (M - 1) : 0;
        const int khi = (sidx - 1) >> 1;  // largest k with k < s-k
        double cr = 0.0, ci = 0.0;
        for (int k = klo; k <= khi; ++k) {
            const double2 a = s.At[k], b = s.At[sidx - k];
            cr = fma(a.x, b.x, fma(-a.y, b.y, cr));
            ci = fma(a.x, b.y, fma(a.y, b.x, ci));
        }
        cr += cr;
        ci += ci;
        if ((sidx & 1) == 0) {
            const double2 a = s.At[sidx >> 1];
            cr = fma(a.x, a.x, fma(-a.y, a.y, cr));
            ci = fma(a.x + a.x, a.y, ci);
        }
        s.C[sidx] = make_double2(cr, ci);
    }
    __syncthreads();

    for (int j = tid; j < N; j += nt) {
        const int n = s.slot[j];
        double rr = 0.0, ri = 0.0;
        for (int m = 0; m < M; ++m) {
            const double2 a = s.At[m], c = s.C[n + m];
            // conj(a) * c
            rr = fma(a.x, c.x, fma(a.y, c.y, rr));
            ri = fma(a.x, c.y, fma(-a.y, c.x, ri));
        }
        const double2 e = s.E[j], x = s.ys[j];
        // F = conj(E_n) * R ;  k = nha*x + i*gamma*F
        const double fr = fma(ri, e.y, rr * e.x);
        const double fi = fma(ri, e.x, -(rr * e.y));
        const double kr = fma(nha, x.x, -(gamma * fi));
        const double ki = fma(nha, x.y, gamma * fr);
        if (last) {
            const double2 acc = s.yn[j];
            s.y[j] = make_double2(fma(wa, kr, acc.x), fma(wa, ki, acc.y));
        } else {
            const double2 y0 = s.y[j], acc = s.yn[j];
            s.yn[j] = make_double2(fma(wa, kr, acc.x), fma(wa, ki, acc.y));
            s.ys[j] = make_double2(fma(wb, kr, y0.x), fma(wb, ki, y0.y));
        }
    }
    __syncthreads();
}

__global__ void nwave_comb_kernel(const CombParams p) {
    extern __shared__ double comb_smem_raw[];
    const int     N = p.n_waves, M = p.span, tid = threadIdx.x, nt = blockDim.x;
    const int64_t b = blockIdx.x;
    CombSmem      s = comb_carve(comb_smem_raw, N, M);

    const double gamma = p.gamma[b * p.gamma_stride];
    const double nha   = -0.5 * p.alpha[b * p.alpha_stride];
    for (int j = tid; j < N; j += nt) {
        s.beta[j] = p.beta[b * p.beta_stride * N + j];
        s.slot[j] = p.slot[j];
        const double2* a0 = reinterpret_cast<const double2*>(p.A0) + b * p.A0_stride * N;
        s.y[j] = a0[j];
    }
    for (int m = tid; m < M; m += nt) s.At[m] = make_double2(0.0, 0.0);  // empty grid slots stay 0
    __syncthreads();

    double2* tr = p.A_trace ? reinterpret_cast<double2*>(p.A_trace) + b * p.n_saved * N : nullptr;
    if (tr) {
        for (int j = tid; j < N; j += nt) tr[j] = s.y[j];
        tr += N;
    }
    double pm[4] = {0.0, 0.0, 0.0, 0.0};  // waves tid, tid+nt, ... (N <= 128, nt >= 32)
    if (p.Pmax) {
        int q = 0;
        for (int j = tid; j < N; j += nt, ++q) pm[q] = fma(s.y[j].y, s.y[j].y, s.y[j].x * s.y[j].x);
    }

    const int    n_steps = p.n_steps;
    const double z0 = p.z0, z_max = p.z_max;
    const double step = (z_max - z0) / (double)n_steps;  // numpy.linspace arithmetic
    double       zi = z0, di = 0.0;
    int          save_ctr = p.save_every;
    int32_t      bad = FPA_POINT_OK;

    for (int i = 0; i < n_steps; ++i) {
        di += 1.0;
        const double zn = (i + 1 == n_steps) ? z_max : __dadd_rn(__dmul_rn(di, step), z0);
        const double h = zn - zi, hh = 0.5 * h, h6 = h / 6.0, h3 = h6 + h6;
        for (int j = tid; j < N; j += nt) {
            const double2 v = s.y[j];
            s.ys[j] = v;
            s.yn[j] = v;
        }
        __syncthreads();
        int nf = 0;
        comb_stage(s, p, zi, gamma, nha, h6, hh, false, p.check && i > 0 && bad == FPA_POINT_OK, nf);
        if (p.check && i > 0 && bad == FPA_POINT_OK) {
            if (__syncthreads_or(nf)) bad = i - 1;  // the state produced by step i-1 was not finite
        }
        comb_stage(s, p, zi + hh, gamma, nha, h3, hh, false, false, nf);
        comb_stage(s, p, zi + hh, gamma, nha, h3, h, false, false, nf);
        comb_stage(s, p, zi + h, gamma, nha, h6, 0.0, true, false, nf);
        zi = zn;

        if (--save_ctr == 0) {
            save_ctr = p.save_every;
            if (tr) {
                for (int j = tid; j < N; j += nt) tr[j] = s.y[j];
                tr += N;
            }
            if (p.Pmax) {
                int q = 0;
                for (int j = tid; j < N; j += nt, ++q) {
                    const double P = fma(s.y[j].y, s.y[j].y, s.y[j].x * s.y[j].x);
                    pm[q] = (P != P || pm[q] != pm[q]) ? qnan() : fmax(pm[q], P);
                }
            }
        }
    }
    if (p.check && bad == FPA_POINT_OK) {
        int nf = 0;
        for (int j = tid; j < N; j += nt) nf |= (nonfinite(s.y[j].x) || nonfinite(s.y[j].y)) ? 1 : 0;
        if (__syncthreads_or(nf)) bad = n_steps - 1;
    }
    if (p.status && tid == 0) p.status[b] = bad;
    if (p.A_end) {
        double2* o = reinterpret_cast<double2*>(p.A_end) + b * N;
        for (int j = tid; j < N; j += nt) o[j] = s.y[j];
    }
    if (p.Pmax) {
        int q = 0;
        for (int j = tid; j < N; j += nt, ++q) p.Pmax[b * N + j] = pm[q];
    }
}

// d->grid_slot: device pointer to the N grid slots; span = number of grid slots covered.
int nwave_comb_launch(const fpa_nwave_desc* d, cudaStream_t st) {
    const int N = d->n_waves, M = d->grid_span;
    CombParams p;
    p.n_points     = d->n_points;
    p.n_waves      = N;
    p.span         = M;
    p.beta_stride  = (int)d->beta_stride;
    p.gamma_stride = (int)d->gamma_stride;
    p.alpha_stride = (int)d->alpha_stride;
    p.A0_stride    = (int)d->A0_stride;
    p.beta         = d->beta;
    p.gamma        = d->gamma;
    p.alpha        = d->alpha;
    p.A0           = d->A0;
    p.slot         = d->grid_slot;
    p.z0           = d->z0;
    p.z_max        = d->z_max;
    p.n_steps      = (int)d->n_steps;
    p.save_every   = (int)(d->save_every > d->n_steps ? d->n_steps + 1 : d->save_every);
    p.n_saved      = fpa_n_saved(d->n_steps, d->save_every);
    p.A_trace      = (d->flags & FPA_OUT_TRACE) ? d->A_trace : nullptr;
    p.A_end        = (d->flags & FPA_OUT_END) ? d->A_end : nullptr;
    p.Pmax         = (d->flags & FPA_OUT_PMAX) ? d->Pmax : nullptr;
    p.status       = d->status;
    p.check        = (d->flags & FPA_CHECK_NAN) ? 1 : 0;

    const size_t smem = comb_smem_bytes(N, M);
    int threads = ((2 * M - 1) + 31) / 32 * 32;  // one thread per C_s
    if (threads > 1024) threads = 1024;
    cudaError_t e = cudaFuncSetAttribute(nwave_comb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(nwave_comb_kernel)");
    nwave_comb_kernel<<<(unsigned)d->n_points, threads, smem, st>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "nwave_comb_kernel launch");
    return FPA_OK;
}

}  // namespace fpa

extern "C" double fpa_nwave_comb_flops_per_step(int32_t n_waves, int32_t grid_span) {
    // per RHS: C (half the ordered pairs, complex MAC = 8 flops) 4*M^2, correlation 8*N*M, and 30*N for
    // phases, rotation, conj(E)*R and the assembly; per step 4 RHS + 26*N for the RK4 combination.
    const double M = grid_span, N = n_waves;
    return 4.0 * (4.0 * M * M + 8.0 * N * M + 30.0 * N) + 26.0 * N;
}
