// linear.cu -- the integrator on the built-in linear test system  y_j' = lambda_j * y_j.
//
// The reference's own integrator tests (tests.py:146-226) drive integrators.rk4_step /
// integrate_fixed_step / integrate_interval with the Python callable y' = y on a real state of
// dimension 1.  A device integrator cannot call back into Python, so the host mirror registers this
// RHS kind and the same assertions are replayed on the GPU through this kernel.  Grid, saving and
// finite-check semantics are those of integrators.py:68-142 (see yaman4.cu).
#include "fpa_common.cuh"

namespace fpa {

struct LinearParams {
    int64_t       n_threads;  // B * dim
    int           dim;
    const double* y0;      // [B,dim] complex128
    const double* lam;     // [dim] complex128
    const double* z_grid;  // NULL or [n_steps+1]
    double        z0, z_max;
    int           n_steps, save_every;
    int64_t       n_saved;
    double*       y_trace;  // [B,n_saved,dim] complex128 or NULL
    double*       y_end;    // [B,dim] or NULL
    int32_t*      bad;      // [B*dim] scratch: first bad step per component (or -1)
    int           check;
};

__global__ void linear_rk4_kernel(const LinearParams p) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.n_threads) return;
    const int64_t b = t / p.dim;
    const int     j = (int)(t - b * p.dim);
    const double  lr = p.lam[2 * j], li = p.lam[2 * j + 1];
    double        yr = p.y0[2 * t], yi = p.y0[2 * t + 1];

    double* tr = p.y_trace ? p.y_trace + (b * p.n_saved * p.dim + j) * 2 : nullptr;
    if (tr) {
        store_c128(tr, yr, yi);
        tr += 2 * p.dim;
    }
    const double step = (p.z_max - p.z0) / (double)p.n_steps;
    double       zi = p.z_grid ? p.z_grid[0] : p.z0, di = 0.0;
    int          save_ctr = p.save_every;
    int32_t      bad = FPA_POINT_OK;

    for (int i = 0; i < p.n_steps; ++i) {
        double zn;
        if (p.z_grid) {
            zn = p.z_grid[i + 1];
        } else {
            di += 1.0;
            zn = (i + 1 == p.n_steps) ? p.z_max : __dadd_rn(__dmul_rn(di, step), p.z0);
        }
        const double h = zn - zi, hh = 0.5 * h, h6 = h / 6.0;
        // k = lam * y   (complex)
        const double k1r = fma(-li, yi, lr * yr), k1i = fma(lr, yi, li * yr);
        double       sr = fma(hh, k1r, yr), si = fma(hh, k1i, yi);
        const double k2r = fma(-li, si, lr * sr), k2i = fma(lr, si, li * sr);
        sr = fma(hh, k2r, yr);
        si = fma(hh, k2i, yi);
        const double k3r = fma(-li, si, lr * sr), k3i = fma(lr, si, li * sr);
        sr = fma(h, k3r, yr);
        si = fma(h, k3i, yi);
        const double k4r = fma(-li, si, lr * sr), k4i = fma(lr, si, li * sr);
        // y + (h/6)*(k1 + 2k2 + 2k3 + k4)   (integrators.py:59)
        yr = fma(h6, (k1r + 2.0 * k2r) + (2.0 * k3r + k4r), yr);
        yi = fma(h6, (k1i + 2.0 * k2i) + (2.0 * k3i + k4i), yi);
        zi = zn;
        if (p.check && bad == FPA_POINT_OK && (nonfinite(yr) || nonfinite(yi))) bad = i;
        if (--save_ctr == 0) {
            save_ctr = p.save_every;
            if (tr) {
                store_c128(tr, yr, yi);
                tr += 2 * p.dim;
            }
        }
    }
    if (p.y_end) store_c128(p.y_end + 2 * t, yr, yi);
    if (p.bad) p.bad[t] = bad;
}

// status[b] = min over components of the first bad step (or -1)
__global__ void linear_status_kernel(int64_t B, int dim, const int32_t* bad, int32_t* status) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    int32_t s = FPA_POINT_OK;
    for (int j = 0; j < dim; ++j) {
        const int32_t v = bad[b * dim + j];
        if (v != FPA_POINT_OK && (s == FPA_POINT_OK || v < s)) s = v;
    }
    status[b] = s;
}

int linear_launch(int64_t B, int dim, const double* y0, const double* lam, double z0, double z_max,
                  int64_t n_steps, int64_t save_every, const double* z_grid, uint32_t flags,
                  double* y_trace, double* y_end, int32_t* bad_scratch, int32_t* status,
                  cudaStream_t st) {
    FPA_REQUIRE(B >= 0 && dim >= 1, "need B >= 0 and dim >= 1");
    FPA_REQUIRE(n_steps >= 1 && n_steps < 2147483647LL, "n_steps must be in [1, 2^31)");
    FPA_REQUIRE(save_every >= 1, "save_every must be a positive integer");
    FPA_REQUIRE(y0 && lam, "y0 and lam must be set");
    if (B == 0) return FPA_OK;
    LinearParams p;
    p.n_threads  = B * dim;
    p.dim        = dim;
    p.y0         = y0;
    p.lam        = lam;
    p.z_grid     = z_grid;
    p.z0         = z0;
    p.z_max      = z_max;
    p.n_steps    = (int)n_steps;
    p.save_every = (int)(save_every > n_steps ? n_steps + 1 : save_every);
    p.n_saved    = fpa_n_saved(n_steps, save_every);
    p.y_trace    = (flags & FPA_OUT_TRACE) ? y_trace : nullptr;
    p.y_end      = (flags & FPA_OUT_END) ? y_end : nullptr;
    p.bad        = bad_scratch;
    p.check      = (flags & FPA_CHECK_NAN) ? 1 : 0;
    const int threads = 128;
    linear_rk4_kernel<<<(unsigned)((p.n_threads + threads - 1) / threads), threads, 0, st>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "linear_rk4_kernel launch");
    if (status && bad_scratch) {
        linear_status_kernel<<<(unsigned)((B + threads - 1) / threads), threads, 0, st>>>(B, dim, bad_scratch,
                                                                                        status);
        e = cudaGetLastError();
        if (e != cudaSuccess) return cuda_fail(e, "linear_status_kernel launch");
    }
    return FPA_OK;
}

}  // namespace fpa
