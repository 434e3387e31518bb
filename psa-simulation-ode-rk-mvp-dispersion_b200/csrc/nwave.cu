// nwave.cu -- N-wave generalisation of the fused RK4 x FWM integrator (NOT in the reference).
//
//   dA_n/dz = -(alpha/2) A_n + i*gamma * [ (2*sum_j P_j - P_n) A_n
//               + conj(E_n) * sum_{e in row n} D_e * At_k At_l conj(At_m) ],   At_j = A_j E_j,
//   E_j = exp(i*beta_j*z)
// which is the reference's 4-wave system (yaman_model.py:22-25, :135-186) for the fixed table
// {0:(2,3;1) 1:(2,3;0) 2:(0,1;3) 3:(0,1;2)}, weight 2 and beta = [0,0,0,dbeta].
//
// Mapping: one CTA per scan point (a single warp when the plan is small).  The point's complex
// amplitudes, its per-wave phase table and -- when it fits -- the frequency plan's triplet
// index/weight list live in shared memory for all z-steps; each warp owns rows n = w, w+W, ...
// of the triplet sum, lanes stride over the row's entries and the partial sums are combined
// with warp shuffles.  The four RK4 stages are fused: nothing goes to HBM between stages.
//
// The triplet enumerator (integer-grid matching, canonical order n,k,l,m) is host code in this
// file as well; tests compare it bit-for-bit with the Python restatement in oracle/.
#include "fpa_common.cuh"

#include <math.h>
#include <vector>

namespace fpa {

struct NwaveParams {
    int64_t            n_points;
    int                n_waves;
    int                beta_stride, gamma_stride, alpha_stride, A0_stride;
    const double*      beta;
    const double*      gamma;
    const double*      alpha;
    const double*      A0;
    const fpa_triplet* triplets;
    const int64_t*     row_ptr;
    int64_t            n_triplets;
    int                table_in_smem;
    int                table_in_smem_only_one_cta;
    double             z0, z_max;
    int                n_steps, save_every;
    int64_t            n_saved;
    const double*      z_grid;
    double*            A_trace;
    double*            A_end;
    double*            Pmax;
    int32_t*           status;
    int                check;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Shared-memory layout (doubles): y[2N] ys[2N] yn[2N] At[2N] E[2N] beta[N] P[N] red[34] | rows[N+1] | table
struct Smem {
    double *y, *ys, *yn, *At, *E, *beta, *P, *red;
    int*                rows;
    const fpa_triplet*  table;
};

__device__ __forceinline__ Smem carve(double* base, int N, const NwaveParams& p) {
    Smem s;
    s.y    = base;
    s.ys   = s.y + 2 * N;
    s.yn   = s.ys + 2 * N;
    s.At   = s.yn + 2 * N;
    s.E    = s.At + 2 * N;
    s.beta = s.E + 2 * N;
    s.P    = s.beta + N;
    s.red  = s.P + N;
    s.rows = reinterpret_cast<int*>(s.red + 34);
    // table is 8-byte aligned: (N+1) ints rounded up to an even count
    s.table = reinterpret_cast<const fpa_triplet*>(s.rows + ((N + 2) & ~1));
    return s;
}

// One RHS evaluation at abscissa z on stage state s.ys; afterwards, for every wave n, the row owner
// applies   yn_n += wa*k_n   and   ys_n = y_n + wb*k_n   (or y_n = yn_n + wa*k_n when `last`).
// Returns S = sum |ys|^2 (used by the finite check).
__device__ double rhs_stage(const Smem& s, const NwaveParams& p, const fpa_triplet* __restrict__ table,
                            double z, double gamma, double nha, double wa, double wb, bool last) {
    const int N = p.n_waves;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;

    // phases, rotated amplitudes, powers
    double part = 0.0;
    for (int j = tid; j < N; j += blockDim.x) {
        double sn, cs;
        sincos(s.beta[j] * z, &sn, &cs);
        const double xr = s.ys[2 * j], xi = s.ys[2 * j + 1];
        s.E[2 * j]      = cs;
        s.E[2 * j + 1]  = sn;
        s.At[2 * j]     = fma(-xi, sn, xr * cs);
        s.At[2 * j + 1] = fma(xr, sn, xi * cs);
        const double P  = fma(xi, xi, xr * xr);
        s.P[j] = P;
        part += P;
    }
    part = warp_sum(part);
    if (lane == 0) s.red[warp] = part;
    __syncthreads();
    double S = 0.0;
    for (int w = 0; w < nwarps; ++w) S += s.red[w];

    // triplet sums, one warp per row
    for (int n = warp; n < N; n += nwarps) {
        const int e0 = s.rows[n], e1 = s.rows[n + 1];
        double rr = 0.0, ri = 0.0;
        // one entry: D * At_k At_l conj(At_m)
        auto term = [&](const fpa_triplet t, double& ar, double& ai) {
            const double kr = s.At[2 * t.k], ki = s.At[2 * t.k + 1];
            const double lr = s.At[2 * t.l], li = s.At[2 * t.l + 1];
            const double mr = s.At[2 * t.m], mi = s.At[2 * t.m + 1];
            const double w  = (double)t.weight;
            const double qr = w * fma(-ki, li, kr * lr);
            const double qi = w * fma(kr, li, ki * lr);
            ar = fma(qr, mr, fma(qi, mi, ar));   // q * conj(m)
            ai = fma(qi, mr, fma(-qr, mi, ai));
        };
        // A table that does not fit into shared memory streams from L2 (N = 64: 674 KB per RHS and scan point):
        // four entries per lane are requested before the first one is used, and two accumulator pairs keep the
        // FMA chains apart -- the rolled loop ran at the latency of one load per entry (~800 cycles).
        double r2 = 0.0, i2 = 0.0;
        int    e = e0 + lane;
        for (; e + 96 < e1; e += 128) {
            const fpa_triplet t0 = table[e], t1 = table[e + 32], t2 = table[e + 64], t3 = table[e + 96];
            term(t0, rr, ri);
            term(t1, r2, i2);
            term(t2, rr, ri);
            term(t3, r2, i2);
        }
        for (; e < e1; e += 32) term(table[e], rr, ri);
        rr += r2;
        ri += i2;
        rr = warp_sum(rr);
        ri = warp_sum(ri);
        if (lane == 0) {
            const double xr = s.ys[2 * n], xi = s.ys[2 * n + 1];
            const double er = s.E[2 * n], ei = s.E[2 * n + 1];
            // F = conj(E_n) * R
            const double fr = fma(ri, ei, rr * er);
            const double fi = fma(ri, er, -(rr * ei));
            const double G  = gamma * ((S + S) - s.P[n]);
            // k = nha*x + i*(G*x + gamma*F)
            const double kr = fma(nha, xr, -fma(G, xi, gamma * fi));
            const double ki = fma(nha, xi, fma(G, xr, gamma * fr));
            if (last) {
                s.y[2 * n]     = fma(wa, kr, s.yn[2 * n]);
                s.y[2 * n + 1] = fma(wa, ki, s.yn[2 * n + 1]);
            } else {
                const double a = s.y[2 * n], bq = s.y[2 * n + 1];
                // first stage starts the accumulator from y (wb == 0 marks "yn not yet initialised")
                s.yn[2 * n]     = fma(wa, kr, s.yn[2 * n]);
                s.yn[2 * n + 1] = fma(wa, ki, s.yn[2 * n + 1]);
                s.ys[2 * n]     = fma(wb, kr, a);
                s.ys[2 * n + 1] = fma(wb, ki, bq);
            }
        }
    }
    __syncthreads();
    return S;
}

__global__ void nwave_rk4_kernel(const NwaveParams p) {
    extern __shared__ double smem_raw[];
    const int     N = p.n_waves;
    const int64_t b = blockIdx.x;
    const int     tid = threadIdx.x;
    Smem          s = carve(smem_raw, N, p);

    const double gamma = p.gamma[b * p.gamma_stride];
    const double nha   = -0.5 * p.alpha[b * p.alpha_stride];

    for (int j = tid; j < N; j += blockDim.x) {
        s.beta[j] = p.beta[b * p.beta_stride * N + j];
        const double re = p.A0[(b * p.A0_stride * N + j) * 2];
        const double im = p.A0[(b * p.A0_stride * N + j) * 2 + 1];
        s.y[2 * j] = re;
        s.y[2 * j + 1] = im;
    }
    for (int j = tid; j <= N; j += blockDim.x) s.rows[j] = (int)p.row_ptr[j];
    const fpa_triplet* table = p.triplets;
    if (p.table_in_smem) {
        fpa_triplet* dst = const_cast<fpa_triplet*>(s.table);
        for (int64_t e = tid; e < p.n_triplets; e += blockDim.x) dst[e] = p.triplets[e];
        table = s.table;
    }
    __syncthreads();

    double* tr = p.A_trace ? p.A_trace + b * p.n_saved * 2 * N : nullptr;
    if (tr) {
        for (int j = tid; j < 2 * N; j += blockDim.x) tr[j] = s.y[j];
        tr += 2 * N;
    }
    // per-thread running max for the waves this thread owns (j = tid, tid+blockDim, ...): at most
    // ceil(128/32) = 4 waves per thread
    double pm[4] = {0.0, 0.0, 0.0, 0.0};
    if (p.Pmax) {
        int q = 0;
        for (int j = tid; j < N; j += blockDim.x, ++q) pm[q] = fma(s.y[2 * j + 1], s.y[2 * j + 1], s.y[2 * j] * s.y[2 * j]);
    }

    const int    n_steps = p.n_steps;
    const double z0 = p.z0, z_max = p.z_max;
    const double step = (z_max - z0) / (double)n_steps;
    double       zi = p.z_grid ? p.z_grid[0] : z0;
    double       di = 0.0;
    int          save_ctr = p.save_every;
    int32_t      bad = FPA_POINT_OK;

    for (int i = 0; i < n_steps; ++i) {
        double zn;
        if (p.z_grid) {
            zn = p.z_grid[i + 1];
        } else {
            di += 1.0;
            zn = (i + 1 == n_steps) ? z_max : __dadd_rn(__dmul_rn(di, step), z0);
        }
        const double h = zn - zi, hh = 0.5 * h, h6 = h / 6.0, h3 = h6 + h6;

        // stage state and accumulator start from y
        for (int j = tid; j < 2 * N; j += blockDim.x) {
            const double v = s.y[j];
            s.ys[j] = v;
            s.yn[j] = v;
        }
        __syncthreads();

        const double S = rhs_stage(s, p, table, zi, gamma, nha, h6, hh, false);
        if (p.check && i > 0 && bad == FPA_POINT_OK && nonfinite(S)) {
            int nf = 0;
            for (int j = tid; j < 2 * N; j += blockDim.x) nf |= nonfinite(s.y[j]) ? 1 : 0;
            if (__syncthreads_or(nf)) bad = i - 1;
        }
        rhs_stage(s, p, table, zi + hh, gamma, nha, h3, hh, false);
        rhs_stage(s, p, table, zi + hh, gamma, nha, h3, h, false);
        rhs_stage(s, p, table, zi + h, gamma, nha, h6, 0.0, true);
        zi = zn;

        if (--save_ctr == 0) {
            save_ctr = p.save_every;
            if (tr) {
                for (int j = tid; j < 2 * N; j += blockDim.x) tr[j] = s.y[j];
                tr += 2 * N;
            }
            if (p.Pmax) {
                int q = 0;
                for (int j = tid; j < N; j += blockDim.x, ++q) {
                    const double P = fma(s.y[2 * j + 1], s.y[2 * j + 1], s.y[2 * j] * s.y[2 * j]);
                    pm[q] = (P != P || pm[q] != pm[q]) ? qnan() : fmax(pm[q], P);
                }
            }
        }
    }

    if (p.check && bad == FPA_POINT_OK) {
        int nf = 0;
        for (int j = tid; j < 2 * N; j += blockDim.x) nf |= nonfinite(s.y[j]) ? 1 : 0;
        if (__syncthreads_or(nf)) bad = n_steps - 1;
    }
    if (p.status && tid == 0) p.status[b] = bad;
    if (p.A_end)
        for (int j = tid; j < 2 * N; j += blockDim.x) p.A_end[b * 2 * N + j] = s.y[j];
    if (p.Pmax) {
        int q = 0;
        for (int j = tid; j < N; j += blockDim.x, ++q) p.Pmax[b * N + j] = pm[q];
    }
}

static size_t nwave_smem_bytes(int N, int64_t n_table_entries) {
    size_t doubles = (size_t)(5 * 2 * N + 2 * N + 34);
    size_t bytes   = doubles * sizeof(double) + (size_t)((N + 2) & ~1) * sizeof(int);
    return bytes + (size_t)n_table_entries * sizeof(fpa_triplet);
}

int nwave_launch(const fpa_nwave_desc* d, cudaStream_t st) {
    FPA_REQUIRE(d != nullptr, "descriptor is NULL");
    FPA_REQUIRE(d->n_points >= 0, "n_points must be >= 0");
    FPA_REQUIRE(d->n_waves >= 1, "n_waves must be >= 1");
    if (d->n_waves > 128) {
        set_error("n_waves = %d exceeds the kernel limit of 128", d->n_waves);
        return FPA_ERR_UNSUPPORTED;
    }
    FPA_REQUIRE(d->n_steps >= 1 && d->n_steps < 2147483647LL, "n_steps must be in [1, 2^31)");
    FPA_REQUIRE(d->save_every >= 1, "save_every must be a positive integer");
    FPA_REQUIRE(d->beta && d->gamma && d->alpha && d->A0, "beta/gamma/alpha/A0 must be set");
    FPA_REQUIRE(d->n_triplets >= 0 && d->n_triplets < 2147483647LL, "bad n_triplets");
    FPA_REQUIRE(d->row_ptr != nullptr, "row_ptr must be set");
    FPA_REQUIRE(d->n_triplets == 0 || d->triplets, "triplets must be set");
    FPA_REQUIRE((d->gamma_stride | 1) == 1 && (d->alpha_stride | 1) == 1 && (d->A0_stride | 1) == 1 &&
                    (d->beta_stride | 1) == 1,
                "strides must be 0 (broadcast) or 1 (per point)");
    const bool trace = (d->flags & FPA_OUT_TRACE) != 0;
    const bool pmax  = (d->flags & FPA_OUT_PMAX) != 0;
    const bool endo  = (d->flags & FPA_OUT_END) != 0;
    FPA_REQUIRE(!trace || d->A_trace, "FPA_OUT_TRACE needs A_trace");
    FPA_REQUIRE(!pmax || d->Pmax, "FPA_OUT_PMAX needs Pmax");
    FPA_REQUIRE(!endo || d->A_end, "FPA_OUT_END needs A_end");
    if (d->n_points == 0) return FPA_OK;

    NwaveParams p;
    p.n_points     = d->n_points;
    p.n_waves      = d->n_waves;
    p.beta_stride  = (int)d->beta_stride;
    p.gamma_stride = (int)d->gamma_stride;
    p.alpha_stride = (int)d->alpha_stride;
    p.A0_stride    = (int)d->A0_stride;
    p.beta         = d->beta;
    p.gamma        = d->gamma;
    p.alpha        = d->alpha;
    p.A0           = d->A0;
    p.triplets     = d->triplets;
    p.row_ptr      = d->row_ptr;
    p.n_triplets   = d->n_triplets;
    p.z0           = d->z0;
    p.z_max        = d->z_max;
    p.n_steps      = (int)d->n_steps;
    p.save_every   = (int)(d->save_every > d->n_steps ? d->n_steps + 1 : d->save_every);
    p.n_saved      = fpa_n_saved(d->n_steps, d->save_every);
    p.z_grid       = nullptr;  // the C ABI exposes linspace grids only for the N-wave model
    p.A_trace      = trace ? d->A_trace : nullptr;
    p.A_end        = endo ? d->A_end : nullptr;
    p.Pmax         = pmax ? d->Pmax : nullptr;
    p.status       = d->status;
    p.check        = (d->flags & FPA_CHECK_NAN) ? 1 : 0;

    // the triplet list stays in shared memory when it fits beside the state (<= 200 KB in total)
    const size_t with_table = nwave_smem_bytes(d->n_waves, d->n_triplets);
    p.table_in_smem = with_table <= 200 * 1024 ? 1 : 0;
    const size_t smem = p.table_in_smem ? with_table : nwave_smem_bytes(d->n_waves, 0);
    p.table_in_smem_only_one_cta = smem > 100 * 1024 ? 1 : 0;
    // Warps per point: one warp per row of the triplet sum (up to 32 warps) when the batch is too
    // small to fill the GPU with points -- a single run (B = 1) is one CTA and its only parallelism
    // is across rows and entries; for large batches 8 warps per point leave room for several CTAs
    // per SM.
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int warps = d->n_waves < 32 ? d->n_waves : 32;
    if (d->n_points >= 4 * (int64_t)sms && !p.table_in_smem_only_one_cta) warps = warps < 8 ? warps : 8;
    const int threads = 32 * warps;

    cudaError_t e = cudaFuncSetAttribute(nwave_rk4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(nwave_rk4_kernel)");
    FPA_REQUIRE(d->n_points < 2147483647LL, "n_points too large for one launch");
    nwave_rk4_kernel<<<(unsigned)d->n_points, threads, smem, st>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "nwave_rk4_kernel launch");
    return FPA_OK;
}

}  // namespace fpa

// ------------------------------------------------------------------ host: triplet enumeration
extern "C" int64_t fpa_enumerate_triplets(int32_t N, const int32_t* g, fpa_triplet* out, int64_t cap,
                                          int64_t* row_ptr) {
    if (N < 1 || N > 32767 || g == nullptr) {
        fpa::set_error("fpa_enumerate_triplets: need 1 <= N <= 32767 and a grid index array");
        return -1;
    }
    int64_t count = 0;
    for (int32_t n = 0; n < N; ++n) {
        if (row_ptr) row_ptr[n] = count;
        for (int32_t k = 0; k < N; ++k) {
            for (int32_t l = k; l < N; ++l) {
                const int64_t target = (int64_t)g[k] + (int64_t)g[l] - (int64_t)g[n];
                for (int32_t m = 0; m < N; ++m) {
                    if (m == k || m == l) continue;
                    if ((int64_t)g[m] != target) continue;
                    if (out) {
                        if (count >= cap) {
                            fpa::set_error("fpa_enumerate_triplets: output capacity %lld too small",
                                           (long long)cap);
                            return -1;
                        }
                        out[count].k = (int16_t)k;
                        out[count].l = (int16_t)l;
                        out[count].m = (int16_t)m;
                        out[count].weight = (int16_t)(k == l ? 1 : 2);
                    }
                    ++count;
                }
            }
        }
    }
    if (row_ptr) row_ptr[N] = count;
    return count;
}

extern "C" int64_t fpa_enumerate_triplets_omega(int32_t N, const double* omega, double atol, double rtol, fpa_triplet* out,
                                                int64_t cap, int64_t* row_ptr) {
    if (N < 1 || N > 32767 || omega == nullptr || !(atol >= 0.0) || !(rtol >= 0.0)) {
        fpa::set_error("fpa_enumerate_triplets_omega: need 1 <= N <= 32767, an omega array and non-negative tolerances");
        return -1;
    }
    int64_t count = 0;
    for (int32_t n = 0; n < N; ++n) {
        if (row_ptr) row_ptr[n] = count;
        for (int32_t k = 0; k < N; ++k) {
            for (int32_t l = k; l < N; ++l) {
                const volatile double lhs = omega[k] + omega[l];   // rounded to double, as numpy does
                for (int32_t m = 0; m < N; ++m) {
                    if (m == k || m == l) continue;
                    const volatile double rhs = omega[m] + omega[n];
                    // numpy.isclose for finite values: |a - b| <= atol + rtol * |b|
                    const volatile double diff = lhs - rhs;
                    const volatile double tol = rtol * fabs(rhs);
                    if (!(fabs(diff) <= atol + tol)) continue;
                    if (out) {
                        if (count >= cap) {
                            fpa::set_error("fpa_enumerate_triplets_omega: output capacity %lld too small", (long long)cap);
                            return -1;
                        }
                        out[count].k = (int16_t)k;
                        out[count].l = (int16_t)l;
                        out[count].m = (int16_t)m;
                        out[count].weight = (int16_t)(k == l ? 1 : 2);
                    }
                    ++count;
                }
            }
        }
    }
    if (row_ptr) row_ptr[N] = count;
    return count;
}

extern "C" double fpa_nwave_flops_per_step(int32_t n_waves, int64_t n_triplets, int64_t n_pairs) {
    // per RHS 8*T_terms + 6*T_pairs + c*N with c = 30 (phase rotate 6, power 3,
    // Kerr factor 3, conj(E_n)*R 6, assemble 12); per step 4 RHS + 26*N for the RK4 combine.
    const double rhs = 8.0 * (double)n_triplets + 6.0 * (double)n_pairs + 30.0 * (double)n_waves;
    return 4.0 * rhs + 26.0 * (double)n_waves;
}
