// tools/tune_yaman4.cu -- launch-shape sweep of the fused RK4 kernel (threads per block x resident
// blocks per SM), on the headline workload (1e6 points x 2500 steps, reduce mode, uniform physics)
// and on a tail-free batch (an exact multiple of the resident thread count).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I. -o /tmp/tune_yaman4 tools/tune_yaman4.cu
#include <cstdarg>
#include <cstdio>
#include <vector>

#include "../psa-simulation-ode-rk-mvp-dispersion_b200/csrc/yaman4.cu"

namespace fpa {
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    fputc('\n', stderr);
}
int cuda_fail(cudaError_t e, const char* what) {
    fprintf(stderr, "CUDA error %s in %s\n", cudaGetErrorString(e), what);
    return FPA_ERR_CUDA;
}
}  // namespace fpa
extern "C" int64_t fpa_n_saved(int64_t n, int64_t s) { return n / s + 1; }
extern "C" int64_t fpa_interval_steps(double z_max, double dz) { return (int64_t)nearbyint(z_max / dz); }
namespace fpa {
int plan_fill(const fpa_plan_desc*, PlanParams&) { return FPA_ERR_UNSUPPORTED; }  // sweeps are not tuned here
}

using namespace fpa;

template <int THREADS, int MB>
float time_variant(const Yaman4Params& p, int reps) {
    const long blocks = (long)((p.n_points + THREADS - 1) / THREADS);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    yaman4_fast_kernel<false, true, 1, THREADS, MB><<<blocks, THREADS>>>(p);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0);
        yaman4_fast_kernel<false, true, 1, THREADS, MB><<<blocks, THREADS>>>(p);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("  !! %s\n", cudaGetErrorString(e));
    return best;
}

template <int THREADS, int MB>
void report(Yaman4Params p, int sms, double peak_tf) {
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, yaman4_fast_kernel<false, true, 1, THREADS, MB>);
    int resident = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, yaman4_fast_kernel<false, true, 1, THREADS, MB>,
                                                  THREADS, 0);
    const int64_t full = p.n_points;
    const float   ms_full = time_variant<THREADS, MB>(p, 3);
    // tail-free batch: a whole number of waves
    const int64_t per_wave = (int64_t)sms * resident * THREADS;
    p.n_points = (full / per_wave) * per_wave;
    if (p.n_points == 0) p.n_points = per_wave;
    const float  ms_even = time_variant<THREADS, MB>(p, 3);
    const double tf_full = 568.0 * full * p.n_steps / (ms_full * 1e-3) / 1e12;
    const double tf_even = 568.0 * p.n_points * p.n_steps / (ms_even * 1e-3) / 1e12;
    printf("threads=%3d min_blocks=%d regs=%3d resident=%2d blocks (%2d warps/SMSP) | full batch: %7.3f ms %6.2f TF "
           "(%4.1f%%) | %lld pts (whole waves): %7.3f ms %6.2f TF (%4.1f%%)\n",
           THREADS, MB, fa.numRegs, resident, resident * THREADS / 128, ms_full, tf_full, 100 * tf_full / peak_tf,
           (long long)p.n_points, ms_even, tf_even, 100 * tf_even / peak_tf);
}

int main(int argc, char** argv) {
    const double peak_tf = argc > 1 ? atof(argv[1]) : 37.0;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int     sms = prop.multiProcessorCount;
    const int64_t B = argc > 2 ? atoll(argv[2]) : 1000000;
    const int     n_steps = argc > 3 ? atoi(argv[3]) : 2500;
    std::vector<double> dbeta(B);
    for (int64_t i = 0; i < B; ++i) dbeta[i] = -0.015 + 0.03 * (double)i / (double)B;
    const double consts[10] = {11.5e-3, 1.1512925464970228e-4, 0.31622776601683794, 0, 0.31622776601683794, 0,
                               3.1622776601683794e-4, 0, 3.1622776601683794e-4, 0};
    double *d_dbeta, *d_consts, *d_pmax;
    int32_t* d_status;
    cudaMalloc(&d_dbeta, B * 8);
    cudaMalloc(&d_consts, sizeof(consts));
    cudaMalloc(&d_pmax, B * 32);
    cudaMalloc(&d_status, B * 4);
    cudaMemcpy(d_dbeta, dbeta.data(), B * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(d_consts, consts, sizeof(consts), cudaMemcpyHostToDevice);

    Yaman4Params p{};
    p.n_points = B;
    p.dbeta = d_dbeta;
    p.gamma = d_consts;
    p.alpha = d_consts + 1;
    p.A0 = d_consts + 2;
    p.Pmax = d_pmax;
    p.status = d_status;
    p.z0 = 0.0;
    p.z_max = 500.0;
    p.h = 500.0 / n_steps;
    p.n_steps = n_steps;
    p.save_every = 10;
    p.check = 1;
    p.n_saved = n_steps / 10 + 1;
    p.coef = make_coef(consts[0], consts[1], p.h);

    printf("%s, %d SMs; percentages are of %.2f TFLOP/s\n", prop.name, sms, peak_tf);
    report<128, 3>(p, sms, peak_tf);
    report<128, 4>(p, sms, peak_tf);
    report<128, 5>(p, sms, peak_tf);
    report<128, 6>(p, sms, peak_tf);
    report<64, 8>(p, sms, peak_tf);
    report<64, 10>(p, sms, peak_tf);
    report<64, 12>(p, sms, peak_tf);
    report<256, 2>(p, sms, peak_tf);
    report<256, 3>(p, sms, peak_tf);
    report<32, 16>(p, sms, peak_tf);
    report<32, 20>(p, sms, peak_tf);
    report<32, 24>(p, sms, peak_tf);
    return 0;
}
