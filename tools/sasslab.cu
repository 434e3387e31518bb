// tools/sasslab.cu -- load cubin variants of the fast RK4 kernel with the driver API, time each on
// the headline workload and compare the outputs bit for bit with the first variant.  Used to test
// what the SASS post-pass (tools/sass_sched.py: operand-reuse flags, instruction order) changes.
//
// device code:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --fmad=true \
//                    -DSASSLAB_DEVICE -cubin -o /tmp/lab.cubin tools/sasslab.cu
// host program: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/sasslab.bin \
//                    tools/sasslab.cu -lcuda
// run:          tools/sasslab.bin <kernel-name> a.cubin [b.cubin ...]
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda.h>

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
int plan_fill(const fpa_plan_desc*, PlanParams&) { return FPA_ERR_UNSUPPORTED; }
}  // namespace fpa
extern "C" int64_t fpa_n_saved(int64_t n, int64_t s) { return n / s + 1; }
extern "C" int64_t fpa_interval_steps(double z_max, double dz) { return (int64_t)nearbyint(z_max / dz); }

using namespace fpa;

#ifndef SASSLAB_DEVICE
#define CU(x)                                                              \
    do {                                                                   \
        CUresult r_ = (x);                                                 \
        if (r_ != CUDA_SUCCESS) {                                          \
            const char* s_ = nullptr;                                      \
            cuGetErrorString(r_, &s_);                                     \
            fprintf(stderr, "%s failed: %s\n", #x, s_ ? s_ : "?");         \
            return 1;                                                      \
        }                                                                  \
    } while (0)

static uint64_t fnv(const void* p, size_t n) {
    const unsigned char* b = (const unsigned char*)p;
    uint64_t             h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) h = (h ^ b[i]) * 1099511628211ull;
    return h;
}

int main(int argc, char** argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s <kernel-name> a.cubin [b.cubin ...]\n", argv[0]);
        return 2;
    }
    const char* kname = argv[1];
    cudaFree(0);  // primary context
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int64_t B = getenv("LAB_POINTS") ? atoll(getenv("LAB_POINTS")) : 1000000;
    const int     n_steps = getenv("LAB_STEPS") ? atoi(getenv("LAB_STEPS")) : 2500;
    std::vector<double> dbeta(B);
    for (int64_t i = 0; i < B; ++i) dbeta[i] = -0.015 + 0.03 * (double)i / (double)B;
    const double consts[10] = {11.5e-3, 1.1512925464970228e-4, 0.31622776601683794, 0, 0.31622776601683794, 0,
                               3.1622776601683794e-4, 0, 3.1622776601683794e-4, 0};
    double *d_dbeta, *d_consts, *d_pmax, *d_end;
    int32_t* d_status;
    cudaMalloc(&d_dbeta, B * 8);
    cudaMalloc(&d_consts, sizeof(consts));
    cudaMalloc(&d_pmax, B * 32);
    cudaMalloc(&d_end, B * 64);
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
    p.A_end = d_end;
    p.status = d_status;
    p.z0 = 0.0;
    p.z_max = 500.0;
    p.h = 500.0 / n_steps;
    p.n_steps = n_steps;
    p.save_every = 10;
    p.check = 1;
    p.n_saved = n_steps / 10 + 1;
    p.coef = make_coef(consts[0], consts[1], p.h);

    printf("%s, %d SMs; kernel %s\n", prop.name, prop.multiProcessorCount, kname);
    std::vector<double> out(B * 12);
    uint64_t            h0 = 0;
    double              ms0 = 0;
    cudaEvent_t         e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int a = 2; a < argc; ++a) {
        CUmodule   mod;
        CUfunction fn;
        CU(cuModuleLoad(&mod, argv[a]));
        CU(cuModuleGetFunction(&fn, mod, kname));
        int regs = 0;
        cuFuncGetAttribute(&regs, CU_FUNC_ATTRIBUTE_NUM_REGS, fn);
        void*          args[] = {&p};
        const unsigned threads = getenv("LAB_THREADS") ? (unsigned)atoi(getenv("LAB_THREADS")) : 128u;
        const unsigned blocks = (unsigned)((B + threads - 1) / threads);
        cudaMemset(d_pmax, 0, B * 32);
        cudaMemset(d_end, 0, B * 64);
        float best = 1e30f;
        for (int r = 0; r < 4; ++r) {
            cudaEventRecord(e0);
            CU(cuLaunchKernel(fn, blocks, 1, 1, threads, 1, 1, 0, 0, args, nullptr));
            cudaEventRecord(e1);
            if (cudaEventSynchronize(e1) != cudaSuccess) {
                fprintf(stderr, "%s: kernel failed: %s\n", argv[a], cudaGetErrorString(cudaGetLastError()));
                return 1;
            }
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (r > 0 && ms < best) best = ms;
        }
        cudaMemcpy(out.data(), d_pmax, B * 32, cudaMemcpyDeviceToHost);
        cudaMemcpy(out.data() + B * 4, d_end, B * 64, cudaMemcpyDeviceToHost);
        const uint64_t h = fnv(out.data(), B * 96);
        if (a == 2) {
            h0 = h;
            ms0 = best;
        }
        const double tf = 568.0 * B * n_steps / (best * 1e-3) / 1e12;
        printf("%-40s regs=%3d  %8.3f ms  %6.2f TFLOP/s  x%.4f vs first  outputs %s (fnv %016llx)\n", argv[a], regs,
               best, tf, ms0 / best, h == h0 ? "BIT-IDENTICAL" : "DIFFER", (unsigned long long)h);
        fflush(stdout);
        cuModuleUnload(mod);
    }
    return 0;
}
#else
// device-code build: instantiate the kernels the lab loads by name
namespace fpa {
template __global__ void yaman4_fast_kernel<false, true, 1, 128, 3>(const Yaman4Params);
template __global__ void yaman4_fast_kernel<false, true, 1, 128, 4>(const Yaman4Params);
}  // namespace fpa
#endif
