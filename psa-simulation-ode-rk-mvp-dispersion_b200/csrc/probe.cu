// probe.cu -- FP64 FMA peak micro-benchmark: the roofline denominator of this solver.
//
// MEASURED_PEAKS.json holds HBM and bf16 figures only, so the FP64 FMA peak is measured on the
// box: every thread keeps 16 independent DFMA chains in registers (no memory traffic in the loop),
// enough resident warps to cover the FP64 pipe latency on every SM sub-partition.
#include "fpa_common.cuh"

namespace fpa {

constexpr int kChains = 16;
constexpr int kInner  = 64;  // FMAs per chain per outer iteration

__global__ void __launch_bounds__(256) dfma_probe_kernel(int iters, double a, double b, double* sink) {
    double v[kChains];
#pragma unroll
    for (int c = 0; c < kChains; ++c) v[c] = a + (double)(threadIdx.x + c) * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < kInner; ++u) {
#pragma unroll
            for (int c = 0; c < kChains; ++c) v[c] = fma(v[c], b, a);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < kChains; ++c) s += v[c];
    if (s == 123456.789) sink[0] = s;  // never true; keeps the chains alive
}

int probe_run(int device, int iters, double* tflops, double* ms_out) {
    FPA_REQUIRE(iters >= 1, "iters must be >= 1");
    int rc = use_device(device);
    if (rc != FPA_OK) return rc;
    cudaDeviceProp prop;
    FPA_CUDA(cudaGetDeviceProperties(&prop, device));
    void* sink = nullptr;
    rc = workspace(device, 15, 64, &sink);
    if (rc != FPA_OK) return rc;
    const int threads = 256;
    const int blocks  = prop.multiProcessorCount * 4;
    cudaEvent_t e0, e1;
    FPA_CUDA(cudaEventCreate(&e0));
    FPA_CUDA(cudaEventCreate(&e1));
    // warm-up, then the timed launch
    dfma_probe_kernel<<<blocks, threads>>>(iters / 8 + 1, 0.999999, 1.0000001, (double*)sink);
    FPA_CUDA(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        FPA_CUDA(cudaEventRecord(e0));
        dfma_probe_kernel<<<blocks, threads>>>(iters, 0.999999, 1.0000001, (double*)sink);
        FPA_CUDA(cudaEventRecord(e1));
        FPA_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        FPA_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double fmas = (double)blocks * threads * (double)iters * kInner * kChains;
    if (tflops) *tflops = 2.0 * fmas / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return FPA_OK;
}

}  // namespace fpa
