// tools/dfma_warp_probe.cu -- FP64 FMA issue rate of ONE warp versus several warps per SM sub-partition.
// Each warp runs ILP independent DFMA chains (operands: one reused register, accumulators), long enough
// that only the issue rate matters.  Prints cycles per warp-wide DFMA seen by one warp and the SM-wide rate.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dfma_warp_probe.bin tools/dfma_warp_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void probe(double* out, long long* cyc, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) acc[k] = threadIdx.x * 1e-9 + k;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int k = 0; k < ILP; ++k) acc[k] = fma(acc[k], a, b);
        }
    }
    const long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP>
void run(int warps_per_sm) {
    double*    out;
    long long* cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(double));
    cudaMalloc(&cyc, sizeof(long long));
    const int iters = 20000;
    const int threads = 32 * warps_per_sm;  // one CTA per SM; warps spread round-robin over the 4 sub-partitions
    probe<ILP><<<148, threads>>>(out, cyc, iters, 0.999999, 1e-7);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<ILP><<<148, threads>>>(out, cyc, iters, 0.999999, 1e-7);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long c;
    cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    const double n = (double)iters * 4 * ILP;
    printf("ILP %2d  warps/SM %2d (%.2f per sub-partition): %.2f cycles per DFMA of one warp, %.2f TFLOP/s\n", ILP, warps_per_sm,
           warps_per_sm / 4.0, c / n, 148.0 * warps_per_sm * 32 * n * 2 / (ms * 1e-3) / 1e12);
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    for (int w : {4, 8, 12, 16}) run<4>(w);
    for (int w : {4, 8, 12, 16}) run<8>(w);
    for (int w : {4, 8, 12, 16}) run<16>(w);
    return 0;
}
