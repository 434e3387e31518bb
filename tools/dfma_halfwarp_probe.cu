// tools/dfma_halfwarp_probe.cu -- does a DFMA of a warp with only 16 (8, 4, 1) active lanes leave the FP64 pipe
// sooner than a full one?  One warp per SM sub-partition, ILP 8, lanes >= `active` masked off by a branch.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dfma_halfwarp_probe.bin tools/dfma_halfwarp_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void probe(double* out, long long* cyc, int iters, double a, double b, int active, int upper) {
    constexpr int ILP = 8;
    double        acc[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) acc[k] = threadIdx.x * 1e-9 + k;
    const int lane = threadIdx.x & 31;
    const bool on = upper ? (lane >= 32 - active) : (lane < active);
    long long t0 = 0, t1 = 0;
    if (on) {
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
#pragma unroll
                for (int k = 0; k < ILP; ++k) acc[k] = fma(acc[k], a, b);
            }
        }
        t1 = clock64();
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (on && (lane == 0 || lane == 31) && blockIdx.x == 0 && threadIdx.x < 32) *cyc = t1 - t0;
}

int main() {
    double*    out;
    long long* cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(double));
    cudaMalloc(&cyc, sizeof(long long));
    const int iters = 20000;
    for (int warps : {4, 8, 16})
        for (int upper : {0, 1})
            for (int active : {32, 16, 8, 1}) {
                probe<<<148, 32 * warps>>>(out, cyc, iters, 0.999999, 1e-7, active, upper);
                cudaDeviceSynchronize();
                long long c;
                cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
                printf("warps/SM %2d  active lanes %2d (%s half): %.2f cycles per DFMA of one warp\n", warps, active,
                       upper ? "upper" : "lower", c / ((double)iters * 4 * 8));
            }
    return 0;
}
