// tools/corr_fma_probe.cu -- why does the correlation body of the N-wave comb kernel cost ~3 FP64-pipe cycles
// per FMA?  One warp per sub-partition runs the 8-term x 8-output complex MAC block from registers only
// (no loads), in several orderings of its 256 FMAs; prints cycles per warp-wide DFMA.
//   V0: as shipped in round 2 first version: per output  re += -ay*wy ; im += ay*wx ; then re += ax*wx ; im += ax*wy
//   V1: negated copy of a.y in its own register: four runs of 8 FMAs, each sharing ONE register operand, no modifiers
//   V2: like V1 but runs of 16 (re and im of the same a-register interleaved)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/corr_fma_probe tools/corr_fma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int V, int K>
__device__ __forceinline__ void term(double (&re)[8], double (&im)[8], double ax, double ay, double nay, const double2 (&win)[8]) {
    if (V == 0) {
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            re[t] = fma(-ay, win[(K + t) & 7].y, re[t]);
            im[t] = fma(ay, win[(K + t) & 7].x, im[t]);
        }
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            re[t] = fma(ax, win[(K + t) & 7].x, re[t]);
            im[t] = fma(ax, win[(K + t) & 7].y, im[t]);
        }
    } else if (V == 1) {
#pragma unroll
        for (int t = 0; t < 8; ++t) re[t] = fma(nay, win[(K + t) & 7].y, re[t]);
#pragma unroll
        for (int t = 0; t < 8; ++t) im[t] = fma(ay, win[(K + t) & 7].x, im[t]);
#pragma unroll
        for (int t = 0; t < 8; ++t) re[t] = fma(ax, win[(K + t) & 7].x, re[t]);
#pragma unroll
        for (int t = 0; t < 8; ++t) im[t] = fma(ax, win[(K + t) & 7].y, im[t]);
    } else {
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            re[t] = fma(nay, win[(K + t) & 7].y, re[t]);
            im[t] = fma(ay, win[(K + t) & 7].x, im[t]);
        }
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            re[t] = fma(ax, win[(K + t) & 7].x, re[t]);
            im[t] = fma(ax, win[(K + t) & 7].y, im[t]);
        }
    }
}

template <int V>
__global__ void probe(double* out, long long* cyc, int iters, double s) {
    double  re[8], im[8];
    double2 win[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        re[t] = threadIdx.x * 1e-9 + t;
        im[t] = threadIdx.x * 2e-9 - t;
        win[t] = make_double2(s + t * 1e-3 + threadIdx.x * 1e-7, s - t * 1e-3);
    }
    double ax = s * 0.5 + threadIdx.x * 1e-8, ay = s * 0.25;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#define T(K)                                                                  \
    {                                                                         \
        const double nay = -ay;                                               \
        term<V, K>(re, im, ax, ay, nay, win);                                 \
        win[K] = make_double2(win[K].y * 0.999, win[K].x);                    \
        const double tmp = ax;                                                \
        ax = ay;                                                              \
        ay = tmp * 0.999;                                                     \
    }
        T(0) T(1) T(2) T(3) T(4) T(5) T(6) T(7)
#undef T
    }
    const long long t1 = clock64();
    double acc = 0.0;
#pragma unroll
    for (int t = 0; t < 8; ++t) acc += re[t] + im[t];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int V>
void run(int warps_per_sm) {
    double*    out;
    long long* cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(double));
    cudaMalloc(&cyc, sizeof(long long));
    const int iters = 4000;
    probe<V><<<148, 32 * warps_per_sm>>>(out, cyc, iters, 0.7);
    cudaDeviceSynchronize();
    probe<V><<<148, 32 * warps_per_sm>>>(out, cyc, iters, 0.7);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    printf("V%d  warps per sub-partition %d: %.3f cycles per DFMA of one warp (x%d warps = %.3f pipe cycles per DFMA)\n", V,
           warps_per_sm / 4, c / (double)(iters * 264.0), warps_per_sm / 4, c / (double)(iters * 264.0) / (warps_per_sm / 4));
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    for (int w : {4, 16}) {
        run<0>(w);
        run<1>(w);
        run<2>(w);
    }
    return 0;
}
