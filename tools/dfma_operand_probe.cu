// tools/dfma_operand_probe.cu -- does the FP64 pipe rate depend on how many DISTINCT register
// operands a DFMA reads?  (register-file bandwidth vs the 16-lane FP64 pipe)
//   mode 0: v = fma(v, b, a)      a, b shared by all chains (operand-reuse friendly)
//   mode 1: v = fma(v, w_c, u_c)  three distinct 64-bit register operands per instruction
//   mode 2: v = fma(v, w_c, a)    two distinct register operands + one shared
//   mode 3: v = v * w_c           DMUL, two register operands
//   mode 4: v = fma(v, K1, K2)    constant-bank operands (kernel parameters)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dfma_probe tools/dfma_operand_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kChains = 12;
constexpr int kInner  = 32;

template <int MODE>
__global__ void __launch_bounds__(128) probe(int iters, double a, double b, double* sink) {
    double v[kChains], w[kChains], u[kChains];
#pragma unroll
    for (int c = 0; c < kChains; ++c) {
        v[c] = a + (threadIdx.x + c) * 1e-9;
        w[c] = b + (threadIdx.x * 3 + c) * 1e-12;
        u[c] = a * 1e-3 + c * 1e-10 * threadIdx.x;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < kInner; ++k) {
#pragma unroll
            for (int c = 0; c < kChains; ++c) {
                if (MODE == 0) v[c] = fma(v[c], b, a);
                if (MODE == 1) v[c] = fma(v[c], w[c], u[c]);
                if (MODE == 2) v[c] = fma(v[c], w[c], a);
                if (MODE == 3) v[c] = v[c] * w[c];
                if (MODE == 4) v[c] = fma(v[c], b, a);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < kChains; ++c) s += v[c] + w[c] + u[c];
    if (s == 123.456) sink[0] = s;
}

template <int MODE>
void run(const char* name, int blocks_per_sm, int sms) {
    double* sink;
    cudaMalloc(&sink, 64);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 4000, blocks = sms * blocks_per_sm;
    probe<MODE><<<blocks, 128>>>(iters / 10, 0.999999, 1.0000001, sink);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    probe<MODE><<<blocks, 128>>>(iters, 0.999999, 1.0000001, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)blocks * 128 * iters * kInner * kChains;
    printf("%-44s warps/SMSP=%d  %8.3f ms  %7.2f G inst-lanes/s  = %6.2f TFLOP/s-equivalent (x2)\n", name,
           blocks_per_sm, ms, ops / ms / 1e6, 2 * ops / ms / 1e9);
    cudaFree(sink);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    for (int bps : {1, 2, 4}) {
        run<0>("mode0 fma(v, b, a)   shared b,a (registers)", bps, p.multiProcessorCount);
        run<1>("mode1 fma(v, w_c, u_c) 3 distinct registers", bps, p.multiProcessorCount);
        run<2>("mode2 fma(v, w_c, a)   2 distinct + shared", bps, p.multiProcessorCount);
        run<3>("mode3 v * w_c          DMUL 2 registers", bps, p.multiProcessorCount);
    }
    return 0;
}
