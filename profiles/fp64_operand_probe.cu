// FP64 FMA issue rate as a function of where the operands come from (one B200, standalone):
//   A  x_i = fma(x_i, a, b)        two operands shared by all chains   (the in-library peak benchmark)
//   B  x_i = fma(a_i, b_i, x_i)    three distinct registers per FMA, a_i / b_i loop-invariant
//   C  x_i = fma(a_i, y_j, x_i)    the tangent kernel's pattern: coefficient a_i used once, y_j shared by a few chains
// Build + run:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_probe fp64_operand_probe.cu && ./fp64_probe
#include <cstdio>
#include <cuda_runtime.h>

constexpr int NCH = 8;

template <int MODE>
__global__ void __launch_bounds__(256) probe(double* out, int iters, double a0, double b0) {
    double x[NCH], a[NCH], b[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) { x[i] = threadIdx.x * 1e-3 + i; a[i] = a0 + i * 1e-9; b[i] = b0 + i * 1e-10; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < NCH; ++i) {
                if (MODE == 0) x[i] = fma(x[i], a0, b0);
                else if (MODE == 1) x[i] = fma(a[i], b[i], x[i]);
                else x[i] = fma(a[(i + r) % NCH], b[i / 2], x[i]);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
double run(int sms, double* d) {
    const int blocks = sms * 8, iters = 4000;
    probe<MODE><<<blocks, 256>>>(d, 10, 0.999999, 1e-9);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<MODE><<<blocks, 256>>>(d, iters, 0.999999, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return 2.0 * NCH * 8 * (double)iters * blocks * 256 / (ms * 1e-3) * 1e-12;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double* d; cudaMalloc(&d, (size_t)p.multiProcessorCount * 8 * 256 * 8);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    printf("A  shared operands        : %.2f TFLOP/s\n", run<0>(p.multiProcessorCount, d));
    printf("B  three distinct operands: %.2f TFLOP/s\n", run<1>(p.multiProcessorCount, d));
    printf("C  coefficient x stage    : %.2f TFLOP/s\n", run<2>(p.multiProcessorCount, d));
    return 0;
}
