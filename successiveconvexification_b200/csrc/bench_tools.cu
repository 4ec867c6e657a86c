// bench_tools.cu — measurement helpers for bench.py and profiles/ (libscvx_benchtools.so).  NOT part of the product
// ABI: nothing in include/scvx_b200.h, the Julia shim or the package's host mirror refers to this library.
//
//   scvx_bench_fp64_peak   DFMA issue-rate microbenchmark in the tangent kernel's FMA shape: the measured FP64 roofline
//                          denominator (MEASURED_PEAKS.json holds no FP64 figure).
//   scvx_bench_dmma_probe  the DMMA experiment BASELINE.json's north_star asks for: does the FP64 tensor path
//                          (mma.sync m8n8k4 / m16n8k8 .f64) add throughput to, or only compete with, the FP64 FMA pipe?
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>

namespace {

// 8 independent chains per thread, x_i = fma(coefficient, stage value, x_i) with the chain through the addend
// (chains through the multiplicand with two shared operands stop ~4 % lower, profiles/fp64_operand_probe.cu)
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double a0, double b0) {
    double x[8], a[8], y[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 1e-3 + i; a[i] = a0 + i * 1e-9; }
#pragma unroll
    for (int i = 0; i < 4; ++i) y[i] = b0 + i * 1e-10;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = fma(a[(i + r) & 7], y[i >> 1], x[i]);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x[0] + x[1]) + (x[2] + x[3])) + ((x[4] + x[5]) + (x[6] + x[7]));
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

// MODE 0: DFMA only (64 FMA per thread and iteration)
// MODE 1: DMMA m8n8k4 only, 8 independent accumulator tiles per warp (8 x 512 flop per iteration)
// MODE 2: DMMA m16n8k8 only, 4 independent accumulator tiles per warp (4 x 2048 flop per iteration)
// MODE 3: even warps DFMA (as MODE 0), odd warps DMMA m8n8k4 (as MODE 1): co-issue from different warps of one SM
// MODE 4: every warp interleaves 32 DFMA with 4 DMMA m8n8k4 per iteration (co-issue inside one warp)
template <int MODE>
__global__ void __launch_bounds__(256) dmma_probe_kernel(double* out, int iters, double a0, double b0) {
    const int warp = threadIdx.x >> 5;
    double x[8], a[8], y[4], c[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 1e-3 + i; a[i] = a0 + i * 1e-9; }
#pragma unroll
    for (int i = 0; i < 4; ++i) y[i] = b0 + i * 1e-10;
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = 0.0;
    const bool fma_warp = (MODE == 0) || (MODE == 3 && (warp & 1) == 0);
    const bool mma_warp = (MODE == 1) || (MODE == 3 && (warp & 1) == 1);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 3) {
            if (fma_warp) {
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int i = 0; i < 8; ++i) x[i] = fma(a[(i + r) & 7], y[i >> 1], x[i]);
            }
        }
        if (MODE == 1 || MODE == 3) {
            if (mma_warp) {
#pragma unroll
                for (int t = 0; t < 8; ++t) dmma884(c[2 * t], c[2 * t + 1], a[t], y[t & 3]);
            }
        }
        if (MODE == 2) {
            double af[4] = { a[0], a[1], a[2], a[3] };
            double bf[2] = { y[0], y[1] };
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                double cc[4] = { c[4 * t], c[4 * t + 1], c[4 * t + 2], c[4 * t + 3] };
                dmma1688(cc, af, bf);
                c[4 * t] = cc[0]; c[4 * t + 1] = cc[1]; c[4 * t + 2] = cc[2]; c[4 * t + 3] = cc[3];
            }
        }
        if (MODE == 4) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = fma(a[(i + r) & 7], y[i >> 1], x[i]);
                dmma884(c[2 * r], c[2 * r + 1], a[r], y[r]);
            }
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
cudaError_t time_probe(double* buf, int blocks, int iters, float* ms) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    dmma_probe_kernel<MODE><<<blocks, 256>>>(buf, 16, 0.999999, 1e-9);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        dmma_probe_kernel<MODE><<<blocks, 256>>>(buf, iters, 0.999999, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float t = 0.f;
        cudaEventElapsedTime(&t, e0, e1);
        best = std::min(best, t);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *ms = best;
    return cudaGetLastError();
}

}  // namespace

extern "C" {

// sustained DFMA rate of `device` in TFLOP/s (2 flop per FMA), best of 4 timed launches
int scvx_bench_fp64_peak(int device, double* tflops) {
    if (!tflops || cudaSetDevice(device) != cudaSuccess) return -1;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -2;
    const int blocks = prop.multiProcessorCount * 8, iters = 4096;
    double* buf = nullptr;
    if (cudaMalloc((void**)&buf, (size_t)blocks * 256 * 8) != cudaSuccess) return -2;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        fp64_peak_kernel<<<blocks, 256>>>(buf, iters, 0.999999, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double fl = 2.0 * 64.0 * (double)iters * 256.0 * (double)blocks;
        if (rep > 0) best = std::max(best, fl / (ms * 1e-3) * 1e-12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(buf);
    *tflops = best;
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// out[0] DFMA only TF | out[1] DMMA m8n8k4 only TF | out[2] DMMA m16n8k8 only TF
// out[3], out[4]  MODE 3 (alternate warps): DFMA TF and DMMA TF achieved concurrently
// out[5], out[6]  MODE 4 (same warp):       DFMA TF and DMMA TF achieved concurrently
int scvx_bench_dmma_probe(int device, double* out) {
    if (!out || cudaSetDevice(device) != cudaSuccess) return -1;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -2;
    const int blocks = prop.multiProcessorCount * 8, iters = 2048;
    double* buf = nullptr;
    if (cudaMalloc((void**)&buf, (size_t)blocks * 256 * 8) != cudaSuccess) return -2;
    const double threads = 256.0 * blocks, warps = threads / 32.0;
    float ms = 0.f;
    cudaError_t e;
    e = time_probe<0>(buf, blocks, iters, &ms); out[0] = 2.0 * 64.0 * iters * threads / (ms * 1e-3) * 1e-12;
    if (e == cudaSuccess) { e = time_probe<1>(buf, blocks, iters, &ms); out[1] = 8.0 * 512.0 * iters * warps / (ms * 1e-3) * 1e-12; }
    if (e == cudaSuccess) { e = time_probe<2>(buf, blocks, iters, &ms); out[2] = 4.0 * 2048.0 * iters * warps / (ms * 1e-3) * 1e-12; }
    if (e == cudaSuccess) {
        e = time_probe<3>(buf, blocks, iters, &ms);
        out[3] = 2.0 * 64.0 * iters * (threads / 2) / (ms * 1e-3) * 1e-12;
        out[4] = 8.0 * 512.0 * iters * (warps / 2) / (ms * 1e-3) * 1e-12;
    }
    if (e == cudaSuccess) {
        e = time_probe<4>(buf, blocks, iters, &ms);
        out[5] = 2.0 * 32.0 * iters * threads / (ms * 1e-3) * 1e-12;
        out[6] = 4.0 * 512.0 * iters * warps / (ms * 1e-3) * 1e-12;
    }
    cudaFree(buf);
    return e == cudaSuccess ? 0 : -2;
}

}  // extern "C"
