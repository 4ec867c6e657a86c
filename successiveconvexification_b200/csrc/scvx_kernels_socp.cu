// scvx_kernels_socp.cu — value writer for the fixed-pattern sparse SOCP rows.
//
// Pure data movement (HBM bound): per trajectory 325K+3 matrix values and 15K+1 constants are gathered from the
// linearisation outputs.  A thread owns one value index p, decodes it once (scvx_socp_pattern.h) and then walks over
// trajectories, so consecutive threads write consecutive addresses (coalesced 8-byte stores) and read runs of 14
// doubles of one block column.
#include "scvx_kernels.h"
#include "scvx_socp_pattern.h"

namespace {

constexpr int TRAJ_PER_THREAD = 8;   // a thread amortises its index decode over 8 consecutive trajectories

__global__ void __launch_bounds__(256)
socp_values_kernel(const double* __restrict__ blocks, const double* __restrict__ lin_err, const double* __restrict__ tlb,
                   int K, int B, double* __restrict__ vals, double* __restrict__ rhs) {
    const int nnz = socp_nnz(K), nr = socp_rows(K);
    const int p = blockIdx.y * blockDim.x + threadIdx.x;
    if (p >= nnz + nr) return;
    const size_t blk_stride = (size_t)K * SCVX_BLOCK_DOUBLES, tlb_stride = (size_t)4 * (K + 1), err_stride = (size_t)14 * K;
    if (p < nnz) {
        const SocpEntry e = socp_decode(p, K);
        const int b0 = blockIdx.x * TRAJ_PER_THREAD;
        double* dst = vals + (size_t)b0 * nnz + p;
        if (e.src == SOCP_SRC_BLOCK || e.src == SOCP_SRC_TLB) {
            const size_t stride = e.src == SOCP_SRC_BLOCK ? blk_stride : tlb_stride;
            const double* src = (e.src == SOCP_SRC_BLOCK ? blocks : tlb) + e.offset + (size_t)b0 * stride;
            double v[TRAJ_PER_THREAD];
            #pragma unroll
            for (int k = 0; k < TRAJ_PER_THREAD; ++k) v[k] = (b0 + k < B) ? __ldg(src + k * stride) : 0.0;   // all loads in flight
            #pragma unroll
            for (int k = 0; k < TRAJ_PER_THREAD; ++k) if (b0 + k < B) dst[(size_t)k * nnz] = v[k];
        } else {
            const double v = e.src == SOCP_SRC_PLUS1 ? 1.0 : -1.0;
            #pragma unroll
            for (int k = 0; k < TRAJ_PER_THREAD; ++k) if (b0 + k < B) dst[(size_t)k * nnz] = v;
        }
    } else if (rhs) {
        const int r = p - nnz;
        const int b0 = blockIdx.x * TRAJ_PER_THREAD;
        // lin_err_n = endpoint_n - x_{n+1} (rocketland.jl:129, 256);  h_n = Tmin - |u_n| (rocketland.jl:200, 263)
        const bool is_err = r < 14 * K;
        const size_t stride = is_err ? err_stride : tlb_stride;
        const double* src = is_err ? lin_err + r : tlb + 4 * (r - 14 * K) + 3;
        #pragma unroll
        for (int k = 0; k < TRAJ_PER_THREAD; ++k)
            if (b0 + k < B) rhs[(size_t)(b0 + k) * nr + r] = __ldg(src + (size_t)(b0 + k) * stride);
    }
}

}  // namespace

cudaError_t scvx_launch_socp_values(const double* blocks, const double* lin_err, const double* tlb, int n_nodes, int B,
                                    double* vals, double* rhs, cudaStream_t s) {
    const int K = n_nodes - 1;
    const int total = socp_nnz(K) + socp_rows(K);
    const int gx = (total + 255) / 256;
    if (gx > 65535) return cudaErrorInvalidValue;
    const int gy = (B + TRAJ_PER_THREAD - 1) / TRAJ_PER_THREAD;
    socp_values_kernel<<<dim3(gy, gx), 256, 0, s>>>(blocks, lin_err, tlb, K, B, vals, rhs);
    return cudaGetLastError();
}
