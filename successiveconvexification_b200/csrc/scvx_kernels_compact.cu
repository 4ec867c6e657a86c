// scvx_kernels_compact.cu — compact result records (scvx_linearize_batch_compact, include/scvx_b200.h).
//
// The host path is PCIe bound (2 576 B of block per interval at ~55 GB/s), and 92 of the 322 entries of a block are
// structural constants (scvx_compact.h).  The pack kernel gathers the 229 data entries of every block into a 230-double
// record (slot 229 = per-interval non-finite flag) so that only data crosses the bus; the host expander
// (scvx_expand_compact in scvx_api.cu) restores the dense block bit for bit.  Pure data movement, HBM bound:
// 2 576 B read + 1 840 B written per interval; one warp per interval, persistent blocks, coalesced 8-byte stores.
#include "scvx_kernels.h"
#include "scvx_compact.h"

namespace {

// n_data = 229 (all data entries) or 215 (without the z column, which is the last run of slots); record = n_data + 1.
__global__ void __launch_bounds__(256) compact_pack_kernel(const double* __restrict__ blocks, long n_int, int n_data,
                                                           double* __restrict__ out) {
    __shared__ short src[SCVX_COMPACT_DATA + 3];
    if (threadIdx.x < 23) {                      // one thread per block column fills its run of slots
        const int c = threadIdx.x;
        int n = 0;
        for (int cc = 0; cc < c; ++cc) { int lo, hi; compact_rows(cc, lo, hi); n += hi - lo; }
        int lo, hi;
        compact_rows(c, lo, hi);
        for (int r = lo; r < hi; ++r) src[n++] = (short)(c * 14 + r);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (long itv = (long)blockIdx.x * 8 + warp; itv < n_int; itv += (long)gridDim.x * 8) {
        const double* blk = blocks + itv * SCVX_BLOCK_DOUBLES;
        double* rec = out + itv * (n_data + 1);
        bool bad = false;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int slot = lane + 32 * k;
            if (slot < n_data) {
                const double v = __ldg(blk + src[slot]);
                bad |= !isfinite(v);
                rec[slot] = v;
            }
        }
        // the constants of a block are right whenever its Jacobian blocks were finite; a NaN/Inf there shows in the data too,
        // but the flag is cheap to make exact: scan the 93 non-data entries as well (same sectors, already in L1)
        for (int o = lane; o < SCVX_BLOCK_DOUBLES; o += 32) bad |= !isfinite(__ldg(blk + o));
        bad = __any_sync(0xffffffffu, bad);
        if (lane == 0) rec[n_data] = bad ? 1.0 : 0.0;
    }
}

}  // namespace

cudaError_t scvx_launch_compact_pack(const double* blocks, long n_intervals, int record_doubles, double* out, int sm_count,
                                     cudaStream_t s) {
    if (n_intervals <= 0) return cudaSuccess;
    long grid = (n_intervals + 7) / 8;
    const long cap = (long)sm_count * 8;         // 8 resident blocks of 256 threads per SM
    if (grid > cap) grid = cap;
    compact_pack_kernel<<<(unsigned)grid, 256, 0, s>>>(blocks, n_intervals, record_doubles - 1, out);
    return cudaGetLastError();
}
