// scvx_socp_pattern.h — fixed sparsity pattern of the trajectory-dependent SOCP rows.
//
// The reference refreshes, every SCvx iteration, the K dynamics equality blocks (rocketland.jl:117-133, 251-258) and
// the K+1 linearised thrust-lower-bound rows (rocketland.jl:194-201, 260-265) through K*21 + 3(K+1) MOI.modify calls.
// The pattern of those rows never changes; only values do.  This header is the single definition of that pattern in
// compressed-sparse-column form, shared by the host (pattern arrays) and the device (value gather).
//
//   rows    0 .. 14K-1      : dynamics row i of interval n at 14n + i            (Zeros cone, rocketland.jl:131)
//           14K .. 14K+K    : thrust lower bound of node n at 14K + n            (Nonpositives, rocketland.jl:201)
//   columns, in the reference's variable creation order dxv, duv, dsig, nuv (rocketland.jl:73-76; they are consecutive
//   there, so local column j is reference variable 17(K+1) + j, 0-based):
//           dxv[j,n] -> 14n + j ; duv[j,n] -> 14(K+1) + 3n + j ; dsig -> 17(K+1) ; nuv[j,n] -> 17(K+1) + 1 + 14n + j
//   entries per interval n (0-based; D = block columns 1..21): A_n on dxv[:,n], B-_n on duv[:,n], B+_n on duv[:,n+1],
//   Sigma_n on dsig, +1 on nuv[:,n+1], -1 on dxv[:,n+1] (rocketland.jl:125-128); all 14x21 entries of D are stored,
//   structural zeros included, exactly as `eachcol(derivative)` emits them.  Per node n: H_n = -u/|u| on duv[:,n].
//   Within a column rows ascend.  nnz = 210K + (87K+3) + 14K + 14K = 325K + 3.
#pragma once

#if defined(__CUDACC__)
#define SCVX_HD __host__ __device__ __forceinline__
#else
#define SCVX_HD inline
#endif

enum { SOCP_SRC_BLOCK = 0, SOCP_SRC_PLUS1 = 1, SOCP_SRC_MINUS1 = 2, SOCP_SRC_TLB = 3 };

struct SocpEntry {
    int row, col;
    int src;       // SOCP_SRC_*
    int offset;    // SRC_BLOCK: offset into the trajectory's blocks (322 doubles per interval); SRC_TLB: into its 4 x n_nodes tlb
};

SCVX_HD int socp_nnz(int K) { return 325 * K + 3; }
SCVX_HD int socp_rows(int K) { return 15 * K + 1; }
SCVX_HD int socp_cols(int K) { return 31 * (K + 1) + 1; }

SCVX_HD int socp_blk(int n, int c, int i) { return n * 322 + c * 14 + i; }

// Decode value index p (0 <= p < socp_nnz(K)) of the CSC value array.
SCVX_HD SocpEntry socp_decode(int p, int K) {
    SocpEntry e;
    e.src = SOCP_SRC_BLOCK; e.offset = 0; e.row = 0; e.col = 0;
    // --- dxv columns: 210 K entries
    if (p < 210 * K) {
        if (p < 196) {                                   // node 0: A_0 only
            const int j = p / 14, i = p - 14 * j;
            e.col = j; e.row = i; e.offset = socp_blk(0, 1 + j, i);
            return e;
        }
        const int q = p - 196;
        if (q < 210 * (K - 1)) {                         // nodes 1..K-1: -1 of interval n-1, then A_n
            const int n = 1 + q / 210, r = q % 210, j = r / 15, t = r - 15 * j;
            e.col = 14 * n + j;
            if (t == 0) { e.row = 14 * (n - 1) + j; e.src = SOCP_SRC_MINUS1; }
            else { e.row = 14 * n + t - 1; e.offset = socp_blk(n, 1 + j, t - 1); }
            return e;
        }
        const int j = q - 210 * (K - 1);                 // node K: -1 only
        e.col = 14 * K + j; e.row = 14 * (K - 1) + j; e.src = SOCP_SRC_MINUS1;
        return e;
    }
    p -= 210 * K;
    const int cu = 14 * (K + 1);
    // --- duv columns: 87 K + 3 entries
    if (p < 87 * K + 3) {
        if (p < 45) {                                    // node 0: B-_0, H_0
            const int j = p / 15, t = p - 15 * j;
            e.col = cu + j;
            if (t < 14) { e.row = t; e.offset = socp_blk(0, 15 + j, t); }
            else { e.row = 14 * K; e.src = SOCP_SRC_TLB; e.offset = j; }
            return e;
        }
        const int q = p - 45;
        if (q < 87 * (K - 1)) {                          // nodes 1..K-1: B+_{n-1}, B-_n, H_n
            const int n = 1 + q / 87, r = q % 87, j = r / 29, t = r - 29 * j;
            e.col = cu + 3 * n + j;
            if (t < 14) { e.row = 14 * (n - 1) + t; e.offset = socp_blk(n - 1, 18 + j, t); }
            else if (t < 28) { e.row = 14 * n + t - 14; e.offset = socp_blk(n, 15 + j, t - 14); }
            else { e.row = 14 * K + n; e.src = SOCP_SRC_TLB; e.offset = 4 * n + j; }
            return e;
        }
        const int r = q - 87 * (K - 1), j = r / 15, t = r - 15 * j;   // node K: B+_{K-1}, H_K
        e.col = cu + 3 * K + j;
        if (t < 14) { e.row = 14 * (K - 1) + t; e.offset = socp_blk(K - 1, 18 + j, t); }
        else { e.row = 14 * K + K; e.src = SOCP_SRC_TLB; e.offset = 4 * K + j; }
        return e;
    }
    p -= 87 * K + 3;
    const int cs = 17 * (K + 1);
    // --- dsig column: 14 K entries
    if (p < 14 * K) {
        const int n = p / 14, i = p - 14 * n;
        e.col = cs; e.row = p; e.offset = socp_blk(n, 21, i);
        return e;
    }
    p -= 14 * K;
    // --- nuv columns of nodes 1..K: +1 (node 0 has no entry)
    e.col = cs + 1 + 14 + p; e.row = p; e.src = SOCP_SRC_PLUS1;
    return e;
}
