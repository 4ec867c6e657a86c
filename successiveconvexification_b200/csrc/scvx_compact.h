// scvx_compact.h — the compact result record (include/scvx_b200.h, scvx_linearize_batch_compact): the single definition
// of which entries of the 14 x 23 block are data and which are structural constants of the dynamics.
//
// Block column c (0 endpoint, 1..21 = D columns for inp 0..20, 22 z), row r (state index 0 m, 1..3 r, 4..6 v, 7..10 q,
// 11..13 w).  For the reference's right-hand side (dynamics.jl:54-77):
//   * nothing depends on position            -> D[:, r_j] = e_{r_j}                         (3 constant columns)
//   * mdot = -a |u| depends on u only         -> row m of the state columns is e_m           (D[m,m] = 1, zeros elsewhere)
//   * qdot, wdot do not depend on m or v      -> rows q, w of the columns d/dm, d/dv are 0
//   * wdot does not depend on q               -> rows w of the columns d/dq are 0
// Everything else (endpoint, the r / v rows of every state column, q rows of d/dq and d/dw, w rows of d/dw, the control
// columns B-, B+, Sigma, z) is data: 229 of 322 entries.
#pragma once

#if defined(__CUDACC__)
#define SCVX_CHD __host__ __device__ __forceinline__
#else
#define SCVX_CHD inline
#endif

// row range [lo, hi) of the data entries of block column c
SCVX_CHD void compact_rows(int c, int& lo, int& hi) {
    if (c == 0 || c >= 15) { lo = 0; hi = 14; }       // endpoint, B-, B+, Sigma, z
    else if (c == 1) { lo = 1; hi = 7; }              // d/dm: r, v rows
    else if (c <= 4) { lo = 0; hi = 0; }              // d/dr: constant e_r
    else if (c <= 7) { lo = 1; hi = 7; }              // d/dv: r, v rows
    else if (c <= 11) { lo = 1; hi = 11; }            // d/dq: r, v, q rows
    else { lo = 1; hi = 14; }                         // d/dw: r, v, q, w rows
}

// value of a structural constant of the block at (c, r) — only meaningful outside compact_rows(c)
SCVX_CHD double compact_constant(int c, int r) {
    if (c == 1) return r == 0 ? 1.0 : 0.0;            // D[m, m] = 1
    if (c >= 2 && c <= 4) return r == c - 1 ? 1.0 : 0.0;   // D[r_j, r_j] = 1
    return 0.0;
}

// fills index[0..228] with the dense offsets of the compact slots; returns the count (229)
SCVX_CHD int compact_fill_layout(int* index) {
    int n = 0;
    for (int c = 0; c < 23; ++c) {
        int lo, hi;
        compact_rows(c, lo, hi);
        for (int r = lo; r < hi; ++r) index[n++] = c * 14 + r;
    }
    return n;
}
