// scvx_kernels_basic.cu — first-generation kernels of the linearise-and-discretise path (sm_100a, FP64):
//   * linearize_dualwarp_kernel : one warp per (trajectory, interval); lane L < 21 carries the tangent
//                                 d/d inp[L] as a one-partial dual number through the whole rk4
//                                 (= Zygote.forward_jacobian of rk4, reference dynamics.jl:311-313).
//                                 Exact forward mode of the executed arithmetic by construction; kept as
//                                 the on-device cross-check of the faster STAGED kernel.
//   * predict_kernel            : one thread per interval, value only (predict_state / simulate_zygote,
//                                 dynamics.jl:308-310, 315-317).
//   * prefilter kernels         : Interpolations.jl cubic-B-spline prefilter, one thread per grid line.
#include "scvx_common.cuh"
#include "scvx_kernels.h"

// ---------------------------------------------------------------------------------------------
// thrust-lower-bound rows of one node (rocketland.jl:199-200, 261-263):  H = -u/|u| (3), h = Tmin - |u|
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void write_tlb(const ScvxBatch& bt, int b, int node) {
    const double* u = bt.U + ((size_t)b * bt.n_nodes + node) * 3;
    const double u0 = u[0], u1 = u[1], u2 = u[2];
    const double nu = sqrt(u0 * u0 + u1 * u1 + u2 * u2);
    double* o = bt.out_tlb + ((size_t)b * bt.n_nodes + node) * 4;
    o[0] = -(u0 / nu); o[1] = -(u1 / nu); o[2] = -(u2 / nu);
    o[3] = bt.P[bt.n_params == 1 ? 0 : b].Tmin - nu;
}

__global__ void __launch_bounds__(128)
linearize_dualwarp_kernel(ScvxBatch bt, ScvxTables tb) {
    const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int ni = bt.n_nodes - 1;
    const long total = (long)ni * bt.B;
    if (warp >= total) return;
    const int b = (int)(warp / ni), i = (int)(warp % ni);
    const scvx_probinfo& P = bt.P[bt.n_params == 1 ? 0 : b];
    const double* xin = bt.X + ((size_t)b * bt.n_nodes + i) * 14;
    const double* uin = bt.U + ((size_t)b * bt.n_nodes + i) * 3;

    D1 st[14], um[3], up[3];
#pragma unroll
    for (int r = 0; r < 14; ++r) st[r] = D1(xin[r], lane == r ? 1.0 : 0.0);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        um[c] = D1(uin[c], lane == 14 + c ? 1.0 : 0.0);
        up[c] = D1(uin[3 + c], lane == 17 + c ? 1.0 : 0.0);
    }
    const D1 sg(bt.sigma[b], lane == 20 ? 1.0 : 0.0);
    // this lane's own input value (for z = endpoint - D*inp)
    double my_inp = 0.0;
    if (lane < 14) my_inp = xin[lane];
    else if (lane < 20) my_inp = uin[lane - 14];
    else if (lane == 20) my_inp = bt.sigma[b];

    rk4_t<D1>(P, tb, st, um, up, sg, bt.dt, bt.npts, bt.mode);

    double* blk = bt.out_blocks + (size_t)warp * SCVX_BLOCK_DOUBLES;
#pragma unroll
    for (int r = 0; r < 14; ++r) {
        if (lane < 21) blk[14 * (1 + lane) + r] = st[r].d;
        double t = st[r].d * my_inp;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) {
            blk[r] = st[r].v;
            blk[14 * 22 + r] = st[r].v - t;
            if (bt.out_lin_err) bt.out_lin_err[(size_t)warp * 14 + r] = st[r].v - xin[14 + r];
        }
    }
    if (bt.out_tlb) {
        if (lane == 31) write_tlb(bt, b, i);
        if (lane == 30 && i == ni - 1) write_tlb(bt, b, i + 1);
    }
}

// SURVEY.md §8f-4 variant (fin forces + aero torque): control_dim = 5, inp = [x(14); u_k(5); u_{k+1}(5); sigma] (25),
// lane L < 25 carries d/d inp[L]; block 14 x 27 = [endpoint | D (25) | z].  bt.U is 5 x n_nodes x B here.
__global__ void __launch_bounds__(128)
linearize_dualwarp_fins_kernel(ScvxBatch bt, ScvxTables tb) {
    const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int ni = bt.n_nodes - 1;
    const long total = (long)ni * bt.B;
    if (warp >= total) return;
    const int b = (int)(warp / ni), i = (int)(warp % ni);
    const scvx_probinfo& P = bt.P[bt.n_params == 1 ? 0 : b];
    const double* xin = bt.X + ((size_t)b * bt.n_nodes + i) * 14;
    const double* uin = bt.U + ((size_t)b * bt.n_nodes + i) * 5;
    D1 st[14], um[5], up[5];
#pragma unroll
    for (int r = 0; r < 14; ++r) st[r] = D1(xin[r], lane == r ? 1.0 : 0.0);
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        um[c] = D1(uin[c], lane == 14 + c ? 1.0 : 0.0);
        up[c] = D1(uin[5 + c], lane == 19 + c ? 1.0 : 0.0);
    }
    const D1 sg(bt.sigma[b], lane == 24 ? 1.0 : 0.0);
    double my_inp = 0.0;
    if (lane < 14) my_inp = xin[lane];
    else if (lane < 24) my_inp = uin[lane - 14];
    else if (lane == 24) my_inp = bt.sigma[b];
    rk4_fins_t<D1>(P, tb, st, um, up, sg, bt.dt, bt.npts, bt.mode);
    double* blk = bt.out_blocks + (size_t)warp * (14 * 27);
#pragma unroll
    for (int r = 0; r < 14; ++r) {
        if (lane < 25) blk[14 * (1 + lane) + r] = st[r].d;
        double t = st[r].d * my_inp;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) {
            blk[r] = st[r].v;
            blk[14 * 26 + r] = st[r].v - t;
            if (bt.out_lin_err) bt.out_lin_err[(size_t)warp * 14 + r] = st[r].v - xin[14 + r];
        }
    }
}

__global__ void __launch_bounds__(128)
predict_kernel(ScvxBatch bt, ScvxTables tb) {
    const long w = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int ni = bt.n_nodes - 1;
    const long total = (long)ni * bt.B;
    if (w >= total) return;
    const int b = (int)(w / ni), i = (int)(w % ni);
    const scvx_probinfo& P = bt.P[bt.n_params == 1 ? 0 : b];
    const double* xin = bt.X + ((size_t)b * bt.n_nodes + i) * 14;
    const double* uin = bt.U + ((size_t)b * bt.n_nodes + i) * 3;
    double st[14], um[3], up[3];
#pragma unroll
    for (int r = 0; r < 14; ++r) st[r] = xin[r];
#pragma unroll
    for (int c = 0; c < 3; ++c) { um[c] = uin[c]; up[c] = uin[3 + c]; }
    rk4_t<double>(P, tb, st, um, up, bt.sigma[b], bt.dt, bt.npts, bt.mode);
    double* o = bt.out_endpoints + (size_t)w * 14;
#pragma unroll
    for (int r = 0; r < 14; ++r) o[r] = st[r];
}

// ---------------------------------------------------------------------------------------------
// Cubic B-spline prefilter, BSpline(Cubic(Line(OnGrid()))) (Interpolations.jl; call sites
// aerodynamics.jl:19-21).  Per line of n samples d_1..d_n solve for c_0..c_{n+1}:
//     c_{k-1}/6 + 2 c_k/3 + c_{k+1}/6 = d_k  (k=1..n),   c_0 - 2c_1 + c_2 = 0,   c_{n-1} - 2c_n + c_{n+1} = 0.
// Subtracting the boundary rows from rows 1 and n gives c_1 = d_1 and c_n = d_n exactly; the interior
// unknowns c_2..c_{n-1} then satisfy the diagonally dominant tridiagonal system [1 4 1] c = 6 d, solved
// by the Thomas algorithm (its elimination factors `cp` are data independent and precomputed once).
// ---------------------------------------------------------------------------------------------
__device__ void prefilter_line(const double* __restrict__ in, int n, size_t sin, double* __restrict__ out,
                               size_t sout, const double* __restrict__ cp) {
    // out index g = 0..n+1 ; in index k-1 for grid point k
    const double d1 = in[0], dn = in[(size_t)(n - 1) * sin];
    out[1 * sout] = d1;
    out[(size_t)n * sout] = dn;
    if (n >= 3) {
        // forward sweep over k = 2..n-1 (m = k-2 = 0..n-3); modified rhs stored in place
        double prev = 0.0;
        for (int k = 2; k <= n - 1; ++k) {
            double rhs = 6.0 * in[(size_t)(k - 1) * sin];
            if (k == 2) rhs -= d1;
            if (k == n - 1) rhs -= dn;
            const double denom_inv = cp[k - 2];                 // 1 / (4 - cp'[m-1]) precomputed
            prev = (rhs - prev) * denom_inv;
            out[(size_t)k * sout] = prev;
        }
        // back substitution: c_k = d'_k - cp[m] * c_{k+1}
        double next = 0.0;
        for (int k = n - 1; k >= 2; --k) {
            double v = out[(size_t)k * sout];
            if (k < n - 1) v -= cp[k - 2] * next;
            out[(size_t)k * sout] = v;
            next = v;
        }
    }
    out[0] = 2.0 * out[1 * sout] - out[2 * sout];
    out[(size_t)(n + 1) * sout] = 2.0 * out[(size_t)n * sout] - out[(size_t)(n - 1) * sout];
}

// pass 1: along axis 0 (length n1) for each of n2 columns: samples (n1 x n2) -> tmp ((n1+2) x n2)
__global__ void prefilter_axis0_kernel(const double* samples, int n1, int n2, double* tmp, const double* cp) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n2) return;
    prefilter_line(samples + (size_t)j * n1, n1, 1, tmp + (size_t)j * (n1 + 2), 1, cp);
}
// pass 2: along axis 1 (length n2) for each of n1+2 rows: tmp ((n1+2) x n2) -> coef ((n1+2) x (n2+2))
__global__ void prefilter_axis1_kernel(const double* tmp, int n1, int n2, double* coef, const double* cp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n1 + 2) return;
    prefilter_line(tmp + i, n2, (size_t)(n1 + 2), coef + i, (size_t)(n1 + 2), cp);
}

// ---------------------------------------------------------------------------------------------
// Fused cost / defect evaluation (rocketland.jl:289-290): one warp per trajectory reduces its K defect vectors.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) defect_cost_kernel(const double* __restrict__ X, const double* __restrict__ lin_err,
                                                          int n_nodes, int B, double wNu, double* __restrict__ out_defect,
                                                          double* __restrict__ out_cost) {
    const int b = (int)(((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    const int n = (n_nodes - 1) * 14;
    const double* e = lin_err + (size_t)b * n;
    double acc = 0.0;
    for (int k = lane; k < n; k += 32) { const double v = e[k]; acc = fma(v, v, acc); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
        const double d = sqrt(acc);
        out_defect[b] = d;
        if (out_cost) out_cost[b] = fma(wNu, d, -X[((size_t)b * n_nodes + (n_nodes - 1)) * 14]);
    }
}

// ---------------------------------------------------------------------------------------------
// Batched initial guess: linear_points (initial_solve.jl:113-129), one thread per (trajectory, node).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) linear_points_kernel(const double* __restrict__ rIi, const double* __restrict__ vIi,
                                                            const double* __restrict__ mwet, double mwet_shared, double mdry,
                                                            double rf0, double rf1, double rf2, double vf0, double vf1,
                                                            double vf2, double g, int K, int B, double* __restrict__ X,
                                                            double* __restrict__ U) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int n = K + 1;
    if (t >= (long)n * B) return;
    const int b = (int)(t / n), k = (int)(t - (long)b * n);
    const double wa = (double)(K - k) / (double)K, wb = (double)k / (double)K;       // (K-k)/K and k/K, as the reference
    const double mw = mwet ? mwet[b] : mwet_shared;
    const double mk = wa * mw + wb * mdry;
    const double rf[3] = { rf0, rf1, rf2 }, vf[3] = { vf0, vf1, vf2 };
    double r[3], v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { r[c] = wa * rIi[(size_t)b * 3 + c] + wb * rf[c]; v[c] = wa * vIi[(size_t)b * 3 + c] + wb * vf[c]; }
    // rotation_between(u = [1,0,0], w = -v): q = normalize([ |u||w| + u.w ; u x w ])
    const double wv[3] = { -v[0], -v[1], -v[2] };
    const double normprod = sqrt(wv[0] * wv[0] + wv[1] * wv[1] + wv[2] * wv[2]);
    double qw = normprod + wv[0];
    double ax = 0.0, ay = -wv[2], az = wv[1];
    if (fabs(qw) < 100.0 * 2.220446049250313e-16) { ax = 0.0; ay = 0.0; az = 1.0; }    // antiparallel: any axis perpendicular to u
    const double qn = 1.0 / sqrt(qw * qw + ax * ax + ay * ay + az * az);
    double* x = X + (size_t)t * 14;
    x[0] = mk; x[1] = r[0]; x[2] = r[1]; x[3] = r[2]; x[4] = v[0]; x[5] = v[1]; x[6] = v[2];
    x[7] = qw * qn; x[8] = ax * qn; x[9] = ay * qn; x[10] = az * qn; x[11] = 0.0; x[12] = 0.0; x[13] = 0.0;
    double* u = U + (size_t)t * 3;
    u[0] = mk * g; u[1] = 0.0; u[2] = 0.0;
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
cudaError_t scvx_launch_dualwarp(const ScvxBatch& bt, const ScvxTables& tb, cudaStream_t s) {
    const long total = (long)(bt.n_nodes - 1) * bt.B;
    if (total <= 0) return cudaSuccess;
    const long blocks = (total + 3) / 4;
    linearize_dualwarp_kernel<<<(unsigned)blocks, 128, 0, s>>>(bt, tb);
    return cudaGetLastError();
}

// fin-force table lookup: both splines at n (mach, deflection) pairs
__global__ void __launch_bounds__(128) fin_force_kernel(ScvxTables lift_tb, ScvxTables drag_tb, const double* __restrict__ mach,
                                                        const double* __restrict__ defl, int n, double* __restrict__ out_lift,
                                                        double* __restrict__ out_drag) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const double m = mach[t], d = defl[t];
    if (out_lift) out_lift[t] = spline_eval<double>(lift_tb.drag, lift_tb, m, d);
    if (out_drag) out_drag[t] = spline_eval<double>(drag_tb.drag, drag_tb, m, d);
}

cudaError_t scvx_launch_fin_force(const ScvxTables& lift_tb, const ScvxTables& drag_tb, const double* mach, const double* defl,
                                  int n, double* out_lift, double* out_drag, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    fin_force_kernel<<<(n + 127) / 128, 128, 0, s>>>(lift_tb, drag_tb, mach, defl, n, out_lift, out_drag);
    return cudaGetLastError();
}

cudaError_t scvx_launch_dualwarp_fins(const ScvxBatch& bt, const ScvxTables& tb, cudaStream_t s) {
    const long total = (long)(bt.n_nodes - 1) * bt.B;
    if (total <= 0) return cudaSuccess;
    linearize_dualwarp_fins_kernel<<<(unsigned)((total + 3) / 4), 128, 0, s>>>(bt, tb);
    return cudaGetLastError();
}

cudaError_t scvx_launch_predict(const ScvxBatch& bt, const ScvxTables& tb, cudaStream_t s) {
    const long total = (long)(bt.n_nodes - 1) * bt.B;
    if (total <= 0) return cudaSuccess;
    predict_kernel<<<(unsigned)((total + 127) / 128), 128, 0, s>>>(bt, tb);
    return cudaGetLastError();
}

cudaError_t scvx_launch_prefilter(const double* d_samples, int n1, int n2, double* d_tmp, double* d_coef,
                                  const double* d_cp, cudaStream_t s) {
    prefilter_axis0_kernel<<<(n2 + 63) / 64, 64, 0, s>>>(d_samples, n1, n2, d_tmp, d_cp);
    prefilter_axis1_kernel<<<(n1 + 2 + 63) / 64, 64, 0, s>>>(d_tmp, n1, n2, d_coef, d_cp);
    return cudaGetLastError();
}

cudaError_t scvx_launch_defect_cost(const double* X, const double* lin_err, int n_nodes, int B, double wNu,
                                    double* out_defect, double* out_cost, cudaStream_t s) {
    if (B <= 0) return cudaSuccess;
    const long threads = (long)B * 32;
    defect_cost_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(X, lin_err, n_nodes, B, wNu, out_defect, out_cost);
    return cudaGetLastError();
}

cudaError_t scvx_launch_linear_points(const double* rIi, const double* vIi, const double* mwet, double mwet_shared, double mdry,
                                      const double* rIf, const double* vIf, double g, int K, int B, double* X, double* U,
                                      cudaStream_t s) {
    const long threads = (long)(K + 1) * B;
    if (threads <= 0) return cudaSuccess;
    linear_points_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(rIi, vIi, mwet, mwet_shared, mdry, rIf[0], rIf[1],
                                                                          rIf[2], vIf[0], vIf[1], vIf[2], g, K, B, X, U);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Dispersion set-up: per-trajectory normalize_problem (sample_problems.jl:5-23) + ProbInfo (master.jl:73-83) +
// linear_points (initial_solve.jl:113-129), one thread per (trajectory, node).  Every scaling is written as the
// reference writes it (multiply by a reciprocal where it broadcasts `*`, divide where it divides) so the results match
// a host evaluation of those lines to the last bit wherever IEEE arithmetic is deterministic.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void inverse3x3_lu(const double A[9], double Ai[9]) {
    // LU with partial pivoting, then three unit right-hand sides (what `inv` does); column-major
    double a[3][3];
    int piv[3] = { 0, 1, 2 };
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) a[r][c] = A[r + 3 * c];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        int p = k;
#pragma unroll
        for (int r = k + 1; r < 3; ++r) if (fabs(a[r][k]) > fabs(a[p][k])) p = r;
        if (p != k) {
#pragma unroll
            for (int c = 0; c < 3; ++c) { const double t = a[k][c]; a[k][c] = a[p][c]; a[p][c] = t; }
            const int t = piv[k]; piv[k] = piv[p]; piv[p] = t;
        }
#pragma unroll
        for (int r = k + 1; r < 3; ++r) {
            a[r][k] /= a[k][k];
#pragma unroll
            for (int c = k + 1; c < 3; ++c) a[r][c] -= a[r][k] * a[k][c];
        }
    }
#pragma unroll
    for (int col = 0; col < 3; ++col) {
        double y[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) y[r] = (piv[r] == col) ? 1.0 : 0.0;
        y[1] -= a[1][0] * y[0];
        y[2] -= a[2][0] * y[0] + a[2][1] * y[1];
        y[2] = y[2] / a[2][2];
        y[1] = (y[1] - a[1][2] * y[2]) / a[1][1];
        y[0] = (y[0] - a[0][1] * y[1] - a[0][2] * y[2]) / a[0][0];
#pragma unroll
        for (int r = 0; r < 3; ++r) Ai[r + 3 * col] = y[r];
    }
}

__global__ void __launch_bounds__(128) dispersed_setup_kernel(scvx_dim_problem base, const double* __restrict__ rIi,
                                                              const double* __restrict__ vIi, const double* __restrict__ mwet,
                                                              int B, double* __restrict__ X, double* __restrict__ U,
                                                              double* __restrict__ sigma, double* __restrict__ scales,
                                                              scvx_probinfo* __restrict__ P0, scvx_probinfo* __restrict__ P1) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int K = base.K, n = K + 1;
    if (t >= (long)n * B) return;
    const int b = (int)(t / n), k = (int)(t - (long)b * n);
    const double ri[3] = { rIi[(size_t)b * 3], rIi[(size_t)b * 3 + 1], rIi[(size_t)b * 3 + 2] };
    const double vi[3] = { vIi[(size_t)b * 3], vIi[(size_t)b * 3 + 1], vIi[(size_t)b * 3 + 2] };
    const double Ul = fmax(fmax(ri[0], ri[1]), ri[2]);                   // sample_problems.jl:6
    const double Ut = base.tf_guess;                                     // :7
    const double Um = mwet ? mwet[b] : base.mwet;                        // :8
    const double acc = Ul / (Ut * Ut);                                   // Ul/Ut^2
    const double g = base.g / acc;                                       // :10
    const double mdry = base.mdry / Um, mw = Um / Um;
    const double il = 1.0 / Ul, ivel = 1.0 / (Ul / Ut);
    // linear_points of the normalised problem (vIf is the normalised vIi, sample_problems.jl:15)
    const double wa = (double)(K - k) / (double)K, wb = (double)k / (double)K;
    const double mk = wa * mw + wb * mdry;
    double r[3], v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        r[c] = wa * (ri[c] * il) + wb * (base.rIf[c] * il);
        v[c] = wa * (vi[c] * ivel) + wb * (vi[c] * ivel);
    }
    const double wv[3] = { -v[0], -v[1], -v[2] };
    const double normprod = sqrt(wv[0] * wv[0] + wv[1] * wv[1] + wv[2] * wv[2]);
    double qw = normprod + wv[0];
    double ax = 0.0, ay = -wv[2], az = wv[1];
    if (fabs(qw) < 100.0 * 2.220446049250313e-16) { ax = 0.0; ay = 0.0; az = 1.0; }
    const double qn = 1.0 / sqrt(qw * qw + ax * ax + ay * ay + az * az);
    double* x = X + (size_t)t * 14;
    x[0] = mk; x[1] = r[0]; x[2] = r[1]; x[3] = r[2]; x[4] = v[0]; x[5] = v[1]; x[6] = v[2];
    x[7] = qw * qn; x[8] = ax * qn; x[9] = ay * qn; x[10] = az * qn; x[11] = 0.0; x[12] = 0.0; x[13] = 0.0;
    double* u = U + (size_t)t * 3;
    u[0] = mk * g; u[1] = 0.0; u[2] = 0.0;
    if (k != 0) return;
    sigma[b] = base.tf_guess / Ut;                                       // :20
    if (scales) { scales[(size_t)b * 3] = Ul; scales[(size_t)b * 3 + 1] = Ut; scales[(size_t)b * 3 + 2] = Um; }
    if (!P0 && !P1) return;
    scvx_probinfo p;
    p.a = base.alpha / (Ut * Ut / Ul);                                   // :19
    p.g0 = g;
    p.sos = base.sos / (Ul / Ut);                                        // :22
    const double ij = 1.0 / (Um * (Ul * Ul));                            // :13
#pragma unroll
    for (int e = 0; e < 9; ++e) p.jB[e] = base.jB[e] * ij;
    inverse3x3_lu(p.jB, p.jBi);                                          // master.jl:82
#pragma unroll
    for (int e = 0; e < 3; ++e) { p.rTB[e] = base.rTB[e] * il; p.rFB[e] = base.rFB[e] * (1.0 / Ut); }   // :14, :17 (1/Ut as there)
    p.force_scalar = 1.0 / (Ul * Um / (Ut * Ut));                        // aerodynamics.jl:31
    p.length_scalar = 1.0 / Ul;
    p.Tmin = base.Tmin / (Um * Ul / (Ut * Ut));                          // :11
    p.aero_kind = base.aero_kind; p._pad = 0;
    if (P0) P0[b] = p;
    if (P1) P1[b] = p;
}

cudaError_t scvx_launch_dispersed_setup(const scvx_dim_problem& base, const double* rIi, const double* vIi, const double* mwet,
                                        int B, double* X, double* U, double* sigma, double* scales, scvx_probinfo* P0,
                                        scvx_probinfo* P1, cudaStream_t s) {
    const long threads = (long)(base.K + 1) * B;
    if (threads <= 0) return cudaSuccess;
    dispersed_setup_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(base, rIi, vIi, mwet, B, X, U, sigma, scales, P0, P1);
    return cudaGetLastError();
}
