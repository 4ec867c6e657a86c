// scvx_kernels_staged.cu — the STAGED linearise-and-discretise path (sm_100a, FP64).
//
// The exact forward-mode Jacobian of the reference's rk4 (dynamics.jl:112-134, 311-313) is the tangent
// recursion  K_s = J_x(Y_s) * Yt_s + J_u(Y_s) * U_s + e_sigma f(Y_s)  over the 4*npts stages.
// The value trajectory does not depend on the tangents, so the work is split in two kernels:
//
//  A  stage_value_kernel : one THREAD per interval.  Integrates the 14-state value with rk4, writes the
//     endpoint (block column 0), lin_err, the thrust-lower-bound rows, and a 25-double "stage record"
//     (stage state m,v,q,w; stage control u; unscaled rhs f) per stage, laid out
//     [group of 32 intervals][stage][entry][32 lanes] so that one stage of one group is one contiguous
//     6400-byte slab.
//
//  B  tangent_kernel : persistent, one 256-thread CTA per SM, 32 intervals in flight per CTA.
//     * Tangent propagation ("consumer" role): 8 lanes per interval, every lane owns two full tangent
//       columns (rows m,v,q,w + the r rows as pure quadrature) and one light column (inputs m, v), all in
//       registers.  Structure used: nothing depends on r; m, q, w rows of the m/v columns vanish.
//     * Jacobian production ("producer" role): the 8 warps take turns (stage T -> warp T mod 8); the producing
//       warp works with lane = interval (no redundancy), pulls its stage record with one TMA bulk copy
//       (cp.async.bulk -> mbarrier), forms the sigma-scaled Jacobian blocks (78 doubles per interval) and
//       stores them into a shared-memory ring; consumers read them back as broadcast 128-bit loads.
//     * Ring slots are handed over with mbarriers (full/empty), so a warp that is busy producing does not
//       stall the others until they are a whole ring ahead.
//     Outputs [A|B-|B+|Sigma|z] go straight from registers to the 14x23 block with 16-byte stores.
#include "scvx_common.cuh"
#include "scvx_kernels.h"

namespace {

constexpr int REC_EXO = 25;      // stage record entries: m, v(3), q(4), w(3), u(3), f_m, f_v(3), f_q(4), f_w(3)
constexpr int REC_AERO = 43;     // + dF_aero/dv (9, row-major) + dF_aero/db (9), b = C(q) e1
constexpr int REC_MAX = REC_AERO;
constexpr int NJ = 78;           // Jacobian record entries per interval per stage (2 x odd: conflict-free STS.128)
#ifndef SCVX_RING
#define SCVX_RING 6
#endif
#ifndef SCVX_LOOKAHEAD
#define SCVX_LOOKAHEAD 3
#endif
constexpr int RING = SCVX_RING;  // ring slots
constexpr int LOOKAHEAD = SCVX_LOOKAHEAD;   // producer runs this many stages ahead of the consumers (< RING)
constexpr int GROUP = 32;        // intervals per CTA pass
constexpr int NWARP = 8;

// Jacobian record layout (doubles)
constexpr int J_WW = 0;          // 9  sigma * d(wdot)/dw, row-major
constexpr int J_HW = 9;          // 3  sigma*w/2
constexpr int J_HQ = 12;         // 4  sigma*q/2
constexpr int J_V = 16;          // 3 rows x 8: [d(vdot_r)/dm, d(vdot_r)/dv (3), d(vdot_r)/dq (4)], all times sigma
constexpr int J_G = 40;          // 4 columns (u0,u1,u2,f) x 7 rows (m, v0..2, w0..2)
constexpr int J_FRQ = 68;        // 7  f_r (= v) and f_q of the unscaled rhs (sigma column only)
constexpr int J_SIG = 75;        // 1  sigma

struct StagedArgs {
    ScvxBatch bt;
    ScvxTables tb;
    double* rec;                 // stage records of this chunk
    int rec_n;                   // entries per stage record (REC_EXO or REC_AERO)
    int first;                   // first interval (global index) of this chunk (total intervals < 2^31)
    int count;                   // intervals in this chunk
    int n_groups;                // ceil(count / 32)
};

// ------------------------------------------------------------------------------------------------
// Kernel A: value trajectory + stage records.
// ------------------------------------------------------------------------------------------------
// spline value and gradient w.r.t. the physical coordinates (gradient 0 when strictly outside: Flat extrapolation);
// one pass over the 16 coefficients serves all three.
__device__ __forceinline__ void spline_val_grad(const double* __restrict__ coef, const ScvxTables& t, double x, double y,
                                                double& val, double& gx, double& gy) {
    const int L1 = t.n1 + 2;
    double xi = (x - t.x0) * t.inv_dx + 1.0, yi = (y - t.y0) * t.inv_dy + 1.0;
    double sx = t.inv_dx, sy = t.inv_dy;
    if (xi > (double)t.n1) { xi = (double)t.n1; sx = 0.0; } else if (xi < 1.0) { xi = 1.0; sx = 0.0; }
    if (yi > (double)t.n2) { yi = (double)t.n2; sy = 0.0; } else if (yi < 1.0) { yi = 1.0; sy = 0.0; }
    int i = (int)floor(xi); i = max(min(i, t.n1 - 1), 1);
    int j = (int)floor(yi); j = max(min(j, t.n2 - 1), 1);
    const double dx = xi - (double)i, dy = yi - (double)j, ox = 1.0 - dx, oy = 1.0 - dy;
    const double wx[4] = { ox * ox * ox * (1.0 / 6.0), (2.0 / 3.0) - dx * dx + 0.5 * dx * dx * dx,
                           (2.0 / 3.0) - ox * ox + 0.5 * ox * ox * ox, dx * dx * dx * (1.0 / 6.0) };
    const double gxw[4] = { -0.5 * ox * ox, -2.0 * dx + 1.5 * dx * dx, 2.0 * ox - 1.5 * ox * ox, 0.5 * dx * dx };
    const double wy[4] = { oy * oy * oy * (1.0 / 6.0), (2.0 / 3.0) - dy * dy + 0.5 * dy * dy * dy,
                           (2.0 / 3.0) - oy * oy + 0.5 * oy * oy * oy, dy * dy * dy * (1.0 / 6.0) };
    const double gyw[4] = { -0.5 * oy * oy, -2.0 * dy + 1.5 * dy * dy, 2.0 * oy - 1.5 * oy * oy, 0.5 * dy * dy };
    const double* base = coef + (i - 1) + (size_t)(j - 1) * L1;
    double av = 0.0, ax = 0.0, ay = 0.0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const double* p = base + (size_t)b * L1;
        const double c0 = __ldg(p), c1 = __ldg(p + 1), c2 = __ldg(p + 2), c3 = __ldg(p + 3);
        const double rv = wx[0] * c0 + wx[1] * c1 + wx[2] * c2 + wx[3] * c3;
        const double rg = gxw[0] * c0 + gxw[1] * c1 + gxw[2] * c2 + gxw[3] * c3;
        av = fma(wy[b], rv, av);
        ax = fma(wy[b], rg, ax);
        ay = fma(gyw[b], rv, ay);
    }
    val = av; gx = ax * sx; gy = ay * sy;
}

// Jacobian of the aerodynamic force F(b, v) (aerodynamics.jl:38-58) w.r.t. v and b = C(q) e1: exact derivative
// of the executed branch (|dp| >= 0.95 drag only; clamp active only strictly outside [-1,1]).
__device__ __forceinline__ void aero_force_jac(const scvx_probinfo& P, const ScvxTables& tb, const double b[3],
                                               const double v[3], double F[3], double Fv[3][3], double Fb[3][3]) {
    const double vv = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    const double nv = sqrt(vv), inv = 1.0 / nv;
    const double vh[3] = { v[0] * inv, v[1] * inv, v[2] * inv };
    const double bvdot = b[0] * v[0] + b[1] * v[1] + b[2] * v[2];
    const double dp = bvdot * inv;
    const double nb = sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]), inb = 1.0 / nb;
    const double car = dp * inb;
    double ca = car, mc = 1.0;
    if (car > 1.0) { ca = 1.0; mc = 0.0; } else if (car < -1.0) { ca = -1.0; mc = 0.0; }
    const double mach = nv * (1.0 / __ldg(&P.sos));
    // d(ca)/dv, d(ca)/db ; d(mach)/dv
    double cav[3], cab[3], mv[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        cav[k] = mc * (b[k] - dp * vh[k]) * inv * inb;
        cab[k] = mc * (vh[k] - car * b[k] * inb) * inb;
        mv[k] = vh[k] * (1.0 / __ldg(&P.sos));
    }
    const double fs = __ldg(&P.force_scalar);
    double drag, gx, gy;
    spline_val_grad(tb.drag, tb, ca, mach, drag, gx, gy);
    drag *= fs; gx *= fs; gy *= fs;
    double dv[3], db[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { dv[k] = gx * cav[k] + gy * mv[k]; db[k] = gx * cab[k]; }
    const double dn = drag * inv;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        F[r] = dn * v[r];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            Fv[r][c] = vh[r] * dv[c] + dn * ((r == c ? 1.0 : 0.0) - vh[r] * vh[c]);
            Fb[r][c] = vh[r] * db[c];
        }
    }
    if (fabs(dp) >= 0.95) return;
    double lift;
    spline_val_grad(tb.lift, tb, ca, mach, lift, gx, gy);
    lift *= fs; gx *= fs; gy *= fs;
    double lv[3], lb[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { lv[k] = gx * cav[k] + gy * mv[k]; lb[k] = gx * cab[k]; }
    // l = (-(v x b)) x v = v (v.b) - b (v.v)
    const double l[3] = { v[0] * bvdot - b[0] * vv, v[1] * bvdot - b[1] * vv, v[2] * bvdot - b[2] * vv };
    const double nl = sqrt(l[0] * l[0] + l[1] * l[1] + l[2] * l[2]), inl = 1.0 / nl;
    const double lh[3] = { l[0] * inl, l[1] * inl, l[2] * inl };
    // dl/dv = (v.b) I + v b^T - 2 b v^T ;  dl/db = v v^T - (v.v) I
    double Lv[3][3], Lb[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            Lv[r][c] = (r == c ? bvdot : 0.0) + v[r] * b[c] - 2.0 * b[r] * v[c];
            Lb[r][c] = v[r] * v[c] - (r == c ? vv : 0.0);
        }
    const double ln = lift * inl;
#pragma unroll
    for (int r = 0; r < 3; ++r) F[r] = fma(ln, l[r], F[r]);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        // (I - lh lh^T) * L[:,c]
        const double pv = lh[0] * Lv[0][c] + lh[1] * Lv[1][c] + lh[2] * Lv[2][c];
        const double pb = lh[0] * Lb[0][c] + lh[1] * Lb[1][c] + lh[2] * Lb[2][c];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            Fv[r][c] += lh[r] * lv[c] + ln * (Lv[r][c] - lh[r] * pv);
            Fb[r][c] += lh[r] * lb[c] + ln * (Lb[r][c] - lh[r] * pb);
        }
    }
}

// unscaled f(x,u) (dx_static without the `.* mult`, dynamics.jl:54-77) together with the Jacobians of the
// aerodynamic force w.r.t. v and b = C(q) e1 (zero for the exo-atmospheric variant).
__device__ __forceinline__ void rhs_value(const scvx_probinfo& P, const ScvxTables& tb, const double x[14],
                                          const double u[3], double f[14], double Fv[3][3], double Fb[3][3]) {
    const double q0 = x[7], q1 = x[8], q2 = x[9], q3 = x[10];
    const double w0 = x[11], w1 = x[12], w2 = x[13];
    const double p1 = q1 * q2, p2 = q0 * q3, p3 = q1 * q3, p4 = q0 * q2, p5 = q2 * q3, p6 = q0 * q1;
    const double c00 = 1.0 - 2.0 * (q2 * q2 + q3 * q3), c01 = 2.0 * (p1 - p2), c02 = 2.0 * (p3 + p4);
    const double c10 = 2.0 * (p1 + p2), c11 = 1.0 - 2.0 * (q1 * q1 + q3 * q3), c12 = 2.0 * (p5 - p6);
    const double c20 = 2.0 * (p3 - p4), c21 = 2.0 * (p5 + p6), c22 = 1.0 - 2.0 * (q1 * q1 + q2 * q2);
    double F[3] = { 0.0, 0.0, 0.0 };
    if (__ldg(&P.aero_kind) == SCVX_AERO_TABLE) {
        const double bv[3] = { c00, c10, c20 };
        aero_force_jac(P, tb, bv, x + 4, F, Fv, Fb);
    } else {
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) { Fv[r][c] = 0.0; Fb[r][c] = 0.0; }
    }
    const double im = 1.0 / x[0];
    f[0] = -__ldg(&P.a) * sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    f[1] = x[4]; f[2] = x[5]; f[3] = x[6];
    f[4] = (c00 * u[0] + c01 * u[1] + c02 * u[2] + F[0]) * im - __ldg(&P.g0);
    f[5] = (c10 * u[0] + c11 * u[1] + c12 * u[2] + F[1]) * im;
    f[6] = (c20 * u[0] + c21 * u[1] + c22 * u[2] + F[2]) * im;
    f[7]  = 0.5 * (-(w0 * q1) - w1 * q2 - w2 * q3);
    f[8]  = 0.5 * (w0 * q0 + w2 * q2 - w1 * q3);
    f[9]  = 0.5 * (w1 * q0 - w2 * q1 + w0 * q3);
    f[10] = 0.5 * (w2 * q0 + w1 * q1 - w0 * q2);
    const double h0 = __ldg(&P.jB[0]) * w0 + __ldg(&P.jB[3]) * w1 + __ldg(&P.jB[6]) * w2;
    const double h1 = __ldg(&P.jB[1]) * w0 + __ldg(&P.jB[4]) * w1 + __ldg(&P.jB[7]) * w2;
    const double h2 = __ldg(&P.jB[2]) * w0 + __ldg(&P.jB[5]) * w1 + __ldg(&P.jB[8]) * w2;
    const double m0 = (__ldg(&P.rTB[1]) * u[2] - __ldg(&P.rTB[2]) * u[1]) - (w1 * h2 - w2 * h1);
    const double m1 = (__ldg(&P.rTB[2]) * u[0] - __ldg(&P.rTB[0]) * u[2]) - (w2 * h0 - w0 * h2);
    const double m2 = (__ldg(&P.rTB[0]) * u[1] - __ldg(&P.rTB[1]) * u[0]) - (w0 * h1 - w1 * h0);
    f[11] = __ldg(&P.jBi[0]) * m0 + __ldg(&P.jBi[3]) * m1 + __ldg(&P.jBi[6]) * m2;
    f[12] = __ldg(&P.jBi[1]) * m0 + __ldg(&P.jBi[4]) * m1 + __ldg(&P.jBi[7]) * m2;
    f[13] = __ldg(&P.jBi[2]) * m0 + __ldg(&P.jBi[5]) * m1 + __ldg(&P.jBi[8]) * m2;
}

#ifndef SCVX_A_MINBLOCKS
#define SCVX_A_MINBLOCKS 2
#endif
__global__ void __launch_bounds__(128, SCVX_A_MINBLOCKS) stage_value_kernel(StagedArgs a) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.n_groups * GROUP) return;
    const ScvxBatch& bt = a.bt;
    const int ni = bt.n_nodes - 1;
    const bool live = t < a.count;
    const int w = a.first + (live ? t : a.count - 1);          // padded lanes recompute the last interval
    const int b = (int)(w / ni), i = (int)(w % ni);
    const scvx_probinfo& P = bt.P[bt.n_params == 1 ? 0 : b];
    const double* xin = bt.X + ((size_t)b * bt.n_nodes + i) * 14;
    const double* uin = bt.U + ((size_t)b * bt.n_nodes + i) * 3;
    const double sigma = bt.sigma[b];
    double x[14], um[3], up[3];
#pragma unroll
    for (int r = 0; r < 14; ++r) x[r] = xin[r];
#pragma unroll
    for (int c = 0; c < 3; ++c) { um[c] = uin[c]; up[c] = uin[3 + c]; }

    const int nst = 4 * bt.npts;
    double* rec = a.rec + ((size_t)(t >> 5) * nst) * ((size_t)a.rec_n * GROUP) + (t & 31);
    const double h = bt.dt / (double)bt.npts;
    const double pcs = 1.0 / (double)bt.npts;
    const double s = (bt.mode == SCVX_MODE_LITERAL) ? 1.0 : h;
    double pca = 0.0;
    for (int it = 0; it < bt.npts; ++it) {
        double acc[14], y[14];
#pragma unroll
        for (int r = 0; r < 14; ++r) { y[r] = x[r]; acc[r] = 0.0; }
#pragma unroll 1
        for (int st = 0; st < 4; ++st) {
            const double pc = (st == 0) ? pca : (st == 3 ? pca + pcs : pca + 0.5 * pcs);
            double uc[3], f[14], Fv[3][3], Fb[3][3];
#pragma unroll
            for (int c = 0; c < 3; ++c) uc[c] = (1.0 - pc) * um[c] + pc * up[c];
            rhs_value(P, a.tb, y, uc, f, Fv, Fb);
            // record: m, v, q, w, u, f_m, f_v, f_q, f_w [, dF/dv, dF/db]
            double* rp = rec + (size_t)(it * 4 + st) * ((size_t)a.rec_n * GROUP);
            if (a.rec_n == REC_AERO) {
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        rp[(25 + 3 * r + c) * GROUP] = Fv[r][c];
                        rp[(34 + 3 * r + c) * GROUP] = Fb[r][c];
                    }
            }
            rp[0 * GROUP] = y[0];
#pragma unroll
            for (int r = 0; r < 10; ++r) rp[(1 + r) * GROUP] = y[4 + r];
#pragma unroll
            for (int c = 0; c < 3; ++c) rp[(11 + c) * GROUP] = uc[c];
            rp[14 * GROUP] = f[0];
#pragma unroll
            for (int r = 0; r < 10; ++r) rp[(15 + r) * GROUP] = f[4 + r];
            const double wgt = (st == 0 || st == 3) ? 1.0 : 2.0;
            const double cy = (st == 2) ? s : 0.5 * s;
#pragma unroll
            for (int r = 0; r < 14; ++r) {
                const double k = f[r] * sigma;
                acc[r] = fma(wgt, k, acc[r]);
                y[r] = fma(cy, k, x[r]);
            }
        }
        pca += pcs;
#pragma unroll
        for (int r = 0; r < 14; ++r) x[r] = fma(h * (1.0 / 6.0), acc[r], x[r]);
    }
    if (!live) return;
    double* blk = bt.out_blocks + (size_t)w * SCVX_BLOCK_DOUBLES;
#pragma unroll
    for (int r = 0; r < 14; r += 2) *reinterpret_cast<double2*>(blk + r) = make_double2(x[r], x[r + 1]);
    if (bt.out_lin_err) {
        double* e = bt.out_lin_err + (size_t)w * 14;
#pragma unroll
        for (int r = 0; r < 14; r += 2) *reinterpret_cast<double2*>(e + r) = make_double2(x[r] - xin[14 + r], x[r + 1] - xin[15 + r]);
    }
    if (bt.out_tlb) {
        const int last = (i == ni - 1) ? 2 : 1;
        for (int k = 0; k < last; ++k) {
            const double* u = uin + 3 * k;
            const double nu = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
            double* o = bt.out_tlb + ((size_t)b * bt.n_nodes + i + k) * 4;
            *reinterpret_cast<double2*>(o) = make_double2(-(u[0] / nu), -(u[1] / nu));
            *reinterpret_cast<double2*>(o + 2) = make_double2(-(u[2] / nu), __ldg(&P.Tmin) - nu);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Kernel A2: the four "light" tangent columns d/d(m, v0, v1, v2) — one THREAD per interval.
// Only the v and r rows of these columns are non-trivial: K_v = Jvv Y_v (+ Jvm for the mass
// column), r rows are a quadrature of the v rows.  Reads m, f_v and dF/dv from the stage records, writes the
// seven columns d/d(m, r, v) of the block and the partial z = endpoint - D[:, m r v] * inp[m r v].
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 3) light_columns_kernel(StagedArgs a) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.count) return;
    const ScvxBatch& bt = a.bt;
    const int ni = bt.n_nodes - 1;
    const int w = a.first + t;
    const int b = w / ni, i = w % ni;
    const scvx_probinfo& P = bt.P[bt.n_params == 1 ? 0 : b];
    const double sigma = __ldg(bt.sigma + b), g0 = __ldg(&P.g0);
    const bool aero = (a.rec_n == REC_AERO);
    const int nst = 4 * bt.npts;
    const double* rec = a.rec + ((size_t)(t >> 5) * nst) * ((size_t)a.rec_n * GROUP) + (t & 31);
    const double h = bt.dt / (double)bt.npts;
    const double sstep = (bt.mode == SCVX_MODE_LITERAL) ? 1.0 : h;
    const double h6 = h * (1.0 / 6.0);
    double S[4][3], A[4][3], Y[4][3], Sr[4][3];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r) { S[c][r] = (c == r + 1) ? 1.0 : 0.0; Y[c][r] = S[c][r]; A[c][r] = 0.0; Sr[c][r] = 0.0; }
    // software-pipelined record reads: the 13 values of stage s+1 are in flight while stage s is computed
    const size_t rstride = (size_t)a.rec_n * GROUP;
    double nx[13];
    auto fetch = [&](int s) {
        const double* rp = rec + (size_t)s * rstride;
        nx[0] = __ldg(rp);
#pragma unroll
        for (int r = 0; r < 3; ++r) nx[1 + r] = __ldg(rp + (15 + r) * GROUP);
#pragma unroll
        for (int k = 0; k < 9; ++k) nx[4 + k] = aero ? __ldg(rp + (25 + k) * GROUP) : 0.0;
    };
    fetch(0);
#pragma unroll 1
    for (int s = 0; s < nst; ++s) {
        const int st = s & 3;
        double cu[13];
#pragma unroll
        for (int k = 0; k < 13; ++k) cu[k] = nx[k];
        if (s + 1 < nst) fetch(s + 1);
        const double sm = sigma / cu[0];
        double Jvv[3][3], Jvm[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            Jvm[r] = -sm * (cu[1 + r] + (r == 0 ? g0 : 0.0));
#pragma unroll
            for (int c = 0; c < 3; ++c) Jvv[r][c] = sm * cu[4 + 3 * r + c];
        }
        const double wgt = (st == 0 || st == 3) ? 1.0 : 2.0;
        const double cy = (st == 2) ? sstep : 0.5 * sstep;
        const double csg = h6 * wgt * sigma;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            double K[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                Sr[c][r] = fma(csg, Y[c][r], Sr[c][r]);
                K[r] = fma(Jvv[r][0], Y[c][0], fma(Jvv[r][1], Y[c][1], fma(Jvv[r][2], Y[c][2], c == 0 ? Jvm[r] : 0.0)));
            }
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                if (st != 3) { A[c][r] = fma(wgt, K[r], A[c][r]); Y[c][r] = fma(cy, K[r], S[c][r]); }
                else { S[c][r] = fma(h6, A[c][r] + K[r], S[c][r]); Y[c][r] = S[c][r]; A[c][r] = 0.0; }
            }
        }
    }
    double* blk = bt.out_blocks + (size_t)w * SCVX_BLOCK_DOUBLES;
    const double* xin = bt.X + ((size_t)b * bt.n_nodes + i) * 14;
    // columns: inp 0 (m), 1..3 (r), 4..6 (v)
#pragma unroll
    for (int c = 0; c < 7; ++c) {
        double col[14];
#pragma unroll
        for (int r = 0; r < 14; ++r) col[r] = 0.0;
        if (c == 0) { col[0] = 1.0; for (int r = 0; r < 3; ++r) { col[1 + r] = Sr[0][r]; col[4 + r] = S[0][r]; } }
        else if (c < 4) col[c] = 1.0;
        else { for (int r = 0; r < 3; ++r) { col[1 + r] = Sr[c - 3][r]; col[4 + r] = S[c - 3][r]; } }
        double* o = blk + 14 * (1 + c);
#pragma unroll
        for (int r = 0; r < 14; r += 2) *reinterpret_cast<double2*>(o + r) = make_double2(col[r], col[r + 1]);
    }
    // partial z (kernel B subtracts the remaining 14 columns)
    double z[14];
#pragma unroll
    for (int r = 0; r < 14; ++r) z[r] = blk[r];
    z[0] -= xin[0];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        z[1 + r] -= Sr[0][r] * xin[0] + xin[1 + r] + Sr[1][r] * xin[4] + Sr[2][r] * xin[5] + Sr[3][r] * xin[6];
        z[4 + r] -= S[0][r] * xin[0] + S[1][r] * xin[4] + S[2][r] * xin[5] + S[3][r] * xin[6];
    }
    double* o = blk + 14 * 22;
#pragma unroll
    for (int r = 0; r < 14; r += 2) *reinterpret_cast<double2*>(o + r) = make_double2(z[r], z[r + 1]);
}

// ------------------------------------------------------------------------------------------------
// mbarrier / TMA bulk-copy helpers (PTX)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n }" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n }" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        " .reg .pred p;\n"
        "WAIT_LOOP:\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        " @p bra WAIT_DONE;\n"
        " bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// Producer role: Jacobian blocks of one stage for 32 intervals (lane = interval).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void st2(double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }

// rec: this lane's stage record (stride GROUP doubles between entries, in shared memory);
// out: this lane's NJ-double Jacobian record in the ring.
__device__ __noinline__ void produce_stage(const scvx_probinfo& P, bool aero_rec, double sigma,
                                              const double* __restrict__ rec, double* __restrict__ out) {
    // parameters first, as one batch of independent read-only loads (one latency, not one per use)
    const double Pa = __ldg(&P.a), Pg0 = __ldg(&P.g0);
    double jB[9], jBi[9], rT[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) { jB[k] = __ldg(&P.jB[k]); jBi[k] = __ldg(&P.jBi[k]); }
#pragma unroll
    for (int k = 0; k < 3; ++k) rT[k] = __ldg(&P.rTB[k]);
    const double m = rec[0];
    const double v[3] = { rec[1 * GROUP], rec[2 * GROUP], rec[3 * GROUP] };
    const double q0 = rec[4 * GROUP], q1 = rec[5 * GROUP], q2 = rec[6 * GROUP], q3 = rec[7 * GROUP];
    const double w0 = rec[8 * GROUP], w1 = rec[9 * GROUP], w2 = rec[10 * GROUP];
    const double u0 = rec[11 * GROUP], u1 = rec[12 * GROUP], u2 = rec[13 * GROUP];
    const double fm = rec[14 * GROUP];
    const double fv[3] = { rec[15 * GROUP], rec[16 * GROUP], rec[17 * GROUP] };
    const double fq[4] = { rec[18 * GROUP], rec[19 * GROUP], rec[20 * GROUP], rec[21 * GROUP] };
    const double fw[3] = { rec[22 * GROUP], rec[23 * GROUP], rec[24 * GROUP] };
    const double sm = sigma / m;

    // ---- rotational block: Jww = -sigma * jBi * ([w]x jB - [jB w]x)
    {
        const double L0 = jB[0] * w0 + jB[3] * w1 + jB[6] * w2;
        const double L1 = jB[1] * w0 + jB[4] * w1 + jB[7] * w2;
        const double L2 = jB[2] * w0 + jB[5] * w1 + jB[8] * w2;
        double M[3][3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {          // column c of [w]x jB = w x jB[:,c]
            const double a0 = jB[3 * c], a1 = jB[3 * c + 1], a2 = jB[3 * c + 2];
            M[0][c] = w1 * a2 - w2 * a1; M[1][c] = w2 * a0 - w0 * a2; M[2][c] = w0 * a1 - w1 * a0;
        }
        // minus [L]x = [[0,-L2,L1],[L2,0,-L0],[-L1,L0,0]]
        M[0][1] += L2; M[0][2] -= L1; M[1][0] -= L2; M[1][2] += L0; M[2][0] += L1; M[2][1] -= L0;
        double Jw[9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                Jw[3 * r + c] = -sigma * (jBi[r] * M[0][c] + jBi[r + 3] * M[1][c] + jBi[r + 6] * M[2][c]);
        const double hs = 0.5 * sigma;
        st2(out + J_WW + 0, Jw[0], Jw[1]); st2(out + J_WW + 2, Jw[2], Jw[3]); st2(out + J_WW + 4, Jw[4], Jw[5]);
        st2(out + J_WW + 6, Jw[6], Jw[7]); st2(out + J_WW + 8, Jw[8], hs * w0);
        st2(out + J_HW + 1, hs * w1, hs * w2);
        st2(out + J_HQ + 0, hs * q0, hs * q1); st2(out + J_HQ + 2, hs * q2, hs * q3);
    }
    // ---- translational block
    const double c00 = 1.0 - 2.0 * (q2 * q2 + q3 * q3), c01 = 2.0 * (q1 * q2 - q0 * q3), c02 = 2.0 * (q1 * q3 + q0 * q2);
    const double c10 = 2.0 * (q1 * q2 + q0 * q3), c11 = 1.0 - 2.0 * (q1 * q1 + q3 * q3), c12 = 2.0 * (q2 * q3 - q0 * q1);
    const double c20 = 2.0 * (q1 * q3 - q0 * q2), c21 = 2.0 * (q2 * q3 + q0 * q1), c22 = 1.0 - 2.0 * (q1 * q1 + q2 * q2);
    // d(C u)/dq, row-major 3x4
    double Jq[12];
    Jq[0] = 2.0 * (q2 * u2 - q3 * u1);             Jq[1] = 2.0 * (q2 * u1 + q3 * u2);
    Jq[2] = 2.0 * (q1 * u1 + q0 * u2) - 4.0 * q2 * u0; Jq[3] = 2.0 * (q1 * u2 - q0 * u1) - 4.0 * q3 * u0;
    Jq[4] = 2.0 * (q3 * u0 - q1 * u2);             Jq[5] = 2.0 * (q2 * u0 - q0 * u2) - 4.0 * q1 * u1;
    Jq[6] = 2.0 * (q1 * u0 + q3 * u2);             Jq[7] = 2.0 * (q0 * u0 + q2 * u2) - 4.0 * q3 * u1;
    Jq[8] = 2.0 * (q1 * u1 - q2 * u0);             Jq[9] = 2.0 * (q3 * u0 + q0 * u1) - 4.0 * q1 * u2;
    Jq[10] = 2.0 * (q3 * u1 - q0 * u0) - 4.0 * q2 * u2; Jq[11] = 2.0 * (q1 * u0 + q2 * u1);
    double Jvv[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
    if (aero_rec) {
        double Fv[3][3], Fb[3][3];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) { Fv[r][c] = rec[(25 + 3 * r + c) * GROUP]; Fb[r][c] = rec[(34 + 3 * r + c) * GROUP]; }
        // db/dq
        const double B[3][4] = { { 0.0, 0.0, -4.0 * q2, -4.0 * q3 },
                                 { 2.0 * q3, 2.0 * q2, 2.0 * q1, 2.0 * q0 },
                                 { -2.0 * q2, 2.0 * q3, -2.0 * q0, 2.0 * q1 } };
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int c = 0; c < 4; ++c) Jq[4 * r + c] += Fb[r][0] * B[0][c] + Fb[r][1] * B[1][c] + Fb[r][2] * B[2][c];
#pragma unroll
            for (int c = 0; c < 3; ++c) Jvv[3 * r + c] = sm * Fv[r][c];
        }
    }
    // Jvm = -sigma * ((C u + F)/m) / m = -(sigma/m) * (f_v + g0 e1); one 8-double row per v component
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        double* o = out + J_V + 8 * r;
        st2(o + 0, -sm * (fv[r] + (r == 0 ? Pg0 : 0.0)), Jvv[3 * r]);
        st2(o + 2, Jvv[3 * r + 1], Jvv[3 * r + 2]);
        st2(o + 4, sm * Jq[4 * r], sm * Jq[4 * r + 1]);
        st2(o + 6, sm * Jq[4 * r + 2], sm * Jq[4 * r + 3]);
    }
    // ---- direct (control / sigma) columns: G[col][row], rows m, v0..2, w0..2
    const double nu = sqrt(u0 * u0 + u1 * u1 + u2 * u2);
    const double gm = -sigma * Pa / nu;
    // jBi * (rTB x e_j): rTB x e0 = (0, r2, -r1); x e1 = (-r2, 0, r0); x e2 = (r1, -r0, 0)
    double G[28];
    G[0] = gm * u0; G[1] = sm * c00; G[2] = sm * c10; G[3] = sm * c20;
    G[4] = sigma * (jBi[3] * rT[2] - jBi[6] * rT[1]); G[5] = sigma * (jBi[4] * rT[2] - jBi[7] * rT[1]); G[6] = sigma * (jBi[5] * rT[2] - jBi[8] * rT[1]);
    G[7] = gm * u1; G[8] = sm * c01; G[9] = sm * c11; G[10] = sm * c21;
    G[11] = sigma * (jBi[6] * rT[0] - jBi[0] * rT[2]); G[12] = sigma * (jBi[7] * rT[0] - jBi[1] * rT[2]); G[13] = sigma * (jBi[8] * rT[0] - jBi[2] * rT[2]);
    G[14] = gm * u2; G[15] = sm * c02; G[16] = sm * c12; G[17] = sm * c22;
    G[18] = sigma * (jBi[0] * rT[1] - jBi[3] * rT[0]); G[19] = sigma * (jBi[1] * rT[1] - jBi[4] * rT[0]); G[20] = sigma * (jBi[2] * rT[1] - jBi[5] * rT[0]);
    G[21] = fm; G[22] = fv[0]; G[23] = fv[1]; G[24] = fv[2]; G[25] = fw[0]; G[26] = fw[1]; G[27] = fw[2];
#pragma unroll
    for (int k = 0; k < 28; k += 2) st2(out + J_G + k, G[k], G[k + 1]);
    st2(out + J_FRQ + 0, v[0], v[1]); st2(out + J_FRQ + 2, v[2], fq[0]);
    st2(out + J_FRQ + 4, fq[1], fq[2]); st2(out + J_FRQ + 6, fq[3], sigma);
    st2(out + J_FRQ + 8, 0.0, 0.0);
}

// ------------------------------------------------------------------------------------------------
// Kernel B
// ------------------------------------------------------------------------------------------------
struct __align__(16) TangentSmem {
    double ring[RING][GROUP][NJ];            // Jacobian records
    double recbuf[NWARP][REC_MAX * GROUP];   // per-warp stage-record staging (TMA destination)
    uint64_t full[RING];
    uint64_t empty[RING];
    uint64_t recfull[NWARP];
};

__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }

// one full tangent column: rows m, v(3), q(4), w(3) carried as (S, acc, Y); r rows as a pure quadrature
struct FullCol {
    double S[11], A[11], Y[11], Sr[3];
};


// One rk4 stage of TWO full tangent columns (8-lane variant).  LAST = final stage of the step (compile time, so
// there is no control flow inside the stage; two instantiations keep the hot loop inside the instruction cache).
// Rows are processed in cascade order (r, v, m, q, w): a row block is overwritten only after every block that
// reads its old stage value has been formed.
template <bool LAST>
__device__ __forceinline__ void consume_stage8(FullCol& FA, FullCol& FB, const double* __restrict__ J, const int gcol,
                                               const int l8, const double pc, const double wgt, const double cy,
                                               const double h6, uint64_t* empty_bar, const int lane) {
    constexpr bool last = LAST;
    const double alA = (l8 < 3) ? 1.0 - pc : (l8 == 3 ? 1.0 : 0.0), alB = (l8 < 3) ? pc : 0.0;
    const double dsA = (l8 == 3) ? 1.0 : 0.0;
    const double* Gc = J + J_G + 7 * gcol;
    auto upd = [&](FullCol& F, const int idx, const double K) {
        if constexpr (!last) { F.A[idx] = fma(wgt, K, F.A[idx]); F.Y[idx] = fma(cy, K, F.S[idx]); }
        else { F.S[idx] = fma(h6, F.A[idx] + K, F.S[idx]); F.Y[idx] = F.S[idx]; F.A[idx] = 0.0; }
    };
    // ---- r rows (pure quadrature): S_r += cr * (sigma * Y_v + dsigma * f_r)
    {
        const double2 fr01 = ld2(J + J_FRQ);
        const double fr2 = J[J_FRQ + 2];
        const double sg = J[J_FRQ + 7];
        const double cr = h6 * wgt;
        const double csg = cr * sg, cds = cr * dsA;
        FA.Sr[0] = fma(csg, FA.Y[1], fma(cds, fr01.x, FA.Sr[0]));
        FA.Sr[1] = fma(csg, FA.Y[2], fma(cds, fr01.y, FA.Sr[1]));
        FA.Sr[2] = fma(csg, FA.Y[3], fma(cds, fr2, FA.Sr[2]));
#pragma unroll
        for (int r = 0; r < 3; ++r) FB.Sr[r] = fma(csg, FB.Y[1 + r], FB.Sr[r]);
    }
    // ---- v rows: K_v = Jvm Y_m + Jvv Y_v + Jvq Y_q + alpha * G_v
    {
        double kA[3], kB[3];
#pragma unroll
        for (int row = 0; row < 3; ++row) {
            const double2 c01 = ld2(J + J_V + 8 * row), c23 = ld2(J + J_V + 8 * row + 2);
            const double2 qa = ld2(J + J_V + 8 * row + 4), qb = ld2(J + J_V + 8 * row + 6);
            const double gg = Gc[1 + row];
            kA[row] = fma(c01.x, FA.Y[0], fma(c01.y, FA.Y[1], fma(c23.x, FA.Y[2], fma(c23.y, FA.Y[3], alA * gg)))) +
                      fma(qa.x, FA.Y[4], fma(qa.y, FA.Y[5], fma(qb.x, FA.Y[6], qb.y * FA.Y[7])));
            kB[row] = fma(c01.x, FB.Y[0], fma(c01.y, FB.Y[1], fma(c23.x, FB.Y[2], fma(c23.y, FB.Y[3], alB * gg)))) +
                      fma(qa.x, FB.Y[4], fma(qa.y, FB.Y[5], fma(qb.x, FB.Y[6], qb.y * FB.Y[7])));
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) { upd(FA, 1 + r, kA[r]); upd(FB, 1 + r, kB[r]); }
    }
    // ---- m row: K_m = alpha * G_m
    {
        const double gmv = Gc[0];
        upd(FA, 0, alA * gmv); upd(FB, 0, alB * gmv);
    }
    // ---- q rows: K_q = Omega(hw) Y_q + Omega(Y_w) hq + dsigma * f_q
    {
        const double hw0 = J[J_HW], hw1 = J[J_HW + 1], hw2 = J[J_HW + 2];
        const double2 hq01 = ld2(J + J_HQ), hq23 = ld2(J + J_HQ + 2);
        const double hq0 = hq01.x, hq1 = hq01.y, hq2 = hq23.x, hq3 = hq23.y;
        const double fq0 = J[J_FRQ + 3];
        const double2 fq12 = ld2(J + J_FRQ + 4);
        const double fq3 = J[J_FRQ + 6];
        double kA[4], kB[4];
#define QROWS(F, K, ds)                                                                                                                     \
        K[0] = fma(-hw0, F.Y[5], fma(-hw1, F.Y[6], fma(-hw2, F.Y[7], ds * fq0))) + fma(-hq1, F.Y[8], fma(-hq2, F.Y[9], -hq3 * F.Y[10]));     \
        K[1] = fma(hw0, F.Y[4], fma(hw2, F.Y[6], fma(-hw1, F.Y[7], ds * fq12.x))) + fma(hq0, F.Y[8], fma(hq2, F.Y[10], -hq3 * F.Y[9]));      \
        K[2] = fma(hw1, F.Y[4], fma(-hw2, F.Y[5], fma(hw0, F.Y[7], ds * fq12.y))) + fma(hq0, F.Y[9], fma(-hq1, F.Y[10], hq3 * F.Y[8]));      \
        K[3] = fma(hw2, F.Y[4], fma(hw1, F.Y[5], fma(-hw0, F.Y[6], ds * fq3))) + fma(hq0, F.Y[10], fma(hq1, F.Y[9], -hq2 * F.Y[8]));
        QROWS(FA, kA, dsA)
        QROWS(FB, kB, 0.0)
#undef QROWS
#pragma unroll
        for (int r = 0; r < 4; ++r) { upd(FA, 4 + r, kA[r]); upd(FB, 4 + r, kB[r]); }
    }
    // ---- w rows: K_w = Jww * Y_w + alpha * G_w
    {
        const double2 j01 = ld2(J + J_WW), j23 = ld2(J + J_WW + 2), j45 = ld2(J + J_WW + 4), j67 = ld2(J + J_WW + 6);
        const double j8 = J[J_WW + 8];
        const double g0 = Gc[4], g1 = Gc[5], g2 = Gc[6];
        double kA[3], kB[3];
        kA[0] = fma(j01.x, FA.Y[8], fma(j01.y, FA.Y[9], fma(j23.x, FA.Y[10], alA * g0)));
        kA[1] = fma(j23.y, FA.Y[8], fma(j45.x, FA.Y[9], fma(j45.y, FA.Y[10], alA * g1)));
        kA[2] = fma(j67.x, FA.Y[8], fma(j67.y, FA.Y[9], fma(j8, FA.Y[10], alA * g2)));
        kB[0] = fma(j01.x, FB.Y[8], fma(j01.y, FB.Y[9], fma(j23.x, FB.Y[10], alB * g0)));
        kB[1] = fma(j23.y, FB.Y[8], fma(j45.x, FB.Y[9], fma(j45.y, FB.Y[10], alB * g1)));
        kB[2] = fma(j67.x, FB.Y[8], fma(j67.y, FB.Y[9], fma(j8, FB.Y[10], alB * g2)));
        // all reads of the ring slot are done: hand it back
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar);
#pragma unroll
        for (int r = 0; r < 3; ++r) { upd(FA, 8 + r, kA[r]); upd(FB, 8 + r, kB[r]); }
    }
}

__global__ void __launch_bounds__(256, 1) tangent_kernel(StagedArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TangentSmem& sm = *reinterpret_cast<TangentSmem*>(smem_raw);
    const ScvxBatch& bt = a.bt;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int l8 = lane & 7;                   // column-group lane within the interval
    const int sub = lane >> 3;                 // interval within the warp (0..3)
    const int ni = bt.n_nodes - 1;
    const int nst = 4 * bt.npts;
    const double h = bt.dt / (double)bt.npts;
    const double pcs = 1.0 / (double)bt.npts;
    const double sstep = (bt.mode == SCVX_MODE_LITERAL) ? 1.0 : h;
    const double h6 = h * (1.0 / 6.0);

    if (tid == 0) {
        for (int r = 0; r < RING; ++r) { mbar_init(&sm.full[r], 32); mbar_init(&sm.empty[r], NWARP); }
        for (int wq = 0; wq < NWARP; ++wq) mbar_init(&sm.recfull[wq], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // groups handled by this CTA: g = blockIdx.x, blockIdx.x + gridDim.x, ...
    const int my_groups = (a.n_groups > (int)blockIdx.x) ? (a.n_groups - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int total_stages = my_groups * nst;               // global stage counter T = it * nst + s

    // ---- static per-lane column configuration
    // full slots: lanes 0..2: (u-_j, u+_j); lane 3: (sigma, -); lane 4: (w0,w1); 5: (w2,q0); 6: (q1,q2); 7: (q3,-)
    int colA = -1, colB = -1, gcol = 3;
    if (l8 < 3) { colA = 14 + l8; colB = 17 + l8; gcol = l8; }
    else if (l8 == 3) colA = 20;
    else if (l8 == 4) { colA = 11; colB = 12; }
    else if (l8 == 5) { colA = 13; colB = 7; }
    else if (l8 == 6) { colA = 8; colB = 9; }
    else colA = 10;
    // direct-term coefficients, re-derived from l8 where used (keeps them out of long-lived registers):
    //   alpha_A = 1-pc (u- columns) | 1 (sigma column) | 0 ;  alpha_B = pc (u+ columns) | 0 ;
    //   dsigma_A = 1 for the sigma column

    // producer bookkeeping: this warp produces global stages T with T % NWARP == warp
    int nextP = warp;                                // next global stage this warp has to produce
    int p_it = 0, p_s = warp;                        // ... as (group pass, stage) ; nst >= 4
    while (p_s >= nst) { p_s -= nst; ++p_it; }
    uint32_t rec_phase = 0;
    auto issue_record = [&](int it, int s) {         // lane 0 only: TMA the stage record of (pass it, stage s)
        const int g = blockIdx.x + it * gridDim.x;
        const uint32_t bytes = (uint32_t)a.rec_n * GROUP * 8;
        const double* src = a.rec + ((size_t)g * nst + s) * ((size_t)a.rec_n * GROUP);
        mbar_expect_tx(&sm.recfull[warp], bytes);
        bulk_g2s(sm.recbuf[warp], src, bytes, &sm.recfull[warp]);
    };
    if (lane == 0 && nextP < total_stages) issue_record(p_it, p_s);

    auto produce = [&]() {                           // produce global stage nextP = (p_it, p_s), then advance
        const int g = blockIdx.x + p_it * gridDim.x;
        int t = g * GROUP + lane; if (t >= a.count) t = a.count - 1;
        const int b = (a.first + t) / ni;
        const scvx_probinfo& P = bt.P[bt.n_params == 1 ? 0 : b];
        const double sigma = bt.sigma[b];
        const int slot = nextP % RING;
        const int use = nextP / RING;
        mbar_wait(&sm.recfull[warp], rec_phase); rec_phase ^= 1;
        if (use > 0) mbar_wait(&sm.empty[slot], (uint32_t)((use - 1) & 1));
        produce_stage(P, a.rec_n == REC_AERO, sigma, sm.recbuf[warp] + lane, &sm.ring[slot][lane][0]);
        mbar_arrive(&sm.full[slot]);
        nextP += NWARP; p_s += NWARP;
        while (p_s >= nst) { p_s -= nst; ++p_it; }
        __syncwarp();
        if (lane == 0 && nextP < total_stages) { fence_proxy_async(); issue_record(p_it, p_s); }
    };

    // prologue: stages 0..LOOKAHEAD-1
    for (int k = 0; k < LOOKAHEAD; ++k)
        if (nextP == k && nextP < total_stages) produce();

    FullCol FA, FB;
    double park[72];
    volatile double* vp = park;
    int T = 0;
    int c_slot = 0;
    uint32_t c_phase = 0;
    for (int it = 0; it < my_groups; ++it) {
        const int g = blockIdx.x + it * gridDim.x;
        // ---- initial tangent: S = [I | 0]
#pragma unroll
        for (int r = 0; r < 11; ++r) {
            // local row order of a full column: 0 m, 1..3 v, 4..7 q, 8..10 w  (inp column of local row r >= 4 is r + 3)
            FA.S[r] = (r >= 4 && colA == r + 3) ? 1.0 : 0.0;
            FB.S[r] = (r >= 4 && colB == r + 3) ? 1.0 : 0.0;
            FA.A[r] = 0.0; FB.A[r] = 0.0;
            FA.Y[r] = FA.S[r]; FB.Y[r] = FB.S[r];
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) FA.Sr[r] = FB.Sr[r] = 0.0;

        double pca = 0.0;
#pragma unroll 1
        for (int s = 0; s < nst; ++s, ++T) {
            if (nextP == T + LOOKAHEAD && nextP < total_stages) {
                // manual live-range split: park the tangent state in local memory across the producer call so
                // that it never competes with the producer for registers inside the hot loop
#pragma unroll
                for (int r = 0; r < 11; ++r) {
                    vp[r] = FA.S[r]; vp[11 + r] = FA.A[r]; vp[22 + r] = FA.Y[r];
                    vp[36 + r] = FB.S[r]; vp[47 + r] = FB.A[r]; vp[58 + r] = FB.Y[r];
                }
#pragma unroll
                for (int r = 0; r < 3; ++r) { vp[33 + r] = FA.Sr[r]; vp[69 + r] = FB.Sr[r]; }
                produce();
#pragma unroll
                for (int r = 0; r < 11; ++r) {
                    FA.S[r] = vp[r]; FA.A[r] = vp[11 + r]; FA.Y[r] = vp[22 + r];
                    FB.S[r] = vp[36 + r]; FB.A[r] = vp[47 + r]; FB.Y[r] = vp[58 + r];
                }
#pragma unroll
                for (int r = 0; r < 3; ++r) { FA.Sr[r] = vp[33 + r]; FB.Sr[r] = vp[69 + r]; }
            }
            const int st = s & 3;
            const int slot = c_slot;
            mbar_wait(&sm.full[slot], c_phase);
            if (++c_slot == RING) { c_slot = 0; c_phase ^= 1; }
            const double* J = &sm.ring[slot][warp * 4 + sub][0];
            if (st == 3) {
                consume_stage8<true>(FA, FB, J, gcol, l8, pca + pcs, 1.0, 0.0, h6, &sm.empty[slot], lane);
                pca += pcs;
            } else {
                const double pc = (st == 0) ? pca : pca + 0.5 * pcs;
                consume_stage8<false>(FA, FB, J, gcol, l8, pc, st == 0 ? 1.0 : 2.0, st == 2 ? sstep : 0.5 * sstep, h6,
                                      &sm.empty[slot], lane);
            }
        }

        // ---- epilogue: write D columns and z for interval (g*32 + warp*4 + sub)
        const int t = g * GROUP + warp * 4 + sub;
        const bool live = t < a.count;
        const int wi = a.first + (live ? t : a.count - 1);
        const int b = (int)(wi / ni), i = (int)(wi % ni);
        double* blk = bt.out_blocks + (size_t)wi * SCVX_BLOCK_DOUBLES;
        const double* xin = bt.X + ((size_t)b * bt.n_nodes + i) * 14;
        const double* uin = bt.U + ((size_t)b * bt.n_nodes + i) * 3;
        auto inp_of = [&](int c) -> double {
            if (c < 0) return 0.0;
            if (c < 14) return xin[c];
            if (c < 20) return uin[c - 14];
            return bt.sigma[b];
        };
        double zp[14];
#pragma unroll
        for (int r = 0; r < 14; ++r) zp[r] = 0.0;
        auto emit_full = [&](const FullCol& F, int c) {
            if (c < 0) return;
            // state row order: m, r(3), v(3), q(4), w(3)
            const double col[14] = { F.S[0], F.Sr[0], F.Sr[1], F.Sr[2], F.S[1], F.S[2], F.S[3], F.S[4], F.S[5], F.S[6], F.S[7],
                                     F.S[8], F.S[9], F.S[10] };
            const double xc = inp_of(c);
            double* o = blk + 14 * (1 + c);
#pragma unroll
            for (int r = 0; r < 14; r += 2) {
                if (live) *reinterpret_cast<double2*>(o + r) = make_double2(col[r], col[r + 1]);
                zp[r] = fma(col[r], xc, zp[r]); zp[r + 1] = fma(col[r + 1], xc, zp[r + 1]);
            }
        };
        emit_full(FA, colA);
        emit_full(FB, colB);
        // z = (partial z of kernel A2) - D[:, heavy columns] * inp  (sum over the 8 lanes of this interval)
#pragma unroll
        for (int r = 0; r < 14; ++r) {
            double v = zp[r];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            zp[r] = v;
        }
        if (live && l8 == 7) {
            double* o = blk + 14 * 22;
#pragma unroll
            for (int r = 0; r < 14; r += 2) {
                const double2 e = *reinterpret_cast<const double2*>(o + r);        // partial z written by kernel A2
                *reinterpret_cast<double2*>(o + r) = make_double2(e.x - zp[r], e.y - zp[r + 1]);
            }
        }
    }
}


// ------------------------------------------------------------------------------------------------
// Kernel B, 16-lane variant: ONE full tangent column per lane, 16 lanes per interval, 2 intervals per warp,
// 16 warps (512 threads, <= 128 registers) per CTA, still 32 intervals in flight per SM.  Twice the warps per
// scheduler of the 8-lane variant for the same FP64 work: the stage loop is latency bound, not issue bound.
// Stage records are staged in a pool of NB buffers indexed by the global stage number; the warp that will
// produce stage T issues its TMA bulk copy PREFETCH stages before it produces.
// ------------------------------------------------------------------------------------------------
constexpr int NWARP16 = 16;
constexpr int NB16 = 6;            // record buffers; must be >= LOOKAHEAD + PREFETCH + 1
constexpr int PREFETCH16 = 2;
static_assert(NB16 >= LOOKAHEAD + PREFETCH16 + 1, "record buffer pool too small");

struct __align__(16) Tangent16Smem {
    double ring[RING][GROUP][NJ];
    double recbuf[NB16][REC_MAX * GROUP];
    uint64_t full[RING];
    uint64_t empty[RING];
    uint64_t recfull[NB16];
};

__global__ void __launch_bounds__(512, 1) tangent16_kernel(StagedArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Tangent16Smem& sm = *reinterpret_cast<Tangent16Smem*>(smem_raw);
    const ScvxBatch& bt = a.bt;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int l16 = lane & 15;                 // column lane within the interval
    const int sub = lane >> 4;                 // interval within the warp (0..1)
    const int ni = bt.n_nodes - 1;
    const int nst = 4 * bt.npts;
    const double h = bt.dt / (double)bt.npts;
    const double pcs = 1.0 / (double)bt.npts;
    const double sstep = (bt.mode == SCVX_MODE_LITERAL) ? 1.0 : h;
    const double h6 = h * (1.0 / 6.0);

    if (tid == 0) {
        for (int r = 0; r < RING; ++r) { mbar_init(&sm.full[r], 32); mbar_init(&sm.empty[r], NWARP16); }
        for (int q = 0; q < NB16; ++q) mbar_init(&sm.recfull[q], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int my_groups = (a.n_groups > (int)blockIdx.x) ? (a.n_groups - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int total_stages = my_groups * nst;

    // column of this lane: 0..2 u-_j | 3..5 u+_j | 6 sigma | 7..9 w_j | 10..13 q_j | 14,15 idle
    int col = -1, gcol = 3;
    if (l16 < 3) { col = 14 + l16; gcol = l16; }
    else if (l16 < 6) { col = 14 + l16; gcol = l16 - 3; }
    else if (l16 == 6) col = 20;
    else if (l16 < 10) col = 4 + l16;            // w_j -> inp columns 11..13
    else if (l16 < 14) col = l16 - 3;            // q_j -> inp columns 7..10

    auto issue_record = [&](int Tp) {             // one lane: TMA the stage record of global stage Tp
        const int it = Tp / nst, s = Tp - it * nst;
        const int g = blockIdx.x + it * gridDim.x;
        const uint32_t bytes = (uint32_t)a.rec_n * GROUP * 8;
        const double* src = a.rec + ((size_t)g * nst + s) * ((size_t)a.rec_n * GROUP);
        const int q = Tp % NB16;
        fence_proxy_async();
        mbar_expect_tx(&sm.recfull[q], bytes);
        bulk_g2s(sm.recbuf[q], src, bytes, &sm.recfull[q]);
    };
    auto produce = [&](int Tp) {                  // whole warp, lane = interval
        const int it = Tp / nst;
        const int g = blockIdx.x + it * gridDim.x;
        int t = g * GROUP + lane; if (t >= a.count) t = a.count - 1;
        const int b = (a.first + t) / ni;
        const scvx_probinfo& P = bt.P[bt.n_params == 1 ? 0 : b];
        const double sigma = bt.sigma[b];
        const int slot = Tp % RING, use = Tp / RING;
        const int q = Tp % NB16;
        mbar_wait(&sm.recfull[q], (uint32_t)((Tp / NB16) & 1));
        if (use > 0) mbar_wait(&sm.empty[slot], (uint32_t)((use - 1) & 1));
        produce_stage(P, a.rec_n == REC_AERO, sigma, sm.recbuf[q] + lane, &sm.ring[slot][lane][0]);
        mbar_arrive(&sm.full[slot]);
    };

    // prologue: records of stages 0..LOOKAHEAD+PREFETCH-1 in flight, stages 0..LOOKAHEAD-1 produced
    for (int k = 0; k < LOOKAHEAD + PREFETCH16; ++k)
        if ((k % NWARP16) == warp && lane == 0 && k < total_stages) issue_record(k);
    for (int k = 0; k < LOOKAHEAD; ++k)
        if ((k % NWARP16) == warp && k < total_stages) produce(k);

    FullCol F;
    double park[36];
    volatile double* vp = park;
    int T = 0, c_slot = 0;
    uint32_t c_phase = 0;
    for (int it = 0; it < my_groups; ++it) {
        const int g = blockIdx.x + it * gridDim.x;
#pragma unroll
        for (int r = 0; r < 11; ++r) {
            F.S[r] = (r >= 4 && col == r + 3) ? 1.0 : 0.0;       // local rows: 0 m, 1..3 v, 4..7 q, 8..10 w
            F.A[r] = 0.0; F.Y[r] = F.S[r];
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) F.Sr[r] = 0.0;

        double pca = 0.0;
#pragma unroll 1
        for (int s = 0; s < nst; ++s, ++T) {
            {   // producer duties of this warp
                const int Ti = T + LOOKAHEAD + PREFETCH16;
                if ((Ti & (NWARP16 - 1)) == warp && lane == 0 && Ti < total_stages) issue_record(Ti);
                const int Tp = T + LOOKAHEAD;
                if ((Tp & (NWARP16 - 1)) == warp && Tp < total_stages) {
                    // manual live-range split: park the tangent state in local memory across the producer call
#pragma unroll
                    for (int r = 0; r < 11; ++r) { vp[r] = F.S[r]; vp[11 + r] = F.A[r]; vp[22 + r] = F.Y[r]; }
#pragma unroll
                    for (int r = 0; r < 3; ++r) vp[33 + r] = F.Sr[r];
                    produce(Tp);
#pragma unroll
                    for (int r = 0; r < 11; ++r) { F.S[r] = vp[r]; F.A[r] = vp[11 + r]; F.Y[r] = vp[22 + r]; }
#pragma unroll
                    for (int r = 0; r < 3; ++r) F.Sr[r] = vp[33 + r];
                }
            }
            const int st = s & 3;
            const bool last = (st == 3);
            const double pc = (st == 0) ? pca : (last ? pca + pcs : pca + 0.5 * pcs);
            const double wgt = (st == 0 || last) ? 1.0 : 2.0;
            const double cy = (st == 2) ? sstep : 0.5 * sstep;
            const double cr = h6 * wgt;
            const double al = (l16 < 3) ? 1.0 - pc : (l16 < 6 ? pc : (l16 == 6 ? 1.0 : 0.0));
            const double ds = (l16 == 6) ? 1.0 : 0.0;
            const int slot = c_slot;
            mbar_wait(&sm.full[slot], c_phase);
            if (++c_slot == RING) { c_slot = 0; c_phase ^= 1; }
            const double* J = &sm.ring[slot][warp * 2 + sub][0];
            const double* Gc = J + J_G + 7 * gcol;
#define UPD1(idx, Kv)                                                                               \
            if (!last) { F.A[idx] = fma(wgt, (Kv), F.A[idx]); F.Y[idx] = fma(cy, (Kv), F.S[idx]); }  \
            else { F.S[idx] = fma(h6, F.A[idx] + (Kv), F.S[idx]); F.Y[idx] = F.S[idx]; F.A[idx] = 0.0; }
            // ---- r rows (quadrature)
            {
                const double2 fr01 = ld2(J + J_FRQ);
                const double fr2 = J[J_FRQ + 2];
                const double sg = J[J_FRQ + 7];
                const double csg = cr * sg, cds = cr * ds;
                F.Sr[0] = fma(csg, F.Y[1], fma(cds, fr01.x, F.Sr[0]));
                F.Sr[1] = fma(csg, F.Y[2], fma(cds, fr01.y, F.Sr[1]));
                F.Sr[2] = fma(csg, F.Y[3], fma(cds, fr2, F.Sr[2]));
            }
            // ---- v rows
            {
                double k[3];
#define VROW1(row)                                                                                                      \
                {                                                                                                         \
                    const double2 c01 = ld2(J + J_V + 8 * row), c23 = ld2(J + J_V + 8 * row + 2);                         \
                    const double2 qa = ld2(J + J_V + 8 * row + 4), qb = ld2(J + J_V + 8 * row + 6);                       \
                    const double t0 = fma(c01.x, F.Y[0], fma(c01.y, F.Y[1], fma(c23.x, F.Y[2], fma(c23.y, F.Y[3], al * Gc[1 + row])))); \
                    const double t1 = fma(qa.x, F.Y[4], fma(qa.y, F.Y[5], fma(qb.x, F.Y[6], qb.y * F.Y[7])));             \
                    k[row] = t0 + t1;                                                                                     \
                }
                VROW1(0)
                VROW1(1)
                VROW1(2)
#undef VROW1
#pragma unroll
                for (int r = 0; r < 3; ++r) { UPD1(1 + r, k[r]) }
            }
            // ---- m row
            { const double gmv = Gc[0]; UPD1(0, al * gmv) }
            // ---- q rows
            {
                const double hw0 = J[J_HW], hw1 = J[J_HW + 1], hw2 = J[J_HW + 2];
                const double2 hq01 = ld2(J + J_HQ), hq23 = ld2(J + J_HQ + 2);
                const double hq0 = hq01.x, hq1 = hq01.y, hq2 = hq23.x, hq3 = hq23.y;
                const double fq0 = J[J_FRQ + 3];
                const double2 fq12 = ld2(J + J_FRQ + 4);
                const double fq3 = J[J_FRQ + 6];
                double k[4];
                k[0] = fma(-hw0, F.Y[5], fma(-hw1, F.Y[6], fma(-hw2, F.Y[7], ds * fq0))) + fma(-hq1, F.Y[8], fma(-hq2, F.Y[9], -hq3 * F.Y[10]));
                k[1] = fma(hw0, F.Y[4], fma(hw2, F.Y[6], fma(-hw1, F.Y[7], ds * fq12.x))) + fma(hq0, F.Y[8], fma(hq2, F.Y[10], -hq3 * F.Y[9]));
                k[2] = fma(hw1, F.Y[4], fma(-hw2, F.Y[5], fma(hw0, F.Y[7], ds * fq12.y))) + fma(hq0, F.Y[9], fma(-hq1, F.Y[10], hq3 * F.Y[8]));
                k[3] = fma(hw2, F.Y[4], fma(hw1, F.Y[5], fma(-hw0, F.Y[6], ds * fq3))) + fma(hq0, F.Y[10], fma(hq1, F.Y[9], -hq2 * F.Y[8]));
#pragma unroll
                for (int r = 0; r < 4; ++r) { UPD1(4 + r, k[r]) }
            }
            // ---- w rows
            {
                const double2 j01 = ld2(J + J_WW), j23 = ld2(J + J_WW + 2), j45 = ld2(J + J_WW + 4), j67 = ld2(J + J_WW + 6);
                const double j8 = J[J_WW + 8];
                const double g0 = Gc[4], g1 = Gc[5], g2 = Gc[6];
                double k[3];
                k[0] = fma(j01.x, F.Y[8], fma(j01.y, F.Y[9], fma(j23.x, F.Y[10], al * g0)));
                k[1] = fma(j23.y, F.Y[8], fma(j45.x, F.Y[9], fma(j45.y, F.Y[10], al * g1)));
                k[2] = fma(j67.x, F.Y[8], fma(j67.y, F.Y[9], fma(j8, F.Y[10], al * g2)));
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.empty[slot]);       // all reads of the ring slot are done
#pragma unroll
                for (int r = 0; r < 3; ++r) { UPD1(8 + r, k[r]) }
            }
#undef UPD1
            if (last) pca += pcs;
        }

        // ---- epilogue: this lane's column of D, and z
        const int t = g * GROUP + warp * 2 + sub;
        const bool live = t < a.count;
        const int wi = a.first + (live ? t : a.count - 1);
        const int b = wi / ni, i = wi - b * ni;
        double* blk = bt.out_blocks + (size_t)wi * SCVX_BLOCK_DOUBLES;
        double xc = 0.0;
        if (col >= 0) {
            if (col < 14) xc = bt.X[((size_t)b * bt.n_nodes + i) * 14 + col];
            else if (col < 20) xc = bt.U[((size_t)b * bt.n_nodes + i) * 3 + (col - 14)];
            else xc = bt.sigma[b];
        }
        const double cv[14] = { F.S[0], F.Sr[0], F.Sr[1], F.Sr[2], F.S[1], F.S[2], F.S[3], F.S[4], F.S[5], F.S[6], F.S[7],
                                F.S[8], F.S[9], F.S[10] };
        if (live && col >= 0) {
            double* o = blk + 14 * (1 + col);
#pragma unroll
            for (int r = 0; r < 14; r += 2) *reinterpret_cast<double2*>(o + r) = make_double2(cv[r], cv[r + 1]);
        }
        double zlast = 0.0, zprev = 0.0;
#pragma unroll
        for (int r = 0; r < 14; ++r) {
            double v = (col >= 0) ? cv[r] * xc : 0.0;
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            // lane r (of the 16) keeps row r
            if (l16 == r) zlast = v;
            (void)zprev;
        }
        if (live && l16 < 14) {
            double* o = blk + 14 * 22 + l16;
            *o = *o - zlast;                                        // partial z written by kernel A2
        }
    }
}

}  // namespace

size_t scvx_staged_scratch_bytes(int npts, int chunk_intervals) {
    const size_t groups = ((size_t)chunk_intervals + GROUP - 1) / GROUP;
    return groups * (size_t)(4 * npts) * REC_MAX * GROUP * sizeof(double);
}

int scvx_staged_chunk_intervals(int sm_count) { return sm_count * GROUP * 14; }

cudaError_t scvx_launch_staged(const ScvxBatch& bt, const ScvxTables& tb, bool any_aero, void* scratch,
                               int chunk_intervals, int sm_count, cudaStream_t s, int* launches, int lanes) {
    const long total = (long)(bt.n_nodes - 1) * bt.B;
    const size_t smem = (lanes == 8) ? sizeof(TangentSmem) : sizeof(Tangent16Smem);
    {
        cudaError_t e = (lanes == 8)
            ? cudaFuncSetAttribute(tangent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
            : cudaFuncSetAttribute(tangent16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    for (long first = 0; first < total; first += chunk_intervals) {
        StagedArgs a;
        a.bt = bt; a.tb = tb; a.rec = (double*)scratch; a.first = (int)first;
        a.rec_n = any_aero ? REC_AERO : REC_EXO;
        a.count = (int)((total - first < chunk_intervals) ? (total - first) : chunk_intervals);
        a.n_groups = (a.count + GROUP - 1) / GROUP;
        const int threads = a.n_groups * GROUP;
        stage_value_kernel<<<(threads + 127) / 128, 128, 0, s>>>(a);
        light_columns_kernel<<<(a.count + 127) / 128, 128, 0, s>>>(a);
        const int grid = a.n_groups < sm_count ? a.n_groups : sm_count;
        if (lanes == 8) tangent_kernel<<<grid, 256, smem, s>>>(a);
        else tangent16_kernel<<<grid, 512, smem, s>>>(a);
        if (launches) *launches += 3;
    }
    return cudaGetLastError();
}
