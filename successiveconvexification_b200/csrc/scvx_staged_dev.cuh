// scvx_staged_dev.cuh — device-side building blocks of the STAGED linearise-and-discretise path
// (scvx_kernels_staged.cu): Jacobian-record layout, spline value+gradient, aero force Jacobian, the per-stage
// Jacobian producer, mbarrier / TMA helpers and the two-column tangent stage.
#pragma once
#include "scvx_common.cuh"

namespace {


// Stage record (value kernel -> producers).  The value kernel is bound by the HBM WRITE rate of these records (round 2:
// 1.62 GB per chunk in 0.42 ms = 3.9 TB/s), so the record holds only what cannot be re-formed cheaply: the stage state,
// f_v (it contains the aero force) and the aero Jacobians.  The stage control (FOH of the interval's two node controls,
// which the producers keep in shared memory), f_m = -a |u|, f_q = Omega(w) q / 2 and f_w = jBi (rTB x u - w x jB w) are
// re-formed by the producers (~45 FP64 operations per record).
#ifndef SCVX_A_LEAN_LIFT
#define SCVX_A_LEAN_LIFT 1     // leaner lift-branch Jacobian algebra: -44 FP64 operations, +0.55 % (A/B: profiles/r2_ab_lean_lift.txt)
#endif
constexpr int REC_EXO = 14;      // stage record entries: m, v(3), q(4), w(3), f_v(3)
constexpr int REC_AERO = 32;     // + dF_aero/dv (9, row-major) + dF_aero/db (9), b = C(q) e1
constexpr int R_FV = 11, R_AV = 14, R_AB = 23;
constexpr int REC_MAX = REC_AERO;
constexpr int NJ = 78;           // Jacobian record entries per interval per stage (2 x odd: conflict-free STS.128)
constexpr int GROUP = 32;        // intervals per CTA pass
constexpr int NWARP = 8;

// Jacobian record layout (doubles)
// (entries that enter a stage increment carry sigma * c_i, c_i = the stage's rk4 factor; see consume_stage8)
constexpr int J_WW = 0;          // 9  sigma * d(wdot)/dw, row-major
constexpr int J_HW = 9;          // 3  sigma*w/2
constexpr int J_HQ = 12;         // 4  sigma*q/2
constexpr int J_V = 16;          // 3 rows x 8: [d(vdot_r)/dm, d(vdot_r)/dv (3), d(vdot_r)/dq (4)], all times sigma
constexpr int J_G = 40;          // 4 columns (u0,u1,u2,f) x 7 rows (m, v0..2, w0..2)
constexpr int J_FRQ = 68;        // 7  f_r (= v, unscaled) and c_i * f_q (sigma column only)
constexpr int J_SIG = 75;        // 1  sigma (unscaled: the r-row quadrature uses it)

struct StagedArgs {
    ScvxBatch bt;
    ScvxTables tb;
    scvx_probinfo Pc;            // the parameter record itself when all trajectories share one (kernel-argument space:
                                 // its fields become constant-bank operands, no loads; kernels instantiated with SP = true)
    double Kw[9];                // with Pc: jBi * [rTB]x, the d(wdot)/du block (column-major 3 x 3), formed once on the host
    double Tw[27];               // with Pc: d(wdot)/dw = sum_k w_k Tw[9k + 3r + c], Tw_k = -jBi ([e_k]x jB - [jB e_k]x)
    double* rec;                 // stage records of this chunk
    int rec_n;                   // entries per stage record (REC_EXO or REC_AERO)
    int first;                   // first interval (global index) of this chunk (total intervals < 2^31)
    int count;                   // intervals in this chunk
    int n_groups;                // ceil(count / 32)
};

// Parameter access.  SP = the record is shared by every trajectory and sits in the kernel arguments (constant bank):
// plain reads, which the compiler folds into instruction operands.  Otherwise one record per trajectory in global
// memory, read through the read-only path.
template <int SP> __device__ __forceinline__ double ldp(const double* p) {
    if constexpr (SP != 0) return *p; else return __ldg(p);
}
template <int SP> __device__ __forceinline__ int ldpi(const int32_t* p) {
    if constexpr (SP != 0) return *p; else return __ldg(p);
}

// ------------------------------------------------------------------------------------------------
// Kernel A: value trajectory + stage records.
// ------------------------------------------------------------------------------------------------
// spline value and gradient w.r.t. the physical coordinates (gradient 0 when strictly outside: Flat extrapolation);
// one pass over the 16 coefficients serves all three.
#ifndef SCVX_A_BRANCHLESS_LIFT
#define SCVX_A_BRANCHLESS_LIFT 1
#endif
#if SCVX_A_BRANCHLESS_LIFT && !SCVX_A_LEAN_LIFT
#error "SCVX_A_BRANCHLESS_LIFT needs SCVX_A_LEAN_LIFT"
#endif
#ifndef SCVX_A_HOIST_LIFT
#define SCVX_A_HOIST_LIFT 0
#endif
#ifndef SCVX_A_EVICT_LAST
#define SCVX_A_EVICT_LAST 0
#endif
// spline coefficient from global memory through the read-only path; SCVX_A_EVICT_LAST marks the line evict-last in L1
// (A/B: profiles/r2_ab_evict_last.txt)
__device__ __forceinline__ double ld_coef(const double* p) {
#if SCVX_A_EVICT_LAST
    double v;
    asm volatile("ld.global.nc.L1::evict_last.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}

// SMEM: the coefficient array lives in shared memory (plain loads) instead of global memory (read-only path).
// WIN: a WIN_I x WIN_J window of the coefficients around the block's starting cells is staged in shared memory
// (`wcoef`, origin t.wi0 / t.wj0); a 4 x 4 patch inside the window is read from there, any other from global memory.
constexpr int WIN_I = 48, WIN_J = 24;
template <bool SMEM = false, bool WIN = false>
__device__ __forceinline__ void spline_val_grad(const double* __restrict__ coef, const double* __restrict__ wcoef,
                                                const ScvxTables& t, double x, double y,
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
    int stride = L1;
    if constexpr (WIN) {
        const int li = i - 1 - t.wi0, lj = j - 1 - t.wj0;
        if ((unsigned)li <= (unsigned)(WIN_I - 4) && (unsigned)lj <= (unsigned)(WIN_J - 4)) {
            base = wcoef + li + lj * WIN_I;
            stride = WIN_I;
        }
    }
    double av = 0.0, ax = 0.0, ay = 0.0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const double* p = base + (size_t)b * stride;
        double c0, c1, c2, c3;
        if constexpr (SMEM || WIN) { c0 = p[0]; c1 = p[1]; c2 = p[2]; c3 = p[3]; }      // WIN: generic loads (shared or global)
        else { c0 = ld_coef(p); c1 = ld_coef(p + 1); c2 = ld_coef(p + 2); c3 = ld_coef(p + 3); }
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
// TS = tables staged in shared memory: 0 none, 1 drag, 2 drag + lift (tb.drag / tb.lift then point there), 3 a window of both.
template <int TS = 0, int SP = 0>
__device__ __forceinline__ void aero_force_jac(const scvx_probinfo& P, const ScvxTables& tb, const double b[3],
                                               const double v[3], double F[3], double Fv[3][3], double Fb[3][3]) {
    const double vv = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    const double inv = rsqrt(vv), nv = vv * inv;              // one long-latency op instead of sqrt + divide
    const double vh[3] = { v[0] * inv, v[1] * inv, v[2] * inv };
    const double bvdot = b[0] * v[0] + b[1] * v[1] + b[2] * v[2];
    const double dp = bvdot * inv;
    const double bb = b[0] * b[0] + b[1] * b[1] + b[2] * b[2];
    const double inb = rsqrt(bb);
    const double car = dp * inb;
    double ca = car, mc = 1.0;
    if (car > 1.0) { ca = 1.0; mc = 0.0; } else if (car < -1.0) { ca = -1.0; mc = 0.0; }
    const double isos = 1.0 / ldp<SP>(&P.sos);
    const double mach = nv * isos;
    // d(ca)/dv, d(ca)/db ; d(mach)/dv
    double cav[3], cab[3], mv[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        cav[k] = mc * (b[k] - dp * vh[k]) * inv * inb;
        cab[k] = mc * (vh[k] - car * b[k] * inb) * inb;
        mv[k] = vh[k] * isos;
    }
    const double fs = ldp<SP>(&P.force_scalar);
    double drag, gx, gy;
    spline_val_grad<(TS == 1 || TS == 2), (TS == 3)>(tb.drag, tb.wdrag, tb, ca, mach, drag, gx, gy);
    drag *= fs; gx *= fs; gy *= fs;
#if SCVX_A_HOIST_LIFT
    // the lift spline is evaluated here, ahead of the |dp| branch that decides whether it is used: its loads then overlap
    // the drag Jacobian below instead of standing alone behind the branch
    double lift, lgx, lgy;
    spline_val_grad<(TS == 2), (TS == 3)>(tb.lift, tb.wlift, tb, ca, mach, lift, lgx, lgy);
#endif
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
#if SCVX_A_BRANCHLESS_LIFT
    // no branch: the lift terms are always formed and selected at the end, so that the whole stage is one basic block and
    // the scheduler can overlap the lift loads and chains with the rest of the stage (A/B: profiles/r2_ab_branchless_lift.txt)
    const bool take = !(fabs(dp) >= 0.95);
#else
    if (fabs(dp) >= 0.95) return;
#endif
    // (the lift coefficients are read behind this branch with nothing left to overlap their latency: 18 % of the kernel's
    // stall samples.  prefetch.global.L1 of their four rows before the drag evaluation: -6 % LITERAL, -37 % TEXTBOOK,
    // profiles/r2_ab_prefetch_lift.txt)
#if SCVX_A_HOIST_LIFT
    lift *= fs; gx = lgx * fs; gy = lgy * fs;
#else
    double lift;
    spline_val_grad<(TS == 2), (TS == 3)>(tb.lift, tb.wlift, tb, ca, mach, lift, gx, gy);
    lift *= fs; gx *= fs; gy *= fs;
#endif
    double lv[3], lb[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { lv[k] = gx * cav[k] + gy * mv[k]; lb[k] = gx * cab[k]; }
    // l = (-(v x b)) x v = v (v.b) - b (v.v)
    const double l[3] = { v[0] * bvdot - b[0] * vv, v[1] * bvdot - b[1] * vv, v[2] * bvdot - b[2] * vv };
    const double inl = rsqrt(l[0] * l[0] + l[1] * l[1] + l[2] * l[2]);
    const double lh[3] = { l[0] * inl, l[1] * inl, l[2] * inl };
#if SCVX_A_LEAN_LIFT
    // dl/dv = Lv = (v.b) I + v b^T - 2 b v^T ;  dl/db = Lb = v v^T - (v.v) I ;  d(lh)/d. = (I - lh lh^T) L. / |l|, so
    //   dF/dv += lh lv^T + ln (Lv - lh (lh^T Lv)),   dF/db += lh lb^T + ln (Lb - lh (lh^T Lb)),   ln = lift / |l|.
    // The projections are formed from two scalars instead of from the 3 x 3 matrices:
    //   lh^T Lv = (v.b) lh^T + (lh.v) b^T - 2 (lh.b) v^T ,   lh^T Lb = (lh.v) v^T - (v.v) lh^T
    const double ln = lift * inl;
#if SCVX_A_BRANCHLESS_LIFT
#pragma unroll
    for (int r = 0; r < 3; ++r) F[r] = take ? fma(ln, l[r], F[r]) : F[r];
#else
#pragma unroll
    for (int r = 0; r < 3; ++r) F[r] = fma(ln, l[r], F[r]);
#endif
    const double lhv = lh[0] * v[0] + lh[1] * v[1] + lh[2] * v[2];
    const double lhb = lh[0] * b[0] + lh[1] * b[1] + lh[2] * b[2];
    double pjv[3], pjb[3], lnv[3], lnb2[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double pv = fma(bvdot, lh[c], fma(lhv, b[c], -2.0 * lhb * v[c]));
        const double pb = fma(lhv, v[c], -vv * lh[c]);
        pjv[c] = fma(-ln, pv, lv[c]);
        pjb[c] = fma(-ln, pb, lb[c]);
        lnv[c] = ln * v[c];
        lnb2[c] = -2.0 * ln * b[c];
    }
    const double dgv = ln * bvdot, dgb = -ln * vv;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
#if SCVX_A_BRANCHLESS_LIFT
            const double nv_ = Fv[r][c] + fma(lh[r], pjv[c], fma(lnv[r], b[c], fma(lnb2[r], v[c], r == c ? dgv : 0.0)));
            const double nb_ = Fb[r][c] + fma(lh[r], pjb[c], fma(lnv[r], v[c], r == c ? dgb : 0.0));
            Fv[r][c] = take ? nv_ : Fv[r][c];
            Fb[r][c] = take ? nb_ : Fb[r][c];
#else
            Fv[r][c] += fma(lh[r], pjv[c], fma(lnv[r], b[c], fma(lnb2[r], v[c], r == c ? dgv : 0.0)));
            Fb[r][c] += fma(lh[r], pjb[c], fma(lnv[r], v[c], r == c ? dgb : 0.0));
#endif
        }
#else
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
#endif
}

// unscaled f(x,u) (dx_static without the `.* mult`, dynamics.jl:54-77) together with the Jacobians of the
// aerodynamic force w.r.t. v and b = C(q) e1 (zero for the exo-atmospheric variant).
template <bool JAC, int TS = 0, int SP = 0>
__device__ __forceinline__ void rhs_value(const scvx_probinfo& P, double pa, const ScvxTables& tb, const double x[14],
                                          const double u[3], double f[14], double Fv[3][3], double Fb[3][3]) {
    const double q0 = x[7], q1 = x[8], q2 = x[9], q3 = x[10];
    const double w0 = x[11], w1 = x[12], w2 = x[13];
    const double p1 = q1 * q2, p2 = q0 * q3, p3 = q1 * q3, p4 = q0 * q2, p5 = q2 * q3, p6 = q0 * q1;
    const double c00 = 1.0 - 2.0 * (q2 * q2 + q3 * q3), c01 = 2.0 * (p1 - p2), c02 = 2.0 * (p3 + p4);
    const double c10 = 2.0 * (p1 + p2), c11 = 1.0 - 2.0 * (q1 * q1 + q3 * q3), c12 = 2.0 * (p5 - p6);
    const double c20 = 2.0 * (p3 - p4), c21 = 2.0 * (p5 + p6), c22 = 1.0 - 2.0 * (q1 * q1 + q2 * q2);
    double F[3] = { 0.0, 0.0, 0.0 };
    if (ldpi<SP>(&P.aero_kind) == SCVX_AERO_TABLE) {
        const double bv[3] = { c00, c10, c20 };
        if constexpr (JAC) aero_force_jac<TS, SP>(P, tb, bv, x + 4, F, Fv, Fb);
        else aero_force_t<double>(P, tb, bv, x + 4, F);
    } else if constexpr (JAC) {
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) { Fv[r][c] = 0.0; Fb[r][c] = 0.0; }
    }
    const double im = 1.0 / x[0];
    f[0] = -pa * sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    f[1] = x[4]; f[2] = x[5]; f[3] = x[6];
    f[4] = (c00 * u[0] + c01 * u[1] + c02 * u[2] + F[0]) * im - ldp<SP>(&P.g0);
    f[5] = (c10 * u[0] + c11 * u[1] + c12 * u[2] + F[1]) * im;
    f[6] = (c20 * u[0] + c21 * u[1] + c22 * u[2] + F[2]) * im;
    f[7]  = 0.5 * (-(w0 * q1) - w1 * q2 - w2 * q3);
    f[8]  = 0.5 * (w0 * q0 + w2 * q2 - w1 * q3);
    f[9]  = 0.5 * (w1 * q0 - w2 * q1 + w0 * q3);
    f[10] = 0.5 * (w2 * q0 + w1 * q1 - w0 * q2);
    const double h0 = ldp<SP>(&P.jB[0]) * w0 + ldp<SP>(&P.jB[3]) * w1 + ldp<SP>(&P.jB[6]) * w2;
    const double h1 = ldp<SP>(&P.jB[1]) * w0 + ldp<SP>(&P.jB[4]) * w1 + ldp<SP>(&P.jB[7]) * w2;
    const double h2 = ldp<SP>(&P.jB[2]) * w0 + ldp<SP>(&P.jB[5]) * w1 + ldp<SP>(&P.jB[8]) * w2;
    const double m0 = (ldp<SP>(&P.rTB[1]) * u[2] - ldp<SP>(&P.rTB[2]) * u[1]) - (w1 * h2 - w2 * h1);
    const double m1 = (ldp<SP>(&P.rTB[2]) * u[0] - ldp<SP>(&P.rTB[0]) * u[2]) - (w2 * h0 - w0 * h2);
    const double m2 = (ldp<SP>(&P.rTB[0]) * u[1] - ldp<SP>(&P.rTB[1]) * u[0]) - (w0 * h1 - w1 * h0);
    f[11] = ldp<SP>(&P.jBi[0]) * m0 + ldp<SP>(&P.jBi[3]) * m1 + ldp<SP>(&P.jBi[6]) * m2;
    f[12] = ldp<SP>(&P.jBi[1]) * m0 + ldp<SP>(&P.jBi[4]) * m1 + ldp<SP>(&P.jBi[7]) * m2;
    f[13] = ldp<SP>(&P.jBi[2]) * m0 + ldp<SP>(&P.jBi[5]) * m1 + ldp<SP>(&P.jBi[8]) * m2;
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
// arrive from one elected lane WITHOUT a divergent branch (predicated instruction: no BSSY/BSYNC reconvergence)
__device__ __forceinline__ void mbar_arrive_lane0(uint64_t* bar, int lane) {
    asm volatile("{\n .reg .pred p;\n .reg .b64 st;\n setp.eq.s32 p, %1, 0;\n @p mbarrier.arrive.shared::cta.b64 st, [%0];\n }"
                 ::"r"(smem_u32(bar)), "r"(lane) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n }" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// SCVX_MBAR_HINT: suspend-time hint (ns) on try_wait — the warp stays suspended until the phase completes or the time is
// up, instead of re-issuing the instruction (A/B: profiles/r2_ab_mbar_wait.txt)
#ifndef SCVX_MBAR_HINT
#define SCVX_MBAR_HINT 0
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#if SCVX_MBAR_HINT
    asm volatile(
        "{\n"
        " .reg .pred p;\n"
        "WAIT_LOOP:\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        " @p bra WAIT_DONE;\n"
        " bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"((uint32_t)SCVX_MBAR_HINT) : "memory");
#else
    asm volatile(
        "{\n"
        " .reg .pred p;\n"
        "WAIT_LOOP:\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        " @p bra WAIT_DONE;\n"
        " bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
#endif
}
// Wait of a warp that is usually far ahead (the producers, on a free ring slot): back off between polls.  A polling
// producer re-issued try_wait 35 times per step — 7.4 % of the instructions of the kernel (ncu source view), one polling
// warp on every scheduler beside the two consumer warps it was waiting for.  With the back-off the step is 2.5 % faster;
// 200 ns to 3 us all measure the same, a suspend-time hint on try_wait does not help (profiles/r2_ab_mbar_wait*.txt).
#ifndef SCVX_PRODUCER_SLEEP_NS
#define SCVX_PRODUCER_SLEEP_NS 200
#endif
template <int NS = SCVX_PRODUCER_SLEEP_NS>
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
    if constexpr (NS > 0) {
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n"
            " .reg .pred p;\n"
            " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            " selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return;
        __nanosleep(NS);
    }
    } else {
        mbar_wait(bar, parity);
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// Producer role: Jacobian blocks of one stage for 32 intervals (lane = interval).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void st2(double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }

// Register-lean producer for the dedicated producer warps of tangent_kernel (they run with 88 registers after the
// setmaxnreg re-partition).  rec: this lane's stage record (stride GROUP doubles between entries, in shared memory);
// out: this lane's NJ-double Jacobian record in the ring (the caller has waited for the slot).  The record is formed
// and stored block by block, so only a few values are live at any time; every block that enters a stage increment
// carries sigma * scale (scale = the stage's rk4 factor, see consume_stage8), the quadrature entries (v, sigma) do not.
template <int SP = 0>
__device__ __forceinline__ void produce_lean(const scvx_probinfo& P, double pa, const double* __restrict__ Kw,
                                             const double* __restrict__ Tw, bool aero_rec, double sigma,
                                             double scale, const double* __restrict__ rec, const double* __restrict__ unode,
                                             double pc, double* __restrict__ out) {
    const double ss = sigma * scale;
    const double hs = 0.5 * ss;
    const double sm = ss / rec[0];
    // ---- quadrature entries, hw, hq; f_q = Omega(w) q / 2 is re-formed here (dynamics.jl:46-52, 68)
    {
        const double q0 = rec[4 * GROUP], q1 = rec[5 * GROUP], q2 = rec[6 * GROUP], q3 = rec[7 * GROUP];
        st2(out + J_HQ + 0, hs * q0, hs * q1); st2(out + J_HQ + 2, hs * q2, hs * q3);
        const double v0 = rec[1 * GROUP], v1 = rec[2 * GROUP], v2 = rec[3 * GROUP];
        const double w0 = rec[8 * GROUP], w1 = rec[9 * GROUP], w2 = rec[10 * GROUP];
        const double hc = 0.5 * scale;
        const double f0 = hc * (-(w0 * q1) - w1 * q2 - w2 * q3);
        const double f1 = hc * (w0 * q0 + w2 * q2 - w1 * q3);
        const double f2 = hc * (w1 * q0 - w2 * q1 + w0 * q3);
        const double f3 = hc * (w2 * q0 + w1 * q1 - w0 * q2);
        st2(out + J_FRQ + 0, v0, v1); st2(out + J_FRQ + 2, v2, f0);
        st2(out + J_FRQ + 4, f1, f2); st2(out + J_FRQ + 6, f3, sigma);
        st2(out + J_FRQ + 8, 0.0, 0.0);
    }
    // ---- v rows: [d/dm, d/dv (3), d/dq (4)] per row.  The factor 2 of d(C u)/dq and of db/dq is folded into sm2 = 2 sm;
    // the structurally zero entries of db/dq (row 0: {0, 0, -4 q2, -4 q3}) are not multiplied out.
    {
        const double q0 = rec[4 * GROUP], q1 = rec[5 * GROUP], q2 = rec[6 * GROUP], q3 = rec[7 * GROUP];
        // stage control: current_control (dynamics.jl:108-110) of the interval's node controls (unode: u-(3), u+(3), stride GROUP)
        const double opc = 1.0 - pc;
        const double u0 = opc * unode[0 * GROUP] + pc * unode[3 * GROUP];
        const double u1 = opc * unode[1 * GROUP] + pc * unode[4 * GROUP];
        const double u2 = opc * unode[2 * GROUP] + pc * unode[5 * GROUP];
        const double Pg0 = ldp<SP>(&P.g0);
        const double sm2 = sm + sm;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            // half of d(C u)_r / dq
            double j0, j1, j2, j3;
            if (r == 0) {
                j0 = q2 * u2 - q3 * u1;                     j1 = q2 * u1 + q3 * u2;
                j2 = fma(-2.0 * q2, u0, q1 * u1 + q0 * u2); j3 = fma(-2.0 * q3, u0, q1 * u2 - q0 * u1);
            } else if (r == 1) {
                j0 = q3 * u0 - q1 * u2;                     j1 = fma(-2.0 * q1, u1, q2 * u0 - q0 * u2);
                j2 = q1 * u0 + q3 * u2;                     j3 = fma(-2.0 * q3, u1, q0 * u0 + q2 * u2);
            } else {
                j0 = q1 * u1 - q2 * u0;                     j1 = fma(-2.0 * q1, u2, q3 * u0 + q0 * u1);
                j2 = fma(-2.0 * q2, u2, q3 * u1 - q0 * u0); j3 = q1 * u0 + q2 * u1;
            }
            double vv0 = 0.0, vv1 = 0.0, vv2 = 0.0;
            if (aero_rec) {
                const double b0 = rec[(R_AB + 3 * r) * GROUP], b1 = rec[(R_AB + 1 + 3 * r) * GROUP], b2 = rec[(R_AB + 2 + 3 * r) * GROUP];
                // half of dF_r/db * db/dq, db/dq rows: {0, 0, -4 q2, -4 q3}, {2 q3, 2 q2, 2 q1, 2 q0}, {-2 q2, 2 q3, -2 q0, 2 q1}
                j0 = fma(b1, q3, fma(-b2, q2, j0));
                j1 = fma(b1, q2, fma(b2, q3, j1));
                j2 = fma(-2.0 * b0, q2, fma(b1, q1, fma(-b2, q0, j2)));
                j3 = fma(-2.0 * b0, q3, fma(b1, q0, fma(b2, q1, j3)));
                vv0 = sm * rec[(R_AV + 3 * r) * GROUP]; vv1 = sm * rec[(R_AV + 1 + 3 * r) * GROUP]; vv2 = sm * rec[(R_AV + 2 + 3 * r) * GROUP];
            }
            const double fvr = rec[(R_FV + r) * GROUP];
            st2(out + J_V + 8 * r + 0, -sm * (fvr + (r == 0 ? Pg0 : 0.0)), vv0);
            st2(out + J_V + 8 * r + 2, vv1, vv2);
            st2(out + J_V + 8 * r + 4, sm2 * j0, sm2 * j1);
            st2(out + J_V + 8 * r + 6, sm2 * j2, sm2 * j3);
        }
        // ---- direct (control / sigma) columns G[col][row], rows m, v0..2, w0..2 (7 per column)
        const double c00 = 1.0 - 2.0 * (q2 * q2 + q3 * q3), c01 = 2.0 * (q1 * q2 - q0 * q3), c02 = 2.0 * (q1 * q3 + q0 * q2);
        const double c10 = 2.0 * (q1 * q2 + q0 * q3), c11 = 1.0 - 2.0 * (q1 * q1 + q3 * q3), c12 = 2.0 * (q2 * q3 - q0 * q1);
        const double c20 = 2.0 * (q1 * q3 - q0 * q2), c21 = 2.0 * (q2 * q3 + q0 * q1), c22 = 1.0 - 2.0 * (q1 * q1 + q2 * q2);
        const double nu = sqrt(u0 * u0 + u1 * u1 + u2 * u2);
        const double gm = -ss * pa / nu;
        const double r0 = ldp<SP>(&P.rTB[0]), r1 = ldp<SP>(&P.rTB[1]), r2 = ldp<SP>(&P.rTB[2]);
        double* G = out + J_G;
        // jBi * (rTB x e_j): rTB x e0 = (0, r2, -r1); x e1 = (-r2, 0, r0); x e2 = (r1, -r0, 0) — constant per parameter
        // record: taken from the kernel arguments when the record is shared (SP), formed here otherwise
        double kw[9];
        if constexpr (SP != 0) {
#pragma unroll
            for (int k = 0; k < 9; ++k) kw[k] = Kw[k];
        } else {
            double bi[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) bi[k] = ldp<SP>(&P.jBi[k]);
            kw[0] = bi[3] * r2 - bi[6] * r1; kw[1] = bi[4] * r2 - bi[7] * r1; kw[2] = bi[5] * r2 - bi[8] * r1;
            kw[3] = bi[6] * r0 - bi[0] * r2; kw[4] = bi[7] * r0 - bi[1] * r2; kw[5] = bi[8] * r0 - bi[2] * r2;
            kw[6] = bi[0] * r1 - bi[3] * r0; kw[7] = bi[1] * r1 - bi[4] * r0; kw[8] = bi[2] * r1 - bi[5] * r0;
        }
        st2(G + 0, gm * u0, sm * c00); st2(G + 2, sm * c10, sm * c20);
        st2(G + 4, ss * kw[0], ss * kw[1]);
        st2(G + 6, ss * kw[2], gm * u1);
        st2(G + 8, sm * c01, sm * c11);
        st2(G + 10, sm * c21, ss * kw[3]);
        st2(G + 12, ss * kw[4], ss * kw[5]);
        st2(G + 14, gm * u2, sm * c02); st2(G + 16, sm * c12, sm * c22);
        st2(G + 18, ss * kw[6], ss * kw[7]);
        // ---- rotational block: Jww = -ss * jBi * ([w]x jB - [jB w]x), and hw.  The block is linear in w; with a shared
        // parameter record its three constant 3 x 3 factors come from the kernel arguments (27 FMAs instead of 54 operations).
        double fwq[3];
        {
            const double w0 = rec[8 * GROUP], w1 = rec[9 * GROUP], w2 = rec[10 * GROUP];
            // fwq = -jBi (w x jB w), the rate-dependent part of f_w (the producer adds jBi (rTB x u) below).  d(wdot)/dw is
            // linear in w and  (d(wdot)/dw) w = -2 jBi (w x jB w),  so fwq = (Jww / ss) w / 2
            if constexpr (SP != 0) {
                const double s0 = ss * w0, s1 = ss * w1, s2 = ss * w2;
                auto jw = [&](int k) { return fma(s2, Tw[18 + k], fma(s1, Tw[9 + k], s0 * Tw[k])); };
                const double j0 = jw(0), j1 = jw(1), j2 = jw(2), j3 = jw(3), j4 = jw(4), j5 = jw(5), j6 = jw(6), j7 = jw(7), j8 = jw(8);
                st2(out + J_WW + 0, j0, j1); st2(out + J_WW + 2, j2, j3); st2(out + J_WW + 4, j4, j5);
                st2(out + J_WW + 6, j6, j7); st2(out + J_WW + 8, j8, hs * w0);
                const double hi = 0.5 / ss;
                fwq[0] = hi * fma(j2, w2, fma(j1, w1, j0 * w0));
                fwq[1] = hi * fma(j5, w2, fma(j4, w1, j3 * w0));
                fwq[2] = hi * fma(j8, w2, fma(j7, w1, j6 * w0));
            } else {
                double Jw[9];
                double M[3][3];
                {
                    double jB[9];
    #pragma unroll
                    for (int k = 0; k < 9; ++k) jB[k] = ldp<SP>(&P.jB[k]);
                    const double L0 = jB[0] * w0 + jB[3] * w1 + jB[6] * w2;
                    const double L1 = jB[1] * w0 + jB[4] * w1 + jB[7] * w2;
                    const double L2 = jB[2] * w0 + jB[5] * w1 + jB[8] * w2;
    #pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const double a0 = jB[3 * c], a1 = jB[3 * c + 1], a2 = jB[3 * c + 2];
                        M[0][c] = w1 * a2 - w2 * a1; M[1][c] = w2 * a0 - w0 * a2; M[2][c] = w0 * a1 - w1 * a0;
                    }
                    M[0][1] += L2; M[0][2] -= L1; M[1][0] -= L2; M[1][2] += L0; M[2][0] += L1; M[2][1] -= L0;
                }
    #pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const double b0 = ldp<SP>(&P.jBi[r]), b1 = ldp<SP>(&P.jBi[r + 3]), b2 = ldp<SP>(&P.jBi[r + 6]);
    #pragma unroll
                    for (int c = 0; c < 3; ++c) Jw[3 * r + c] = -ss * (b0 * M[0][c] + b1 * M[1][c] + b2 * M[2][c]);
                }
                st2(out + J_WW + 0, Jw[0], Jw[1]); st2(out + J_WW + 2, Jw[2], Jw[3]); st2(out + J_WW + 4, Jw[4], Jw[5]);
                st2(out + J_WW + 6, Jw[6], Jw[7]); st2(out + J_WW + 8, Jw[8], hs * w0);
                const double hi = 0.5 / ss;
                fwq[0] = hi * fma(Jw[2], w2, fma(Jw[1], w1, Jw[0] * w0));
                fwq[1] = hi * fma(Jw[5], w2, fma(Jw[4], w1, Jw[3] * w0));
                fwq[2] = hi * fma(Jw[8], w2, fma(Jw[7], w1, Jw[6] * w0));
            }
            st2(out + J_HW + 1, hs * w1, hs * w2);
        }
        // sigma column (the unscaled rhs): f_m = -a |u|, f_v from the record, f_w = jBi (rTB x u) + fwq
        const double fw0 = fma(kw[6], u2, fma(kw[3], u1, fma(kw[0], u0, fwq[0])));
        const double fw1 = fma(kw[7], u2, fma(kw[4], u1, fma(kw[1], u0, fwq[1])));
        const double fw2 = fma(kw[8], u2, fma(kw[5], u1, fma(kw[2], u0, fwq[2])));
        st2(G + 20, ss * kw[8], -(scale * pa) * nu);
        st2(G + 22, scale * rec[R_FV * GROUP], scale * rec[(R_FV + 1) * GROUP]);
        st2(G + 24, scale * rec[(R_FV + 2) * GROUP], scale * fw0);
        st2(G + 26, scale * fw1, scale * fw2);
    }
}

__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }

// one full tangent column: rows m, v(3), q(4), w(3) carried as (S, acc, Y); r rows as a pure quadrature
struct FullCol {
    double S[11], A[11], Y[11], Sr[3];
};


// One rk4 stage of TWO full tangent columns (8-lane variant).  The position of the stage in its step is a compile-time
// parameter (KIND below), so there is no control flow inside the stage; the caller unrolls the four stages of a step.
// Rows are processed in cascade order (r, v, m, q, w): a row block is overwritten only after every block that
// reads its old stage value has been formed.
//
// Slot roles.  Slot A of every lane carries a CONTROL-type column (B- / B+ / Sigma, or the light column d/dm): the only
// columns with a direct term alpha * G and a non-zero mass row.  Slot B carries a STATE-type column (d/dw, d/dq, or the
// light column d/dv0): mdot does not depend on the state, so its mass row is identically zero and it has no direct
// term — the code of slot B omits the J_vm * Y_m products, the G terms and the whole m row (11 of 163 FMAs per stage).
//
// Stage-tangent form of rk4.  The producer scales the Jacobian blocks of stage i by c_i (the factor of
// Y_{i+1} = S + c_i K_i; h/6 for the final stage), so a row's chain of FMAs started FROM S yields the next stage
// tangent directly (no separate K, no multiply to start the chain).  With T = Y_2 + 2 Y_3 + Y_4 and kappa = h/(3 s)
// (s = 1 LITERAL, h TEXTBOOK):   S_new = S + kappa (T - 4 S) + (h/6) K_4,  the last term again a chain started from
// the first two.  T - 4S cancels exactly for entries that never move, so constants of D stay exact.
// 11 % fewer FP64 instructions per stage than forming K, acc += w K, Y = S + c K.
// KIND: 0 = first stage of a step (Y == S: the stage reads S, and A is written rather than accumulated, so that the last
// stage of the previous step neither copies S into Y nor clears A), 1 = middle stage, 2 = last stage, 3 = last stage that
// also sets Y = S and A = 0 (for a step whose first stage runs the middle-stage body: SCVX_T_FIRST_BODY = 0).
template <int KIND>
__device__ __forceinline__ void consume_stage8(FullCol& FA, FullCol& FB, const double* __restrict__ J, const int gcol,
                                               const double alA, const double dsA, const double tw, const double cr,
                                               const double kappa, uint64_t* empty_bar, const int lane) {
    constexpr bool last = (KIND >= 2), first = (KIND == 0);
    const double* const YA = first ? FA.S : FA.Y;
    const double* const YB = first ? FB.S : FB.Y;
    // slot A: lanes 0..2 B- (alA = 1 - pc), 3..5 B+ (alA = pc), 6 Sigma (direct term = the f column, alA = 1, dsA = 1),
    // 7 the light column d/dm (no direct term): alA = cA0 + cA1 * pc with per-lane constants, formed by the caller
    const double* Gc = J + J_G + 7 * gcol;
    auto init = [&](const FullCol& F, const int idx) -> double {
        if constexpr (!last) return F.S[idx];
        else return fma(kappa, fma(-4.0, F.S[idx], F.A[idx]), F.S[idx]);
    };
    auto fin = [&](FullCol& F, const int idx, const double yn) {
        if constexpr (first) { F.A[idx] = tw * yn; F.Y[idx] = yn; }
        else if constexpr (!last) { F.A[idx] = fma(tw, yn, F.A[idx]); F.Y[idx] = yn; }
        else { F.S[idx] = yn; if constexpr (KIND == 3) { F.Y[idx] = yn; F.A[idx] = 0.0; } }
    };
    // ---- r rows (pure quadrature): S_r += cr * (sigma * Y_v + dsigma * f_r),  cr = h/6 * rk4 weight
    {
        const double2 fr01 = ld2(J + J_FRQ);
        const double fr2 = J[J_FRQ + 2];
        const double sg = J[J_FRQ + 7];
        const double csg = cr * sg, cds = cr * dsA;
        FA.Sr[0] = fma(csg, YA[1], fma(cds, fr01.x, FA.Sr[0]));
        FA.Sr[1] = fma(csg, YA[2], fma(cds, fr01.y, FA.Sr[1]));
        FA.Sr[2] = fma(csg, YA[3], fma(cds, fr2, FA.Sr[2]));
#pragma unroll
        for (int r = 0; r < 3; ++r) FB.Sr[r] = fma(csg, YB[1 + r], FB.Sr[r]);
    }
    // ---- v rows: Jvm Y_m + Jvv Y_v + Jvq Y_q + alpha * G_v
    {
        double kA[3], kB[3];
#pragma unroll
        for (int row = 0; row < 3; ++row) {
            const double2 c01 = ld2(J + J_V + 8 * row), c23 = ld2(J + J_V + 8 * row + 2);
            const double2 qa = ld2(J + J_V + 8 * row + 4), qb = ld2(J + J_V + 8 * row + 6);
            const double gg = Gc[1 + row];
            kA[row] = fma(c01.x, YA[0], fma(c01.y, YA[1], fma(c23.x, YA[2], fma(c23.y, YA[3],
                      fma(qa.x, YA[4], fma(qa.y, YA[5], fma(qb.x, YA[6], fma(qb.y, YA[7], fma(alA, gg, init(FA, 1 + row))))))))));
            kB[row] = fma(c01.y, YB[1], fma(c23.x, YB[2], fma(c23.y, YB[3],
                      fma(qa.x, YB[4], fma(qa.y, YB[5], fma(qb.x, YB[6], fma(qb.y, YB[7], init(FB, 1 + row))))))));
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) { fin(FA, 1 + r, kA[r]); fin(FB, 1 + r, kB[r]); }
    }
    // ---- m row: alpha * G_m (slot A only: the mass row of a state-type column is identically zero)
    {
        const double gmv = Gc[0];
        const double a0 = fma(alA, gmv, init(FA, 0));
        fin(FA, 0, a0);
    }
    // ---- q rows: Omega(hw) Y_q + Omega(Y_w) hq + dsigma * f_q   (only slot A can hold the sigma column)
    {
        const double hw0 = J[J_HW], hw1 = J[J_HW + 1], hw2 = J[J_HW + 2];
        const double2 hq01 = ld2(J + J_HQ), hq23 = ld2(J + J_HQ + 2);
        const double hq0 = hq01.x, hq1 = hq01.y, hq2 = hq23.x, hq3 = hq23.y;
        const double fq0 = J[J_FRQ + 3];
        const double2 fq12 = ld2(J + J_FRQ + 4);
        const double fq3 = J[J_FRQ + 6];
        double kA[4], kB[4];
#define QROWS(Yp, K, i0, i1, i2, i3)                                                                                                  \
        K[0] = fma(-hw0, Yp[5], fma(-hw1, Yp[6], fma(-hw2, Yp[7], fma(-hq1, Yp[8], fma(-hq2, Yp[9], fma(-hq3, Yp[10], i0))))));   \
        K[1] = fma(hw0, Yp[4], fma(hw2, Yp[6], fma(-hw1, Yp[7], fma(hq0, Yp[8], fma(hq2, Yp[10], fma(-hq3, Yp[9], i1))))));       \
        K[2] = fma(hw1, Yp[4], fma(-hw2, Yp[5], fma(hw0, Yp[7], fma(hq0, Yp[9], fma(-hq1, Yp[10], fma(hq3, Yp[8], i2))))));       \
        K[3] = fma(hw2, Yp[4], fma(hw1, Yp[5], fma(-hw0, Yp[6], fma(hq0, Yp[10], fma(hq1, Yp[9], fma(-hq2, Yp[8], i3))))));
        QROWS(YA, kA, fma(dsA, fq0, init(FA, 4)), fma(dsA, fq12.x, init(FA, 5)), fma(dsA, fq12.y, init(FA, 6)), fma(dsA, fq3, init(FA, 7)))
        QROWS(YB, kB, init(FB, 4), init(FB, 5), init(FB, 6), init(FB, 7))
#undef QROWS
#pragma unroll
        for (int r = 0; r < 4; ++r) { fin(FA, 4 + r, kA[r]); fin(FB, 4 + r, kB[r]); }
    }
    // ---- w rows: Jww * Y_w + alpha * G_w
    {
        const double2 j01 = ld2(J + J_WW), j23 = ld2(J + J_WW + 2), j45 = ld2(J + J_WW + 4), j67 = ld2(J + J_WW + 6);
        const double j8 = J[J_WW + 8];
        const double g0 = Gc[4], g1 = Gc[5], g2 = Gc[6];
        double kA[3], kB[3];
        kA[0] = fma(j01.x, YA[8], fma(j01.y, YA[9], fma(j23.x, YA[10], fma(alA, g0, init(FA, 8)))));
        kA[1] = fma(j23.y, YA[8], fma(j45.x, YA[9], fma(j45.y, YA[10], fma(alA, g1, init(FA, 9)))));
        kA[2] = fma(j67.x, YA[8], fma(j67.y, YA[9], fma(j8, YA[10], fma(alA, g2, init(FA, 10)))));
        kB[0] = fma(j01.x, YB[8], fma(j01.y, YB[9], fma(j23.x, YB[10], init(FB, 8))));
        kB[1] = fma(j23.y, YB[8], fma(j45.x, YB[9], fma(j45.y, YB[10], init(FB, 9))));
        kB[2] = fma(j67.x, YB[8], fma(j67.y, YB[9], fma(j8, YB[10], init(FB, 10))));
        // all reads of the ring slot are done: hand it back (per-stage hand-over only; the step-synchronised kernel
        // passes a null barrier and releases a whole step at once)
        if (empty_bar != nullptr) {
            __syncwarp();
            mbar_arrive_lane0(empty_bar, lane);
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) { fin(FA, 8 + r, kA[r]); fin(FB, 8 + r, kB[r]); }
    }
}

}  // namespace
