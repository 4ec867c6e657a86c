// scvx_common.cuh — shared device-side definitions of the linearise-and-discretise kernels (sm_100a, FP64).
//
// Reference semantics implemented here (file:line into the reference repository):
//   DCM                 dynamics.jl:29-44        Omega              dynamics.jl:46-52
//   dx_static           dynamics.jl:54-77        current_control    dynamics.jl:108-110
//   rk4                 dynamics.jl:112-134      aero_force         aerodynamics.jl:38-58
//   spline tables       aerodynamics.jl:17-21 (Interpolations.jl cubic B-spline, Line/OnGrid, Flat)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/scvx_b200.h"

#define SCVX_NX 14
#define SCVX_NU 3
#define SCVX_NINP 21
#define SCVX_BLOCK_COLS 23
#define SCVX_BLOCK_DOUBLES (SCVX_NX * SCVX_BLOCK_COLS)   // 322 doubles = 2576 B per interval

// Prefiltered cubic-B-spline coefficient tables staged in device memory.
// coef layout: (n1+2) x (n2+2) column-major; coefficient with padded grid index g (0..n+1) sits at offset g.
struct ScvxTables {
    const double* drag;
    const double* lift;
    const double* trq;    // torque table: only the fins variant reads it
    // shared-memory window of the drag / lift coefficients (value kernel of the STAGED path): WIN_I x WIN_J padded
    // coefficients starting at (wi0, wj0); nullptr / unused elsewhere
    const double* wdrag;
    const double* wlift;
    int wi0, wj0;
    int n1, n2;           // n_cos, n_mach
    double x0, inv_dx;    // cos axis:  index coordinate = (x - x0) * inv_dx + 1
    double y0, inv_dy;    // mach axis
};

// One launch worth of work: B trajectories x (n_nodes-1) intervals.
struct ScvxBatch {
    const double* X;            // 14 x n_nodes x B
    const double* U;            //  3 x n_nodes x B
    const double* sigma;        //  B
    const scvx_probinfo* P;     //  n_params records
    int n_params;               //  1 (shared) or B
    int n_nodes;
    int B;
    double dt;                  // base_dt
    int npts;                   // RK4 sub-steps (reference default 10, dynamics.jl:112)
    int mode;                   // SCVX_MODE_*
    double* out_blocks;         // 14 x 23 x (n_nodes-1) x B
    double* out_lin_err;        // 14 x (n_nodes-1) x B   or nullptr
    double* out_tlb;            //  4 x n_nodes x B       or nullptr
    double* out_endpoints;      // 14 x (n_nodes-1) x B   (predict kernel)
};

// ---------------------------------------------------------------------------------------------
// Scalar types: plain double, and a dual number with ONE tangent (value + one directional
// derivative).  The DUALWARP kernel gives each lane of a warp its own seed direction.
// ---------------------------------------------------------------------------------------------
struct D1 {
    double v, d;
    __device__ __forceinline__ D1() {}
    __device__ __forceinline__ D1(double a) : v(a), d(0.0) {}
    __device__ __forceinline__ D1(double a, double b) : v(a), d(b) {}
};
__device__ __forceinline__ D1 operator+(D1 a, D1 b) { return D1(a.v + b.v, a.d + b.d); }
__device__ __forceinline__ D1 operator-(D1 a, D1 b) { return D1(a.v - b.v, a.d - b.d); }
__device__ __forceinline__ D1 operator-(D1 a) { return D1(-a.v, -a.d); }
__device__ __forceinline__ D1 operator*(D1 a, D1 b) { return D1(a.v * b.v, fma(a.d, b.v, a.v * b.d)); }
__device__ __forceinline__ D1 operator/(D1 a, D1 b) {
    const double ib = 1.0 / b.v; const double r = a.v * ib; return D1(r, (a.d - r * b.d) * ib); }
__device__ __forceinline__ D1 operator+(D1 a, double b) { return D1(a.v + b, a.d); }
__device__ __forceinline__ D1 operator+(double a, D1 b) { return D1(a + b.v, b.d); }
__device__ __forceinline__ D1 operator-(D1 a, double b) { return D1(a.v - b, a.d); }
__device__ __forceinline__ D1 operator-(double a, D1 b) { return D1(a - b.v, -b.d); }
__device__ __forceinline__ D1 operator*(D1 a, double b) { return D1(a.v * b, a.d * b); }
__device__ __forceinline__ D1 operator*(double a, D1 b) { return D1(a * b.v, a * b.d); }
__device__ __forceinline__ D1 sqrt_t(D1 a) { const double r = sqrt(a.v); return D1(r, a.d * (0.5 / r)); }
__device__ __forceinline__ double sqrt_t(double a) { return sqrt(a); }
__device__ __forceinline__ double val(double a) { return a; }
__device__ __forceinline__ double val(D1 a) { return a.v; }

// Base.clamp semantics: the argument itself (tangent kept) unless STRICTLY outside.
template <class T> __device__ __forceinline__ T clamp_strict(T x, double lo, double hi) {
    if (val(x) > hi) return T(hi);
    if (val(x) < lo) return T(lo);
    return x;
}

// ---------------------------------------------------------------------------------------------
// Flat-extrapolated, scaled tensor-product cubic B-spline (value; tangent flows through T).
// ---------------------------------------------------------------------------------------------
template <class T>
__device__ __forceinline__ void bspline_weights(T d, T w[4]) {
    const T o = 1.0 - d;
    const T d2 = d * d, o2 = o * o;
    w[0] = (o2 * o) * (1.0 / 6.0);
    w[1] = (2.0 / 3.0) - d2 + (d2 * d) * 0.5;
    w[2] = (2.0 / 3.0) - o2 + (o2 * o) * 0.5;
    w[3] = (d2 * d) * (1.0 / 6.0);
}

template <class T>
__device__ __forceinline__ T spline_eval(const double* __restrict__ coef, const ScvxTables& t, T x, T y) {
    const int L1 = t.n1 + 2;
    T xi = clamp_strict((x - t.x0) * t.inv_dx + 1.0, 1.0, (double)t.n1);
    T yi = clamp_strict((y - t.y0) * t.inv_dy + 1.0, 1.0, (double)t.n2);
    int i = (int)floor(val(xi)); i = min(i, t.n1 - 1); i = max(i, 1);
    int j = (int)floor(val(yi)); j = min(j, t.n2 - 1); j = max(j, 1);
    T wx[4], wy[4];
    bspline_weights(xi - (double)i, wx);
    bspline_weights(yi - (double)j, wy);
    const double* base = coef + (i - 1) + (size_t)(j - 1) * L1;
    T acc(0.0);
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const double* p = base + (size_t)b * L1;
        T row = wx[0] * __ldg(p) + wx[1] * __ldg(p + 1) + wx[2] * __ldg(p + 2) + wx[3] * __ldg(p + 3);
        acc = acc + wy[b] * row;
    }
    return acc;
}

// ---------------------------------------------------------------------------------------------
// Right-hand side  F(x,u) = sigma * f(x,u)   (dx_static), generic in the scalar type.
// ---------------------------------------------------------------------------------------------
template <class T>
__device__ __forceinline__ void aero_force_t(const scvx_probinfo& P, const ScvxTables& tb, const T bv[3],
                                             const T v[3], T F[3]) {
    const T nv = sqrt_t(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    const T inv_nv = T(1.0) / nv;
    const T dp = (bv[0] * v[0] + bv[1] * v[1] + bv[2] * v[2]) * inv_nv;
    const T nb = sqrt_t(bv[0] * bv[0] + bv[1] * bv[1] + bv[2] * bv[2]);
    const T cosa = clamp_strict(dp / nb, -1.0, 1.0);
    const T mach = nv * (1.0 / P.sos);
    const T drag = spline_eval(tb.drag, tb, cosa, mach) * P.force_scalar;
    const T dn = drag * inv_nv;
    F[0] = dn * v[0]; F[1] = dn * v[1]; F[2] = dn * v[2];
    if (fabs(val(dp)) >= 0.95) return;                       // aerodynamics.jl:42-45
    const T lift = spline_eval(tb.lift, tb, cosa, mach) * P.force_scalar;
    // trqd = v x bv ; liftd = (-trqd) x v, normalised      aerodynamics.jl:50-52
    T t[3] = { v[1] * bv[2] - v[2] * bv[1], v[2] * bv[0] - v[0] * bv[2], v[0] * bv[1] - v[1] * bv[0] };
    T l[3] = { -(t[1] * v[2] - t[2] * v[1]), -(t[2] * v[0] - t[0] * v[2]), -(t[0] * v[1] - t[1] * v[0]) };
    const T ln = lift / sqrt_t(l[0] * l[0] + l[1] * l[1] + l[2] * l[2]);
    F[0] = F[0] + ln * l[0]; F[1] = F[1] + ln * l[1]; F[2] = F[2] + ln * l[2];
}

template <class T>
__device__ __forceinline__ void rhs_t(const scvx_probinfo& P, const ScvxTables& tb, const T x[14], const T u[3],
                                      T sigma, T out[14]) {
    const T q0 = x[7], q1 = x[8], q2 = x[9], q3 = x[10];
    const T w0 = x[11], w1 = x[12], w2 = x[13];
    // DCM (dynamics.jl:29-44), no normalisation of q
    const T p1 = q1 * q2, p2 = q0 * q3, p3 = q1 * q3, p4 = q0 * q2, p5 = q2 * q3, p6 = q0 * q1;
    const T c00 = 1.0 - 2.0 * (q2 * q2 + q3 * q3), c01 = 2.0 * (p1 - p2), c02 = 2.0 * (p3 + p4);
    const T c10 = 2.0 * (p1 + p2), c11 = 1.0 - 2.0 * (q1 * q1 + q3 * q3), c12 = 2.0 * (p5 - p6);
    const T c20 = 2.0 * (p3 - p4), c21 = 2.0 * (p5 + p6), c22 = 1.0 - 2.0 * (q1 * q1 + q2 * q2);
    T F[3] = { T(0.0), T(0.0), T(0.0) };
    if (P.aero_kind == SCVX_AERO_TABLE) {
        const T bv[3] = { c00, c10, c20 };                   // DCM(q) * [1,0,0]   dynamics.jl:58
        aero_force_t(P, tb, bv, x + 4, F);
    }
    const T im = T(1.0) / x[0];
    const T a0 = (c00 * u[0] + c01 * u[1] + c02 * u[2] + F[0]) * im;
    const T a1 = (c10 * u[0] + c11 * u[1] + c12 * u[2] + F[1]) * im;
    const T a2 = (c20 * u[0] + c21 * u[1] + c22 * u[2] + F[2]) * im;
    // rotational dynamics: jBi * (rTB x u - w x (jB w))     dynamics.jl:70
    const T h0 = P.jB[0] * w0 + P.jB[3] * w1 + P.jB[6] * w2;
    const T h1 = P.jB[1] * w0 + P.jB[4] * w1 + P.jB[7] * w2;
    const T h2 = P.jB[2] * w0 + P.jB[5] * w1 + P.jB[8] * w2;
    const T m0 = (P.rTB[1] * u[2] - P.rTB[2] * u[1]) - (w1 * h2 - w2 * h1);
    const T m1 = (P.rTB[2] * u[0] - P.rTB[0] * u[2]) - (w2 * h0 - w0 * h2);
    const T m2 = (P.rTB[0] * u[1] - P.rTB[1] * u[0]) - (w0 * h1 - w1 * h0);
    out[0] = (sqrt_t(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]) * (-P.a)) * sigma;   // dynamics.jl:71
    out[1] = x[4] * sigma; out[2] = x[5] * sigma; out[3] = x[6] * sigma;
    out[4] = (a0 - P.g0) * sigma; out[5] = a1 * sigma; out[6] = a2 * sigma;
    const T hs = sigma * 0.5;                                // 0.5 * Omega(w) * q   dynamics.jl:46-52, 68
    out[7]  = (-(w0 * q1) - w1 * q2 - w2 * q3) * hs;
    out[8]  = (w0 * q0 + w2 * q2 - w1 * q3) * hs;
    out[9]  = (w1 * q0 - w2 * q1 + w0 * q3) * hs;
    out[10] = (w2 * q0 + w1 * q1 - w0 * q2) * hs;
    out[11] = (P.jBi[0] * m0 + P.jBi[3] * m1 + P.jBi[6] * m2) * sigma;
    out[12] = (P.jBi[1] * m0 + P.jBi[4] * m1 + P.jBi[7] * m2) * sigma;
    out[13] = (P.jBi[2] * m0 + P.jBi[5] * m1 + P.jBi[8] * m2) * sigma;
}

// rk4 (dynamics.jl:112-134).  inp = [x ; u- ; u+ ; sigma]; `state` enters as x and leaves as the endpoint.
template <class T>
__device__ __forceinline__ void rk4_t(const scvx_probinfo& P, const ScvxTables& tb, T state[14], const T um[3],
                                      const T up[3], T sigma, double dt, int npts, int mode) {
    const double h = dt / (double)npts;
    const double pcs = 1.0 / (double)npts;
    const double s = (mode == SCVX_MODE_LITERAL) ? 1.0 : h;   // LITERAL: stage increments not scaled (dynamics.jl:126-128)
    const double s2 = 0.5 * s;
    double pca = 0.0;                                          // running sum like the reference (dynamics.jl:120, 130)
    for (int it = 0; it < npts; ++it) {
        T uc[3], k[14], y[14], acc[14];
        const double pm = pca + 0.5 * pcs, pe = pca + pcs;
#pragma unroll
        for (int c = 0; c < 3; ++c) uc[c] = (1.0 - pca) * um[c] + pca * up[c];
        rhs_t(P, tb, state, uc, sigma, k);
#pragma unroll
        for (int r = 0; r < 14; ++r) { acc[r] = k[r]; y[r] = state[r] + k[r] * s2; }
#pragma unroll
        for (int c = 0; c < 3; ++c) uc[c] = (1.0 - pm) * um[c] + pm * up[c];
        rhs_t(P, tb, y, uc, sigma, k);
#pragma unroll
        for (int r = 0; r < 14; ++r) { acc[r] = acc[r] + k[r] * 2.0; y[r] = state[r] + k[r] * s2; }
        rhs_t(P, tb, y, uc, sigma, k);
#pragma unroll
        for (int r = 0; r < 14; ++r) { acc[r] = acc[r] + k[r] * 2.0; y[r] = state[r] + k[r] * s; }
#pragma unroll
        for (int c = 0; c < 3; ++c) uc[c] = (1.0 - pe) * um[c] + pe * up[c];
        rhs_t(P, tb, y, uc, sigma, k);
        pca += pcs;
#pragma unroll
        for (int r = 0; r < 14; ++r) state[r] = state[r] + (acc[r] + k[r]) * (h * (1.0 / 6.0));
    }
}

// ---------------------------------------------------------------------------------------------
// SURVEY.md §8f-4 variant: the fin-force and aero-torque terms the reference carries as comments, restored
// (control_dim 3 -> 5; dynamics.jl:60-63, 66, 69; torque of aerodynamics.jl:45, 49-56).  No live consumer in the
// reference.  Generic in the scalar type like the rest, so the DUALWARP kernel differentiates it exactly.
// ---------------------------------------------------------------------------------------------
template <class T>
__device__ __forceinline__ void aero_force_trq_t(const scvx_probinfo& P, const ScvxTables& tb, const T bv[3],
                                                 const T v[3], T F[3], T Tq[3]) {
    const T nv = sqrt_t(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    const T inv_nv = T(1.0) / nv;
    const T dp = (bv[0] * v[0] + bv[1] * v[1] + bv[2] * v[2]) * inv_nv;
    const T nb = sqrt_t(bv[0] * bv[0] + bv[1] * bv[1] + bv[2] * bv[2]);
    const T cosa = clamp_strict(dp / nb, -1.0, 1.0);
    const T mach = nv * (1.0 / P.sos);
    const T drag = spline_eval(tb.drag, tb, cosa, mach) * P.force_scalar;
    const T dn = drag * inv_nv;
    F[0] = dn * v[0]; F[1] = dn * v[1]; F[2] = dn * v[2];
    Tq[0] = T(0.0); Tq[1] = T(0.0); Tq[2] = T(0.0);
    if (fabs(val(dp)) >= 0.95) return;                       // aerodynamics.jl:42-45: drag only, zero torque
    const T lift = spline_eval(tb.lift, tb, cosa, mach) * P.force_scalar;
    const T trq = spline_eval(tb.trq, tb, cosa, mach) * (P.length_scalar * P.force_scalar);
    T t[3] = { v[1] * bv[2] - v[2] * bv[1], v[2] * bv[0] - v[0] * bv[2], v[0] * bv[1] - v[1] * bv[0] };   // v x bv
    T l[3] = { -(t[1] * v[2] - t[2] * v[1]), -(t[2] * v[0] - t[0] * v[2]), -(t[0] * v[1] - t[1] * v[0]) };
    const T ln = lift / sqrt_t(l[0] * l[0] + l[1] * l[1] + l[2] * l[2]);
    const T tn = trq / sqrt_t(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
    F[0] = F[0] + ln * l[0]; F[1] = F[1] + ln * l[1]; F[2] = F[2] + ln * l[2];
    Tq[0] = tn * t[0]; Tq[1] = tn * t[1]; Tq[2] = tn * t[2];
}

template <class T>
__device__ __forceinline__ void rhs_fins_t(const scvx_probinfo& P, const ScvxTables& tb, const T x[14], const T u[5],
                                           T sigma, T out[14]) {
    const T q0 = x[7], q1 = x[8], q2 = x[9], q3 = x[10];
    const T w0 = x[11], w1 = x[12], w2 = x[13];
    const T p1 = q1 * q2, p2 = q0 * q3, p3 = q1 * q3, p4 = q0 * q2, p5 = q2 * q3, p6 = q0 * q1;
    const T c00 = 1.0 - 2.0 * (q2 * q2 + q3 * q3), c01 = 2.0 * (p1 - p2), c02 = 2.0 * (p3 + p4);
    const T c10 = 2.0 * (p1 + p2), c11 = 1.0 - 2.0 * (q1 * q1 + q3 * q3), c12 = 2.0 * (p5 - p6);
    const T c20 = 2.0 * (p3 - p4), c21 = 2.0 * (p5 + p6), c22 = 1.0 - 2.0 * (q1 * q1 + q2 * q2);
    const T bv[3] = { c00, c10, c20 };
    T F[3], Tq[3];
    aero_force_trq_t(P, tb, bv, x + 4, F, Tq);
    // fin directions (dynamics.jl:60-62): fd1 = normalize((C e2) x v), fd2 = fd1 x v;  ff = u4 fd1 + u5 fd2 (:63)
    const T v0 = x[4], v1 = x[5], v2 = x[6];
    T d1[3] = { c11 * v2 - c21 * v1, c21 * v0 - c01 * v2, c01 * v1 - c11 * v0 };
    const T in1 = T(1.0) / sqrt_t(d1[0] * d1[0] + d1[1] * d1[1] + d1[2] * d1[2]);
    d1[0] = d1[0] * in1; d1[1] = d1[1] * in1; d1[2] = d1[2] * in1;
    const T d2[3] = { d1[1] * v2 - d1[2] * v1, d1[2] * v0 - d1[0] * v2, d1[0] * v1 - d1[1] * v0 };
    const T ff[3] = { u[3] * d1[0] + u[4] * d2[0], u[3] * d1[1] + u[4] * d2[1], u[3] * d1[2] + u[4] * d2[2] };
    const T im = T(1.0) / x[0];
    const T a0 = (c00 * u[0] + c01 * u[1] + c02 * u[2] + (F[0] + ff[0])) * im;
    const T a1 = (c10 * u[0] + c11 * u[1] + c12 * u[2] + (F[1] + ff[1])) * im;
    const T a2 = (c20 * u[0] + c21 * u[1] + c22 * u[2] + (F[2] + ff[2])) * im;
    const T h0 = P.jB[0] * w0 + P.jB[3] * w1 + P.jB[6] * w2;
    const T h1 = P.jB[1] * w0 + P.jB[4] * w1 + P.jB[7] * w2;
    const T h2 = P.jB[2] * w0 + P.jB[5] * w1 + P.jB[8] * w2;
    // rTB x u + (rFB x ff + bdy_trq) - w x (jB w)     dynamics.jl:69-70
    const T m0 = (P.rTB[1] * u[2] - P.rTB[2] * u[1]) + ((P.rFB[1] * ff[2] - P.rFB[2] * ff[1]) + Tq[0]) - (w1 * h2 - w2 * h1);
    const T m1 = (P.rTB[2] * u[0] - P.rTB[0] * u[2]) + ((P.rFB[2] * ff[0] - P.rFB[0] * ff[2]) + Tq[1]) - (w2 * h0 - w0 * h2);
    const T m2 = (P.rTB[0] * u[1] - P.rTB[1] * u[0]) + ((P.rFB[0] * ff[1] - P.rFB[1] * ff[0]) + Tq[2]) - (w0 * h1 - w1 * h0);
    out[0] = (sqrt_t(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]) * (-P.a)) * sigma;
    out[1] = x[4] * sigma; out[2] = x[5] * sigma; out[3] = x[6] * sigma;
    out[4] = (a0 - P.g0) * sigma; out[5] = a1 * sigma; out[6] = a2 * sigma;
    const T hs = sigma * 0.5;
    out[7]  = (-(w0 * q1) - w1 * q2 - w2 * q3) * hs;
    out[8]  = (w0 * q0 + w2 * q2 - w1 * q3) * hs;
    out[9]  = (w1 * q0 - w2 * q1 + w0 * q3) * hs;
    out[10] = (w2 * q0 + w1 * q1 - w0 * q2) * hs;
    out[11] = (P.jBi[0] * m0 + P.jBi[3] * m1 + P.jBi[6] * m2) * sigma;
    out[12] = (P.jBi[1] * m0 + P.jBi[4] * m1 + P.jBi[7] * m2) * sigma;
    out[13] = (P.jBi[2] * m0 + P.jBi[5] * m1 + P.jBi[8] * m2) * sigma;
}

template <class T>
__device__ __forceinline__ void rk4_fins_t(const scvx_probinfo& P, const ScvxTables& tb, T state[14], const T um[5],
                                           const T up[5], T sigma, double dt, int npts, int mode) {
    const double h = dt / (double)npts;
    const double pcs = 1.0 / (double)npts;
    const double s = (mode == SCVX_MODE_LITERAL) ? 1.0 : h;
    const double s2 = 0.5 * s;
    double pca = 0.0;
    for (int it = 0; it < npts; ++it) {
        T uc[5], k[14], y[14], acc[14];
        const double pm = pca + 0.5 * pcs, pe = pca + pcs;
#pragma unroll
        for (int c = 0; c < 5; ++c) uc[c] = (1.0 - pca) * um[c] + pca * up[c];
        rhs_fins_t(P, tb, state, uc, sigma, k);
#pragma unroll
        for (int r = 0; r < 14; ++r) { acc[r] = k[r]; y[r] = state[r] + k[r] * s2; }
#pragma unroll
        for (int c = 0; c < 5; ++c) uc[c] = (1.0 - pm) * um[c] + pm * up[c];
        rhs_fins_t(P, tb, y, uc, sigma, k);
#pragma unroll
        for (int r = 0; r < 14; ++r) { acc[r] = acc[r] + k[r] * 2.0; y[r] = state[r] + k[r] * s2; }
        rhs_fins_t(P, tb, y, uc, sigma, k);
#pragma unroll
        for (int r = 0; r < 14; ++r) { acc[r] = acc[r] + k[r] * 2.0; y[r] = state[r] + k[r] * s; }
#pragma unroll
        for (int c = 0; c < 5; ++c) uc[c] = (1.0 - pe) * um[c] + pe * up[c];
        rhs_fins_t(P, tb, y, uc, sigma, k);
        pca += pcs;
#pragma unroll
        for (int r = 0; r < 14; ++r) state[r] = state[r] + (acc[r] + k[r]) * (h * (1.0 / 6.0));
    }
}
