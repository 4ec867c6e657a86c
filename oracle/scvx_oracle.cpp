// scvx_oracle.cpp — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
//
// A CPU restatement of the per-interval linearise-and-discretise path of
// BenChung/SuccessiveConvexification, variant "V2" of SURVEY.md §8a:
//   fixed-step rk4 (reference dynamics.jl:112-134) differentiated exactly by
//   forward-mode dual numbers (what Zygote.forward_jacobian / ForwardDiff do at
//   dynamics.jl:311-313), over the 6-DoF right-hand side dx_static
//   (dynamics.jl:54-77) with DCM (29-44), Omega (46-52), current_control
//   (108-110) and the table aerodynamics aero_force (aerodynamics.jl:38-58).
//
// Third-party algorithms that are NOT in /root/reference and are restated here
// from their published definitions (versions are unpinned: the reference ships
// no Project.toml / Manifest.toml):
//   * Interpolations.jl  `extrapolate(scale(interpolate(A, BSpline(Cubic(Line(OnGrid())))), r1, r2), Flat())`
//     (call sites aerodynamics.jl:19-21, 43, 47-49): separable cubic B-spline
//     prefilter with zero second derivative at the first/last grid point, one
//     padding coefficient per side, Flat (clamp) extrapolation.
//   * ForwardDiff.jl / Zygote.jl forward mode (call site dynamics.jl:312): the exact
//     derivative of the executed arithmetic (taken branches, active clamps).
//
// PARITY UNPINNED: the reference has no tests, golden vectors or fixtures for
// this path and Julia is not installed, so this oracle cannot
// be checked against reference-produced numbers.  It is pinned instead by
// independent means (tests/test_oracle.py): a numpy complex-step restatement
// (oracle/py_restatement.py), scipy's natural cubic spline, finite differences
// and the survey's scratch anchors.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load this library.  The product (.so under
// successiveconvexification_b200/csrc) never links, loads or calls it.
//
// Build: oracle/build.py  (g++ -O2 -ffp-contract=off -fopenmp -shared -fPIC ... -lquadmath)

#include <cmath>
#include <quadmath.h>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ----------------------------------------------------------------------------------------------
// Real types.  The path is evaluated in IEEE double (the reference's Float64) and, for the
// conditioning-aware parity checks, in IEEE binary128 (__float128, libquadmath): same operation
// sequence, same double-valued constants, parameters, inputs and spline coefficients, 113-bit
// arithmetic — i.e. the exact value (to ~1e-33) of the function the FP64 code approximates.  The
// distance of the FP64 result from it is a direct measurement of kappa * eps for that interval.
// ----------------------------------------------------------------------------------------------
typedef __float128 quad;
inline double real_sqrt(double a) { return std::sqrt(a); }
inline quad real_sqrt(quad a) { return sqrtq(a); }
inline double real_floor(double a) { return std::floor(a); }
inline quad real_floor(quad a) { return floorq(a); }
inline double real_abs(double a) { return std::fabs(a); }
inline quad real_abs(quad a) { return fabsq(a); }

// ----------------------------------------------------------------------------------------------
// Dual numbers with N partials (ForwardDiff.Dual{T,Float64,N} semantics) over the real type R.
// ----------------------------------------------------------------------------------------------
template <class R, int N>
struct Dual {
    R v;
    R d[N];
    Dual() : v(0.0) { for (int i = 0; i < N; ++i) d[i] = 0.0; }
    Dual(R x) : v(x) { for (int i = 0; i < N; ++i) d[i] = 0.0; }
};

#define DT template <class R, int N> inline Dual<R, N>
DT operator+(const Dual<R, N>& a, const Dual<R, N>& b) {
    Dual<R, N> r; r.v = a.v + b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
DT operator-(const Dual<R, N>& a, const Dual<R, N>& b) {
    Dual<R, N> r; r.v = a.v - b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
DT operator-(const Dual<R, N>& a) {
    Dual<R, N> r; r.v = -a.v; for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r; }
DT operator*(const Dual<R, N>& a, const Dual<R, N>& b) {
    Dual<R, N> r; r.v = a.v * b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
DT operator/(const Dual<R, N>& a, const Dual<R, N>& b) {
    Dual<R, N> r; r.v = a.v / b.v; const R ib = 1.0 / b.v;
    for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * ib;
    return r; }
DT operator+(const Dual<R, N>& a, double b) { Dual<R, N> r = a; r.v += b; return r; }
DT operator+(double a, const Dual<R, N>& b) { return b + a; }
DT operator-(const Dual<R, N>& a, double b) { Dual<R, N> r = a; r.v -= b; return r; }
DT operator-(double a, const Dual<R, N>& b) { return (-b) + a; }
DT operator*(const Dual<R, N>& a, double b) {
    Dual<R, N> r; r.v = a.v * b; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b; return r; }
DT operator*(double a, const Dual<R, N>& b) { return b * a; }
DT operator/(const Dual<R, N>& a, double b) {
    Dual<R, N> r; r.v = a.v / b; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] / b; return r; }
DT operator/(double a, const Dual<R, N>& b) { return Dual<R, N>(a) / b; }
DT sqrt(const Dual<R, N>& a) {
    Dual<R, N> r; r.v = real_sqrt(a.v); const R s = 0.5 / r.v;
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * s;
    return r; }
#undef DT
inline double sqrt(double a) { return std::sqrt(a); }

inline double value_of(double x) { return x; }
template <class R, int N> inline R value_of(const Dual<R, N>& x) { return x.v; }

// Base.clamp(x, lo, hi): returns x itself (partials kept) unless STRICTLY outside,
// in which case the bound is returned as a constant (zero partials).
template <class T> inline T clamp_like_julia(const T& x, double lo, double hi) {
    if (value_of(x) > hi) return T(hi);
    if (value_of(x) < lo) return T(lo);
    return x;
}

// Branch signature of one evaluation: every data-dependent decision of the path (the |dp| >= 0.95 branch, active
// clamps, spline cell indices) is folded into a 64-bit hash.  Two evaluations of the same interval (FP64 / binary128)
// that took different decisions differentiate DIFFERENT smooth pieces of the map and cannot be compared entry by entry.
thread_local uint64_t g_sig = 0;
inline void sig_mix(int token) { g_sig = (g_sig ^ (uint64_t)(uint32_t)token) * 1099511628211ULL; }

// ----------------------------------------------------------------------------------------------
// Problem parameters: mirror of ProbInfo (master.jl:73-83) + AtmosphericData scalars
// (master.jl:10-16) + Tmin (master.jl:21).  Same memory layout as scvx_probinfo in
// include/scvx_b200.h (kept in sync by tests/test_abi.py).
// ----------------------------------------------------------------------------------------------
struct ProbInfo {
    double a, g0, sos;
    double jB[9], jBi[9];   // column-major 3x3
    double rTB[3], rFB[3];
    double force_scalar, length_scalar;
    double Tmin;
    int32_t aero_kind;      // 0 = ExoatmosphericData (zero aero force), 1 = AtmosphericData tables
    int32_t pad_;
};

// Spline coefficient table: (n1+2) x (n2+2) column-major, index coordinate on axis a is
// (x - x0)/dx + 1 (Interpolations.jl `scale` over a StepRangeLen).
struct Table {
    const double* coef;
    int n1, n2;
    double x0, dx, y0, dy;
};

struct Tables {
    Table drag, lift, trq;
    int have, have_trq;
};

// ----------------------------------------------------------------------------------------------
// Interpolations.jl restatement.
// ----------------------------------------------------------------------------------------------

// Solve the (n+2)x(n+2) prefilter system for one line:
//   row 0      :  c[0] - 2 c[1] + c[2]            = 0      (Line BC, OnGrid: zero 2nd derivative at grid point 1)
//   row k=1..n :  c[k-1]/6 + 2 c[k]/3 + c[k+1]/6  = data[k-1]
//   row n+1    :  c[n-1] - 2 c[n] + c[n+1]        = 0
// Dense Gaussian elimination with partial pivoting, deliberately NOT the closed-form
// shortcut the product uses (independence of the two implementations).
void solve_prefilter_line(const double* data, int n, int stride_in, double* c, int stride_out) {
    const int m = n + 2;
    std::vector<double> A((size_t)m * m, 0.0), b(m, 0.0);
    A[0 * m + 0] = 1.0; A[0 * m + 1] = -2.0; A[0 * m + 2] = 1.0;
    for (int k = 1; k <= n; ++k) {
        A[k * m + (k - 1)] = 1.0 / 6.0; A[k * m + k] = 2.0 / 3.0; A[k * m + (k + 1)] = 1.0 / 6.0;
        b[k] = data[(size_t)(k - 1) * stride_in];
    }
    A[(m - 1) * m + (m - 3)] = 1.0; A[(m - 1) * m + (m - 2)] = -2.0; A[(m - 1) * m + (m - 1)] = 1.0;
    for (int col = 0; col < m; ++col) {
        int piv = col; double best = std::fabs(A[col * m + col]);
        const int rmax = std::min(m, col + 3);   // band structure: only nearby rows can be non-zero
        for (int r = col + 1; r < rmax; ++r) if (std::fabs(A[r * m + col]) > best) { best = std::fabs(A[r * m + col]); piv = r; }
        if (piv != col) { for (int j = 0; j < m; ++j) std::swap(A[col * m + j], A[piv * m + j]); std::swap(b[col], b[piv]); }
        const double inv = 1.0 / A[col * m + col];
        for (int r = col + 1; r < rmax; ++r) {
            const double f = A[r * m + col] * inv;
            if (f == 0.0) continue;
            for (int j = col; j < std::min(m, col + 5); ++j) A[r * m + j] -= f * A[col * m + j];
            b[r] -= f * b[col];
        }
    }
    for (int r = m - 1; r >= 0; --r) {
        double s = b[r];
        for (int j = r + 1; j < std::min(m, r + 5); ++j) s -= A[r * m + j] * c[(size_t)j * stride_out];
        c[(size_t)r * stride_out] = s / A[r * m + r];
    }
}

// Value (and exact derivative through T) of the Flat-extrapolated, scaled cubic B-spline.
template <class T>
T spline_eval(const Table& t, const T& x, const T& y) {
    const int L1 = t.n1 + 2;
    // scale(): index coordinate; extrapolate(Flat()): clamp to [1, n] (constant when strictly outside)
    const T xr = (x - t.x0) / t.dx + 1.0, yr = (y - t.y0) / t.dy + 1.0;
    T xi = clamp_like_julia(xr, 1.0, (double)t.n1);
    T yi = clamp_like_julia(yr, 1.0, (double)t.n2);
    int i = (int)real_floor(value_of(xi)); if (i > t.n1 - 1) i = t.n1 - 1; if (i < 1) i = 1;
    int j = (int)real_floor(value_of(yi)); if (j > t.n2 - 1) j = t.n2 - 1; if (j < 1) j = 1;
    sig_mix(i); sig_mix(j);
    sig_mix((value_of(xr) > (double)t.n1 ? 1 : value_of(xr) < 1.0 ? 2 : 0) | (value_of(yr) > (double)t.n2 ? 4 : value_of(yr) < 1.0 ? 8 : 0));
    const T dx = xi - (double)i, dy = yi - (double)j;
    const T ox = 1.0 - dx, oy = 1.0 - dy;
    // value_weights(::Cubic, δ)
    T wx[4] = { (ox * ox * ox) * (1.0 / 6.0),
                2.0 / 3.0 - dx * dx + (dx * dx * dx) * 0.5,
                2.0 / 3.0 - ox * ox + (ox * ox * ox) * 0.5,
                (dx * dx * dx) * (1.0 / 6.0) };
    T wy[4] = { (oy * oy * oy) * (1.0 / 6.0),
                2.0 / 3.0 - dy * dy + (dy * dy * dy) * 0.5,
                2.0 / 3.0 - oy * oy + (oy * oy * oy) * 0.5,
                (dy * dy * dy) * (1.0 / 6.0) };
    // coefficient with grid index g (1-based, padded) lives at storage offset g (0-based storage holds index 0..n+1)
    T acc(0.0);
    for (int b = 0; b < 4; ++b) {
        T row(0.0);
        for (int a = 0; a < 4; ++a) row = row + wx[a] * t.coef[(size_t)(i - 1 + a) + (size_t)(j - 1 + b) * L1];
        acc = acc + wy[b] * row;
    }
    return acc;
}

// ----------------------------------------------------------------------------------------------
// dynamics.jl restatement (generic in the scalar type, like the Julia source).
// ----------------------------------------------------------------------------------------------

// DCM(quat)  dynamics.jl:29-44 — row-major C[r][c]; no normalisation of q.
template <class T>
void DCM(const T q[4], T C[3][3]) {
    const T q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3];
    const T p1 = q1 * q2, p2 = q0 * q3, p3 = q1 * q3, p4 = q0 * q2, p5 = q2 * q3, p6 = q0 * q1;
    C[0][0] = 1.0 - 2.0 * (q2 * q2 + q3 * q3); C[0][1] = 2.0 * (p1 - p2);               C[0][2] = 2.0 * (p3 + p4);
    C[1][0] = 2.0 * (p1 + p2);               C[1][1] = 1.0 - 2.0 * (q1 * q1 + q3 * q3); C[1][2] = 2.0 * (p5 - p6);
    C[2][0] = 2.0 * (p3 - p4);               C[2][1] = 2.0 * (p5 + p6);               C[2][2] = 1.0 - 2.0 * (q1 * q1 + q2 * q2);
}

template <class T> inline void cross3(const T a[3], const T b[3], T o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
template <class T> inline T norm3(const T a[3]) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

// Aerodynamics.aero_force numeric method, aerodynamics.jl:38-58 (force only; the torque it
// also returns is discarded by the caller, dynamics.jl:69).
template <class T>
void aero_force(const ProbInfo& P, const Tables& tb, const T bv[3], const T vel[3], T F[3]) {
    const T nv = norm3(vel);
    const T dp = (bv[0] * vel[0] + bv[1] * vel[1] + bv[2] * vel[2]) / nv;          // :39
    const T car = dp / norm3(bv);
    const T cos_aoa = clamp_like_julia(car, -1.0, 1.0);                              // :40
    sig_mix((real_abs(value_of(dp)) >= 0.95 ? 16 : 0) | (value_of(car) > 1.0 ? 32 : value_of(car) < -1.0 ? 64 : 0));
    const T mach = nv / P.sos;                                                        // :41
    const T drag = spline_eval(tb.drag, cos_aoa, mach) * P.force_scalar;             // :43 / :47
    if (real_abs(value_of(dp)) >= 0.95) {                                            // :42
        for (int k = 0; k < 3; ++k) F[k] = drag * vel[k] / nv;                        // :44
        return;
    }
    const T lift = spline_eval(tb.lift, cos_aoa, mach) * P.force_scalar;             // :48
    T trqd[3]; cross3(vel, bv, trqd);                                                 // :50
    T ntrqd[3] = { -trqd[0], -trqd[1], -trqd[2] };
    T liftd[3]; cross3(ntrqd, vel, liftd);                                            // :51
    const T nl = norm3(liftd);
    for (int k = 0; k < 3; ++k) liftd[k] = liftd[k] / nl;                             // :52
    for (int k = 0; k < 3; ++k) F[k] = drag * vel[k] / nv + lift * liftd[k];          // :54-56
}

// dx_static numeric method, dynamics.jl:54-77.
template <class T>
void dx_static(const ProbInfo& P, const Tables& tb, const T x[14], const T u[3], const T& mult, T out[14]) {
    const T* q = x + 7;  const T* w = x + 11;  const T* v = x + 4;
    T C[3][3]; DCM(q, C);
    T aerf[3] = { T(0.0), T(0.0), T(0.0) };
    if (P.aero_kind == 1) {
        T bv[3] = { C[0][0] * 1.0 + C[0][1] * 0.0 + C[0][2] * 0.0,                    // DCM(qbi) * [1,0,0]  :58
                    C[1][0] * 1.0 + C[1][1] * 0.0 + C[1][2] * 0.0,
                    C[2][0] * 1.0 + C[2][1] * 0.0 + C[2][2] * 0.0 };
        aero_force(P, tb, bv, v, aerf);
    }   // ExoatmosphericData: zero force (SURVEY.md §8a a5 "Exo caveat", Appendix B6)
    T thr[3], acc[3];
    for (int r = 0; r < 3; ++r) thr[r] = C[r][0] * u[0] + C[r][1] * u[1] + C[r][2] * u[2];   // :65
    for (int r = 0; r < 3; ++r) acc[r] = (thr[r] + aerf[r]) / x[0];                              // :67
    // rot_vel = 0.5 * Omega(omb) * qbi   :46-52, :68
    T rv[4];
    rv[0] = 0.5 * (-(w[0] * q[1]) - w[1] * q[2] - w[2] * q[3]);
    rv[1] = 0.5 * (w[0] * q[0] + w[2] * q[2] - w[1] * q[3]);
    rv[2] = 0.5 * (w[1] * q[0] - w[2] * q[1] + w[0] * q[3]);
    rv[3] = 0.5 * (w[2] * q[0] + w[1] * q[1] - w[0] * q[2]);
    // rot_acc = jBi * (cross(rTB,u) + 0 - cross(omb, jB*omb))   :70
    T rTB[3] = { T(P.rTB[0]), T(P.rTB[1]), T(P.rTB[2]) };
    T t1[3]; cross3(rTB, u, t1);
    T Jw[3];
    for (int r = 0; r < 3; ++r) Jw[r] = P.jB[r + 0] * w[0] + P.jB[r + 3] * w[1] + P.jB[r + 6] * w[2];
    T t2[3]; cross3(w, Jw, t2);
    T rhs[3] = { t1[0] + 0.0 - t2[0], t1[1] + 0.0 - t2[1], t1[2] + 0.0 - t2[2] };
    T ra[3];
    for (int r = 0; r < 3; ++r) ra[r] = P.jBi[r + 0] * rhs[0] + P.jBi[r + 3] * rhs[1] + P.jBi[r + 6] * rhs[2];
    out[0] = (-P.a * sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2])) * mult;            // :71
    out[1] = x[4] * mult; out[2] = x[5] * mult; out[3] = x[6] * mult;                  // :72
    out[4] = (acc[0] - P.g0) * mult; out[5] = acc[1] * mult; out[6] = acc[2] * mult;   // :73
    for (int k = 0; k < 4; ++k) out[7 + k] = rv[k] * mult;                             // :74
    for (int k = 0; k < 3; ++k) out[11 + k] = ra[k] * mult;                            // :74
}

// current_control, dynamics.jl:108-110
template <class T>
inline void current_control(double pc, const T a[3], const T b[3], T o[3]) {
    for (int k = 0; k < 3; ++k) o[k] = (1.0 - pc) * a[k] + pc * b[k];
}

// rk4, dynamics.jl:112-134.  mode 0 = LITERAL (stage increments not scaled by the step, :126-128),
// mode 1 = TEXTBOOK (stage increments scaled by idt).
template <class T>
void rk4(const ProbInfo& P, const Tables& tb, const T inp[21], double dt, int npts, int mode, T state[14]) {
    for (int k = 0; k < 14; ++k) state[k] = inp[k];
    const T* su = inp + 14; const T* eu = inp + 17;
    const double idt = dt / npts;
    const double pcs = 1.0 / npts;
    double pca = 0.0;
    const double s = (mode == 0) ? 1.0 : idt;
    for (int i = 0; i < npts; ++i) {
        T ict[3], mct[3], ect[3];
        current_control(pca, su, eu, ict);
        current_control(pca + pcs / 2, su, eu, mct);
        current_control(pca + pcs, su, eu, ect);
        T k1[14], k2[14], k3[14], k4[14], tmp[14];
        dx_static(P, tb, state, ict, inp[20], k1);
        for (int k = 0; k < 14; ++k) tmp[k] = state[k] + (mode == 0 ? k1[k] / 2.0 : (k1[k] * s) / 2.0);
        dx_static(P, tb, tmp, mct, inp[20], k2);
        for (int k = 0; k < 14; ++k) tmp[k] = state[k] + (mode == 0 ? k2[k] / 2.0 : (k2[k] * s) / 2.0);
        dx_static(P, tb, tmp, mct, inp[20], k3);
        for (int k = 0; k < 14; ++k) tmp[k] = state[k] + (mode == 0 ? k3[k] : k3[k] * s);
        dx_static(P, tb, tmp, ect, inp[20], k4);
        pca += pcs;
        for (int k = 0; k < 14; ++k)
            state[k] = state[k] + idt * (k1[k] / 6.0 + k2[k] / 3.0 + k3[k] / 3.0 + k4[k] / 6.0);   // :131
    }
}

// ----------------------------------------------------------------------------------------------
// the fin-force and aero-torque terms the reference carries as COMMENTS, restored (control_dim 3 -> 5).
// There is no live consumer in the reference (`control_dim = 3`, rocketland.jl:17; `aero_trq = [0,0,0]`,
// dynamics.jl:69), so nothing here can be checked against reference behaviour: the formulas are the commented-out
// expressions themselves,
//     ff       = u[4]*fd1 + u[5]*fd2                       dynamics.jl:60-63   (fd1 = normalize(C e2 x v), fd2 = fd1 x v)
//     aero_frc = aerf + ff                                 dynamics.jl:66
//     aero_trq = cross(rFB, ff) + bdy_trq                  dynamics.jl:69
// with bdy_trq the torque `aero_force` already returns (aerodynamics.jl:45, 49-56): zero in the |dp| >= 0.95 branch,
// trq_itrp(cos_aoa, mach) * length_scalar * force_scalar * normalize(v x bv) otherwise.
// ----------------------------------------------------------------------------------------------
template <class T>
void aero_force_trq(const ProbInfo& P, const Tables& tb, const T bv[3], const T vel[3], T F[3], T Tq[3]) {
    const T nv = norm3(vel);
    const T dp = (bv[0] * vel[0] + bv[1] * vel[1] + bv[2] * vel[2]) / nv;
    const T car = dp / norm3(bv);
    const T cos_aoa = clamp_like_julia(car, -1.0, 1.0);
    const T mach = nv / P.sos;
    const T drag = spline_eval(tb.drag, cos_aoa, mach) * P.force_scalar;
    if (real_abs(value_of(dp)) >= 0.95) {
        for (int k = 0; k < 3; ++k) { F[k] = drag * vel[k] / nv; Tq[k] = T(0.0); }                // :44-45
        return;
    }
    const T lift = spline_eval(tb.lift, cos_aoa, mach) * P.force_scalar;
    const T trq = spline_eval(tb.trq, cos_aoa, mach) * P.length_scalar * P.force_scalar;          // :49
    T trqd[3]; cross3(vel, bv, trqd);
    T ntrqd[3] = { -trqd[0], -trqd[1], -trqd[2] };
    T liftd[3]; cross3(ntrqd, vel, liftd);
    const T nl = norm3(liftd), nt = norm3(trqd);
    for (int k = 0; k < 3; ++k) {
        F[k] = drag * vel[k] / nv + lift * (liftd[k] / nl);                                       // :52-56
        Tq[k] = (trqd[k] / nt) * trq;                                                             // :53, :56
    }
}

template <class T>
void dx_static_fins(const ProbInfo& P, const Tables& tb, const T x[14], const T u[5], const T& mult, T out[14]) {
    const T* q = x + 7;  const T* w = x + 11;  const T* v = x + 4;
    T C[3][3]; DCM(q, C);
    T bv[3] = { C[0][0], C[1][0], C[2][0] };
    T aerf[3], btrq[3];
    aero_force_trq(P, tb, bv, v, aerf, btrq);
    T by[3] = { C[0][1], C[1][1], C[2][1] };                                                      // DCM(qbi) * [0,1,0]   :60
    T fd1[3]; cross3(by, v, fd1);
    const T n1 = sqrt(fd1[0] * fd1[0] + fd1[1] * fd1[1] + fd1[2] * fd1[2]);
    for (int k = 0; k < 3; ++k) fd1[k] = fd1[k] / n1;                                             // :61
    T fd2[3]; cross3(fd1, v, fd2);                                                                // :62
    T ff[3];
    for (int k = 0; k < 3; ++k) ff[k] = u[3] * fd1[k] + u[4] * fd2[k];                            // :63
    T thr[3], acc[3];
    for (int r = 0; r < 3; ++r) thr[r] = C[r][0] * u[0] + C[r][1] * u[1] + C[r][2] * u[2];
    for (int r = 0; r < 3; ++r) acc[r] = (thr[r] + (aerf[r] + ff[r])) / x[0];                     // :66-67
    T rv[4];
    rv[0] = 0.5 * (-(w[0] * q[1]) - w[1] * q[2] - w[2] * q[3]);
    rv[1] = 0.5 * (w[0] * q[0] + w[2] * q[2] - w[1] * q[3]);
    rv[2] = 0.5 * (w[1] * q[0] - w[2] * q[1] + w[0] * q[3]);
    rv[3] = 0.5 * (w[2] * q[0] + w[1] * q[1] - w[0] * q[2]);
    T rTB[3] = { T(P.rTB[0]), T(P.rTB[1]), T(P.rTB[2]) };
    T rFB[3] = { T(P.rFB[0]), T(P.rFB[1]), T(P.rFB[2]) };
    T t1[3]; cross3(rTB, u, t1);
    T tf[3]; cross3(rFB, ff, tf);
    T Jw[3];
    for (int r = 0; r < 3; ++r) Jw[r] = P.jB[r + 0] * w[0] + P.jB[r + 3] * w[1] + P.jB[r + 6] * w[2];
    T t2[3]; cross3(w, Jw, t2);
    T rhs[3];
    for (int k = 0; k < 3; ++k) rhs[k] = t1[k] + (tf[k] + btrq[k]) - t2[k];                       // :69-70
    T ra[3];
    for (int r = 0; r < 3; ++r) ra[r] = P.jBi[r + 0] * rhs[0] + P.jBi[r + 3] * rhs[1] + P.jBi[r + 6] * rhs[2];
    out[0] = (-P.a * sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2])) * mult;
    out[1] = x[4] * mult; out[2] = x[5] * mult; out[3] = x[6] * mult;
    out[4] = (acc[0] - P.g0) * mult; out[5] = acc[1] * mult; out[6] = acc[2] * mult;
    for (int k = 0; k < 4; ++k) out[7 + k] = rv[k] * mult;
    for (int k = 0; k < 3; ++k) out[11 + k] = ra[k] * mult;
}

// rk4 (dynamics.jl:112-134) over inp = [x(14); u_k(5); u_{k+1}(5); sigma] (25)
template <class T>
void rk4_fins(const ProbInfo& P, const Tables& tb, const T inp[25], double dt, int npts, int mode, T state[14]) {
    for (int k = 0; k < 14; ++k) state[k] = inp[k];
    const T* su = inp + 14; const T* eu = inp + 19;
    const double idt = dt / npts, pcs = 1.0 / npts, s = (mode == 0) ? 1.0 : idt;
    double pca = 0.0;
    for (int i = 0; i < npts; ++i) {
        T ict[5], mct[5], ect[5];
        for (int k = 0; k < 5; ++k) {
            ict[k] = (1.0 - pca) * su[k] + pca * eu[k];
            mct[k] = (1.0 - (pca + pcs / 2)) * su[k] + (pca + pcs / 2) * eu[k];
            ect[k] = (1.0 - (pca + pcs)) * su[k] + (pca + pcs) * eu[k];
        }
        T k1[14], k2[14], k3[14], k4[14], tmp[14];
        dx_static_fins(P, tb, state, ict, inp[24], k1);
        for (int k = 0; k < 14; ++k) tmp[k] = state[k] + (mode == 0 ? k1[k] / 2.0 : (k1[k] * s) / 2.0);
        dx_static_fins(P, tb, tmp, mct, inp[24], k2);
        for (int k = 0; k < 14; ++k) tmp[k] = state[k] + (mode == 0 ? k2[k] / 2.0 : (k2[k] * s) / 2.0);
        dx_static_fins(P, tb, tmp, mct, inp[24], k3);
        for (int k = 0; k < 14; ++k) tmp[k] = state[k] + (mode == 0 ? k3[k] : k3[k] * s);
        dx_static_fins(P, tb, tmp, ect, inp[24], k4);
        pca += pcs;
        for (int k = 0; k < 14; ++k) state[k] = state[k] + idt * (k1[k] / 6.0 + k2[k] / 3.0 + k3[k] / 3.0 + k4[k] / 6.0);
    }
}

// block: 14 x 27 column-major, col 0 endpoint, cols 1..25 D = d endpoint / d inp (25), col 26 z  (acc_width =
// state_dim + 2*control_dim + 3 with control_dim = 5, rocketland.jl:22)
void linearize_interval_fins(const ProbInfo& P, const Tables& tb, const double inp[25], double dt, int npts, int mode,
                             double* block) {
    typedef Dual<double, 25> D25;
    D25 din[25], out[14];
    for (int k = 0; k < 25; ++k) { din[k] = D25(inp[k]); din[k].d[k] = 1.0; }
    rk4_fins<D25>(P, tb, din, dt, npts, mode, out);
    for (int r = 0; r < 14; ++r) {
        block[r] = out[r].v;
        double zr = out[r].v;
        for (int c = 0; c < 25; ++c) { block[r + 14 * (1 + c)] = out[r].d[c]; zr -= out[r].d[c] * inp[c]; }
        block[r + 14 * 26] = zr;
    }
}

// sensitivity_zygote (dynamics.jl:311-313) for one interval + named outputs of old_dynamics.jl:84-98.
// block: 14 x 23 column-major, col 0 = endpoint, cols 1..21 = D = d endpoint / d inp, col 22 = z.
// R = double: the reference's arithmetic.  R = quad: the same operation sequence in binary128, results rounded to
// double on output (z is formed in R before rounding).  Returns the branch signature of the evaluation.
template <class R>
uint64_t linearize_interval_t(const ProbInfo& P, const Tables& tb, const double inp[21], double dt, int npts, int mode,
                              double* block) {
    typedef Dual<R, 21> D21;
    D21 din[21], out[14];
    for (int k = 0; k < 21; ++k) { din[k] = D21((R)inp[k]); din[k].d[k] = 1.0; }
    g_sig = 1469598103934665603ULL;
    rk4<D21>(P, tb, din, dt, npts, mode, out);
    for (int r = 0; r < 14; ++r) {
        block[r] = (double)out[r].v;
        R zr = out[r].v;
        for (int c = 0; c < 21; ++c) { block[r + 14 * (1 + c)] = (double)out[r].d[c]; zr -= out[r].d[c] * inp[c]; }
        block[r + 14 * 22] = (double)zr;
    }
    return g_sig;
}
inline void linearize_interval(const ProbInfo& P, const Tables& tb, const double inp[21], double dt, int npts, int mode,
                               double* block) {
    linearize_interval_t<double>(P, tb, inp, dt, npts, mode, block);
}

Tables make_tables(const double* drag_coef, const double* lift_coef, const double* geom) {
    Tables tb; std::memset(&tb, 0, sizeof(tb));
    if (drag_coef && lift_coef && geom) {
        const int n1 = (int)geom[0], n2 = (int)geom[1];
        tb.drag = Table{ drag_coef, n1, n2, geom[2], geom[3], geom[4], geom[5] };
        tb.lift = Table{ lift_coef, n1, n2, geom[2], geom[3], geom[4], geom[5] };
        tb.have = 1;
    }
    return tb;
}

Tables make_tables3(const double* drag_coef, const double* lift_coef, const double* trq_coef, const double* geom) {
    Tables tb = make_tables(drag_coef, lift_coef, geom);
    if (tb.have && trq_coef) {
        tb.trq = Table{ trq_coef, tb.drag.n1, tb.drag.n2, geom[2], geom[3], geom[4], geom[5] };
        tb.have_trq = 1;
    }
    return tb;
}

}  // namespace

// ----------------------------------------------------------------------------------------------
// C entry points (ctypes).  `geom` = {n_cos, n_mach, cos0, dcos, mach0, dmach}.
// ----------------------------------------------------------------------------------------------
extern "C" {

int oracle_sizeof_probinfo(void) { return (int)sizeof(ProbInfo); }

// samples: n1 x n2 column-major (as reshape() at aerodynamics.jl:19-21); coef: (n1+2) x (n2+2) column-major.
int oracle_prefilter(const double* samples, int n1, int n2, double* coef) {
    const int L1 = n1 + 2;
    std::vector<double> tmp((size_t)L1 * n2);
    for (int j = 0; j < n2; ++j) solve_prefilter_line(samples + (size_t)j * n1, n1, 1, tmp.data() + (size_t)j * L1, 1);
    for (int i = 0; i < L1; ++i) solve_prefilter_line(tmp.data() + i, n2, L1, coef + i, L1);
    return 0;
}

// value and gradient (d/dx, d/dy) of the scaled, Flat-extrapolated spline at (x, y)
double oracle_spline_eval(const double* coef, const double* geom, double x, double y, double* grad) {
    Table t{ coef, (int)geom[0], (int)geom[1], geom[2], geom[3], geom[4], geom[5] };
    Dual<double, 2> dx(x), dy(y); dx.d[0] = 1.0; dy.d[1] = 1.0;
    Dual<double, 2> r = spline_eval(t, dx, dy);
    if (grad) { grad[0] = r.d[0]; grad[1] = r.d[1]; }
    return r.v;
}

// f(x,u,sigma) = dx_static(...) (already multiplied by sigma), plain doubles
void oracle_rhs(const ProbInfo* P, const double* drag_coef, const double* lift_coef, const double* geom,
                const double* x, const double* u, double sigma, double* out) {
    Tables tb = make_tables(drag_coef, lift_coef, geom);
    dx_static<double>(*P, tb, x, u, sigma, out);
}

// aero_force(bv, vel) in plain doubles (test hook)
void oracle_aero_force(const ProbInfo* P, const double* drag_coef, const double* lift_coef, const double* geom,
                       const double* bv, const double* vel, double* F) {
    Tables tb = make_tables(drag_coef, lift_coef, geom);
    aero_force<double>(*P, tb, bv, vel, F);
}

// simulate_zygote (dynamics.jl:308-310): value only
void oracle_rk4(const ProbInfo* P, const double* drag_coef, const double* lift_coef, const double* geom,
                const double* inp, double dt, int npts, int mode, double* out14) {
    Tables tb = make_tables(drag_coef, lift_coef, geom);
    rk4<double>(*P, tb, inp, dt, npts, mode, out14);
}

void oracle_linearize_interval(const ProbInfo* P, const double* drag_coef, const double* lift_coef, const double* geom,
                               const double* inp, double dt, int npts, int mode, double* block /*14x23*/) {
    Tables tb = make_tables(drag_coef, lift_coef, geom);
    linearize_interval(*P, tb, inp, dt, npts, mode, block);
}

// Batched form with the same array layout as scvx_linearize_batch (include/scvx_b200.h):
//   X 14 x n_nodes x B, U 3 x n_nodes x B, sigma B, params n_params in {1, B};
//   out_blocks 14 x 23 x (n_nodes-1) x B; out_lin_err 14 x (n_nodes-1) x B (endpoint_n - xbar_{n+1},
//   rocketland.jl:130,256); out_tlb 4 x n_nodes x B = [-u/|u| ; Tmin - |u|] (rocketland.jl:199-200,261-263).
// Returns the number of threads used.
// precision: 0 = IEEE double (the reference's arithmetic), 1 = IEEE binary128 evaluation rounded to double on output.
// out_sig (optional, (n_nodes-1) x B): branch signature of every interval (see sig_mix).
int oracle_linearize_batch_ex(const ProbInfo* P, int n_params, const double* drag_coef, const double* lift_coef,
                              const double* geom, const double* X, const double* U, const double* sigma, double dt,
                              int npts, int mode, int n_nodes, int B, double* out_blocks, double* out_lin_err,
                              double* out_tlb, int nthreads, int precision, uint64_t* out_sig) {
    Tables tb = make_tables(drag_coef, lift_coef, geom);
    const int ni = n_nodes - 1;
    const long total = (long)ni * B;
    int used = 1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
    used = nthreads > 0 ? nthreads : omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 4)
#endif
    for (long w = 0; w < total; ++w) {
        const int b = (int)(w / ni), i = (int)(w % ni);
        const ProbInfo& Pb = P[n_params == 1 ? 0 : b];
        const double* xb = X + ((size_t)b * n_nodes + i) * 14;
        const double* ub = U + ((size_t)b * n_nodes + i) * 3;
        double inp[21];
        for (int k = 0; k < 14; ++k) inp[k] = xb[k];
        for (int k = 0; k < 3; ++k) { inp[14 + k] = ub[k]; inp[17 + k] = ub[3 + k]; }
        inp[20] = sigma[b];
        double* blk = out_blocks + (size_t)w * 14 * 23;
        const uint64_t sg = precision == 1 ? linearize_interval_t<quad>(Pb, tb, inp, dt, npts, mode, blk)
                                           : linearize_interval_t<double>(Pb, tb, inp, dt, npts, mode, blk);
        if (out_sig) out_sig[w] = sg;
        if (out_lin_err) for (int k = 0; k < 14; ++k) out_lin_err[(size_t)w * 14 + k] = blk[k] - xb[14 + k];
    }
    if (out_tlb) {
        for (long w = 0; w < (long)n_nodes * B; ++w) {
            const int b = (int)(w / n_nodes);
            const double* ub = U + (size_t)w * 3;
            const double nu = std::sqrt(ub[0] * ub[0] + ub[1] * ub[1] + ub[2] * ub[2]);
            for (int k = 0; k < 3; ++k) out_tlb[(size_t)w * 4 + k] = -(ub[k] / nu);
            out_tlb[(size_t)w * 4 + 3] = P[n_params == 1 ? 0 : b].Tmin - nu;
        }
    }
    return used;
}

int oracle_linearize_batch(const ProbInfo* P, int n_params, const double* drag_coef, const double* lift_coef,
                           const double* geom, const double* X, const double* U, const double* sigma, double dt,
                           int npts, int mode, int n_nodes, int B, double* out_blocks, double* out_lin_err,
                           double* out_tlb, int nthreads) {
    return oracle_linearize_batch_ex(P, n_params, drag_coef, lift_coef, geom, X, U, sigma, dt, npts, mode, n_nodes, B,
                                     out_blocks, out_lin_err, out_tlb, nthreads, 0, nullptr);
}

// SURVEY.md §8f-4 variant (fin forces + aero torque, control_dim = 5): X 14 x n_nodes x B, U 5 x n_nodes x B;
// out_blocks 14 x 27 x (n_nodes-1) x B; out_lin_err 14 x (n_nodes-1) x B (optional).  Needs all three tables.
int oracle_linearize_batch_fins(const ProbInfo* P, int n_params, const double* drag_coef, const double* lift_coef,
                                const double* trq_coef, const double* geom, const double* X, const double* U,
                                const double* sigma, double dt, int npts, int mode, int n_nodes, int B,
                                double* out_blocks, double* out_lin_err, int nthreads) {
    Tables tb = make_tables3(drag_coef, lift_coef, trq_coef, geom);
    if (!tb.have_trq) return -1;
    const int ni = n_nodes - 1;
    const long total = (long)ni * B;
    int used = 1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
    used = nthreads > 0 ? nthreads : omp_get_max_threads();
#pragma omp parallel for schedule(static)
#endif
    for (long w = 0; w < total; ++w) {
        const int b = (int)(w / ni), i = (int)(w % ni);
        const double* xb = X + ((size_t)b * n_nodes + i) * 14;
        const double* ub = U + ((size_t)b * n_nodes + i) * 5;
        double inp[25];
        for (int k = 0; k < 14; ++k) inp[k] = xb[k];
        for (int k = 0; k < 10; ++k) inp[14 + k] = ub[k];
        inp[24] = sigma[b];
        double* blk = out_blocks + (size_t)w * 14 * 27;
        linearize_interval_fins(P[n_params == 1 ? 0 : b], tb, inp, dt, npts, mode, blk);
        if (out_lin_err) for (int k = 0; k < 14; ++k) out_lin_err[(size_t)w * 14 + k] = blk[k] - xb[14 + k];
    }
    return used;
}

// predict_state / simulate_zygote batched: endpoints 14 x (n_nodes-1) x B
int oracle_predict_batch(const ProbInfo* P, int n_params, const double* drag_coef, const double* lift_coef,
                         const double* geom, const double* X, const double* U, const double* sigma, double dt,
                         int npts, int mode, int n_nodes, int B, double* out, int nthreads) {
    Tables tb = make_tables(drag_coef, lift_coef, geom);
    const int ni = n_nodes - 1;
    const long total = (long)ni * B;
    int used = 1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
    used = nthreads > 0 ? nthreads : omp_get_max_threads();
#pragma omp parallel for schedule(static)
#endif
    for (long w = 0; w < total; ++w) {
        const int b = (int)(w / ni), i = (int)(w % ni);
        const double* xb = X + ((size_t)b * n_nodes + i) * 14;
        const double* ub = U + ((size_t)b * n_nodes + i) * 3;
        double inp[21];
        for (int k = 0; k < 14; ++k) inp[k] = xb[k];
        for (int k = 0; k < 3; ++k) { inp[14 + k] = ub[k]; inp[17 + k] = ub[3 + k]; }
        inp[20] = sigma[b];
        rk4<double>(P[n_params == 1 ? 0 : b], tb, inp, dt, npts, mode, out + (size_t)w * 14);
    }
    return used;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
