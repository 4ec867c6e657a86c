/* scvx_b200.h — C ABI of the B200-native linearise-and-discretise path.
 *
 * The reference (BenChung/SuccessiveConvexification) has no FFI for this path: the boundary is
 * the pair of Julia functions
 *     Dynamics.linearize_dynamics(states, tf_guess, base_dt, cache)   reference dynamics.jl:321-334
 *     Dynamics.predict_state(x, uk, up, sigma, dt, pinfo, cache)       reference dynamics.jl:315-317
 * (+ simulate_zygote / sensitivity_zygote, dynamics.jl:308-313).  The entry points below are what a
 * Julia `ccall` shim (julia/SCvxB200.jl) binds to replace them.
 *
 * Conventions
 *   - Everything is FP64.  Arrays are column-major with Julia's index order, i.e. a Julia
 *     Array{Float64,3} of size (14, n_nodes, B) is passed as-is.
 *   - State x(14) = [m, r(3), v(3), q(4, scalar first), w(3)]  (dynamics.jl:13-19); control u(3).
 *   - One unit of work = one interval (trajectory b, nodes i and i+1):
 *       inp = [x_i ; u_i ; u_{i+1} ; sigma_b]   (21)            dynamics.jl:318-320, 136-139
 *   - Output block per interval: 14 x 23 column-major  [ endpoint | A(14) | B-(3) | B+(3) | Sigma | z ]
 *     (accumulator layout of old_dynamics.jl:84-98; rocketland.jl:22-23 acc_width = 23).
 *     Columns 1..21 are LinRes.derivative (master.jl:90-93), column 0 is LinRes.endpoint,
 *     z = endpoint - D*inp (old_dynamics.jl:139, 150-153).
 *   - Pointers may be host or device memory (detected with cudaPointerGetAttributes); device
 *     pointers must live on one of the context's devices (the call runs there).  Host pointers stay caller-owned and must
 *     remain valid until the call returns (calls are synchronous for host pointers).  For device
 *     pointers the work is enqueued on the context stream (scvx_set_stream) and the call returns
 *     without synchronising.
 *   - Return value: 0 on success, negative error code otherwise; message via scvx_last_error()
 *     (thread-local).  No C++ exception crosses the ABI.  There is no CPU fallback.
 */
#ifndef SCVX_B200_H
#define SCVX_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct scvx_ctx scvx_ctx;

enum { SCVX_OK = 0, SCVX_ERR_ARG = -1, SCVX_ERR_CUDA = -2, SCVX_ERR_STATE = -3, SCVX_ERR_NOMEM = -4 };

/* aero_kind: ExoatmosphericData (master.jl:8) vs AtmosphericData (master.jl:10-16) */
enum { SCVX_AERO_EXO = 0, SCVX_AERO_TABLE = 1 };
/* RK4 stage rule: LITERAL reproduces dynamics.jl:126-128 (stage increments not scaled by the
 * sub-step); TEXTBOOK scales them (classical RK4).  SURVEY.md §0.4 */
enum { SCVX_MODE_LITERAL = 0, SCVX_MODE_TEXTBOOK = 1 };
/* tables of AtmosphericData (drag_itrp, lift_itrp, trq_itrp; aerodynamics.jl:19-21) */
enum { SCVX_TABLE_DRAG = 0, SCVX_TABLE_LIFT = 1, SCVX_TABLE_TORQUE = 2 };
/* kernel selection: AUTO picks the fastest validated path */
enum { SCVX_KERNEL_AUTO = 0, SCVX_KERNEL_DUALWARP = 1, SCVX_KERNEL_STAGED = 2 };

/* Mirror of ProbInfo (master.jl:73-83) + AtmosphericData scalars (master.jl:14-15) + Tmin
 * (master.jl:21, used by the thrust-lower-bound rows rocketland.jl:199-200).  3x3 column-major. */
typedef struct scvx_probinfo {
    double a, g0, sos;
    double jB[9], jBi[9];
    double rTB[3], rFB[3];
    double force_scalar, length_scalar;
    double Tmin;
    int32_t aero_kind;
    int32_t _pad;
} scvx_probinfo;

/* The dimensional DescentProblem fields (master.jl:17-71) that normalize_problem (sample_problems.jl:5-23), ProbInfo
 * (master.jl:73-83) and linear_points (initial_solve.jl:113-129) consume; shared by a dispersion batch. */
typedef struct scvx_dim_problem {
    double g, mdry, mwet, Tmin, Tmax, alpha, sos, tf_guess;
    double jB[9], rTB[3], rFB[3], rIf[3];
    int32_t aero_kind;
    int32_t K;
} scvx_dim_problem;

/* Number of visible CUDA devices (0 if none / no driver). */
int scvx_device_count(void);
/* Thread-local message of the last failing call on this thread. */
const char* scvx_last_error(void);
/* Library / ABI version (major*1000 + minor). */
int scvx_version(void);
/* sizeof(scvx_probinfo) the library was built with (binding sanity check). */
int scvx_sizeof_probinfo(void);

/* Opaque context: owns device memory, streams and the staged tables of every listed device.
 * Stands in for the reference's IntegratorCache (master.jl:113-120; ctor dynamics.jl:258-286).
 * n_dev > 1 shards the trajectories of host-pointer calls over the devices in contiguous blocks. */
int scvx_create(scvx_ctx** out, const int* device_ids, int n_dev);
void scvx_destroy(scvx_ctx* ctx);

/* n == 1: parameters shared by all trajectories; n == B: one record per trajectory.  `p` is a host pointer.
 * Records that differ in `a` and `Tmin` only (mass / thrust-bound sweeps) run at the speed of shared parameters; records
 * that differ in any other field take a slower path that reads whole records from device memory (about -8 %). */
int scvx_set_params(scvx_ctx* ctx, const scvx_probinfo* p, int n);

/* Upload one aero table.  `samples` (host): n_cos x n_mach column-major exactly as the reshape at
 * aerodynamics.jl:19-21; axes are the StepRangeLen's cos0:dcos:.. and mach0:dmach:..
 * (aerodynamics.jl:17-18).  prefiltered == 0: the library runs the cubic-B-spline prefilter
 * (Interpolations.jl BSpline(Cubic(Line(OnGrid())))) on the device; prefiltered != 0: `samples`
 * already holds the (n_cos+2) x (n_mach+2) coefficients. */
int scvx_set_aero_table(scvx_ctx* ctx, int which, const double* samples, int n_cos, int n_mach,
                        double cos0, double dcos, double mach0, double dmach, int prefiltered);
/* Read back the staged coefficients ((n_cos+2) x (n_mach+2) doubles, host pointer). */
int scvx_get_aero_coefficients(scvx_ctx* ctx, int which, double* out);

/* linearize_dynamics for B trajectories of n_nodes nodes each (dynamics.jl:321-334 with the
 * rk4 + forward-Jacobian body of dynamics.jl:112-134, 311-313).
 *   X          14 x n_nodes x B       U  3 x n_nodes x B       sigma  B
 *   out_blocks 14 x 23 x (n_nodes-1) x B                          (required)
 *   out_lin_err 14 x (n_nodes-1) x B   endpoint_n - x_{n+1}       (optional, may be NULL; rocketland.jl:130, 256)
 *   out_tlb     4 x n_nodes x B        [-u/|u| ; Tmin - |u|]      (optional, may be NULL; rocketland.jl:199-200, 261-263)
 * B == 1 reproduces one call of the reference function. */
int scvx_linearize_batch(scvx_ctx* ctx, const double* X, const double* U, const double* sigma,
                         double base_dt, int npts, int mode, int n_nodes, int B,
                         double* out_blocks, double* out_lin_err, double* out_tlb);

/* SURVEY.md §8f-4 — fin forces and aero torque (control_dim 3 -> 5).  The reference carries these terms as COMMENTS and
 * never runs them (control_dim = 3, rocketland.jl:17; ff and bdy_trq commented out at dynamics.jl:63, 66, 69), so there is
 * NO REFERENCE CONSUMER and no reference behaviour to match; the entry point restores exactly the commented expressions:
 *   ff = u[4]*fd1 + u[5]*fd2, fd1 = normalize((C(q) e2) x v), fd2 = fd1 x v          dynamics.jl:60-63
 *   aero_frc = aerf + ff;  aero_trq = cross(rFB, ff) + bdy_trq                           dynamics.jl:66, 69
 *   bdy_trq = trq_itrp(cos_aoa, mach) * length_scalar * force_scalar * normalize(v x bv) (zero when |dp| >= 0.95)
 *                                                                                         aerodynamics.jl:45, 49-56
 * Needs aero_kind = TABLE and all three tables (drag, lift, torque).  The torque makes wdot depend on q and v, so the
 * structural zeros the STAGED kernels exploit are gone: this variant runs on the generic forward-mode kernel (one warp
 * per interval, lane L < 25 carries d/d inp[L]), exact by construction, about an order of magnitude slower than the STAGED path.
 *   X 14 x n_nodes x B    U5 5 x n_nodes x B    sigma B      inp = [x; u_k(5); u_{k+1}(5); sigma] (25)
 *   out_blocks 14 x 27 x (n_nodes-1) x B = [ endpoint | D (25 columns) | z ]   (acc_width = 14 + 2*5 + 3, rocketland.jl:22)
 *   out_lin_err 14 x (n_nodes-1) x B (optional).  All host or all device pointers. */
int scvx_linearize_batch_fins(scvx_ctx* ctx, const double* X, const double* U5, const double* sigma,
                              double base_dt, int npts, int mode, int n_nodes, int B,
                              double* out_blocks, double* out_lin_err);

/* Fin-force tables (aero/fin.csv: columns lift, drag over 60 Mach numbers x 901 deflection angles, written by
 * aero/AeroTable.jl:94-112; read and dropped at aerodynamics.jl:23-26 — no consumer in the reference).  Staged like the
 * other tables (same cubic B-spline prefilter and Flat extrapolation): which = 0 lift, 1 drag; samples n_mach x n_defl
 * column-major (Mach fastest, the file's row order).  scvx_fin_force_batch evaluates both splines for n (mach,
 * deflection) pairs: the lookup a fin-deflection control model would build on.  Host or device pointers. */
int scvx_set_fin_table(scvx_ctx* ctx, int which, const double* samples, int n_mach, int n_defl,
                       double mach0, double dmach, double defl0, double ddefl, int prefiltered);
int scvx_fin_force_batch(scvx_ctx* ctx, const double* mach, const double* deflection, int n,
                         double* out_lift, double* out_drag);

/* predict_state / simulate_zygote for every interval (value only): out 14 x (n_nodes-1) x B. */
int scvx_predict_batch(scvx_ctx* ctx, const double* X, const double* U, const double* sigma,
                       double base_dt, int npts, int mode, int n_nodes, int B, double* out_endpoints);

/* Compact result format for hosts on the far side of PCIe.  92 of the 322 entries of a block are structural constants of
 * the dynamics (nothing depends on position, the mass row and the q / w rows have fixed zero
 * patterns), so the host path is bound by bytes it does not need.  The compact record of one interval holds only the
 * data entries, in block (column-major) order, SCVX_COMPACT_DOUBLES = 230 doubles:
 *   slots 0..228  the data entries of the block (scvx_compact_layout gives the dense offset of each slot):
 *                 endpoint 14 | d/dm rows r,v 6 | d/dv rows r,v 18 | d/dq rows r,v,q 40 | d/dw rows r,v,q,w 39 |
 *                 B- 42 | B+ 42 | Sigma 14 | z 14
 *   slot  229     status word: 0.0 if every entry of the dense block is finite, 1.0 otherwise (per-interval non-finite
 *                 flag; the constants of a flagged interval's dense block are not guaranteed)
 * lin_err is not shipped: it is endpoint - x_{n+1}, one IEEE subtraction the host repeats exactly (scvx_expand_compact).
 * layout SCVX_COMPACT_NO_Z drops the z column as well (the reference's SOCP consumes D and lin_err only,
 * rocketland.jl:123-133, 251-258; z is the named output of old_dynamics.jl:139): 215 data entries + status word =
 * SCVX_COMPACT_NO_Z_DOUBLES = 216 doubles; the expander then re-forms z = endpoint - D*inp on the host in FP64 (same
 * value to rounding, not the same bits: the summation order differs from the device's).
 *   out_compact R x (n_nodes-1) x B, R = scvx_compact_record_doubles(layout)      out_tlb 4 x n_nodes x B (optional)
 * Host or device pointers, as scvx_linearize_batch. */
enum { SCVX_COMPACT_FULL = 0, SCVX_COMPACT_NO_Z = 1 };
#define SCVX_COMPACT_DOUBLES 230
#define SCVX_COMPACT_DATA 229
#define SCVX_COMPACT_NO_Z_DOUBLES 216
int scvx_compact_record_doubles(int layout);      /* 230, 216, or a negative error code */
int scvx_linearize_batch_compact(scvx_ctx* ctx, const double* X, const double* U, const double* sigma,
                                 double base_dt, int npts, int mode, int n_nodes, int B, int layout,
                                 double* out_compact, double* out_tlb);
/* Dense offset (column * 14 + row, 0 <= offset < 322) of compact slots 0..228; no device work.  `index` has room for
 * SCVX_COMPACT_DATA int32. */
int scvx_compact_layout(int32_t* index);
/* Host-side expander (plain C loop on `n_threads` host threads, no device work; all pointers are HOST memory):
 * compact R x K x B  ->  out_blocks 14 x 23 x K x B (may be NULL) with the structural constants filled in, equal to what
 * scvx_linearize_batch writes for every interval whose status word is 0 (layout NO_Z: except the z column, see above);
 * out_lin_err 14 x K x B (may be NULL).  X 14 x n_nodes x B is needed for out_lin_err and for NO_Z; U 3 x n_nodes x B
 * and sigma B for NO_Z only (may be NULL otherwise).
 * Returns the number of intervals whose status word is non-zero (>= 0), or a negative error code. */
int64_t scvx_expand_compact(const double* compact, int layout, const double* X, const double* U, const double* sigma,
                            int n_nodes, int B, double* out_blocks, double* out_lin_err, int n_threads);

/* Page-locked host memory.  The chunked H2D / kernel / D2H pipeline of host-pointer calls overlaps copies with compute
 * only for page-locked buffers (pageable memory makes every cudaMemcpyAsync a staged, serialising copy): allocate result
 * arrays here, or pin arrays the host language owns (e.g. Julia Arrays) for the lifetime of the calls. */
int scvx_host_alloc(void** out, uint64_t bytes);
int scvx_host_free(void* p);
int scvx_host_register(void* p, uint64_t bytes);
int scvx_host_unregister(void* p);

/* Fused cost / defect evaluation of the SCvx ratio test (rocketland.jl:289-290): per trajectory
 *   defect_b = sqrt( sum_k || x_{k+1} - endpoint_k ||^2 )   (= Julia's norm over the K defect vectors, = ||lin_err||_F)
 *   cost_b   = -X[0, n_nodes-1, b] + wNu * defect_b          (J_k; mass of the last node enters with weight -1)
 * from the lin_err array produced by scvx_linearize_batch at the same inputs, so the accept/reject test needs one
 * tiny transfer.  lin_err 14 x (n_nodes-1) x B, X 14 x n_nodes x B; out_defect B (required), out_cost B (may be NULL). */
int scvx_defect_cost_batch(scvx_ctx* ctx, const double* X, const double* lin_err, int n_nodes, int B, double wNu,
                           double* out_defect, double* out_cost);

/* Batched initial guess: `linear_points` of the reference (initial_solve.jl:113-129) for B
 * dispersed initial conditions.  Per trajectory b and node k = 0..K: mass, position and velocity on a straight line from
 * (mwet_b, rIi_b, vIi_b) to (mdry, rIf, vIf); attitude = rotation_between([1,0,0], -v_k) (Rotations.jl, scalar-first
 * quaternion); zero angular rate; control = [m_k * g, 0, 0].
 *   rIi, vIi 3 x B;  mwet B (NULL: mwet_shared for every trajectory);  rIf, vIf 3 (host pointers, always)
 *   X 14 x (K+1) x B, U 3 x (K+1) x B.  rIi/vIi/mwet/X/U are all host or all device pointers. */
int scvx_linear_points_batch(scvx_ctx* ctx, const double* rIi, const double* vIi, const double* mwet, double mwet_shared,
                             double mdry, const double* rIf, const double* vIf, double g, int K, int B,
                             double* X, double* U);

/* Per-trajectory problem set-up for Monte-Carlo dispersions of a DIMENSIONAL problem, one launch:
 *   normalize_problem (sample_problems.jl:5-23) with Ul_b = max(rIi_b), Ut = tf_guess, Um = mwet_b (vIf := vIi as there),
 *   ProbInfo of the normalised problem (master.jl:73-83: a, g0, sos, jB, inv(jB), rTB, rFB; aero scalars of
 *   rescale_aerodata, aerodynamics.jl:30-36) and its linear_points initial guess (initial_solve.jl:113-129).
 * In : rIi, vIi 3 x B, mwet B (NULL: base->mwet) — dimensional.
 * Out: X 14 x (K+1) x B, U 3 x (K+1) x B, sigma B (= tf_guess / Ut), scales 3 x B = [Ul, Ut, Um] (may be NULL),
 *      out_params B records (may be NULL).  install != 0 makes the records the context's per-trajectory parameters, as if
 *      scvx_set_params(ctx, records, B) had been called (no host round trip).  All arrays host or all device pointers. */
int scvx_dispersed_setup_batch(scvx_ctx* ctx, const scvx_dim_problem* base, const double* rIi, const double* vIi,
                               const double* mwet, int B, double* X, double* U, double* sigma, double* scales,
                               scvx_probinfo* out_params, int install);

/* Fixed-pattern sparse SOCP rows.  The reference refreshes the K dynamics equality blocks
 * (rocketland.jl:117-133, 251-258) and the K+1 linearised thrust-lower-bound rows (rocketland.jl:194-201, 260-265)
 * with K*21 + 3(K+1) MOI.modify calls per iteration; their sparsity never changes.  These entry points give that
 * sub-matrix in compressed-sparse-column form for a direct conic-solver interface:
 *   rows     14n + i (dynamics row i of interval n, Zeros cone), then 14K + n (thrust lower bound of node n, Nonpositives)
 *   columns  dxv[j,n] -> 14n + j;  duv[j,n] -> 14(K+1) + 3n + j;  dsig -> 17(K+1);  nuv[j,n] -> 17(K+1) + 1 + 14n + j
 *            (the reference's variable creation order, rocketland.jl:73-76: local column j = its variable 17(K+1) + j)
 *   values   A_n, B-_n, B+_n, Sigma_n (all 14 x 21 entries, structural zeros kept as eachcol() emits them), +1 on
 *            nuv[:,n+1], -1 on dxv[:,n+1], H_n = -u_n/|u_n| on duv[:,n];  constants  const = [lin_err ; Tmin - |u_n|]
 * with K = n_nodes - 1, n_rows = 15K + 1, n_cols = 31(K+1) + 1, nnz = 325K + 3.
 * SIGN CONVENTION: `const` is the constant term of the MOI VectorAffineFunction, i.e. the constraints read
 *   M x + const  in Zeros        (rows 0 .. 14K-1,   rocketland.jl:129-131)
 *   M x + const  in Nonpositives (rows 14K .. 15K,   rocketland.jl:199-201).
 * A solver interface of the form  A x = b,  G x <= h  (ECOS / SCS style) takes  b = -const[0:14K],  h = -const[14K:]. */
int scvx_socp_dims(int n_nodes, int* n_rows, int* n_cols, int* nnz);
/* Pattern (host arrays, no device work): colptr n_cols + 1, rowind nnz; rows ascend inside a column. */
int scvx_socp_pattern(int n_nodes, int32_t* colptr, int32_t* rowind);
/* Values for B trajectories from the outputs of scvx_linearize_batch at the same inputs:
 *   blocks 14 x 23 x K x B, lin_err 14 x K x B, tlb 4 x n_nodes x B  ->  out_vals nnz x B, out_const n_rows x B (may be
 *   NULL; lin_err may be NULL then).  All host or all device pointers (device: enqueued on the context's stream).
 *   n_nodes <= 50000 (one launch covers 340 K + 4 value/constant indices in blocks of 256 along gridDim.y). */
int scvx_socp_values_batch(scvx_ctx* ctx, const double* blocks, const double* lin_err, const double* tlb, int n_nodes, int B,
                           double* out_vals, double* out_const);

/* Stream used for device-pointer calls (a cudaStream_t of the device the pointers live on).  NULL = back to the library's
 * own non-blocking stream; to run on the legacy default stream pass cudaStreamLegacy ((cudaStream_t)0x1), which is what
 * frameworks mean by "stream 0". */
int scvx_set_stream(scvx_ctx* ctx, void* cuda_stream);
/* Kernel selection (SCVX_KERNEL_*). */
int scvx_set_kernel(scvx_ctx* ctx, int which);
/* Wait for everything enqueued by this context. */
int scvx_synchronize(scvx_ctx* ctx);
/* Number of kernels the library launched since the context was created (all devices). */
int64_t scvx_launch_count(scvx_ctx* ctx);
/* Device-event duration (ms) of the kernels of the last device-pointer linearize/predict call on the
 * first device; synchronises the stream. */
int scvx_last_kernel_ms(scvx_ctx* ctx, double* ms);

#ifdef __cplusplus
}
#endif
#endif /* SCVX_B200_H */
