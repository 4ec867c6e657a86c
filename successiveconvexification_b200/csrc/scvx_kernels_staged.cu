// scvx_kernels_staged.cu — the STAGED linearise-and-discretise path (sm_100a, FP64): the default.
//
// The exact forward-mode Jacobian of the reference's rk4 (dynamics.jl:112-134, 311-313) is the tangent
// recursion  K_s = J_x(Y_s) * Yt_s + J_u(Y_s) * U_s + e_sigma f(Y_s)  over the 4*npts stages.
// The value trajectory does not depend on the tangents, so the work is split in two kernels, each in the shape in
// which it parallelises:
//
//  A   stage_value_kernel   : one THREAD per interval.  Integrates the 14-state value with rk4 and evaluates the aero
//      force together with its Jacobians dF/dv, dF/db (b = C(q) e1).  Writes the endpoint (block column 0), lin_err,
//      the thrust-lower-bound rows and a 43-double "stage record" per stage (stage state m,v,q,w; stage control u;
//      unscaled rhs f; dF/dv; dF/db), laid out [group of 32 intervals][stage][entry][32 lanes] so that one stage of
//      one group is one contiguous 11 KB slab (= one TMA bulk copy).  It also carries two of the four light tangent
//      columns (d/dv1, d/dv2; state in shared memory), the constant position columns and the partial z; the other two
//      (d/dm, d/dv0) ride in the two otherwise idle column slots of the tangent kernel.
//
//  B   tangent_kernel       : persistent, warp-specialised, one 384-thread CTA per SM, 32 intervals per pass.  The 14
//      heavy tangent columns (+ 2 light ones): 8 consumer warps, 8 lanes per interval, two full columns per lane,
//      tangent state (S, stage sum T, stage tangent, r-row quadrature = 156 registers) in registers.  The
//      sigma*c_i-scaled Jacobian blocks (78 doubles per interval and stage) are formed by 4 dedicated producer warps with
//      lane = interval (no redundancy) from TMA-staged stage records into a shared-memory ring and read back by the 8
//      lanes of an interval as broadcast 128-bit loads.  Ring slabs are handed over with mbarriers once per rk4 step;
//      registers are re-partitioned between the roles with setmaxnreg.  [A|B-|B+|Sigma] go from registers to the
//      14x23 block with 16-byte stores; z is reduced over the 8 lanes and accumulated in place.
#include "scvx_staged_dev.cuh"
#include <cstdlib>
#include "scvx_kernels.h"

namespace {

#ifndef SCVX_A_MINBLOCKS
#define SCVX_A_MINBLOCKS 2
#endif
// The kernel also propagates two of the four light tangent columns d/d(m, v0, v1, v2) — d/dv1 and d/dv2 (only their v
// and r rows are non-trivial: K_v = Jvv Y_v, r rows a quadrature of the v rows; d/dm and d/dv0 ride in the two idle
// slots of the tangent kernel).  Their 24 doubles of state live in shared memory, lane-private [entry][thread]
// (conflict-free), because the value chain already uses every register;
// the Jacobian blocks they need (dF/dv, f_v, m) are at hand here, so no separate pass re-reads the stage records
// (a separate one-thread-per-interval kernel was latency bound on those reads: 0.13 ms per chunk vs +0.03 ms here).
#ifndef SCVX_A_PARK
#define SCVX_A_PARK 0
#endif
// SCVX_A_PARK: the step-start state x and the rk4 accumulator (28 doubles) also live in lane-private shared memory, so
// that the kernel fits 168 registers and a third block per SM (3 warps per scheduler hide the dependent-issue latency of
// the serial value chain better than 2).
// SCVX_A_STAGE_UNROLL: unrolling of the four-stage loop of the value kernel (1 = none)
#ifndef SCVX_A_STAGE_UNROLL
#define SCVX_A_STAGE_UNROLL 1
#endif
#ifndef SCVX_A_PREFETCH_EPILOGUE
#define SCVX_A_PREFETCH_EPILOGUE 1
#endif
#ifndef SCVX_A_SMEM_TABLES
#define SCVX_A_SMEM_TABLES 0
#endif
// Table staging of the value kernel, template parameter TS (BASELINE.json's north_star: "the aero tables are staged into
// shared memory"; A/Bs in profiles/r2_smem_tables_variants.txt, r2_smem_tables_small_record.txt, r2_smem_window_variant.txt):
//   0  spline coefficients read through the read-only path (ld.global.nc), L1 resident; two blocks of 128 threads per SM.
//      The default (SCVX_A_SMEM_TABLES); always compiled: the fallback when the tables do not fit, and the path of
//      exo-atmospheric batches;
//   1  DRAG table (92 KB: every stage reads it) staged in shared memory by every block, one block of 256 threads per SM,
//      the lift table (one branch only) through L1: +0.8 % with the LITERAL rule at sigma ~ U(1, 15), -2.6 % with the
//      TEXTBOOK rule (one 256-thread block per SM drains less evenly than two of 128) — not the default;
//   2  drag + lift tables (184 KB) staged; one block of 224 threads per SM (what fits beside the light-column state): -1.7 %;
//   (a PATCH-EXPANDED global copy — the 16 coefficients of a cell and table in one 128-byte line — was measured too:
//      -6 % / -8 %, profiles/r2_ab_patch_table.txt: neighbouring cells no longer share lines, the copy is 16x the table)
//   3  a WIN_I x WIN_J WINDOW of both tables around the block's starting (cos aoa, Mach) cells (18 KB per block; two blocks
//      of 128 threads per SM as in 0), 4 x 4 patches outside the window from global memory: -6.8 % (generic loads, spills).
#ifndef SCVX_A_VT
#define SCVX_A_VT 128
#endif
__host__ __device__ constexpr int value_threads(int ts) { return (ts == 1) ? 256 : (ts == 2 ? 224 : SCVX_A_VT); }
__host__ __device__ constexpr int value_minblocks(int ts) { return (ts == 1 || ts == 2) ? 1 : SCVX_A_MINBLOCKS; }
constexpr int VALUE_SMEM_DOUBLES = 24 + (SCVX_A_PARK ? 28 : 0);
__host__ __device__ constexpr size_t value_smem_bytes(int ts) {
    return (size_t)VALUE_SMEM_DOUBLES * value_threads(ts) * sizeof(double) + (ts == 3 ? 2 * WIN_I * WIN_J * sizeof(double) : 0);
}
// Programmatic dependent launch: the tangent kernel carries cudaLaunchAttributeProgrammaticStreamSerialization and the
// value kernel releases its dependents as its blocks start, so the blocks of T_c become resident — barriers initialised —
// on the SMs that V_c's last wave leaves, and wait there for V_c to complete: +0.15 %.  Measured and not adopted
// (profiles/r2_ab_pdl_*.txt): T_c starting its passes as the value-kernel blocks publish their records through per-block
// flags (+0.05 % more, for a spin-wait in the product); V_{c+1} moving into the SMs that T_c's blocks leave, over two
// record buffers (-1.3 to -1.5 %, wherever T_c releases its dependents).
// Without the launch attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <int SP, int TS>
__global__ void __launch_bounds__(value_threads(TS), value_minblocks(TS)) stage_value_kernel(const __grid_constant__ StagedArgs a) {
    constexpr int VT = value_threads(TS);
    extern __shared__ double light_smem[];            // [24][VT] doubles (+ the staged tables)
    pdl_launch_dependents();                          // T_c may take the SMs this grid's last wave leaves (it waits for this grid)
    ScvxTables tbl = a.tb;
    if constexpr (TS == 1 || TS == 2) {
        const int ncoef = (a.tb.n1 + 2) * (a.tb.n2 + 2);
        double* sd = light_smem + VALUE_SMEM_DOUBLES * VT;
        for (int k = threadIdx.x; k < ncoef; k += VT) sd[k] = __ldg(a.tb.drag + k);
        tbl.drag = sd;
        if constexpr (TS == 2) {
            for (int k = threadIdx.x; k < ncoef; k += VT) sd[ncoef + k] = __ldg(a.tb.lift + k);
            tbl.lift = sd + ncoef;
        }
        __syncthreads();
    }
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if constexpr (TS != 3) {
        if (t >= a.n_groups * GROUP) return;
    }
    const ScvxBatch& bt = a.bt;
    const int ni = bt.n_nodes - 1;
    const bool live = t < a.count;
    const int w = a.first + (live ? t : a.count - 1);          // padded lanes recompute the last interval
    const int b = (int)(w / ni), i = (int)(w % ni);
    const scvx_probinfo& P = SP ? a.Pc : bt.P[bt.n_params == 1 ? 0 : b];
    // SP == 2 (a mass / thrust-bound sweep): the records differ in `a` and `Tmin` only; those two come from the
    // trajectory's own record, everything else from the shared one in the kernel arguments
    const double pa = (SP == 2) ? __ldg(&bt.P[b].a) : ldp<SP>(&P.a);
    const double* xin = bt.X + ((size_t)b * bt.n_nodes + i) * 14;
    const double* uin = bt.U + ((size_t)b * bt.n_nodes + i) * 3;
    const double sigma = bt.sigma[b];
    double x[14], um[3], up[3];
#pragma unroll
    for (int r = 0; r < 14; ++r) x[r] = xin[r];
#pragma unroll
    for (int c = 0; c < 3; ++c) { um[c] = uin[c]; up[c] = uin[3 + c]; }
    if constexpr (TS == 3) {
        // window of the spline coefficients around the block's starting (cos aoa, Mach) cells
        __shared__ int wmm[4];                               // min i, max i, min j, max j over the block
        if (threadIdx.x == 0) { wmm[0] = 1 << 30; wmm[1] = -1; wmm[2] = 1 << 30; wmm[3] = -1; }
        __syncthreads();
        if (a.tb.drag != nullptr && ldpi<SP>(&P.aero_kind) == SCVX_AERO_TABLE) {
            const double q0 = x[7], q1 = x[8], q2 = x[9], q3 = x[10];
            const double b0 = 1.0 - 2.0 * (q2 * q2 + q3 * q3), b1 = 2.0 * (q1 * q2 + q0 * q3), b2 = 2.0 * (q1 * q3 - q0 * q2);
            const double vv = x[4] * x[4] + x[5] * x[5] + x[6] * x[6], nv = sqrt(vv);
            double ca = (b0 * x[4] + b1 * x[5] + b2 * x[6]) / (nv * sqrt(b0 * b0 + b1 * b1 + b2 * b2));
            ca = fmin(fmax(ca, -1.0), 1.0);
            const double mach = nv / ldp<SP>(&P.sos);
            double xi = (ca - a.tb.x0) * a.tb.inv_dx + 1.0, yi = (mach - a.tb.y0) * a.tb.inv_dy + 1.0;
            xi = fmin(fmax(xi, 1.0), (double)a.tb.n1); yi = fmin(fmax(yi, 1.0), (double)a.tb.n2);
            const int ci = (xi == xi) ? (int)xi : 1, cj = (yi == yi) ? (int)yi : 1;
            atomicMin(&wmm[0], ci); atomicMax(&wmm[1], ci); atomicMin(&wmm[2], cj); atomicMax(&wmm[3], cj);
        }
        __syncthreads();
        const int L1 = a.tb.n1 + 2, L2 = a.tb.n2 + 2;
        int wi0 = -(1 << 29), wj0 = -(1 << 29);              // "no window": no patch is ever inside
        double* wd = light_smem + VALUE_SMEM_DOUBLES * VT;
        if (wmm[1] >= 0 && L1 >= WIN_I && L2 >= WIN_J) {
            wi0 = min(max((wmm[0] + wmm[1]) / 2 - WIN_I / 2, 0), L1 - WIN_I);
            wj0 = min(max((wmm[2] + wmm[3]) / 2 - WIN_J / 2, 0), L2 - WIN_J);
            for (int k = threadIdx.x; k < WIN_I * WIN_J; k += VT) {
                const int li = k % WIN_I, lj = k / WIN_I;
                const size_t g = (size_t)(wi0 + li) + (size_t)(wj0 + lj) * L1;
                wd[k] = __ldg(a.tb.drag + g);
                wd[WIN_I * WIN_J + k] = __ldg(a.tb.lift + g);
            }
        }
        __syncthreads();
        tbl.wdrag = wd; tbl.wlift = wd + WIN_I * WIN_J; tbl.wi0 = wi0; tbl.wj0 = wj0;
        if (t >= a.n_groups * GROUP) return;
    }

    const int nst = 4 * bt.npts;
    double* rec = a.rec + ((size_t)(t >> 5) * nst) * ((size_t)a.rec_n * GROUP) + (t & 31);
    const double h = bt.dt / (double)bt.npts;
    const double pcs = 1.0 / (double)bt.npts;
    const double s = (bt.mode == SCVX_MODE_LITERAL) ? 1.0 : h;
    // light-column state (columns d/dv1, d/dv2): L[(c*12 + kind*3 + r)*128 + tid], kind 0 = S, 1 = acc, 2 = Y, 3 = r-row sum
    double* L = light_smem + threadIdx.x;
    {
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const double s0 = (c + 1 == r) ? 1.0 : 0.0;
                L[(c * 12 + 0 + r) * VT] = s0; L[(c * 12 + 3 + r) * VT] = 0.0;
                L[(c * 12 + 6 + r) * VT] = s0; L[(c * 12 + 9 + r) * VT] = 0.0;
            }
    }
    double pca = 0.0;
#if SCVX_A_PARK
    double* PX = L + 24 * VT;            // step-start state [14][128]
    double* PA = L + 38 * VT;            // rk4 accumulator  [14][128]
#pragma unroll
    for (int r = 0; r < 14; ++r) PX[r * VT] = x[r];
#endif
    for (int it = 0; it < bt.npts; ++it) {
#if SCVX_A_PREFETCH_EPILOGUE
        // the epilogue reads this node's and the next node's state and controls again (z, lin_err, thrust rows): by then
        // they have left the caches, and nothing is left to overlap the loads with (6 % of the kernel's stall samples).
        // One L2 prefetch per line before the last step brings them back in time.
        if (it == bt.npts - 1) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(xin));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(xin + 14));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(xin + 27));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(uin + 5));
        }
#endif
        double y[14];
#if SCVX_A_PARK
#pragma unroll
        for (int r = 0; r < 14; ++r) { y[r] = PX[r * VT]; PA[r * VT] = 0.0; }
#else
        double acc[14];
#pragma unroll
        for (int r = 0; r < 14; ++r) { y[r] = x[r]; acc[r] = 0.0; }
#endif
#if SCVX_A_STAGE_UNROLL == 4
#pragma unroll
#elif SCVX_A_STAGE_UNROLL == 2
#pragma unroll 2
#else
#pragma unroll 1
#endif
        for (int st = 0; st < 4; ++st) {
            const double pc = (st == 0) ? pca : (st == 3 ? pca + pcs : pca + 0.5 * pcs);
            double uc[3], f[14], Fv[3][3], Fb[3][3];
#pragma unroll
            for (int c = 0; c < 3; ++c) uc[c] = (1.0 - pc) * um[c] + pc * up[c];
            rhs_value<true, TS, SP>(P, pa, tbl, y, uc, f, Fv, Fb);
            // record: m, v, q, w, f_v [, dF/dv, dF/db]  (u, f_m, f_q, f_w are re-formed by the producers)
            double* rp = rec + (size_t)(it * 4 + st) * ((size_t)a.rec_n * GROUP);
            if (a.rec_n == REC_AERO) {
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        rp[(R_AV + 3 * r + c) * GROUP] = Fv[r][c];
                        rp[(R_AB + 3 * r + c) * GROUP] = Fb[r][c];
                    }
            }
            rp[0 * GROUP] = y[0];
#pragma unroll
            for (int r = 0; r < 10; ++r) rp[(1 + r) * GROUP] = y[4 + r];
#pragma unroll
            for (int c = 0; c < 3; ++c) rp[(R_FV + c) * GROUP] = f[4 + c];
            const double wgt = (st == 0 || st == 3) ? 1.0 : 2.0;
            const double cy = (st == 2) ? s : 0.5 * s;
            {
                const double smv = sigma / y[0];
                const double csg = h * (1.0 / 6.0) * wgt * sigma;
                double Jvv[3][3];
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c = 0; c < 3; ++c) Jvv[r][c] = smv * Fv[r][c];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    double* Lc = L + c * 12 * VT;
                    const double y0 = Lc[6 * VT], y1 = Lc[7 * VT], y2 = Lc[8 * VT];
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        const double yr = (r == 0) ? y0 : (r == 1 ? y1 : y2);
                        Lc[(9 + r) * VT] = fma(csg, yr, Lc[(9 + r) * VT]);
                        const double K = fma(Jvv[r][0], y0, fma(Jvv[r][1], y1, Jvv[r][2] * y2));
                        if (st != 3) {
                            Lc[(3 + r) * VT] = fma(wgt, K, Lc[(3 + r) * VT]);
                            Lc[(6 + r) * VT] = fma(cy, K, Lc[r * VT]);
                        } else {
                            const double sn = fma(h * (1.0 / 6.0), Lc[(3 + r) * VT] + K, Lc[r * VT]);
                            Lc[r * VT] = sn; Lc[(6 + r) * VT] = sn; Lc[(3 + r) * VT] = 0.0;
                        }
                    }
                }
            }
#if SCVX_A_PARK
#pragma unroll
            for (int r = 0; r < 14; ++r) {
                const double k = f[r] * sigma;
                PA[r * VT] = fma(wgt, k, PA[r * VT]);
                y[r] = fma(cy, k, PX[r * VT]);
            }
#else
#pragma unroll
            for (int r = 0; r < 14; ++r) {
                const double k = f[r] * sigma;
                acc[r] = fma(wgt, k, acc[r]);
                y[r] = fma(cy, k, x[r]);
            }
#endif
        }
        pca += pcs;
#if SCVX_A_PARK
#pragma unroll
        for (int r = 0; r < 14; ++r) PX[r * VT] = fma(h * (1.0 / 6.0), PA[r * VT], PX[r * VT]);
#else
#pragma unroll
        for (int r = 0; r < 14; ++r) x[r] = fma(h * (1.0 / 6.0), acc[r], x[r]);
#endif
    }
#if SCVX_A_PARK
#pragma unroll
    for (int r = 0; r < 14; ++r) x[r] = PX[r * VT];
#endif
    if (!live) return;
    double* blk = bt.out_blocks + (size_t)w * SCVX_BLOCK_DOUBLES;
#pragma unroll
    for (int r = 0; r < 14; r += 2) *reinterpret_cast<double2*>(blk + r) = make_double2(x[r], x[r + 1]);
    {
        // columns d/d(r, v1, v2) of the block and the partial z = endpoint - D[:, r v1 v2] * inp[r v1 v2]
        // (the columns d/dm and d/dv0 ride in the two otherwise idle slots of the tangent kernel)
        double z[7];
#pragma unroll
        for (int r = 0; r < 7; ++r) z[r] = x[r];
#pragma unroll
        for (int c = 1; c < 7; ++c) {               // inp 1..3 (r), 5..6 (v1, v2)
            if (c == 4) continue;
            double col[14];
#pragma unroll
            for (int r = 0; r < 14; ++r) col[r] = 0.0;
            if (c >= 5) {
                const double* Lc = L + (c - 5) * 12 * VT;
                const double xc = xin[c];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    col[1 + r] = Lc[(9 + r) * VT]; col[4 + r] = Lc[r * VT];
                    z[1 + r] = fma(-col[1 + r], xc, z[1 + r]); z[4 + r] = fma(-col[4 + r], xc, z[4 + r]);
                }
            } else {
                col[c] = 1.0;                        // nothing depends on position: D[:, r_j] = e_{r_j}
                z[c] -= xin[c];
            }
            double* o = blk + 14 * (1 + c);
#pragma unroll
            for (int r = 0; r < 14; r += 2) *reinterpret_cast<double2*>(o + r) = make_double2(col[r], col[r + 1]);
        }
        double* zo = blk + 14 * 22;
#pragma unroll
        for (int r = 0; r < 7; ++r) zo[r] = z[r];
#pragma unroll
        for (int r = 7; r < 14; ++r) zo[r] = x[r];
    }
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
            *reinterpret_cast<double2*>(o + 2) = make_double2(-(u[2] / nu), ((SP == 2) ? __ldg(&bt.P[b].Tmin) : ldp<SP>(&P.Tmin)) - nu);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Kernel B: tangent propagation of the 14 heavy columns (+ 2 light ones in the idle slots).  Warp-specialised,
// persistent, one CTA of 12 warps per SM:
//   * warps 0..7  CONSUMERS: 8 lanes per interval, two full columns per lane (4 intervals per warp, 32 per pass); they
//     run nothing but the FP64 chains of consume_stage8 on register state;
//   * warps 8..11 PRODUCERS: warp 8+k forms the Jacobian blocks of rk4 stage k of every step with lane = interval, from
//     the stage record the value kernel wrote (one TMA bulk copy per record, mbarrier complete_tx; the next record is
//     pulled into L2 a step ahead because the staging buffer is single);
//   * registers are re-partitioned between the warpgroups with setmaxnreg (168 at launch -> 208 per consumer thread,
//     88 per producer thread): the consumers keep their 156 registers of tangent state without spills while a third
//     warp per scheduler fills the issue slots their FP64 dependency chains leave empty;
//   * hand-over once per rk4 STEP through a shared-memory ring two steps deep (2 x 4 stage slabs): consumers wait ONE
//     "full" barrier per step and release the four slabs with ONE "empty" arrive per step (a per-stage hand-over cost
//     17 % of the kernel in barrier waits and loop drain); producers run as far ahead as the ring allows and back off
//     between polls of a free ring slot (mbar_wait_relaxed: a polling producer took issue slots from the consumers);
//   * the four stages of a step are four instantiations of consume_stage8 (first / middle / middle / last), the step
//     loop is unrolled by two.
// (Until the producers got their own warps the consumer warps took turns producing, inlined at step boundaries where
//  only S is live; the consumers then spent 18 % of their issue slots on it: 9.8e7 -> 1.02e8 intervals/s.)
// ------------------------------------------------------------------------------------------------
constexpr int TANGENT_THREADS = (NWARP + 4) * 32;
// setmaxnreg re-partition: 128 * PRODUCER_REGS + 256 * CONSUMER_REGS <= 384 * 168 (the launch-time allocation)
constexpr int CONSUMER_REGS = 208;
constexpr int PRODUCER_REGS = (384 * 168 - 256 * CONSUMER_REGS) / 128;
static_assert(PRODUCER_REGS == 88, "register split");

struct __align__(16) StepSmem {
    double ring[8][GROUP][NJ];               // [half * 4 + stage][interval][entry]
    double recbuf[4][REC_MAX * GROUP];       // stage records of the step being produced (TMA destination)
    double unode[4][6][GROUP];               // per producer warp: node controls u-(3), u+(3) of the pass's 32 intervals
    uint64_t full_step[2];
    uint64_t empty_step[2];
    uint64_t recfull[2][4];                  // [half][stage]: one waiting warp per barrier (it observes every phase)
};

// Compile-time switches of the tangent kernel (A/B builds: profiles/build_variants.py; defaults = the measured best).
// SCVX_T_FIRST_BODY: 0 = one body for the first three stages of a step, 1 = own body for the first stage (see
// consume_stage8), 2 = and the two middle stages unrolled
#ifndef SCVX_T_FIRST_BODY
#define SCVX_T_FIRST_BODY 2
#endif
// SCVX_T_SPLIT_REDUCE: halving butterfly for the z partials of the epilogue (see there)
#ifndef SCVX_T_SPLIT_REDUCE
#define SCVX_T_SPLIT_REDUCE 1
#endif
// SCVX_T_PREFETCH_PASS: a producer warp pulls the next pass's sigma / control lines into L2
#ifndef SCVX_T_PREFETCH_PASS
#define SCVX_T_PREFETCH_PASS 1
#endif
// SCVX_T_STEP_UNROLL: 2 = the consumers' step loop unrolled by two
#ifndef SCVX_T_STEP_UNROLL
#define SCVX_T_STEP_UNROLL 2
#endif
// SCVX_RECORD_SLEEP_NS: back-off of the producers' wait for their TMA'd record (0 = spin; no measurable effect)
#ifndef SCVX_RECORD_SLEEP_NS
#define SCVX_RECORD_SLEEP_NS 0
#endif
template <int SP>
__global__ void __launch_bounds__(TANGENT_THREADS, 1) tangent_kernel(const __grid_constant__ StagedArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    StepSmem& sm = *reinterpret_cast<StepSmem*>(smem_raw);
    const ScvxBatch& bt = a.bt;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int l8 = lane & 7, sub = lane >> 3;
    const int ni = bt.n_nodes - 1;
    const int npts = bt.npts, nst = 4 * npts;
    const double h = bt.dt / (double)npts;
    const double pcs = 1.0 / (double)npts;
    const double sstep = (bt.mode == SCVX_MODE_LITERAL) ? 1.0 : h;
    const double h6 = h * (1.0 / 6.0);

    if (tid == 0) {
        for (int r = 0; r < 2; ++r) { mbar_init(&sm.full_step[r], 4 * 32); mbar_init(&sm.empty_step[r], NWARP); }
        for (int k = 0; k < 8; ++k) mbar_init(&sm.recfull[k >> 2][k & 3], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_wait();                                       // V_c is complete, its stage records and partial z are visible

    const int my_groups = (a.n_groups > (int)blockIdx.x) ? (a.n_groups - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int total_steps = my_groups * npts;          // global step counter n = pass * npts + local step

    // slot A: control-type columns (B-0..2, B+0..2, Sigma, light d/dm); slot B: state-type columns (w0..2, q0..3, light
    // d/dv0) — see consume_stage8
    const int colA = (l8 < 7) ? 14 + l8 : 0;
    const int colB = (l8 < 3) ? 11 + l8 : (l8 < 7 ? 4 + l8 : 4);
    const int gcol = (l8 < 3) ? l8 : (l8 < 6 ? l8 - 3 : 3);
    // FOH weight of slot A's direct term: alpha = cA0 + cA1 * pc  (B-: 1 - pc, B+: pc, Sigma: 1, d/dm: 0)
    const double cA0 = (l8 < 3 || l8 == 6) ? 1.0 : 0.0, cA1 = (l8 < 3) ? -1.0 : (l8 < 6 ? 1.0 : 0.0);
    const double dsA = (l8 == 6) ? 1.0 : 0.0;

    const int kq = warp & 3;                            // stage (within a step) this warp produces
    // rk4 factor folded into the Jacobian blocks of that stage (consume_stage8): Y_{i+1} = S + c_i K_i, h/6 for the last
    const double stage_scale = (kq == 3) ? h6 : (kq == 2 ? sstep : 0.5 * sstep);
    const double kappa = (bt.mode == SCVX_MODE_LITERAL) ? h * (1.0 / 3.0) : (1.0 / 3.0);
    const uint32_t rec_bytes = (uint32_t)a.rec_n * GROUP * 8;
    // stage record of (pass it, local step ls, stage kq); (it, ls) are tracked incrementally — no integer divisions in the
    // producer loop
    auto record_src = [&](int it, int ls) -> const double* {
        const int g = blockIdx.x + it * gridDim.x;
        return a.rec + ((size_t)g * nst + 4 * ls + kq) * ((size_t)a.rec_n * GROUP);
    };
    auto issue_record = [&](int n, int it, int ls) {    // one lane: TMA the record of global step n = it * npts + ls
        fence_proxy_async();
        mbar_expect_tx(&sm.recfull[n & 1][kq], rec_bytes);
        bulk_g2s(sm.recbuf[kq], record_src(it, ls), rec_bytes, &sm.recfull[n & 1][kq]);
    };

    // dedicated producer warps: warp NWARP + kq forms stage kq of every step, as far ahead as the ring allows
    if (warp >= NWARP) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PRODUCER_REGS));
        if (total_steps > 0) {
            if (lane == 0) issue_record(0, 0, 0);
            __syncwarp();
            int n = 0;
#pragma unroll 1
            for (int it = 0; it < my_groups; ++it) {
                // whole warp, lane = interval: this pass's interval, its trajectory and parameters (once per pass)
                const int g = blockIdx.x + it * gridDim.x;
                int t = g * GROUP + lane; if (t >= a.count) t = a.count - 1;
                const int b = (a.first + t) / ni;
                const scvx_probinfo& P = SP ? a.Pc : bt.P[bt.n_params == 1 ? 0 : b];
                const double pa = (SP == 2) ? __ldg(&bt.P[b].a) : ldp<SP>(&P.a);
                const double sigma = __ldg(bt.sigma + b);
#if SCVX_T_PREFETCH_PASS
                if (kq == 0 && it + 1 < my_groups) {
                    // the next pass starts with these loads on the consumers' critical path (they wait for its first
                    // records at the pass boundary): have the lines in L2 by then
                    int t2 = (g + (int)gridDim.x) * GROUP + lane; if (t2 >= a.count) t2 = a.count - 1;
                    const int b2 = (a.first + t2) / ni, i2 = (a.first + t2) - b2 * ni;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(bt.sigma + b2));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(bt.U + ((size_t)b2 * bt.n_nodes + i2) * 3));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(bt.U + ((size_t)b2 * bt.n_nodes + i2) * 3 + 5));
                }
#endif
                {   // the interval's node controls (the stage control is their FOH blend, re-formed per record)
                    const int ii = (a.first + t) - b * ni;
                    const double* uin = bt.U + ((size_t)b * bt.n_nodes + ii) * 3;
#pragma unroll
                    for (int c = 0; c < 6; ++c) sm.unode[kq][c][lane] = __ldg(uin + c);
                    __syncwarp();
                }
                double pca = 0.0;
#pragma unroll 1
                for (int ls = 0; ls < npts; ++ls, ++n) {
                    const bool more = n + 1 < total_steps;
                    const int nit = (ls + 1 < npts) ? it : it + 1, nls = (ls + 1 < npts) ? ls + 1 : 0;
                    // the record buffer is single (shared memory is full), so the TMA of step n+1 can only be issued once
                    // this step's record has been read: pull it into L2 now, the copy then costs an L2 hit instead of a
                    // DRAM round trip
                    if (lane == 0 && more) bulk_prefetch_l2(record_src(nit, nls), rec_bytes);
                    const int half = n & 1, use = n >> 1;
                    mbar_wait_relaxed<SCVX_RECORD_SLEEP_NS>(&sm.recfull[half][kq], (uint32_t)(use & 1));      // this warp is the only waiter of recfull[half][kq]
                    if (use > 0) mbar_wait_relaxed(&sm.empty_step[half], (uint32_t)((use - 1) & 1));
                    const double pc = (kq == 0) ? pca : (kq == 3 ? pca + pcs : pca + 0.5 * pcs);     // as the value kernel
                    pca += pcs;
                    produce_lean<SP>(P, pa, a.Kw, a.Tw, a.rec_n == REC_AERO, sigma, stage_scale, sm.recbuf[kq] + lane,
                                     &sm.unode[kq][0][lane], pc, &sm.ring[half * 4 + kq][lane][0]);
                    mbar_arrive(&sm.full_step[half]);
                    __syncwarp();                        // every lane has finished reading recbuf[kq]
                    if (lane == 0 && more) issue_record(n + 1, nit, nls);
                }
            }
        }
        return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(CONSUMER_REGS));

    FullCol FA, FB;
    int n = 0;                                           // global consumer step
    for (int it = 0; it < my_groups; ++it) {
        const int g = blockIdx.x + it * gridDim.x;
#pragma unroll
        for (int r = 0; r < 11; ++r) {
            // identity part of S(0) = [I | 0]; local rows 0 m, 1..3 v, 4..7 q, 8..10 w <-> inp columns 0, 4..6, 7..10, 11..13
            FA.S[r] = (colA == 0 && r == 0) ? 1.0 : 0.0;        // control columns start at zero; d/dm starts at e_m
            FB.S[r] = (colB == r + 3 && r >= 1) ? 1.0 : 0.0;    // state columns start at their unit vector
            FA.A[r] = 0.0; FB.A[r] = 0.0;
            FA.Y[r] = FA.S[r]; FB.Y[r] = FB.S[r];
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) FA.Sr[r] = FB.Sr[r] = 0.0;
        // this lane's interval and the two input values its columns multiply in z = endpoint - D * inp: loaded now,
        // so the epilogue has no dependent global loads
        const int t = g * GROUP + warp * 4 + sub;
        const bool live = t < a.count;
        const int wi = a.first + (live ? t : a.count - 1);
        const int bi = wi / ni, ii = wi - bi * ni;
        auto inp_of = [&](int c) -> double {
            if (c < 0) return 0.0;
            if (c < 14) return __ldg(bt.X + ((size_t)bi * bt.n_nodes + ii) * 14 + c);
            if (c < 20) return __ldg(bt.U + ((size_t)bi * bt.n_nodes + ii) * 3 + (c - 14));
            return __ldg(bt.sigma + bi);
        };
        const double xcA = inp_of(colA), xcB = inp_of(colB);

        double pca = 0.0;
#if SCVX_T_STEP_UNROLL > 1
#pragma unroll 2
#else
#pragma unroll 1
#endif
        for (int ls = 0; ls < npts; ++ls, ++n) {
            // ---- consume the four stages of step n
            const int half = n & 1;
            mbar_wait(&sm.full_step[half], (uint32_t)((n >> 1) & 1));      // (backing off here changes nothing: r2_ab_consumer_sleep.txt)
            const double* J0 = &sm.ring[half * 4][warp * 4 + sub][0];
#if SCVX_T_FIRST_BODY
            consume_stage8<0>(FA, FB, J0, gcol, fma(cA1, pca, cA0), dsA, 1.0, h6, kappa, nullptr, lane);
#if SCVX_T_FIRST_BODY == 2
#pragma unroll
#else
#pragma unroll 1
#endif
            for (int k = 1; k < 3; ++k)
                consume_stage8<1>(FA, FB, J0 + k * (GROUP * NJ), gcol, fma(cA1, pca + 0.5 * pcs, cA0), dsA, k == 1 ? 2.0 : 1.0,
                                  2.0 * h6, kappa, nullptr, lane);
            consume_stage8<2>(FA, FB, J0 + 3 * (GROUP * NJ), gcol, fma(cA1, pca + pcs, cA0), dsA, 1.0, h6, kappa, nullptr, lane);
#else
#pragma unroll 1
            for (int k = 0; k < 3; ++k) {
                const double pc = (k == 0) ? pca : pca + 0.5 * pcs;
                consume_stage8<1>(FA, FB, J0 + k * (GROUP * NJ), gcol, fma(cA1, pc, cA0), dsA, k == 1 ? 2.0 : 1.0,
                                  k == 0 ? h6 : 2.0 * h6, kappa, nullptr, lane);
            }
            consume_stage8<3>(FA, FB, J0 + 3 * (GROUP * NJ), gcol, fma(cA1, pca + pcs, cA0), dsA, 1.0, h6, kappa, nullptr, lane);
#endif
            __syncwarp();
            mbar_arrive_lane0(&sm.empty_step[half], lane);           // the four slabs of this step are free again
            pca += pcs;
        }

        // ---- epilogue: write D columns; z -= D[:, heavy] * inp with one fire-and-forget reduction per row
        double* blk = bt.out_blocks + (size_t)wi * SCVX_BLOCK_DOUBLES;
        double zp[14];
#pragma unroll
        for (int r = 0; r < 14; ++r) zp[r] = 0.0;
        auto emit_full = [&](const FullCol& F, int c, double xc) {
            if (c < 0) return;
            const double col[14] = { F.S[0], F.Sr[0], F.Sr[1], F.Sr[2], F.S[1], F.S[2], F.S[3], F.S[4], F.S[5], F.S[6], F.S[7],
                                     F.S[8], F.S[9], F.S[10] };
            double* o = blk + 14 * (1 + c);
#pragma unroll
            for (int r = 0; r < 14; r += 2) {
                if (live) *reinterpret_cast<double2*>(o + r) = make_double2(col[r], col[r + 1]);
                zp[r] = fma(col[r], xc, zp[r]); zp[r + 1] = fma(col[r + 1], xc, zp[r + 1]);
            }
        };
        emit_full(FA, colA, xcA);
        emit_full(FB, colB, xcB);
#if SCVX_T_SPLIT_REDUCE
        // Sum of the 14 partials over the 8 lanes of the interval, halving the rows a lane carries in every round (7 + 4 + 2
        // exchanged values instead of 3 x 14: the epilogue is bound by the shuffle queue).  The pairing is the butterfly
        // 1, 2, 4 of the plain reduction, so every sum is formed in the same order: bit-identical.  Lane l8 ends up with the
        // rows 7 b0 + 4 b1 + 2 b2 + {0, 1} (b = the bits of l8; rows past the seventh of a half do not exist).
        {
            const bool b0 = l8 & 1, b1 = l8 & 2, b2 = l8 & 4;
            double a7[7], a4[4], a2[2];
#pragma unroll
            for (int r = 0; r < 7; ++r) {
                const double send = b0 ? zp[r] : zp[7 + r], keep = b0 ? zp[7 + r] : zp[r];
                a7[r] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const double hi = (r < 3) ? a7[4 + r] : 0.0;
                const double send = b1 ? a7[r] : hi, keep = b1 ? hi : a7[r];
                a4[r] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const double send = b2 ? a4[r] : a4[2 + r], keep = b2 ? a4[2 + r] : a4[r];
                a2[r] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            }
            // the z column holds the partial z written by stage_value_kernel; exactly one addend per entry -> deterministic
            const int half_row = (b1 ? 4 : 0) + (b2 ? 2 : 0);
            double* o = blk + 14 * 22 + (b0 ? 7 : 0) + half_row;
            if (live && half_row < 7) atomicAdd(o, -a2[0]);
            if (live && half_row + 1 < 7) atomicAdd(o + 1, -a2[1]);
        }
#else
#pragma unroll
        for (int r = 0; r < 14; ++r) {
            double v = zp[r];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            zp[r] = v;
        }
        if (live && l8 == 7) {
            // the z column holds the partial z written by stage_value_kernel; exactly one addend per entry -> deterministic
            double* o = blk + 14 * 22;
#pragma unroll
            for (int r = 0; r < 14; ++r) atomicAdd(o + r, -zp[r]);
        }
#endif
    }
}

}  // namespace

size_t scvx_staged_scratch_bytes(int npts, int chunk_intervals) {
    const size_t groups = ((size_t)chunk_intervals + GROUP - 1) / GROUP;
    return groups * (size_t)(4 * npts) * REC_MAX * GROUP * sizeof(double);
}

// chunk = whole waves of both kernels: the value kernel keeps 2 x 128 threads per SM resident (255 registers), the
// tangent kernel 32 intervals per pass.  Every chunk costs two kernel boundaries (ramp, tail, launch gap), so chunks are
// long: 2304 intervals per SM = 9 waves / 72 passes, 3.5 GB of stage records at npts = 10 (768 per SM: -2.0 %, 1536:
// -0.9 %, 3072: +0.3 %; profiles/r2_ab_chunk.txt).  SCVX_CHUNK_PER_SM overrides (A/B).
int scvx_staged_chunk_intervals(int sm_count) {
    static const int per_sm = getenv("SCVX_CHUNK_PER_SM") ? atoi(getenv("SCVX_CHUNK_PER_SM")) : (SCVX_A_SMEM_TABLES == 2 ? 2688 : 2304);
    return sm_count * (per_sm > 0 ? (per_sm + 31) / 32 * 32 : 2304);
}

// the table-staging mode a launch uses: the compiled preference if the tables fit beside the light-column state, else 0
static int value_table_mode(const ScvxTables& tb, bool any_aero) {
    const int pref = SCVX_A_SMEM_TABLES;
    if (pref == 0 || !any_aero || !tb.drag || !tb.lift) return 0;
    if (pref == 3) return 3;
    const size_t need = value_smem_bytes(pref) + (size_t)pref * (tb.n1 + 2) * (tb.n2 + 2) * sizeof(double);
    return need <= 232448 ? pref : 0;
}

template <int TS>
static cudaError_t value_kernel_attributes() {
    const int vs = (TS == 1 || TS == 2) ? 232448 : (int)value_smem_bytes(TS);
    cudaError_t e = cudaFuncSetAttribute(stage_value_kernel<0, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, vs);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(stage_value_kernel<1, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, vs);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(stage_value_kernel<2, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, vs);
    return e;
}

cudaError_t scvx_staged_init() {
    cudaError_t e = value_kernel_attributes<0>();
    if (e == cudaSuccess && SCVX_A_SMEM_TABLES != 0) e = value_kernel_attributes<SCVX_A_SMEM_TABLES>();
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tangent_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StepSmem));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tangent_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StepSmem));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tangent_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StepSmem));
    return e;
}

// Launch with or without programmatic stream serialisation (see pdl_wait above).
template <typename K>
static cudaError_t launch_pdl(K kernel, unsigned grid, unsigned block, size_t smem, cudaStream_t s, bool programmatic,
                              const StagedArgs& a) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = programmatic ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, a);
}

template <int TS>
static cudaError_t launch_value(const StagedArgs& a, const ScvxTables& tb, int sp, cudaStream_t s, bool programmatic) {
    constexpr int VT = value_threads(TS);
    const unsigned grid = (unsigned)((a.n_groups * GROUP + VT - 1) / VT);
    size_t vsmem = value_smem_bytes(TS);
    if (TS == 1 || TS == 2) vsmem += (size_t)TS * (tb.n1 + 2) * (tb.n2 + 2) * sizeof(double);
    if (sp == 2) return launch_pdl(stage_value_kernel<2, TS>, grid, VT, vsmem, s, programmatic, a);
    if (sp == 1) return launch_pdl(stage_value_kernel<1, TS>, grid, VT, vsmem, s, programmatic, a);
    return launch_pdl(stage_value_kernel<0, TS>, grid, VT, vsmem, s, programmatic, a);
}

cudaError_t scvx_launch_staged(const ScvxBatch& bt, const ScvxTables& tb, bool any_aero, const scvx_probinfo* shared_params,
                               bool sweep, void* scratch, int chunk_intervals, int sm_count, cudaStream_t s, int* launches) {
    const long total = (long)(bt.n_nodes - 1) * bt.B;
    const size_t smem = sizeof(StepSmem);
    static const bool pdl = !(getenv("SCVX_PDL") && atoi(getenv("SCVX_PDL")) == 0);     // SCVX_PDL=0: plain launches (A/B)
    for (long first = 0; first < total; first += chunk_intervals) {
        StagedArgs a;
        a.bt = bt; a.tb = tb; a.rec = (double*)scratch; a.first = (int)first;
        const int ts = value_table_mode(tb, any_aero);
        if (shared_params) {
            a.Pc = *shared_params;
            const double* bi = a.Pc.jBi;
            const double r0 = a.Pc.rTB[0], r1 = a.Pc.rTB[1], r2 = a.Pc.rTB[2];
            a.Kw[0] = bi[3] * r2 - bi[6] * r1; a.Kw[1] = bi[4] * r2 - bi[7] * r1; a.Kw[2] = bi[5] * r2 - bi[8] * r1;
            a.Kw[3] = bi[6] * r0 - bi[0] * r2; a.Kw[4] = bi[7] * r0 - bi[1] * r2; a.Kw[5] = bi[8] * r0 - bi[2] * r2;
            a.Kw[6] = bi[0] * r1 - bi[3] * r0; a.Kw[7] = bi[1] * r1 - bi[4] * r0; a.Kw[8] = bi[2] * r1 - bi[5] * r0;
            const double* jB = a.Pc.jB;
            for (int k = 0; k < 3; ++k) {                  // Tw_k = -jBi ([e_k]x jB - [jB e_k]x), row-major 3 x 3
                double w[3] = { 0.0, 0.0, 0.0 }, M[3][3];
                w[k] = 1.0;
                const double L[3] = { jB[0] * w[0] + jB[3] * w[1] + jB[6] * w[2], jB[1] * w[0] + jB[4] * w[1] + jB[7] * w[2],
                                      jB[2] * w[0] + jB[5] * w[1] + jB[8] * w[2] };
                for (int c = 0; c < 3; ++c) {
                    const double a0 = jB[3 * c], a1 = jB[3 * c + 1], a2 = jB[3 * c + 2];
                    M[0][c] = w[1] * a2 - w[2] * a1; M[1][c] = w[2] * a0 - w[0] * a2; M[2][c] = w[0] * a1 - w[1] * a0;
                }
                M[0][1] += L[2]; M[0][2] -= L[1]; M[1][0] -= L[2]; M[1][2] += L[0]; M[2][0] += L[1]; M[2][1] -= L[0];
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c)
                        a.Tw[9 * k + 3 * r + c] = -(bi[r] * M[0][c] + bi[r + 3] * M[1][c] + bi[r + 6] * M[2][c]);
            }
        }
        a.rec_n = any_aero ? REC_AERO : REC_EXO;
        a.count = (int)((total - first < chunk_intervals) ? (total - first) : chunk_intervals);
        a.n_groups = (a.count + GROUP - 1) / GROUP;
        const int grid = a.n_groups < sm_count ? a.n_groups : sm_count;
        // 0: one record per trajectory in global memory; 1: one shared record in the kernel arguments;
        // 2: shared record + per-trajectory `a` / `Tmin` (the records of bt.P differ in those two fields only)
        const int sp = !shared_params ? 0 : (sweep ? 2 : 1);
        cudaError_t e;
        if (ts != 0) e = launch_value<SCVX_A_SMEM_TABLES>(a, tb, sp, s, false);
        else e = launch_value<0>(a, tb, sp, s, false);
        if (e != cudaSuccess) return e;
        if (sp == 2) e = launch_pdl(tangent_kernel<2>, grid, TANGENT_THREADS, smem, s, pdl, a);
        else if (sp == 1) e = launch_pdl(tangent_kernel<1>, grid, TANGENT_THREADS, smem, s, pdl, a);
        else e = launch_pdl(tangent_kernel<0>, grid, TANGENT_THREADS, smem, s, pdl, a);
        if (e != cudaSuccess) return e;
        if (launches) *launches += 2;
    }
    return cudaGetLastError();
}
