// scvx_kernels_fused.cu — the FUSED linearise-and-discretise path (sm_100a, FP64): the default.
//
// Same mathematics as the STAGED path (exact forward-mode tangent of the reference's rk4, dynamics.jl:112-134,
// 311-313), reorganised so that every piece of work runs in the shape in which it parallelises:
//
//  A1  value_record_kernel  : one THREAD per interval — the only inherently serial part: the 14-state value through
//      the 4*npts stages.  Writes endpoint, lin_err, thrust-lower-bound rows and a 25-double stage record per stage.
//
//  B2  tangent_fused_kernel : persistent, one 256-thread CTA per SM, 32 intervals per pass, two phases per pass.
//      phase 1  Jacobian records for ALL stages of the 32 intervals at once: one warp-task per stage with
//               lane = interval (no redundancy, no serial chain, 40 independent tasks over 8 warps): aero force
//               Jacobians (two spline value+gradient evaluations), rotational / quaternion / thrust blocks,
//               sigma-scaled, 78 doubles per (interval, stage), stored to a per-CTA slab of global scratch that stays
//               L2-resident (written and re-read by the same SM within one pass).
//      phase 2  tangent propagation: 8 lanes per interval, two full tangent columns per lane in registers.  The
//               stage slabs (32 x 78 doubles, contiguous) are streamed back into a shared-memory ring by TMA bulk
//               copies (cp.async.bulk + mbarrier complete_tx), issued a few stages ahead by one elected thread;
//               consumers read them as broadcast 128-bit loads and hand slots back through an "empty" mbarrier.
//               No thread computes anything but tangent FMAs in this phase.
//      [A|B-|B+|Sigma] columns go straight from registers to the 14x23 block; z is accumulated in place.
//
//  A2  light_columns_fused_kernel : one THREAD per interval — the four light columns d/d(m, v): only their v and r
//      rows are non-trivial; reads m, f_v from the stage record and dF/dv saved by phase 1.
#include "scvx_staged_dev.cuh"
#include "scvx_kernels.h"

namespace {

constexpr int FRING = 8;           // ring slots of phase 2 (8 x 19.5 KB)
constexpr int FLOOK = 5;           // TMA runs this many stages ahead of the consumers (< FRING)
constexpr int NFV = 9;             // dF/dv entries kept for the light columns

struct FusedArgs {
    ScvxBatch bt;
    ScvxTables tb;
    double* rec;                   // [n_groups][nst][25][32]  stage records (value kernel -> phase 1, A2)
    double* fvrec;                 // [n_groups][nst][9][32]   dF_aero/dv (phase 1 -> A2); unused when !aero
    double* jscr;                  // [gridDim][nst][32][NJ]   per-CTA Jacobian slabs (phase 1 -> phase 2)
    int aero;                      // any trajectory uses the aero tables
    int first, count, n_groups;
};

// ------------------------------------------------------------------------------------------------
// A1: value trajectory + stage records
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 3) value_record_kernel(FusedArgs a) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.n_groups * GROUP) return;
    const ScvxBatch& bt = a.bt;
    const int ni = bt.n_nodes - 1;
    const bool live = t < a.count;
    const int w = a.first + (live ? t : a.count - 1);          // padded lanes recompute the last interval
    const int b = w / ni, i = w - b * ni;
    const scvx_probinfo& P = bt.P[bt.n_params == 1 ? 0 : b];
    const double* xin = bt.X + ((size_t)b * bt.n_nodes + i) * 14;
    const double* uin = bt.U + ((size_t)b * bt.n_nodes + i) * 3;
    const double sigma = __ldg(bt.sigma + b);
    double x[14], um[3], up[3];
#pragma unroll
    for (int r = 0; r < 14; ++r) x[r] = xin[r];
#pragma unroll
    for (int c = 0; c < 3; ++c) { um[c] = uin[c]; up[c] = uin[3 + c]; }

    const int nst = 4 * bt.npts;
    double* rec = a.rec + ((size_t)(t >> 5) * nst) * (REC_EXO * GROUP) + (t & 31);
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
            rhs_value<false>(P, a.tb, y, uc, f, Fv, Fb);
            double* rp = rec + (size_t)(it * 4 + st) * (REC_EXO * GROUP);
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
    for (int r = 0; r < 14; r += 2) {
        *reinterpret_cast<double2*>(blk + r) = make_double2(x[r], x[r + 1]);                 // endpoint
        *reinterpret_cast<double2*>(blk + 14 * 22 + r) = make_double2(x[r], x[r + 1]);       // z starts as the endpoint
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
            *reinterpret_cast<double2*>(o + 2) = make_double2(-(u[2] / nu), __ldg(&P.Tmin) - nu);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// B2: fused Jacobian production (phase 1) + TMA-fed tangent propagation (phase 2)
// ------------------------------------------------------------------------------------------------
struct __align__(16) FusedSmem {
    double ring[FRING][GROUP][NJ];
    uint64_t full[FRING];
    uint64_t empty[FRING];
};

// phase-1 task: Jacobian record of one stage for 32 intervals (lane = interval)
__device__ __noinline__ void produce_fused(const scvx_probinfo& P, const ScvxTables& tb, bool aero, double sigma,
                                           const double* __restrict__ rec, double* __restrict__ fv_out,
                                           double* __restrict__ out) {
    produce_core(P, aero, sigma, rec, out,
                 [&](const double v[3], double c00, double c10, double c20, double Fv[3][3], double Fb[3][3]) {
                     const double bvec[3] = { c00, c10, c20 };
                     double F[3];
                     aero_force_jac(P, tb, bvec, v, F, Fv, Fb);
#pragma unroll
                     for (int r = 0; r < 3; ++r)
#pragma unroll
                         for (int c = 0; c < 3; ++c) fv_out[(3 * r + c) * GROUP] = Fv[r][c];
                 });
}

__global__ void __launch_bounds__(256, 1) tangent_fused_kernel(FusedArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FusedSmem& sm = *reinterpret_cast<FusedSmem*>(smem_raw);
    const ScvxBatch& bt = a.bt;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int l8 = lane & 7, sub = lane >> 3;
    const int ni = bt.n_nodes - 1;
    const int nst = 4 * bt.npts;
    const double h = bt.dt / (double)bt.npts;
    const double pcs = 1.0 / (double)bt.npts;
    const double sstep = (bt.mode == SCVX_MODE_LITERAL) ? 1.0 : h;
    const double h6 = h * (1.0 / 6.0);
    constexpr uint32_t SLAB_BYTES = GROUP * NJ * 8;

    if (tid == 0) {
        for (int r = 0; r < FRING; ++r) { mbar_init(&sm.full[r], 1); mbar_init(&sm.empty[r], NWARP); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    int colA = -1, colB = -1, gcol = 3;
    if (l8 < 3) { colA = 14 + l8; colB = 17 + l8; gcol = l8; }
    else if (l8 == 3) colA = 20;
    else if (l8 == 4) { colA = 11; colB = 12; }
    else if (l8 == 5) { colA = 13; colB = 7; }
    else if (l8 == 6) { colA = 8; colB = 9; }
    else colA = 10;

    double* jslab = a.jscr + (size_t)blockIdx.x * nst * (GROUP * NJ);
    // ring bookkeeping runs across passes: stage counter of the whole CTA life
    int c_slot = 0; uint32_t c_phase = 0;          // consumer side
    int p_slot = 0; uint32_t p_use = 0;            // TMA issue side (thread 0): slot and number of completed laps

    for (int g = blockIdx.x; g < a.n_groups; g += gridDim.x) {
        // ================= phase 1: Jacobian records of all stages of this group =================
        {
            int t = g * GROUP + lane; if (t >= a.count) t = a.count - 1;
            const int b = (a.first + t) / ni;
            const scvx_probinfo& P = bt.P[bt.n_params == 1 ? 0 : b];
            const double sigma = __ldg(bt.sigma + b);
            const double* rec0 = a.rec + ((size_t)g * nst) * (REC_EXO * GROUP) + lane;
            double* fv0 = a.fvrec + ((size_t)g * nst) * (NFV * GROUP) + lane;
            for (int s = warp; s < nst; s += NWARP)
                produce_fused(P, a.tb, a.aero != 0, sigma, rec0 + (size_t)s * (REC_EXO * GROUP),
                              fv0 + (size_t)s * (NFV * GROUP), jslab + ((size_t)s * GROUP + lane) * NJ);
        }
        // make the slabs visible to the async proxy (TMA reads them back) and to the whole CTA
        __threadfence();
        asm volatile("fence.proxy.async;" ::: "memory");
        __syncthreads();

        // ================= phase 2: tangent propagation =================
        FullCol FA, FB;
#pragma unroll
        for (int r = 0; r < 11; ++r) {
            FA.S[r] = (r >= 4 && colA == r + 3) ? 1.0 : 0.0;       // local rows: 0 m, 1..3 v, 4..7 q, 8..10 w
            FB.S[r] = (r >= 4 && colB == r + 3) ? 1.0 : 0.0;
            FA.A[r] = 0.0; FB.A[r] = 0.0;
            FA.Y[r] = FA.S[r]; FB.Y[r] = FB.S[r];
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) FA.Sr[r] = FB.Sr[r] = 0.0;

        auto issue = [&](int s) {                    // thread 0: TMA slab of stage s into the next ring slot
            if (p_use > 0) mbar_wait(&sm.empty[p_slot], (p_use - 1) & 1);
            mbar_expect_tx(&sm.full[p_slot], SLAB_BYTES);
            bulk_g2s(&sm.ring[p_slot][0][0], jslab + (size_t)s * (GROUP * NJ), SLAB_BYTES, &sm.full[p_slot]);
            if (++p_slot == FRING) { p_slot = 0; ++p_use; }
        };
        if (tid == 0)
            for (int s = 0; s < FLOOK && s < nst; ++s) issue(s);

        double pca = 0.0;
#pragma unroll 1
        for (int s = 0; s < nst; ++s) {
            if (tid == 0 && s + FLOOK < nst) issue(s + FLOOK);
            const int st = s & 3;
            const int slot = c_slot;
            mbar_wait(&sm.full[slot], c_phase);
            if (++c_slot == FRING) { c_slot = 0; c_phase ^= 1; }
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

        // ---- epilogue: D columns of this lane, z -= D[:, heavy] * inp
        const int t = g * GROUP + warp * 4 + sub;
        const bool live = t < a.count;
        const int wi = a.first + (live ? t : a.count - 1);
        const int b = wi / ni, i = wi - b * ni;
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
                const double2 e = *reinterpret_cast<const double2*>(o + r);        // endpoint (value kernel)
                *reinterpret_cast<double2*>(o + r) = make_double2(e.x - zp[r], e.y - zp[r + 1]);
            }
        }
        // every warp must be done with the ring / slab of this group before phase 1 of the next group overwrites
        // the slab (all TMA copies of this group have completed: every stage was consumed)
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// A2: light tangent columns d/d(m, v0, v1, v2) + position columns; z -= D[:, m r v] * inp[m r v]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 3) light_columns_fused_kernel(FusedArgs a) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.count) return;
    const ScvxBatch& bt = a.bt;
    const int ni = bt.n_nodes - 1;
    const int w = a.first + t;
    const int b = w / ni, i = w - b * ni;
    const scvx_probinfo& P = bt.P[bt.n_params == 1 ? 0 : b];
    const double sigma = __ldg(bt.sigma + b), g0 = __ldg(&P.g0);
    const bool aero = a.aero != 0;
    const int nst = 4 * bt.npts;
    const double* rec = a.rec + ((size_t)(t >> 5) * nst) * (REC_EXO * GROUP) + (t & 31);
    const double* fvr = a.fvrec + ((size_t)(t >> 5) * nst) * (NFV * GROUP) + (t & 31);
    const double h = bt.dt / (double)bt.npts;
    const double sstep = (bt.mode == SCVX_MODE_LITERAL) ? 1.0 : h;
    const double h6 = h * (1.0 / 6.0);
    double S[4][3], A[4][3], Y[4][3], Sr[4][3];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r) { S[c][r] = (c == r + 1) ? 1.0 : 0.0; Y[c][r] = S[c][r]; A[c][r] = 0.0; Sr[c][r] = 0.0; }
    // software-pipelined record reads: the 13 values of stage s+1 are in flight while stage s is computed
    double nx[13];
    auto fetch = [&](int s) {
        const double* rp = rec + (size_t)s * (REC_EXO * GROUP);
        const double* fp = fvr + (size_t)s * (NFV * GROUP);
        nx[0] = __ldg(rp);
#pragma unroll
        for (int r = 0; r < 3; ++r) nx[1 + r] = __ldg(rp + (15 + r) * GROUP);
#pragma unroll
        for (int k = 0; k < 9; ++k) nx[4 + k] = aero ? __ldg(fp + k * GROUP) : 0.0;
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
#pragma unroll
    for (int c = 0; c < 7; ++c) {           // columns: inp 0 (m), 1..3 (r), 4..6 (v)
        double col[14];
#pragma unroll
        for (int r = 0; r < 14; ++r) col[r] = 0.0;
        if (c == 0) { col[0] = 1.0; for (int r = 0; r < 3; ++r) { col[1 + r] = Sr[0][r]; col[4 + r] = S[0][r]; } }
        else if (c < 4) col[c] = 1.0;       // nothing depends on position: D[:, r_j] = e_{r_j}
        else { for (int r = 0; r < 3; ++r) { col[1 + r] = Sr[c - 3][r]; col[4 + r] = S[c - 3][r]; } }
        double* o = blk + 14 * (1 + c);
#pragma unroll
        for (int r = 0; r < 14; r += 2) *reinterpret_cast<double2*>(o + r) = make_double2(col[r], col[r + 1]);
    }
    double* zo = blk + 14 * 22;             // z already holds endpoint - D[:, heavy] * inp (value + tangent kernels)
    double z[7];
#pragma unroll
    for (int r = 0; r < 7; ++r) z[r] = zo[r];
    z[0] -= xin[0];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        z[1 + r] -= Sr[0][r] * xin[0] + xin[1 + r] + Sr[1][r] * xin[4] + Sr[2][r] * xin[5] + Sr[3][r] * xin[6];
        z[4 + r] -= S[0][r] * xin[0] + S[1][r] * xin[4] + S[2][r] * xin[5] + S[3][r] * xin[6];
    }
#pragma unroll
    for (int r = 0; r < 7; ++r) zo[r] = z[r];
}

}  // namespace

// scratch: stage records + dF/dv records for `chunk_intervals`, Jacobian slabs for `sm_count` CTAs
size_t scvx_fused_scratch_bytes(int npts, int chunk_intervals, int sm_count) {
    const size_t groups = ((size_t)chunk_intervals + GROUP - 1) / GROUP;
    const size_t nst = (size_t)4 * npts;
    return (groups * nst * (REC_EXO + NFV) * GROUP + (size_t)sm_count * nst * GROUP * NJ) * sizeof(double);
}

cudaError_t scvx_launch_fused(const ScvxBatch& bt, const ScvxTables& tb, bool any_aero, void* scratch,
                              int chunk_intervals, int sm_count, cudaStream_t s, int* launches) {
    const long total = (long)(bt.n_nodes - 1) * bt.B;
    const size_t smem = sizeof(FusedSmem);
    {
        cudaError_t e = cudaFuncSetAttribute(tangent_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const size_t groups_cap = ((size_t)chunk_intervals + GROUP - 1) / GROUP;
    const size_t nst = (size_t)4 * bt.npts;
    double* rec = (double*)scratch;
    double* fvrec = rec + groups_cap * nst * REC_EXO * GROUP;
    double* jscr = fvrec + groups_cap * nst * NFV * GROUP;
    for (long first = 0; first < total; first += chunk_intervals) {
        FusedArgs a;
        a.bt = bt; a.tb = tb; a.rec = rec; a.fvrec = fvrec; a.jscr = jscr; a.aero = any_aero ? 1 : 0;
        a.first = (int)first;
        a.count = (int)((total - first < chunk_intervals) ? (total - first) : chunk_intervals);
        a.n_groups = (a.count + GROUP - 1) / GROUP;
        const int threads = a.n_groups * GROUP;
        value_record_kernel<<<(threads + 127) / 128, 128, 0, s>>>(a);
        const int grid = a.n_groups < sm_count ? a.n_groups : sm_count;
        tangent_fused_kernel<<<grid, 256, smem, s>>>(a);
        light_columns_fused_kernel<<<(a.count + 127) / 128, 128, 0, s>>>(a);
        if (launches) *launches += 3;
    }
    return cudaGetLastError();
}
