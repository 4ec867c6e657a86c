// scvx_api.cu — context management and the C ABI declared in include/scvx_b200.h.
//
// A context owns, per device: the staged spline tables, the parameter records, NSLOT pipeline slots of
// staging buffers + streams (host-pointer calls are cut into trajectory chunks so that the H2D copy of
// chunk c+1, the kernels of chunk c and the D2H copy of chunk c-1 overlap), and launch bookkeeping.
// Trajectories are independent (reference dynamics.jl:324-332 reads only nodes i, i+1), so multi-device
// contexts shard them in contiguous blocks with no exchange step.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <algorithm>
#include <atomic>
#include <thread>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "scvx_kernels.h"
#include "scvx_socp_pattern.h"
#include "scvx_compact.h"

static_assert(sizeof(scvx_probinfo) == 248 && sizeof(scvx_dim_problem) == 216, "ABI record sizes are part of the contract");

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    g_err = buf;
    return code;
}

#define CK(expr)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (expr);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(SCVX_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// Pipeline depth of the host-pointer path.  Three slots: when the D2H copy of chunk c ends, the copy of chunk c+1 is running
// and chunk c+2 has already been computed, so the D2H engine never waits for a kernel (with two slots it idled for
// the ~1.3 ms of H2D + kernels of every chunk: 43 instead of 54 GB/s on this pool's PCIe).
constexpr int NSLOT = 3;

struct Slot {
    cudaStream_t stream = nullptr;
    double *dX = nullptr, *dU = nullptr, *dS = nullptr, *dOut = nullptr, *dErr = nullptr, *dTlb = nullptr, *dEnd = nullptr,
           *dCmp = nullptr;
    size_t capX = 0, capU = 0, capS = 0, capOut = 0, capErr = 0, capTlb = 0, capEnd = 0, capCmp = 0;
};

// NVTX range over a scope (host timeline of the enqueue; the kernels and copies it covers show up under it in a trace)
struct Range {
    explicit Range(const char* name) { nvtxRangePushA(name); }
    ~Range() { nvtxRangePop(); }
};

// device temporaries that must not leak on an early return
struct DevTmp {
    void* p = nullptr;
    ~DevTmp() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes); }
    double* d() const { return (double*)p; }
};

struct Dev {
    int id = 0;
    Slot slot[NSLOT];
    double* coef[3] = { nullptr, nullptr, nullptr };
    int n1 = 0, n2 = 0;
    double x0 = 0, dx = 1, y0 = 0, dy = 1;
    scvx_probinfo* dP = nullptr;
    int nP = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    void* scratch = nullptr;           // STAGED kernel scratch (stage records)
    size_t scratch_cap = 0;
    int sm_count = 0;
    cudaStream_t scratch_user = nullptr;   // stream that last used the scratch
    cudaEvent_t ev_scratch = nullptr;
    double* dense = nullptr;               // dense blocks behind a device-pointer compact call
    size_t dense_cap = 0;
    double* fin[2] = { nullptr, nullptr }; // fin-force tables (lift, drag), SURVEY.md §8f-4
    int fn1 = 0, fn2 = 0;
    double fx0 = 0, fdx = 1, fy0 = 0, fdy = 1;
};

}  // namespace

struct scvx_ctx {
    std::vector<Dev> devs;
    cudaStream_t user_stream = nullptr;
    bool have_user_stream = false;
    int kernel = SCVX_KERNEL_AUTO;
    int64_t launches = 0;
    bool timed = false;
    std::vector<scvx_probinfo> hP;
    bool any_aero = false;
    bool hP_values = false;      // hP holds the records' values (set by scvx_set_params), not just their count / aero kind
    bool hP_sweep = false;       // > 1 records that differ in `a` and `Tmin` only (a mass / thrust-bound sweep)
};

namespace {

int grow(double** p, size_t* cap, size_t need_doubles) {
    if (need_doubles <= *cap) return 0;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    cudaError_t e = cudaMalloc((void**)p, need_doubles * sizeof(double));
    if (e != cudaSuccess) return fail(SCVX_ERR_NOMEM, "cudaMalloc(%zu B) failed: %s", need_doubles * sizeof(double), cudaGetErrorString(e));
    *cap = need_doubles;
    return 0;
}

bool is_device_ptr(const void* p) {
    if (!p) return false;
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// index (into c->devs) of the device a device pointer lives on, or -1
int dev_index_of(const scvx_ctx* c, const void* p);

ScvxTables tables_of(const Dev& d) {
    ScvxTables t;
    t.drag = d.coef[SCVX_TABLE_DRAG]; t.lift = d.coef[SCVX_TABLE_LIFT]; t.trq = d.coef[SCVX_TABLE_TORQUE];
    t.wdrag = nullptr; t.wlift = nullptr; t.wi0 = 0; t.wj0 = 0;
    t.n1 = d.n1; t.n2 = d.n2; t.x0 = d.x0; t.inv_dx = 1.0 / d.dx; t.y0 = d.y0; t.inv_dy = 1.0 / d.dy;
    return t;
}

int check_ready(scvx_ctx* c, int B) {
    if (c->hP.empty()) return fail(SCVX_ERR_STATE, "scvx_set_params has not been called");
    if ((int)c->hP.size() != 1 && (int)c->hP.size() != B)
        return fail(SCVX_ERR_ARG, "parameter count %d must be 1 or B=%d", (int)c->hP.size(), B);
    bool need_tables = false;
    for (const auto& p : c->hP) if (p.aero_kind == SCVX_AERO_TABLE) { need_tables = true; break; }
    if (need_tables) {
        const Dev& d = c->devs[0];
        if (!d.coef[SCVX_TABLE_DRAG] || !d.coef[SCVX_TABLE_LIFT])
            return fail(SCVX_ERR_STATE, "aero_kind=TABLE needs the drag and lift tables (scvx_set_aero_table)");
    }
    return 0;
}

// Launch the linearisation kernels for `bt` on device `d`, stream `s`.
int launch_linearize(scvx_ctx* c, Dev& d, const ScvxBatch& bt, cudaStream_t s) {
    const ScvxTables tb = tables_of(d);
    int k = c->kernel;
    if (k == SCVX_KERNEL_AUTO) k = SCVX_KERNEL_STAGED;
    if (k == SCVX_KERNEL_DUALWARP) {
        CK(scvx_launch_dualwarp(bt, tb, s));
        c->launches += 1;
        return 0;
    }
    // STAGED: the stage-record scratch is shared by everything enqueued on this device, so work on the other
    // pipeline stream must have consumed it before it is overwritten.
    const long total = (long)(bt.n_nodes - 1) * bt.B;
    int chunk = scvx_staged_chunk_intervals(d.sm_count);
    // keep the stage-record scratch below ~4 GiB for long integrations (npts >> 10): whole tangent passes per SM
    while (chunk > d.sm_count * 32 && scvx_staged_scratch_bytes(bt.npts, chunk) > ((size_t)4 << 30)) chunk -= d.sm_count * 32;
    if (total < chunk) chunk = (int)((total + 31) / 32 * 32);
    else {
        // equal chunks instead of full ones and a remainder: whole waves of the value kernel (2 x 128 threads per SM)
        const long n = (total + chunk - 1) / chunk, wave = (long)d.sm_count * 256;
        const long per = ((total + n - 1) / n + wave - 1) / wave * wave;
        if (per < chunk) chunk = (int)per;
    }
    const size_t need = scvx_staged_scratch_bytes(bt.npts, chunk);
    if (need > d.scratch_cap) {
        CK(cudaDeviceSynchronize());
        if (d.scratch) cudaFree(d.scratch);
        d.scratch = nullptr; d.scratch_cap = 0;
        cudaError_t e = cudaMalloc(&d.scratch, need);
        if (e != cudaSuccess) return fail(SCVX_ERR_NOMEM, "cudaMalloc(%zu B) for the stage-record scratch failed: %s", need, cudaGetErrorString(e));
        d.scratch_cap = need;
    }
    // ev_scratch marks the end of the KERNELS of the previous user (recorded right behind them, below): waiting on an
    // event recorded now on that stream would also wait for its D2H copy and serialise copy and compute
    if (d.scratch_user && d.scratch_user != s) CK(cudaStreamWaitEvent(s, d.ev_scratch, 0));
    int n = 0;
    // one record for all (n_params == 1), or a sweep whose records differ in a / Tmin only (the kernels then read those
    // two per trajectory and take the rest from record 0): the record travels in the kernel arguments
    const scvx_probinfo* shared = (c->hP_values && (c->hP.size() == 1 ? bt.n_params == 1 : c->hP_sweep)) ? c->hP.data() : nullptr;
    CK(scvx_launch_staged(bt, tb, c->any_aero, shared, c->hP.size() > 1, d.scratch, chunk, d.sm_count, s, &n));
    c->launches += n;
    CK(cudaEventRecord(d.ev_scratch, s));
    d.scratch_user = s;
    return 0;
}

int launch_predict(scvx_ctx* c, Dev& d, const ScvxBatch& bt, cudaStream_t s) {
    CK(scvx_launch_predict(bt, tables_of(d), s));
    c->launches += 1;
    return 0;
}

// Drain every pipeline stream of the context (after a failure: async copies into caller-owned buffers must not outlive
// the call).  Errors are ignored — the caller already reports one.
void drain(scvx_ctx* c) {
    for (Dev& d : c->devs) {
        cudaSetDevice(d.id);
        for (int q = 0; q < NSLOT; ++q) cudaStreamSynchronize(d.slot[q].stream);
    }
    cudaGetLastError();
}

int run_impl(scvx_ctx* c, bool predict, const double* X, const double* U, const double* sigma, double dt, int npts,
             int mode, int n_nodes, int B, double* out_blocks, double* out_lin_err, double* out_tlb, double* out_end,
             double* out_compact, int rec);

// Run one call. predict=false: linearize. Handles host pointers (chunked, all devices) and device pointers.
// out_compact != nullptr: compact records instead of dense blocks (out_blocks / out_lin_err must be null then).
int run(scvx_ctx* c, bool predict, const double* X, const double* U, const double* sigma, double dt, int npts,
        int mode, int n_nodes, int B, double* out_blocks, double* out_lin_err, double* out_tlb, double* out_end,
        double* out_compact = nullptr, int rec = SCVX_COMPACT_DOUBLES) {
    const int rc = run_impl(c, predict, X, U, sigma, dt, npts, mode, n_nodes, B, out_blocks, out_lin_err, out_tlb, out_end,
                            out_compact, rec);
    if (rc != 0 && c) drain(c);
    return rc;
}

int run_impl(scvx_ctx* c, bool predict, const double* X, const double* U, const double* sigma, double dt, int npts,
             int mode, int n_nodes, int B, double* out_blocks, double* out_lin_err, double* out_tlb, double* out_end,
             double* out_compact, int rec) {
    if (!c) return fail(SCVX_ERR_ARG, "null context");
    if (!X || !U || !sigma) return fail(SCVX_ERR_ARG, "X, U and sigma must be non-null");
    double* const out_main = predict ? out_end : (out_compact ? out_compact : out_blocks);
    if (!out_main) return fail(SCVX_ERR_ARG, "output pointer is null");
    if (n_nodes < 2) return fail(SCVX_ERR_ARG, "n_nodes=%d: at least two nodes (one interval) are required", n_nodes);
    if (B < 0) return fail(SCVX_ERR_ARG, "B=%d is negative", B);
    if (npts < 1) return fail(SCVX_ERR_ARG, "npts=%d must be >= 1", npts);
    if (mode != SCVX_MODE_LITERAL && mode != SCVX_MODE_TEXTBOOK) return fail(SCVX_ERR_ARG, "unknown mode %d", mode);
    if (!(dt > 0.0)) return fail(SCVX_ERR_ARG, "base_dt must be positive");
    if (B == 0) return 0;
    if ((long)(n_nodes - 1) * (long)B >= (1L << 31) - 64)
        return fail(SCVX_ERR_ARG, "(n_nodes-1)*B = %ld intervals exceed the 2^31 limit of one call", (long)(n_nodes - 1) * (long)B);
    if (int rc = check_ready(c, B)) return rc;
    const int ni = n_nodes - 1;
    const int nP = (int)c->hP.size();

    const bool dev_in = is_device_ptr(X);
    if (dev_in != is_device_ptr(U) || dev_in != is_device_ptr(sigma) || dev_in != is_device_ptr(out_main) ||
        (out_lin_err && dev_in != is_device_ptr(out_lin_err)) || (out_tlb && dev_in != is_device_ptr(out_tlb)))
        return fail(SCVX_ERR_ARG, "all array arguments must be either host or device pointers, not a mix");

    if (dev_in) {
        const int di = dev_index_of(c, X);
        if (di < 0 || di != dev_index_of(c, out_main))
            return fail(SCVX_ERR_ARG, "device pointers must live on one device of this context");
        Dev& d = c->devs[di];
        CK(cudaSetDevice(d.id));
        cudaStream_t s = c->have_user_stream ? c->user_stream : d.slot[0].stream;
        ScvxBatch bt;
        bt.X = X; bt.U = U; bt.sigma = sigma; bt.P = d.dP; bt.n_params = nP; bt.n_nodes = n_nodes; bt.B = B;
        bt.dt = dt; bt.npts = npts; bt.mode = mode;
        bt.out_blocks = out_blocks; bt.out_lin_err = out_lin_err; bt.out_tlb = out_tlb; bt.out_endpoints = out_end;
        if (out_compact) {
            // dense blocks go to a context-owned buffer, the pack kernel gathers the data entries behind them
            const size_t need = (size_t)ni * B * SCVX_BLOCK_DOUBLES;
            if (need > d.dense_cap) { CK(cudaDeviceSynchronize()); }
            if (grow(&d.dense, &d.dense_cap, need)) return SCVX_ERR_NOMEM;
            bt.out_blocks = d.dense;
        }
        Range r(predict ? "scvx:predict(device)" : "scvx:linearize(device)");
        CK(cudaEventRecord(d.ev0, s));
        if (int rc = predict ? launch_predict(c, d, bt, s) : launch_linearize(c, d, bt, s)) return rc;
        if (out_compact) {
            CK(scvx_launch_compact_pack(d.dense, (long)ni * B, rec, out_compact, d.sm_count, s));
            c->launches += 1;
        }
        CK(cudaEventRecord(d.ev1, s));
        c->timed = true;
        return 0;
    }

    // Host pointers: contiguous trajectory blocks per device, chunked + double-buffered inside a device.
    const int nd = (int)c->devs.size();
    const size_t per_traj_out = predict ? (size_t)ni * 14 : (size_t)ni * (out_compact ? rec : SCVX_BLOCK_DOUBLES);
    Range whole(predict ? "scvx:predict(host)" : (out_compact ? "scvx:linearize_compact(host)" : "scvx:linearize(host)"));
    static const long chunk_mb = getenv("SCVX_HOST_CHUNK_MB") ? atol(getenv("SCVX_HOST_CHUNK_MB")) : 256;
    long chunk = (long)(((size_t)chunk_mb << 20) / (per_traj_out * sizeof(double)));   // output per pipeline chunk (D2H saturates this pool's PCIe at ~43 GB/s from 128 MiB up, profiles/e2e_sweep.py)
    chunk = std::max(1L, chunk);
    // chunk index outermost, devices innermost: every device gets its c-th chunk enqueued before any device gets its
    // (c+1)-th, so the slot-reuse waits of one device never hold back the others (all devices run concurrently)
    // the D2H engine is the bottleneck and sits idle until the first chunk has been computed: a short first chunk (1/8)
    // starts it early (the pipeline prologue was 3-4 % of a 12-chunk call)
    const long first = std::max(1L, chunk / 8);
    long max_chunks = 0;
    for (int di = 0; di < nd; ++di) {
        const long len = (long)B * (di + 1) / nd - (long)B * di / nd;
        max_chunks = std::max(max_chunks, len <= chunk ? (len > 0 ? 1L : 0L) : 1 + (len - first + chunk - 1) / chunk);
    }
    for (long ci = 0; ci < max_chunks; ++ci) {
        for (int di = 0; di < nd; ++di) {
            Dev& d = c->devs[di];
            const long b0 = (long)B * di / nd, b1 = (long)B * (di + 1) / nd;
            const bool split = (b1 - b0) > chunk;
            const long cb = (ci == 0 || !split) ? b0 + ci * chunk : b0 + first + (ci - 1) * chunk;
            if (cb >= b1) continue;
            CK(cudaSetDevice(d.id));
            const int nb = (int)std::min((ci == 0 && split) ? first : chunk, b1 - cb);
            const size_t cap_nb = (size_t)std::min(chunk, b1 - b0);      // every slot is sized for a full chunk at once
            Slot& sl = d.slot[ci % NSLOT];
            CK(cudaStreamSynchronize(sl.stream));       // previous user of this slot has drained its D2H
            if (grow(&sl.dX, &sl.capX, cap_nb * n_nodes * 14) || grow(&sl.dU, &sl.capU, cap_nb * n_nodes * 3) ||
                grow(&sl.dS, &sl.capS, cap_nb))
                return SCVX_ERR_NOMEM;
            Range chunk_range("scvx:chunk h2d+kernels+d2h");
            CK(cudaMemcpyAsync(sl.dX, X + (size_t)cb * n_nodes * 14, (size_t)nb * n_nodes * 14 * 8, cudaMemcpyHostToDevice, sl.stream));
            CK(cudaMemcpyAsync(sl.dU, U + (size_t)cb * n_nodes * 3, (size_t)nb * n_nodes * 3 * 8, cudaMemcpyHostToDevice, sl.stream));
            CK(cudaMemcpyAsync(sl.dS, sigma + cb, (size_t)nb * 8, cudaMemcpyHostToDevice, sl.stream));
            ScvxBatch bt;
            bt.X = sl.dX; bt.U = sl.dU; bt.sigma = sl.dS; bt.P = (nP == 1) ? d.dP : d.dP + cb; bt.n_params = (nP == 1) ? 1 : nb;
            bt.n_nodes = n_nodes; bt.B = nb; bt.dt = dt; bt.npts = npts; bt.mode = mode;
            bt.out_blocks = nullptr; bt.out_lin_err = nullptr; bt.out_tlb = nullptr; bt.out_endpoints = nullptr;
            if (predict) {
                if (grow(&sl.dEnd, &sl.capEnd, cap_nb * ni * 14)) return SCVX_ERR_NOMEM;
                bt.out_endpoints = sl.dEnd;
                if (int rc = launch_predict(c, d, bt, sl.stream)) return rc;
                CK(cudaMemcpyAsync(out_end + (size_t)cb * ni * 14, sl.dEnd, (size_t)nb * ni * 14 * 8, cudaMemcpyDeviceToHost, sl.stream));
            } else {
                if (grow(&sl.dOut, &sl.capOut, cap_nb * ni * SCVX_BLOCK_DOUBLES)) return SCVX_ERR_NOMEM;
                bt.out_blocks = sl.dOut;
                if (out_lin_err) { if (grow(&sl.dErr, &sl.capErr, cap_nb * ni * 14)) return SCVX_ERR_NOMEM; bt.out_lin_err = sl.dErr; }
                if (out_tlb) { if (grow(&sl.dTlb, &sl.capTlb, cap_nb * n_nodes * 4)) return SCVX_ERR_NOMEM; bt.out_tlb = sl.dTlb; }
                if (int rc = launch_linearize(c, d, bt, sl.stream)) return rc;
                if (out_compact) {
                    if (grow(&sl.dCmp, &sl.capCmp, cap_nb * ni * rec)) return SCVX_ERR_NOMEM;
                    CK(scvx_launch_compact_pack(sl.dOut, (long)nb * ni, rec, sl.dCmp, d.sm_count, sl.stream));
                    c->launches += 1;
                    CK(cudaMemcpyAsync(out_compact + (size_t)cb * ni * rec, sl.dCmp, (size_t)nb * ni * rec * 8,
                                       cudaMemcpyDeviceToHost, sl.stream));
                } else {
                    CK(cudaMemcpyAsync(out_blocks + (size_t)cb * ni * SCVX_BLOCK_DOUBLES, sl.dOut,
                                       (size_t)nb * ni * SCVX_BLOCK_DOUBLES * 8, cudaMemcpyDeviceToHost, sl.stream));
                }
                if (out_lin_err)
                    CK(cudaMemcpyAsync(out_lin_err + (size_t)cb * ni * 14, sl.dErr, (size_t)nb * ni * 14 * 8, cudaMemcpyDeviceToHost, sl.stream));
                if (out_tlb)
                    CK(cudaMemcpyAsync(out_tlb + (size_t)cb * n_nodes * 4, sl.dTlb, (size_t)nb * n_nodes * 4 * 8, cudaMemcpyDeviceToHost, sl.stream));
            }
        }
    }
    for (int di = 0; di < nd; ++di) {
        Dev& d = c->devs[di];
        CK(cudaSetDevice(d.id));
        for (int q = 0; q < NSLOT; ++q) CK(cudaStreamSynchronize(d.slot[q].stream));
    }
    c->timed = false;
    return 0;
}

}  // namespace

namespace {
int dev_index_of(const scvx_ctx* c, const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return -1; }
    for (size_t i = 0; i < c->devs.size(); ++i) if (c->devs[i].id == a.device) return (int)i;
    return -1;
}

// the stream device-pointer work of entry points other than linearize / predict runs on, and its device
int device_call_target(scvx_ctx* c, const void* p, Dev** d, cudaStream_t* s) {
    const int di = dev_index_of(c, p);
    if (di < 0) return fail(SCVX_ERR_ARG, "device pointers must live on one device of this context");
    *d = &c->devs[di];
    CK(cudaSetDevice((*d)->id));
    *s = c->have_user_stream ? c->user_stream : (*d)->slot[0].stream;
    return 0;
}
}  // namespace

extern "C" {

int scvx_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* scvx_last_error(void) { return g_err.c_str(); }
int scvx_version(void) { return 1000; }
int scvx_sizeof_probinfo(void) { return (int)sizeof(scvx_probinfo); }

int scvx_create(scvx_ctx** out, const int* device_ids, int n_dev) {
    if (!out) return fail(SCVX_ERR_ARG, "out is null");
    *out = nullptr;
    int have = scvx_device_count();
    if (have <= 0) return fail(SCVX_ERR_CUDA, "no CUDA device is visible: this library has no CPU fallback");
    std::vector<int> ids;
    if (!device_ids || n_dev <= 0) ids.push_back(0);
    else ids.assign(device_ids, device_ids + n_dev);
    for (int id : ids) if (id < 0 || id >= have) return fail(SCVX_ERR_ARG, "device id %d out of range (have %d)", id, have);
    scvx_ctx* c = new (std::nothrow) scvx_ctx();
    if (!c) return fail(SCVX_ERR_NOMEM, "out of host memory");
    c->devs.resize(ids.size());
    for (size_t i = 0; i < ids.size(); ++i) {
        Dev& d = c->devs[i];
        d.id = ids[i];
        cudaError_t e = cudaSetDevice(d.id);
        for (int s = 0; s < NSLOT && e == cudaSuccess; ++s) e = cudaStreamCreateWithFlags(&d.slot[s].stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreate(&d.ev0);
        if (e == cudaSuccess) e = cudaEventCreate(&d.ev1);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d.ev_scratch, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, d.id);
        if (e == cudaSuccess) e = scvx_staged_init();          // per-device kernel attributes, once per context
        if (e != cudaSuccess) { scvx_destroy(c); return fail(SCVX_ERR_CUDA, "context setup on device %d failed: %s", d.id, cudaGetErrorString(e)); }
    }
    *out = c;
    return 0;
}

void scvx_destroy(scvx_ctx* c) {
    if (!c) return;
    for (Dev& d : c->devs) {
        cudaSetDevice(d.id);
        for (int s = 0; s < NSLOT; ++s) {
            Slot& sl = d.slot[s];
            if (sl.stream) { cudaStreamSynchronize(sl.stream); cudaStreamDestroy(sl.stream); }
            double* bufs[] = { sl.dX, sl.dU, sl.dS, sl.dOut, sl.dErr, sl.dTlb, sl.dEnd, sl.dCmp };
            for (double* b : bufs) if (b) cudaFree(b);
        }
        for (int t = 0; t < 3; ++t) if (d.coef[t]) cudaFree(d.coef[t]);
        if (d.dP) cudaFree(d.dP);
        if (d.scratch) cudaFree(d.scratch);
        if (d.dense) cudaFree(d.dense);
        for (int t = 0; t < 2; ++t) if (d.fin[t]) cudaFree(d.fin[t]);
        if (d.ev0) cudaEventDestroy(d.ev0);
        if (d.ev1) cudaEventDestroy(d.ev1);
        if (d.ev_scratch) cudaEventDestroy(d.ev_scratch);
    }
    delete c;
}

int scvx_set_params(scvx_ctx* c, const scvx_probinfo* p, int n) {
    if (!c || !p || n < 1) return fail(SCVX_ERR_ARG, "scvx_set_params: bad arguments");
    for (int i = 0; i < n; ++i)
        if (p[i].aero_kind != SCVX_AERO_EXO && p[i].aero_kind != SCVX_AERO_TABLE)
            return fail(SCVX_ERR_ARG, "record %d: unknown aero_kind %d", i, p[i].aero_kind);
    c->hP.assign(p, p + n);
    c->hP_values = true;
    c->hP_sweep = n > 1;
    for (int i = 1; i < n && c->hP_sweep; ++i) {
        scvx_probinfo r = p[i];
        r.a = p[0].a; r.Tmin = p[0].Tmin; r._pad = p[0]._pad;
        c->hP_sweep = std::memcmp(&r, &p[0], sizeof r) == 0;
    }
    c->any_aero = false;
    for (int i = 0; i < n; ++i) if (p[i].aero_kind == SCVX_AERO_TABLE) c->any_aero = true;
    for (Dev& d : c->devs) {
        CK(cudaSetDevice(d.id));
        // kernels still in flight (on the pipeline streams or on a non-blocking user stream) read the records
        CK(cudaDeviceSynchronize());
        if (d.nP < n) {
            if (d.dP) cudaFree(d.dP);
            d.dP = nullptr; d.nP = 0;
            CK(cudaMalloc((void**)&d.dP, (size_t)n * sizeof(scvx_probinfo)));
            d.nP = n;
        }
        CK(cudaMemcpy(d.dP, p, (size_t)n * sizeof(scvx_probinfo), cudaMemcpyHostToDevice));
    }
    return 0;
}

// upload (and, unless prefiltered, prefilter on the device) one n1 x n2 table into *dst (freshly allocated)
static int upload_table(scvx_ctx* c, Dev& d, double** dst, const double* samples, int n1, int n2, int prefiltered) {
    const size_t ncoef = (size_t)(n1 + 2) * (n2 + 2);
    const int nmax = std::max(n1, n2);
    if (*dst) { cudaFree(*dst); *dst = nullptr; }
    CK(cudaMalloc((void**)dst, ncoef * sizeof(double)));
    cudaStream_t s = d.slot[0].stream;
    if (prefiltered) {
        CK(cudaMemcpyAsync(*dst, samples, ncoef * sizeof(double), cudaMemcpyHostToDevice, s));
    } else {
        std::vector<double> cp(nmax);
        cp[0] = 0.25;
        for (int m = 1; m < nmax; ++m) cp[m] = 1.0 / (4.0 - cp[m - 1]);
        DevTmp ds, dt, dcp;                              // freed on every exit path
        CK(ds.alloc((size_t)n1 * n2 * 8));
        CK(dt.alloc((size_t)(n1 + 2) * n2 * 8));
        CK(dcp.alloc((size_t)nmax * 8));
        CK(cudaMemcpyAsync(ds.d(), samples, (size_t)n1 * n2 * 8, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(dcp.d(), cp.data(), (size_t)nmax * 8, cudaMemcpyHostToDevice, s));
        CK(scvx_launch_prefilter(ds.d(), n1, n2, dt.d(), *dst, dcp.d(), s));
        c->launches += 2;
        CK(cudaStreamSynchronize(s));
    }
    CK(cudaStreamSynchronize(s));
    return 0;
}

int scvx_set_fin_table(scvx_ctx* c, int which, const double* samples, int n_mach, int n_defl, double mach0, double dmach,
                       double defl0, double ddefl, int prefiltered) {
    if (!c || !samples) return fail(SCVX_ERR_ARG, "scvx_set_fin_table: null argument");
    if (which < 0 || which > 1) return fail(SCVX_ERR_ARG, "unknown fin table id %d (0 lift, 1 drag)", which);
    if (n_mach < 2 || n_defl < 2) return fail(SCVX_ERR_ARG, "table needs at least 2 samples per axis");
    if (!(dmach > 0.0) || !(ddefl > 0.0)) return fail(SCVX_ERR_ARG, "axis steps must be positive");
    for (Dev& d : c->devs) {
        CK(cudaSetDevice(d.id));
        if (d.fn1 && (d.fn1 != n_mach || d.fn2 != n_defl || d.fx0 != mach0 || d.fdx != dmach || d.fy0 != defl0 || d.fdy != ddefl)) {
            const int o = 1 - which;
            if (d.fin[o]) { cudaFree(d.fin[o]); d.fin[o] = nullptr; }       // a new geometry invalidates the other table
        }
        if (int rc = upload_table(c, d, &d.fin[which], samples, n_mach, n_defl, prefiltered)) return rc;
        d.fn1 = n_mach; d.fn2 = n_defl; d.fx0 = mach0; d.fdx = dmach; d.fy0 = defl0; d.fdy = ddefl;
    }
    return 0;
}

int scvx_fin_force_batch(scvx_ctx* c, const double* mach, const double* deflection, int n, double* out_lift, double* out_drag) {
    if (!c || !mach || !deflection) return fail(SCVX_ERR_ARG, "scvx_fin_force_batch: null argument");
    if (!out_lift && !out_drag) return fail(SCVX_ERR_ARG, "no output requested");
    if (n < 0) return fail(SCVX_ERR_ARG, "n=%d is negative", n);
    if (n == 0) return 0;
    const bool dev = is_device_ptr(mach);
    if (dev != is_device_ptr(deflection) || (out_lift && dev != is_device_ptr(out_lift)) || (out_drag && dev != is_device_ptr(out_drag)))
        return fail(SCVX_ERR_ARG, "all array arguments must be either host or device pointers, not a mix");
    Dev* dp = &c->devs[0];
    cudaStream_t s = dp->slot[0].stream;
    if (dev) { if (int rc = device_call_target(c, mach, &dp, &s)) return rc; }
    Dev& d = *dp;
    CK(cudaSetDevice(d.id));
    if ((out_lift && !d.fin[0]) || (out_drag && !d.fin[1])) return fail(SCVX_ERR_STATE, "fin table not uploaded (scvx_set_fin_table)");
    ScvxTables lt, dt;
    lt.drag = d.fin[0]; lt.lift = nullptr; lt.trq = nullptr; lt.n1 = d.fn1; lt.n2 = d.fn2;
    lt.wdrag = nullptr; lt.wlift = nullptr; lt.wi0 = 0; lt.wj0 = 0;
    lt.x0 = d.fx0; lt.inv_dx = 1.0 / d.fdx; lt.y0 = d.fy0; lt.inv_dy = 1.0 / d.fdy;
    dt = lt; dt.drag = d.fin[1];
    if (dev) {
        CK(scvx_launch_fin_force(lt, dt, mach, deflection, n, out_lift, out_drag, s));
        c->launches += 1;
        return 0;
    }
    Slot& sl = d.slot[0];
    CK(cudaStreamSynchronize(sl.stream));
    if (grow(&sl.dS, &sl.capS, (size_t)4 * n)) return SCVX_ERR_NOMEM;
    double *dm = sl.dS, *dd = sl.dS + n, *dl = sl.dS + 2 * (size_t)n, *dg = sl.dS + 3 * (size_t)n;
    CK(cudaMemcpyAsync(dm, mach, (size_t)n * 8, cudaMemcpyHostToDevice, sl.stream));
    CK(cudaMemcpyAsync(dd, deflection, (size_t)n * 8, cudaMemcpyHostToDevice, sl.stream));
    CK(scvx_launch_fin_force(lt, dt, dm, dd, n, out_lift ? dl : nullptr, out_drag ? dg : nullptr, sl.stream));
    c->launches += 1;
    if (out_lift) CK(cudaMemcpyAsync(out_lift, dl, (size_t)n * 8, cudaMemcpyDeviceToHost, sl.stream));
    if (out_drag) CK(cudaMemcpyAsync(out_drag, dg, (size_t)n * 8, cudaMemcpyDeviceToHost, sl.stream));
    CK(cudaStreamSynchronize(sl.stream));
    return 0;
}

// SURVEY.md §8f-4 variant on the generic forward-mode kernel (see the header)
static int linearize_fins_impl(scvx_ctx* c, const double* X, const double* U5, const double* sigma, double dt, int npts, int mode,
                               int n_nodes, int B, double* out_blocks, double* out_lin_err) {
    if (!c) return fail(SCVX_ERR_ARG, "null context");
    if (!X || !U5 || !sigma || !out_blocks) return fail(SCVX_ERR_ARG, "X, U5, sigma and out_blocks must be non-null");
    if (n_nodes < 2) return fail(SCVX_ERR_ARG, "n_nodes=%d: at least two nodes (one interval) are required", n_nodes);
    if (B < 0) return fail(SCVX_ERR_ARG, "B=%d is negative", B);
    if (npts < 1) return fail(SCVX_ERR_ARG, "npts=%d must be >= 1", npts);
    if (mode != SCVX_MODE_LITERAL && mode != SCVX_MODE_TEXTBOOK) return fail(SCVX_ERR_ARG, "unknown mode %d", mode);
    if (!(dt > 0.0)) return fail(SCVX_ERR_ARG, "base_dt must be positive");
    if (B == 0) return 0;
    if ((long)(n_nodes - 1) * (long)B >= (1L << 31) - 64) return fail(SCVX_ERR_ARG, "too many intervals for one call");
    if (int rc = check_ready(c, B)) return rc;
    for (const auto& p : c->hP) if (p.aero_kind != SCVX_AERO_TABLE) return fail(SCVX_ERR_STATE, "the fins variant needs aero_kind = TABLE");
    const int ni = n_nodes - 1, nP = (int)c->hP.size();
    const bool dev = is_device_ptr(X);
    if (dev != is_device_ptr(U5) || dev != is_device_ptr(sigma) || dev != is_device_ptr(out_blocks) ||
        (out_lin_err && dev != is_device_ptr(out_lin_err)))
        return fail(SCVX_ERR_ARG, "all array arguments must be either host or device pointers, not a mix");
    Dev* dp = &c->devs[0];
    cudaStream_t s = dp->slot[0].stream;
    if (dev) { if (int rc = device_call_target(c, X, &dp, &s)) return rc; }
    Dev& d = *dp;
    CK(cudaSetDevice(d.id));
    if (!d.coef[SCVX_TABLE_TORQUE]) return fail(SCVX_ERR_STATE, "the fins variant needs the torque table (scvx_set_aero_table, which = 2)");
    ScvxBatch bt;
    bt.P = d.dP; bt.n_params = nP; bt.n_nodes = n_nodes; bt.dt = dt; bt.npts = npts; bt.mode = mode;
    bt.out_tlb = nullptr; bt.out_endpoints = nullptr;
    if (dev) {
        bt.X = X; bt.U = U5; bt.sigma = sigma; bt.B = B; bt.out_blocks = out_blocks; bt.out_lin_err = out_lin_err;
        CK(scvx_launch_dualwarp_fins(bt, tables_of(d), s));
        c->launches += 1;
        return 0;
    }
    // host pointers: sequential trajectory chunks through the first pipeline slot (a feature variant, not the hot path)
    Slot& sl = d.slot[0];
    const size_t per_traj = (size_t)ni * 14 * 27;
    const long chunk = std::max(1L, (long)(((size_t)256 << 20) / (per_traj * 8)));
    for (long cb = 0; cb < B; cb += chunk) {
        const int nb = (int)std::min(chunk, (long)B - cb);
        CK(cudaStreamSynchronize(sl.stream));
        if (grow(&sl.dX, &sl.capX, (size_t)nb * n_nodes * 14) || grow(&sl.dU, &sl.capU, (size_t)nb * n_nodes * 5) ||
            grow(&sl.dS, &sl.capS, (size_t)nb) || grow(&sl.dOut, &sl.capOut, (size_t)nb * per_traj) ||
            (out_lin_err && grow(&sl.dErr, &sl.capErr, (size_t)nb * ni * 14)))
            return SCVX_ERR_NOMEM;
        CK(cudaMemcpyAsync(sl.dX, X + (size_t)cb * n_nodes * 14, (size_t)nb * n_nodes * 14 * 8, cudaMemcpyHostToDevice, sl.stream));
        CK(cudaMemcpyAsync(sl.dU, U5 + (size_t)cb * n_nodes * 5, (size_t)nb * n_nodes * 5 * 8, cudaMemcpyHostToDevice, sl.stream));
        CK(cudaMemcpyAsync(sl.dS, sigma + cb, (size_t)nb * 8, cudaMemcpyHostToDevice, sl.stream));
        bt.X = sl.dX; bt.U = sl.dU; bt.sigma = sl.dS; bt.B = nb; bt.out_blocks = sl.dOut; bt.out_lin_err = out_lin_err ? sl.dErr : nullptr;
        bt.P = (nP == 1) ? d.dP : d.dP + cb; bt.n_params = (nP == 1) ? 1 : nb;
        CK(scvx_launch_dualwarp_fins(bt, tables_of(d), sl.stream));
        c->launches += 1;
        CK(cudaMemcpyAsync(out_blocks + (size_t)cb * per_traj, sl.dOut, (size_t)nb * per_traj * 8, cudaMemcpyDeviceToHost, sl.stream));
        if (out_lin_err)
            CK(cudaMemcpyAsync(out_lin_err + (size_t)cb * ni * 14, sl.dErr, (size_t)nb * ni * 14 * 8, cudaMemcpyDeviceToHost, sl.stream));
    }
    CK(cudaStreamSynchronize(sl.stream));
    return 0;
}

int scvx_linearize_batch_fins(scvx_ctx* c, const double* X, const double* U5, const double* sigma, double base_dt, int npts,
                              int mode, int n_nodes, int B, double* out_blocks, double* out_lin_err) {
    try {
        const int rc = linearize_fins_impl(c, X, U5, sigma, base_dt, npts, mode, n_nodes, B, out_blocks, out_lin_err);
        if (rc != 0 && c) drain(c);
        return rc;
    } catch (...) { return fail(SCVX_ERR_STATE, "unexpected C++ exception"); }
}

int scvx_set_aero_table(scvx_ctx* c, int which, const double* samples, int n_cos, int n_mach, double cos0,
                        double dcos, double mach0, double dmach, int prefiltered) {
    if (!c || !samples) return fail(SCVX_ERR_ARG, "scvx_set_aero_table: null argument");
    if (which < 0 || which > 2) return fail(SCVX_ERR_ARG, "unknown table id %d", which);
    if (n_cos < 2 || n_mach < 2) return fail(SCVX_ERR_ARG, "table needs at least 2 samples per axis");
    if (!(dcos > 0.0) || !(dmach > 0.0)) return fail(SCVX_ERR_ARG, "axis steps must be positive");
    for (Dev& d : c->devs) {
        CK(cudaSetDevice(d.id));
        CK(cudaDeviceSynchronize());                   // nothing in flight may still read the table being replaced
        if (d.n1 && (d.n1 != n_cos || d.n2 != n_mach || d.x0 != cos0 || d.dx != dcos || d.y0 != mach0 || d.dy != dmach)) {
            // a new geometry invalidates tables uploaded with the old one
            for (int t = 0; t < 3; ++t) if (t != which && d.coef[t]) { cudaFree(d.coef[t]); d.coef[t] = nullptr; }
        }
        if (int rc = upload_table(c, d, &d.coef[which], samples, n_cos, n_mach, prefiltered)) return rc;
        d.n1 = n_cos; d.n2 = n_mach; d.x0 = cos0; d.dx = dcos; d.y0 = mach0; d.dy = dmach;
    }
    return 0;
}

int scvx_get_aero_coefficients(scvx_ctx* c, int which, double* out) {
    if (!c || !out || which < 0 || which > 2) return fail(SCVX_ERR_ARG, "scvx_get_aero_coefficients: bad arguments");
    Dev& d = c->devs[0];
    if (!d.coef[which]) return fail(SCVX_ERR_STATE, "table %d has not been uploaded", which);
    CK(cudaSetDevice(d.id));
    CK(cudaMemcpy(out, d.coef[which], (size_t)(d.n1 + 2) * (d.n2 + 2) * 8, cudaMemcpyDeviceToHost));
    return 0;
}

int scvx_linearize_batch(scvx_ctx* c, const double* X, const double* U, const double* sigma, double base_dt,
                         int npts, int mode, int n_nodes, int B, double* out_blocks, double* out_lin_err,
                         double* out_tlb) {
    try {
        return run(c, false, X, U, sigma, base_dt, npts, mode, n_nodes, B, out_blocks, out_lin_err, out_tlb, nullptr);
    } catch (...) { return fail(SCVX_ERR_STATE, "unexpected C++ exception"); }
}

int scvx_predict_batch(scvx_ctx* c, const double* X, const double* U, const double* sigma, double base_dt,
                       int npts, int mode, int n_nodes, int B, double* out_endpoints) {
    try {
        return run(c, true, X, U, sigma, base_dt, npts, mode, n_nodes, B, nullptr, nullptr, nullptr, out_endpoints);
    } catch (...) { return fail(SCVX_ERR_STATE, "unexpected C++ exception"); }
}

int scvx_defect_cost_batch(scvx_ctx* c, const double* X, const double* lin_err, int n_nodes, int B, double wNu,
                           double* out_defect, double* out_cost) {
    if (!c) return fail(SCVX_ERR_ARG, "null context");
    if (!X || !lin_err || !out_defect) return fail(SCVX_ERR_ARG, "X, lin_err and out_defect must be non-null");
    if (n_nodes < 2) return fail(SCVX_ERR_ARG, "n_nodes=%d: at least two nodes are required", n_nodes);
    if (B < 0) return fail(SCVX_ERR_ARG, "B=%d is negative", B);
    if (B == 0) return 0;
    const bool dev = is_device_ptr(X);
    if (dev != is_device_ptr(lin_err) || dev != is_device_ptr(out_defect) || (out_cost && dev != is_device_ptr(out_cost)))
        return fail(SCVX_ERR_ARG, "all array arguments must be either host or device pointers, not a mix");
    if (dev) {
        Dev* dd; cudaStream_t s;
        if (int rc = device_call_target(c, X, &dd, &s)) return rc;
        CK(scvx_launch_defect_cost(X, lin_err, n_nodes, B, wNu, out_defect, out_cost, s));
        c->launches += 1;
        return 0;
    }
    Dev& d = c->devs[0];
    CK(cudaSetDevice(d.id));
    // host pointers: small arrays, one staging round trip on the first device
    Slot& sl = d.slot[0];
    CK(cudaStreamSynchronize(sl.stream));
    const size_t nX = (size_t)B * n_nodes * 14, nE = (size_t)B * (n_nodes - 1) * 14;
    if (grow(&sl.dX, &sl.capX, nX) || grow(&sl.dErr, &sl.capErr, nE) || grow(&sl.dS, &sl.capS, (size_t)2 * B)) return SCVX_ERR_NOMEM;
    CK(cudaMemcpyAsync(sl.dX, X, nX * 8, cudaMemcpyHostToDevice, sl.stream));
    CK(cudaMemcpyAsync(sl.dErr, lin_err, nE * 8, cudaMemcpyHostToDevice, sl.stream));
    CK(scvx_launch_defect_cost(sl.dX, sl.dErr, n_nodes, B, wNu, sl.dS, out_cost ? sl.dS + B : nullptr, sl.stream));
    c->launches += 1;
    CK(cudaMemcpyAsync(out_defect, sl.dS, (size_t)B * 8, cudaMemcpyDeviceToHost, sl.stream));
    if (out_cost) CK(cudaMemcpyAsync(out_cost, sl.dS + B, (size_t)B * 8, cudaMemcpyDeviceToHost, sl.stream));
    CK(cudaStreamSynchronize(sl.stream));
    return 0;
}

int scvx_linear_points_batch(scvx_ctx* c, const double* rIi, const double* vIi, const double* mwet, double mwet_shared,
                             double mdry, const double* rIf, const double* vIf, double g, int K, int B, double* X,
                             double* U) {
    if (!c) return fail(SCVX_ERR_ARG, "null context");
    if (!rIi || !vIi || !rIf || !vIf || !X || !U) return fail(SCVX_ERR_ARG, "null array argument");
    if (K < 1) return fail(SCVX_ERR_ARG, "K=%d must be >= 1", K);
    if (B < 0) return fail(SCVX_ERR_ARG, "B=%d is negative", B);
    if (B == 0) return 0;
    const bool dev = is_device_ptr(rIi);
    if (dev != is_device_ptr(vIi) || dev != is_device_ptr(X) || dev != is_device_ptr(U) || (mwet && dev != is_device_ptr(mwet)))
        return fail(SCVX_ERR_ARG, "all array arguments must be either host or device pointers, not a mix");
    const size_t n = (size_t)(K + 1) * B;
    if (dev) {
        Dev* dd; cudaStream_t s;
        if (int rc = device_call_target(c, X, &dd, &s)) return rc;
        CK(scvx_launch_linear_points(rIi, vIi, mwet, mwet_shared, mdry, rIf, vIf, g, K, B, X, U, s));
        c->launches += 1;
        return 0;
    }
    Dev& d = c->devs[0];
    CK(cudaSetDevice(d.id));
    Slot& sl = d.slot[0];
    CK(cudaStreamSynchronize(sl.stream));
    if (grow(&sl.dX, &sl.capX, n * 14) || grow(&sl.dU, &sl.capU, n * 3) || grow(&sl.dS, &sl.capS, (size_t)7 * B)) return SCVX_ERR_NOMEM;
    double *dr = sl.dS, *dv = sl.dS + (size_t)3 * B, *dm = sl.dS + (size_t)6 * B;
    CK(cudaMemcpyAsync(dr, rIi, (size_t)3 * B * 8, cudaMemcpyHostToDevice, sl.stream));
    CK(cudaMemcpyAsync(dv, vIi, (size_t)3 * B * 8, cudaMemcpyHostToDevice, sl.stream));
    if (mwet) CK(cudaMemcpyAsync(dm, mwet, (size_t)B * 8, cudaMemcpyHostToDevice, sl.stream));
    CK(scvx_launch_linear_points(dr, dv, mwet ? dm : nullptr, mwet_shared, mdry, rIf, vIf, g, K, B, sl.dX, sl.dU, sl.stream));
    c->launches += 1;
    CK(cudaMemcpyAsync(X, sl.dX, n * 14 * 8, cudaMemcpyDeviceToHost, sl.stream));
    CK(cudaMemcpyAsync(U, sl.dU, n * 3 * 8, cudaMemcpyDeviceToHost, sl.stream));
    CK(cudaStreamSynchronize(sl.stream));
    return 0;
}

int scvx_dispersed_setup_batch(scvx_ctx* c, const scvx_dim_problem* base, const double* rIi, const double* vIi,
                               const double* mwet, int B, double* X, double* U, double* sigma, double* scales,
                               scvx_probinfo* out_params, int install) {
    if (!c || !base) return fail(SCVX_ERR_ARG, "null context or base problem");
    if (!rIi || !vIi || !X || !U || !sigma) return fail(SCVX_ERR_ARG, "null array argument");
    if (base->K < 1) return fail(SCVX_ERR_ARG, "K=%d must be >= 1", base->K);
    if (base->aero_kind != SCVX_AERO_EXO && base->aero_kind != SCVX_AERO_TABLE) return fail(SCVX_ERR_ARG, "unknown aero_kind %d", base->aero_kind);
    if (!(base->tf_guess > 0.0)) return fail(SCVX_ERR_ARG, "tf_guess must be positive");
    if (B < 0) return fail(SCVX_ERR_ARG, "B=%d is negative", B);
    if (B == 0) return 0;
    const bool dev = is_device_ptr(rIi);
    if (dev != is_device_ptr(vIi) || dev != is_device_ptr(X) || dev != is_device_ptr(U) || dev != is_device_ptr(sigma) ||
        (mwet && dev != is_device_ptr(mwet)) || (scales && dev != is_device_ptr(scales)) ||
        (out_params && dev != is_device_ptr(out_params)))
        return fail(SCVX_ERR_ARG, "all array arguments must be either host or device pointers, not a mix");
    if (dev && dev_index_of(c, rIi) != 0)
        return fail(SCVX_ERR_ARG, "scvx_dispersed_setup_batch: device pointers must live on the context's first device");
    Dev& d = c->devs[0];
    CK(cudaSetDevice(d.id));
    const size_t n = (size_t)(base->K + 1) * B;
    cudaStream_t s = (dev && c->have_user_stream) ? c->user_stream : d.slot[0].stream;
    scvx_probinfo* dP = nullptr;
    if (install) {
        // the records are written straight into the context's parameter array (nothing in flight may still read it)
        CK(cudaDeviceSynchronize());
        if (d.nP < B) {
            if (d.dP) cudaFree(d.dP);
            d.dP = nullptr; d.nP = 0;
            CK(cudaMalloc((void**)&d.dP, (size_t)B * sizeof(scvx_probinfo)));
            d.nP = B;
        }
        dP = d.dP;
    }
    if (dev) {
        CK(scvx_launch_dispersed_setup(*base, rIi, vIi, mwet, B, X, U, sigma, scales, dP, out_params, s));
        c->launches += 1;
    } else {
        Slot& sl = d.slot[0];
        CK(cudaStreamSynchronize(sl.stream));
        if (grow(&sl.dX, &sl.capX, n * 14) || grow(&sl.dU, &sl.capU, n * 3) || grow(&sl.dS, &sl.capS, (size_t)11 * B)) return SCVX_ERR_NOMEM;
        double *dr = sl.dS, *dv = sl.dS + (size_t)3 * B, *dm = sl.dS + (size_t)6 * B, *dsg = sl.dS + (size_t)7 * B, *dsc = sl.dS + (size_t)8 * B;
        scvx_probinfo* dOutP = nullptr;
        if (out_params && !dP) {
            if (grow(&sl.dEnd, &sl.capEnd, ((size_t)B * sizeof(scvx_probinfo) + 7) / 8)) return SCVX_ERR_NOMEM;
            dOutP = reinterpret_cast<scvx_probinfo*>(sl.dEnd);
        }
        CK(cudaMemcpyAsync(dr, rIi, (size_t)3 * B * 8, cudaMemcpyHostToDevice, sl.stream));
        CK(cudaMemcpyAsync(dv, vIi, (size_t)3 * B * 8, cudaMemcpyHostToDevice, sl.stream));
        if (mwet) CK(cudaMemcpyAsync(dm, mwet, (size_t)B * 8, cudaMemcpyHostToDevice, sl.stream));
        CK(scvx_launch_dispersed_setup(*base, dr, dv, mwet ? dm : nullptr, B, sl.dX, sl.dU, dsg, dsc, dP, dOutP, sl.stream));
        c->launches += 1;
        CK(cudaMemcpyAsync(X, sl.dX, n * 14 * 8, cudaMemcpyDeviceToHost, sl.stream));
        CK(cudaMemcpyAsync(U, sl.dU, n * 3 * 8, cudaMemcpyDeviceToHost, sl.stream));
        CK(cudaMemcpyAsync(sigma, dsg, (size_t)B * 8, cudaMemcpyDeviceToHost, sl.stream));
        if (scales) CK(cudaMemcpyAsync(scales, dsc, (size_t)3 * B * 8, cudaMemcpyDeviceToHost, sl.stream));
        if (out_params)
            CK(cudaMemcpyAsync(out_params, dP ? dP : dOutP, (size_t)B * sizeof(scvx_probinfo), cudaMemcpyDeviceToHost, sl.stream));
        CK(cudaStreamSynchronize(sl.stream));
    }
    if (install) {
        CK(cudaStreamSynchronize(s));       // a set-up call: later work on any stream sees the installed records
        // other devices of the context receive a copy; the host mirror only carries what the argument checks read
        // (count and aero kind) — the values live on the devices
        for (size_t i = 1; i < c->devs.size(); ++i) {
            Dev& o = c->devs[i];
            CK(cudaSetDevice(o.id));
            for (int q = 0; q < NSLOT; ++q) CK(cudaStreamSynchronize(o.slot[q].stream));
            if (o.nP < B) {
                if (o.dP) cudaFree(o.dP);
                o.dP = nullptr; o.nP = 0;
                CK(cudaMalloc((void**)&o.dP, (size_t)B * sizeof(scvx_probinfo)));
                o.nP = B;
            }
            CK(cudaMemcpyPeer(o.dP, o.id, d.dP, d.id, (size_t)B * sizeof(scvx_probinfo)));
        }
        CK(cudaSetDevice(d.id));
        scvx_probinfo tmpl;
        memset(&tmpl, 0, sizeof(tmpl));
        tmpl.aero_kind = base->aero_kind;
        c->hP.assign((size_t)B, tmpl);
        c->hP_values = false; c->hP_sweep = false;
        c->any_aero = base->aero_kind == SCVX_AERO_TABLE;
    }
    return 0;
}

int scvx_socp_dims(int n_nodes, int* n_rows, int* n_cols, int* nnz) {
    if (n_nodes < 2) return fail(SCVX_ERR_ARG, "n_nodes=%d: at least two nodes are required", n_nodes);
    if (n_nodes > 6000000) return fail(SCVX_ERR_ARG, "n_nodes=%d: value indices would exceed 32 bits", n_nodes);
    const int K = n_nodes - 1;
    if (n_rows) *n_rows = socp_rows(K);
    if (n_cols) *n_cols = socp_cols(K);
    if (nnz) *nnz = socp_nnz(K);
    return 0;
}

int scvx_socp_pattern(int n_nodes, int32_t* colptr, int32_t* rowind) {
    if (!colptr || !rowind) return fail(SCVX_ERR_ARG, "colptr and rowind must be non-null");
    if (int rc = scvx_socp_dims(n_nodes, nullptr, nullptr, nullptr)) return rc;
    const int K = n_nodes - 1, nnz = socp_nnz(K), nc = socp_cols(K);
    for (int j = 0; j <= nc; ++j) colptr[j] = 0;
    for (int p = 0; p < nnz; ++p) {
        const SocpEntry e = socp_decode(p, K);
        rowind[p] = e.row;
        colptr[e.col + 1] += 1;
    }
    for (int j = 0; j < nc; ++j) colptr[j + 1] += colptr[j];
    return 0;
}

static int socp_values_impl(scvx_ctx* c, const double* blocks, const double* lin_err, const double* tlb, int n_nodes, int B,
                            double* out_vals, double* out_rhs);

int scvx_socp_values_batch(scvx_ctx* c, const double* blocks, const double* lin_err, const double* tlb, int n_nodes, int B,
                           double* out_vals, double* out_const) {
    const int rc = socp_values_impl(c, blocks, lin_err, tlb, n_nodes, B, out_vals, out_const);
    if (rc != 0 && c) drain(c);        // no async copy into caller-owned buffers may outlive a failed call
    return rc;
}

static int socp_values_impl(scvx_ctx* c, const double* blocks, const double* lin_err, const double* tlb, int n_nodes, int B,
                            double* out_vals, double* out_rhs) {
    if (!c) return fail(SCVX_ERR_ARG, "null context");
    if (!blocks || !tlb || !out_vals) return fail(SCVX_ERR_ARG, "blocks, tlb and out_vals must be non-null");
    if (out_rhs && !lin_err) return fail(SCVX_ERR_ARG, "out_rhs needs lin_err");
    if (int rc = scvx_socp_dims(n_nodes, nullptr, nullptr, nullptr)) return rc;
    if (B < 0) return fail(SCVX_ERR_ARG, "B=%d is negative", B);
    if (B == 0) return 0;
    const bool dev = is_device_ptr(blocks);
    if (dev != is_device_ptr(tlb) || dev != is_device_ptr(out_vals) || (lin_err && dev != is_device_ptr(lin_err)) ||
        (out_rhs && dev != is_device_ptr(out_rhs)))
        return fail(SCVX_ERR_ARG, "all array arguments must be either host or device pointers, not a mix");
    const int K = n_nodes - 1;
    if (n_nodes > 50000)
        return fail(SCVX_ERR_ARG, "n_nodes=%d: scvx_socp_values_batch covers the 340K+4 value indices of a trajectory with one "
                                  "launch (gridDim.y <= 65535 blocks of 256), i.e. n_nodes <= 50000", n_nodes);
    const size_t nnz = socp_nnz(K), nr = socp_rows(K), nblk = (size_t)K * SCVX_BLOCK_DOUBLES, nerr = (size_t)14 * K, ntlb = (size_t)4 * n_nodes;
    if (dev) {
        Dev* dd; cudaStream_t s;
        if (int rc = device_call_target(c, blocks, &dd, &s)) return rc;
        CK(scvx_launch_socp_values(blocks, lin_err, tlb, n_nodes, B, out_vals, out_rhs, s));
        c->launches += 1;
        return 0;
    }
    Dev& d = c->devs[0];
    CK(cudaSetDevice(d.id));
    // host pointers: trajectory chunks through the two pipeline slots of the first device
    long chunk = std::max(1L, (long)(((size_t)128 << 20) / (nblk * sizeof(double))));
    for (long cb = 0, ci = 0; cb < B; cb += chunk, ++ci) {
        const int nb = (int)std::min(chunk, (long)B - cb);
        Slot& sl = d.slot[ci % NSLOT];
        CK(cudaStreamSynchronize(sl.stream));
        if (grow(&sl.dOut, &sl.capOut, nb * nblk) || grow(&sl.dTlb, &sl.capTlb, nb * ntlb) || grow(&sl.dEnd, &sl.capEnd, nb * nnz) ||
            (out_rhs && (grow(&sl.dErr, &sl.capErr, nb * nerr) || grow(&sl.dX, &sl.capX, nb * nr))))
            return SCVX_ERR_NOMEM;
        CK(cudaMemcpyAsync(sl.dOut, blocks + cb * nblk, nb * nblk * 8, cudaMemcpyHostToDevice, sl.stream));
        CK(cudaMemcpyAsync(sl.dTlb, tlb + cb * ntlb, nb * ntlb * 8, cudaMemcpyHostToDevice, sl.stream));
        if (out_rhs) CK(cudaMemcpyAsync(sl.dErr, lin_err + cb * nerr, nb * nerr * 8, cudaMemcpyHostToDevice, sl.stream));
        CK(scvx_launch_socp_values(sl.dOut, out_rhs ? sl.dErr : nullptr, sl.dTlb, n_nodes, nb, sl.dEnd, out_rhs ? sl.dX : nullptr, sl.stream));
        c->launches += 1;
        CK(cudaMemcpyAsync(out_vals + cb * nnz, sl.dEnd, nb * nnz * 8, cudaMemcpyDeviceToHost, sl.stream));
        if (out_rhs) CK(cudaMemcpyAsync(out_rhs + cb * nr, sl.dX, nb * nr * 8, cudaMemcpyDeviceToHost, sl.stream));
    }
    for (int q = 0; q < NSLOT; ++q) CK(cudaStreamSynchronize(d.slot[q].stream));
    return 0;
}

int scvx_set_stream(scvx_ctx* c, void* stream) {
    if (!c) return fail(SCVX_ERR_ARG, "null context");
    c->user_stream = (cudaStream_t)stream;
    c->have_user_stream = stream != nullptr;        // NULL = back to the library's own stream
    return 0;
}

int scvx_set_kernel(scvx_ctx* c, int which) {
    if (!c) return fail(SCVX_ERR_ARG, "null context");
    if (which < SCVX_KERNEL_AUTO || which > SCVX_KERNEL_STAGED) return fail(SCVX_ERR_ARG, "unknown kernel id %d", which);
    c->kernel = which;
    return 0;
}

int scvx_synchronize(scvx_ctx* c) {
    if (!c) return fail(SCVX_ERR_ARG, "null context");
    for (Dev& d : c->devs) {
        CK(cudaSetDevice(d.id));
        for (int q = 0; q < NSLOT; ++q) CK(cudaStreamSynchronize(d.slot[q].stream));
    }
    if (c->have_user_stream) { CK(cudaSetDevice(c->devs[0].id)); CK(cudaStreamSynchronize(c->user_stream)); }
    return 0;
}

int64_t scvx_launch_count(scvx_ctx* c) { return c ? c->launches : 0; }

int scvx_last_kernel_ms(scvx_ctx* c, double* ms) {
    if (!c || !ms) return fail(SCVX_ERR_ARG, "null argument");
    if (!c->timed) return fail(SCVX_ERR_STATE, "no device-pointer call has been made yet");
    Dev& d = c->devs[0];
    CK(cudaSetDevice(d.id));
    CK(cudaEventSynchronize(d.ev1));
    float f = 0.f;
    CK(cudaEventElapsedTime(&f, d.ev0, d.ev1));
    *ms = (double)f;
    return 0;
}

int scvx_compact_record_doubles(int layout) {
    if (layout == SCVX_COMPACT_FULL) return SCVX_COMPACT_DOUBLES;
    if (layout == SCVX_COMPACT_NO_Z) return SCVX_COMPACT_NO_Z_DOUBLES;
    return fail(SCVX_ERR_ARG, "unknown compact layout %d", layout);
}

int scvx_linearize_batch_compact(scvx_ctx* c, const double* X, const double* U, const double* sigma, double base_dt,
                                 int npts, int mode, int n_nodes, int B, int layout, double* out_compact, double* out_tlb) {
    try {
        if (!out_compact) return fail(SCVX_ERR_ARG, "out_compact is null");
        const int rec = scvx_compact_record_doubles(layout);
        if (rec < 0) return rec;
        return run(c, false, X, U, sigma, base_dt, npts, mode, n_nodes, B, nullptr, nullptr, out_tlb, nullptr, out_compact, rec);
    } catch (...) { return fail(SCVX_ERR_STATE, "unexpected C++ exception"); }
}

int scvx_compact_layout(int32_t* index) {
    if (!index) return fail(SCVX_ERR_ARG, "index is null");
    int tmp[SCVX_COMPACT_DATA];
    const int n = compact_fill_layout(tmp);
    for (int i = 0; i < n; ++i) index[i] = tmp[i];
    return n == SCVX_COMPACT_DATA ? 0 : fail(SCVX_ERR_STATE, "compact layout has %d entries", n);
}

int64_t scvx_expand_compact(const double* compact, int layout, const double* X, const double* U, const double* sigma,
                            int n_nodes, int B, double* out_blocks, double* out_lin_err, int n_threads) {
    if (!compact) return fail(SCVX_ERR_ARG, "compact is null");
    if (n_nodes < 2 || B < 0) return fail(SCVX_ERR_ARG, "bad n_nodes / B");
    const int rec_n = scvx_compact_record_doubles(layout);
    if (rec_n < 0) return rec_n;
    const bool no_z = layout == SCVX_COMPACT_NO_Z;
    if (out_lin_err && !X) return fail(SCVX_ERR_ARG, "out_lin_err needs X");
    if (no_z && out_blocks && (!X || !U || !sigma)) return fail(SCVX_ERR_ARG, "layout NO_Z needs X, U and sigma to re-form z");
    if (is_device_ptr(compact) || is_device_ptr(out_blocks) || is_device_ptr(out_lin_err))
        return fail(SCVX_ERR_ARG, "scvx_expand_compact works on host memory");
    try {
        const int ni = n_nodes - 1, n_data = rec_n - 1;
        const long total = (long)ni * B;
        // template block: the structural constants; the data entries are overwritten per interval
        double tmpl[SCVX_BLOCK_DOUBLES];
        int idx[SCVX_COMPACT_DATA];
        compact_fill_layout(idx);
        for (int cc = 0; cc < 23; ++cc) for (int r = 0; r < 14; ++r) tmpl[cc * 14 + r] = compact_constant(cc, r);
        int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
        nt = (int)std::max(1L, std::min((long)nt, (total + 4095) / 4096));
        std::atomic<int64_t> flagged(0);
        auto work = [&](long w0, long w1) {
            int64_t bad = 0;
            for (long w = w0; w < w1; ++w) {
                const double* rec = compact + (size_t)w * rec_n;
                if (rec[n_data] != 0.0) ++bad;
                const long b = w / ni, i = w - b * ni;
                if (out_blocks) {
                    double* blk = out_blocks + (size_t)w * SCVX_BLOCK_DOUBLES;
                    memcpy(blk, tmpl, sizeof(tmpl));
                    for (int k = 0; k < n_data; ++k) blk[idx[k]] = rec[k];
                    if (no_z) {
                        // z = endpoint - D * inp, inp = [x_i; u_i; u_{i+1}; sigma]   (old_dynamics.jl:139, 150-153)
                        const double* x = X + ((size_t)b * n_nodes + i) * 14;
                        const double* u = U + ((size_t)b * n_nodes + i) * 3;
                        double inp[21];
                        for (int k = 0; k < 14; ++k) inp[k] = x[k];
                        for (int k = 0; k < 6; ++k) inp[14 + k] = u[k];
                        inp[20] = sigma[b];
                        for (int r = 0; r < 14; ++r) {
                            double acc = blk[r];
                            for (int cc = 0; cc < 21; ++cc) acc -= blk[14 * (1 + cc) + r] * inp[cc];
                            blk[14 * 22 + r] = acc;
                        }
                    }
                }
                if (out_lin_err) {
                    const double* xn = X + ((size_t)b * n_nodes + i + 1) * 14;       // x_{n+1} (rocketland.jl:130, 256)
                    double* e = out_lin_err + (size_t)w * 14;
                    for (int r = 0; r < 14; ++r) e[r] = rec[r] - xn[r];
                }
            }
            flagged += bad;
        };
        if (nt == 1) work(0, total);
        else {
            std::vector<std::thread> pool;
            for (int t = 0; t < nt; ++t) pool.emplace_back(work, total * t / nt, total * (t + 1) / nt);
            for (auto& th : pool) th.join();
        }
        return flagged.load();
    } catch (...) { return fail(SCVX_ERR_STATE, "unexpected C++ exception"); }
}

int scvx_host_alloc(void** out, uint64_t bytes) {
    if (!out) return fail(SCVX_ERR_ARG, "out is null");
    *out = nullptr;
    if (bytes == 0) return 0;
    cudaError_t e = cudaHostAlloc(out, (size_t)bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(SCVX_ERR_NOMEM, "cudaHostAlloc(%llu B) failed: %s", (unsigned long long)bytes, cudaGetErrorString(e)); }
    return 0;
}

int scvx_host_free(void* p) {
    if (!p) return 0;
    CK(cudaFreeHost(p));
    return 0;
}

int scvx_host_register(void* p, uint64_t bytes) {
    if (!p || bytes == 0) return fail(SCVX_ERR_ARG, "scvx_host_register: null pointer or zero size");
    CK(cudaHostRegister(p, (size_t)bytes, cudaHostRegisterPortable));
    return 0;
}

int scvx_host_unregister(void* p) {
    if (!p) return 0;
    CK(cudaHostUnregister(p));
    return 0;
}

}  // extern "C"
