#!/usr/bin/env python
"""bench.py — FOH interval discretisations / second (FP64, K=50) on N B200s, beside the CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[4] = SURVEY.md §8d C5, the configuration the metric is quoted on): Monte-Carlo
dispersion of the aero-table sample problem, K=50 (51 nodes, 50 intervals per trajectory), 32768 trajectories PER
GPU (262144 on 8 GPUs -> weak scaling), per-iteration discretisation + SOCP-assembly extras (lin_err, thrust-LB).
One "step" = one pass of the hot path over that batch.  No data-path collective: trajectories are sharded.

  value  : intervals/s, inputs and outputs resident in HBM, CUDA events on the launching stream, max over ranks
  e2e    : same metric through the C-ABI call with pinned HOST buffers (H2D + kernels + D2H inside the timed region)
  e2e    : same metric through the C-ABI call a Monte-Carlo host makes, scvx_linearize_batch_compact, with pinned HOST
           buffers (H2D + kernels + pack + D2H inside the timed region); the dense-block call is timed beside it
  parity : conditioning-aware check of a random sample of the TIMED outputs against the CPU oracle in FP64 and binary128
  roofline : binding resource is the FP64 FMA pipe; `peak` is measured in this run by a DFMA
             microbenchmark (libscvx_benchtools.so; MEASURED_PEAKS.json holds no FP64 figure), the fraction of the
             nominal 37.2 TFLOP/s is given beside it, and so is the HBM view
  extra  : (N=1) the other BASELINE configs: C2 single-trajectory call latency, C3 / C4 throughput + roofline
  cpu_baseline : the CPU oracle (C++ restatement of the reference's Julia path; Julia is not installed) on the
             box's host cores, bounded sample.  A reported baseline, not the target.
`--impl reference` times that CPU implementation alone (rank 0 only) and prints the same JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

AERO_NPZ = os.path.join(ROOT, "tests", "golden", "aero_lift_drag.npz")
K_NODES = 50                    # problem.K -> 51 nodes, 50 intervals (reference convention, initial_solve.jl:115-116)
TRAJ_PER_GPU = 32768            # C5: 262144 trajectories over 8 GPUs
SEED = 1003
NPTS = 10
METRIC = "FOH interval discretisations/sec (FP64, K=50)"
UNIT = "intervals/s"
# algorithmic work per interval (SURVEY.md §8d / BASELINE.md §3), aero-table variant, npts = 10
FLOP_PER_INTERVAL_AERO = 10 * (4 * (300 + 600 + 2 * 21 * 60 + 2 * 6 * 21 + 28) + 4620)     # 194 200
BYTES_PER_INTERVAL = 8 * (21 + 14 * 23) + 8 * (14 + 4)                                        # 2 888 incl. lin_err + tlb
FP64_NOMINAL_TF = 148 * 64 * 2 * 1.965e9 * 1e-12                                              # 37.2
PARITY_SAMPLE = 64              # trajectories of the timed outputs checked against the oracle (3 200 intervals)


def fp64_peak_tf(device_index):
    """DFMA issue-rate microbenchmark (measurement helper library, not part of the product ABI)."""
    import ctypes
    from successiveconvexification_b200 import _lib
    tf = ctypes.c_double()
    rc = _lib.load_benchtools().scvx_bench_fp64_peak(int(device_index), ctypes.byref(tf))
    if rc != 0:
        raise RuntimeError(f"scvx_bench_fp64_peak failed ({rc})")
    return tf.value


def parity_of_timed_outputs(prob, X, U, sigma, P, dOut, dt, mode, threads):
    """Conditioning-aware parity of a random sample of the timed outputs (tests/conftest.py: 1e-10 against the binary128
    evaluation wherever FP64 can hold it; within K_COND x the reference arithmetic's own distance elsewhere)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import conftest
    from oracle import oracle
    tb = oracle.OracleTables.from_aero(prob.aero)
    B = X.shape[0]
    pick = np.sort(np.random.default_rng(SEED).choice(B, min(PARITY_SAMPLE, B), replace=False))
    got = dOut[torch.from_numpy(pick).to(dOut.device)].cpu().numpy()
    Pp = P if P.shape[0] == 1 else P[pick]
    ref64, refq, same, kap = conftest.reference_resolution(Pp, tb, X[pick], U[pick], sigma[pick], dt, NPTS, mode,
                                                           nthreads=threads)
    rep = conftest.conditioned_parity(got, refq, same, kap)
    rep["trajectories_sampled"] = int(len(pick))
    rep["tolerance"] = conftest.PARITY_TOL
    rep["k_cond"] = conftest.K_COND
    rep["pass"] = bool(rep["max_metric_well_conditioned"] <= conftest.PARITY_TOL and
                       rep["max_err_over_kappa_eps_ill_conditioned"] <= conftest.K_COND)
    rep["vs_fp64_oracle_max_metric"] = float(conftest.parity_metric_per_interval(got, ref64).max())
    rep["note"] = ("sample of the TIMED device outputs vs the CPU oracle in IEEE binary128; LITERAL rk4 at sigma up to 15 is "
                   "ill-conditioned, so intervals are split by the measured resolution kappa*eps of the reference's FP64 "
                   "arithmetic (spread of 9 FP64 oracle evaluations at inputs within one ulp); pass = 1e-10 on the "
                   "well-conditioned ones and <= k_cond x kappa*eps on the others")
    return rep


class ClockSampler:
    """SM clock / throttle reasons DURING the timed region (B200_PROFILING.md recipe).  NVML in-process (a sample
    every ~20 ms); falls back to polling nvidia-smi when pynvml is missing."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
        except Exception:
            self.nvml = None

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[index])
            except Exception:
                return index
        return index

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)
        pw = n.nvmlDeviceGetPowerUsage(self.h) / 1000.0
        get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        r = get(self.h)
        flags = [("Active" if r & m else "Not Active") for m in (0x8, 0x40, 0x20, 0x4)]   # hw_slowdown, hw_thermal, sw_thermal, sw_power_cap
        self.rows.append([str(sm), str(mx), str(pw)] + flags)

    def _run(self):
        while not self._stop.is_set():
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                          str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append([c.strip() for c in out.splitlines()[0].split(",")])
            except Exception:
                pass
            self._stop.wait(0.02 if self.nvml is not None else 0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows), "power_w_max": max(float(r[2]) for r in self.rows),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def bind_to_gpu_numa(index):
    """Pin this process to the CPU cores NVML reports as local to GPU `index` (one process per GPU: keeps the pinned
    staging buffers and the D2H traffic on the GPU's own NUMA node).  Best effort; returns a note for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(ClockSampler._physical_index(index))
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"bound to {len(cpus)} cores local to GPU {index}"
    except Exception as e:          # noqa: BLE001
        return f"not bound ({type(e).__name__})"
    return "not bound"


def cpu_reference_run(prob, steps, warmup, traj_per_thread):
    """The reference's CPU implementation of the path (oracle port), all host threads, bounded sample per step."""
    from oracle import oracle
    from successiveconvexification_b200 import workloads
    tb = oracle.OracleTables.from_aero(prob.aero)
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core it is allowed to run on
    threads = max(oracle.max_threads(), len(os.sched_getaffinity(0)))
    n_traj = max(8, traj_per_thread * threads)
    X, U, sigma, P = workloads.monte_carlo_batch(prob, K_NODES, n_traj, SEED)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        oracle.linearize_batch(P, tb, X, U, sigma, 1.0 / (K_NODES + 1), NPTS, 0, True, True, nthreads=threads)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    value = n_traj * K_NODES * len(times) / total
    sample = (f"{n_traj} trajectories x {K_NODES} intervals of the C5 batch per step, {len(times)} timed steps "
              f"({total:.1f} s), OpenMP over intervals")
    return value, threads, sample, total / len(times)


def other_configs(prob, ctx, dev, args, fp64_peak):
    """The other BASELINE configs on one GPU (parity-test cases; reported for completeness, not the headline):
      C2  one trajectory x 50 intervals through the host-pointer ABI: call latency (median of 200) — the only number the
          unmodified SCvx loop (rocketland.jl:318, one linearize_dynamics per iteration) would feel;
      C3  aero K=100, 4096 trajectories;  C4  K=400, 16384 trajectories, per-trajectory parameter sweep:
          device-resident throughput + FP64 roofline fraction."""
    import torch
    from successiveconvexification_b200 import dynamics, workloads
    out = {}
    # ---- C2
    X, U, sigma, dt = workloads.sample_trajectory(prob)
    hX, hU, hS = (torch.from_numpy(a).pin_memory() for a in (X, U, sigma))
    hOut = torch.empty((1, 50, 23, 14), dtype=torch.float64).pin_memory()
    lat = []
    for it in range(220):
        t0 = time.perf_counter()
        ctx.linearize_ptr(hX.data_ptr(), hU.data_ptr(), hS.data_ptr(), dt, NPTS, args.mode, 51, 1, hOut.data_ptr())
        if it >= 20:
            lat.append(time.perf_counter() - t0)
    lat = np.array(lat) * 1e6
    out["c2"] = {"workload": "C2: sample trajectory, 1 x 50 intervals, host pointers, dense blocks", "calls": 200,
                 "latency_us_median": float(np.median(lat)), "latency_us_p90": float(np.percentile(lat, 90)),
                 "intervals_per_s": 50.0 / (float(np.median(lat)) * 1e-6)}
    # ---- C3, C4 (device resident)
    stream = torch.cuda.current_stream()
    for name, K, B, seed, sweep in (("c3", 100, 4096, 1001, False), ("c4", 400, 16384, 1002, True)):
        Xc, Uc, sc, Pc = workloads.monte_carlo_batch(prob, K, B, seed, sweep=sweep)
        if sweep:
            ptr, n, keep = workloads.as_c_params(Pc)
            ctx.set_params_raw(ptr, n)
        dX, dU, dS = (torch.from_numpy(a).to(dev) for a in (Xc, Uc, sc))
        dO = torch.empty((B, K, 23, 14), dtype=torch.float64, device=dev)
        dT = torch.empty((B, K + 1, 4), dtype=torch.float64, device=dev)

        def step():
            ctx.linearize_ptr(dX.data_ptr(), dU.data_ptr(), dS.data_ptr(), 1.0 / (K + 1), NPTS, args.mode, K + 1, B,
                              dO.data_ptr(), 0, dT.data_ptr())
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        steps = 10 if name == "c3" else 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        tf = FLOP_PER_INTERVAL_AERO * B * K / (ms * 1e-3) * 1e-12
        out[name] = {"workload": f"{name.upper()}: K={K}, {B} trajectories{', per-trajectory parameter sweep' if sweep else ''}, "
                                 f"device resident, mode={'LITERAL' if args.mode == 0 else 'TEXTBOOK'}, sigma ~ U(1, 15)",
                     "intervals_per_s": B * K / (ms * 1e-3), "ms_per_step": ms, "steps": steps,
                     "roofline": {"bound": "fp64", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tf / fp64_peak},
                     "results_finite": bool(torch.isfinite(dO[:: max(1, B // 64)]).all().item()),
                     "l2": f"outputs {dO.nbytes / 1e9:.1f} GB exceed L2"}
        del dX, dU, dS, dO, dT
        torch.cuda.empty_cache()
        if sweep:
            from successiveconvexification_b200.defns import ProbInfo
            ctx.set_params(ProbInfo(prob))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--traj-per-gpu", type=int, default=TRAJ_PER_GPU)
    ap.add_argument("--kernel", type=int, default=0, help="0 auto (= staged), 1 dualwarp, 2 staged")
    ap.add_argument("--mode", type=int, default=0, help="0 LITERAL (parity contract), 1 TEXTBOOK")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-assembly", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--sigma-range", default="1,15", help="sigma ~ U(lo,hi) of the synthetic batch (C5: 1,15)")
    args = ap.parse_args()
    sig_lo, sig_hi = (float(v) for v in args.sigma_range.split(","))
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    from successiveconvexification_b200 import sample_problems as sp, workloads
    prob = sp.base_prob_aero_scaled(AERO_NPZ)
    config = {"workload": f"C5 Monte-Carlo dispersion, aero-table 6-DoF sample problem, K={K_NODES} "
                          f"({K_NODES + 1} nodes, {K_NODES} intervals/trajectory), {args.traj_per_gpu} trajectories per GPU, "
                          f"discretise [A|B-|B+|Sigma|z] + lin_err + thrust-LB rows, npts={NPTS}, "
                          f"mode={'LITERAL' if args.mode == 0 else 'TEXTBOOK'}, sigma ~ U({sig_lo:g}, {sig_hi:g})",
              "trajectories_per_gpu": args.traj_per_gpu, "intervals_per_trajectory": K_NODES, "seed": SEED,
              "sharding": f"trajectory shards, {world} rank(s), no data-path collective",
              "l2": "inputs (227 MB/GPU) and outputs (4.7 GB/GPU) exceed the 126 MB L2; no flush needed"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        value, threads, sample, step_s = cpu_reference_run(prob, args.steps, args.warmup, traj_per_thread=120)
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                                 "sample": sample + "; C++ restatement of the reference Julia path (Julia unavailable)"},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return 0

    # keep stdout for the single JSON line: anything libraries print meanwhile (e.g. the NCCL version banner) goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from successiveconvexification_b200 import dynamics, sharding
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_note = bind_to_gpu_numa(local_rank)     # before any pinned allocation: host buffers land next to the GPU
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the single JSON line
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- this rank's shard (independent reproducible stream per shard; C5 = contiguous shards of 32768)
    B = args.traj_per_gpu
    X, U, sigma, P = workloads.monte_carlo_batch(prob, K_NODES, B, SEED, shard=rank, sigma_range=(sig_lo, sig_hi))
    n_nodes, ni = K_NODES + 1, K_NODES
    dt = 1.0 / (K_NODES + 1)
    cache = dynamics.make_cache(prob, device_ids=[local_rank])
    ctx = cache.sim_prob
    ctx.set_kernel(args.kernel)
    fp64_peak = fp64_peak_tf(local_rank)

    dX, dU, dS = (torch.from_numpy(a).to(dev) for a in (X, U, sigma))
    dOut = torch.empty((B, ni, 23, 14), dtype=torch.float64, device=dev)
    dErr = torch.empty((B, ni, 14), dtype=torch.float64, device=dev)
    dTlb = torch.empty((B, n_nodes, 4), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)

    def step_device():
        ctx.linearize_ptr(dX.data_ptr(), dU.data_ptr(), dS.data_ptr(), dt, NPTS, args.mode, n_nodes, B,
                          dOut.data_ptr(), dErr.data_ptr(), dTlb.data_ptr())

    for _ in range(args.warmup):
        step_device()
    barrier()
    launches0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern_ms = []
    with ClockSampler(local_rank) as clk:
        e0.record(stream)
        for _ in range(args.steps):
            step_device()
        e1.record(stream)
        barrier()
    elapsed_ms = e0.elapsed_time(e1)
    rank_ms_per_step = elapsed_ms / args.steps            # this rank's device time per step over the timed region
    launches = ctx.launch_count() - launches0
    # per-launch kernel time of the dominant kernel, measured live (events inside the library, same stream)
    for _ in range(3):
        step_device()
        torch.cuda.synchronize()
        kern_ms.append(ctx.last_kernel_ms())
    single_step_ms = float(np.median(kern_ms))            # one isolated step (library events), for reference
    kern_ms = rank_ms_per_step                            # roofline: average over the timed region (CUDA events)
    finite = bool(torch.isfinite(dOut[:: max(1, B // 64)]).all().item())
    checksum = float(dOut[:: max(1, B // 64), :, 0, :].sum().item())

    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok, _ = sharding.reduce_status(finite, checksum, dev)      # NCCL is used only for status / result gathering
    else:
        ok = finite
    elapsed_ms = float(t.item())
    # result collection over NCCL: not part of the timed step — a slab of every rank's blocks is sent to rank 0
    # (sharding.gather_to_root: send/recv into one result tensor, no padding, nothing on the other ranks) and checked there
    gather = None
    if world > 1:
        slab = dOut[:64].contiguous()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        g0.record()
        rooted = sharding.gather_to_root(slab, dst=0)
        g1.record()
        torch.cuda.synchronize()
        if rank == 0:
            same = bool(torch.equal(rooted[:64], slab)) and rooted.shape[0] == 64 * world and bool(torch.isfinite(rooted).all())
            gather = {"call": "sharding.gather_to_root (NCCL send/recv)", "trajectories_per_rank": 64,
                      "bytes_collected": int(rooted.nbytes), "ms": g0.elapsed_time(g1), "ok": same}
            del rooted
    total_intervals = B * ni * world
    value = total_intervals * args.steps / (elapsed_ms * 1e-3)

    # ---- auxiliary: fixed-pattern sparse SOCP value writer on the same device-resident results.
    # Pure data movement: algorithmic bytes = 8 B read + 8 B written per value and per constant.
    assembly = None
    if rank == 0 and not args.no_assembly:
        from successiveconvexification_b200 import rocketland
        nr, _, nnz = rocketland.socp_dims(n_nodes)
        dVals = torch.empty((B, nnz), dtype=torch.float64, device=dev)
        dRhs = torch.empty((B, nr), dtype=torch.float64, device=dev)

        def step_assembly():
            ctx.socp_values_ptr(dOut.data_ptr(), dErr.data_ptr(), dTlb.data_ptr(), n_nodes, B, dVals.data_ptr(), dRhs.data_ptr())
        for _ in range(3):
            step_assembly()
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        for _ in range(10):
            step_assembly()
        a1.record(stream)
        torch.cuda.synchronize()
        a_ms = a0.elapsed_time(a1) / 10
        a_bytes = 16.0 * (nnz + nr) * B
        assembly = {"kernel": "socp_values_kernel", "ms": a_ms, "trajectories": B, "nnz_per_trajectory": nnz,
                    "bound": "hbm", "achieved": a_bytes / (a_ms * 1e-3) * 1e-9, "unit": "GB/s",
                    "note": "in+out 8.7 GB per launch exceed L2; not part of the timed step"}
        del dVals, dRhs

    # ---- end to end through the C ABI with pinned host buffers: the compact call (what a Monte-Carlo host uses: only
    # the 229 data entries of a block + a status word cross PCIe) and, beside it, the dense-block call
    e2e = None
    if not args.no_e2e:
        hX, hU, hS = (torch.from_numpy(a).pin_memory() for a in (X, U, sigma))
        hTlb = torch.empty((B, n_nodes, 4), dtype=torch.float64).pin_memory()

        def timed_host(step, steps):
            for _ in range(2):
                step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                step()
            barrier()
            el = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(el, op=dist.ReduceOp.MAX)
            return total_intervals * steps / float(el.item())

        e2e_steps = max(1, min(args.steps, 5))
        hCmp = torch.empty((B, ni, 230), dtype=torch.float64).pin_memory()

        def step_compact():
            ctx.linearize_compact_ptr(hX.data_ptr(), hU.data_ptr(), hS.data_ptr(), dt, NPTS, args.mode, n_nodes, B,
                                      hCmp.data_ptr(), hTlb.data_ptr())
        v_compact = timed_host(step_compact, e2e_steps)
        # the host expander restores the dense blocks bit for bit: check a slice against the device-pointer result
        eb, ee, flagged = dynamics.expand_compact(hCmp[:4].numpy(), X[:4])
        assert np.array_equal(eb, dOut[:4].cpu().numpy()), "compact host path and device-pointer path disagree"
        assert np.array_equal(ee, dErr[:4].cpu().numpy()), "lin_err recomputed on the host differs"
        d2h_compact = int(hCmp.nbytes + hTlb.nbytes)
        n_flag = int((hCmp[..., 229] != 0).sum().item())
        del hCmp
        # layout NO_Z (216-double records: the reference's SOCP consumes D and lin_err, not z)
        hC2 = torch.empty((B, ni, 216), dtype=torch.float64).pin_memory()

        def step_no_z():
            ctx.linearize_compact_ptr(hX.data_ptr(), hU.data_ptr(), hS.data_ptr(), dt, NPTS, args.mode, n_nodes, B,
                                      hC2.data_ptr(), hTlb.data_ptr(), dynamics.COMPACT_NO_Z)
        v_no_z = timed_host(step_no_z, e2e_steps)
        d2h_no_z = int(hC2.nbytes + hTlb.nbytes)
        del hC2
        hOut = torch.empty((B, ni, 23, 14), dtype=torch.float64).pin_memory()
        hErr = torch.empty((B, ni, 14), dtype=torch.float64).pin_memory()

        def step_host():
            ctx.linearize_ptr(hX.data_ptr(), hU.data_ptr(), hS.data_ptr(), dt, NPTS, args.mode, n_nodes, B,
                              hOut.data_ptr(), hErr.data_ptr(), hTlb.data_ptr())
        v_dense = timed_host(step_host, e2e_steps)
        assert torch.equal(hOut[:4], dOut[:4].cpu()), "host-pointer path and device-pointer path disagree"
        e2e = {"value": v_compact, "unit": UNIT,
               "h2d_bytes_per_step": int(hX.nbytes + hU.nbytes + hS.nbytes),
               "d2h_bytes_per_step": d2h_compact, "steps": e2e_steps,
               "call": "scvx_linearize_batch_compact (230-double records: 229 data entries + status word; lin_err recomputed "
                       "exactly by scvx_expand_compact on the host)",
               "intervals_flagged_non_finite": n_flag,
               "d2h_gb_per_s_per_gpu": d2h_compact * v_compact / total_intervals * 1e-9,
               "no_z": {"value": v_no_z, "call": "scvx_linearize_batch_compact, layout NO_Z (216-double records; z re-formed on the host)",
                        "d2h_bytes_per_step": d2h_no_z},
               "dense": {"value": v_dense, "call": "scvx_linearize_batch (14x23 blocks + lin_err + tlb)",
                         "d2h_bytes_per_step": int(hOut.nbytes + hErr.nbytes + hTlb.nbytes)},
               "note": "pinned host buffers; chunked H2D/kernel/D2H pipeline; PCIe-bound; " + numa_note}
        del hOut, hErr, hTlb

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    achieved_tf = FLOP_PER_INTERVAL_AERO * (B * ni) / (kern_ms * 1e-3) * 1e-12
    if achieved_tf > 1.2 * FP64_NOMINAL_TF:
        raise SystemExit(f"device timing is implausible ({achieved_tf:.1f} TFLOP/s FP64): the kernels did not run on the timed stream")
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved_gbs = BYTES_PER_INTERVAL * (B * ni) / (kern_ms * 1e-3) * 1e-9
    traffic = None
    try:
        # DRAM bytes of one tangent_kernel launch of THIS run: the per-interval figure of the ncu capture x the intervals
        # an average launch of the timed region covers (the library splits a step into equal chunks)
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        per_interval = tj.get("dram_bytes_per_interval") or tj["dram_bytes_per_launch"] / tj["intervals_per_launch"]
        chunks_per_step = max(1, int(launches) // max(1, args.steps) // 2)
        traffic = per_interval * (B * ni) / chunks_per_step
    except Exception:
        pass
    roofline = {"bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": achieved_tf / fp64_peak, "traffic": traffic,
                "peak_nominal": FP64_NOMINAL_TF, "frac_nominal": achieved_tf / FP64_NOMINAL_TF,
                "flop_per_interval": FLOP_PER_INTERVAL_AERO, "kernel_ms": kern_ms, "isolated_step_ms": single_step_ms,
                "kernels": "stage_value_kernel + tangent_kernel, all launches of one step (W spans both)",
                "peak_source": "in-run DFMA microbenchmark (libscvx_benchtools.so: scvx_bench_fp64_peak); MEASURED_PEAKS.json "
                               "has no FP64 figure; nominal = 148 SM x 64 FMA/clk x 2 x 1.965 GHz",
                "hbm": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                        "frac": achieved_gbs / hbm_peak, "bytes_per_interval": BYTES_PER_INTERVAL,
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}}

    if assembly:
        assembly["peak"] = hbm_peak
        assembly["frac"] = assembly["achieved"] / hbm_peak
    threads_avail = len(os.sched_getaffinity(0))
    parity = None
    if not args.no_parity:
        parity = parity_of_timed_outputs(prob, X, U, sigma, P, dOut, dt, args.mode, threads_avail)
    extra = None
    if world == 1 and not args.no_extra:
        del dOut, dErr, dTlb
        torch.cuda.empty_cache()
        extra = other_configs(prob, ctx, dev, args, fp64_peak)
    cpu = None
    if not args.no_cpu and world == 1:
        v, threads, sample, _ = cpu_reference_run(prob, steps=1, warmup=1, traj_per_thread=400)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": sample + "; C++ restatement of the reference Julia path (Julia unavailable)"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "clocks": clk.summary(), "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": cpu, "parity": parity, "assembly": assembly, "extra": extra, "gather": gather, "results_finite": ok,
            "kernel": args.kernel}
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
