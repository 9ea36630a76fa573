#!/usr/bin/env python
"""Benchmark of the artgpn likelihood hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this implementation
    python bench.py --impl reference --gpus N ...            # the reference algorithm on host cores

A "step" is one pass of the hot path over one batch of synthetic walkers: W parameter rows ->
W log-likelihoods (mean residual, covariance assembly, blocked FP64 Cholesky, solve, log-det).
Workload (default C3 = BASELINE.json configs[2], the one the metric is quoted on): synthetic
3 x 2048 points, 2 QuasiPeriodic nodes, Constant weights, Linear means, 128 walkers PER GPU
(weak scaling: walkers are independent); `--scaling strong` keeps 128
walkers in total.  Rank 0 prints ONE JSON line.

`value`  : evals/s, theta already resident in HBM, CUDA-event timed, max over ranks.
`e2e`    : evals/s through network.log_likelihood_batch(host numpy theta) -- H2D of theta and
           D2H of the results inside the timed region.
`roofline`: the dominant kernel -- the trailing update.  Default path: syrk_i8_kernel, exact integer products
           of 7-bit slices on the INT8 tensor path (tcgen05.mma.kind::i8), executed INT8 ops against the
           kind::i8 issue peak measured in this run (agpn_i8_peak); with AGPN_SYRK=dmma: syrk_kernel against the
           FP64 tensor (DMMA.8x8x4) issue peak measured in this run.
`roofline_cholesky`: the whole batched Cholesky (diagonal-block + panel-solve + digit-slicing + trailing-update
           kernels), ALGORITHMIC FP64 flops p N^3 / 3 per eval against the DMMA.8x8x4 peak (BASELINE's
           "% FP64 TC peak"; above 1.0 on the INT8-sliced path).  `roofline_assembly`: the covariance
           assembly kernel against the measured HBM bandwidth.
`cpu_baseline`: the oracle port of the reference (numpy + scipy LAPACK) on the host cores, a bounded
           sample of the same walkers, reference-style process pool with 1 BLAS thread each.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "log-likelihood evals/sec at 3x2048 points, batch of walkers; % FP64 TC peak"


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm (the only place bench.py executes oracle/)
# ---------------------------------------------------------------------------------------------
def _cpu_worker(args):
    os.environ["OMP_NUM_THREADS"] = "1"
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, ROOT)
    import artgpn_oracle as oracle
    from artgpn_b200 import synth
    N, p, q, wk, rows = args
    t, y, sig = synth.make_data(N, p, seed=0)

    class M(object):
        def __getattr__(self, name):
            return lambda *pars: (name, list(pars))
    m = M()
    out = []
    for r in rows:
        out.append(oracle.log_likelihood(t, y, sig, q, *synth.objects_from_row(m, m, m, r, p, q, wk)))
    return out


def cpu_evals_per_sec(cfg, wk, n_evals, ncores, W_total=None):
    """Reference-style parallel mode (examples/Sun/01_emcee.py:3-4,119-122): a process pool with
    OMP_NUM_THREADS=1, one walker per task.  Returns (evals/s, values, seconds)."""
    import multiprocessing as mp
    from artgpn_b200 import synth
    # the FIRST n_evals rows of the very batch the GPU arm evaluates (make_walkers draws column-wise)
    theta = synth.make_walkers(max(W_total or cfg["W"], n_evals), p=cfg["p"], q=cfg["q"], weights=wk, seed=1)[:n_evals]
    old = os.environ.get("OMP_NUM_THREADS")
    os.environ["OMP_NUM_THREADS"] = "1"          # inherited by the spawned workers before they import numpy
    try:
        ctx = mp.get_context("spawn")
        with ctx.Pool(ncores) as pool:
            tasks = [(cfg["N"], cfg["p"], cfg["q"], wk, [theta[i]]) for i in range(n_evals)]
            pool.map(_cpu_worker, [(64, cfg["p"], cfg["q"], wk, [theta[0]])] * ncores)      # import / warm-up
            t0 = time.perf_counter()
            res = pool.map(_cpu_worker, tasks, chunksize=1)
            dt = time.perf_counter() - t0
    finally:
        if old is None:
            os.environ.pop("OMP_NUM_THREADS", None)
        else:
            os.environ["OMP_NUM_THREADS"] = old
    return n_evals / dt, [r[0] for r in res], dt


# ---------------------------------------------------------------------------------------------
class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for ts, r in self.rows if t0 <= ts <= t1 + 0.2] or [r for _, r in self.rows]
        sm, mx, reasons = [], None, set()
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C3", choices=["C2", "C3", "C4"])
    ap.add_argument("--weights", default="Constant", choices=["Constant", "SquaredExponential"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--walkers", type=int, default=0, help="override walkers per GPU (weak) / total (strong)")
    ap.add_argument("--cpu-evals", type=int, default=0, help="CPU baseline sample size (default: host cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    from artgpn_b200 import synth
    cfg = dict(synth.CONFIGS[args.workload])
    if args.walkers:
        cfg["W"] = args.walkers
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    ncores = os.cpu_count() or 1
    workload_name = "%s: synthetic %dx%d points, %d QuasiPeriodic nodes, %s weights, Linear means, %d walkers" % (
        args.workload, cfg["p"], cfg["N"], cfg["q"], args.weights, cfg["W"])

    # ---------------------------------------------------------------- reference arm (CPU only)
    if args.impl == "reference":
        if rank != 0:
            return 0
        n = args.cpu_evals or ncores
        for _ in range(max(args.warmup, 0) and 1):
            cpu_evals_per_sec(cfg, args.weights, min(n, ncores), ncores)
        vals, secs = [], 0.0
        for _ in range(args.steps):
            v, _, dt = cpu_evals_per_sec(cfg, args.weights, n, ncores)
            vals.append(v)
            secs += dt
        value = n * args.steps / secs
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": workload_name, "sample": "%d walkers of the workload per step" % n},
                "cpu_baseline": {"value": value, "unit": "evals/s", "cores": ncores, "kind": "port",
                                 "sample": "%d evals/step x %d steps, process pool of %d, OMP_NUM_THREADS=1 "
                                           "(reference's own parallel mode); oracle port = numpy + scipy LAPACK as the "
                                           "reference" % (n, args.steps, ncores)},
                "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ---------------------------------------------------------------- B200 arm
    # Libraries (NCCL prints its version banner) may write to stdout: keep fd 1 for the JSON line only.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # before CUDA is initialised in this process; bounded sample (one eval per host core)
        n = args.cpu_evals or ncores
        v, cpu_vals, dt = cpu_evals_per_sec(cfg, args.weights, n, ncores)
        cpu_base = {"value": v, "unit": "evals/s", "cores": ncores, "kind": "port",
                    "sample": "%d walkers of the workload, process pool of %d, OMP_NUM_THREADS=1, %.1f s" % (n, ncores, dt),
                    "_vals": cpu_vals}

    import torch
    import torch.distributed as dist
    from artgpn_b200 import node, weight, mean, _lib
    from artgpn_b200.arttwo import network
    from artgpn_b200 import sharding
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    t, y, sig = synth.make_data(cfg["N"], cfg["p"], seed=0)
    if args.scaling == "weak":
        W_total = cfg["W"] * world
    else:
        W_total = cfg["W"]
    theta_all = synth.make_walkers(W_total, p=cfg["p"], q=cfg["q"], weights=args.weights, seed=1)
    lo, hi = sharding.shard_bounds(W_total, world, rank)
    theta_h = np.ascontiguousarray(theta_all[lo:hi])
    W_local = hi - lo
    dargs = []
    for i in range(cfg["p"]):
        dargs += [y[i], sig[i]]
    net = network(cfg["q"], t, *dargs, device=local_rank)
    o = synth.objects_from_row(node, weight, mean, theta_all[0], cfg["p"], cfg["q"], args.weights)
    S = net.set_structure(*o[:3])
    theta_d = torch.from_numpy(theta_h).to(dev)
    theta_pin = torch.from_numpy(theta_h).pin_memory()
    lib = _lib.load()

    def step_device():
        local = net.log_likelihood_batch(theta_d)
        if world > 1:
            return sharding.gather_loglikes(local, W_total)
        return local

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        out = step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = lib.agpn_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    w0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        out = step_device()
    e1.record()
    barrier()
    w1 = time.perf_counter()
    launches = lib.agpn_launch_count() - launches0
    ms = e0.elapsed_time(e1)
    if world > 1:
        tms = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    clocks = sampler.stop(w0, w1) if rank == 0 else None
    value = W_total * args.steps / (ms * 1e-3)
    result_host = out.cpu().numpy()

    # ---- end to end: host theta (pinned) -> H2D -> kernels -> D2H of the log-likelihoods ----------
    theta_pin_np = theta_pin.numpy()
    for _ in range(2):
        net.log_likelihood_batch(theta_pin_np)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ll_host = net.log_likelihood_batch(theta_pin_np)          # returns after the D2H copy completed
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        ts = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        e2e_s = float(ts.item())
    e2e_value = W_total * args.steps / e2e_s
    assert np.array_equal(ll_host, result_host[lo:hi]), "host-buffer and device-buffer paths disagree"

    # ---- per-phase device times (separate profiled passes; events around every launch) ------------
    net.set_profiling(True)
    phase = np.zeros(6)
    nl = np.zeros(6, dtype=np.int64)
    slice_ms, slice_n = 0.0, 0
    PROF_STEPS = 2
    for _ in range(PROF_STEPS):
        net.log_likelihood_batch(theta_d)
        msv, nv = net.phase_times()
        phase += msv
        nl = nv
        sm_, slice_n = net.slice_times()
        slice_ms += sm_
    phase /= PROF_STEPS
    slice_ms /= PROF_STEPS
    net.set_profiling(False)
    update_path = net.update_path()

    if rank == 0:
        peak_tf = np.zeros(1)
        _lib.check(lib.agpn_dmma_peak(local_rank, _lib.ptr(peak_tf)))
        peaks, peak_src = load_peaks()
        p, N = cfg["p"], cfg["N"]
        chol_flops = W_local * p * N ** 3 / 3.0                   # F_chol = p N^3 / 3 per eval
        chol_ms = float(phase[2] + phase[3] + phase[4] + slice_ms)
        step_prof_ms = float(phase.sum() + slice_ms)
        chol_tf = chol_flops / (chol_ms * 1e-3) / 1e12
        nblk = (N + 127) // 128
        syrk_tiles = sum((nblk - k - 1) * (nblk - k) // 2 for k in range(nblk))
        syrk_tf = W_local * p * syrk_tiles * 2.0 * 128 ** 3 / (phase[4] * 1e-3) / 1e12 if phase[4] > 0 else None
        asm_bytes = W_local * 4.0 * p * N * (N + 1)               # triangle-only store: B_tri = 4 p N (N+1)
        asm_gbs = asm_bytes / (phase[1] * 1e-3) / 1e9
        chol_traffic = asm_traffic = syrk_traffic = None          # DRAM bytes from the committed ncu pass
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                tr = json.load(f)
            if tr.get("workload") == args.workload and args.weights == "Constant" and tr.get("update_path") == update_path:
                chol_traffic = tr["cholesky_dram_bytes_per_step"] * W_local / tr["walkers"]
                asm_traffic = tr["assembly_dram_bytes_per_launch"] * W_local / tr["walkers"]
                syrk_traffic = tr["syrk_dram_bytes_per_step"] * W_local / tr["walkers"]
        except Exception:
            pass
        dmma_peak = float(peak_tf[0])
        syrk_share = float(phase[4] / max(step_prof_ms, 1e-9))
        if update_path:
            i8_peak = np.zeros(1)
            _lib.check(lib.agpn_i8_peak(local_rank, _lib.ptr(i8_peak)))
            i8_peak = float(i8_peak[0])
            # 36 slice pairs (s + t <= 7) per FP64 product, 2 ops per int8 multiply-add
            syrk_tops = 36.0 * syrk_tf if syrk_tf else None
            roofline = {
                "bound": "tensor",
                "kernel": "syrk_i8_persistent_kernel: trailing update as exact integer products of 8 x 7-bit slices, "
                          "tcgen05.mma.kind::i8 with TMEM int32 accumulators (all launches of a step)",
                "achieved": syrk_tops, "peak": i8_peak, "unit": "TOP/s",
                "frac": (syrk_tops / i8_peak) if syrk_tops else None,
                "peak_source": "kind::i8 issue peak (M=128, N=256, operands resident in shared memory) measured in this run "
                               "(agpn_i8_peak); MEASURED_PEAKS.json has no INT8 figure (2 x its bf16 burst = %.0f TOP/s)"
                               % (2.0 * peaks.get("bf16_tflops", float("nan"))),
                "ops_per_step": (36.0 * W_local * p * syrk_tiles * 2.0 * 128 ** 3),
                "ms": float(phase[4]), "launches": int(nl[4]), "share_of_step_ms": syrk_share,
                "traffic": syrk_traffic,
                "traffic_note": "DRAM bytes of all syrk_i8_persistent_kernel launches of one step, ncu pass in profiles/ (scaled by walkers)",
                "fp64_equivalent": {"tflops": syrk_tf, "dmma_peak_tflops": dmma_peak,
                                    "ratio_to_dmma_peak": (syrk_tf / dmma_peak) if syrk_tf else None,
                                    "note": "the same update on the FP64 tensor instruction cannot exceed 1.0"},
                "note": "executed ops (full diagonal tiles counted), CUDA-event time of the launches in a profiled pass"}
        else:
            roofline = {
                "bound": "tensor", "kernel": "syrk_kernel (DMMA.8x8x4; all narrow + wide launches of a step)",
                "achieved": syrk_tf, "peak": dmma_peak, "unit": "TFLOP/s",
                "frac": (syrk_tf / dmma_peak) if syrk_tf else None,
                "peak_source": "DMMA.8x8x4 issue peak measured in this run (agpn_dmma_peak); MEASURED_PEAKS.json has no FP64 figure",
                "ms": float(phase[4]), "launches": int(nl[4]), "share_of_step_ms": syrk_share, "traffic": syrk_traffic,
                "note": "flops actually executed by the tile GEMMs (full diagonal tiles counted), CUDA-event time of the syrk launches"}
        line = {
            "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name + (" per GPU" if args.scaling == "weak" else " in total"),
                       "walkers_total": W_total, "walkers_per_gpu": W_local, "theta_dim": S.D,
                       "l2": "working set %.1f GB of covariance blocks per step per GPU >> 126 MB L2, rewritten "
                             "every step (no flush needed)" % (W_local * p * (nblk * 128) ** 2 * 8 / 1e9),
                       "parallelism": "walkers block-partitioned over %d GPU(s); one all_gather of the "
                                      "log-likelihoods per step" % world},
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": int(theta_h.nbytes),
                    "d2h_bytes_per_step": int(W_local * 12)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "roofline_cholesky": {"bound": "tensor",
                                  "kernel": "batched Cholesky = potrf_diag + trsm + digit slicing + trailing-update kernels",
                                  "achieved": chol_tf, "peak": dmma_peak, "unit": "TFLOP/s", "frac": chol_tf / dmma_peak,
                                  "peak_source": "DMMA.8x8x4 issue peak measured in this run (agpn_dmma_peak); "
                                                 "MEASURED_PEAKS.json has no FP64 figure",
                                  "flops_per_step": chol_flops, "ms": chol_ms, "traffic": chol_traffic,
                                  "traffic_note": "DRAM bytes of all Cholesky launches of one step, ncu pass in profiles/ (scaled by walkers)",
                                  "note": "ALGORITHMIC FP64 flops p N^3 / 3 per eval (SURVEY 8(d)) over the device time of all "
                                          "factorisation kernels; a fraction above 1.0 means faster than the FP64 tensor "
                                          "instruction can go: the trailing update runs FP64-equivalent on the INT8 tensor path"},
            "update_path": {0: "dmma", 1: "i8 narrow", 2: "i8 wide", 3: "i8"}.get(update_path, str(update_path)),
            "roofline_assembly": {"bound": "hbm", "kernel": "assemble_fast_kernel (lower block triangle only; generic assemble_kernel for other kernel classes)",
                                  "achieved": asm_gbs, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                                  "frac": asm_gbs / peaks.get("hbm_gbs"), "peak_source": peak_src,
                                  "bytes_per_launch": asm_bytes, "ms": float(phase[1]), "traffic": asm_traffic},
            "phase_ms": {"mean_residual": float(phase[0]), "assembly": float(phase[1]), "potrf_diag": float(phase[2]),
                         "trsm": float(phase[3]), "syrk": float(phase[4]), "finalize": float(phase[5]),
                         "digit_slicing": float(slice_ms), "launches": [int(x) for x in nl] + [int(slice_n)]},
        }
        if cpu_base is not None:
            ref_vals = np.array(cpu_base.pop("_vals"))
            got = result_host[:len(ref_vals)]
            line["cpu_baseline"] = cpu_base
            line["parity_vs_cpu_sample"] = {"max_rel_err": float(np.max(np.abs(got - ref_vals) / np.abs(ref_vals))),
                                            "n": int(len(ref_vals)), "tolerance": 1e-9}
        else:
            line["cpu_baseline"] = None
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
