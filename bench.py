#!/usr/bin/env python
"""Benchmark of the artgpn likelihood hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this implementation
    python bench.py --impl reference --gpus N ...            # the reference algorithm on host cores

A "step" is one pass of the hot path over one batch of synthetic walkers: W parameter rows ->
W log-likelihoods (mean residual, covariance assembly, blocked FP64 Cholesky, solve, log-det).
Workload (default C3 = BASELINE.json configs[2], the one the metric is quoted on): synthetic
3 x 2048 points, 2 QuasiPeriodic nodes, Constant weights, Linear means, 128 walkers IN TOTAL, block-
partitioned over the GPUs (strong scaling: BASELINE config 3 is "128 walkers at 1/2/4/8 B200");
`--scaling weak` gives every GPU 128 walkers instead.  Rank 0 prints ONE JSON line.

`value`  : evals/s, theta already resident in HBM, CUDA-event timed, max over ranks; at N > 1 one call into the C ABI
           per step (agpn_loglike_batch_sharded: local rows + one ncclAllGather).
`e2e`    : evals/s through network.log_likelihood_batch(host numpy theta) -- H2D of theta and
           D2H of the results inside the timed region.
`roofline`: the kernel with the largest share of the step.  On the default path that is either the panel solve
           (trsm_kernel: DMMA.8x8x4 GEMM against the explicit inverse with digit extraction fused; algorithmic flops
           of the triangular solves that were executed, against the FP64 tensor issue peak measured in this run) or the trailing update (syrk_i8_persistent_kernel:
           exact integer products of 7-bit slices on tcgen05.mma.kind::i8; EXECUTED INT8 ops against the kind::i8
           issue peak measured in this run); the other one is reported as `roofline_update` / `roofline_panel_solve`.
           Both kernels leave out work that is exactly / provably nothing: K steps whose digit chunks are all zero,
           panel tile halves bounded below 2^-96 of their row scale.  Only executed work is counted; the fraction is
           in the object, and `dense_schedule` (extras) times the same step with nothing left out.
           With AGPN_SYRK=dmma: syrk_kernel against the FP64 tensor (DMMA.8x8x4) issue peak.
`roofline_cholesky`: the whole batched Cholesky, ALGORITHMIC FP64 flops p N^3 / 3 per eval against the DMMA.8x8x4 peak
           (BASELINE's "% FP64 TC peak"; above 1.0 on the INT8-sliced path).  `roofline_dmma_path`: the same figure
           with every kernel of the factorisation on the FP64 tensor instruction (second context, AGPN_SYRK=dmma).
`roofline_assembly` (+ `_generic`, `_se_weights`): the covariance assembly kernels against the measured HBM bandwidth.
`other_configs`: the other BASELINE.json configs (C2, C4, C5) and a 16-walker single-GPU latency, each with a parity
           check of one row against the CPU port.
`cpu_baseline`: the oracle port of the reference (numpy + scipy LAPACK) on the host cores, a bounded
           sample of the same walkers, reference-style process pool with 1 BLAS thread each (rank 0, any N).
"""
import argparse
import contextlib
import io
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
class _Spec(object):
    """Stand-in for the node / weight / mean modules: Class(*pars) -> ('Class', [pars]) (the oracle's object form)."""
    def __getattr__(self, name):
        return lambda *pars: (name, list(pars))


def _cpu_worker(args):
    """One task of the CPU pool.  ("ll", N, p, q, weights, rows) -> log-likelihoods of `rows`;
    ("ll_m52rq", N, p, rows) -> same for the Matern52 + RationalQuadratic structure;
    ("pred", N, p, q, M, idx) -> (mean, std) of prediction-grid points idx of every series (whole-grid mean centring)."""
    os.environ["OMP_NUM_THREADS"] = "1"
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, ROOT)
    import artgpn_oracle as oracle
    from artgpn_b200 import synth
    m = _Spec()
    kind = args[0]
    if kind == "ll":
        _, N, p, q, wk, rows = args
        t, y, sig = synth.make_data(N, p, seed=0)
        if N >= 4096:      # the one C4 reference row (3 x 8192: 550 GFLOP of dpotrf) may use every core's BLAS thread
            from threadpoolctl import threadpool_limits
            with threadpool_limits(limits=os.cpu_count() or 1):
                return [oracle.log_likelihood(t, y, sig, q, *synth.objects_from_row(m, m, m, r, p, q, wk)) for r in rows]
        return [oracle.log_likelihood(t, y, sig, q, *synth.objects_from_row(m, m, m, r, p, q, wk)) for r in rows]
    if kind == "ll_m52rq":
        _, N, p, rows = args
        t, y, sig = synth.make_data(N, p, seed=0)
        return [oracle.log_likelihood(t, y, sig, 2, *synth.objects_from_row_m52rq(m, m, m, r, p)) for r in rows]
    if kind == "pred":
        _, N, p, q, M, idx, series = args
        t, y, sig = synth.make_data(N, p, seed=0)
        theta = synth.make_walkers(1, p=p, q=q, seed=1)[0]
        o = synth.objects_from_row(m, m, m, theta, p, q)
        grid = synth.prediction_grid(t, M)
        sub = grid[idx]
        mu, sd = oracle.prediction(t, y, sig, q, *o, sub, series)
        # mean.Linear is centred on the mean of the grid it is given (mean.py:92): move from the subset's to the whole grid's
        mu = mu - oracle.mean_vector(o[2], sub, p)[series - 1] + oracle.mean_vector(o[2], grid, p)[series - 1][idx]
        return mu, sd
    raise ValueError(kind)


def cpu_evals_per_sec(cfg, wk, n_evals, ncores, W_total=None, extra_tasks=None):
    """Reference-style parallel mode (examples/Sun/01_emcee.py:3-4,119-122): a process pool with
    OMP_NUM_THREADS=1, one walker per task.  Returns (evals/s, values, seconds, extra results); the extra tasks
    (references for the other configs) run after the timed sample, all at once."""
    import multiprocessing as mp
    from artgpn_b200 import synth
    # the FIRST n_evals rows of the very batch the GPU arm evaluates (make_walkers draws column-wise)
    theta = synth.make_walkers(max(W_total or cfg["W"], n_evals), p=cfg["p"], q=cfg["q"], weights=wk, seed=1)[:n_evals]
    old = os.environ.get("OMP_NUM_THREADS")
    os.environ["OMP_NUM_THREADS"] = "1"          # inherited by the spawned workers before they import numpy
    extra = None
    try:
        ctx = mp.get_context("spawn")
        with ctx.Pool(ncores) as pool:
            tasks = [("ll", cfg["N"], cfg["p"], cfg["q"], wk, [theta[i]]) for i in range(n_evals)]
            pool.map(_cpu_worker, [("ll", 64, cfg["p"], cfg["q"], wk, [theta[0]])] * ncores)      # import / warm-up
            t0 = time.perf_counter()
            res = pool.map(_cpu_worker, tasks, chunksize=1)
            dt = time.perf_counter() - t0
            if extra_tasks:
                extra = pool.map(_cpu_worker, extra_tasks, chunksize=1)
    finally:
        if old is None:
            os.environ.pop("OMP_NUM_THREADS", None)
        else:
            os.environ["OMP_NUM_THREADS"] = old
    return n_evals / dt, [r[0] for r in res], dt, extra


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


def workload_config(args, cfg, world):
    """The `config` object -- identical in both arms (this implementation and --impl reference)."""
    W_total = cfg["W"] * world if args.scaling == "weak" else cfg["W"]
    W_local = -(-W_total // world)
    nblk = (cfg["N"] + 127) // 128
    name = "%s: synthetic %dx%d points, %d QuasiPeriodic nodes, %s weights, Linear means, %d walkers %s" % (
        args.workload, cfg["p"], cfg["N"], cfg["q"], args.weights, cfg["W"],
        "per GPU" if args.scaling == "weak" else "in total")
    return {"workload": name, "walkers_total": W_total, "walkers_per_gpu": W_local, "scaling": args.scaling,
            "l2": "working set %.1f GB of covariance blocks per step per GPU >> 126 MB L2, rewritten "
                  "every step (no flush needed)" % (W_local * cfg["p"] * (nblk * 128) ** 2 * 8 / 1e9),
            "parallelism": "walkers block-partitioned over %d GPU(s); one all-gather of the "
                           "log-likelihoods per step" % world}, W_total


def rel_err(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C3", choices=["C2", "C3", "C4"])
    ap.add_argument("--weights", default="Constant", choices=["Constant", "SquaredExponential"])
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"])
    ap.add_argument("--walkers", type=int, default=0, help="override walkers per GPU (weak) / total (strong)")
    ap.add_argument("--cpu-evals", type=int, default=0, help="CPU baseline sample size (default: host cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip roofline_dmma_path / other_configs / extra assembly figures")
    args = ap.parse_args()

    from artgpn_b200 import synth
    cfg = dict(synth.CONFIGS[args.workload])
    if args.walkers:
        cfg["W"] = args.walkers
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    ncores = os.cpu_count() or 1
    config, W_total = workload_config(args, cfg, world)

    # ---------------------------------------------------------------- reference arm (CPU only)
    if args.impl == "reference":
        if rank != 0:
            return 0
        n = args.cpu_evals or ncores
        for _ in range(max(args.warmup, 0) and 1):
            cpu_evals_per_sec(cfg, args.weights, min(n, ncores), ncores, W_total)
        secs = 0.0
        for _ in range(args.steps):
            _, _, dt, _ = cpu_evals_per_sec(cfg, args.weights, n, ncores, W_total)
            secs += dt
        value = n * args.steps / secs
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config,
                "cpu_baseline": {"value": value, "unit": "evals/s", "cores": ncores, "kind": "port",
                                 "sample": "each step = the first %d walkers of the workload, %d steps, process pool of %d, "
                                           "OMP_NUM_THREADS=1 (reference's own parallel mode); oracle port = numpy + scipy "
                                           "LAPACK as the reference" % (n, args.steps, ncores)},
                "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ---------------------------------------------------------------- B200 arm
    # Libraries (NCCL prints its version banner) may write to stdout: keep fd 1 for the JSON line only.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    extras = not args.no_extras and args.workload == "C3" and args.weights == "Constant"
    cpu_base, cpu_extra = None, None
    C2, C4, C5 = synth.CONFIGS["C2"], synth.CONFIGS["C4"], synth.CONFIGS["C5"]
    pred_idx = np.sort(np.random.default_rng(7).choice(C5["M"], 1500, replace=False))
    if rank == 0 and not args.no_cpu_baseline:
        # before CUDA is initialised in this process; bounded sample (one eval per host core); then, in the same pool,
        # one reference row of each of the other configs
        n = args.cpu_evals or ncores
        tasks = None
        if extras:
            tasks = [("ll", C4["N"], C4["p"], C4["q"], "Constant", [synth.make_walkers(C4["W"], p=C4["p"], q=C4["q"], seed=1)[0]]),
                     ("pred", C5["N"], C5["p"], C5["q"], C5["M"], pred_idx, 1),
                     ("ll", C2["N"], C2["p"], C2["q"], "Constant", list(synth.make_walkers(C2["W"], p=C2["p"], q=C2["q"], seed=1)[:4])),
                     ("ll_m52rq", cfg["N"], cfg["p"], list(synth.make_walkers_m52rq(32, p=cfg["p"], seed=1)[:1])),
                     ("ll", cfg["N"], cfg["p"], cfg["q"], "SquaredExponential",
                      list(synth.make_walkers(32, p=cfg["p"], q=cfg["q"], weights="SquaredExponential", seed=1)[:1]))]
        t_cpu0 = time.perf_counter()
        v, cpu_vals, dt, cpu_extra = cpu_evals_per_sec(cfg, args.weights, n, ncores, W_total, tasks)
        cpu_base = {"value": v, "unit": "evals/s", "cores": ncores, "kind": "port",
                    "sample": "the first %d walkers of the workload, process pool of %d, OMP_NUM_THREADS=1, %.1f s "
                              "(+ %.1f s for one reference row of each other config)" % (n, ncores, dt, time.perf_counter() - t_cpu0 - dt),
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

    def make_net(c, **kw):
        t_, y_, s_ = synth.make_data(c["N"], c["p"], seed=0)
        dargs = []
        for i in range(c["p"]):
            dargs += [y_[i], s_[i]]
        return network(c["q"], t_, *dargs, device=local_rank, **kw), t_

    theta_all = synth.make_walkers(W_total, p=cfg["p"], q=cfg["q"], weights=args.weights, seed=1)
    lo, hi = sharding.shard_bounds(W_total, world, rank)
    theta_h = np.ascontiguousarray(theta_all[lo:hi])
    W_local = hi - lo
    net, _ = make_net(cfg)
    o = synth.objects_from_row(node, weight, mean, theta_all[0], cfg["p"], cfg["q"], args.weights)
    S = net.set_structure(*o[:3])
    theta_d = torch.from_numpy(theta_h).to(dev)
    theta_all_d = torch.from_numpy(theta_all).to(dev)
    theta_pin = torch.from_numpy(theta_h).pin_memory()
    lib = _lib.load()

    def step_device():
        if world > 1:
            return sharding.log_likelihood_batch_sharded(net, theta_all_d)      # C ABI: local rows + ncclAllGather
        return net.log_likelihood_batch(theta_d)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            tms = torch.tensor([x], dtype=torch.float64, device=dev)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            return float(tms.item())
        return float(x)

    def timed(fn, steps, warm):
        """CUDA-event time of `steps` calls after `warm` warm-up calls, barrier + synchronize on both sides, max over ranks."""
        res = None
        for _ in range(warm):
            res = fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            res = fn()
        b.record()
        barrier()
        return max_over_ranks(a.elapsed_time(b)) / steps, res

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
    ms = max_over_ranks(e0.elapsed_time(e1))
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
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = W_total * args.steps / e2e_s
    assert np.array_equal(ll_host, result_host[lo:hi]), "host-buffer and device-buffer paths disagree"

    # ---- per-phase device times (separate profiled passes; events around every launch) ------------
    def profiled(nt, th, steps=2):
        nt.set_profiling(True)
        phase = np.zeros(6)
        nl = np.zeros(6, dtype=np.int64)
        slice_ms, slice_n = 0.0, 0
        for _ in range(steps):
            nt.log_likelihood_batch(th)
            msv, nv = nt.phase_times()
            phase += msv
            nl = nv
            sm_, slice_n = nt.slice_times()
            slice_ms += sm_
        nt.set_profiling(False)
        return phase / steps, nl, slice_ms / steps, slice_n

    phase, nl, slice_ms, slice_n = profiled(net, theta_d)
    update_path = net.update_path()
    skip = net.skip_stats()          # (dense K steps, executed K steps, panel tile halves, halves left out, right halves among them)
    p, N = cfg["p"], cfg["N"]
    peaks, peak_src = load_peaks()

    # ---- extras: the north_star numbers beyond the headline (every rank takes part; rank 0 reports) -------------------
    extra_out = {}
    if extras:
        peak_tf = np.zeros(1)
        _lib.check(lib.agpn_dmma_peak(local_rank, _lib.ptr(peak_tf)))
        dmma_peak_x = float(peak_tf[0])
        # (a) the all-FP64 factorisation: every kernel on the FP64 tensor instruction (second context)
        old = os.environ.get("AGPN_SYRK")
        os.environ["AGPN_SYRK"] = "dmma"
        try:
            net_d, _ = make_net(cfg)
            net_d._ctx()                              # the context reads AGPN_SYRK when it is created
        finally:
            if old is None:
                os.environ.pop("AGPN_SYRK", None)
            else:
                os.environ["AGPN_SYRK"] = old
        net_d.set_structure(*o[:3])
        ms_d, out_d = timed(lambda: net_d.log_likelihood_batch(theta_d), 3, 2)
        ph_d, _, sl_d, _ = profiled(net_d, theta_d)
        chol_ms_d = float(ph_d[2] + ph_d[3] + ph_d[4] + sl_d)
        chol_tf_d = W_local * p * N ** 3 / 3.0 / (chol_ms_d * 1e-3) / 1e12
        agree = rel_err(out_d.cpu().numpy(), result_host[lo:hi])
        assert net_d.update_path() == 0
        net_d.close()
        extra_out["roofline_dmma_path"] = {
            "bound": "tensor", "kernel": "potrf_diag + trsm + syrk_kernel, all on DMMA.8x8x4 (AGPN_SYRK=dmma, second context)",
            "achieved": chol_tf_d, "peak": dmma_peak_x, "unit": "TFLOP/s", "frac": chol_tf_d / dmma_peak_x,
            "flops_per_step": W_local * p * N ** 3 / 3.0, "ms": chol_ms_d, "step_ms": ms_d,
            "evals_per_s": W_total / (ms_d * 1e-3), "phase_ms": {"potrf_diag": float(ph_d[2]), "trsm": float(ph_d[3]), "syrk": float(ph_d[4])},
            "max_rel_diff_vs_default_path": agree,
            "note": "ALGORITHMIC flops p N^3 / 3 per eval over the device time of the factorisation kernels; the north_star's "
                    "'>= 50 % of FP64 tensor-core peak on the batched Cholesky' on the path that uses only the FP64 tensor instruction"}

        # (a2) the dense schedule: no K step / tile half left out (what a covariance that does not decay would cost)
        dense_extra = None
        if update_path:                      # same decision on every rank (timed() holds barriers)
            os.environ["AGPN_I8_SKIP"] = "0"
            try:
                net_n, _ = make_net(cfg)
                net_n._ctx()
            finally:
                os.environ.pop("AGPN_I8_SKIP", None)
            net_n.set_structure(*o[:3])
            ms_n, out_n = timed(lambda: net_n.log_likelihood_batch(theta_d), 5, 3)
            ph_n, _, _, _ = profiled(net_n, theta_d)
            dense_extra = {"what": "AGPN_I8_SKIP=0, second context: every K step and every panel-solve tile executed",
                           "ms_per_step": ms_n, "evals_per_s": W_total / (ms_n * 1e-3),
                           "phase_ms": {"potrf_diag": float(ph_n[2]), "trsm": float(ph_n[3]), "syrk": float(ph_n[4])},
                           "bit_identical_to_the_default_schedule": bool(np.array_equal(out_n.cpu().numpy(), result_host[lo:hi]))}
            net_n.close()
            extra_out["dense_schedule"] = dense_extra

        # (b) 16-walker single-GPU latency (the reference's own caller runs half-ensembles of 15, 01_emcee.py:111-122)
        th16 = torch.from_numpy(np.ascontiguousarray(theta_all[:16])).to(dev)
        t16, o16 = timed(lambda: net.log_likelihood_batch(th16), 10, 3)
        lat16 = {"shape": "3 x 2048, 16 walkers on ONE GPU", "ms_per_batch": t16, "evals_per_s": 16 / (t16 * 1e-3),
                 "bit_identical_to_the_headline_batch": bool(np.array_equal(o16.cpu().numpy(), result_host[:16]))}

        # (c) assembly kernels other than the headline's: generic kernel (Matern52 + RationalQuadratic nodes) and the
        #     specialised kernel with SquaredExponential weights; 32 walkers, profiled pass, parity of row 0
        asm_extra = {}
        for key, th_np, objs in (
                ("roofline_assembly_generic", synth.make_walkers_m52rq(32, p=p, seed=1),
                 lambda r: synth.objects_from_row_m52rq(node, weight, mean, r, p)),
                ("roofline_assembly_se_weights", synth.make_walkers(32, p=p, q=cfg["q"], weights="SquaredExponential", seed=1),
                 lambda r: synth.objects_from_row(node, weight, mean, r, p, cfg["q"], "SquaredExponential"))):
            ob = objs(th_np[0])
            net.set_structure(*ob[:3])
            th_x = torch.from_numpy(th_np).to(dev)
            got = net.log_likelihood_batch(th_x).cpu().numpy()
            ph_x, _, _, _ = profiled(net, th_x, 1)
            b = 32 * 4.0 * p * N * (N + 1)
            gbs = b / (ph_x[1] * 1e-3) / 1e9
            asm_extra[key] = {"bound": "hbm", "achieved": gbs, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                              "frac": gbs / peaks.get("hbm_gbs"), "bytes_per_launch": b, "ms": float(ph_x[1]),
                              "walkers": 32, "_got0": float(got[0])}
        asm_extra["roofline_assembly_generic"]["kernel"] = "assemble_kernel (any of the 14 x 15 kernel classes): nodes Matern52 + RationalQuadratic, Constant weights"
        asm_extra["roofline_assembly_se_weights"]["kernel"] = "assemble_fast_kernel<3, WSE>: QuasiPeriodic nodes, SquaredExponential weights (one exponential per series and node)"
        net.set_structure(*o[:3])
        extra_out.update(asm_extra)
        net.close()                                   # the C4 workspace needs the memory

        # (d) other BASELINE configs, walkers block-partitioned over the ranks like the headline
        other = {"C3_16_walkers_one_gpu": lat16}

        def run_ll_config(c, label, steps, warm):
            Wt = c["W"]
            th_np = synth.make_walkers(Wt, p=c["p"], q=c["q"], seed=1)
            nt, _ = make_net(c)
            ob = synth.objects_from_row(node, weight, mean, th_np[0], c["p"], c["q"])
            nt.set_structure(*ob[:3])
            th_full = torch.from_numpy(th_np).to(dev)
            l2, h2 = sharding.shard_bounds(Wt, world, rank)
            th_loc = th_full[l2:h2].contiguous()
            fn = (lambda: sharding.log_likelihood_batch_sharded(nt, th_full)) if world > 1 else (lambda: nt.log_likelihood_batch(th_loc))
            ms_c, res = timed(fn, steps, warm)
            vals = res.cpu().numpy()
            nt.close()
            return {"shape": label, "walkers_total": Wt, "walkers_per_gpu": h2 - l2, "ms_per_batch": ms_c,
                    "evals_per_s": Wt / (ms_c * 1e-3)}, vals

        other["C2"], v2 = run_ll_config(C2, "3 x 167 points, 1 QuasiPeriodic node, 256 walkers (fused small-N kernel)", 20, 3)
        other["C4"], v4 = run_ll_config(C4, "3 x 8192 points, 3 QuasiPeriodic nodes, 32 walkers", 3, 1)
        f4 = C4["W"] * C4["p"] * C4["N"] ** 3 / 3.0
        other["C4"]["cholesky_equiv_tflops_per_gpu"] = f4 / world / (other["C4"]["ms_per_batch"] * 1e-3) / 1e12
        other["C4"]["ratio_to_dmma_peak"] = other["C4"]["cholesky_equiv_tflops_per_gpu"] / dmma_peak_x

        # C5: prediction on a 10^5-point grid per series from a 3 x 4096 fitted network (grid split over the ranks)
        n5, t5 = make_net(C5)
        th5 = synth.make_walkers(1, p=C5["p"], q=C5["q"], seed=1)[0]
        o5 = synth.objects_from_row(node, weight, mean, th5, C5["p"], C5["q"])
        grid = synth.prediction_grid(t5, C5["M"])

        def predict_all():
            res = []
            with contextlib.redirect_stdout(io.StringIO()):
                for d in (1, 2, 3):
                    if world > 1:
                        res.append(sharding.prediction_sharded(n5, *o5, grid, dataset=d)[:2])
                    else:
                        res.append(n5.prediction(nodes=o5[0], weights=o5[1], means=o5[2], jitters=o5[3], time=grid, dataset=d)[:2])
            return res
        predict_all()
        barrier()
        tp0 = time.perf_counter()
        pr = predict_all()
        torch.cuda.synchronize()
        t_pred = max_over_ranks(time.perf_counter() - tp0)
        ppath = n5.predict_path()
        n5.close()
        fl5 = C5["N"] ** 2 * float(C5["M"]) + C5["N"] ** 3 / 3.0
        other["C5"] = {"shape": "prediction: 3 x 4096 fitted network, mean and std on a 100000-point grid per series",
                       "ms_per_series": 1e3 * t_pred / 3, "fp64_equiv_tflops": 3 * fl5 / t_pred / 1e12,
                       "timing": "wall clock around network.prediction (host grid in, host mean / std out: end to end)",
                       "predict_path": "factorisation kernels ([K ; K*] stacked, INT8-sliced updates)" if ppath == 1 else "predict_kernel (DMMA)"}
        extra_out["other_configs"] = other
        extra_out["_parity_inputs"] = (v2, v4, pr[0])

    if rank == 0:
        peak_tf = np.zeros(1)
        _lib.check(lib.agpn_dmma_peak(local_rank, _lib.ptr(peak_tf)))
        chol_flops = W_local * p * N ** 3 / 3.0                   # F_chol = p N^3 / 3 per eval
        chol_ms = float(phase[2] + phase[3] + phase[4] + slice_ms)
        step_prof_ms = float(phase.sum() + slice_ms)
        chol_tf = chol_flops / (chol_ms * 1e-3) / 1e12
        nblk = (N + 127) // 128
        syrk_tiles = sum((nblk - k - 1) * (nblk - k) // 2 for k in range(nblk))
        syrk_tf = W_local * p * syrk_tiles * 2.0 * 128 ** 3 / (phase[4] * 1e-3) / 1e12 if phase[4] > 0 else None
        asm_bytes = W_local * 4.0 * p * N * (N + 1)               # triangle-only store: B_tri = 4 p N (N+1)
        asm_gbs = asm_bytes / (phase[1] * 1e-3) / 1e9 if phase[1] > 0 else None
        chol_traffic = asm_traffic = syrk_traffic = trsm_traffic = None          # DRAM bytes from the committed ncu pass
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                tr = json.load(f)
            if tr.get("workload") == args.workload and args.weights == "Constant" and tr.get("update_path") == update_path:
                chol_traffic = tr["cholesky_dram_bytes_per_step"] * W_local / tr["walkers"]
                asm_traffic = tr["assembly_dram_bytes_per_launch"] * W_local / tr["walkers"]
                syrk_traffic = tr["syrk_dram_bytes_per_step"] * W_local / tr["walkers"]
                trsm_traffic = tr["trsm_dram_bytes_per_step"] * W_local / tr["walkers"]
        except Exception:
            pass
        dmma_peak = float(peak_tf[0])
        second_roofline = None
        syrk_share = float(phase[4] / max(step_prof_ms, 1e-9))
        if update_path:
            i8_peak = np.zeros(1)
            _lib.check(lib.agpn_i8_peak(local_rank, _lib.ptr(i8_peak)))
            i8_peak = float(i8_peak[0])
            # 36 slice pairs (s + t <= 7) per FP64 product, 2 ops per int8 multiply-add.  EXECUTED ops: the kernel leaves out
            # the K steps (128 x 128 x 32) in which either operand's digit chunk is all zero (exact), so the dense count is
            # scaled by the fraction of K steps the masks of this run's last wave let through (agpn_skip_stats)
            active = skip[1] / skip[0] if skip[0] > 0 else 1.0
            dense_ops = 36.0 * W_local * p * syrk_tiles * 2.0 * 128 ** 3
            exec_ops = dense_ops * active
            syrk_tops = exec_ops / (phase[4] * 1e-3) / 1e12 if phase[4] > 0 else None
            syrk_tf_exec = syrk_tf * active if syrk_tf else None
            roofline_update = {
                "bound": "tensor",
                "kernel": "syrk_i8_persistent_kernel: trailing update as exact integer products of 8 x 7-bit slices, "
                          "tcgen05.mma.kind::i8 with TMEM int32 accumulators (all launches of a step)",
                "achieved": syrk_tops, "peak": i8_peak, "unit": "TOP/s",
                "frac": (syrk_tops / i8_peak) if syrk_tops else None,
                "peak_source": "kind::i8 issue peak (M=128, N=256, operands resident in shared memory) measured in this run "
                               "(agpn_i8_peak); MEASURED_PEAKS.json has no INT8 figure (2 x its bf16 burst = %.0f TOP/s)"
                               % (2.0 * peaks.get("bf16_tflops", float("nan"))),
                "ops_per_step": exec_ops, "dense_ops_per_step": dense_ops, "active_k_step_fraction": active,
                "ms": float(phase[4]), "launches": int(nl[4]), "share_of_step_ms": syrk_share,
                "traffic": syrk_traffic,
                "traffic_note": "DRAM bytes of all syrk_i8_persistent_kernel launches of one step, ncu pass in profiles/ (scaled by walkers)",
                "fp64_equivalent": {"tflops_executed": syrk_tf_exec, "tflops_dense_schedule_equivalent": syrk_tf,
                                    "dmma_peak_tflops": dmma_peak,
                                    "note": "the same update on the FP64 tensor instruction cannot exceed the DMMA peak"},
                "note": "EXECUTED ops only (full diagonal tiles counted): K steps whose operand digit chunks are all zero are "
                        "left out -- exact, the results are bit-identical to the dense schedule (extras: dense_schedule); "
                        "CUDA-event time of the launches in a profiled pass"}
            # the panel solve: executed DMMA flops of the tile halves that were solved (right half K = 128, left half K = 64)
            halves = skip[2] / 2.0
            scale = W_local * p / (skip[2] / (2.0 * sum(nblk - k - 1 for k in range(nblk)))) if skip[2] > 0 else 1.0
            right_done, left_done = halves - skip[4], halves - (skip[3] - skip[4])
            # ALGORITHMIC flops of the triangular solve X L^T = C for the halves that were solved: column c of X costs 2 (c + 1) flops
            # per row -> right half (c = 64 .. 127) 2 x 128 x 64 x 96.5, left half (c = 0 .. 63) 2 x 128 x 64 x 32.5 (the kernel's GEMM
            # against the explicit inverse executes k extents 104 and 40 on average)
            trsm_flops = scale * (right_done * 2.0 * 128 * 64 * 96.5 + left_done * 2.0 * 128 * 64 * 32.5)
            trsm_tf = trsm_flops / (phase[3] * 1e-3) / 1e12 if phase[3] > 0 else None
            roofline_panel = {
                "bound": "tensor",
                "kernel": "trsm_kernel: A_ik Linv_kk^T on DMMA.8x8x4, digit extraction and plane stores fused (all launches of a step)",
                "achieved": trsm_tf, "peak": dmma_peak, "unit": "TFLOP/s", "frac": (trsm_tf / dmma_peak) if trsm_tf else None,
                "peak_source": "DMMA.8x8x4 issue peak measured in this run (agpn_dmma_peak); MEASURED_PEAKS.json has no FP64 figure",
                "flops_per_step": trsm_flops, "tile_halves": skip[2] * scale, "tile_halves_left_out": skip[3] * scale,
                "right_halves_solved": right_done * scale, "left_halves_solved": left_done * scale,
                "ms": float(phase[3]), "launches": int(nl[3]), "share_of_step_ms": float(phase[3] / max(step_prof_ms, 1e-9)),
                "traffic": trsm_traffic,
                "traffic_note": "DRAM bytes of all trsm_kernel launches of one step, ncu pass in profiles/ (scaled by walkers)",
                "note": "ALGORITHMIC flops of the triangular solve (2 (c + 1) per row and column c) for the tile halves that were "
                        "solved; halves whose entries are bounded below 2^-96 of their row scale are left out (AGPN_TRSM_SKIP=0: none)"}
            # `roofline` = the kernel with the largest share of the step
            roofline = roofline_panel if phase[3] > phase[4] else roofline_update
            second_roofline = ("roofline_update", roofline_update) if roofline is roofline_panel else ("roofline_panel_solve", roofline_panel)
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
            "config": config, "theta_dim": S.D,
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
            "banded_schedule": {
                "what": "work that adds nothing is not executed: K steps of the INT8-sliced update whose digit chunks are all zero "
                        "(exact), panel-solve tile halves bounded below 2^-96 of their row scale; AGPN_I8_SKIP=0 = dense schedule",
                "executed_k_step_fraction": (skip[1] / skip[0]) if skip[0] > 0 else 1.0,
                "panel_halves_left_out_fraction": (skip[3] / skip[2]) if skip[2] > 0 else 0.0,
                "bit_identical_to_dense_schedule": (extra_out.get("dense_schedule") or {}).get("bit_identical_to_the_default_schedule"),
                "dense_schedule_ms_per_step": (extra_out.get("dense_schedule") or {}).get("ms_per_step")},
            "roofline_assembly": {"bound": "hbm", "kernel": "assemble_fast_kernel (lower block triangle only; generic assemble_kernel for other kernel classes)",
                                  "achieved": asm_gbs, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                                  "frac": (asm_gbs / peaks.get("hbm_gbs")) if asm_gbs else None, "peak_source": peak_src,
                                  "bytes_per_launch": asm_bytes, "ms": float(phase[1]), "traffic": asm_traffic},
            "phase_ms": {"mean_residual": float(phase[0]), "assembly": float(phase[1]), "potrf_diag": float(phase[2]),
                         "trsm": float(phase[3]), "syrk": float(phase[4]), "finalize": float(phase[5]),
                         "digit_slicing": float(slice_ms), "launches": [int(x) for x in nl] + [int(slice_n)]},
        }
        if second_roofline is not None:
            line[second_roofline[0]] = second_roofline[1]
        if cpu_base is not None:
            ref_vals = np.array(cpu_base.pop("_vals"))
            got = result_host[:len(ref_vals)]
            line["cpu_baseline"] = cpu_base
            line["parity_vs_cpu_sample"] = {"max_rel_err": rel_err(got, ref_vals), "n": int(len(ref_vals)), "tolerance": 1e-9}
        else:
            line["cpu_baseline"] = None
        if extras:
            v2, v4, pr0 = extra_out.pop("_parity_inputs")
            g_gen = extra_out["roofline_assembly_generic"].pop("_got0")
            g_se = extra_out["roofline_assembly_se_weights"].pop("_got0")
            if cpu_extra is not None:
                r4, r5, r2, rgen, rse = cpu_extra
                oc = extra_out["other_configs"]
                oc["C4"]["parity_row0_vs_cpu_port"] = rel_err(v4[:1], r4)
                oc["C2"]["parity_rows0_3_vs_cpu_port"] = rel_err(v2[:4], r2)
                mu_g, sd_g = pr0
                mu_r, sd_r = r5
                oc["C5"]["parity_vs_cpu_port"] = {
                    "points": int(len(pred_idx)), "series": 1,
                    "mean_max_abs_err_over_max_abs_mean": float(np.max(np.abs(mu_g[pred_idx] - mu_r)) / np.max(np.abs(mu_r))),
                    "var_max_abs_err_over_max_var": float(np.max(np.abs(sd_g[pred_idx] ** 2 - sd_r ** 2)) / np.max(sd_r ** 2)),
                    "tolerance": 1e-9}
                extra_out["roofline_assembly_generic"]["parity_row0_vs_cpu_port"] = rel_err([g_gen], rgen)
                extra_out["roofline_assembly_se_weights"]["parity_row0_vs_cpu_port"] = rel_err([g_se], rse)
            line.update(extra_out)
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
