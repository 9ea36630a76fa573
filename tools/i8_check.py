"""INT8-sliced trailing update vs the DMMA path on the same inputs (GPU): log-likelihood agreement and timing.
usage: python tools/i8_check.py            (contexts read AGPN_SYRK when they are created)"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from artgpn_b200 import synth, node, weight, mean
from artgpn_b200.arttwo import network


def make(N, W, mode, q=2, group=None):
    if mode:
        os.environ["AGPN_SYRK"] = mode
    else:
        os.environ.pop("AGPN_SYRK", None)
    if group:
        os.environ["AGPN_GROUP"] = str(group)
    else:
        os.environ.pop("AGPN_GROUP", None)
    t, y, sig = synth.make_data(N, 3, seed=0)
    theta = synth.make_walkers(W, p=3, q=q, weights="Constant", seed=1)
    net = network(q, t, y[0], sig[0], y[1], sig[1], y[2], sig[2])
    o = synth.objects_from_row(node, weight, mean, theta[0], 3, q, "Constant")
    net.set_structure(*o[:3])
    return net, torch.from_numpy(theta).cuda()


def timeit(net, th, steps=3):
    for _ in range(3):
        out = net.log_likelihood_batch(th)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = net.log_likelihood_batch(th)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


for N, W in ((300, 6), (1000, 6), (2048, 8)):
    ref_net, th = make(N, W, None)
    ref = ref_net.log_likelihood_batch(th).cpu().numpy()
    for mode in ("i8", "i8_wide", "i8_narrow"):
        for group in (None, 2, 3):
            net, th2 = make(N, W, mode, group=group)
            got = net.log_likelihood_batch(th2).cpu().numpy()
            got2 = net.log_likelihood_batch(th2).cpu().numpy()
            rel = np.max(np.abs(got - ref) / np.abs(ref))
            print("N=%d W=%d %s group=%s: max rel diff vs dmma %.3e  bit-reproducible %s  ll[0]=%.15g ref %.15g" % (
                N, W, mode, group, rel, bool(np.all(got == got2)), got[0], ref[0]), flush=True)

if len(sys.argv) > 1 and sys.argv[1] == "time":
    N, W = 2048, 128
    for mode in (None, "i8", "i8_wide"):
        for group in (8, 4, 2):
            net, th = make(N, W, mode, group=group)
            # warm the clocks up
            t0 = time.perf_counter()
            while time.perf_counter() - t0 < 0.5:
                net.log_likelihood_batch(th)
                torch.cuda.synchronize()
            ms = timeit(net, th, steps=8)
            print("time N=%d W=%d mode=%s group=%d: %.2f ms/step %.1f evals/s" % (N, W, mode, group, ms, W / ms * 1e3), flush=True)
            del net
            torch.cuda.empty_cache()

if len(sys.argv) > 1 and sys.argv[1] == "prof":
    N, W = 2048, 128
    for mode, group in ((None, 8), ("i8", 8), ("i8", 16), ("i8", 4), ("i8_wide", 4)):
        net, th = make(N, W, mode, group=group)
        for _ in range(3):
            net.log_likelihood_batch(th)
        torch.cuda.synchronize()
        net.set_profiling(True)
        net.log_likelihood_batch(th)
        ms, n = net.phase_times()
        sm, sn = net.slice_times()
        net.set_profiling(False)
        ms_step = timeit(net, th, steps=8)
        print("prof mode=%s group=%d: step %.2f ms | mean %.2f asm %.2f diag %.2f trsm %.2f syrk %.2f final %.2f slice %.2f (launches %s, slice %d)" % (
            mode, group, ms_step, ms[0], ms[1], ms[2], ms[3], ms[4], ms[5], sm, n.tolist(), sn), flush=True)
        del net
        torch.cuda.empty_cache()

if len(sys.argv) > 1 and sys.argv[1] == "fuse":
    for N, W, group in ((300, 5, None), (1000, 6, None), (2048, 8, None), (2048, 8, 4), (1000, 6, 3)):
        res = {}
        for fuse in ("0", "1"):
            os.environ["AGPN_I8_FUSE"] = fuse
            net, th = make(N, W, "i8", group=group)
            res[fuse] = net.log_likelihood_batch(th).cpu().numpy()
        ref_net, th = make(N, W, "dmma")
        ref = ref_net.log_likelihood_batch(th).cpu().numpy()
        print("N=%d W=%d group=%s: fused == separate slicing: %s ; max rel diff vs dmma %.3e" % (
            N, W, group, bool(np.array_equal(res["0"], res["1"])), np.max(np.abs(res["1"] - ref) / np.abs(ref))), flush=True)
    for fuse in ("0", "1"):
        os.environ["AGPN_I8_FUSE"] = fuse
        for W in (128, 16):
            net, th = make(2048, W, "i8")
            t0 = time.perf_counter()
            while time.perf_counter() - t0 < 0.5:
                net.log_likelihood_batch(th)
                torch.cuda.synchronize()
            ms = timeit(net, th, steps=8)
            net.set_profiling(True)
            net.log_likelihood_batch(th)
            pm, n = net.phase_times()
            sm, sn = net.slice_times()
            net.set_profiling(False)
            print("fuse=%s W=%d: %.3f ms/step %.1f evals/s | asm %.2f diag %.2f trsm %.2f syrk %.2f slice %.2f" % (
                fuse, W, ms, W / ms * 1e3, pm[1], pm[2], pm[3], pm[4], sm), flush=True)
            del net
            torch.cuda.empty_cache()

if len(sys.argv) > 1 and sys.argv[1] == "persist":
    for N, W, group in ((300, 5, None), (1000, 6, None), (2048, 8, None), (2048, 8, 4), (1000, 6, 3), (2048, 40, None)):
        res = {}
        for pe in ("0", "1"):
            os.environ["AGPN_I8_PERSIST"] = pe
            net, th = make(N, W, "i8", group=group)
            res[pe] = net.log_likelihood_batch(th).cpu().numpy()
            again = net.log_likelihood_batch(th).cpu().numpy()
            assert np.array_equal(again, res[pe])
        print("N=%d W=%d group=%s: persistent == per-tile kernel: %s  ll[0]=%.15g" % (
            N, W, group, bool(np.array_equal(res["0"], res["1"])), res["1"][0]), flush=True)
    for pe in ("0", "1"):
        os.environ["AGPN_I8_PERSIST"] = pe
        for N, W in ((2048, 128), (2048, 16), (8192, 4)):
            net, th = make(N, W, "i8", q=3 if N == 8192 else 2)
            t0 = time.perf_counter()
            while time.perf_counter() - t0 < 0.5:
                net.log_likelihood_batch(th)
                torch.cuda.synchronize()
            ms = timeit(net, th, steps=8 if N == 2048 else 2)
            net.set_profiling(True)
            net.log_likelihood_batch(th)
            pm, n = net.phase_times()
            net.set_profiling(False)
            print("persist=%s N=%d W=%d: %.3f ms/step %.1f evals/s | asm %.2f diag %.2f trsm %.2f syrk %.2f" % (
                pe, N, W, ms, W / ms * 1e3, pm[1], pm[2], pm[3], pm[4]), flush=True)
            del net
            torch.cuda.empty_cache()
