"""A/B of the zero-chunk skipping knobs at the BASELINE config-3 shape: ms per step, executed-work fractions, bit identity."""
import os, sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import torch
from artgpn_b200 import synth, node, weight, mean
from artgpn_b200.arttwo import network
N, p, q = 2048, 3, 2
t, y, sig = synth.make_data(N, p, seed=0)
def run(W, env):
    for k in ("AGPN_I8_SKIP", "AGPN_TRSM_SKIP", "AGPN_TILE_BOUNDS"): os.environ.pop(k, None)
    os.environ.update(env)
    theta = synth.make_walkers(W, p=p, q=q, seed=1)
    net = network(q, t, y[0], sig[0], y[1], sig[1], y[2], sig[2])
    net.set_structure(*synth.objects_from_row(node, weight, mean, theta[0], p, q)[:3])
    for _ in range(4): out = net.log_likelihood_batch(theta)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): out = net.log_likelihood_batch(theta)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20 * 1e3
    st = net.skip_stats()
    net.set_profiling(True) if hasattr(net, 'set_profiling') else None
    net.close()
    return dt, st, out
ref = None
for env in ({}, {"AGPN_TILE_BOUNDS": "0"}, {"AGPN_TRSM_SKIP": "0"}, {"AGPN_I8_SKIP": "0"}):
    for W in (128, 16):
        dt, st, out = run(W, env)
        print(env, "W=%d %.3f ms" % (W, dt), "active %.4f skipped halves %.4f (right %.0f)" % (st[1] / st[0], st[3] / st[2], st[4]), flush=True)
        if W == 128:
            if ref is None: ref = out
            print("  identical to first:", np.array_equal(ref, out), flush=True)
