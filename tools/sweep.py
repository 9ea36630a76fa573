"""Quick device-only timing of log_likelihood_batch for tuning (not the bench): prints evals/s for
the environment's AGPN_GROUP / AGPN_STREAMS.  usage: python tools/sweep.py N W [steps]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from artgpn_b200 import synth, node, weight, mean
from artgpn_b200.arttwo import network

N, W = int(sys.argv[1]), int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
q = int(os.environ.get("SWEEP_Q", "2"))
wk = os.environ.get("SWEEP_WEIGHTS", "Constant")
t, y, sig = synth.make_data(N, 3, seed=0)
theta = synth.make_walkers(W, p=3, q=q, weights=wk, seed=1)
net = network(q, t, y[0], sig[0], y[1], sig[1], y[2], sig[2])
o = synth.objects_from_row(node, weight, mean, theta[0], 3, q, wk)
net.set_structure(*o[:3])
th = torch.from_numpy(theta).cuda()
import time
t_w = time.perf_counter()
while time.perf_counter() - t_w < 0.4:          # the GPU idles at 120 MHz: warm up until clocks are up
    for _ in range(3):
        out = net.log_likelihood_batch(th)
    torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ms = 0.0
while ms * steps < 300.0:                        # time at least 0.3 s
    e0.record()
    for _ in range(steps):
        out = net.log_likelihood_batch(th)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if ms * steps < 300.0:
        steps = int(steps * max(2.0, 350.0 / max(ms * steps, 1e-3)))
print("N=%d W=%d q=%d %s group=%s streams=%s : %.3f ms/step  %.1f evals/s  chol-equivalent %.2f TF/s  ll[0]=%.12g" % (
    N, W, q, wk, os.environ.get("AGPN_GROUP", "def"), os.environ.get("AGPN_STREAMS", "def"), ms, W / ms * 1e3,
    W * 3 * N ** 3 / 3 / (ms * 1e-3) / 1e12, float(out[0])))
