"""Per-phase device times (profiled, serialised pass) and executed-work statistics for a few workload shapes."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from artgpn_b200 import synth, node, weight, mean
from artgpn_b200.arttwo import network
names = ["mean", "asm", "potrf", "trsm", "syrk", "final"]
shapes = [(2048, 3, 2, 128), (2048, 3, 2, 16), (8192, 3, 3, 32)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for N, p, q, W in shapes:
    t, y, sig = synth.make_data(N, p, seed=0)
    theta = synth.make_walkers(W, p=p, q=q, seed=1)
    args = []
    for i in range(p):
        args += [y[i], sig[i]]
    net = network(q, t, *args)
    net.set_structure(*synth.objects_from_row(node, weight, mean, theta[0], p, q)[:3])
    th = torch.from_numpy(theta).cuda()
    for _ in range(3):
        net.log_likelihood_batch(th)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        net.log_likelihood_batch(th)
    b.record(); torch.cuda.synchronize()
    net.set_profiling(True)
    net.log_likelihood_batch(th)
    ms, n = net.phase_times()
    net.set_profiling(False)
    st = net.skip_stats()
    print("N=%d p=%d q=%d W=%d: %.3f ms/step | " % (N, p, q, W, a.elapsed_time(b) / 5) +
          " ".join("%s %.3f(%d)" % (names[i], ms[i], n[i]) for i in range(6)) +
          " | active K steps %.4f, panel halves left out %.4f" % (st[1] / max(st[0], 1), st[3] / max(st[2], 1)), flush=True)
    net.close()
