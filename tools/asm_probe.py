import os, sys, json
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from artgpn_b200 import synth, node, weight, mean
from artgpn_b200.arttwo import network
t, y, sig = synth.make_data(2048, 3, seed=0)
net = network(2, t, y[0], sig[0], y[1], sig[1], y[2], sig[2])
th = synth.make_walkers_m52rq(32, p=3, seed=1)
net.set_structure(*synth.objects_from_row_m52rq(node, weight, mean, th[0], 3)[:3])
x = torch.from_numpy(th).cuda()
got = net.log_likelihood_batch(x).cpu().numpy()
net.set_profiling(True)
net.log_likelihood_batch(x); ms, _ = net.phase_times()
b = 32 * 4.0 * 3 * 2048 * 2049
print(os.environ.get("AGPN_LIB", "default"), os.environ.get("AGPN_GENERIC_ASM", ""), "generic asm ms %.3f  GB/s %.0f  frac %.3f  ll0 %.12g" % (ms[1], b / ms[1] / 1e6, b / ms[1] / 1e6 / 6534.1, got[0]))
