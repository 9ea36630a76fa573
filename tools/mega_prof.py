"""Where the persistent task kernel spends its time: python tools/mega_prof.py W   (sets AGPN_MEGA_PROF=1)."""
import os, sys
os.environ["AGPN_MEGA_PROF"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from artgpn_b200 import synth, node, weight, mean, _lib
from artgpn_b200.arttwo import network
W = int(sys.argv[1])
t, y, sig = synth.make_data(2048, 3, seed=0)
theta = synth.make_walkers(W, p=3, q=2, seed=1)
net = network(2, t, y[0], sig[0], y[1], sig[1], y[2], sig[2])
net.set_structure(*synth.objects_from_row(node, weight, mean, theta[0], 3, 2)[:3])
th = torch.from_numpy(theta).cuda()
for _ in range(3):
    net.log_likelihood_batch(th)
torch.cuda.synchronize()
out = np.zeros(16)
_lib.check(_lib.load().agpn_mega_prof(net._ctx(), _lib.ptr(out)))
names = ["U runs us", "U tasks", "D body us", "D wait us", "D tasks", "P body us", "P wait us", "P tasks(g0)", "kernel us", "cluster sync us", "producer wait us"]
print("W=%d per-CTA means:" % W)
for n_, v in zip(names, out):
    print("  %-18s %10.1f" % (n_, v))
print("  U per task %.2f us, D per task %.1f us, P per task %.1f us" % (out[0] / max(out[1], 1), out[2] / max(out[4], 1), out[5] / max(out[7], 1)))
