"""Concurrency picture of one log_likelihood_batch: python tools/timeline.py W out.csv   (AGPN_TIMELINE=1 is set here)."""
import os, sys
os.environ["AGPN_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from artgpn_b200 import synth, node, weight, mean, _lib
from artgpn_b200.arttwo import network
W, out = int(sys.argv[1]), sys.argv[2]
t, y, sig = synth.make_data(2048, 3, seed=0)
theta = synth.make_walkers(W, p=3, q=2, seed=1)
net = network(2, t, y[0], sig[0], y[1], sig[1], y[2], sig[2])
net.set_structure(*synth.objects_from_row(node, weight, mean, theta[0], 3, 2)[:3])
for _ in range(3):
    net.log_likelihood_batch(theta)
_lib.check(_lib.load().agpn_timeline_dump(net._ctx(), out.encode()))
rows = np.loadtxt(out, delimiter=",", skiprows=1)
names = ["mean", "asm", "potrf", "trsm", "syrk", "final", "slice"]
print("W=%d: %d launches, span %.3f ms" % (W, len(rows), rows[:, 4].max()))
for ph in range(7):
    r = rows[rows[:, 1] == ph]
    if len(r):
        print("  %-6s n=%3d  sum of durations %.3f ms  mean %.1f us" % (names[ph], len(r), (r[:, 4] - r[:, 3]).sum(), 1e3 * (r[:, 4] - r[:, 3]).mean()))
