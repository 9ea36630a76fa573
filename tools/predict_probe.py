"""Timing / sanity probe of network.prediction at BASELINE config 5 shape (3 x N fitted network,
M-point grid per series).  usage: python tools/predict_probe.py N M"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from artgpn_b200 import synth, node, weight, mean
from artgpn_b200.arttwo import network
N, M = int(sys.argv[1]), int(sys.argv[2])
t, y, sig = synth.make_data(N, 3, seed=0)
theta = synth.make_walkers(1, p=3, q=2, seed=1)
net = network(2, t, y[0], sig[0], y[1], sig[1], y[2], sig[2])
nodes, ws, ms, jit = synth.objects_from_row(node, weight, mean, theta[0], 3, 2)
ts = np.linspace(t.min() - 5, t.max() + 5, M)
import io, contextlib
for rep in range(2):
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        out = [net.prediction(nodes=nodes, weights=ws, means=ms, jitters=jit, time=ts, dataset=d) for d in (1, 2, 3)]
    dt = time.perf_counter() - t0
    print("N=%d M=%d: 3 series in %.3f s (%.1f ms/series), trsm-equivalent %.2f TF/s" % (N, M, dt, dt / 3 * 1e3, 3 * (N * N * M + N ** 3 / 3) / dt / 1e12))
mu, sd, _ = out[0]
print("finite:", np.isfinite(mu).all(), np.isfinite(sd).all(), "mu[:3]", mu[:3], "sd[:3]", sd[:3])
if N <= 1024 and M <= 4000:
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import artgpn_oracle as oracle
    class Mm(object):
        def __getattr__(self, name): return lambda *pars: (name, list(pars))
    m = Mm()
    o = synth.objects_from_row(m, m, m, theta[0], 3, 2)
    emu, esd = oracle.prediction(t, y, sig, 2, *o, ts, 1)
    print("vs oracle: mean", np.max(np.abs(mu - emu)) / np.max(np.abs(emu)), "var", np.max(np.abs(sd**2 - esd**2)) / np.max(esd**2))
