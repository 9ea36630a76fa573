"""A/B of the prediction path (BASELINE config 5 shape) with / without zero-chunk skipping: wall time per series."""
import os, sys, time, io, contextlib
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from artgpn_b200 import synth, node, weight, mean
from artgpn_b200.arttwo import network
N, p, q, M = 4096, 3, 2, 100000
t, y, sig = synth.make_data(N, p, seed=0)
theta = synth.make_walkers(1, p=p, q=q, seed=1)
grid = synth.prediction_grid(t, M) if hasattr(synth, "prediction_grid") else np.linspace(t.min(), t.max(), M)
out = {}
for env in ({}, {"AGPN_TRSM_SKIP": "0"}, {"AGPN_I8_SKIP": "0"}):
    for k in ("AGPN_I8_SKIP", "AGPN_TRSM_SKIP"): os.environ.pop(k, None)
    os.environ.update(env)
    net = network(q, t, y[0], sig[0], y[1], sig[1], y[2], sig[2])
    o = synth.objects_from_row(node, weight, mean, theta[0], p, q)
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            mu, sd, _ = net.prediction(nodes=o[0], weights=o[1], means=o[2], jitters=o[3], time=grid, dataset=1)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(env, "%.2f ms per series, path %d" % (dt * 1e3, net.predict_path()), flush=True)
    out[str(env)] = (mu, sd)
    net.close()
ks = list(out)
for k in ks[1:]:
    print(k, "max |dmu| / max|mu| = %.3g, max |dvar| / max var = %.3g" % (
        np.max(np.abs(out[k][0] - out[ks[0]][0])) / np.max(np.abs(out[ks[0]][0])),
        np.max(np.abs(out[k][1] ** 2 - out[ks[0]][1] ** 2)) / np.max(out[ks[0]][1] ** 2)))
