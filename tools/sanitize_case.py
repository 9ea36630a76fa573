"""Small end-to-end pass over every kernel, for compute-sanitizer (one tool per gpurun call):
    compute-sanitizer --tool memcheck python tools/sanitize_case.py
    compute-sanitizer --tool racecheck python tools/sanitize_case.py"""
import io, contextlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from artgpn_b200 import synth, node, weight, mean
from artgpn_b200.arttwo import network

def run(N, W, wk, q=2):
    t, y, sig = synth.make_data(N, 3, seed=1)
    th = synth.make_walkers(W, p=3, q=q, weights=wk, seed=2)
    net = network(q, t, y[0], sig[0], y[1], sig[1], y[2], sig[2])
    o = synth.objects_from_row(node, weight, mean, th[0], 3, q, wk)
    net.set_structure(*o[:3])
    out = net.log_likelihood_batch(th)
    assert np.all(np.isfinite(out)), out
    return net, o

variant = os.environ.get("AGPN_SYRK", "cpasync")
net, o = run(300, 3, "Constant")                   # tiled path, fast assembly, 3 block columns
net, o = run(300, 2, "SquaredExponential")         # fast assembly with SE weights
net, o = run(1400, 2, "Constant")                  # 11 block columns: banded schedule (zero-chunk masks, panel tiles left out)
st = net.skip_stats()
assert st[1] < st[0] and st[3] > 0, st
ts = np.linspace(56990, 58420, 1500)
with contextlib.redirect_stdout(io.StringIO()):
    mu, sd, _ = net.prediction(nodes=o[0], weights=o[1], means=o[2], jitters=o[3], time=ts, dataset=1)   # stacked, masks
assert np.all(np.isfinite(mu)) and np.all(np.isfinite(sd))
net, o = run(50, 3, "Constant")                    # K0 fused kernel
net.set_exact(True)
net.log_likelihood(*o)                             # exact evaluation mode
net, o = run(200, 1, "Constant")
ts = np.linspace(57000, 57210, 300)
with contextlib.redirect_stdout(io.StringIO()):
    mu, sd, _ = net.prediction(nodes=o[0], weights=o[1], means=o[2], jitters=o[3], time=ts, dataset=2)
    mu2, sd2, cov = net.prediction_cov(nodes=o[0], weights=o[1], means=o[2], jitters=o[3], time=ts[:150], dataset=2)
assert np.all(np.isfinite(mu)) and np.all(np.isfinite(sd)) and np.all(np.isfinite(cov))
# generic assembly path (non exp-family kernels) through the tiled factorisation
t, y, sig = synth.make_data(260, 2, seed=3)
net = network(2, t, y[0], sig[0], y[1], sig[1])
ll = net.log_likelihood([node.Matern52(8.0), node.RationalQuadratic(1.2, 9.0)],
                        [weight.Cosine(0.8, 30.0), weight.Constant(1.0), weight.Exponential(1.1, 12.0), weight.Matern32(0.7, 6.0)],
                        [mean.Keplerian(11.0, 0.7, 0.2, 0.4, 1.1), None], [0.5, 0.6])
assert np.isfinite(ll)
print("sanitize case ok (syrk variant %s)" % variant)
