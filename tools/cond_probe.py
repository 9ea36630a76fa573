"""How far does parity hold as the covariance becomes ill-conditioned?  (explicit-inverse TRSM vs LAPACK)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np
import artgpn_oracle as oracle
from artgpn_b200 import node, weight, mean
from artgpn_b200.arttwo import network
rng = np.random.default_rng(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
t = 57000.0 + np.sort(rng.uniform(0, 1.0 * N, N))
y = rng.normal(0, 2, (3, N))
for scale in (1.0, 1e-1, 1e-2, 1e-3, 1e-4, 1e-5):
    sig = rng.uniform(0.1, 0.3, (3, N)) * scale
    jit = [0.8 * scale, 0.37 * scale, 1.06 * scale]
    nodes = [node.QuasiPeriodic(18.56, 28.9, 0.7), node.QuasiPeriodic(60.0, 11.0, 1.2)]
    ws = [weight.Constant(c) for c in (1.99, 4.53, 6.94, 0.7, 1.1, 2.0)]
    ms = [mean.Linear(0.0, 0.0)] * 3
    net = network(2, t, y[0], sig[0], y[1], sig[1], y[2], sig[2])
    got = net.log_likelihood(nodes, ws, ms, jit)
    net.set_exact(True)
    got_exact = net.log_likelihood(nodes, ws, ms, jit)
    sp = lambda k: (type(k).__name__, list(k.pars))
    exp = oracle.log_likelihood(t, y, sig, 2, [sp(k) for k in nodes], [sp(k) for k in ws], [sp(k) for k in ms], jit)
    K = oracle.covariance_block(t, [sp(k) for k in nodes], [sp(k) for k in ws], 2, 2)
    K[np.diag_indices(N)] += sig[2] ** 2 + jit[2] ** 2
    Kd = net._covariance_matrix(nodes, ws, t, 3)
    Ko = oracle.covariance_block(t, [sp(k) for k in nodes], [sp(k) for k in ws], 2, 2)
    print("noise scale %g: cond(K_3) = %.2e  ll = %.10g  |rel diff vs LAPACK| fast %.2e  exact %.2e   max entry rel diff (exact mode) %.1e" % (
        scale, np.linalg.cond(K), exp, abs(got - exp) / abs(exp), abs(got_exact - exp) / abs(exp), np.max(np.abs(Kd - Ko) / np.abs(Ko))))
