"""GPU: analytic gradient of the log-likelihood (agpn_loglike_grad) against central finite differences of the CPU oracle
and of the GPU likelihood itself."""
import numpy as np
import pytest

import helpers as H
from helpers import oracle

pytestmark = pytest.mark.gpu


def _objs(mod_node, mod_weight, mod_mean, th):
    """theta row -> objects of the test network: nodes QuasiPeriodic + Matern52, weights mixed, means Linear / Constant / None."""
    th = list(map(float, th))
    nodes = [mod_node.QuasiPeriodic(th[0], th[1], th[2]), mod_node.Matern52(th[3])]
    ws = [mod_weight.Constant(th[4]), mod_weight.SquaredExponential(th[5], th[6]),
          mod_weight.Matern32(th[7], th[8]), mod_weight.Constant(th[9]),
          mod_weight.Periodic(th[10], th[11], th[12]), mod_weight.Exponential(th[13], th[14])]
    means = [mod_mean.Linear(th[15], th[16]), mod_mean.Constant(th[17]), None]
    return nodes, ws, means, th[18:21]


THETA0 = np.array([25.0, 17.0, 0.9, 12.0, 1.5, 0.8, 30.0, 1.1, 9.0, 0.7, 0.9, 14.0, 1.3, 0.6, 20.0,
                   0.01, 0.3, -0.2, 0.7, 0.5, 0.9])


class _Spec(object):
    def __getattr__(self, name):
        return lambda *pars: (name, list(pars))


def _oracle_ll(t, y, sig, th):
    m = _Spec()
    nodes, ws, means, jit = _objs(m, m, m, th)
    return oracle.log_likelihood(t, y, sig, 2, nodes, ws, means, jit)


@pytest.mark.parametrize("N", [96, 300])
def test_gradient_matches_finite_differences_of_the_oracle(N):
    from artgpn_b200 import synth, node, weight, mean
    from artgpn_b200.arttwo import network
    t, y, sig = synth.make_data(N, 3, seed=11)
    net = network(2, t, y[0], sig[0], y[1], sig[1], y[2], sig[2])
    nodes, ws, means, jit = _objs(node, weight, mean, THETA0)
    ll, grad = net.log_likelihood_grad(nodes, ws, means, jit)
    s = net._structure
    assert grad.shape == (s.D,) and s.D == THETA0.size
    assert np.array_equal(s.pack(nodes, ws, means, jit), THETA0)
    assert abs(ll - _oracle_ll(t, y, sig, THETA0)) <= 1e-9 * abs(ll)
    assert abs(ll - net.log_likelihood(nodes, ws, means, jit)) <= 1e-10 * abs(ll)
    fd = np.zeros_like(THETA0)
    for k in range(THETA0.size):
        h = 1e-5 * max(abs(THETA0[k]), 0.1)
        tp, tm = THETA0.copy(), THETA0.copy()
        tp[k] += h
        tm[k] -= h
        fd[k] = (_oracle_ll(t, y, sig, tp) - _oracle_ll(t, y, sig, tm)) / (2 * h)
    scale = np.maximum(np.abs(fd), 1e-3 * np.max(np.abs(fd)))
    assert np.max(np.abs(grad - fd) / scale) <= 2e-5, (grad, fd)
    net.close()


def test_gradient_on_the_bench_structure_block_size():
    """N = 600 (five block columns: the augmented system runs through the tiled factorisation kernels), the bench's
    structure (two QuasiPeriodic nodes, Constant weights, Linear means), finite differences of the GPU likelihood."""
    from artgpn_b200 import synth, node, weight, mean
    from artgpn_b200.arttwo import network
    t, y, sig = synth.make_data(600, 3, seed=5)
    theta = synth.make_walkers(1, p=3, q=2, seed=2)[0]
    net = network(2, t, y[0], sig[0], y[1], sig[1], y[2], sig[2])
    o = synth.objects_from_row(node, weight, mean, theta, 3, 2)
    ll, grad = net.log_likelihood_grad(*o)
    net.set_structure(*o[:3])
    rows = []
    hs = []
    for k in range(theta.size):
        h = 1e-5 * max(abs(theta[k]), 0.1)
        for sgn in (1, -1):
            r = theta.copy()
            r[k] += sgn * h
            rows.append(r)
        hs.append(h)
    vals = net.log_likelihood_batch(np.array(rows))
    fd = (vals[0::2] - vals[1::2]) / (2 * np.array(hs))
    scale = np.maximum(np.abs(fd), 1e-3 * np.max(np.abs(fd)))
    assert np.max(np.abs(grad - fd) / scale) <= 5e-5, (grad, fd)
    assert abs(ll - net.log_likelihood_batch(theta[None, :])[0]) <= 1e-10 * abs(ll)
    net.close()


def test_gradient_rejects_unsupported_classes():
    from artgpn_b200 import synth, node, weight, mean, _lib
    from artgpn_b200.arttwo import network
    t, y, sig = synth.make_data(50, 1, seed=1)
    net = network(1, t, y[0], sig[0])
    with pytest.raises(_lib.AgpnError):
        net.log_likelihood_grad([node.RationalQuadratic(1.5, 10.0)], [weight.Constant(1.0)], [None], [0.5])
    with pytest.raises(_lib.AgpnError):
        net.log_likelihood_grad([node.SquaredExponential(10.0)], [weight.Constant(1.0)], [mean.Sine(1.0, 9.0, 0.1, 0.0)], [0.5])
    ll, g = net.log_likelihood_grad([node.SquaredExponential(10.0)], [weight.Constant(1.0)], None, [0.5])
    assert np.isfinite(ll) and g.shape == (3,)
    net.close()
