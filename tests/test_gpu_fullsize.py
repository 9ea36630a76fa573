"""GPU: BASELINE.json sizes (3 x 2048, 2 nodes) -- one oracle evaluation plus size-independent
properties (walker-order invariance, batch-composition invariance, jitter monotonicity)."""
import numpy as np
import pytest

import helpers as H
from helpers import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c3():
    from artgpn_b200 import synth, node, weight, mean
    from artgpn_b200.arttwo import network
    t, y, sig = synth.make_data(2048, 3, seed=0)
    theta = synth.make_walkers(12, p=3, q=2, weights="Constant", seed=1)
    net = network(2, t, y[0], sig[0], y[1], sig[1], y[2], sig[2])
    o = synth.objects_from_row(node, weight, mean, theta[0], 3, 2, "Constant")
    net.set_structure(*o[:3])
    return net, (t, y, sig), theta


def test_c3_one_walker_vs_oracle(c3):
    from artgpn_b200 import synth
    net, (t, y, sig), theta = c3

    class M(object):
        def __getattr__(self, name):
            return lambda *pars: (name, list(pars))
    m = M()
    o = synth.objects_from_row(m, m, m, theta[0], 3, 2, "Constant")
    exp = oracle.log_likelihood(t, y, sig, 2, *o)
    got = net.log_likelihood_batch(theta[:1])[0]
    assert abs(got - exp) <= 1e-9 * abs(exp), (got, exp)


def test_c3_batch_invariances(c3):
    net, _, theta = c3
    full = net.log_likelihood_batch(theta)
    assert np.all(np.isfinite(full))
    perm = np.random.default_rng(0).permutation(theta.shape[0])
    assert np.array_equal(net.log_likelihood_batch(theta[perm].copy()), full[perm])      # order
    assert np.array_equal(net.log_likelihood_batch(theta[3:7].copy()), full[3:7])         # composition
    again = net.log_likelihood_batch(theta)
    assert np.array_equal(again, full)                                                   # run-to-run


def test_c3_se_weights_one_walker_vs_oracle():
    from artgpn_b200 import synth, node, weight, mean
    from artgpn_b200.arttwo import network
    t, y, sig = synth.make_data(1024, 3, seed=0)
    theta = synth.make_walkers(2, p=3, q=2, weights="SquaredExponential", seed=1)
    net = network(2, t, y[0], sig[0], y[1], sig[1], y[2], sig[2])
    o = synth.objects_from_row(node, weight, mean, theta[1], 3, 2, "SquaredExponential")
    got = net.log_likelihood(*o)

    class M(object):
        def __getattr__(self, name):
            return lambda *pars: (name, list(pars))
    m = M()
    exp = oracle.log_likelihood(t, y, sig, 2, *synth.objects_from_row(m, m, m, theta[1], 3, 2, "SquaredExponential"))
    assert abs(got - exp) <= 1e-9 * abs(exp), (got, exp)


def test_ragged_3000_points_three_groups_vs_oracle():
    """N = 3000 (24 block columns = three groups of 8, ragged last tile), 2 walkers, SE weights."""
    from artgpn_b200 import synth, node, weight, mean
    from artgpn_b200.arttwo import network
    t, y, sig = synth.make_data(3000, 2, seed=3)
    theta = synth.make_walkers(2, p=2, q=2, weights="SquaredExponential", seed=4)
    net = network(2, t, y[0], sig[0], y[1], sig[1])
    net.set_structure(*synth.objects_from_row(node, weight, mean, theta[0], 2, 2, "SquaredExponential")[:3])
    got = net.log_likelihood_batch(theta)

    class M(object):
        def __getattr__(self, name):
            return lambda *pars: (name, list(pars))
    m = M()
    for w in range(2):
        exp = oracle.log_likelihood(t, y, sig, 2, *synth.objects_from_row(m, m, m, theta[w], 2, 2, "SquaredExponential"))
        assert abs(got[w] - exp) <= 1e-9 * abs(exp), (w, got[w], exp)


def test_c3_update_paths_agree_and_match_oracle(monkeypatch):
    """BASELINE size: INT8-sliced update (default left-looking, groups of 4 with wide updates) and the FP64
    tensor instruction give the same log-likelihoods to round-off, each within 1e-9 of the oracle."""
    from artgpn_b200 import synth, node, weight, mean
    from artgpn_b200.arttwo import network
    t, y, sig = synth.make_data(2048, 3, seed=0)
    theta = synth.make_walkers(3, p=3, q=2, weights="Constant", seed=1)

    class M(object):
        def __getattr__(self, name):
            return lambda *pars: (name, list(pars))
    m = M()
    exp = oracle.log_likelihood(t, y, sig, 2, *synth.objects_from_row(m, m, m, theta[2], 3, 2, "Constant"))
    got = {}
    for mode, group, path in (("i8", None, 3), ("i8", "4", 3), ("i8_wide", "4", 2), ("dmma", None, 0)):
        monkeypatch.setenv("AGPN_SYRK", mode)
        if group:
            monkeypatch.setenv("AGPN_GROUP", group)
        else:
            monkeypatch.delenv("AGPN_GROUP", raising=False)
        net = network(2, t, y[0], sig[0], y[1], sig[1], y[2], sig[2])
        net.set_structure(*synth.objects_from_row(node, weight, mean, theta[0], 3, 2, "Constant")[:3])
        assert net.update_path() == path
        got[(mode, group)] = net.log_likelihood_batch(theta)
        assert np.array_equal(net.log_likelihood_batch(theta), got[(mode, group)])
        net.close()
    # the panel solve writes the digit planes itself (default); the separate slicing kernel gives the same bits
    monkeypatch.setenv("AGPN_SYRK", "i8")
    monkeypatch.delenv("AGPN_GROUP", raising=False)
    monkeypatch.setenv("AGPN_I8_FUSE", "0")
    net = network(2, t, y[0], sig[0], y[1], sig[1], y[2], sig[2])
    net.set_structure(*synth.objects_from_row(node, weight, mean, theta[0], 3, 2, "Constant")[:3])
    assert np.array_equal(net.log_likelihood_batch(theta), got[("i8", None)])
    net.close()
    monkeypatch.delenv("AGPN_I8_FUSE", raising=False)
    # persistent clusters walking over the tiles (default) vs one CTA pair per tile: same bits, also with wide updates
    for group in (None, "4"):
        monkeypatch.setenv("AGPN_I8_PERSIST", "0")
        if group:
            monkeypatch.setenv("AGPN_GROUP", group)
        else:
            monkeypatch.delenv("AGPN_GROUP", raising=False)
        net = network(2, t, y[0], sig[0], y[1], sig[1], y[2], sig[2])
        net.set_structure(*synth.objects_from_row(node, weight, mean, theta[0], 3, 2, "Constant")[:3])
        assert np.array_equal(net.log_likelihood_batch(theta), got[("i8", group)])
        net.close()
    monkeypatch.delenv("AGPN_I8_PERSIST", raising=False)
    monkeypatch.delenv("AGPN_GROUP", raising=False)
    # look-ahead (diagonal tile updated and factorised while a side stream updates the tiles below it): same bits
    for la in ("2", "0"):
        monkeypatch.setenv("AGPN_LOOKAHEAD", la)
        net = network(2, t, y[0], sig[0], y[1], sig[1], y[2], sig[2])
        net.set_structure(*synth.objects_from_row(node, weight, mean, theta[0], 3, 2, "Constant")[:3])
        big = np.concatenate([theta] * 6)                    # 18 walkers = 54 matrices: the sub-batch streams are in use
        out = net.log_likelihood_batch(big)
        assert np.array_equal(out[:3], got[("i8", None)]) and np.array_equal(out[15:], got[("i8", None)])
        net.close()
    monkeypatch.delenv("AGPN_LOOKAHEAD", raising=False)
    ref = got[("dmma", None)]
    for k, v in got.items():
        assert abs(v[2] - exp) <= 1e-9 * abs(exp), (k, v[2], exp)
        assert H.rel_err(v, ref) <= 1e-12, (k, v, ref)
