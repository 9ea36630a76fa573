"""GPU: edge cases of the boundary -- limits, structure switching, several contexts, non-default
streams, memory-sized waves -- each checked against the oracle."""
import numpy as np
import pytest

import helpers as H
from helpers import oracle

pytestmark = pytest.mark.gpu
LL_RTOL = 1e-9


class _M(object):
    def __getattr__(self, name):
        return lambda *pars: (name, list(pars))


def _oracle_ll(t, y, sig, q, nodes, weights, means, jit):
    sp = lambda k: None if k is None else (type(k).__name__, list(k.pars))
    return oracle.log_likelihood(t, y, sig, q, [sp(k) for k in nodes], [sp(k) for k in weights],
                                 None if means is None else [sp(k) for k in means], jit)


@pytest.mark.parametrize("p,q,N", [(1, 1, 40), (1, 8, 33), (8, 1, 21), (4, 3, 200), (2, 5, 300)])
def test_shape_limits_vs_oracle(p, q, N):
    """1..8 series, 1..8 nodes (limits of the C ABI), both the fused and the tiled size range."""
    from artgpn_b200 import node, weight, mean
    from artgpn_b200.arttwo import network
    rng = np.random.default_rng(100 * p + q)
    t = np.sort(rng.uniform(0, 2.0 * N, N)) + 5000.0
    y = rng.normal(0, 1, (p, N))
    sig = rng.uniform(0.1, 0.3, (p, N))
    kinds = [lambda: node.QuasiPeriodic(rng.uniform(10, 30), rng.uniform(8, 25), rng.uniform(0.5, 1.2)),
             lambda: node.SquaredExponential(rng.uniform(3, 20)), lambda: node.Matern32(rng.uniform(3, 20)),
             lambda: node.Periodic(rng.uniform(8, 25), rng.uniform(0.5, 1.2)), lambda: node.Exponential(rng.uniform(3, 9))]
    nodes = [kinds[j % len(kinds)]() for j in range(q)]
    weights = [weight.Constant(rng.uniform(0.3, 1.5)) if (i + j) % 2 else weight.SquaredExponential(rng.uniform(0.3, 1.5), rng.uniform(20, 60))
               for i in range(p) for j in range(q)]
    means = [mean.Linear(rng.normal(0, 0.01), rng.normal(0, 0.3)) if i % 2 == 0 else None for i in range(p)]
    jit = list(rng.uniform(0.2, 0.8, p))
    args = []
    for i in range(p):
        args += [y[i], sig[i]]
    net = network(q, t, *args)
    got = net.log_likelihood(nodes, weights, means, jit)
    exp = _oracle_ll(t, y, sig, q, nodes, weights, means, jit)
    assert abs(got - exp) <= LL_RTOL * abs(exp), (got, exp)


def test_too_many_series_or_nodes_is_an_error():
    from artgpn_b200 import AgpnError, node, weight
    from artgpn_b200.arttwo import network
    t = np.linspace(0, 10, 12)
    args = []
    for i in range(9):
        args += [np.zeros(12), np.ones(12)]
    net = network(1, t, *args)
    with pytest.raises(AgpnError):
        net.log_likelihood([node.Constant(1.0)], [weight.Constant(1.0)] * 9, None, [0.1] * 9)


def test_structure_switching_and_two_contexts():
    """The reference re-reads the kernel classes on every call: alternate structures on one network (same
    q, different kernel / mean classes) and interleave two networks (two device contexts) -- every value
    must match its golden / the oracle."""
    from artgpn_b200 import node, weight, mean
    a = H.loglike_case("G4_sun50_initial")
    b = H.loglike_case("G4_sun50_means_none")
    c = H.loglike_case("G1_corot20_verbatim")
    na, nc = H.product_network(a), H.product_network(c)
    oa, ob, oc = H.product_objs(a), H.product_objs(b), H.product_objs(c)
    d = H.dataset(a["data"])
    custom = ([node.Matern52(7.0)], [weight.SquaredExponential(1.3, 25.0), weight.Cosine(0.8, 40.0), weight.Constant(2.0)],
              [mean.Constant(-16.0), None, mean.Linear(0.0, 7784.0)])
    jit = [0.9, 1.1, 1.3]
    exp_custom = _oracle_ll(d["time"], d["y"], d["yerr"], 1, custom[0], custom[1], custom[2], jit)
    for _ in range(3):
        for net, objs, case in ((na, oa, a), (nc, oc, c), (na, ob, b), (nc, oc, c), (na, oa, a)):
            got = net.log_likelihood(objs[0], objs[1], objs[2], case["jitters"])
            assert abs(got - case["ll"]) <= LL_RTOL * abs(case["ll"]), case["name"]
        got = na.log_likelihood(custom[0], custom[1], custom[2], jit)
        assert abs(got - exp_custom) <= LL_RTOL * abs(exp_custom)


def test_non_default_stream_and_device_status():
    import torch
    from artgpn_b200 import synth, node, weight, mean
    b = H.cases()["batch"][1]
    arr = H.arrays()
    theta, exp = arr[b["name"] + "_theta"], arr[b["name"] + "_ll"]
    net = H.product_network(b["data"], b["q"])
    net.set_structure(*synth.objects_from_row(node, weight, mean, theta[0], b["p"], b["q"], b["weights"])[:3])
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        th = torch.from_numpy(theta).cuda()
        out, status = net.log_likelihood_batch(th, return_status=True)
    s.synchronize()
    assert np.all(status == 0) and H.rel_err(out.cpu().numpy(), exp) <= LL_RTOL


def test_large_batch_matches_small_batches_bitwise():
    """2000 walkers at N = 64 (K0) and 700 at N = 257 (tiled, several sub-streams): a row's value does not
    depend on what else is in the batch."""
    from artgpn_b200 import synth, node, weight, mean
    for dname, W in (("synth64", 2000), ("synth257", 700)):
        d = H.dataset(dname)
        net = H.product_network(dname, 2)
        theta = synth.make_walkers(W, p=3, q=2, weights="Constant", seed=11)
        net.set_structure(*synth.objects_from_row(node, weight, mean, theta[0], 3, 2)[:3])
        full = net.log_likelihood_batch(theta)
        assert np.all(np.isfinite(full))
        for lo, hi in ((0, 1), (5, 9), (W - 3, W)):
            assert np.array_equal(net.log_likelihood_batch(theta[lo:hi].copy()), full[lo:hi])
        m = _M()
        exp = oracle.log_likelihood(d["time"], d["y"], d["yerr"], 2, *synth.objects_from_row(m, m, m, theta[W // 2], 3, 2))
        assert abs(full[W // 2] - exp) <= LL_RTOL * abs(exp)


def test_prediction_small_grids_and_training_grid():
    """M < 128, M == 1, and predicting on the training grid itself (the WhiteNoise square-lag quirk then
    fires inside K*, node.py:68-72)."""
    case = H.loglike_case("mixed_a_sun50")
    d = H.dataset(case["data"])
    net = H.product_network(case)
    nodes, weights, means = H.product_objs(case)
    onodes, oweights, omeans = H.oracle_objs(case)
    for ts in (d["time"][:1] + 0.5, np.linspace(d["time"][0], d["time"][-1], 7), d["time"].copy()):
        mu, sd, _ = net.prediction(nodes=nodes, weights=weights, means=means, jitters=case["jitters"], time=ts, dataset=3)
        emu, esd = oracle.prediction(d["time"], d["y"], d["yerr"], case["q"], onodes, oweights, omeans, case["jitters"], ts, 3)
        assert np.max(np.abs(mu - emu)) <= 1e-9 * np.max(np.abs(emu))
        assert np.max(np.abs(sd ** 2 - esd ** 2)) <= 1e-9 * np.max(esd ** 2)


def test_integration_md_ctypes_stub_runs():
    """The ctypes stub printed in INTEGRATION.md (section B) is executed verbatim against the built library
    with kernel objects that only expose a class name and .pars (what the reference's objects expose)."""
    import os
    import re
    from artgpn_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    md = open(os.path.join(root, "INTEGRATION.md")).read()
    code = re.search(r"```python\n(# artgpn/_b200\.py.*?)```", md, re.S).group(1)
    code = code.replace('ctypes.CDLL("libartgpn_b200.so")', 'ctypes.CDLL(%r)' % _lib.LIB_PATH)
    ns = {}
    exec(compile(code, "INTEGRATION.md", "exec"), ns)

    def obj(name, pars):
        return type(name, (), {"pars": list(pars)})()
    case = H.loglike_case("G4_sun50_initial")
    d = H.dataset(case["data"])
    fake_net = type("network", (), {"time": d["time"], "y": d["y"], "yerr": d["yerr"], "p": 3, "q": 1})()
    be = ns["B200Backend"](fake_net)
    nodes = [obj(s[0], s[1]) for s in case["nodes"]]
    weights = [obj(s[0], s[1]) for s in case["weights"]]
    means = [obj(s[0], s[1]) for s in case["means"]]
    ll = be.log_likelihood(nodes, weights, means, case["jitters"])
    assert abs(ll - case["ll"]) <= LL_RTOL * abs(case["ll"])
    a = H.arrays()
    mu, sd, _ = be.prediction(nodes, weights, means, case["jitters"], a["pred_G4_ds2_tstar"], 2)
    assert np.max(np.abs(mu - a["pred_G4_ds2_mean"])) <= 1e-9 * np.max(np.abs(a["pred_G4_ds2_mean"]))
    assert np.max(np.abs(sd ** 2 - a["pred_G4_ds2_std"] ** 2)) <= 1e-9 * np.max(a["pred_G4_ds2_std"] ** 2)


def test_exact_mode_entries_and_conditioning():
    """set_exact(True): covariance entries in the reference's operation order agree with numpy to a few ulp
    (fast mode: ~1e-13); and parity of the log-likelihood degrades like cond(K) x (entry differences) -- at
    cond ~ 3e6 both modes stay within 1e-8 of LAPACK, at cond ~ 3e2 within 1e-13."""
    from artgpn_b200 import node, weight, mean
    from artgpn_b200.arttwo import network
    rng = np.random.default_rng(0)
    N = 300
    t = 57000.0 + np.sort(rng.uniform(0, 1.0 * N, N))
    y = rng.normal(0, 2, (3, N))
    nodes = [node.QuasiPeriodic(18.56, 28.9, 0.7), node.Matern52(9.0)]
    ws = [weight.Constant(1.99), weight.SquaredExponential(0.7, 40.0), weight.Periodic(4.53, 13.0, 0.9), weight.Constant(1.1),
          weight.RationalQuadratic(6.94, 1.5, 20.0), weight.Exponential(2.0, 30.0)]
    ms = [mean.Linear(0.01, 0.1)] * 3
    sp = lambda k: (type(k).__name__, list(k.pars))
    for scale, tol in ((1.0, 1e-13), (1e-2, 1e-8)):
        sig = rng.uniform(0.1, 0.3, (3, N)) * scale
        jit = [0.8 * scale, 0.37 * scale, 1.06 * scale]
        exp = oracle.log_likelihood(t, y, sig, 2, [sp(k) for k in nodes], [sp(k) for k in ws], [sp(k) for k in ms], jit)
        for exact in (False, True):
            net = network(2, t, y[0], sig[0], y[1], sig[1], y[2], sig[2], exact=exact)
            got = net.log_likelihood(nodes, ws, ms, jit)
            assert abs(got - exp) <= tol * abs(exp), (scale, exact, got, exp)
    net = network(2, t, y[0], sig[0], y[1], sig[1], y[2], sig[2])
    for pos in (1, 2, 3):
        ref = oracle.covariance_block(t, [sp(k) for k in nodes], [sp(k) for k in ws], pos - 1, 2)
        net.set_exact(False)
        fast = net._covariance_matrix(nodes, ws, t, pos)
        net.set_exact(True)
        exact = net._covariance_matrix(nodes, ws, t, pos)
        assert np.max(np.abs(fast - ref) / np.abs(ref)) <= 1e-11
        assert np.max(np.abs(exact - ref) / np.abs(ref)) <= 5e-14
        assert np.max(np.abs(exact - ref) / np.abs(ref)) <= np.max(np.abs(fast - ref) / np.abs(ref))
