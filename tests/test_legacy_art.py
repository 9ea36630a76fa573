"""Legacy art.network (art.py) and the full predictive covariance: oracle vs reference goldens on the
CPU, product (C ABI) vs goldens on the GPU."""
import numpy as np
import pytest

import helpers as H
from helpers import oracle


def _sp(s):
    return None if s is None else (s[0], list(s[1]))


@pytest.mark.parametrize("c", H.cases()["legacy_art"], ids=lambda c: c["name"])
def test_oracle_legacy_art_vs_reference(c):
    a = H.arrays()
    d = H.dataset(c["data"])
    nodes = [_sp(s) for s in c["nodes"]]
    means = [_sp(s) for s in c["means"]]
    q = len(nodes)
    ll = oracle.art_log_likelihood(d["time"], d["y"], d["yerr"], q, nodes, _sp(c["weight"]), c["weights_values"], means, c["jitters"])
    assert abs(ll - c["ll"]) <= 1e-12 * abs(c["ll"])
    for ds in c["datasets"]:
        key = "%s_ds%d" % (c["name"], ds)
        mu, sd, cov = oracle.art_prediction(d["time"], d["y"], d["yerr"], q, nodes, _sp(c["weight"]), c["weights_values"],
                                            means, c["jitters"], c["ctor_jitters"], a[key + "_tstar"], ds)
        assert np.allclose(mu, a[key + "_mean"], rtol=1e-10, atol=1e-10 * np.max(np.abs(a[key + "_mean"])))
        assert np.allclose(cov, a[key + "_cov"], rtol=0, atol=1e-10 * np.max(np.abs(a[key + "_cov"])))


@pytest.mark.gpu
@pytest.mark.parametrize("c", H.cases()["legacy_art"], ids=lambda c: c["name"])
def test_legacy_art_network_vs_reference(c, capsys):
    from artgpn_b200 import node, weight, mean
    from artgpn_b200.art import network
    a = H.arrays()
    d = H.dataset(c["data"])
    p = d["y"].shape[0]
    nodes = [H.build(node, s) for s in c["nodes"]]
    means = [H.build(mean, s) for s in c["means"]]
    args = []
    for i in range(p):
        args += [d["y"][i].copy(), d["yerr"][i].copy()]
    net = network(nodes, [H.build(weight, c["weight"])], list(c["weights_values"]), means, list(c["ctor_jitters"]),
                  d["time"].copy(), *args)
    assert np.allclose(net.yerr, d["yerr"] ** 2)                                  # art.py:58
    w0 = H.build(weight, c["weight"])
    ll = net.log_likelihood(nodes, [w0], list(c["weights_values"]), means, list(c["jitters"]))
    assert abs(ll - c["ll"]) <= 1e-9 * abs(c["ll"]), (ll, c["ll"])
    assert w0.pars[0] == c["weights_values"][-1]                                  # in-place overwrite (art.py:259)
    for ds in c["datasets"]:
        key = "%s_ds%d" % (c["name"], ds)
        mu, sd, cov = net.prediction(nodes=nodes, weights=[H.build(weight, c["weight"])], weights_values=list(c["weights_values"]),
                                     means=means, jitters=list(c["jitters"]), time=a[key + "_tstar"].copy(), dataset=ds)
        assert "Working with dataset %d" % ds in capsys.readouterr().out
        scale = np.max(np.abs(a[key + "_cov"]))
        assert np.max(np.abs(mu - a[key + "_mean"])) <= 1e-9 * np.max(np.abs(a[key + "_mean"]))
        assert np.max(np.abs(cov - a[key + "_cov"])) <= 1e-9 * scale
        assert np.max(np.abs(sd ** 2 - a[key + "_std"] ** 2)) <= 1e-9 * scale
        assert np.array_equal(cov, cov.T)
    exp = a[c["name"] + "_compute"]
    got = net.compute_matrix(nodes, H.build(weight, c["weight"]), list(c["weights_values"]), d["time"])
    assert np.allclose(got, exp, rtol=1e-11, atol=1e-13 * np.max(np.abs(exp)))


@pytest.mark.gpu
@pytest.mark.parametrize("M", [1, 50, 131, 300])
def test_prediction_cov_vs_oracle_and_diag_path(M):
    """arttwo extension: full predictive covariance = Schur complement of the augmented matrix; its
    diagonal must agree with network.prediction's std and the whole matrix with the reference form."""
    case = H.loglike_case("mixed_a_sun50")
    d = H.dataset(case["data"])
    net = H.product_network(case)
    nodes, weights, means = H.product_objs(case)
    onodes, oweights, omeans = H.oracle_objs(case)
    ts = np.linspace(d["time"].min() - 3, d["time"].max() + 3, M) if M != 50 else d["time"].copy()
    for ds in (1, 3):
        mu, sd, cov = net.prediction_cov(nodes=nodes, weights=weights, means=means, jitters=case["jitters"], time=ts, dataset=ds)
        emu, esd, ecov = oracle.prediction_reference_form(d["time"], d["y"], d["yerr"], case["q"], onodes, oweights, omeans,
                                                          case["jitters"], ts, ds, return_cov=True)
        scale = np.max(np.abs(ecov))
        assert np.max(np.abs(mu - emu)) <= 1e-9 * np.max(np.abs(emu))
        assert np.max(np.abs(cov - ecov)) <= 1e-9 * scale
        mu2, sd2, _ = net.prediction(nodes=nodes, weights=weights, means=means, jitters=case["jitters"], time=ts, dataset=ds)
        assert np.max(np.abs(mu - mu2)) <= 1e-10 * np.max(np.abs(emu))
        assert np.max(np.abs(sd ** 2 - sd2 ** 2)) <= 1e-10 * scale
