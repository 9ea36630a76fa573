"""GPU: every node / weight kernel class and mean function, device value vs the reference
goldens and vs the oracle on random lags (through the C ABI: agpn_kernel_eval / agpn_mean_table)."""
import numpy as np
import pytest

import helpers as H
from helpers import oracle

pytestmark = pytest.mark.gpu

# libdevice sin/exp/pow are within 1-2 ulp of the host libm and the hoisted constants add a few
# ulp of the exponent: 1e-12 relative (+1e-14 absolute for values that cancel to ~0) per entry.
RTOL, ATOL = 1e-12, 1e-14


def _mod(fam):
    from artgpn_b200 import node, weight
    return node if fam == "node" else weight


@pytest.mark.parametrize("kc", H.cases()["kernels"], ids=lambda c: c["key"])
def test_kernel_classes_vs_reference_goldens(kc):
    a = H.arrays()
    t, s5, s9 = a["kern_t"], a["kern_ts5"], a["kern_ts9"]
    net = H.product_network("sun50", 1)
    net.time = t
    k = getattr(_mod(kc["family"]), kc["name"])(*kc["pars"])
    got = {"_train": net._kernel_matrix(k, t), "_s5": net._predict_kernel_matrix(k, s5),
           "_s9": net._predict_kernel_matrix(k, s9)}
    for key, g in got.items():
        exp = a[kc["key"] + key]
        assert g.shape == exp.shape
        assert np.allclose(g, exp, rtol=RTOL, atol=ATOL * max(1.0, np.max(np.abs(exp)))), (kc["key"] + key, np.max(np.abs(g - exp)))


@pytest.mark.parametrize("fam", ["node", "weight"])
def test_kernel_classes_vs_oracle_random_lags(fam):
    rng = np.random.default_rng(7)
    r = rng.uniform(-40, 40, size=(37, 53))
    r[0, 0] = 0.0
    t1 = 55000.0 + rng.uniform(0, 60, size=(37, 1))
    t2 = 55000.0 + rng.uniform(0, 60, size=(1, 53))
    seen = set()
    for kc in H.cases()["kernels"]:
        if kc["family"] != fam:
            continue
        k = getattr(_mod(fam), kc["name"])(*kc["pars"])
        f = oracle.node_eval if fam == "node" else oracle.weight_eval
        if kc["name"] in ("Linear", "Polynomial"):
            got, exp = k(None, t1, t2), f((kc["name"], kc["pars"]), None, t1, t2)
        else:
            got, exp = k(r), f((kc["name"], kc["pars"]), r)
        assert np.allclose(got, exp, rtol=RTOL, atol=ATOL * max(1.0, np.max(np.abs(exp)))), kc["name"]
        seen.add(kc["name"])
    assert len(seen) >= 14


def test_white_noise_square_quirk():
    from artgpn_b200 import node, weight
    r = np.zeros((5, 5))
    assert np.array_equal(node.WhiteNoise(0.5)(r), 0.25 * np.eye(5))
    assert np.array_equal(weight.WhiteNoise(2.0)(np.ones((4, 4))), 4.0 * np.eye(4))
    assert np.array_equal(node.WhiteNoise(0.5)(np.zeros((5, 4))), np.zeros((5, 4)))


@pytest.mark.parametrize("mc", H.cases()["means"], ids=lambda c: c["key"])
def test_mean_functions_vs_reference_goldens(mc):
    from artgpn_b200 import mean
    a = H.arrays()
    m = H.build(mean, mc["spec"])
    for grid, key in ((a["kern_t"], "_train"), (a["kern_ts5"], "_s5")):
        got = m(grid)
        exp = a[mc["key"] + key]
        assert np.allclose(got, exp, rtol=1e-11, atol=1e-11 * max(1.0, np.max(np.abs(exp)))), (mc["key"], got, exp)


@pytest.mark.parametrize("cc", H.cases()["covariance"], ids=lambda c: c["key"])
def test_covariance_matrix_and_compute_matrix(cc):
    a = H.arrays()
    case = H.loglike_case(cc["case"])
    d = H.dataset(case["data"])
    net = H.product_network(case)
    nodes, weights, means = H.product_objs(case)
    key = cc["key"]
    for pos in range(1, d["y"].shape[0] + 1):
        for suffix, flag in (("", False), ("_err", True)):
            exp = a["%s_pos%d%s" % (key, pos, suffix)]
            got = net._covariance_matrix(nodes, weights, d["time"], pos, add_errors=flag)
            assert np.allclose(got, exp, rtol=1e-11, atol=1e-13 * np.max(np.abs(exp))), (pos, suffix)
    tg = a[key + "_grid33_t"]
    exp = a[key + "_grid33_pos2"]
    assert np.allclose(net._covariance_matrix(nodes, weights, tg, 2), exp, rtol=1e-11, atol=1e-13 * np.max(np.abs(exp)))
    exp = a[key + "_compute"]
    assert np.allclose(net.compute_matrix(nodes, weights, d["time"]), exp, rtol=1e-11, atol=1e-13 * np.max(np.abs(exp)))
    exp = a[key + "_compute_nugget_shift"]
    assert np.allclose(net.compute_matrix(nodes, weights, d["time"], nugget=True, shift=True), exp, rtol=1e-11,
                       atol=1e-13 * np.max(np.abs(exp)))
    exp = a[key + "_meanvec"]
    assert np.allclose(net._mean(means), exp, rtol=1e-11, atol=1e-11 * np.max(np.abs(exp)))
    exp = a[key + "_meanvec_grid33"]
    assert np.allclose(net._mean(means, tg), exp, rtol=1e-11, atol=1e-11 * np.max(np.abs(exp)))
