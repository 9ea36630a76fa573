"""CPU: the oracle (oracle/artgpn_oracle.py) against the golden vectors produced by executing the
unmodified reference (tests/golden/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest

import helpers as H
from helpers import oracle

LL_TOL = 1e-12      # same LAPACK path as the reference: only summation-order noise is expected


@pytest.mark.parametrize("case", H.cases()["loglike"], ids=lambda c: c["name"])
def test_loglike_matches_reference(case):
    exp = case["ll"]
    if exp == "ValueError":
        with pytest.raises(ValueError):
            H.oracle_loglike(case)
    elif exp == "-inf":
        assert np.isneginf(H.oracle_loglike(case))
    else:
        got = H.oracle_loglike(case)
        assert abs(got - exp) <= LL_TOL * abs(exp), (got, exp)


def test_survey_golden_values():
    # SURVEY.md section 8(c) / BASELINE.md section 2, regenerated here and hard-coded as a cross-check
    known = {"G1_corot20_verbatim": -177.9546328644226, "G2_corot167_verbatim": -1355.716666557541,
             "G3_corot20_qp_const": -125.09567824421134, "G3_corot167_qp_const": -1033.1067001340953,
             "G4_sun50_initial": -305.4997408748811}
    for name, v in known.items():
        assert abs(H.loglike_case(name)["ll"] - v) <= 1e-12 * abs(v)
    a = H.arrays()
    assert abs(a["pred_G4_ds1_mean"][100] - (-16.93592831553663)) < 1e-10
    assert abs(a["pred_G4_ds1_std"][100] - 1.9603989751613233) < 1e-10


@pytest.mark.parametrize("b", H.cases()["batch"], ids=lambda c: c["name"])
def test_batch_rows_match_reference(b):
    from artgpn_b200 import synth
    a = H.arrays()
    d = H.dataset(b["data"])
    theta, exp = a[b["name"] + "_theta"], a[b["name"] + "_ll"]

    class M(object):        # module-like triple producing oracle tuples
        def __getattr__(self, name):
            return lambda *pars: (name, list(pars))
    m = M()
    for w in range(theta.shape[0]):
        nodes, ws, ms, jit = synth.objects_from_row(m, m, m, theta[w], b["p"], b["q"], b["weights"])
        got = oracle.log_likelihood(d["time"], d["y"], d["yerr"], b["q"], nodes, ws, ms, jit)
        assert abs(got - exp[w]) <= LL_TOL * abs(exp[w])


@pytest.mark.parametrize("pc", H.cases()["predict"], ids=lambda c: c["name"])
def test_prediction_diag_form_matches_reference(pc):
    a = H.arrays()
    case = H.loglike_case(pc["case"])
    d = H.dataset(case["data"])
    nodes, weights, means = H.oracle_objs(case)
    ts = a[pc["name"] + "_tstar"]
    mu, sd = oracle.prediction(d["time"], d["y"], d["yerr"], case["q"], nodes, weights, means, case["jitters"],
                               ts, pc["dataset"])
    scale = np.max(np.abs(a[pc["name"] + "_mean"]))
    assert np.max(np.abs(mu - a[pc["name"] + "_mean"])) <= 1e-10 * scale
    # variance = k** + s^2 - quad loses digits where it cancels; compare variances on their own scale
    v_ref, v = a[pc["name"] + "_std"] ** 2, sd ** 2
    assert np.max(np.abs(v - v_ref)) <= 1e-9 * np.max(v_ref)


def test_prediction_reference_form_is_same_function():
    case = H.loglike_case("mixed_a_sun50")
    d = H.dataset(case["data"])
    nodes, weights, means = H.oracle_objs(case)
    for ts in (np.linspace(d["time"].min() - 2, d["time"].max() + 2, 37), d["time"] + 0.25):
        m1, s1 = oracle.prediction(d["time"], d["y"], d["yerr"], 3, nodes, weights, means, case["jitters"], ts, 2)
        m2, s2 = oracle.prediction_reference_form(d["time"], d["y"], d["yerr"], 3, nodes, weights, means,
                                                  case["jitters"], ts, 2)
        assert np.allclose(m1, m2, rtol=1e-10, atol=1e-10 * np.max(np.abs(m2)))
        assert np.allclose(s1 ** 2, s2 ** 2, rtol=0, atol=1e-9 * np.max(s2 ** 2))


@pytest.mark.parametrize("kc", H.cases()["kernels"], ids=lambda c: c["key"])
def test_single_kernels_match_reference(kc):
    a = H.arrays()
    t, s5, s9 = a["kern_t"], a["kern_ts5"], a["kern_ts9"]
    f = oracle.node_matrix if kc["family"] == "node" else oracle.weight_matrix
    k = (kc["name"], kc["pars"])
    for got, key in ((f(k, t), "_train"), (f(k, s5, t), "_s5"), (f(k, s9, t), "_s9")):
        exp = a[kc["key"] + key]
        assert got.shape == exp.shape
        assert np.allclose(got, exp, rtol=1e-14, atol=1e-300), kc["key"] + key


@pytest.mark.parametrize("mc", H.cases()["means"], ids=lambda c: c["key"])
def test_means_match_reference(mc):
    a = H.arrays()
    nodes, _, means = H.oracle_objs({"nodes": [], "weights": [], "means": [mc["spec"]]})
    for grid, key in ((a["kern_t"], "_train"), (a["kern_ts5"], "_s5")):
        got = oracle.mean_eval(means[0], grid)
        assert np.allclose(got, a[mc["key"] + key], rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("cc", H.cases()["covariance"], ids=lambda c: c["key"])
def test_covariance_blocks_match_reference(cc):
    a = H.arrays()
    case = H.loglike_case(cc["case"])
    d = H.dataset(case["data"])
    nodes, weights, means = H.oracle_objs(case)
    p, q, key = d["y"].shape[0], case["q"], cc["key"]
    for pos in range(1, p + 1):
        got = oracle.covariance_block(d["time"], nodes, weights, pos - 1, q)
        assert np.allclose(got, a["%s_pos%d" % (key, pos)], rtol=1e-13, atol=1e-13)
        got = oracle.covariance_block(d["time"], nodes, weights, pos - 1, q, yerr=d["yerr"], add_errors=True)
        assert np.allclose(got, a["%s_pos%d_err" % (key, pos)], rtol=1e-13, atol=1e-13)
    tg = a[key + "_grid33_t"]
    assert np.allclose(oracle.covariance_block(d["time"], nodes, weights, 1, q, ta=tg), a[key + "_grid33_pos2"],
                       rtol=1e-13, atol=1e-13)
    assert np.allclose(oracle.compute_matrix(d["time"], d["yerr"], q, nodes, weights), a[key + "_compute"],
                       rtol=1e-13, atol=1e-13)
    assert np.allclose(oracle.compute_matrix(d["time"], d["yerr"], q, nodes, weights, True, True),
                       a[key + "_compute_nugget_shift"], rtol=1e-13, atol=1e-13)
    assert np.allclose(oracle.mean_vector(means, d["time"], p).reshape(-1), a[key + "_meanvec"], rtol=1e-13, atol=1e-10)
    assert np.allclose(oracle.mean_vector(means, tg, p).reshape(-1), a[key + "_meanvec_grid33"], rtol=1e-13, atol=1e-10)


def test_oracle_matches_reference_on_ill_conditioned_goldens():
    """tests/golden/make_golden_cond.py: the oracle performs the reference's numpy / LAPACK operations in the
    reference's order, so it reproduces the reference even where cond(K) ~ 1e11 amplifies every rounding."""
    import json
    import os
    from artgpn_b200 import synth
    with open(os.path.join(H.GOLD, "golden_cond.json")) as f:
        g = json.load(f)
    a = dict(np.load(os.path.join(H.GOLD, "golden_cond.npz")))

    class M(object):
        def __getattr__(self, name):
            return lambda *pars: (name, list(pars))
    m = M()
    for k, c in enumerate(g["cases"]):
        o = synth.objects_from_row(m, m, m, a["theta_%d" % k], g["p"], g["q"], g["weights"])
        got = oracle.log_likelihood(a["time"], a["y"], a["yerr"] * c["scale"], g["q"], *o)
        assert abs(got - c["ll"]) <= 1e-12 * abs(c["ll"]) * max(1.0, max(c["cond"]) * 1e-6), (c["name"], got, c["ll"])
