"""GPU: network.prediction through the C ABI vs the reference goldens."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("pc", H.cases()["predict"], ids=lambda c: c["name"])
def test_prediction_vs_reference_goldens(pc, capsys):
    a = H.arrays()
    case = H.loglike_case(pc["case"])
    net = H.product_network(case)
    nodes, weights, means = H.product_objs(case)
    ts = a[pc["name"] + "_tstar"]
    mu, sd, tt = net.prediction(nodes=nodes, weights=weights, means=means, jitters=case["jitters"], time=ts.copy(),
                                dataset=pc["dataset"])
    assert "Working with dataset %d" % pc["dataset"] in capsys.readouterr().out      # arttwo.py:357
    assert np.array_equal(tt, ts)                                                     # third value = time grid
    exp_mu, exp_sd = a[pc["name"] + "_mean"], a[pc["name"] + "_std"]
    assert np.max(np.abs(mu - exp_mu)) <= 1e-9 * np.max(np.abs(exp_mu))
    # variance = k** + s^2 - quad: compare on the scale of the prior variance (1e-9 relative to it)
    assert np.max(np.abs(sd ** 2 - exp_sd ** 2)) <= 1e-9 * np.max(exp_sd ** 2)


def test_prediction_not_positive_definite_raises():
    case = H.loglike_case("notpd_negative_cosine_weight")
    net = H.product_network(case)
    nodes, weights, means = H.product_objs(case)
    with pytest.raises(np.linalg.LinAlgError):
        net.prediction(nodes=nodes, weights=weights, means=means, jitters=case["jitters"],
                       time=np.linspace(57222.0, 57300.0, 10), dataset=1)
