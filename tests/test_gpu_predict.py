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


def test_prediction_slices_reassemble_the_full_grid():
    """Slices of the prediction grid (what each rank computes in sharding.prediction_sharded) give the same
    numbers as one call on the whole grid -- Linear means are centred on the WHOLE grid."""
    case = H.loglike_case("G4_sun50_initial")          # Linear means
    net = H.product_network(case)
    nodes, weights, means = H.product_objs(case)
    means[0].pars[0] = 0.03                              # non-zero slope so that the centring matters
    d = H.dataset(case["data"])
    ts = np.linspace(d["time"].min() - 5, d["time"].max() + 5, 1000)
    mu, sd, _ = net.prediction(nodes=nodes, weights=weights, means=means, jitters=case["jitters"], time=ts, dataset=1)
    parts = [net.prediction_slice(nodes, weights, means, case["jitters"], ts, lo, hi, 1) for lo, hi in ((0, 333), (333, 700), (700, 1000))]
    mu2 = np.concatenate([p[0] for p in parts])
    sd2 = np.concatenate([p[1] for p in parts])
    assert np.allclose(mu, mu2, rtol=0, atol=1e-12 * np.max(np.abs(mu)))
    assert np.allclose(sd, sd2, rtol=0, atol=1e-12 * np.max(sd))


def test_prediction_sharded_single_rank_nccl():
    import os
    import socket
    import torch
    import torch.distributed as dist
    from artgpn_b200 import sharding
    case = H.loglike_case("G4_sun50_initial")
    net = H.product_network(case)
    nodes, weights, means = H.product_objs(case)
    a = H.arrays()
    ts = a["pred_G4_ds1_tstar"]
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        mu, sd, tt = sharding.prediction_sharded(net, nodes, weights, means, case["jitters"], ts, dataset=1)
    finally:
        dist.destroy_process_group()
    assert np.max(np.abs(mu - a["pred_G4_ds1_mean"])) <= 1e-9 * np.max(np.abs(a["pred_G4_ds1_mean"]))
    assert np.max(np.abs(sd ** 2 - a["pred_G4_ds1_std"] ** 2)) <= 1e-9 * np.max(a["pred_G4_ds1_std"] ** 2)
