"""GPU: network.prediction through the C ABI vs the reference goldens."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["stacked", "dmma"])
def predict_path(request, monkeypatch):
    """Default: K* rides along as extra block rows of the factorisation (INT8-sliced left-looking updates, panel
    solves, folded forward solve); AGPN_PREDICT=dmma (read when the context is created): the stand-alone
    predict_kernel on the FP64 tensor instruction -- also the automatic fall-back of the default path."""
    if request.param == "dmma":
        monkeypatch.setenv("AGPN_PREDICT", "dmma")
    else:
        monkeypatch.delenv("AGPN_PREDICT", raising=False)
    return request.param


@pytest.mark.parametrize("pc", H.cases()["predict"], ids=lambda c: c["name"])
def test_prediction_vs_reference_goldens(pc, capsys, predict_path):
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


def test_prediction_paths_agree_on_a_multi_block_network(monkeypatch):
    """N = 700 (6 block columns), M = 3000: the stacked INT8 path, the DMMA predict_kernel and the oracle agree; a
    WhiteNoise node on a grid with M == N (the reference's square-lag quirk, node.py:68-72, puts wn^2 on the
    diagonal of K*) must still match the oracle -- the stacked path detects entries above their row scale and hands
    over to predict_kernel."""
    from artgpn_b200 import synth, node, weight, mean
    from artgpn_b200.arttwo import network
    from helpers import oracle
    N, M = 700, 3000
    t, y, sig = synth.make_data(N, 3, seed=2)
    theta = synth.make_walkers(1, p=3, q=2, seed=3)

    class Mm(object):
        def __getattr__(self, name):
            return lambda *pars: (name, list(pars))
    m = Mm()
    ts = np.linspace(t.min() - 3, t.max() + 3, M)
    o_ref = synth.objects_from_row(m, m, m, theta[0], 3, 2)
    emu, esd = oracle.prediction(t, y, sig, 2, *o_ref, ts, 2)
    got = {}
    for mode in ("stacked", "dmma"):
        if mode == "dmma":
            monkeypatch.setenv("AGPN_PREDICT", "dmma")
        else:
            monkeypatch.delenv("AGPN_PREDICT", raising=False)
        net = network(2, t, y[0], sig[0], y[1], sig[1], y[2], sig[2])
        nodes, ws, ms, jit = synth.objects_from_row(node, weight, mean, theta[0], 3, 2)
        mu, sd, _ = net.prediction(nodes=nodes, weights=ws, means=ms, jitters=jit, time=ts, dataset=2)
        assert np.max(np.abs(mu - emu)) <= 1e-9 * np.max(np.abs(emu)), mode
        assert np.max(np.abs(sd ** 2 - esd ** 2)) <= 1e-9 * np.max(esd ** 2), mode
        assert net.predict_path() == (1 if mode == "stacked" else 0)
        got[mode] = (mu, sd)
        net.close()
    # fault injection: unit row scales for the K* rows -> entries exceed them -> detected -> predict_kernel takes over
    monkeypatch.delenv("AGPN_PREDICT", raising=False)
    monkeypatch.setenv("AGPN_I8_TEST_WRAP", "1")
    net = network(2, t, y[0], sig[0], y[1], sig[1], y[2], sig[2])
    nodes, ws, ms, jit = synth.objects_from_row(node, weight, mean, theta[0], 3, 2)
    mu, sd, _ = net.prediction(nodes=nodes, weights=ws, means=ms, jitters=jit, time=ts, dataset=2)
    assert net.predict_path() == 0
    assert np.array_equal(mu, got["dmma"][0]) and np.array_equal(sd, got["dmma"][1])
    net.close()
    monkeypatch.delenv("AGPN_I8_TEST_WRAP", raising=False)
    assert np.max(np.abs(got["stacked"][0] - got["dmma"][0])) <= 1e-11 * np.max(np.abs(emu))
    # WhiteNoise quirk, M == N
    monkeypatch.delenv("AGPN_PREDICT", raising=False)
    N2 = 300
    t2, y2, sig2 = synth.make_data(N2, 1, seed=5)
    ts2 = t2 + 0.37
    net = network(2, t2, y2[0], sig2[0])
    nodes = [node.QuasiPeriodic(20.0, 25.0, 0.8), node.WhiteNoise(3.0)]
    ws = [weight.Constant(2.0), weight.Constant(1.5)]
    mu, sd, _ = net.prediction(nodes=nodes, weights=ws, means=[None], jitters=[0.4], time=ts2, dataset=1)
    emu2, esd2 = oracle.prediction(t2, y2, sig2, 2, [("QuasiPeriodic", [20.0, 25.0, 0.8]), ("WhiteNoise", [3.0])],
                                   [("Constant", [2.0]), ("Constant", [1.5])], [None], [0.4], ts2, 1)
    assert np.max(np.abs(mu - emu2)) <= 1e-9 * np.max(np.abs(emu2))
    ok = np.isfinite(esd2)
    assert np.array_equal(ok, np.isfinite(sd))
    assert np.max(np.abs(sd[ok] ** 2 - esd2[ok] ** 2)) <= 1e-9 * np.max(esd2[ok] ** 2)
