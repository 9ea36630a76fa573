"""GPU: round-2 additions -- BASELINE sizes C4 / C5 in pytest, multi-rank identity over NCCL, the sharded C-ABI entry,
network.sample, context-free mean evaluation, device restore, the whole-grid WhiteNoise quirk of prediction slices."""
import ctypes
import io
import contextlib
import os
import socket
import sys

import numpy as np
import pytest

import helpers as H
from helpers import oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Spec(object):
    def __getattr__(self, name):
        return lambda *pars: (name, list(pars))


def _net(N, p, q, seed=0, **kw):
    from artgpn_b200 import synth
    from artgpn_b200.arttwo import network
    t, y, sig = synth.make_data(N, p, seed=seed)
    args = []
    for i in range(p):
        args += [y[i], sig[i]]
    return network(q, t, *args, **kw), (t, y, sig)


# ---- BASELINE config 4 size: N = 8192 (64 block columns), one series, one walker, both update paths ---------------
@pytest.mark.parametrize("mode", ["i8", "dmma"])
def test_c4_size_one_block_vs_lapack(monkeypatch, mode):
    from artgpn_b200 import synth, node, weight, mean
    monkeypatch.setenv("AGPN_SYRK", mode)
    net, (t, y, sig) = _net(8192, 1, 3)
    theta = synth.make_walkers(1, p=1, q=3, seed=1)
    o = synth.objects_from_row(node, weight, mean, theta[0], 1, 3)
    net.set_structure(*o[:3])
    assert net.update_path() == (3 if mode == "i8" else 0)
    got = net.log_likelihood_batch(theta)[0]
    m = _Spec()
    exp = oracle.log_likelihood(t, y, sig, 3, *synth.objects_from_row(m, m, m, theta[0], 1, 3))
    assert abs(got - exp) <= 1e-9 * abs(exp), (mode, got, exp)
    net.close()


# ---- BASELINE config 5: 3 x 4096 fitted network, 10^5-point grid; oracle on a random 2000-point subset ------------
@pytest.mark.parametrize("path", ["stacked", "dmma"])
def test_c5_prediction_full_grid_vs_oracle_subset(monkeypatch, path):
    from artgpn_b200 import synth, node, weight, mean
    if path == "dmma":
        monkeypatch.setenv("AGPN_PREDICT", "dmma")
    c = synth.CONFIGS["C5"]
    net, (t, y, sig) = _net(c["N"], c["p"], c["q"])
    theta = synth.make_walkers(1, p=c["p"], q=c["q"], seed=1)[0]
    o = synth.objects_from_row(node, weight, mean, theta, c["p"], c["q"])
    grid = synth.prediction_grid(t, c["M"])
    with contextlib.redirect_stdout(io.StringIO()):
        mu, sd, tt = net.prediction(nodes=o[0], weights=o[1], means=o[2], jitters=o[3], time=grid, dataset=2)
    assert net.predict_path() == (1 if path == "stacked" else 0)
    assert mu.shape == (c["M"],) and np.all(np.isfinite(mu)) and np.all(np.isfinite(sd)) and tt is not None
    idx = np.sort(np.random.default_rng(11).choice(c["M"], 2000, replace=False))
    m = _Spec()
    oo = synth.objects_from_row(m, m, m, theta, c["p"], c["q"])
    emu, esd = oracle.prediction(t, y, sig, c["q"], *oo, grid[idx], 2)
    # mean.Linear is centred on the mean of the grid it is given (mean.py:92): whole grid, not the subset
    emu = emu - oracle.mean_vector(oo[2], grid[idx], c["p"])[1] + oracle.mean_vector(oo[2], grid, c["p"])[1][idx]
    assert np.max(np.abs(mu[idx] - emu)) <= 1e-9 * np.max(np.abs(emu))
    assert np.max(np.abs(sd[idx] ** 2 - esd ** 2)) <= 1e-9 * np.max(esd ** 2)
    net.close()


# ---- sharded entry of the C ABI on a one-rank communicator ---------------------------------------------------------
def test_c_abi_sharded_entry_one_rank():
    from artgpn_b200 import synth, node, weight, mean, _lib
    lib = _lib.load()
    net, _ = _net(600, 3, 2, seed=5)
    theta = synth.make_walkers(7, p=3, q=2, seed=2)
    net.set_structure(*synth.objects_from_row(node, weight, mean, theta[0], 3, 2)[:3])
    ref, ref_st = net.log_likelihood_batch(theta, return_status=True)
    idb = np.zeros(128, dtype=np.uint8)
    _lib.check(lib.agpn_comm_unique_id(_lib.ptr(idb)))
    _lib.check(lib.agpn_comm_init(net._ctx(), _lib.ptr(idb), 0, 1))
    r, n = ctypes.c_int(-1), ctypes.c_int(-1)
    _lib.check(lib.agpn_comm_rank(net._ctx(), ctypes.byref(r), ctypes.byref(n)))
    assert (r.value, n.value) == (0, 1)
    out = np.empty(7)
    st = np.full(7, -1, dtype=np.int32)
    _lib.check(lib.agpn_loglike_batch_sharded(net._ctx(), 7, _lib.ptr(theta), 0, _lib.ptr(out), 0, _lib.ptr(st), None))
    assert np.array_equal(out, ref) and np.array_equal(st, ref_st)
    net.close()


def test_sharded_entry_needs_comm():
    from artgpn_b200 import synth, node, weight, mean, _lib
    lib = _lib.load()
    net, _ = _net(64, 3, 2, seed=5)
    theta = synth.make_walkers(2, p=3, q=2, seed=2)
    net.set_structure(*synth.objects_from_row(node, weight, mean, theta[0], 3, 2)[:3])
    out = np.empty(2)
    assert lib.agpn_loglike_batch_sharded(net._ctx(), 2, _lib.ptr(theta), 0, _lib.ptr(out), 0, None, None) != 0
    assert b"agpn_comm_init" in lib.agpn_last_error()
    net.close()


# ---- two ranks over NCCL: sharded results are bit-identical to one GPU (skipped on a one-GPU box) -----------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _nccl_worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from artgpn_b200 import synth, node, weight, mean, sharding
    from artgpn_b200.arttwo import network
    t, y, sig = synth.make_data(700, 3, seed=5)
    theta = synth.make_walkers(9, p=3, q=2, seed=2)            # uneven: 5 + 4 rows
    net = network(2, t, y[0], sig[0], y[1], sig[1], y[2], sig[2], device=rank)
    o = synth.objects_from_row(node, weight, mean, theta[0], 3, 2)
    net.set_structure(*o[:3])
    full = sharding.log_likelihood_batch_sharded(net, theta)
    even = sharding.log_likelihood_batch_sharded(net, theta[:8])
    grid = np.linspace(t.min() - 3, t.max() + 3, 1001)
    with contextlib.redirect_stdout(io.StringIO()):
        mu, sd, _ = sharding.prediction_sharded(net, *o, grid, dataset=2)
    np.savez(out_path % rank, full=full.cpu().numpy(), even=even.cpu().numpy(), mu=mu, sd=sd)
    dist.barrier()
    net.close()
    dist.destroy_process_group()


def test_two_ranks_nccl_bit_identical_to_one_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from artgpn_b200 import synth, node, weight, mean
    out = str(tmp_path / "r%d.npz")
    mp.spawn(_nccl_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    net, (t, y, sig) = _net(700, 3, 2, seed=5)
    theta = synth.make_walkers(9, p=3, q=2, seed=2)
    o = synth.objects_from_row(node, weight, mean, theta[0], 3, 2)
    net.set_structure(*o[:3])
    one = net.log_likelihood_batch(theta)
    grid = np.linspace(t.min() - 3, t.max() + 3, 1001)
    with contextlib.redirect_stdout(io.StringIO()):
        mu, sd, _ = net.prediction(nodes=o[0], weights=o[1], means=o[2], jitters=o[3], time=grid, dataset=2)
    for r in range(2):
        d = np.load(out % r)
        assert np.array_equal(d["full"], one)                # every rank holds the full vector, same bits as one GPU
        assert np.array_equal(d["even"], one[:8])
        # prediction points are independent: a slice of the grid gives the values of the whole-grid call
        assert np.max(np.abs(d["mu"] - mu)) <= 1e-11 * np.max(np.abs(mu))
        assert np.max(np.abs(d["sd"] ** 2 - sd ** 2)) <= 1e-11 * np.max(sd ** 2)
    net.close()


# ---- network.sample (arttwo.py:129-147) ------------------------------------------------------------------------------
def test_sample_is_chol_times_normal_and_has_the_right_moments():
    from artgpn_b200 import node
    net, (t, _, _) = _net(150, 1, 1, seed=3)
    K = net._kernel_matrix(node.QuasiPeriodic(20.0, 15.0, 0.8), t) + 0.05 * np.identity(t.size)
    # deterministic part: the same z through numpy's factor
    z = np.random.default_rng(5).standard_normal((3, t.size))
    got = net.sample(K, t, size=3, random_state=np.random.default_rng(5))
    exp = z @ np.linalg.cholesky(K).T
    assert got.shape == (3, t.size)
    assert np.max(np.abs(got - exp)) <= 1e-10 * np.max(np.abs(exp))
    assert net._last_sample_nugget == 0.0
    one = net.sample(K, t)
    assert one.shape == (t.size,)
    # moments: empirical covariance of 20000 draws
    S = net.sample(K, t, size=20000, random_state=1)
    C = S.T @ S / S.shape[0]
    assert np.max(np.abs(C - K)) <= 6.0 * np.max(np.diag(K)) / np.sqrt(S.shape[0])
    assert np.max(np.abs(S.mean(axis=0))) <= 5.0 * np.sqrt(np.max(np.diag(K)) / S.shape[0])
    net.close()


def test_sample_allows_a_singular_covariance():
    """scipy's allow_singular=True: a rank-deficient (positive semi-definite) matrix is accepted."""
    net, (t, _, _) = _net(40, 1, 1, seed=3)
    u = np.random.default_rng(2).standard_normal((t.size, 3))
    K = u @ u.T                                              # rank 3
    S = net.sample(K, t, size=5000, random_state=3)
    assert net._last_sample_nugget > 0.0 and net._last_sample_nugget <= 1e-6 * np.max(np.diag(K))
    C = S.T @ S / S.shape[0]
    assert np.max(np.abs(C - K)) <= 6.0 * np.max(np.diag(K)) / np.sqrt(S.shape[0])
    with pytest.raises(Exception):
        net.sample(-np.identity(t.size), t)                  # not positive semi-definite
    net.close()


# ---- mean functions without touching the context's structure --------------------------------------------------------
def test_mean_evaluation_leaves_structure_and_graph_alone():
    from artgpn_b200 import synth, node, weight, mean, _lib
    lib = _lib.load()
    net, (t, y, sig) = _net(300, 3, 2, seed=5)
    theta = synth.make_walkers(4, p=3, q=2, seed=2)
    S = net.set_structure(*synth.objects_from_row(node, weight, mean, theta[0], 3, 2)[:3])
    before = net.log_likelihood_batch(theta)
    ms = [mean.Linear(0.1, 2.0), None, mean.Sine(1.0, 30.0, 0.3, 0.5) + mean.Constant(2.0)]
    got = net._mean(ms)
    m = _Spec()
    exp = np.concatenate([0.1 * (t - t.mean()) + 2.0, np.zeros_like(t), np.sin(2 * np.pi * t / 30.0 + 0.3) + 0.5 + 2.0])
    assert np.allclose(got, exp, rtol=1e-11, atol=1e-11)
    assert net._structure is S and lib.agpn_theta_dim(net._ctx()) == S.D          # untouched
    l0 = lib.agpn_launch_count()
    after = net.log_likelihood_batch(theta)
    assert np.array_equal(before, after)
    # a network without any structure still has none afterwards
    net2, _ = _net(50, 2, 1, seed=1)
    net2._mean([mean.Constant(3.0)])
    with pytest.raises(RuntimeError):
        net2.log_likelihood_batch(np.zeros((1, 5)))
    assert np.allclose(mean.Constant(3.0)(t[:5]), 3.0)
    net.close()
    net2.close()
    assert l0 > 0


def test_current_device_is_restored():
    import torch
    from artgpn_b200 import synth, node, weight, mean
    torch.cuda.set_device(0)
    dev = torch.cuda.device_count() - 1                       # cuda:1 when the box has it
    net, _ = _net(200, 3, 2, seed=5, device=dev)
    theta = synth.make_walkers(2, p=3, q=2, seed=2)
    o = synth.objects_from_row(node, weight, mean, theta[0], 3, 2)
    net.set_structure(*o[:3])
    net.log_likelihood_batch(theta)
    net.log_likelihood(*o)
    net._covariance_matrix(o[0], o[1], net.time, 1)
    with contextlib.redirect_stdout(io.StringIO()):
        net.prediction(nodes=o[0], weights=o[1], means=o[2], jitters=o[3], time=np.linspace(57000, 57100, 50))
    cur = ctypes.c_int(-1)
    rt = ctypes.CDLL("libcudart.so.12")
    assert rt.cudaGetDevice(ctypes.byref(cur)) == 0 and cur.value == 0
    assert torch.cuda.current_device() == 0
    x = torch.zeros(4, device="cuda")
    assert x.device.index == 0
    net.close()


# ---- WhiteNoise's square-lag quirk inside K* is a property of the WHOLE prediction grid ---------------------------
@pytest.mark.parametrize("M", [200, 300])
def test_prediction_slices_see_the_whole_grid_whitenoise_quirk(M):
    """N = 200.  M = 200: the reference's K* lag array is square -> wn^2 on its diagonal, for every slicing of the grid.
    M = 300: not square -> no wn^2, even for a slice of exactly N = 200 points."""
    from artgpn_b200 import node, weight, mean
    net, (t, y, sig) = _net(200, 2, 2, seed=9)
    nodes = [node.QuasiPeriodic(20.0, 25.0, 0.9), node.WhiteNoise(0.7)]
    ws = [weight.Constant(1.5), weight.Constant(0.9), weight.Constant(1.1), weight.Constant(1.3)]
    ms = [mean.Linear(0.01, 0.2), mean.Constant(0.1)]
    jit = [0.6, 0.8]
    grid = np.linspace(t.min() - 2, t.max() + 2, M)
    with contextlib.redirect_stdout(io.StringIO()):
        mu, sd, _ = net.prediction(nodes=nodes, weights=ws, means=ms, jitters=jit, time=grid, dataset=1)
    m = _Spec()
    emu, esd = oracle.prediction(t, y, sig, 2, [("QuasiPeriodic", [20.0, 25.0, 0.9]), ("WhiteNoise", [0.7])],
                                 [("Constant", [1.5]), ("Constant", [0.9]), ("Constant", [1.1]), ("Constant", [1.3])],
                                 [("Linear", [0.01, 0.2]), ("Constant", [0.1])], jit, grid, 1)
    assert np.max(np.abs(mu - emu)) <= 1e-9 * np.max(np.abs(emu))
    for lo, hi in ((0, 100), (100, M), (0, 200), (M - 200, M), (37, 38)):
        smu, ssd = net.prediction_slice(nodes, ws, ms, jit, grid, lo, hi, dataset=1)
        assert np.max(np.abs(smu - mu[lo:hi])) <= 1e-11 * np.max(np.abs(mu)), (M, lo, hi)
        assert np.max(np.abs(ssd ** 2 - sd[lo:hi] ** 2)) <= 1e-11 * np.max(sd ** 2), (M, lo, hi)
    net.close()


def test_posterior_samples_have_the_predictive_moments():
    from artgpn_b200 import node, weight, mean
    net, (t, y, sig) = _net(120, 2, 1, seed=4)
    nodes = [node.QuasiPeriodic(20.0, 15.0, 0.8)]
    ws = [weight.Constant(1.5), weight.Constant(0.9)]
    ms = [mean.Linear(0.01, 0.2), None]
    jit = [0.6, 0.8]
    grid = np.linspace(t.min() - 2, t.max() + 2, 90)
    with contextlib.redirect_stdout(io.StringIO()):
        mu, sd, cov = net.prediction_cov(nodes, ws, ms, jit, grid, 1)
        S = net.posterior_sample(nodes, ws, ms, jit, grid, 1, size=20000, random_state=7)
    assert S.shape == (20000, grid.size)
    assert np.max(np.abs(S.mean(axis=0) - mu)) <= 5.0 * np.max(sd) / np.sqrt(S.shape[0])
    C = np.cov(S.T)
    assert np.max(np.abs(C - cov)) <= 6.0 * np.max(np.diag(cov)) / np.sqrt(S.shape[0])
    net.close()


def test_predictive_covariance_on_the_int8_path_and_its_fallback(monkeypatch):
    """agpn_predict_cov: the Schur update of the trailing block runs on the INT8 tensor path (predict_path 1); forcing the
    fall-back (AGPN_I8_TEST_WRAP=1) or the FP64 instruction (AGPN_PREDICT=dmma) gives the same covariance to round-off,
    and all of them match the reference form of the oracle."""
    from artgpn_b200 import synth, node, weight, mean
    t, y, sig = synth.make_data(700, 2, seed=6)
    theta = synth.make_walkers(1, p=2, q=2, seed=3)[0]
    grid = np.linspace(t.min() - 2, t.max() + 2, 300)
    m = _Spec()
    oo = synth.objects_from_row(m, m, m, theta, 2, 2)
    emu, esd, ecov = oracle.prediction_reference_form(t, y, sig, 2, *oo, grid, 2, return_cov=True)
    got = {}
    for name, env in (("i8", {}), ("fallback", {"AGPN_I8_TEST_WRAP": "1"}), ("dmma", {"AGPN_PREDICT": "dmma"})):
        for k_, v_ in env.items():
            monkeypatch.setenv(k_, v_)
        net, _ = _net(700, 2, 2, seed=6)
        o = synth.objects_from_row(node, weight, mean, theta, 2, 2)
        with contextlib.redirect_stdout(io.StringIO()):
            mu, sd, cov = net.prediction_cov(o[0], o[1], o[2], o[3], grid, 2)
        assert net.predict_path() == (1 if name == "i8" else 0), name
        got[name] = (mu, cov)
        net.close()
        for k_ in env:
            monkeypatch.delenv(k_)
        assert np.max(np.abs(mu - emu)) <= 1e-9 * np.max(np.abs(emu)), name
        assert np.max(np.abs(cov - ecov)) <= 1e-9 * np.max(np.abs(np.diag(ecov))), name
    assert np.max(np.abs(got["i8"][1] - got["dmma"][1])) <= 1e-11 * np.max(np.abs(np.diag(ecov)))
    assert np.array_equal(got["fallback"][1], got["dmma"][1])


def test_sample_from_a_multi_block_covariance():
    """n = 300: three block columns, i.e. the blocked factorisation (panel solves + trailing updates) produces the factor."""
    from artgpn_b200 import node
    net, (t, _, _) = _net(300, 1, 1, seed=8)
    K = net._kernel_matrix(node.Matern52(9.0), t) * 2.5 + 0.02 * np.identity(t.size)
    z = np.random.default_rng(9).standard_normal((2, t.size))
    got = net.sample(K, t, size=2, random_state=np.random.default_rng(9))
    exp = z @ np.linalg.cholesky(K).T
    assert np.max(np.abs(got - exp)) <= 1e-10 * np.max(np.abs(exp))
    net.close()


# ---- banded digit planes: K steps / tile halves with nothing but zero digits are left out --------------------------
def _ll_with_env(monkeypatch, env, N, theta, p=3, q=2, weights="Constant", stats=False):
    from artgpn_b200 import synth, node, weight, mean
    for k in ("AGPN_I8_SKIP", "AGPN_TRSM_SKIP"):
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    net, data = _net(N, p, q)
    net.set_structure(*synth.objects_from_row(node, weight, mean, theta[0], p, q, weights)[:3])
    out = net.log_likelihood_batch(theta)
    assert np.array_equal(out, net.log_likelihood_batch(theta))
    st = net.skip_stats()
    net.close()
    for k in env:
        monkeypatch.delenv(k, raising=False)
    return (out, st, data) if stats else out


def test_zero_chunk_skipping_is_bit_identical_to_the_dense_schedule(monkeypatch):
    """Short correlation lengths (the BASELINE workload): most K steps of the trailing update and about two fifths of the
    panel-solve tile halves hold only zero digits and are left out.  Leaving out K steps is exact; the panel-solve bound
    sits 2^-40 below one digit unit, and the log-likelihoods come out bit-identical to the dense schedule."""
    from artgpn_b200 import synth
    N = 1900
    theta = synth.make_walkers(6, p=3, q=2, seed=1)
    dense, st_d, (t, y, sig) = _ll_with_env(monkeypatch, {"AGPN_I8_SKIP": "0"}, N, theta, stats=True)
    assert st_d[1] == st_d[0] and st_d[3] == 0
    usk, st_u, _ = _ll_with_env(monkeypatch, {"AGPN_TRSM_SKIP": "0"}, N, theta, stats=True)
    assert np.array_equal(usk, dense)
    assert st_u[0] == st_d[0] and st_u[1] < 0.5 * st_u[0] and st_u[3] == 0, st_u
    both, st_b, _ = _ll_with_env(monkeypatch, {}, N, theta, stats=True)
    assert np.array_equal(both, dense)
    assert st_b[1] == st_u[1], (st_b, st_u)                    # the halves left out had no non-zero digit anyway
    assert 0.2 * st_b[2] < st_b[3] < 0.6 * st_b[2], st_b
    m = _Spec()
    exp = oracle.log_likelihood(t, y, sig, 2, *synth.objects_from_row(m, m, m, theta[1], 3, 2))
    assert abs(both[1] - exp) <= 1e-9 * abs(exp)


def test_nothing_is_skipped_when_the_covariance_does_not_decay(monkeypatch):
    """Periodic nodes never decay: every chunk holds non-zero digits, the masks are all ones, no work is left out."""
    from artgpn_b200 import synth, node, weight, mean
    N = 1100
    net, (t, y, sig) = _net(N, 2, 2)
    nodes = [node.Periodic(31.0, 1.1), node.Periodic(47.0, 0.9)]
    weights = [weight.Constant(1.3), weight.Constant(0.8), weight.Constant(0.7), weight.Constant(1.1)]
    means = [mean.Constant(0.1), mean.Constant(-0.2)]
    jitters = [0.3, 0.4]
    ll = net.log_likelihood(nodes, weights, means, jitters)
    st = net.skip_stats()
    net.close()
    assert st[1] == st[0] and st[3] == 0, st
    m = _Spec()
    exp = oracle.log_likelihood(t, y, sig, 2, [m.Periodic(31.0, 1.1), m.Periodic(47.0, 0.9)],
                                [m.Constant(1.3), m.Constant(0.8), m.Constant(0.7), m.Constant(1.1)],
                                [m.Constant(0.1), m.Constant(-0.2)], jitters)
    assert abs(ll - exp) <= 1e-9 * abs(exp)


def test_skipping_with_sub_batch_streams(monkeypatch):
    """144 matrices run as several sub-batches on their own streams, each with its slice of the masks and inverse norms:
    same bits as the dense schedule."""
    from artgpn_b200 import synth
    N = 700
    theta = synth.make_walkers(48, p=3, q=2, seed=3)
    out, st, _ = _ll_with_env(monkeypatch, {}, N, theta, stats=True)
    dense = _ll_with_env(monkeypatch, {"AGPN_I8_SKIP": "0"}, N, theta)
    assert np.array_equal(out, dense)
    assert st[1] < st[0]


def test_prediction_with_and_without_zero_chunk_skipping(monkeypatch):
    """The stacked [K ; K*] factorisation uses the same masks and bound test (K* is banded too: a grid point far from
    every training time adds nothing).  Mean and variance agree with the dense schedule far below the 1e-9 gate, and
    with the oracle."""
    from artgpn_b200 import node, weight, mean
    N, M = 1100, 5000
    res = {}
    for skip in ("1", "0"):
        monkeypatch.setenv("AGPN_I8_SKIP", skip)
        net, (t, y, sig) = _net(N, 2, 2, seed=4)
        nodes = [node.QuasiPeriodic(18.0, 23.0, 0.8), node.SquaredExponential(30.0)]
        ws = [weight.Constant(1.4), weight.Constant(0.8), weight.Constant(1.0), weight.Constant(1.2)]
        ms = [mean.Linear(0.01, 0.2), mean.Constant(0.1)]
        jit = [0.5, 0.7]
        grid = np.linspace(t.min() - 5, t.max() + 5, M)
        with contextlib.redirect_stdout(io.StringIO()):
            mu, sd, _ = net.prediction(nodes=nodes, weights=ws, means=ms, jitters=jit, time=grid, dataset=2)
        assert net.predict_path() == 1
        res[skip] = (mu, sd)
        net.close()
    monkeypatch.delenv("AGPN_I8_SKIP", raising=False)
    mu1, sd1 = res["1"]
    mu0, sd0 = res["0"]
    assert np.max(np.abs(mu1 - mu0)) <= 1e-14 * np.max(np.abs(mu0))
    assert np.max(np.abs(sd1 ** 2 - sd0 ** 2)) <= 1e-14 * np.max(sd0 ** 2)
    idx = np.arange(0, M, 25)
    emu, esd = oracle.prediction(t, y, sig, 2, [("QuasiPeriodic", [18.0, 23.0, 0.8]), ("SquaredExponential", [30.0])],
                                 [("Constant", [1.4]), ("Constant", [0.8]), ("Constant", [1.0]), ("Constant", [1.2])],
                                 [("Linear", [0.01, 0.2]), ("Constant", [0.1])], jit, grid, 2)
    assert np.max(np.abs(mu1 - emu)) <= 1e-9 * np.max(np.abs(emu))
    assert np.max(np.abs(sd1 ** 2 - esd ** 2)) <= 1e-9 * np.max(esd ** 2)
    assert idx.size


def test_results_do_not_depend_on_the_batch_size_or_launch_mode():
    """Small batches split every panel solve over two CTAs, use look-ahead and six sub-batch streams; large ones one CTA per
    tile and three or four streams.  A walker's log-likelihood has the same bits in all of them."""
    from artgpn_b200 import synth, node, weight, mean
    N = 1100
    theta = synth.make_walkers(50, p=3, q=2, seed=5)
    net, _ = _net(N, 3, 2)
    net.set_structure(*synth.objects_from_row(node, weight, mean, theta[0], 3, 2)[:3])
    big = net.log_likelihood_batch(theta)              # 150 matrices
    for W in (1, 5, 16, 33):
        assert np.array_equal(net.log_likelihood_batch(theta[:W]), big[:W]), W
    net.close()
