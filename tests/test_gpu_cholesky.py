"""GPU: the batched blocked Cholesky building block (agpn_potrf_batched) on its own, vs
numpy.linalg on the same matrices; plus the size-independent residual property at full size."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(A_host, rhs_host):
    import torch
    from artgpn_b200 import _lib
    lib = _lib.load()
    nmat, n, _ = A_host.shape
    A = torch.from_numpy(A_host).cuda()
    rhs = torch.from_numpy(rhs_host).cuda()
    logdet = torch.zeros(nmat, dtype=torch.float64, device="cuda")
    quad = torch.zeros(nmat, dtype=torch.float64, device="cuda")
    info = torch.zeros(nmat, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    _lib.check(lib.agpn_potrf_batched(0, nmat, n, A.data_ptr(), rhs.data_ptr(), logdet.data_ptr(), quad.data_ptr(),
                                      info.data_ptr(), None))
    torch.cuda.synchronize()
    return A.cpu().numpy(), rhs.cpu().numpy(), logdet.cpu().numpy(), quad.cpu().numpy(), info.cpu().numpy()


def _spd(rng, nmat, n, cond_jitter):
    X = rng.normal(size=(nmat, n, n // 2))
    return X @ X.transpose(0, 2, 1) + cond_jitter * np.eye(n)


@pytest.fixture(params=["i8", "i8:2", "i8_wide:3", "i8_narrow:4", "i8:3:pertile", "cpasync", "cpasync:2", "tma", "tma_nomc"])
def syrk_variant(request, monkeypatch):
    """Trailing-update kernel (AGPN_SYRK[:AGPN_GROUP], read per call here): the INT8-sliced tcgen05 kernel
    (default; left-looking, or with groups so that narrow and wide updates both run on it, or only one of
    them; persistent clusters or one CTA pair per tile), the FP64 cp.async + mbarrier DMMA kernel, or the TMA (cp.async.bulk.tensor) producer-warp DMMA
    kernel with / without cluster multicast."""
    kind, group, pertile = (request.param.split(":") + ["", ""])[:3]
    monkeypatch.setenv("AGPN_SYRK", kind)
    if group:
        monkeypatch.setenv("AGPN_GROUP", group)
    else:
        monkeypatch.delenv("AGPN_GROUP", raising=False)
    if pertile:
        monkeypatch.setenv("AGPN_I8_PERSIST", "0")      # one CTA pair per tile instead of persistent clusters
    else:
        monkeypatch.delenv("AGPN_I8_PERSIST", raising=False)
    return request.param


@pytest.mark.parametrize("n,nmat", [(128, 5), (256, 3), (640, 2), (1408, 2)])
def test_potrf_batched_vs_numpy(n, nmat, syrk_variant):
    rng = np.random.default_rng(n)
    K = _spd(rng, nmat, n, 1e-2 * n)
    r = rng.normal(size=(nmat, n))
    L, z, logdet, quad, info = _run(K.copy(), r.copy())
    assert np.all(info == 0)
    for m in range(nmat):
        Lref = np.linalg.cholesky(K[m])
        assert np.allclose(np.tril(L[m]), Lref, rtol=1e-9, atol=1e-11 * np.max(np.abs(Lref)))
        zref = np.linalg.solve(Lref, r[m])
        assert np.allclose(z[m], zref, rtol=1e-8, atol=1e-10 * np.max(np.abs(zref)))
        assert abs(logdet[m] - np.sum(np.log(np.diag(Lref)))) <= 1e-11 * abs(logdet[m])
        assert abs(quad[m] - zref @ zref) <= 1e-9 * abs(quad[m])


@pytest.mark.parametrize("mode", ["i8", "dmma"])
def test_potrf_flags_not_positive_definite_only_where_it_is(mode, monkeypatch):
    monkeypatch.setenv("AGPN_SYRK", mode)
    rng = np.random.default_rng(0)
    K = _spd(rng, 4, 256, 1.0)
    K[2, 200, 200] = -5.0            # breaks the second diagonal block of matrix 2 only
    r = rng.normal(size=(4, 256))
    L, z, logdet, quad, info = _run(K.copy(), r.copy())
    assert list(info) == [0, 0, 1, 0]
    for m in (0, 1, 3):
        assert np.allclose(np.tril(L[m]), np.linalg.cholesky(K[m]), rtol=1e-9, atol=1e-10)


@pytest.mark.parametrize("mode", ["i8", "i8:8", "dmma"])
def test_potrf_full_size_residual_2048(mode, monkeypatch):
    """BASELINE size (2048^2): ||L L^T - K|| / ||K|| at round-off, log-det vs numpy -- on the INT8-sliced update
    (left-looking and in groups of 8) and on the FP64 tensor instruction."""
    kind, _, group = mode.partition(":")
    monkeypatch.setenv("AGPN_SYRK", kind)
    if group:
        monkeypatch.setenv("AGPN_GROUP", group)
    else:
        monkeypatch.delenv("AGPN_GROUP", raising=False)
    rng = np.random.default_rng(1)
    n = 2048
    K = _spd(rng, 2, n, 10.0)
    L, z, logdet, quad, info = _run(K.copy(), np.zeros((2, n)))
    assert np.all(info == 0)
    for m in range(2):
        Lm = np.tril(L[m])
        res = np.linalg.norm(Lm @ Lm.T - K[m]) / np.linalg.norm(K[m])
        assert res < 1e-13, res
        sign, ld = np.linalg.slogdet(K[m])
        assert abs(2 * logdet[m] - ld) <= 1e-10 * abs(ld)


def test_i8_update_is_as_accurate_as_the_fp64_instruction(monkeypatch):
    """The INT8-sliced update carries 56 bits below each row's scale: its factor must be as close to a
    long-double reference factor as the DMMA path's, on a matrix with widely varying row scales
    (diag spread 1e-6 .. 1e6) where a per-row scale matters."""
    rng = np.random.default_rng(7)
    n = 768
    X = rng.normal(size=(n, n // 2))
    K0 = X @ X.T + 0.05 * n * np.eye(n)
    d = 10.0 ** rng.uniform(-3, 3, size=n)
    K = (K0 * d[:, None]) * d[None, :]
    Kl = K.astype(np.longdouble)
    # long-double Cholesky (column by column)
    Lr = np.zeros_like(Kl)
    for j in range(n):
        v = Kl[j:, j] - Lr[j:, :j] @ Lr[j, :j]
        Lr[j:, j] = v / np.sqrt(v[0])
    Lr = Lr.astype(np.float64)
    errs = {}
    for mode in ("i8", "dmma"):
        monkeypatch.setenv("AGPN_SYRK", mode)
        monkeypatch.delenv("AGPN_GROUP", raising=False)
        L, z, logdet, quad, info = _run(K[None].copy(), np.zeros((1, n)))
        assert info[0] == 0
        scale = np.sqrt(np.diag(K))[:, None]          # |L_ik| <= sqrt(K_ii): row-relative error
        errs[mode] = np.max(np.abs(np.tril(L[0]) - Lr) / scale)
    assert errs["i8"] < 1e-11 and errs["dmma"] < 1e-11, errs
    assert errs["i8"] <= 4.0 * errs["dmma"] + 1e-15, errs
