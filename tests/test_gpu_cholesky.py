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


@pytest.fixture(params=["cpasync", "tma", "tma_nomc"])
def syrk_variant(request, monkeypatch):
    """Trailing-update kernel: cp.async + mbarrier pipeline (default) or the TMA (cp.async.bulk.tensor)
    producer-warp kernel with / without cluster multicast (AGPN_SYRK, read per call here)."""
    monkeypatch.setenv("AGPN_SYRK", request.param)
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


def test_potrf_flags_not_positive_definite_only_where_it_is():
    rng = np.random.default_rng(0)
    K = _spd(rng, 4, 256, 1.0)
    K[2, 200, 200] = -5.0            # breaks the second diagonal block of matrix 2 only
    r = rng.normal(size=(4, 256))
    L, z, logdet, quad, info = _run(K.copy(), r.copy())
    assert list(info) == [0, 0, 1, 0]
    for m in (0, 1, 3):
        assert np.allclose(np.tril(L[m]), np.linalg.cholesky(K[m]), rtol=1e-9, atol=1e-10)


def test_potrf_full_size_residual_2048():
    """BASELINE size (2048^2): ||L L^T - K|| / ||K|| at round-off, log-det vs numpy."""
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
