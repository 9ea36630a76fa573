"""CPU: the arithmetic facts the banded schedule of the GPU path rests on, checked with numpy on
the oracle's covariance blocks.  Nothing here runs the product path; it pins the reasoning:

  (1) a panel entry below 2^-57 of its row scale has ALL digits zero (the integer it is rounded to is 0), so a K step in
      which either operand's chunk holds only such entries adds integer zeros -- leaving it out is exact;
  (2) |X_ic| <= max_j |C_ij| * ||rows of Linv||_inf for X = C Linv^T (the panel solve's bound test);
  (3) the envelope sum_j w_ij^2 exp(-r_min^2 / (2 ell_j^2)) bounds every entry of an off-diagonal tile of a network of
      QuasiPeriodic / SquaredExponential / Periodic nodes with Constant or SquaredExponential weights;
  (4) on a BASELINE-shaped workload the factor really is banded at that level, and dropping the blocks the bound test
      would drop changes the log-likelihood by nothing a double can see.
"""
import numpy as np

import helpers as H
from helpers import oracle

Q = 126.0 * 2.0 ** 49           # digit unit of the INT8-sliced update: q = rint(x / r_i * Q), eight base-128 digits
TB = 128


class _Spec(object):
    def __getattr__(self, name):
        return lambda *pars: (name, list(pars))


def _digits(q):
    """balanced base-128 digits of the integer q (least significant first), as the kernels' i8_digits2 produces them"""
    out = []
    q = int(q)
    for _ in range(7):
        d = ((q + 64) % 128) - 64
        out.append(d)
        q = (q - d) // 128
    out.append(q)                # the top digit is the signed rest (|q| <= 126 2^49 keeps it within an int8)
    assert -128 <= q <= 127
    return out


def test_entries_below_half_a_digit_unit_have_no_digits():
    rng = np.random.default_rng(0)
    r = 3.7
    for x in np.concatenate([rng.uniform(-1, 1, 200) * 0.49 * r / Q, [0.5 * r / Q, -0.5 * r / Q, 0.0]]):
        q = np.rint(x / r * Q)                    # ties go to the even 0, like cvt.rni
        assert q == 0 and all(d == 0 for d in _digits(q))
    for x in (0.51 * r / Q, -0.75 * r / Q, 1.0 * r / Q, 0.999 * r):
        q = np.rint(x / r * Q)
        ds = _digits(q)
        assert q != 0 and any(d != 0 for d in ds)
        assert sum(d * 128 ** j for j, d in enumerate(ds)) == int(q)      # the digits are the integer
    # 2^-57 of the row scale is below half a unit: 2^-57 * Q = 126 / 256 < 0.5
    assert 2.0 ** -57 * Q < 0.5


def test_panel_bound_inequality():
    rng = np.random.default_rng(1)
    for trial in range(20):
        A = rng.normal(size=(TB, TB))
        L = np.linalg.cholesky(A @ A.T + TB * np.eye(TB))
        Linv = np.linalg.inv(L)
        C = rng.normal(size=(TB, TB)) * 10.0 ** rng.uniform(-30, 3)
        X = C @ Linv.T
        for h, K in ((0, 64), (1, 128)):
            rows = slice(64 * h, 64 * h + 64)
            lnorm = np.abs(Linv[rows, :K]).sum(axis=1).max()
            bound = np.abs(C[:, :K]).max(axis=1) * lnorm
            assert np.all(np.abs(X[:, rows]).max(axis=1) <= bound * (1 + 1e-12))


def _network(seed, N, p=3, q=2, weights="Constant"):
    from artgpn_b200 import synth
    t, y, sig = synth.make_data(N, p, seed=seed)
    theta = synth.make_walkers(4, p=p, q=q, weights=weights, seed=seed + 1)
    m = _Spec()
    return t, y, sig, theta, m


def test_envelope_bounds_every_tile():
    from artgpn_b200 import synth
    for weights in ("Constant", "SquaredExponential"):
        N = 768
        t, y, sig, theta, m = _network(3, N, weights=weights)
        nodes, ws, means, jit = synth.objects_from_row(m, m, m, theta[1], 3, 2, weights)
        nb = N // TB
        lo = t.reshape(nb, TB).min(axis=1)
        hi = t.reshape(nb, TB).max(axis=1)
        for s in range(3):
            K = oracle.covariance_block(t, nodes, ws, s, 2)
            for i in range(nb):
                for j in range(i):
                    rmin = max(0.0, lo[i] - hi[j], lo[j] - hi[i])
                    env = 0.0
                    for jn in range(2):
                        ell_e = nodes[jn][1][0]                         # QuasiPeriodic(ell_e, P, ell_p)
                        w = ws[jn + 2 * s]
                        amp = w[1][0] ** 2
                        c = -0.5 / ell_e ** 2
                        if w[0] == "SquaredExponential":
                            c += -0.5 / w[1][1] ** 2
                        env += amp * np.exp(c * rmin * rmin)
                    tile = np.abs(K[TB * i:TB * i + TB, TB * j:TB * j + TB]).max()
                    assert tile <= env * 1.000001, (weights, s, i, j, tile, env)


def test_factor_is_banded_and_the_dropped_blocks_do_not_matter():
    """3 x 1280 points of the bench's generator: block-eliminate K like the GPU does, drop every 128 x 64 panel half the
    bound test would drop (below 2^-96 of the row scale), and compare log-det and quadratic form with the dense result."""
    from artgpn_b200 import synth
    N = 1280
    t, y, sig, theta, m = _network(0, N)
    nodes, ws, means, jit = synth.objects_from_row(m, m, m, theta[0], 3, 2)
    nb = N // TB
    s = 1
    K = oracle.covariance_block(t, nodes, ws, s, 2)
    K[np.diag_indices(N)] += sig[s] ** 2 + jit[s] ** 2
    r = np.sqrt(np.diag(K))
    rhs = y[s].copy()
    Ld = np.linalg.cholesky(K)
    zd = np.linalg.solve(Ld, rhs)
    dense = (2.0 * np.log(np.diag(Ld)).sum(), zd @ zd)

    A = K.copy()
    z = rhs.copy()
    logdet = 0.0
    dropped = total = zero_digit_halves = 0
    for k in range(nb):
        sl = slice(TB * k, TB * k + TB)
        Lkk = np.linalg.cholesky(A[sl, sl])
        Linv = np.linalg.inv(Lkk)
        logdet += 2.0 * np.log(np.diag(Lkk)).sum()
        z[sl] = Linv @ z[sl]
        X = np.zeros((N - TB * (k + 1), TB))
        for i in range(k + 1, nb):
            ri = slice(TB * i, TB * i + TB)
            C = A[ri, sl]
            for h, Kc in ((0, 64), (1, 128)):
                cols = slice(64 * h, 64 * h + 64)
                lnorm = np.abs(Linv[cols, :Kc]).sum(axis=1).max()
                bound = (np.abs(C[:, :Kc]).max(axis=1) * lnorm * Q / r[ri]).max()
                Xh = C[:, :Kc] @ Linv[cols, :Kc].T
                total += 1
                if np.all(np.rint(Xh * (Q / r[ri])[:, None]) == 0):
                    zero_digit_halves += 1
                if bound < 2.0 ** -40:
                    dropped += 1
                    assert np.all(np.rint(Xh * (Q / r[ri])[:, None]) == 0)       # what is dropped has no digits
                    continue
                X[TB * (i - k - 1):TB * (i - k), cols] = Xh
        z[TB * (k + 1):] -= X @ z[sl]
        A[TB * (k + 1):, TB * (k + 1):] -= X @ X.T
    got = (logdet, z @ z)
    assert dropped > 0.15 * total, (dropped, total)                  # the workload is banded ...
    assert dropped <= zero_digit_halves                               # ... the bound test is conservative ...
    assert abs(got[0] - dense[0]) <= 1e-13 * abs(dense[0])            # ... and what it drops cannot be seen
    assert abs(got[1] - dense[1]) <= 1e-12 * abs(dense[1])
