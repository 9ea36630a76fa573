"""CPU oracle for the artgpn likelihood / prediction hot path.

TEST INFRASTRUCTURE ONLY.  This module is a numpy/scipy restatement of the reference
algorithm (``/root/reference/artgpn``, cited per function as ``file:line``).  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker / reported CPU baseline -- the product package
``artgpn_b200`` never imports anything from ``oracle/`` and has no CPU fallback.

Parity pin: the reference ships no tests or golden vectors, so this oracle is
pinned against outputs of the *unmodified reference executed in the build container*
(``tests/golden/make_golden.py`` imports ``/root/reference`` and writes
``tests/golden/*.json``/``*.npz``; ``tests/test_oracle_golden.py`` replays them).

Kernels and means are described by ``(name, pars)`` pairs, or by any object that has a class
name and a ``.pars`` sequence (the reference's and the product's kernel classes both qualify:
``type(k)(*k.pars)`` is how the reference itself rebuilds them, ``arttwo.py:273-277``).
"""
import numpy as np
from scipy.linalg import cho_factor, cho_solve, solve_triangular, LinAlgError

__all__ = [
    "spec", "node_matrix", "weight_matrix", "node_eval", "weight_eval", "mean_eval", "mean_vector",
    "covariance_block", "compute_matrix", "log_likelihood", "prediction", "prediction_reference_form",
    "art_weights", "art_log_likelihood", "art_prediction",
]

_NONSTATIONARY = ("Linear", "Polynomial")          # arttwo.py:58,66 isinstance test


def spec(k):
    """(name, pars) from a tuple or from an object with ``.pars`` (arttwo.py:73-75)."""
    if k is None:
        return None
    if isinstance(k, tuple) and len(k) == 2 and isinstance(k[0], str):
        return k[0], [float(x) for x in k[1]]
    if type(k).__name__ == "Sum":                  # mean.py:42-65
        return "Sum", [spec(k.m1), spec(k.m2)]
    return type(k).__name__, [float(x) for x in k.pars]


# --------------------------------------------------------------------------------------------
# node kernels (unit amplitude) -- node.py
# --------------------------------------------------------------------------------------------
def _shape_kernel(name, h, r, t1, t2):
    """Amplitude-free part shared by node.py and weight.py. ``h`` = hyper-parameters after the
    weight amplitude has been stripped.  Returns None for kinds handled by the caller."""
    if name == "SquaredExponential":               # node.py:91-92, weight.py:141-142
        (ell,) = h
        return np.exp(-0.5 * r**2 / ell**2)
    if name == "Periodic":                         # node.py:112-113, weight.py:188-189
        P, ell = h
        return np.exp(-2 * np.sin(np.pi * np.abs(r) / P) ** 2 / ell**2)
    if name == "QuasiPeriodic":                    # node.py:137-139, weight.py:259-261
        ell_e, P, ell_p = h
        return np.exp(-2 * np.sin(np.pi * np.abs(r) / P) ** 2 / ell_p**2 - r**2 / (2 * ell_e**2))
    if name == "RationalQuadratic":                # node.py:159-160, weight.py:347-348
        alpha, ell = h
        return 1 / (1 + r**2 / (2 * alpha * ell**2)) ** alpha
    if name == "Cosine":                           # node.py:207-208, weight.py:529-530
        (P,) = h
        return np.cos(2 * np.pi * np.abs(r) / P)
    if name in ("Exponential", "Laplacian"):       # node.py:226-227, weight.py:573-574,618-619
        (ell,) = h
        return np.exp(-np.abs(r) / ell)
    if name == "Matern32":                         # node.py:246-248, weight.py:663-665
        (ell,) = h
        return (1.0 + np.sqrt(3.0) * np.abs(r) / ell) * np.exp(-np.sqrt(3.0) * np.abs(r) / ell)
    if name == "Matern52":                         # node.py:267-270, weight.py:712-715
        (ell,) = h
        return (1.0 + (3 * np.sqrt(5) * ell * np.abs(r) + 5 * np.abs(r) ** 2) / (3 * ell**2)) \
            * np.exp(-np.sqrt(5.0) * np.abs(r) / ell)
    if name == "GammaExp":                         # node.py:307-308, weight.py:809-810
        gamma, ell = h
        return np.exp(-(np.abs(r) / ell) ** gamma)
    if name == "Linear":                           # node.py:287-288, weight.py:764-765
        (c,) = h
        return (t1 - c) * (t2 - c)
    if name == "Polynomial":                       # node.py:326-327, weight.py:871-872
        a, b, c = h
        return (a * t1 * t2 + b) ** c
    return None


def _white_noise(wn, shape):
    """node.py:68-72 / weight.py:106-110: wn^2 on the diagonal iff the lag array is square."""
    out = np.zeros(shape)
    if len(shape) == 2 and shape[0] == shape[1]:
        out[np.diag_indices(shape[0])] = wn**2
    return out


def node_eval(k, r=None, t1=None, t2=None):
    """Value of a node kernel on a 2-D lag array ``r`` (or on broadcastable ``t1, t2`` for the
    non-stationary kinds).  node.py:36-327."""
    name, h = spec(k)
    if name in _NONSTATIONARY:
        t1, t2 = np.asarray(t1, float), np.asarray(t2, float)
        return _shape_kernel(name, h, None, t1, t2) + 0.0 * (t1 + t2)
    r = np.asarray(r, float)
    if name == "Constant":                         # node.py:50-51 (c, NOT squared)
        return h[0] * np.ones_like(r)
    if name == "WhiteNoise":
        return _white_noise(h[0], r.shape)
    if name == "RQP":                              # node.py:186-189 (sign/abs trick)
        alpha, ell_e, P, ell_p = h
        a = np.exp(-2 * np.sin(np.pi * np.abs(r) / P) ** 2 / ell_p**2)
        b = 1 + r**2 / (2 * alpha * ell_e**2)
        return a / (np.sign(b) * np.abs(b) ** alpha)
    out = _shape_kernel(name, h, r, None, None)
    if out is None:
        raise ValueError("unknown node kernel %r" % name)
    return out


def weight_eval(k, r=None, t1=None, t2=None):
    """Value of a weight kernel: ``weight**2`` times the shape, weight first in ``.pars``.
    weight.py:65-872 (value classes only)."""
    name, h = spec(k)
    if name in _NONSTATIONARY:
        t1, t2 = np.asarray(t1, float), np.asarray(t2, float)
        return h[0] ** 2 * _shape_kernel(name, h[1:], None, t1, t2) + 0.0 * (t1 + t2)
    r = np.asarray(r, float)
    if name == "Constant":                         # weight.py:78-79
        return h[0] ** 2 * np.ones_like(r)
    if name == "WhiteNoise":
        return _white_noise(h[0], r.shape)
    if name == "RQP":                              # weight.py:420-423 (plain power)
        w, alpha, ell_e, P, ell_p = h
        return w**2 * np.exp(-2 * np.sin(np.pi * np.abs(r) / P) ** 2 / ell_p**2) \
            / (1 + r**2 / (2 * alpha * ell_e**2)) ** alpha
    out = _shape_kernel(name, h[1:], r, None, None)
    if out is None:
        raise ValueError("unknown weight kernel %r" % name)
    return h[0] ** 2 * out


def _matrix(evalf, k, ta, tb):
    """arttwo.py:53-71: stationary kernels see r = ta[:,None]-tb[None,:]; Linear/Polynomial see
    (None, ta[:,None], tb[None,:])."""
    ta, tb = np.asarray(ta, float), np.asarray(tb, float)
    name, _ = spec(k)
    if name in _NONSTATIONARY:
        return evalf(k, None, ta[:, None], tb[None, :])
    return evalf(k, ta[:, None] - tb[None, :])


def node_matrix(k, ta, tb=None):
    return _matrix(node_eval, k, ta, ta if tb is None else tb)


def weight_matrix(k, ta, tb=None):
    return _matrix(weight_eval, k, ta, ta if tb is None else tb)


# --------------------------------------------------------------------------------------------
# mean functions -- mean.py
# --------------------------------------------------------------------------------------------
def mean_eval(m, t):
    """mean.py:68-202.  ``None`` means zero (arttwo.py:112-113)."""
    t = np.atleast_1d(np.asarray(t, float))
    if m is None:
        return np.zeros_like(t)
    name, h = spec(m)
    if name == "Sum":                              # mean.py:63-65
        return mean_eval(h[0], t) + mean_eval(h[1], t)
    if name == "Constant":                         # mean.py:77-78
        return np.full(t.shape, h[0])
    if name == "Linear":                           # mean.py:91-93 (centred on mean of the grid passed)
        return h[0] * (t - t.mean()) + h[1]
    if name in ("Parabola", "Cubic"):              # mean.py:105-121 (polyval, highest power first)
        out = np.zeros_like(t)
        for c in h:
            out = out * t + c
        return out
    if name == "Sine":                             # mean.py:133-136
        return h[0] * np.sin(2 * np.pi * t / h[1] + h[2]) + h[3]
    if name == "Cosine":                           # mean.py:148-151 (amp squared)
        return h[0] ** 2 * np.cos(2 * np.pi * t / h[1] + h[2]) + h[3]
    if name == "Keplerian":                        # mean.py:173-202
        P, K, e, w, phi = h
        T0 = t[0] - P * phi / (2.0 * np.pi)
        M0 = 2 * np.pi * (t - T0) / P
        E0 = M0 + e * np.sin(M0) + 0.5 * e**2 * np.sin(2 * M0)
        M1 = E0 - e * np.sin(E0) - M0
        if np.any(np.abs(M1) > 1e-10):             # exactly one correction pass (mean.py:184-198)
            M1p = 1 - e * np.cos(E0)
            M1pp = e * np.sin(E0)
            M1ppp = 1 - M1p
            d1 = -M1 / M1p
            d2 = -M1 / (M1p + d1 * M1pp / 2.0)
            d3 = -M1 / (M1p + d2 * M1pp / 2.0 + d2 * d2 * M1ppp / 6.0)
            E0 = E0 + d3
        nu = 2 * np.arctan(np.sqrt((1 + e) / (1 - e)) * np.tan(E0 / 2))
        return K * (e * np.cos(w) + np.cos(w + nu))
    raise ValueError("unknown mean %r" % name)


def mean_vector(means, t, p):
    """arttwo.py:106-126: [p, len(t)] table of mean values (rows of None stay 0)."""
    t = np.asarray(t, float)
    out = np.zeros((p, t.size))
    for i, m in enumerate(means):
        if m is not None:
            out[i] = mean_eval(m, t)
    return out


# --------------------------------------------------------------------------------------------
# network -- arttwo.py
# --------------------------------------------------------------------------------------------
def covariance_block(time, nodes, weights, series, q, ta=None, tb=None, yerr=None, add_errors=False):
    """arttwo.py:151-192: K = sum_j W_{series,j} o F_j on grid ta x tb (0-based ``series``;
    weight index j + q*series, arttwo.py:183)."""
    ta = np.asarray(time if ta is None else ta, float)
    tb = ta if tb is None else np.asarray(tb, float)
    K = np.zeros((ta.size, tb.size))
    for j in range(q):
        K += weight_matrix(weights[j + q * series], ta, tb) * node_matrix(nodes[j], ta, tb)
    if add_errors:
        K += np.diag(np.asarray(yerr[series], float) ** 2)
    return K


def compute_matrix(time, yerr, q, nodes, weights, nugget=False, shift=False):
    """arttwo.py:194-238: block-diagonal pN x pN matrix; sigma (not sigma^2) on the diagonal,
    no jitter; optional nugget / shift."""
    yerr = np.asarray(yerr, float)
    p, N = yerr.shape
    K = np.zeros((p * N, p * N))
    for i in range(p):
        K[i * N:(i + 1) * N, i * N:(i + 1) * N] = covariance_block(time, nodes, weights, i, q)
    K = K + np.diag(yerr.reshape(-1))
    if nugget:
        K = (1 - 0.01) * K + 0.01 * np.diag(np.diag(K))
    if shift:
        K = K + 0.01 * np.identity(p * N)
    return K


def log_likelihood(time, y, yerr, q, nodes, weights, means, jitters):
    """arttwo.py:240-290.  Returns -inf as soon as one block is not positive definite."""
    time = np.asarray(time, float)
    y = np.asarray(y, float)
    yerr = np.asarray(yerr, float)
    p, N = y.shape
    resid = y - mean_vector(means, time, p) if means else y    # arttwo.py:263
    ll = 0.0
    for i in range(p):
        K = covariance_block(time, nodes, weights, i, q)
        K[np.diag_indices(N)] += yerr[i] ** 2 + jitters[i] ** 2
        try:
            U = cho_factor(K, overwrite_a=True, lower=False)   # arttwo.py:284
        except LinAlgError:
            return -np.inf
        ll += -0.5 * resid[i] @ cho_solve(U, resid[i]) - np.sum(np.log(np.diag(U[0]))) \
            - 0.5 * N * np.log(2 * np.pi)
    return ll


def prediction(time, y, yerr, q, nodes, weights, means, jitters, tstar, dataset=1):
    """arttwo.py:329-396 restated in the diagonal-only form: the reference
    builds two M x M matrices and keeps only their diagonal.  Returns (mean, std)."""
    time = np.asarray(time, float)
    tstar = np.asarray(tstar, float)
    y = np.asarray(y, float)
    yerr = np.asarray(yerr, float)
    p, N = y.shape
    d = dataset - 1
    resid = y - mean_vector(means, time, p) if means else y
    K = covariance_block(time, nodes, weights, d, q)
    K[np.diag_indices(N)] += yerr[d] ** 2 + jitters[d] ** 2
    L = np.linalg.cholesky(K)
    alpha = cho_solve((L, True), resid[d])
    Ks = covariance_block(time, nodes, weights, d, q, ta=tstar, tb=time)          # [M, N]
    mu = Ks @ alpha
    if means:
        mu = mu + mean_vector(means, tstar, p)[d]
    # diag of K** : every kernel evaluated at zero lag on a 1x1 (square) grid, arttwo.py:385-387
    kss = np.zeros(tstar.size)
    for j in range(q):
        wk, nk = weights[j + q * d], nodes[j]
        wv = np.array([weight_matrix(wk, tstar[m:m + 1])[0, 0] for m in range(tstar.size)]) \
            if spec(wk)[0] in _NONSTATIONARY else np.full(tstar.size, weight_matrix(wk, tstar[:1])[0, 0])
        nv = np.array([node_matrix(nk, tstar[m:m + 1])[0, 0] for m in range(tstar.size)]) \
            if spec(nk)[0] in _NONSTATIONARY else np.full(tstar.size, node_matrix(nk, tstar[:1])[0, 0])
        kss += wv * nv
    V = solve_triangular(L, Ks.T, lower=True)
    var = kss + jitters[d] ** 2 - np.sum(V * V, axis=0)
    return mu, np.sqrt(var)


def prediction_reference_form(time, y, yerr, q, nodes, weights, means, jitters, tstar, dataset=1,
                              train_jitter=None, return_cov=False):
    """arttwo.py:367-395 as written (M x M intermediates); small M only.  Used to show that the
    diagonal-only form above is the same function.  train_jitter / return_cov cover the legacy
    art.network.prediction (art.py:318-354: constructor jitter in the training block, full y_cov)."""
    time = np.asarray(time, float)
    tstar = np.asarray(tstar, float)
    y = np.asarray(y, float)
    yerr = np.asarray(yerr, float)
    p, N = y.shape
    d = dataset - 1
    resid = y - mean_vector(means, time, p) if means else y
    K = covariance_block(time, nodes, weights, d, q)
    K[np.diag_indices(N)] += yerr[d] ** 2 + (jitters[d] if train_jitter is None else train_jitter) ** 2
    U = cho_factor(K)
    sol = cho_solve(U, resid[d])
    Ks = covariance_block(time, nodes, weights, d, q, ta=tstar, tb=time)
    Kss = covariance_block(time, nodes, weights, d, q, ta=tstar) + jitters[d] ** 2 * np.identity(tstar.size)
    mu = Ks @ sol + (mean_vector(means, tstar, p)[d] if means else 0.0)
    quad = np.array([Ks @ cho_solve(U, Ks[i, :]) for i in range(tstar.size)])
    if return_cov:
        return mu, np.sqrt(np.diag(Kss - quad)), Kss - quad
    return mu, np.sqrt(np.diag(Kss - quad))


def art_weights(q, p, weight_spec, weights_values):
    """Legacy art.network: one weight class, amplitude per (series, node) from weights_values
    (art.py:253-261).  weight_spec = (name, pars) of weights[0]."""
    name, pars = spec(weight_spec)
    return [(name, [float(weights_values[j + q * i])] + list(pars[1:])) for i in range(p) for j in range(q)]


def art_log_likelihood(time, y, yerr, q, nodes, weight_spec, weights_values, means, jitters):
    """art.py:228-276: the constructor squared yerr (art.py:58) and the likelihood squares it again."""
    y = np.asarray(y, float)
    return log_likelihood(time, y, np.asarray(yerr, float) ** 2, q, nodes,
                          art_weights(q, y.shape[0], weight_spec, weights_values), means, jitters)


def art_prediction(time, y, yerr, q, nodes, weight_spec, weights_values, means, jitters, ctor_jitters, tstar, dataset=1):
    """art.py:280-354 -> (mean, std, cov)."""
    y = np.asarray(y, float)
    return prediction_reference_form(time, y, np.asarray(yerr, float) ** 2, q, nodes,
                                     art_weights(q, y.shape[0], weight_spec, weights_values), means, jitters, tstar,
                                     dataset, train_jitter=ctor_jitters[dataset - 1], return_cov=True)
