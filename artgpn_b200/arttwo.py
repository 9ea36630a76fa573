"""GP network -- drop-in for ``artgpn.arttwo.network`` on the likelihood / prediction path.

Same constructor and method signatures as the reference class (arttwo.py:13-396); the work
is done by ``libartgpn_b200.so`` on a B200 (``include/artgpn_b200.h``).  In addition to the
reference API there is a batched entry point, ``log_likelihood_batch(theta)``, that evaluates
many parameter rows (emcee walkers) in one call -- the replacement for the reference's
``multiprocessing.Pool`` over walkers (examples/Sun/01_emcee.py:119-122).

Nothing here computes on the CPU: without the CUDA library / a GPU every method raises.
"""
from copy import copy

import numpy as np

from . import _lib
from ._kernelbase import KernelFunction, get_device
from .kinds import KERNELS, FAMILY_NODE, FAMILY_WEIGHT
from . import node as _node
from . import weight as _weight

STATUS_OK, STATUS_NOT_PD, STATUS_NOT_FINITE = 0, 1, 2


def _kernel_kind(k, family):
    """Kernel id of an object: product classes carry it; any object whose class name matches a
    reference kernel class (and has ``.pars``) is accepted too, because the reference only ever
    uses ``type(k)`` and ``k.pars`` (arttwo.py:273-277)."""
    if isinstance(k, KernelFunction):
        if k._family != family:
            raise TypeError("%r is a %s kernel, expected a %s kernel" % (
                k, "weight" if k._family else "node", "weight" if family else "node"))
        return k._kind
    name = type(k).__name__
    mod = _weight if family == FAMILY_WEIGHT else _node
    cls = getattr(mod, name, None)
    if cls is None or not hasattr(k, "pars"):
        raise TypeError("unsupported %s kernel object %r" % ("weight" if family else "node", k))
    return cls._kind


def _mean_terms(m):
    if m is None:
        return []
    if hasattr(m, "terms"):
        return m.terms()
    # duck-typed reference object
    name = type(m).__name__
    if name == "Sum":
        return _mean_terms(m.m1) + _mean_terms(m.m2)
    from .kinds import MEANS
    if name not in MEANS:
        raise TypeError("unsupported mean function object %r" % (m,))
    return [(MEANS[name][0], [float(x) for x in m.pars])]


class Structure(object):
    """Kernel / mean classes of a network plus the theta-row layout they imply
    (include/artgpn_b200.h: nodes | weights in list order j + q*i | mean terms | jitters)."""

    def __init__(self, p, q, nodes, weights, means):
        if len(nodes) < q:
            raise IndexError("list index out of range")          # as nodes[j - 1] would (arttwo.py:273)
        self.p, self.q = p, q
        self.node_kinds = [_kernel_kind(nodes[j], FAMILY_NODE) for j in range(q)]
        # weights[j + q*i] for i < p, j < q: too few weights -> IndexError like the reference
        self.weight_kinds = [_kernel_kind(weights[j + q * i], FAMILY_WEIGHT) for i in range(p) for j in range(q)]
        self.use_means = bool(means)                              # arttwo.py:263 `if means`
        self.mean_nterms, self.mean_kinds = [], []
        if self.use_means:
            ms = list(means)
            for i in range(p):
                terms = _mean_terms(ms[i]) if i < len(ms) else []
                self.mean_nterms.append(len(terms))
                self.mean_kinds += [t[0] for t in terms]
        else:
            self.mean_nterms = [0] * p
        self.key = (tuple(self.node_kinds), tuple(self.weight_kinds), tuple(self.mean_nterms),
                    tuple(self.mean_kinds), self.use_means)
        lib = _lib.load()
        self.D = sum(lib.agpn_kernel_npars(FAMILY_NODE, k) for k in self.node_kinds) \
            + sum(lib.agpn_kernel_npars(FAMILY_WEIGHT, k) for k in self.weight_kinds) \
            + sum(lib.agpn_mean_npars(k) for k in self.mean_kinds) + p

    def pack(self, nodes, weights, means, jitters):
        """One theta row from object lists (values from ``.pars``, the single source of truth)."""
        row = []
        for j in range(self.q):
            row += [float(x) for x in nodes[j].pars]
        for i in range(self.p):
            for j in range(self.q):
                row += [float(x) for x in weights[j + self.q * i].pars]
        if self.use_means:
            ms = list(means)
            for i in range(self.p):
                for _, pars in (_mean_terms(ms[i]) if i < len(ms) else []):
                    row += pars
        jit = [float(jitters[i]) for i in range(self.p)]
        row += jit
        out = np.array(row, dtype=np.float64)
        if out.size != self.D:
            raise ValueError("parameter count %d does not match the structure (%d)" % (out.size, self.D))
        return out


class network(object):
    """
    Artificial Gaussian-process network (arttwo.py:13-51).

    Parameters
    ----------
    num_nodes: int
        Number of node functions used
    time: array
        Time measurements
    *args: arrays
        data1, data1_error, data2, data2_error, ...
    device: int, keyword only
        CUDA device index (default: LOCAL_RANK or 0)
    exact: bool, keyword only
        see set_exact()
    """

    def __init__(self, num_nodes, time, *args, device=None, exact=False):
        self.q = num_nodes
        self.time = time
        self.args = args
        self.p = int(len(self.args) / 2)
        self.qp = self.q * self.p
        self.tt = np.tile(time, self.p)
        self.y = np.array([a for i, a in enumerate(args) if i % 2 == 0])
        self.yerr = np.array([a for i, a in enumerate(args) if i % 2 == 1])
        assert self.p >= 1 and len(self.y) == len(self.yerr), \
            'Given data and number of components dont match'
        self._device = get_device() if device is None else int(device)
        self._exact = bool(exact)
        self._handle = None
        self._structure_key = None
        self._structure = None

    # ---- device context ---------------------------------------------------------------------
    def _ctx(self):
        if self._handle is None:
            lib = _lib.load()
            t = np.ascontiguousarray(self.time, dtype=np.float64)
            y = np.ascontiguousarray(self.y, dtype=np.float64)
            e = np.ascontiguousarray(self.yerr, dtype=np.float64)
            if y.shape != (self.p, t.size) or e.shape != y.shape:
                raise ValueError("every data / error array must have the length of `time`")
            h = _lib.c_void_p()
            import ctypes
            _lib.check(lib.agpn_create(ctypes.byref(h), self._device, t.size, self.p, self.q,
                                       _lib.ptr(t), _lib.ptr(y), _lib.ptr(e)))
            self._handle = h
            if self._exact:
                _lib.check(lib.agpn_set_exact(h, 1))
        return self._handle

    def close(self):
        if getattr(self, "_handle", None) is not None:
            try:
                _lib.load().agpn_destroy(self._handle)
            finally:
                self._handle = None
                self._structure_key = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_exact(self, enable=True):
        """Evaluate every covariance entry in the reference's own operation order (last-ulp agreement with
        numpy; ~4 x slower assembly) instead of the hoisted / tabulated fast form (~1e-13 relative).  Only
        matters for ill-conditioned covariances, where differences in K are amplified by cond(K)."""
        self._exact = bool(enable)
        if self._handle is not None:
            _lib.check(_lib.load().agpn_set_exact(self._handle, 1 if enable else 0))

    def set_structure(self, nodes, weights, means=None):
        """Fix the kernel / mean classes for ``log_likelihood_batch``; returns the Structure
        (``.D`` = doubles per theta row, ``.pack(...)`` builds a row from objects)."""
        s = Structure(self.p, self.q, nodes, weights, means)
        self._apply_structure(s)
        return s

    def _apply_structure(self, s):
        h = self._ctx()
        if s.key != self._structure_key:
            lib = _lib.load()
            nk = np.array(s.node_kinds, dtype=np.int32)
            wk = np.array(s.weight_kinds, dtype=np.int32)
            mn = np.array(s.mean_nterms, dtype=np.int32)
            mk = np.array(s.mean_kinds if s.mean_kinds else [0], dtype=np.int32)
            _lib.check(lib.agpn_set_structure(h, _lib.ptr(nk), _lib.ptr(wk), _lib.ptr(mn), _lib.ptr(mk),
                                              1 if s.use_means else 0))
            assert lib.agpn_theta_dim(h) == s.D
            self._structure_key = s.key
        self._structure = s

    # ---- reference helper API -----------------------------------------------------------------
    def _kernel_pars(self, kernel):
        """arttwo.py:73-75"""
        return kernel.pars

    def _grid(self, time):
        time = np.asarray(time, dtype=np.float64)
        return time if time.any() else np.asarray(self.time, dtype=np.float64)     # arttwo.py:56

    def _kernel_matrix(self, kernel, time=None):
        """Covariance matrix of one kernel on a square grid (arttwo.py:53-62)."""
        t = self._grid(self.time if time is None else time)
        if kernel._nonstationary:
            return kernel(None, t[:, None], t[None, :])
        return kernel(t[:, None] - t[None, :])

    def _predict_kernel_matrix(self, kernel, tstar):
        """arttwo.py:64-71"""
        tstar = np.asarray(tstar, dtype=np.float64)
        t = np.asarray(self.time, dtype=np.float64)
        if kernel._nonstationary:
            return kernel(None, tstar[:, None], t[None, :])
        return kernel(tstar[:, None] - t[None, :])

    # ---- mean functions (arttwo.py:78-126) ----------------------------------------------------
    @property
    def mean_pars_size(self):
        self._mean_pars_size = 0
        for m in self.means:
            self._mean_pars_size += 0 if m is None else m._parsize
        return self._mean_pars_size

    @property
    def mean_pars(self):
        return self._mean_pars

    @mean_pars.setter
    def mean_pars(self, pars):
        pars = list(pars)
        assert len(pars) == self.mean_pars_size
        self._mean_pars = copy(pars)
        for m in self.means:
            if m is None:
                continue
            for j in range(m._parsize):
                m.pars[j] = pars.pop(0)

    def _mean(self, means, time=None):
        """Mean values of every series, concatenated (arttwo.py:106-126).  The mean classes and parameters
        travel with the call (agpn_mean_eval): the network's structure and cached graphs are not touched."""
        ms = list(means)
        p = self.p
        nterms, kinds, pars = [], [], []
        for i in range(p):
            terms = _mean_terms(ms[i]) if i < len(ms) else []
            nterms.append(len(terms))
            for k, pr in terms:
                kinds.append(k)
                pars += pr
        t = np.ascontiguousarray(self.time if time is None else time, dtype=np.float64)
        out = np.empty((p, t.size))
        nt = np.array(nterms, dtype=np.int32)
        mk = np.array(kinds if kinds else [0], dtype=np.int32)
        pv = np.array(pars if pars else [0.0], dtype=np.float64)
        _lib.check(_lib.load().agpn_mean_eval(self._device, p, _lib.ptr(nt), _lib.ptr(mk), _lib.ptr(pv), _lib.ptr(t),
                                              t.size, _lib.ptr(out)))
        return out.reshape(-1)

    def sample(self, kernel, time, size=None, random_state=None):
        """
        Draw from N(0, kernel) (arttwo.py:129-147; `kernel` is a covariance MATRIX on `time`, e.g. the output of
        _kernel_matrix / _covariance_matrix).  The reference calls scipy's multivariate_normal(..., allow_singular=True)
        .rvs(); here the matrix is factorised on the GPU (batched Cholesky kernels, with a minimal nugget for a
        semi-definite matrix) and the sample is L z.  The standard-normal z come from numpy on the host: the global
        numpy state (like scipy's default), or `random_state` (seed / Generator / RandomState).  size=None -> [n].
        """
        K = np.ascontiguousarray(kernel, dtype=np.float64)
        n = np.asarray(time).size
        if K.shape != (n, n):
            raise ValueError("kernel must be the %d x %d covariance matrix of `time`" % (n, n))
        ns = 1 if size is None else int(size)
        if random_state is None:
            z = np.random.standard_normal((ns, n))
        elif hasattr(random_state, "standard_normal"):
            z = random_state.standard_normal((ns, n))
        else:
            z = np.random.default_rng(random_state).standard_normal((ns, n))
        z = np.ascontiguousarray(z, dtype=np.float64)
        out = np.empty((ns, n))
        nug = np.zeros(1)
        _lib.check(_lib.load().agpn_mvn_sample(self._device, n, _lib.ptr(K), ns, _lib.ptr(z), _lib.ptr(out), _lib.ptr(nug)))
        self._last_sample_nugget = float(nug[0])
        return out[0] if size is None else out

    def marginalLikelihood_2ys(self, *a, **k):
        raise NotImplementedError("marginalLikelihood_2ys is outside the likelihood / prediction path")

    # ---- covariance blocks -------------------------------------------------------------------
    def _covariance_matrix(self, nodes, weights, time, position_p, add_errors=False):
        """Block K_ii = sum_j W_ij o F_j on `time` (arttwo.py:151-192); position_p is 1-based."""
        s = Structure(self.p, self.q, nodes, weights, None)
        self._apply_structure(s)
        theta = s.pack(nodes, weights, None, [0.0] * self.p)
        t = np.ascontiguousarray(self._grid(time), dtype=np.float64)
        out = np.empty((t.size, t.size))
        _lib.check(_lib.load().agpn_covariance_block(self._ctx(), _lib.ptr(theta), position_p - 1, _lib.ptr(t), t.size,
                                                     _lib.ptr(t), t.size, 1, 1 if add_errors else 0, _lib.ptr(out), 0, None))
        return out

    def compute_matrix(self, nodes, weights, time, nugget=False, shift=False):
        """Block-diagonal pN x pN matrix (arttwo.py:194-238): sigma (not sigma^2) on the diagonal,
        no jitter; nugget / shift as in the reference.  Blocks come from the assembly kernel."""
        N = np.asarray(self.time).size
        K = np.zeros((N * self.p, N * self.p))
        for i in range(1, self.p + 1):
            K[(i - 1) * N:i * N, (i - 1) * N:i * N] = self._covariance_matrix(nodes, weights, self.time, i, False)
        K[np.diag_indices(N * self.p)] += np.concatenate(self.yerr)
        if nugget:
            K = (1 - 0.01) * K + 0.01 * np.diag(np.diag(K))
        if shift:
            K = K + 0.01 * np.identity(N * self.p)
        return K

    # ---- likelihood ---------------------------------------------------------------------------
    def log_likelihood(self, nodes, weights, means, jitters):
        """
        Marginal log likelihood of the network (arttwo.py:240-290).

        Returns -inf if a block is not positive definite; raises ValueError if the covariance or
        the residual contains NaN / inf (scipy's check_finite in the reference).
        """
        s = Structure(self.p, self.q, nodes, weights, means)
        self._apply_structure(s)
        theta = s.pack(nodes, weights, means, jitters).reshape(1, -1)
        out = np.empty(1)
        status = np.zeros(1, dtype=np.int32)
        _lib.check(_lib.load().agpn_loglike_batch(self._ctx(), 1, _lib.ptr(theta), 0, _lib.ptr(out), 0,
                                                  _lib.ptr(status), None))
        if status[0] == STATUS_NOT_FINITE:
            raise ValueError("array must not contain infs or NaNs")
        return float(out[0])

    def log_likelihood_grad(self, nodes, weights, means, jitters):
        """
        Log likelihood AND its analytic gradient with respect to every hyper-parameter (not in the reference, which
        only carries derivative kernel classes: weight.py:81-931).  Returns (ll, grad) with grad in theta-row order:
        node .pars | weight .pars (list order j + q*i) | mean terms | jitters -- ``Structure.pack`` of the same
        objects gives the matching parameter vector.  Supported classes: see agpn_loglike_grad.
        """
        s = Structure(self.p, self.q, nodes, weights, means)
        self._apply_structure(s)
        theta = s.pack(nodes, weights, means, jitters)
        ll = np.zeros(1)
        grad = np.zeros(s.D)
        status = np.zeros(1, dtype=np.int32)
        _lib.check(_lib.load().agpn_loglike_grad(self._ctx(), _lib.ptr(theta), _lib.ptr(ll), _lib.ptr(grad),
                                                 _lib.ptr(status), None))
        if status[0] == STATUS_NOT_FINITE:
            raise ValueError("array must not contain infs or NaNs")
        return float(ll[0]), grad

    def log_likelihood_batch(self, theta, return_status=False):
        """
        Log likelihoods of W parameter rows (walkers) in one device pass.

        theta: [W, D] float64 -- numpy array (host; copied inside the call) or CUDA torch tensor on
        this network's device (used in place, result returned as a CUDA tensor on the current
        torch stream).  Structure must have been fixed with ``set_structure``.  Rows whose
        covariance is not positive definite give -inf; rows with NaN / inf give NaN.
        """
        if self._structure is None:
            raise RuntimeError("call set_structure(nodes, weights, means) first")
        D = self._structure.D
        lib = _lib.load()
        if type(theta).__module__.split(".")[0] == "torch":
            import torch
            if not theta.is_cuda or theta.dtype != torch.float64 or theta.dim() != 2 or theta.shape[1] != D:
                raise ValueError("theta tensor must be CUDA float64 of shape [W, %d]" % D)
            if theta.device.index != self._device:
                raise ValueError("theta is on cuda:%s, network is on cuda:%d" % (theta.device.index, self._device))
            theta = theta.contiguous()
            W = theta.shape[0]
            out = torch.empty(W, dtype=torch.float64, device=theta.device)
            stream = torch.cuda.current_stream(theta.device).cuda_stream
            if return_status:
                status = np.zeros(W, dtype=np.int32)
                _lib.check(lib.agpn_loglike_batch(self._ctx(), W, theta.data_ptr(), 1, out.data_ptr(), 1,
                                                  _lib.ptr(status), stream))
                return out, status
            _lib.check(lib.agpn_loglike_batch(self._ctx(), W, theta.data_ptr(), 1, out.data_ptr(), 1, None, stream))
            return out
        th = np.ascontiguousarray(theta, dtype=np.float64)
        if th.ndim != 2 or th.shape[1] != D:
            raise ValueError("theta must have shape [W, %d]" % D)
        W = th.shape[0]
        out = np.empty(W)
        status = np.zeros(W, dtype=np.int32)
        _lib.check(lib.agpn_loglike_batch(self._ctx(), W, _lib.ptr(th), 0, _lib.ptr(out), 0, _lib.ptr(status), None))
        return (out, status) if return_status else out

    # ---- prediction ---------------------------------------------------------------------------
    def prediction(self, nodes=None, weights=None, means=None, jitters=None, time=None, dataset=1):
        """
        Conditional predictive distribution (arttwo.py:329-396).

        Returns (y_mean, y_std, time) exactly like the reference (the third value is the time
        grid).  Raises LinAlgError if the covariance is not positive definite (the reference lets
        scipy's error escape here, arttwo.py:371).
        """
        print('Working with dataset {0}'.format(dataset))
        time = np.asarray(time, dtype=np.float64)
        time = time if time.any() else np.asarray(self.time, dtype=np.float64)      # arttwo.py:362
        s = Structure(self.p, self.q, nodes, weights, means)
        self._apply_structure(s)
        theta = s.pack(nodes, weights, means, jitters)
        ts = np.ascontiguousarray(time, dtype=np.float64)
        mu = np.empty(ts.size)
        sd = np.empty(ts.size)
        status = np.zeros(1, dtype=np.int32)
        _lib.check(_lib.load().agpn_predict(self._ctx(), _lib.ptr(theta), dataset - 1, ts.size, _lib.ptr(ts),
                                            _lib.ptr(mu), _lib.ptr(sd), _lib.ptr(status), None))
        if status[0] == STATUS_NOT_PD:
            raise np.linalg.LinAlgError("matrix is not positive definite")
        if status[0] == STATUS_NOT_FINITE:
            raise ValueError("array must not contain infs or NaNs")
        return mu, sd, time

    def prediction_slice(self, nodes, weights, means, jitters, time, lo, hi, dataset=1):
        """Mean / std of the grid points time[lo:hi] of a larger prediction grid `time` (building block of
        sharding.prediction_sharded; see agpn_predict_slice for the two cases it does not cover)."""
        time = np.ascontiguousarray(time, dtype=np.float64)
        s = Structure(self.p, self.q, nodes, weights, means)
        self._apply_structure(s)
        theta = s.pack(nodes, weights, means, jitters)
        ts = np.ascontiguousarray(time[lo:hi])
        mu, sd = np.empty(ts.size), np.empty(ts.size)
        status = np.zeros(1, dtype=np.int32)
        if ts.size:
            _lib.check(_lib.load().agpn_predict_slice(self._ctx(), _lib.ptr(theta), dataset - 1, ts.size, _lib.ptr(ts),
                                                      time.size, int(lo), float(time.mean()), float(time[0]),
                                                      _lib.ptr(mu), _lib.ptr(sd), _lib.ptr(status), None))
        if status[0] == STATUS_NOT_PD:
            raise np.linalg.LinAlgError("matrix is not positive definite")
        if status[0] == STATUS_NOT_FINITE:
            raise ValueError("array must not contain infs or NaNs")
        return mu, sd

    def prediction_cov(self, nodes=None, weights=None, means=None, jitters=None, time=None, dataset=1,
                       train_jitter=None):
        """
        Predictive mean, standard deviation AND the full covariance matrix (the `y_cov` the reference
        builds at arttwo.py:393 but does not return; legacy art.network.prediction returns it, art.py:354).

        Computed on the GPU as the Schur complement of a partial blocked Cholesky of [K . ; K* K**].
        Dense (N + M)^2 workspace -- meant for grids of up to a few 10^4 points.
        train_jitter: jitter used in the training block only (legacy art.network quirk); default: the
        dataset's entry of `jitters`.  Returns (y_mean, y_std, y_cov).
        """
        time = np.asarray(time, dtype=np.float64)
        time = time if time.any() else np.asarray(self.time, dtype=np.float64)
        s = Structure(self.p, self.q, nodes, weights, means)
        self._apply_structure(s)
        theta = s.pack(nodes, weights, means, jitters)
        ts = np.ascontiguousarray(time, dtype=np.float64)
        mu = np.empty(ts.size)
        cov = np.empty((ts.size, ts.size))
        status = np.zeros(1, dtype=np.int32)
        tj = float("nan") if train_jitter is None else float(train_jitter)
        _lib.check(_lib.load().agpn_predict_cov(self._ctx(), _lib.ptr(theta), dataset - 1, ts.size, _lib.ptr(ts), tj,
                                                _lib.ptr(mu), _lib.ptr(cov), _lib.ptr(status), None))
        if status[0] == STATUS_NOT_PD:
            raise np.linalg.LinAlgError("matrix is not positive definite")
        if status[0] == STATUS_NOT_FINITE:
            raise ValueError("array must not contain infs or NaNs")
        return mu, np.sqrt(np.diag(cov)), cov

    def posterior_sample(self, nodes=None, weights=None, means=None, jitters=None, time=None, dataset=1, size=None,
                         random_state=None):
        """
        Draws from the predictive distribution N(y_mean, y_cov) of `prediction_cov` on the grid `time` (what sampling the
        `y_cov` of arttwo.py:393 would give): the Schur-complement covariance and its factor both come from the batched
        Cholesky kernels; only the standard-normal numbers are drawn on the host.  Returns [len(time)] (size=None) or
        [size, len(time)].
        """
        mu, _, cov = self.prediction_cov(nodes, weights, means, jitters, time, dataset)
        t = np.asarray(time, dtype=np.float64)
        t = t if t.any() else np.asarray(self.time, dtype=np.float64)
        return self.sample(cov, t, size=size, random_state=random_state) + mu

    # ---- profiling hooks used by bench.py -----------------------------------------------------
    def set_profiling(self, enable):
        _lib.check(_lib.load().agpn_set_profiling(self._ctx(), 1 if enable else 0))

    def phase_times(self):
        """(ms[6], launches[6]) of the last log_likelihood_batch: mean residual, assembly,
        diagonal-block factor, panel solve, trailing update, finalise."""
        ms = np.zeros(6)
        n = np.zeros(6, dtype=np.int64)
        _lib.check(_lib.load().agpn_phase_ms(self._ctx(), _lib.ptr(ms), _lib.ptr(n)))
        return ms, n

    def update_path(self):
        """0: every trailing update on the FP64 tensor instruction; bit 0 / bit 1: narrow / wide updates on the
        INT8-sliced path (chosen at construction from AGPN_SYRK)."""
        return int(_lib.load().agpn_update_path(self._ctx()))

    def skip_stats(self):
        """(dense K steps, executed K steps, panel-solve tile halves, halves left out, right halves among them) of the last wave of
        log_likelihood_batch: the INT8-sliced update leaves out K steps whose digit chunks are all zero (exact), the
        panel solve the tile halves bounded below 2^-96 of their row scale."""
        out = np.zeros(5)
        _lib.check(_lib.load().agpn_skip_stats(self._ctx(), _lib.ptr(out)))
        return tuple(float(v) for v in out)

    def predict_path(self):
        """1: the last prediction ran on the factorisation kernels (K* appended as block rows, INT8-sliced updates);
        0: on the stand-alone FP64 predict_kernel (AGPN_PREDICT=dmma or automatic fall-back); -1: none yet."""
        return int(_lib.load().agpn_predict_path(self._ctx()))

    def slice_times(self):
        """(ms, launches) of the digit-slicing kernels of the last profiled log_likelihood_batch."""
        ms = np.zeros(1)
        n = np.zeros(1, dtype=np.int64)
        _lib.check(_lib.load().agpn_slice_ms(self._ctx(), _lib.ptr(ms), _lib.ptr(n)))
        return float(ms[0]), int(n[0])


def evaluate_mean(m, t, device=None):
    """Value of one mean-function object on the grid t, computed on the GPU (no context needed)."""
    t = np.ascontiguousarray(t, dtype=np.float64)
    terms = _mean_terms(m)
    nt = np.array([len(terms)], dtype=np.int32)
    mk = np.array([k for k, _ in terms] or [0], dtype=np.int32)
    pv = np.array([x for _, pr in terms for x in pr] or [0.0], dtype=np.float64)
    out = np.empty(t.size)
    _lib.check(_lib.load().agpn_mean_eval(get_device() if device is None else int(device), 1, _lib.ptr(nt), _lib.ptr(mk),
                                          _lib.ptr(pv), _lib.ptr(t), t.size, _lib.ptr(out)))
    return out
