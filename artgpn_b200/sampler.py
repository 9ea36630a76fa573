"""Sampler-side adapter: the caller that actually produces walker batches.

The reference evaluates one walker per process (`log_transform` + `multiprocessing.Pool`,
examples/Sun/01_emcee.py:76-122; `utils.run_mcmc`, utils.py:123-162).  Here the whole set of
proposals of an ensemble half-step is ONE batched device call:

    post = BatchedPosterior(GPnet, nodes, weights, means, log_prior=my_vectorised_prior)
    sampler = emcee.EnsembleSampler(nwalkers, ndim, post, vectorize=True)        # if emcee is installed
    chain, logp, acc = run_mcmc(prior_func, post, iterations=1000)              # built-in stretch move

`run_mcmc` keeps the reference's argument meaning (prior_func draws one starting point per call,
`2*ndim` walkers, half of the iterations are burn-in) and uses a small built-in Goodman & Weare
stretch-move ensemble sampler (the algorithm emcee implements) because emcee / dynesty are not
part of this package.  The sampler itself is host-side numpy bookkeeping on [W, ndim] arrays; every
log-likelihood comes from the GPU.
"""
import numpy as np


class BatchedPosterior(object):
    """log-prior (host, vectorised) + log-likelihood (GPU) of many parameter rows at once.

    nodes / weights / means: template objects that fix the kernel / mean classes.
    columns: for each of the D entries of a network theta row (nodes | weights j + q*i | mean terms |
        jitters, see include/artgpn_b200.h) the index of the sampler parameter it comes from;
        default: identity (the Sun example's ordering, examples/Sun/01_emcee.py:77).
    fixed: {row position: value} entries of the network row that are not sampled.
    log_prior: callable([W, ndim]) -> [W]; rows with a -inf prior are not evaluated (the early
        `return -np.inf` of the reference's log_transform).
    Rows whose covariance is not positive definite give -inf (arttwo.py:288-289); rows with NaN / inf
    give -inf too unless on_nonfinite='raise' (the reference would raise ValueError inside the sampler).
    """

    def __init__(self, net, nodes, weights, means=None, columns=None, fixed=None, log_prior=None,
                 on_nonfinite="-inf"):
        self.net = net
        self.structure = net.set_structure(nodes, weights, means)
        D = self.structure.D
        self.fixed = dict(fixed or {})
        if columns is None:
            free = [i for i in range(D) if i not in self.fixed]
            columns = {pos: k for k, pos in enumerate(free)}
        elif not isinstance(columns, dict):
            columns = {pos: int(c) for pos, c in enumerate(columns) if c is not None and pos not in self.fixed}
        missing = [i for i in range(D) if i not in columns and i not in self.fixed]
        if missing:
            raise ValueError("network row positions %s are neither sampled nor fixed" % missing)
        self.positions = np.array(sorted(columns), dtype=np.int64)
        self.sources = np.array([columns[p] for p in self.positions], dtype=np.int64)
        self.ndim = int(self.sources.max()) + 1 if self.sources.size else 0
        self.log_prior = log_prior
        self.on_nonfinite = on_nonfinite
        self.n_evaluated = 0

    def rows(self, theta):
        """[W, ndim] sampler parameters -> [W, D] network rows."""
        theta = np.atleast_2d(np.asarray(theta, dtype=np.float64))
        out = np.empty((theta.shape[0], self.structure.D))
        out[:, self.positions] = theta[:, self.sources]
        for pos, v in self.fixed.items():
            out[:, pos] = v
        return out

    def __call__(self, theta):
        single = np.ndim(theta) == 1
        theta = np.atleast_2d(np.asarray(theta, dtype=np.float64))
        W = theta.shape[0]
        lp = np.zeros(W) if self.log_prior is None else np.asarray(self.log_prior(theta), dtype=np.float64).reshape(W)
        out = np.full(W, -np.inf)
        ok = np.isfinite(lp)
        if ok.any():
            self.net._apply_structure(self.structure)
            ll, status = self.net.log_likelihood_batch(self.rows(theta[ok]), return_status=True)
            self.n_evaluated += int(ok.sum())
            if self.on_nonfinite == "raise" and np.any(status == 2):
                raise ValueError("array must not contain infs or NaNs")
            ll = np.where(status == 0, ll, -np.inf)
            out[ok] = lp[ok] + ll
        return float(out[0]) if single else out


def stretch_move(log_prob, p0, nsteps, seed=0, a=2.0, log_prob0=None):
    """Goodman & Weare (2010) affine-invariant stretch move on two half-ensembles (what
    emcee.EnsembleSampler does with vectorize=True): each half-step proposes W/2 points and
    evaluates them with ONE call log_prob([W/2, ndim]).
    Returns chain [nsteps, W, ndim], log-probabilities [nsteps, W], acceptance fraction [W]."""
    rng = np.random.default_rng(seed)
    x = np.array(p0, dtype=np.float64)
    W, ndim = x.shape
    if W % 2 or W < 2 * ndim:
        raise ValueError("need an even number of walkers, at least 2 * ndim")
    lp = np.asarray(log_prob(x), dtype=np.float64) if log_prob0 is None else np.array(log_prob0, dtype=np.float64)
    if not np.all(np.isfinite(lp)):
        raise ValueError("initial walkers must have finite log-probability")
    chain = np.empty((nsteps, W, ndim))
    logp = np.empty((nsteps, W))
    accepted = np.zeros(W)
    half = W // 2
    sets = (np.arange(half), np.arange(half, W))
    for it in range(nsteps):
        for s in (0, 1):
            mine, other = sets[s], sets[1 - s]
            z = ((a - 1.0) * rng.random(half) + 1.0) ** 2 / a
            partner = x[other[rng.integers(0, half, half)]]
            prop = partner + z[:, None] * (x[mine] - partner)
            lp_new = np.asarray(log_prob(prop), dtype=np.float64)
            with np.errstate(invalid="ignore"):
                log_accept = (ndim - 1.0) * np.log(z) + lp_new - lp[mine]
            acc = np.log(rng.random(half)) < log_accept
            x[mine[acc]] = prop[acc]
            lp[mine[acc]] = lp_new[acc]
            accepted[mine[acc]] += 1
        chain[it] = x
        logp[it] = lp
    return chain, logp, accepted / max(nsteps, 1)


def run_mcmc(prior_func, logpost_func, iterations=1000, sampler="emcee", threads=None, seed=0):
    """Counterpart of artgpn.utils.run_mcmc (utils.py:123-162) for a batched posterior.

    prior_func() returns one starting point; 2 * ndim walkers; the first half of `iterations` is
    burn-in; returns the same array layout as the reference's emcee branch: rows = post-burn-in
    samples, columns = parameters followed by the log-probability.  `threads` is ignored (the batch
    replaces the process pool).  sampler = 'dynesty' follows the reference's second branch (utils.py:155-161:
    DynamicNestedSampler, nlive 2500, 'rwalk') when the dynesty package is importable -- prior_func is then the prior
    transform of the unit cube and logpost_func is called with one point at a time ([1, ndim] batches), without the
    reference's process pool (a CUDA context must not be forked); without dynesty it raises ImportError.
    """
    if sampler == "dynesty":
        try:
            import dynesty
        except ImportError as e:
            raise ImportError("run_mcmc(sampler='dynesty') needs the dynesty package (not bundled)") from e
        ndim = np.asarray(prior_func(0)).size
        one = lambda th: float(np.asarray(logpost_func(np.asarray(th, dtype=np.float64)[None, :]))[0])      # noqa: E731
        dsampler = dynesty.DynamicNestedSampler(one, prior_func, ndim=ndim, nlive=2500, sample="rwalk")
        dsampler.run_nested(nlive_init=2500, maxiter=iterations)
        return dsampler.results
    if sampler != "emcee":
        raise ValueError("sampler must be 'emcee' or 'dynesty'")
    ndim = np.asarray(prior_func()).size
    burns, runs = int(iterations / 2), int(iterations / 2)
    nwalkers = 2 * ndim
    p0 = np.array([prior_func() for _ in range(nwalkers)], dtype=np.float64)
    print("Running burn-in")
    chain, logp, _ = stretch_move(logpost_func, p0, burns, seed=seed)
    print("Running production chain")
    chain, logp, _ = stretch_move(logpost_func, chain[-1], runs, seed=seed + 1, log_prob0=logp[-1])
    samples = chain.transpose(1, 0, 2).reshape(-1, ndim)
    lnprob = logp.T.reshape(-1, 1)
    return np.hstack([samples, lnprob])
