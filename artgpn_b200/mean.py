"""Mean functions -- drop-in for ``artgpn.mean`` (mean.py:20-202).

Objects carry ``.pars`` (a plain list, as in the reference) and a class id; the values are
computed on the GPU inside ``network`` (``agpn_mean_table`` / the residual kernel).  Calling an
instance evaluates it on the GPU for the grid passed (Linear is centred on the mean of that grid,
Keplerian is anchored at its first point -- mean.py:92,175).

    Constant(c)  Linear(slope, intercept)  Parabola(quad, slope, intercept)
    Cubic(cub, quad, slope, intercept)  Sine(amp, P, phi, D)  Cosine(amp, P, phi, D)
    Keplerian(P, K, e, w, phi)  and  m1 + m2 -> Sum
"""
import inspect

import numpy as np

from .kinds import MEANS

__all__ = ["Constant", "Linear", "Parabola", "Cubic", "Keplerian"]     # as the reference (mean.py:6)


class MeanModel(object):
    _parsize = 0
    _kind = -1
    _names = ()

    def __init__(self, *pars):
        if self._names and len(pars) != len(self._names):
            raise TypeError("%s() takes %d positional arguments (%s) but %d were given" % (
                type(self).__name__, len(self._names), ", ".join(self._names), len(pars)))
        self.pars = list(pars)

    def __repr__(self):
        return "{0}({1})".format(self.__class__.__name__, ", ".join(map(str, self.pars)))

    @classmethod
    def initialize(cls):
        """All parameters 0 (mean.py:31-34; uses the declared ``_parsize``)."""
        return cls(*([0.] * len(cls._names)))

    def __add__(self, b):
        return Sum(self, b)

    def __radd__(self, b):
        return self.__add__(b)

    def terms(self):
        """Flat list of (kind id, parameter list) -- a Sum contributes several."""
        return [(self._kind, [float(x) for x in self.pars])]

    def __call__(self, t):
        from .arttwo import evaluate_mean
        return evaluate_mean(self, np.atleast_1d(np.asarray(t, dtype=np.float64)))


class Sum(MeanModel):
    """Sum of two mean functions (mean.py:42-65)."""

    def __init__(self, m1, m2):
        self.m1, self.m2 = m1, m2

    @property
    def _parsize(self):
        return self.m1._parsize + self.m2._parsize

    @property
    def pars(self):
        return self.m1.pars + self.m2.pars

    def initialize(self):
        return

    def __repr__(self):
        return "{0} + {1}".format(self.m1, self.m2)

    def terms(self):
        return self.m1.terms() + self.m2.terms()


def _make(name):
    kind, names, parsize = MEANS[name]
    cls = type(name, (MeanModel,), {"_kind": kind, "_names": names, "_parsize": parsize,
                                    "__doc__": "%s mean function, parameters (%s)." % (name, ", ".join(names))})
    cls.__signature__ = inspect.Signature(
        [inspect.Parameter(n, inspect.Parameter.POSITIONAL_OR_KEYWORD) for n in names])
    cls.__module__ = __name__
    return cls


for _name in MEANS:
    globals()[_name] = _make(_name)
del _name
