"""Weight kernels (amplitude ``weight**2`` times the node shapes) -- drop-in for the value
classes of ``artgpn.weight`` (weight.py:65-872).  ``weight`` comes first in ``.pars`` except for
Constant(c) and WhiteNoise(wn).  The reference's ``d*_d*`` derivative classes are never used by
its likelihood or prediction code and are out of scope.

    Constant(c)  WhiteNoise(wn)  SquaredExponential(weight, ell)  Periodic(weight, P, ell)
    QuasiPeriodic(weight, ell_e, P, ell_p)  RationalQuadratic(weight, alpha, ell)
    RQP(weight, alpha, ell_e, P, ell_p)  Cosine(weight, P)  Laplacian(weight, ell)
    Exponential(weight, ell)  Matern32(weight, ell)  Matern52(weight, ell)  Linear(weight, c)
    GammaExp(weight, gamma, ell)  Polynomial(weight, a, b, c)
"""
from ._kernelbase import KernelFunction, make_kernel_class
from .kinds import KERNELS, WEIGHT_ALIASES, FAMILY_WEIGHT


class weightFunction(KernelFunction):
    """Base class of the weight kernels (weight.py:7-34)."""
    _family = FAMILY_WEIGHT


for _name in KERNELS:
    globals()[_name] = make_kernel_class(_name, FAMILY_WEIGHT, weightFunction, module=__name__)
for _name, _target in WEIGHT_ALIASES.items():
    globals()[_name] = make_kernel_class(_name, FAMILY_WEIGHT, weightFunction, alias_of=_target, module=__name__)
del _name, _target

__all__ = ["weightFunction"] + list(KERNELS) + list(WEIGHT_ALIASES)
