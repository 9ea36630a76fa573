"""Node kernels (unit amplitude) -- drop-in for ``artgpn.node`` (node.py:36-327).

Same class names and positional constructor parameters; ``.pars`` is a float64 array in
constructor order.  Calling an instance evaluates it on the GPU.

    Constant(c)  WhiteNoise(wn)  SquaredExponential(ell)  Periodic(P, ell)
    QuasiPeriodic(ell_e, P, ell_p)  RationalQuadratic(alpha, ell)  RQP(alpha, ell_e, P, ell_p)
    Cosine(P)  Exponential(ell)  Matern32(ell)  Matern52(ell)  Linear(c)  GammaExp(gamma, ell)
    Polynomial(a, b, c)
"""
from ._kernelbase import KernelFunction, make_kernel_class
from .kinds import KERNELS, FAMILY_NODE


class nodeFunction(KernelFunction):
    """Base class of the node kernels (node.py:7-32)."""
    _family = FAMILY_NODE


for _name in KERNELS:
    globals()[_name] = make_kernel_class(_name, FAMILY_NODE, nodeFunction, module=__name__)
del _name

__all__ = ["nodeFunction"] + list(KERNELS)
