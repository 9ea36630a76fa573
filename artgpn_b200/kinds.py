"""Kernel / mean class tables shared by ``node.py``, ``weight.py``, ``mean.py`` and the
theta packing.  Ids and parameter orders are those of ``include/artgpn_b200.h``; parameter
names follow the reference constructors (node.py, weight.py, mean.py)."""

# name -> (kind id, constructor parameter names of the NODE version)
KERNELS = {
    "Constant": (0, ("c",)),
    "WhiteNoise": (1, ("wn",)),
    "SquaredExponential": (2, ("ell",)),
    "Periodic": (3, ("P", "ell")),
    "QuasiPeriodic": (4, ("ell_e", "P", "ell_p")),
    "RationalQuadratic": (5, ("alpha", "ell")),
    "RQP": (6, ("alpha", "ell_e", "P", "ell_p")),
    "Cosine": (7, ("P",)),
    "Exponential": (8, ("ell",)),
    "Matern32": (9, ("ell",)),
    "Matern52": (10, ("ell",)),
    "Linear": (11, ("c",)),
    "GammaExp": (12, ("gamma", "ell")),
    "Polynomial": (13, ("a", "b", "c")),
}
# weight-only alias: weight.Laplacian has Exponential's formula (weight.py:573-574, :618-619)
WEIGHT_ALIASES = {"Laplacian": "Exponential"}
NONSTATIONARY = ("Linear", "Polynomial")       # arttwo.py:58,66
NO_AMPLITUDE = ("Constant", "WhiteNoise")      # weight versions keep a single parameter

# name -> (kind id, parameter names, _parsize as DECLARED by the reference (mean.py:129,144 say 3
# for the 4-parameter Sine / Cosine; network.mean_pars honours the declared value))
MEANS = {
    "Constant": (0, ("c",), 1),
    "Linear": (1, ("slope", "intercept"), 2),
    "Parabola": (2, ("quad", "slope", "intercept"), 3),
    "Cubic": (3, ("cub", "quad", "slope", "intercept"), 4),
    "Sine": (4, ("amp", "P", "phi", "D"), 3),
    "Cosine": (5, ("amp", "P", "phi", "D"), 3),
    "Keplerian": (6, ("P", "K", "e", "w", "phi"), 5),
}
FAMILY_NODE, FAMILY_WEIGHT = 0, 1
