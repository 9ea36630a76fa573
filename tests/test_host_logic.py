"""CPU: host-side mirror of the reference API (class names, constructor orders, .pars, theta
packing), the C-ABI library (loads, exports every symbol the header declares) and the absence of
any CPU fallback."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest

import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# constructor parameter orders of the reference (node.py / weight.py / mean.py), written out by hand
NODE_SIGS = {"Constant": ["c"], "WhiteNoise": ["wn"], "SquaredExponential": ["ell"], "Periodic": ["P", "ell"],
             "QuasiPeriodic": ["ell_e", "P", "ell_p"], "RationalQuadratic": ["alpha", "ell"],
             "RQP": ["alpha", "ell_e", "P", "ell_p"], "Cosine": ["P"], "Exponential": ["ell"], "Matern32": ["ell"],
             "Matern52": ["ell"], "Linear": ["c"], "GammaExp": ["gamma", "ell"], "Polynomial": ["a", "b", "c"]}
WEIGHT_SIGS = {"Constant": ["c"], "WhiteNoise": ["wn"], "SquaredExponential": ["weight", "ell"],
               "Periodic": ["weight", "P", "ell"], "QuasiPeriodic": ["weight", "ell_e", "P", "ell_p"],
               "RationalQuadratic": ["weight", "alpha", "ell"], "RQP": ["weight", "alpha", "ell_e", "P", "ell_p"],
               "Cosine": ["weight", "P"], "Laplacian": ["weight", "ell"], "Exponential": ["weight", "ell"],
               "Matern32": ["weight", "ell"], "Matern52": ["weight", "ell"], "Linear": ["weight", "c"],
               "GammaExp": ["weight", "gamma", "ell"], "Polynomial": ["weight", "a", "b", "c"]}
MEAN_SIGS = {"Constant": ["c"], "Linear": ["slope", "intercept"], "Parabola": ["quad", "slope", "intercept"],
             "Cubic": ["cub", "quad", "slope", "intercept"], "Sine": ["amp", "P", "phi", "D"],
             "Cosine": ["amp", "P", "phi", "D"], "Keplerian": ["P", "K", "e", "w", "phi"]}


def test_kernel_and_mean_class_surface():
    from artgpn_b200 import node, weight, mean
    for mod, sigs in ((node, NODE_SIGS), (weight, WEIGHT_SIGS), (mean, MEAN_SIGS)):
        for name, params in sigs.items():
            cls = getattr(mod, name)
            assert list(inspect.signature(cls).parameters) == params, (mod.__name__, name)
            obj = cls(*[float(i + 1) for i in range(len(params))])
            assert list(obj.pars) == [float(i + 1) for i in range(len(params))]
            assert repr(obj).startswith(name + "(")
            with pytest.raises(TypeError):
                cls(*([1.0] * (len(params) + 1)))
    assert len(NODE_SIGS) == 14 and len(WEIGHT_SIGS) == 15
    k = node.QuasiPeriodic(10, 20, 0.5)
    assert (k.ell_e, k.P, k.ell_p) == (10, 20, 0.5) and k.pars.dtype == np.float64
    assert isinstance(mean.Linear(1, 2).pars, list)                # mean.py:23
    s = mean.Linear(1, 2) + mean.Keplerian(3, 4, 5, 6, 7)
    assert s.pars == [1, 2, 3, 4, 5, 6, 7] and s._parsize == 7 and repr(s) == "Linear(1, 2) + Keplerian(3, 4, 5, 6, 7)"
    assert mean.Sine._parsize == 3 and mean.Cosine._parsize == 3     # declared value of the reference (mean.py:129,144)


def test_library_exports_every_declared_symbol():
    from artgpn_b200 import _lib
    header = open(os.path.join(ROOT, "include", "artgpn_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = sorted(set(re.findall(r"\b(agpn_[a-z0-9_]+)\s*\(", header)))
    assert declared == _lib.symbols(), (declared, _lib.symbols())
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    lib = _lib.load()
    assert lib.agpn_version() >= 100
    # parameter-count tables agree with the class surface (no CUDA needed for these calls)
    from artgpn_b200 import node, weight, mean
    for name, params in NODE_SIGS.items():
        assert lib.agpn_kernel_npars(0, getattr(node, name)._kind) == len(params)
    for name, params in WEIGHT_SIGS.items():
        assert lib.agpn_kernel_npars(1, getattr(weight, name)._kind) == len(params)
    for name, params in MEAN_SIGS.items():
        assert lib.agpn_mean_npars(getattr(mean, name)._kind) == len(params)
    assert lib.agpn_kernel_npars(0, 99) == -1 and lib.agpn_mean_npars(-1) == -1


def test_theta_packing_layout():
    from artgpn_b200 import node, weight, mean
    from artgpn_b200.arttwo import Structure
    nodes = [node.QuasiPeriodic(1, 2, 3), node.Periodic(4, 5)]
    weights = [weight.Constant(10 + i) if i % 2 == 0 else weight.SquaredExponential(20 + i, 30 + i) for i in range(6)]
    means = [mean.Linear(0.1, 0.2), None, mean.Constant(0.3) + mean.Sine(1, 2, 3, 4)]
    s = Structure(3, 2, nodes, weights, means)
    row = s.pack(nodes, weights, means, [7, 8, 9])
    assert list(row) == [1, 2, 3, 4, 5, 10, 21, 31, 12, 23, 33, 14, 25, 35, 0.1, 0.2, 0.3, 1, 2, 3, 4, 7, 8, 9]
    assert s.D == row.size and s.mean_nterms == [1, 0, 2] and s.mean_kinds == [1, 0, 4]
    s0 = Structure(3, 2, nodes, weights, None)                         # `if means` falsy (arttwo.py:263)
    assert not s0.use_means and s0.D == 14 + 3
    with pytest.raises(IndexError):
        Structure(3, 2, nodes, weights[:5], means)                     # too few weights (arttwo.py:274)
    with pytest.raises(TypeError):
        Structure(3, 2, [weight.Constant(1)] * 2, weights, means)      # weight object in the node list


def test_reference_objects_are_accepted_by_name():
    """Objects of foreign classes with the reference's class names and .pars (e.g. the reference's own
    kernel objects) map onto the same ids -- the reference itself only uses type(k) and k.pars."""
    from artgpn_b200.arttwo import Structure

    def fake(name, pars):
        return type(name, (), {})() if False else type(name, (), {"pars": pars})()
    nodes = [fake("QuasiPeriodic", [1.0, 2.0, 3.0])]
    weights = [fake("Laplacian", [1.0, 2.0]), fake("Constant", [3.0])]
    s = Structure(2, 1, nodes, weights, [fake("Linear", [0.5, 0.25]), None])
    assert s.node_kinds == [4] and s.weight_kinds == [8, 0] and s.mean_kinds == [1]


def test_network_attributes_match_reference_layout():
    from artgpn_b200.arttwo import network
    d = H.dataset("sun50")
    net = network(1, d["time"], d["y"][0], d["yerr"][0], d["y"][1], d["yerr"][1], d["y"][2], d["yerr"][2])
    assert (net.q, net.p, net.qp) == (1, 3, 3)
    assert net.y.shape == (3, 50) and net.yerr.shape == (3, 50) and net.tt.shape == (150,)
    assert np.array_equal(net.yerr, d["yerr"])              # raw sigma, not squared (arttwo.py:40-48)


def test_no_cpu_fallback():
    """Without a CUDA device the product raises; it never computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from artgpn_b200 import AgpnError
    case = H.loglike_case("G4_sun50_initial")
    net = H.product_network(case)
    nodes, weights, means = H.product_objs(case)
    with pytest.raises(AgpnError):
        net.log_likelihood(nodes, weights, means, case["jitters"])
    with pytest.raises(AgpnError):
        nodes[0](np.zeros((3, 3)))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "artgpn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "artgpn_oracle" not in src and "import oracle" not in src, f
