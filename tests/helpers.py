"""Shared test helpers: golden fixtures (generated from the reference by
tests/golden/make_golden.py), object builders and the CPU oracle."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import artgpn_oracle as oracle          # noqa: E402  (test infrastructure only)

GOLD = os.path.join(HERE, "golden")
_cache = {}


def cases():
    if "cases" not in _cache:
        with open(os.path.join(GOLD, "golden_cases.json")) as f:
            _cache["cases"] = json.load(f)
    return _cache["cases"]


def arrays():
    if "arrays" not in _cache:
        _cache["arrays"] = dict(np.load(os.path.join(GOLD, "golden_arrays.npz")))
    return _cache["arrays"]


def dataset(name):
    if "data" not in _cache:
        _cache["data"] = dict(np.load(os.path.join(GOLD, "golden_data.npz")))
    d = _cache["data"]
    return {"time": d[name + "_time"], "y": d[name + "_y"], "yerr": d[name + "_yerr"]}


def loglike_case(name):
    for c in cases()["loglike"]:
        if c["name"] == name:
            return c
    raise KeyError(name)


def build(mod, spec):
    """object of module `mod` from ["Name", [pars]] (means: nested Sum / None)."""
    if spec is None:
        return None
    name, pars = spec
    if name == "Sum":
        return build(mod, pars[0]) + build(mod, pars[1])
    return getattr(mod, name)(*pars)


def spec_tuple(spec):
    """golden JSON spec -> oracle (name, pars) tuple (Sum nested)."""
    if spec is None:
        return None
    name, pars = spec
    if name == "Sum":
        return ("Sum", [spec_tuple(pars[0]), spec_tuple(pars[1])])
    return (name, list(pars))


class _SumSpec(object):
    pass


def oracle_objs(case):
    """(nodes, weights, means) for the oracle from a golden case."""
    class S(object):            # tiny carrier so oracle.spec() sees a Sum
        def __init__(self, m1, m2):
            self.m1, self.m2 = m1, m2
    S.__name__ = "Sum"

    def mk(spec):
        if spec is None:
            return None
        name, pars = spec
        if name == "Sum":
            return S(mk(pars[0]), mk(pars[1]))
        return (name, list(pars))
    nodes = [mk(s) for s in case["nodes"]]
    weights = [mk(s) for s in case["weights"]]
    means = None if case["means"] is None else [mk(s) for s in case["means"]]
    return nodes, weights, means


def oracle_loglike(case):
    d = dataset(case["data"])
    nodes, weights, means = oracle_objs(case)
    return oracle.log_likelihood(d["time"], d["y"], d["yerr"], case["q"], nodes, weights, means, case["jitters"])


def product_network(case_or_data, q=None):
    """artgpn_b200 network for a golden case (or a dataset name + q)."""
    from artgpn_b200.arttwo import network
    if isinstance(case_or_data, dict):
        d, q = dataset(case_or_data["data"]), case_or_data["q"]
    else:
        d = dataset(case_or_data)
    args = []
    for i in range(d["y"].shape[0]):
        args += [d["y"][i].copy(), d["yerr"][i].copy()]
    return network(q, d["time"].copy(), *args)


def product_objs(case):
    from artgpn_b200 import node, weight, mean
    nodes = [build(node, s) for s in case["nodes"]]
    weights = [build(weight, s) for s in case["weights"]]
    means = None if case["means"] is None else [build(mean, s) for s in case["means"]]
    return nodes, weights, means


def rel_err(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
