"""Generate the golden vectors under tests/golden/ by EXECUTING THE UNMODIFIED REFERENCE.

Run in the build container only (needs /root/reference; the GPU box has no reference):

    python tests/golden/make_golden.py

The reference package imports emcee / dynesty at module level (artgpn/utils.py:120); neither is
installed and neither is on the likelihood path, so two empty stub modules are registered first.
Outputs:  golden_data.npz (example data + synthetic inputs), golden_cases.json (specs and
expected values), golden_arrays.npz (expected arrays).  Everything the tests compare against
comes from calls into the reference made here (numpy / scipy versions are recorded).
"""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"

sys.dont_write_bytecode = True
for _m in ("emcee", "dynesty"):
    sys.modules.setdefault(_m, types.ModuleType(_m))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

from artgpn.arttwo import network as ref_network          # noqa: E402
from artgpn.art import network as ref_art_network         # noqa: E402
from artgpn import node as ref_node, weight as ref_weight, mean as ref_mean   # noqa: E402
import scipy                                                # noqa: E402
from artgpn_b200 import synth                               # noqa: E402  (pure-numpy generators)


def build(mod, spec):
    """object from ["Name", [pars]] (means: nested Sum, None)."""
    if spec is None:
        return None
    name, pars = spec
    if name == "Sum":
        return build(mod, pars[0]) + build(mod, pars[1])
    return getattr(mod, name)(*pars)


def network_of(data, q):
    args = []
    for i in range(data["y"].shape[0]):
        args += [data["y"][i].copy(), data["yerr"][i].copy()]
    return ref_network(q, data["time"].copy(), *args)


def run_ll(data, case):
    net = network_of(data, case["q"])
    nodes = [build(ref_node, s) for s in case["nodes"]]
    weights = [build(ref_weight, s) for s in case["weights"]]
    means = None if case["means"] is None else [build(ref_mean, s) for s in case["means"]]
    try:
        v = net.log_likelihood(nodes, weights, means, case["jitters"])
    except ValueError:
        return "ValueError"
    if np.isneginf(v):
        return "-inf"
    return float(v)


def main():
    data = {}
    arrays = {}
    cases = {"versions": {"numpy": np.__version__, "scipy": scipy.__version__}, "loglike": [], "predict": [],
             "kernels": [], "means": [], "covariance": [], "batch": [], "legacy_art": []}

    # ---- example data -------------------------------------------------------------------------
    corot = np.loadtxt(os.path.join(REF, "examples/CoRoT-7/corot7_data.rdb"), skiprows=1)

    def corot_set(nrows, nseries):
        # examples/CoRoT-7/arttwo_exemple.py:9-32 (err /= val.std() BEFORE val /= val.std())
        d = corot[:nrows].copy()
        t = d[:, 0]
        ys, es = [d[:, 1].copy()], [d[:, 2].copy()]
        for c in (3, 5, 7)[:nseries - 1]:
            v, e = d[:, c].copy(), d[:, c + 1].copy()
            v -= v.mean()
            e /= v.std()
            v /= v.std()
            ys.append(v)
            es.append(e)
        return {"time": t, "y": np.array(ys), "yerr": np.array(es)}

    data["corot20x4"] = corot_set(20, 4)
    data["corot167x4"] = corot_set(167, 4)
    data["corot20x3"] = corot_set(20, 3)
    data["corot167x3"] = corot_set(167, 3)
    sun = np.loadtxt(os.path.join(REF, "examples/Sun/sun50points.txt"), skiprows=1, usecols=(0, 1, 2, 7, 8, 9, 10))
    data["sun50"] = {"time": sun[:, 0], "y": sun[:, [1, 3, 5]].T.copy(), "yerr": sun[:, [2, 4, 6]].T.copy()}
    for N in (64, 257, 512):
        t, y, s = synth.make_data(N, 3, seed=0)
        data["synth%d" % N] = {"time": t, "y": y, "yerr": s}

    def example_case(name, dname, nodes, weights, mean_kind):
        d = data[dname]
        p = d["y"].shape[0]
        if mean_kind == "Constant":
            means = [["Constant", [float(np.mean(d["y"][i]))]] for i in range(p)]
        else:
            means = [["Linear", [0.0, float(np.mean(d["y"][i]))]] for i in range(p)]
        return {"name": name, "data": dname, "q": len(nodes), "nodes": nodes, "weights": weights, "means": means,
                "jitters": [float(np.std(d["y"][i])) for i in range(p)]}

    se4 = [["SquaredExponential", [10.0, 1.0]], ["SquaredExponential", [2.0, 5.0]],
           ["SquaredExponential", [3.0, 20.0]], ["SquaredExponential", [5.0, 30.0]]]
    ll_cases = [
        example_case("G1_corot20_verbatim", "corot20x4", [["Periodic", [20.0, 0.5]]], se4, "Constant"),
        example_case("G2_corot167_verbatim", "corot167x4", [["Periodic", [20.0, 0.5]]], se4, "Constant"),
        example_case("G3_corot20_qp_const", "corot20x3", [["QuasiPeriodic", [20.0, 23.0, 0.5]]],
                     [["Constant", [10.0]], ["Constant", [1.0]], ["Constant", [1.0]]], "Constant"),
        example_case("G3_corot167_qp_const", "corot167x3", [["QuasiPeriodic", [20.0, 23.0, 0.5]]],
                     [["Constant", [10.0]], ["Constant", [1.0]], ["Constant", [1.0]]], "Constant"),
        example_case("G4_sun50_initial", "sun50", [["QuasiPeriodic", [10.0, 20.0, 0.5]]],
                     [["Constant", [1.0]]] * 3, "Linear"),
    ]
    # Sun at the (rounded) published posterior medians
    c = example_case("sun50_posterior_medians", "sun50", [["QuasiPeriodic", [18.56, 28.90, 0.70]]],
                     [["Constant", [1.99]], ["Constant", [4.53]], ["Constant", [6.94]]], "Linear")
    c["means"] = [["Linear", [-0.02, -16.86]], ["Linear", [-0.01, -88.82]], ["Linear", [-0.03, 7784.83]]]
    c["jitters"] = [0.80, 0.37, 1.06]
    ll_cases.append(c)
    # mixed networks: every kernel class appears at least once across these
    mixed_a = {"name": "mixed_a_sun50", "data": "sun50", "q": 3,
               "nodes": [["QuasiPeriodic", [15.0, 25.0, 0.8]], ["Linear", [57230.0]], ["WhiteNoise", [0.3]]],
               "weights": [["Constant", [1.5]], ["Polynomial", [0.02, 1e-9, 0.5, 2.0]], ["SquaredExponential", [0.7, 3.0]],
                           ["Linear", [0.01, 57220.0]], ["WhiteNoise", [0.4]], ["Matern52", [1.1, 6.0]],
                           ["Periodic", [0.9, 12.0, 1.3]], ["Constant", [0.5]], ["Laplacian", [0.8, 9.0]]],
               "means": [["Linear", [0.01, -16.0]], ["Keplerian", [11.0, 0.7, 0.2, 0.4, 1.1]],
                         ["Sum", [["Constant", [7784.0]], ["Sine", [0.5, 9.0, 0.3, 0.1]]]]],
               "jitters": [0.9, 0.5, 1.2]}
    mixed_b = {"name": "mixed_b_corot20", "data": "corot20x3", "q": 4,
               "nodes": [["Periodic", [21.0, 0.6]], ["RationalQuadratic", [1.3, 7.0]], ["Matern32", [4.0]],
                         ["Cosine", [17.0]]],
               "weights": [["QuasiPeriodic", [3.0, 30.0, 22.0, 0.9]], ["RQP", [1.2, 0.8, 12.0, 19.0, 0.7]],
                           ["Exponential", [1.4, 5.0]], ["Constant", [0.3]],
                           ["GammaExp", [0.8, 1.5, 4.0]], ["Cosine", [0.6, 15.0]], ["Matern32", [0.9, 3.0]],
                           ["RationalQuadratic", [0.7, 2.0, 8.0]],
                           ["SquaredExponential", [1.0, 2.5]], ["Constant", [0.4]], ["Matern52", [0.5, 2.0]],
                           ["Periodic", [0.2, 9.0, 1.1]]],
               "means": [["Parabola", [1e-3, -109.6, 3.0e6]], None, ["Cubic", [0.0, 1e-9, 1e-5, -0.5]]],
               "jitters": [2.0, 0.4, 0.6]}
    mixed_c = {"name": "mixed_c_synth64", "data": "synth64", "q": 3,
               "nodes": [["RQP", [1.5, 20.0, 9.0, 0.8]], ["Exponential", [6.0]], ["GammaExp", [1.2, 5.0]]],
               "weights": [["Constant", [1.0]], ["Constant", [0.7]], ["Constant", [0.5]],
                           ["SquaredExponential", [1.2, 10.0]], ["Matern32", [0.8, 7.0]], ["Constant", [0.3]],
                           ["Constant", [0.9]], ["Exponential", [0.6, 12.0]], ["Cosine", [0.4, 30.0]]],
               "means": [["Cosine", [0.7, 14.0, 0.2, 0.05]], ["Constant", [0.1]], None],
               "jitters": [0.5, 0.6, 0.7]}
    mixed_d = {"name": "mixed_d_constnode_matern52_poly", "data": "synth64", "q": 3,
               "nodes": [["Constant", [0.9]], ["Matern52", [8.0]], ["Polynomial", [1e-9, 0.5, 2.0]]],
               "weights": [["SquaredExponential", [1.1, 9.0]], ["Constant", [1.0]], ["Constant", [0.2]]] * 3,
               "means": None, "jitters": [0.5, -0.6, 0.7]}
    ll_cases += [mixed_a, mixed_b, mixed_c, mixed_d]
    # failure semantics (arttwo.py:288-289; scipy check_finite)
    bad = dict(ll_cases[4]); bad["name"] = "notpd_negative_cosine_weight"
    bad["nodes"] = [["Cosine", [3.0]]]; bad["weights"] = [["Constant", [30.0]]] * 3; bad["jitters"] = [0.0, 0.0, 0.0]
    data["sun50_noerr"] = {"time": data["sun50"]["time"], "y": data["sun50"]["y"], "yerr": np.zeros_like(data["sun50"]["yerr"])}
    bad["data"] = "sun50_noerr"
    nan = dict(ll_cases[4]); nan["name"] = "nan_parameter"; nan["nodes"] = [["QuasiPeriodic", [float("nan"), 20.0, 0.5]]]
    pzero = dict(ll_cases[4]); pzero["name"] = "zero_period"; pzero["nodes"] = [["QuasiPeriodic", [10.0, 0.0, 0.5]]]
    nomean = dict(ll_cases[4]); nomean["name"] = "G4_sun50_means_none"; nomean["means"] = None
    # order of failures across series (arttwo.py:270-289: series are processed in order and the
    # first failing one decides): -K (negative definite) via Polynomial(w, 0, -1, 1) on the QP node
    def with_weights(name, repl):
        cs = dict(mixed_a); cs["name"] = name; cs["weights"] = list(mixed_a["weights"])
        for i, wspec in repl.items():
            cs["weights"][i] = wspec
        return cs
    neg = ["Polynomial", [50.0, 0.0, -1.0, 1.0]]
    zero = ["Constant", [0.0]]
    second_bad = with_weights("mixed_a_second_series_notpd", {3: neg, 4: zero, 5: zero})
    notpd_then_nan = with_weights("mixed_a_notpd_then_nan", {0: neg, 1: zero, 2: zero, 3: ["Constant", [float("nan")]]})
    nan_then_notpd = with_weights("mixed_a_nan_then_notpd", {0: ["Constant", [float("nan")]], 3: neg, 4: zero, 5: zero})
    nan_mean = dict(mixed_a); nan_mean["name"] = "mixed_a_nan_mean_third_series"
    nan_mean["means"] = list(mixed_a["means"]); nan_mean["means"][2] = ["Constant", [float("nan")]]
    ll_cases += [bad, nan, pzero, nomean, second_bad, notpd_then_nan, nan_then_notpd, nan_mean]
    for case in ll_cases:
        case["ll"] = run_ll(data[case["data"]], case)
        print("%-40s %r" % (case["name"], case["ll"]))
        cases["loglike"].append(case)

    # ---- batched synthetic walkers (SURVEY section 8(d) generators) -----------------------------
    for N, W, wk in ((64, 6, "Constant"), (257, 4, "SquaredExponential"), (512, 3, "Constant"), (512, 2, "SquaredExponential")):
        dname = "synth%d" % N
        theta = synth.make_walkers(W, p=3, q=2, weights=wk, seed=1)
        net = network_of(data[dname], 2)
        lls = []
        for w in range(W):
            nodes, ws, ms, jit = synth.objects_from_row(ref_node, ref_weight, ref_mean, theta[w], 3, 2, wk)
            lls.append(float(net.log_likelihood(nodes, ws, ms, jit)))
        key = "batch_%s_%s" % (dname, wk)
        arrays[key + "_theta"] = theta
        arrays[key + "_ll"] = np.array(lls)
        cases["batch"].append({"name": key, "data": dname, "p": 3, "q": 2, "weights": wk, "W": W})
        print(key, lls)

    # ---- predictions ---------------------------------------------------------------------------
    def run_pred(case, dataset, tstar, key):
        d = data[case["data"]]
        net = network_of(d, case["q"])
        nodes = [build(ref_node, s) for s in case["nodes"]]
        weights = [build(ref_weight, s) for s in case["weights"]]
        means = None if case["means"] is None else [build(ref_mean, s) for s in case["means"]]
        mu, sd, _ = net.prediction(nodes=nodes, weights=weights, means=means, jitters=case["jitters"],
                                   time=tstar.copy(), dataset=dataset)
        arrays[key + "_tstar"] = tstar
        arrays[key + "_mean"] = mu
        arrays[key + "_std"] = sd
        cases["predict"].append({"name": key, "case": case["name"], "dataset": dataset})

    g1, g4 = ll_cases[0], ll_cases[4]
    t = data["corot20x4"]["time"]
    for ds in (1, 2, 3, 4):
        run_pred(g1, ds, np.linspace(t.min() - 5, t.max() + 5, 500), "pred_G1_ds%d" % ds)
    t = data["sun50"]["time"]
    for ds in (1, 2, 3):
        run_pred(g4, ds, np.linspace(t.min() - 5, t.max() + 5, 5000 if ds == 1 else 300), "pred_G4_ds%d" % ds)
    for ds in (1, 2, 3):
        run_pred(mixed_a, ds, np.linspace(t.min() - 3, t.max() + 3, 131), "pred_mixed_a_M131_ds%d" % ds)
        run_pred(mixed_a, ds, np.linspace(t.min() - 3, t.max() + 3, 50), "pred_mixed_a_MeqN_ds%d" % ds)
    t = data["synth257"]["time"]
    syn = {"name": "synth257_qp2_se", "data": "synth257", "q": 2,
           "nodes": [["QuasiPeriodic", [18.56, 28.90, 0.70]], ["QuasiPeriodic", [60.0, 11.0, 1.2]]],
           "weights": [["SquaredExponential", [1.99, 50.0]], ["SquaredExponential", [4.53, 40.0]],
                       ["SquaredExponential", [6.94, 60.0]], ["SquaredExponential", [0.7, 45.0]],
                       ["SquaredExponential", [1.1, 55.0]], ["SquaredExponential", [2.0, 50.0]]],
           "means": [["Linear", [0.01, 0.2]], ["Linear", [-0.02, 0.1]], ["Linear", [0.0, -0.3]]],
           "jitters": [0.8, 0.37, 1.06]}
    syn["ll"] = run_ll(data["synth257"], syn)
    cases["loglike"].append(syn)
    run_pred(syn, 2, np.linspace(t.min() - 5, t.max() + 5, 400), "pred_synth257_ds2")

    # ---- single kernels (node.py / weight.py __call__ through _kernel_matrix) ------------------
    tk = np.array([55001.2, 55001.9, 55003.5, 55004.0, 55007.7, 55011.1, 55012.0, 55019.4, 55030.0])
    ts5 = np.array([55000.0, 55002.5, 55008.8, 55015.0, 55033.3])
    ts9 = tk + 0.37
    arrays["kern_t"], arrays["kern_ts5"], arrays["kern_ts9"] = tk, ts5, ts9
    knet = ref_network(1, tk, np.zeros_like(tk), np.ones_like(tk))
    node_pars = {"Constant": [0.7], "WhiteNoise": [0.4], "SquaredExponential": [3.0], "Periodic": [7.0, 0.8],
                 "QuasiPeriodic": [9.0, 6.0, 0.7], "RationalQuadratic": [1.4, 5.0], "RQP": [1.6, 8.0, 5.0, 0.9],
                 "Cosine": [6.5], "Exponential": [4.0], "Matern32": [3.5], "Matern52": [4.5], "Linear": [55010.0],
                 "GammaExp": [1.3, 5.0], "Polynomial": [1e-9, 0.3, 2.0]}
    for fam, mod, extra in (("node", ref_node, []), ("weight", ref_weight, ["Laplacian"])):
        for name in list(node_pars) + extra:
            pars = list(node_pars["Exponential" if name == "Laplacian" else name])
            if fam == "weight" and name not in ("Constant", "WhiteNoise"):
                pars = [1.7] + pars
            k = getattr(mod, name)(*pars)
            key = "kern_%s_%s" % (fam, name)
            arrays[key + "_train"] = knet._kernel_matrix(k, tk)
            arrays[key + "_s5"] = knet._predict_kernel_matrix(k, ts5)
            arrays[key + "_s9"] = knet._predict_kernel_matrix(k, ts9)
            cases["kernels"].append({"family": fam, "name": name, "pars": pars, "key": key})

    # ---- mean functions ------------------------------------------------------------------------
    mean_specs = [["Constant", [1.5]], ["Linear", [0.03, -2.0]], ["Parabola", [1e-4, -11.0, 3.0e5]],
                  ["Cubic", [1e-9, 1e-5, -0.5, 2.0]], ["Sine", [0.8, 6.0, 0.4, 0.2]], ["Cosine", [0.8, 6.0, 0.4, 0.2]],
                  ["Keplerian", [9.0, 1.2, 0.3, 0.7, 0.9]], ["Keplerian", [9.0, 1.2, 0.0, 0.7, 0.9]],
                  ["Keplerian", [5.0, 2.0, 0.85, 2.0, 3.0]],
                  ["Sum", [["Linear", [0.01, 1.0]], ["Keplerian", [12.0, 0.5, 0.1, 0.2, 0.3]]]]]
    for i, ms in enumerate(mean_specs):
        key = "mean_%d" % i
        arrays[key + "_train"] = build(ref_mean, ms)(tk)
        arrays[key + "_s5"] = build(ref_mean, ms)(ts5)
        cases["means"].append({"spec": ms, "key": key})

    # ---- _covariance_matrix / compute_matrix / _mean -------------------------------------------
    for case in (mixed_a, mixed_b):
        d = data[case["data"]]
        net = network_of(d, case["q"])
        nodes = [build(ref_node, s) for s in case["nodes"]]
        weights = [build(ref_weight, s) for s in case["weights"]]
        means = [build(ref_mean, s) for s in case["means"]]
        key = "cov_" + case["name"]
        p = d["y"].shape[0]
        for pos in range(1, p + 1):
            arrays["%s_pos%d" % (key, pos)] = net._covariance_matrix(nodes, weights, d["time"], pos, add_errors=False)
            arrays["%s_pos%d_err" % (key, pos)] = net._covariance_matrix(nodes, weights, d["time"], pos, add_errors=True)
        tgrid = np.linspace(d["time"].min(), d["time"].max(), 33)
        arrays[key + "_grid33_t"] = tgrid
        arrays[key + "_grid33_pos2"] = net._covariance_matrix(nodes, weights, tgrid, 2, add_errors=False)
        arrays[key + "_compute"] = net.compute_matrix(nodes, weights, d["time"])
        arrays[key + "_compute_nugget_shift"] = net.compute_matrix(nodes, weights, d["time"], nugget=True, shift=True)
        arrays[key + "_meanvec"] = net._mean(means)
        arrays[key + "_meanvec_grid33"] = net._mean(means, tgrid)
        cases["covariance"].append({"case": case["name"], "key": key})

    # ---- legacy art.network (art.py): shared weight class + weights_values, sigma^4 quirk, (mean, std, cov)
    for name, dname, nspecs, wspec, wvals, ctor_jit, call_jit, mspecs in (
            ("art_sun50_se", "sun50", [["QuasiPeriodic", [15.0, 25.0, 0.8]]], ["SquaredExponential", [1.0, 30.0]],
             [1.0, 2.0, 0.5], [0.9, 0.5, 1.2], [0.7, 0.6, 1.0],
             [["Linear", [0.0, -16.9]], ["Linear", [0.0, -88.0]], ["Linear", [0.0, 7784.0]]]),
            ("art_corot20_const_q2", "corot20x3", [["Periodic", [21.0, 0.6]], ["Matern32", [4.0]]], ["Constant", [1.0]],
             [3.0, 1.0, 0.5, 0.2, 0.7, 0.1], [2.0, 0.4, 0.6], [2.0, 0.4, 0.6],
             [["Constant", [12.0]], None, ["Constant", [0.0]]])):
        d = data[dname]
        p = d["y"].shape[0]
        nodes = [build(ref_node, sp) for sp in nspecs]
        means = [build(ref_mean, sp) for sp in mspecs]
        args = []
        for i in range(p):
            args += [d["y"][i].copy(), d["yerr"][i].copy()]
        anet = ref_art_network(nodes, [build(ref_weight, wspec)], list(wvals), means, list(ctor_jit), d["time"].copy(), *args)
        ll = float(anet.log_likelihood(nodes, [build(ref_weight, wspec)], list(wvals), means, list(call_jit)))
        tsx = np.linspace(d["time"].min() - 2, d["time"].max() + 2, 37)
        entry = {"name": name, "data": dname, "nodes": nspecs, "weight": wspec, "weights_values": wvals,
                 "ctor_jitters": ctor_jit, "jitters": call_jit, "means": mspecs, "ll": ll, "datasets": []}
        for ds in range(1, p + 1):
            mu, sd, cov = anet.prediction(nodes=nodes, weights=[build(ref_weight, wspec)], weights_values=list(wvals),
                                          means=means, jitters=list(call_jit), time=tsx.copy(), dataset=ds)
            key = "%s_ds%d" % (name, ds)
            arrays[key + "_tstar"], arrays[key + "_mean"], arrays[key + "_std"], arrays[key + "_cov"] = tsx, mu, sd, np.array(cov)
            entry["datasets"].append(ds)
        arrays[name + "_compute"] = anet.compute_matrix(nodes, build(ref_weight, wspec), list(wvals), d["time"])
        cases["legacy_art"].append(entry)
        print(name, ll)

    flat = {}
    for k, d in data.items():
        for f in ("time", "y", "yerr"):
            flat["%s_%s" % (k, f)] = d[f]
    np.savez_compressed(os.path.join(HERE, "golden_data.npz"), **flat)
    np.savez_compressed(os.path.join(HERE, "golden_arrays.npz"), **arrays)
    with open(os.path.join(HERE, "golden_cases.json"), "w") as f:
        json.dump(cases, f, indent=1)
    print("wrote", len(cases["loglike"]), "loglike,", len(cases["predict"]), "predict,", len(cases["kernels"]), "kernel cases")


if __name__ == "__main__":
    main()
