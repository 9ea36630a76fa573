"""Synthetic workloads of SURVEY.md section 8(d) (data and walker generators shared by bench.py,
the tests and the golden-vector script).  Pure numpy; no reference or oracle code."""
import numpy as np

# published Sun posterior medians (examples/Sun/12_cornerPlot_logL.png)
QP_CENTRES = [(18.56, 28.90, 0.70), (60.0, 11.0, 1.2), (120.0, 45.0, 0.9)]
WEIGHT_CENTRES = [1.99, 4.53, 6.94, 0.7, 1.1, 2.0, 1.5, 0.9, 3.0]
JITTER_CENTRES = [0.80, 0.37, 1.06]

CONFIGS = {
    # name: N, p, q, W  (BASELINE.json configs[1..3])
    "C2": dict(N=167, p=3, q=1, W=256),
    "C3": dict(N=2048, p=3, q=2, W=128),
    "C4": dict(N=8192, p=3, q=3, W=32),
    # BASELINE.json configs[4]: prediction on a 10^5-point grid per series from a 3 x 4096 fitted network
    "C5": dict(N=4096, p=3, q=2, W=1, M=100000),
}


def prediction_grid(t, M):
    """t* = linspace(t.min() - 5, t.max() + 5, M)."""
    return np.linspace(t.min() - 5.0, t.max() + 5.0, M)


def make_data(N, p=3, seed=0):
    """t[N] irregular (mean spacing 1 d, BJD-like offset), y[p][N] ~ N(0, 2), sigma[p][N] ~ U(0.1, 0.3)."""
    rng = np.random.default_rng(seed)
    t = 57000.0 + np.sort(rng.uniform(0, 1.0 * N, N))
    y = np.empty((p, N))
    sig = np.empty((p, N))
    for i in range(p):
        y[i] = rng.normal(0, 2, N)
        sig[i] = rng.uniform(0.1, 0.3, N)
    return t, y, sig


def make_walkers(W, p=3, q=2, weights="Constant", seed=1):
    """theta[W][D] for q QuasiPeriodic nodes, p*q `weights` kernels (Constant or SquaredExponential),
    p Linear means and p jitters, centred on the published Sun posterior, x lognormal(0, 0.1).
    Row layout = include/artgpn_b200.h (nodes | weights j + q*i | means | jitters)."""
    rng = np.random.default_rng(seed)
    cols = []
    for j in range(q):
        for c in QP_CENTRES[j % len(QP_CENTRES)]:
            cols.append(c * rng.lognormal(0, 0.1, W))
    for i in range(p):
        for j in range(q):
            c = WEIGHT_CENTRES[(j + q * i) % len(WEIGHT_CENTRES)]
            cols.append(c * rng.lognormal(0, 0.1, W))
            if weights == "SquaredExponential":
                cols.append(50.0 * rng.lognormal(0, 0.1, W))
            elif weights != "Constant":
                raise ValueError("weights must be 'Constant' or 'SquaredExponential'")
    for i in range(p):
        cols.append(rng.normal(0, 0.02, W))
        cols.append(rng.normal(0, 0.5, W))
    for i in range(p):
        cols.append(JITTER_CENTRES[i % len(JITTER_CENTRES)] * rng.lognormal(0, 0.1, W))
    return np.ascontiguousarray(np.stack(cols, axis=1))


# A second synthetic structure that the specialised assembly kernel does NOT cover (generic assemble_kernel):
# nodes Matern52(ell) + RationalQuadratic(alpha, ell), Constant weights, Linear means.
M52RQ_CENTRES = [(20.0,), (1.5, 40.0)]


def make_walkers_m52rq(W, p=3, seed=1):
    """theta[W][D] for nodes [Matern52(ell), RationalQuadratic(alpha, ell)], 2p Constant weights, p Linear means,
    p jitters (same centres / spreads as make_walkers for the shared parts)."""
    rng = np.random.default_rng(seed)
    cols = []
    for cs in M52RQ_CENTRES:
        for c in cs:
            cols.append(c * rng.lognormal(0, 0.1, W))
    for i in range(p):
        for j in range(2):
            cols.append(WEIGHT_CENTRES[(j + 2 * i) % len(WEIGHT_CENTRES)] * rng.lognormal(0, 0.1, W))
    for i in range(p):
        cols.append(rng.normal(0, 0.02, W))
        cols.append(rng.normal(0, 0.5, W))
    for i in range(p):
        cols.append(JITTER_CENTRES[i % len(JITTER_CENTRES)] * rng.lognormal(0, 0.1, W))
    return np.ascontiguousarray(np.stack(cols, axis=1))


def objects_from_row_m52rq(mod_node, mod_weight, mod_mean, row, p=3):
    row = list(map(float, row))
    nodes = [mod_node.Matern52(row[0]), mod_node.RationalQuadratic(row[1], row[2])]
    k = 3
    ws = [mod_weight.Constant(row[k + i]) for i in range(2 * p)]
    k += 2 * p
    means = [mod_mean.Linear(row[k + 2 * i], row[k + 2 * i + 1]) for i in range(p)]
    k += 2 * p
    return nodes, ws, means, row[k:k + p]


def structure_names(p=3, q=2, weights="Constant"):
    """Class names of the synthetic network: (node names[q], weight names[p*q], mean names[p])."""
    return ["QuasiPeriodic"] * q, [weights] * (p * q), ["Linear"] * p


def objects_from_row(mod_node, mod_weight, mod_mean, row, p=3, q=2, weights="Constant"):
    """Kernel / mean objects (from any module triple with the reference's class names) + jitters
    for one theta row produced by make_walkers."""
    row = list(map(float, row))
    k = 0
    nodes = []
    for j in range(q):
        nodes.append(mod_node.QuasiPeriodic(*row[k:k + 3]))
        k += 3
    ws = []
    for _ in range(p * q):
        if weights == "Constant":
            ws.append(mod_weight.Constant(row[k]))
            k += 1
        else:
            ws.append(mod_weight.SquaredExponential(row[k], row[k + 1]))
            k += 2
    means = []
    for _ in range(p):
        means.append(mod_mean.Linear(row[k], row[k + 1]))
        k += 2
    jit = row[k:k + p]
    return nodes, ws, means, jit
