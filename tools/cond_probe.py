"""Conditioning study of the likelihood path, one B200.  usage:

    python tools/cond_probe.py golden      log-likelihood vs the reference goldens of tests/golden/golden_cond.* in every
                                           mode: default / exact assembly x INT8-sliced / DMMA update (set by env per run)
    python tools/cond_probe.py factor      the SAME float64 K through agpn_potrf_batched and LAPACK, both against a
                                           long-double Cholesky of that K (isolates the factorisation) + LAPACK on K with
                                           1e-13 relative noise (what entry-level differences do)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np


def golden():
    from artgpn_b200 import synth, node, weight, mean
    from artgpn_b200.arttwo import network
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_cond.json")))
    a = np.load(os.path.join(ROOT, "tests", "golden", "golden_cond.npz"))
    p, q = g["p"], g["q"]
    eps = np.finfo(float).eps
    print("update path from AGPN_SYRK=%s" % os.environ.get("AGPN_SYRK", "(default i8)"))
    for k, c in enumerate(g["cases"]):
        th = a["theta_%d" % k]
        args = []
        for i in range(p):
            args += [a["y"][i], a["yerr"][i] * c["scale"]]
        out = []
        for exact in (False, True):
            net = network(q, a["time"], *args, exact=exact)
            o = synth.objects_from_row(node, weight, mean, th, p, q, g["weights"])
            try:
                v = net.log_likelihood(*o)
            except ValueError:
                v = float("nan")
            out.append(abs(v - c["ll"]) / abs(c["ll"]))
            net.close()
        cond = max(c["cond"])
        print("%-26s cond %.1e  rel diff default %.2e  exact %.2e   (cond eps = %.1e; ratio default %.1f exact %.1f)" % (
            c["name"], cond, out[0], out[1], cond * eps, out[0] / (cond * eps), out[1] / (cond * eps)))


def chol_ld(K, y):
    ld = np.longdouble
    A = K.astype(ld).copy()
    n = A.shape[0]
    z = y.astype(ld).copy()
    logdet = ld(0)
    for j in range(n):
        d = np.sqrt(A[j, j])
        logdet += np.log(d)
        A[j:, j] /= d
        z[j] /= d
        if j + 1 < n:
            A[j + 1:, j + 1:] -= np.outer(A[j + 1:, j], A[j + 1:, j])
            z[j + 1:] -= A[j + 1:, j] * z[j]
    return float(z @ z), float(logdet)


def factor():
    import torch
    from scipy.linalg import cho_factor, cho_solve
    import artgpn_oracle as oracle
    from artgpn_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(0)
    N = 384
    t = 57000.0 + np.sort(rng.uniform(0, 1.0 * N, N))
    y = rng.normal(0, 2, N)
    nodes = [("QuasiPeriodic", [18.56, 28.9, 0.7]), ("QuasiPeriodic", [60.0, 11.0, 1.2])]
    ws = [("Constant", [c]) for c in (1.99, 4.53, 6.94, 0.7, 1.1, 2.0)]
    for scale in (1.0, 1e-2, 1e-3, 1e-4, 1e-5):
        sig = rng.uniform(0.1, 0.3, N) * scale
        K = oracle.covariance_block(t, nodes, ws, 2, 2)
        K[np.diag_indices(N)] += sig ** 2 + (1.06 * scale) ** 2
        q_true, ld_true = chol_ld(K, y)
        U = cho_factor(K.copy(), lower=False)
        q_lap = y @ cho_solve(U, y)
        ld_lap = np.sum(np.log(np.diag(U[0])))
        A = torch.from_numpy(K.copy()[None]).cuda()
        r = torch.from_numpy(y.copy()[None]).cuda()
        ldv = torch.zeros(1, dtype=torch.float64, device="cuda")
        q = torch.zeros(1, dtype=torch.float64, device="cuda")
        info = torch.zeros(1, dtype=torch.int32, device="cuda")
        _lib.check(lib.agpn_potrf_batched(0, 1, N, A.data_ptr(), r.data_ptr(), ldv.data_ptr(), q.data_ptr(), info.data_ptr(), None))
        torch.cuda.synchronize()
        Kp = K * (1 + 1e-13 * rng.standard_normal(K.shape))
        Kp = (Kp + Kp.T) / 2
        q_p = y @ cho_solve(cho_factor(Kp, lower=False), y)
        print("cond %.1e: quad rel err vs long double: LAPACK %.2e  ours %.2e (info %d) | logdet: LAPACK %.2e ours %.2e | "
              "LAPACK on K (1 + 1e-13 noise): quad moves by %.2e" % (
                  np.linalg.cond(K), abs(q_lap - q_true) / q_true, abs(q.item() - q_true) / q_true, info.item(),
                  abs(ld_lap - ld_true) / abs(ld_true), abs(ldv.item() - ld_true) / abs(ld_true), abs(q_p - q_lap) / abs(q_lap)))


if __name__ == "__main__":
    {"golden": golden, "factor": factor}[sys.argv[1] if len(sys.argv) > 1 else "golden"]()
