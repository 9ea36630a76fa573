"""Accuracy against the TRUTH (long-double Cholesky of the same float64 K): LAPACK's error vs ours."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np, torch
from scipy.linalg import cho_factor, cho_solve
import artgpn_oracle as oracle
from artgpn_b200 import _lib
lib = _lib.load()
ld = np.longdouble

def chol_ld(K, y):
    A = K.astype(ld).copy(); n = A.shape[0]; z = y.astype(ld).copy()
    logdet = ld(0)
    for j in range(n):
        d = np.sqrt(A[j, j]); logdet += np.log(d)
        A[j:, j] /= d
        z[j] /= d
        if j + 1 < n:
            A[j + 1:, j + 1:] -= np.outer(A[j + 1:, j], A[j + 1:, j])
            z[j + 1:] -= A[j + 1:, j] * z[j]
    return float(z @ z), float(logdet)

rng = np.random.default_rng(0)
N = 384
t = 57000.0 + np.sort(rng.uniform(0, 1.0 * N, N))
y = rng.normal(0, 2, N)
nodes = [("QuasiPeriodic", [18.56, 28.9, 0.7]), ("QuasiPeriodic", [60.0, 11.0, 1.2])]
ws = [("Constant", [c]) for c in (1.99, 4.53, 6.94, 0.7, 1.1, 2.0)]
for scale in (1.0, 1e-2, 1e-3, 1e-4):
    sig = rng.uniform(0.1, 0.3, N) * scale
    K = oracle.covariance_block(t, nodes, ws, 2, 2)
    K[np.diag_indices(N)] += sig ** 2 + (1.06 * scale) ** 2
    q_true, ld_true = chol_ld(K, y)
    U = cho_factor(K.copy(), lower=False)
    q_lap = y @ cho_solve(U, y); ld_lap = np.sum(np.log(np.diag(U[0])))
    A = torch.from_numpy(K.copy()[None]).cuda(); r = torch.from_numpy(y.copy()[None]).cuda()
    ldv = torch.zeros(1, dtype=torch.float64, device="cuda"); q = torch.zeros(1, dtype=torch.float64, device="cuda")
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    _lib.check(lib.agpn_potrf_batched(0, 1, N, A.data_ptr(), r.data_ptr(), ldv.data_ptr(), q.data_ptr(), info.data_ptr(), None))
    torch.cuda.synchronize()
    print("cond %.1e: quad rel err  LAPACK %.2e  ours %.2e | logdet rel err  LAPACK %.2e  ours %.2e" % (
        np.linalg.cond(K), abs(q_lap - q_true) / q_true, abs(q.item() - q_true) / q_true,
        abs(ld_lap - ld_true) / abs(ld_true), abs(ldv.item() - ld_true) / abs(ld_true)))
