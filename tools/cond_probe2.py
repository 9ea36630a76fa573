"""Isolate the factorisation: the SAME ill-conditioned K (assembled on the CPU by the oracle) goes through
agpn_potrf_batched and through LAPACK; compare log-det and quadratic form."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np, torch
from scipy.linalg import cho_factor, cho_solve
import artgpn_oracle as oracle
from artgpn_b200 import _lib
lib = _lib.load()
rng = np.random.default_rng(0)
N = 512
t = 57000.0 + np.sort(rng.uniform(0, 1.0 * N, N))
y = rng.normal(0, 2, N)
nodes = [("QuasiPeriodic", [18.56, 28.9, 0.7]), ("QuasiPeriodic", [60.0, 11.0, 1.2])]
ws = [("Constant", [c]) for c in (1.99, 4.53, 6.94, 0.7, 1.1, 2.0)]
for scale in (1.0, 1e-2, 1e-3, 1e-4, 1e-5):
    sig = rng.uniform(0.1, 0.3, N) * scale
    K = oracle.covariance_block(t, nodes, ws, 2, 2)
    K[np.diag_indices(N)] += sig ** 2 + (1.06 * scale) ** 2
    U = cho_factor(K.copy(), lower=False)
    quad_ref = y @ cho_solve(U, y)
    ld_ref = np.sum(np.log(np.diag(U[0])))
    # exact-ish reference in long double via numpy cholesky of float128? use mpmath-free: refine with higher precision solve
    A = torch.from_numpy(K.copy()[None]).cuda(); r = torch.from_numpy(y.copy()[None]).cuda()
    ld = torch.zeros(1, dtype=torch.float64, device="cuda"); q = torch.zeros(1, dtype=torch.float64, device="cuda")
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    _lib.check(lib.agpn_potrf_batched(0, 1, N, A.data_ptr(), r.data_ptr(), ld.data_ptr(), q.data_ptr(), info.data_ptr(), None))
    torch.cuda.synchronize()
    # perturbation sensitivity: LAPACK on K with entries perturbed by 1e-13 relative
    Kp = K * (1 + 1e-13 * rng.standard_normal(K.shape)); Kp = (Kp + Kp.T) / 2
    Up = cho_factor(Kp, lower=False); quad_p = y @ cho_solve(Up, y)
    print("scale %g cond %.1e: quad rel diff (ours vs LAPACK, same K) %.2e  logdet rel diff %.2e  info %d | LAPACK on K(1+1e-13 noise): quad rel diff %.2e" % (
        scale, np.linalg.cond(K), abs(q.item() - quad_ref) / abs(quad_ref), abs(ld.item() - ld_ref) / abs(ld_ref), info.item(), abs(quad_p - quad_ref) / abs(quad_ref)))
