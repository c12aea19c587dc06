"""Tiny fits for compute-sanitizer (racecheck / memcheck): python tools/race_check.py <engine>"""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g; g.build()
import demethify_b200
from demethify_b200 import deconvolution as dec
rs = np.random.RandomState(5)
M, N, K, n_u = 700, 12, 5, int(sys.argv[2]) if len(sys.argv) > 2 else 1
Rf = rs.beta(0.5, 0.5, size=(M, K + n_u)); A = rs.dirichlet(np.ones(K + n_u), N).T
D = rs.poisson(40, size=(M, N)) + 1; X = rs.binomial(D, np.clip(Rf @ A, 0, 1)) / D; Rk = np.ascontiguousarray(Rf[:, :K])
demethify_b200.set_engine(sys.argv[1] if len(sys.argv) > 1 else "gram")
u0, R0, a0 = dec.init_BSSMF_md("uniform_", X, D, Rk, n_u, seed=1)
u, a = dec.mdwbssmf_deconv(u0, R0, a0, X, D, Rk, n_u, n_iter1=3, n_iter2=5, tol=0.0)
print("done", dec.last_fit_info())
