import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
import __graft_entry__ as g; g.build()
import torch
from demethify_b200 import bootstrap as bt
import demethify_b200
rs = np.random.RandomState(5)
M, N, K, n_u = 6000, 12, 5, 1
Rf = rs.beta(0.5, 0.5, size=(M, K + n_u)); A = rs.dirichlet(np.ones(K + n_u), N).T
D = rs.poisson(40, size=(M, N)) + 1; X = rs.binomial(D, np.clip(Rf @ A, 0, 1)) / D; Rk = np.ascontiguousarray(Rf[:, :K])
junk = [torch.randn(1 << 22, device='cuda') for _ in range(8)]; del junk
res = []
for trial in range(6):
    os.environ["DMF_HOST_RNG"] = "1" if trial % 2 else "0"
    if len(sys.argv) > 1: demethify_b200.set_engine(sys.argv[1])
    a, u, n = bt.bootstrap_fits(5, n_u, X, D, Rk, "uniform_", 8, 20, 1e-2, None, 3)
    res.append((np.asarray(a).copy(), np.asarray(u).copy(), list(n)))
    junk = torch.full((1 << 24,), float('nan'), device='cuda'); del junk
for t in range(1, 6):
    print(t, "host" if t % 2 else "dev", res[t][2] == res[0][2], np.abs(res[t][0] - res[0][0]).max(), np.abs(res[t][1] - res[0][1]).max())
# single fits, same inputs, repeated: bitwise run-to-run reproducibility per engine
from demethify_b200 import deconvolution as dec
for eng in ("fused", "gram", "stream"):
    demethify_b200.set_engine(eng)
    u0, R0, a0 = dec.init_BSSMF_md("uniform_", X, D, Rk, 1, seed=1)
    outs = []
    for t in range(4):
        u, a = dec.mdwbssmf_deconv(u0, R0, a0, X, D, Rk, 1, n_iter1=8, n_iter2=20, tol=1e-2)
        outs.append((u.copy(), a.copy()))
        junk = torch.full((1 << 22,), float('nan'), device='cuda'); del junk
    print("single", eng, [float(np.abs(o[1] - outs[0][1]).max()) for o in outs[1:]], [float(np.abs(o[0] - outs[0][0]).max()) for o in outs[1:]])
