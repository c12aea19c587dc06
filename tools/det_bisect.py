"""Which stage of the Gram-form outer iteration is not bit-reproducible from run to run? (test harness, GPU box)"""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g; g.build()
import torch
from demethify_b200 import deconvolution as dec, _lib
from demethify_b200.engine import DeviceProblem, FitBatch
rs = np.random.RandomState(5)
M, N, K, n_u = 6000, 12, 5, 1
Rf = rs.beta(0.5, 0.5, size=(M, K + n_u)); A = rs.dirichlet(np.ones(K + n_u), N).T
D = rs.poisson(40, size=(M, N)) + 1; X = rs.binomial(D, np.clip(Rf @ A, 0, 1)) / D; Rk = np.ascontiguousarray(Rf[:, :K])
u0, R0, a0 = dec.init_BSSMF_md("uniform_", X, D, Rk, n_u, seed=1)
def snap(b, tag):
    st = b.states()[0]
    return (tag, b.U.clone().cpu().numpy().tobytes(), b.A.clone().cpu().numpy().tobytes(), (st.cost, st.l_w, st.l_h, st.dmax, st.a1, st.a2), b.ws.clone().cpu().numpy().tobytes())
runs = []
for r in range(5):
    prob = DeviceProblem(X, D, Rk)
    b = FitBatch(prob, n_u, [u0], [a0], engine="gram")
    seq = []
    b.gram_init(); seq.append(snap(b, "init"))
    for it in range(3):
        b.gram_u_inner(20); seq.append(snap(b, f"u_inner{it}"))
        b.gram_panels(False); seq.append(snap(b, f"panels{it}"))
        b.gram_alpha_inner(20); seq.append(snap(b, f"alpha_inner{it}"))
        b.gram_rowgram(False, 0.0); seq.append(snap(b, f"rowgram{it}"))
    runs.append(seq)
    b.close(); del b, prob
    junk = torch.randn(1 << 22, device="cuda"); del junk
for r in range(1, 5):
    for s0, s1 in zip(runs[0], runs[r]):
        if s0[4] != s1[4]:
            a0_, a1_ = np.frombuffer(s0[4], dtype=np.uint8), np.frombuffer(s1[4], dtype=np.uint8)
            w = np.nonzero(a0_ != a1_)[0]
            d0, d1 = np.frombuffer(s0[4][:len(s0[4]) // 8 * 8], dtype=np.float64), np.frombuffer(s1[4][:len(s1[4]) // 8 * 8], dtype=np.float64)
            wd = np.nonzero(d0 != d1)[0]
            print("run", r, "workspace first differs after", s0[0], "bytes", w[:4], "n", len(w), "doubles at", wd[:6], d0[wd[:3]], d1[wd[:3]])
            break
for r in range(1, 5):
    for s0, s1 in zip(runs[0], runs[r]):
        d = [s0[1] != s1[1], s0[2] != s1[2], s0[3] != s1[3], s0[4] != s1[4]]
        if any(d[:3]):
            print("run", r, "first difference after", s0[0], "U/A/state/ws differ:", d, s0[3], s1[3]); break
    else:
        print("run", r, "identical (U, A, state) at every stage; ws differs at:", [s0[0] for s0, s1 in zip(runs[0], runs[r]) if s0[4] != s1[4]][:3])

for r in (1, 2):
    for s0, s1 in zip(runs[0], runs[r]):
        d0, d1 = np.frombuffer(s0[4][:len(s0[4]) // 8 * 8], dtype=np.float64), np.frombuffer(s1[4][:len(s1[4]) // 8 * 8], dtype=np.float64)
        wd = np.nonzero(d0 != d1)[0]
        wd = wd[wd > 64]
        print("run", r, s0[0], "differing doubles beyond the descriptors:", len(wd), wd[:8], (d0[wd[:4]], d1[wd[:4]]))
