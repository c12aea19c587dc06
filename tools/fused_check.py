"""GPU check of the fused engine against the Gram-form engine and the CPU oracle on a few seeded shapes (test harness).
Usage (GPU box): python tools/fused_check.py [--big]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def synth(seed, M, N, K, n_true, depth=50):
    rs = np.random.RandomState(seed)
    a = rs.uniform(0.2, 1.0, size=K + n_true)
    Rf = rs.beta(a, a, size=(M, K + n_true))
    unk = rs.uniform(0, 0.9, size=N)
    Ak = rs.dirichlet(np.ones(max(K, 1)), N).T[:K] * (1 - unk)
    Au = rs.dirichlet(np.ones(n_true), N).T * unk
    D = rs.poisson(depth, size=(M, N)) + 1
    cnt = rs.binomial(D, np.clip(Rf @ np.vstack([Ak, Au]), 0, 1))
    return cnt / D, D.astype(np.int64), np.ascontiguousarray(Rf[:, :K])


def main():
    import torch
    import __graft_entry__ as g
    g.build()
    import demethify_b200
    from demethify_b200 import deconvolution as dec
    from oracle import bssmf_numpy as orc
    shapes = [(33, 2, 2, 1, 3, 3, 1e-9), (999, 7, 1, 1, 4, 7, 1e-9), (5000, 16, 6, 2, 5, 20, 1e-9), (4096, 64, 6, 1, 4, 10, 1e-9),
              (3000, 256, 6, 2, 3, 5, 1e-9), (3001, 200, 8, 2, 3, 5, 1e-9), (2000, 100, 3, 2, 3, 6, 1e-9), (800, 8, 4, 1, 500, 20, 1e-2)]
    ok = True
    for (M, N, K, n_u, it1, it2, tol) in shapes:
        X, D, Rk = synth(M + N, M, N, K, max(n_u, 1))
        u0, R0, a0 = orc.draw_init("uniform_", X, D, Rk, n_u, seed=7)
        tr = {}
        uo, ao = orc.solve_partial_reference(u0.copy(), R0, a0.copy(), X, D.astype(float), Rk, n_u, it1, it2, tol, trace=tr)
        res = {}
        for eng in ("gram", "fused"):
            demethify_b200.set_engine(eng)
            t0 = time.time()
            u, a = dec.mdwbssmf_deconv(u0, R0, a0, X, D, Rk, n_u, n_iter1=it1, n_iter2=it2, tol=tol)
            info = dec.last_fit_info()
            res[eng] = (u, a, info)
            da, du = np.abs(a - ao).max(), np.abs(u - uo).max()
            good = info["n_outer"] == tr["n_outer"] and da <= 1e-6 and du <= 1e-6 and abs(info["cost"] - tr["costs"][-1]) <= 1e-9 * tr["costs"][-1]
            ok &= good
            print(f"{'OK ' if good else 'BAD'} M={M} N={N} K={K} n_u={n_u} {eng:5s} engine={info['engine']} n_outer={info['n_outer']}/{tr['n_outer']} "
                  f"d_alpha={da:.2e} d_u={du:.2e} cost_rel={abs(info['cost'] - tr['costs'][-1]) / tr['costs'][-1]:.1e} {time.time() - t0:.2f}s", flush=True)
    demethify_b200.set_engine("auto")
    print("ALL OK" if ok else "FAILURES")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
