#!/usr/bin/env python
"""Timings of the BASELINE.json configurations other than the bench.py headline (run on the GPU box).  Prints one JSON line
per configuration: wall-clock of the public call, outer iterations executed, fits/s.  Synthetic data, recipe of bench.py.

  python tools/bench_configs.py [c2] [c3] [c4] [c5] [pub] [--resamples B]
"""
import json
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
    Ak = rs.dirichlet(np.ones(K), N).T * (1 - unk)
    Au = rs.dirichlet(np.ones(n_true), N).T * unk
    D = rs.poisson(depth, size=(M, N)) + 1
    X = rs.binomial(D, np.clip(Rf @ np.vstack([Ak, Au]), 0, 1)) / D
    return X, D.astype(np.int64), np.ascontiguousarray(Rf[:, :K]), unk


def main():
    import torch
    import __graft_entry__ as g
    g.build()
    from demethify_b200 import deconvolution as dec
    from demethify_b200.bootstrap import bootstrap_fits
    which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["c2", "c3", "c4"]
    B = int(sys.argv[sys.argv.index("--resamples") + 1]) if "--resamples" in sys.argv else 64
    out = []
    if "c2" in which:      # partial reference, 6 + 2 types x 100k CpGs x 16 samples, CLI defaults (10000 x 20, tol 1e-2)
        X, D, Rk, _ = synth(0, 100_000, 16, 6, 2)
        u0, R0, a0 = dec.init_BSSMF_md("uniform_", X, D, Rk, 2, seed=1)
        for rep in range(2):
            t0 = time.perf_counter()
            dec.mdwbssmf_deconv(u0, R0, a0, X, D, Rk, 2, n_iter1=10000, n_iter2=20, tol=1e-2)
            torch.cuda.synchronize()
            t = time.perf_counter() - t0
        info = dec.last_fit_info()
        out.append({"config": "c2 partial-reference 100k x 16, n_u=2, 10000x20, tol 1e-2", "s_per_fit": t, "n_outer": info["n_outer"],
                    "update_iters_per_s": 40 * info["n_outer"] / t, "engine": info["engine"], "launches": info["launches"]})
    if "c3" in which:      # purity, n_u = 1, 100 x 500, 200k x 64
        X, D, Rk, unk = synth(1, 200_000, 64, 6, 1)
        pur = 1 - unk
        u0, R0, a0 = dec.init_BSSMF_md_p("uniform_", X, D, Rk, 1, pur, seed=1)
        for tol in (1e-2, 0.0):
            for rep in range(2):
                t0 = time.perf_counter()
                dec.mdwbssmf_deconv_p(u0, R0, a0, X, D, Rk, 1, pur, n_iter1=100, n_iter2=500, tol=tol)
                torch.cuda.synchronize()
                t = time.perf_counter() - t0
            info = dec.last_fit_info()
            out.append({"config": f"c3 purity 200k x 64, n_u=1, 100x500, tol {tol}", "s_per_fit": t, "n_outer": info["n_outer"],
                        "update_iters_per_s": 1000 * info["n_outer"] / t, "engine": info["engine"], "launches": info["launches"]})
    if "c4" in which:      # bootstrap: B resamples of 500k x 64, n_u = 1, batched
        X, D, Rk, _ = synth(2, 500_000, 64, 6, 1)
        for n_iter1, tol in ((20, 0.0), (10000, 1e-2)):
            t0 = time.perf_counter()
            alphas, us, n_outer = bootstrap_fits(B, 1, X, D, Rk, "uniform_", n_iter1, 20, tol, None, 1, keep_u=True)
            torch.cuda.synchronize()
            t = time.perf_counter() - t0
            out.append({"config": f"c4 bootstrap {B} resamples of 500k x 64, n_u=1, n_iter1={n_iter1}, tol {tol}", "s_total": t, "fits_per_s": B / t,
                        "mean_outer": float(np.mean(n_outer)), "s_per_1000_resamples": 1000 * t / B})
    if "pub" in which:     # the reference's own published example: 2500 bootstrap resamples of the 350 x 10 fixture (46.47 fits/s)
        fx = dict(np.load(os.path.join(ROOT, "tests", "golden", "fixture_shipped.npz"), allow_pickle=False))
        X, D, Rk = fx["X"], fx["D"], fx["Rk"]
        for rep in range(2):
            t0 = time.perf_counter()
            alphas, us, n_outer = bootstrap_fits(2500, 1, X, D, Rk, "uniform_", 10000, 20, 1e-2, None, 1, keep_u=True)
            torch.cuda.synchronize()
            t = time.perf_counter() - t0
        out.append({"config": "pub bootstrap 2500 resamples of the shipped 350 x 10 fixture, n_u=1, 10000x20, tol 1e-2 (README / notebook cell 29)",
                    "s_total": t, "fits_per_s": 2500 / t, "mean_outer": float(np.mean(n_outer)), "published_fits_per_s_authors_laptop": 46.47})
    if "c5" in which:      # BASELINE config 5: --ic BIC sweep n_u = 0..10 over one matrix of 1M CpGs x 256 samples, K = 6, CLI defaults
        from demethify_b200 import ic as icm
        from tools.time_fused import device_problem
        dev = torch.device("cuda", 0)
        Xd, Dd, Rd = device_problem(torch, dev, 1_000_000, 256, 6, 2, 55)
        X, D, Rk = Xd.cpu().numpy(), Dd.cpu().numpy(), Rd.cpu().numpy()
        del Xd, Dd, Rd
        members = list(range(0, 11))
        t0 = time.perf_counter()
        res = icm.evaluate_best_ic(X, Rk, D, "uniform_", "BIC", 1, 10000, 20, 1e-2, n_restarts=1, n_u_values=members)
        torch.cuda.synchronize()
        t = time.perf_counter() - t0
        out.append({"config": "c5 ic sweep (BIC) n_u = 0..10, 1M x 256, K=6, 10000x20, tol 1e-2, one resident problem, 1 GPU", "s_total": t,
                    "best_n_u": int(res[2]) if res[2] is not None else None, "members": members, "bic": [float(v) for v in res[3]]})
    for o in out:
        print(json.dumps(o), flush=True)




def c4_kernel_times(B=128, M=500_000, N=64, n_u=1):
    """Per-kernel CUDA-event times of the bootstrap batch in multiplicity form (called with `c4prof`)."""
    import torch
    from demethify_b200 import _lib
    from demethify_b200.engine import DeviceProblem, FitBatch
    X, D, Rk, _ = synth(2, M, N, 6, n_u)
    prob = DeviceProblem(X, D, Rk)
    dev = prob.device
    U0, A0, mults, offs, rows = [], [], [], [], []
    for s in range(B):
        rs = np.random.RandomState(s)
        idx = torch.from_numpy(rs.randint(0, M, size=(M,))).to(dev)
        cnt = torch.bincount(idx, minlength=M)
        rows.append(torch.sort(idx).values.to(torch.int32))
        mults.append(cnt.to(torch.int32))
        offs.append(torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(cnt, 0)]).to(torch.int32))
        U0.append(torch.from_numpy(rs.uniform(size=(M, n_u))).to(dev))
        A0.append(rs.dirichlet(np.ones(6 + n_u), N).T)
    batch = FitBatch(prob, n_u, U0, A0, rows=rows, mult=mults, offs=offs)
    batch.gram_init()

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e
    acc = np.zeros(4)
    reps = 6
    for it in range(reps + 2):
        e0 = ev(); batch.gram_u_inner(20); e1 = ev(); batch.gram_panels(False); e2 = ev(); batch.gram_alpha_inner(20); e3 = ev()
        batch.gram_rowgram(False, 0.0); e4 = ev()
        torch.cuda.synchronize()
        if it >= 2:
            acc += [e0.elapsed_time(e1), e1.elapsed_time(e2), e2.elapsed_time(e3), e3.elapsed_time(e4)]
    acc /= reps
    bytes_pass = M * (8 * (N + 6 + n_u) + 2 * N)
    print(json.dumps({"config": f"c4prof {B} fits x {M} x {N}, n_u={n_u}: ms per launch over ALL fits", "geometry": batch.geometry(),
                      "u_inner_mult": acc[0], "panels": acc[1], "alpha_inner": acc[2], "rowgram+cost_cross": acc[3],
                      "ms_per_fit_outer": acc.sum() / B, "panel_GBps_algorithmic": B * bytes_pass / acc[1] / 1e6,
                      "rowgram_GBps_algorithmic": B * bytes_pass / acc[3] / 1e6}))


if __name__ == "__main__" and "c4prof" in sys.argv:
    c4_kernel_times()


if __name__ == "__main__" and "c4prof" not in sys.argv:
    main()
