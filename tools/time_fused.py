#!/usr/bin/env python
"""Kernel-experiment harness (GPU box): CUDA-event times of the fused pass and alpha_inner_kernel for a few shapes / batch sizes,
synthetic data drawn on the device.  One JSON line per case.  Usage:
  [DMF_LIB=demethify_b200/variants/lib_<tag>.so] python tools/time_fused.py [head] [c4] [c4b] [c2] [wave]
    head  1M x 256, K = 6, n_u = 2, one fit                 (the bench.py headline shape)
    c4    500k x 64, K = 6, n_u = 1, one fit                 (one resample of BASELINE config 4)
    c4b   the same, 32 fits of a batch on their own copies  (a bootstrap wave in materialised form)
    c2    100k x 16, K = 6, n_u = 2, one fit
    s2    500k x 128, K = 6, n_u = 2, one fit;  s1u2  500k x 64, K = 6, n_u = 2, one fit
    wave  phase breakdown (host wall clock, synchronised) of one bootstrap_fits wave of 64 resamples of the c4 shape
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def device_problem(torch, dev, M, N, K, n_u, seed):
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    Kt = K + n_u
    conc = torch.rand(Kt, device=dev, generator=gen, dtype=torch.float64) * 0.8 + 0.2
    g1 = torch._standard_gamma(conc.expand(M, Kt).contiguous(), generator=gen)
    g2 = torch._standard_gamma(conc.expand(M, Kt).contiguous(), generator=gen)
    Rf = g1 / (g1 + g2)
    unk = torch.rand(N, device=dev, generator=gen, dtype=torch.float64) * 0.9
    ek = -torch.log1p(-torch.rand(K, N, device=dev, generator=gen, dtype=torch.float64))
    eu = -torch.log1p(-torch.rand(n_u, N, device=dev, generator=gen, dtype=torch.float64))
    A = torch.cat([ek / ek.sum(0) * (1 - unk), eu / eu.sum(0) * unk], 0)
    D = torch.poisson(torch.full((M, N), 50.0, device=dev), generator=gen).to(torch.int64) + 1
    X = torch.binomial(D.to(torch.float64), (Rf @ A).clamp_(0, 1), generator=gen) / D.to(torch.float64)
    return X, D, Rf[:, :K].contiguous()


def time_case(torch, name, M, N, K, n_u, n_fits, reps=20, n_iter2=20):
    from demethify_b200.engine import DeviceProblem, FitBatch
    dev = torch.device("cuda", 0)
    X, D, Rk = device_problem(torch, dev, M, N, K, n_u, 99)
    prob = DeviceProblem(X, D, Rk)
    del X, D
    probs = [prob] + [prob.gathered(torch.arange(M, device=dev, dtype=torch.int32)) for _ in range(n_fits - 1)]
    rs = np.random.RandomState(3)
    U0 = torch.from_numpy(rs.uniform(size=(n_fits, M, n_u))).to(dev)
    A0 = np.stack([rs.dirichlet(np.ones(K + n_u), N).T for _ in range(n_fits)])
    batch = FitBatch(probs, n_u, U0, A0)
    eng = batch.engine
    batch.gram_init()

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e
    acc = np.zeros(2)
    for it in range(reps + 3):
        e0 = ev(); batch.fused_pass(n_iter2, 0.0); e1 = ev(); batch.gram_alpha_inner(n_iter2); e2 = ev()
        torch.cuda.synchronize()
        if it >= 3:
            acc += [e0.elapsed_time(e1), e1.elapsed_time(e2)]
    acc /= reps
    batch.fused_finish(0.0)
    sT, sW = 8, 2
    by = n_fits * M * (sT * (N + K + 4 * n_u) + sW * N)
    tri = n_u * (n_u + 1) // 2
    fl = 2.0 * n_fits * M * N * (K + 2 + n_u + tri + 1 + n_u + n_u * K + tri)
    print(json.dumps({"case": name, "lib": os.environ.get("DMF_LIB", ""), "engine": eng, "M": M, "N": N, "K": K, "n_u": n_u, "fits": n_fits,
                      "geometry": batch.geometry(), "fused_ms": acc[0], "alpha_inner_ms": acc[1], "ms_per_fit_outer": acc.sum() / n_fits,
                      "GBps_algorithmic": by / acc[0] / 1e6, "fp64_tflops_algorithmic": fl / acc[0] / 1e9}), flush=True)
    batch.close()


def wave_breakdown(torch, B=64):
    """Host wall clock of the pieces of one materialised bootstrap wave (each followed by a device synchronisation)."""
    from demethify_b200 import bootstrap as bs, _lib
    from demethify_b200.engine import DeviceProblem, FitBatch
    dev = torch.device("cuda", 0)
    M, N, K, n_u = 500_000, 64, 6, 1
    X, D, Rk = device_problem(torch, dev, M, N, K, n_u, 4321)
    prob = DeviceProblem(X, D, Rk)
    del X, D
    seeds = bs.bootstrap_seeds(1, B)
    chunk = [(s, s) for s in seeds]
    out = {"case": "wave", "resamples": B}

    def lap(key, t0):
        torch.cuda.synchronize()
        out[key] = time.perf_counter() - t0
        return time.perf_counter()
    for rep in range(1 if B > 128 else 2):
        t = time.perf_counter()
        idx_d, u0_d, A0 = bs.device_draws(chunk, M, n_u, K, N, dev, True)
        t = lap("draws_s", t)
        probs = prob.gathered_many(idx_d)
        t = lap("gather_s", t)
        batch = FitBatch(probs, n_u, u0_d, A0)
        t = lap("batch_create_s", t)
        states = batch.fit(10000, 20, 1e-2)
        t = lap("fit_s", t)
        U_d, A_d = batch.stacked_current(states)
        t = lap("stack_s", t)
        out["mean_outer"] = float(np.mean([s.n_outer for s in states]))
        out["max_outer"] = int(max(s.n_outer for s in states))
        out["launches"] = batch.launch_count()
        out["fit_ms_per_fit_outer"] = 1e3 * out["fit_s"] / (B * out["mean_outer"])
        batch.close()
        del batch, probs, U_d, A_d, idx_d, u0_d
    print(json.dumps(out), flush=True)


def main():
    import torch
    import __graft_entry__ as g
    g.build()
    which = sys.argv[1:] or ["head", "c4", "c4b"]
    if "head" in which:
        time_case(torch, "head", 1_000_000, 256, 6, 2, 1)
    if "c4" in which:
        time_case(torch, "c4", 500_000, 64, 6, 1, 1)
    if "c4b" in which:
        time_case(torch, "c4b", 500_000, 64, 6, 1, 32)
    if "c2" in which:
        time_case(torch, "c2", 100_000, 16, 6, 2, 1, reps=50)
    if "s2" in which:
        time_case(torch, "s2", 500_000, 128, 6, 2, 1)
    if "s1u2" in which:
        time_case(torch, "s1u2", 500_000, 64, 6, 2, 1)
    if "wave" in which:
        wave_breakdown(torch, int(os.environ.get("DMF_WAVE_B", "64")))


if __name__ == "__main__":
    main()
