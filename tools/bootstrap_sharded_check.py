#!/usr/bin/env python
"""2+ GPU check of the fit-sharded drivers (run under torchrun on the GPU box): bt_ci with the resamples spread over the ranks
must reproduce the reference's confidence bounds frozen in tests/golden/live_drivers.npz (B = 4, 90 %) on every rank, and
evaluate_best_ic with the n_u sweep spread over the ranks the reference's AIC values and winner.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/bootstrap_sharded_check.py"""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as g
    g.build()
    from demethify_b200.bootstrap import bt_ci
    gold = os.path.join(ROOT, "tests", "golden")
    sh = dict(np.load(os.path.join(gold, "fixture_shipped.npz"), allow_pickle=False))
    lv = dict(np.load(os.path.join(gold, "live_drivers.npz"), allow_pickle=False))
    X, D, Rk = sh["X"], sh["D"], sh["Rk"]
    header = [str(h) for h in sh["header"]]
    names = [f"s{i}" for i in range(10)]
    out = tempfile.mkdtemp()
    res = bt_ci(90, 4, 1, X, D, Rk, "uniform_", 10000, 20, 1e-2, header, out, names, None, 1)
    lo = np.array([[c[0] for c in res[0][n]] for n in names]).T
    hi = np.array([[c[1] for c in res[0][n]] for n in names]).T
    ulo = np.array([c[0] for c in res[1]["unknown_cell_1"]])
    uhi = np.array([c[1] for c in res[1]["unknown_cell_1"]])
    err = max(np.abs(lo - lv["bt_alpha_lo"]).max(), np.abs(hi - lv["bt_alpha_hi"]).max(), np.abs(ulo - lv["bt_u_lo"][:, 0]).max(),
              np.abs(uhi - lv["bt_u_hi"][:, 0]).max())
    wrote = os.path.exists(os.path.join(out, "confidence_interval_celltypes_proportions.csv"))
    ok = err <= 1e-6 and wrote == (rank == 0)
    # the n_u sweep of --ic AIC (ic.py:169-218), sharded over the ranks, against the reference's frozen criteria and winner
    from demethify_b200.ic import evaluate_best_ic
    u, a, best, vals = evaluate_best_ic(X, Rk, D, "uniform_", "AIC", 1, iter1=10000, iter2=20, tol=1e-2, n_restarts=5)
    ic_ok = (best == int(lv["ic_AIC_best"]) and np.allclose(vals, lv["ic_AIC_vals"], rtol=1e-6, atol=1e-9)
             and np.abs(a - lv["ic_AIC_a"]).max() <= 1e-6 and np.abs(u - lv["ic_AIC_u"]).max() <= 1e-6)
    ok = ok and ic_ok
    err = max(err, float(np.abs(a - lv["ic_AIC_a"]).max()))
    flags = [None] * world
    dist.all_gather_object(flags, (rank, float(err), bool(ok)))
    if rank == 0:
        print(f"bootstrap_sharded_check world={world}: per-rank (rank, max err vs reference, ok) = {flags} -> {'OK' if all(f[2] for f in flags) else 'FAIL'}",
              flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
