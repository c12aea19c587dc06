#!/usr/bin/env python
"""2+ GPU check of CpG-row sharding (run under torchrun on the GPU box):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py
Every rank fits its row range over NCCL; rank 0 gathers u and compares alpha, u and the outer-iteration count with the
CPU oracle on the full problem."""
import os
import sys

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
    from demethify_b200 import sharded
    from demethify_b200.sharded import mdwbssmf_deconv_sharded, row_range
    from oracle import bssmf_numpy as orc
    rs = np.random.RandomState(5)
    M, N, K, n_u = 20011, 48, 6, 2
    a = rs.uniform(0.2, 1.0, size=K + n_u)
    Rf = rs.beta(a, a, size=(M, K + n_u))
    A = rs.dirichlet(np.ones(K + n_u), N).T
    D = rs.poisson(50, size=(M, N)) + 1
    X = rs.binomial(D, np.clip(Rf @ A, 0, 1)) / D
    Rk = np.ascontiguousarray(Rf[:, :K])
    u0, R0, a0 = orc.draw_init("uniform_", X, D, Rk, n_u, seed=3)
    lo, hi = row_range(M, rank, world)
    ok = True
    # (iterations, tolerance, peer exchange, CUDA graph): NCCL and the in-kernel NVLink exchange, eager and as a replayed graph
    # (the exchange count lives on the device, so replays advance it)
    cases = [(6, 20, 1e-9, "0", "0"), (400, 20, 1.0, "0", "0"), (6, 20, 1e-9, "1", "0"), (400, 20, 1.0, "1", "0"), (400, 20, 1.0, "1", "1"),
             (400, 20, 1.0, "0", "1")]
    for it1, it2, tol, peer, graph in cases:
        os.environ["DMF_PEER_XCHG"], os.environ["DMF_SHARDED_GRAPH"] = peer, graph
        if graph == "1":
            os.environ.setdefault("NCCL_GRAPH_REGISTER", "0")
        u, al, n_outer, cost = mdwbssmf_deconv_sharded(u0[lo:hi], a0, X[lo:hi], D[lo:hi], Rk[lo:hi], n_u, n_iter1=it1, n_iter2=it2, tol=tol)
        parts = [None] * world
        dist.all_gather_object(parts, (lo, u, al, n_outer))
        if rank == 0:
            tr = {}
            uo, ao = orc.solve_partial_reference(u0.copy(), R0, a0.copy(), X, D.astype(float), Rk, n_u, it1, it2, tol, trace=tr)
            ufull = np.vstack([p[1] for p in sorted(parts, key=lambda p: p[0])])
            same_alpha = all(np.array_equal(p[2], parts[0][2]) for p in parts)
            da, du = np.abs(al - ao).max(), np.abs(ufull - uo).max()
            good = same_alpha and all(p[3] == tr["n_outer"] for p in parts) and da <= 1e-6 and du <= 1e-6
            ok &= good
            print(f"sharded_check world={world} peer_xchg={peer} graph={graph} it1={it1}: n_outer={n_outer} oracle={tr['n_outer']} max|d alpha|={da:.2e} max|d u|={du:.2e} "
                  f"alpha replicated={same_alpha} peer={sharded.last_info} -> {'OK' if good else 'FAIL'}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
