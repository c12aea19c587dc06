"""CPU tests of the multi-GPU host logic (no GPU needed):
  * the Gram-form restatement (oracle/gram_numpy.py, same arithmetic as csrc/dmf_gram.cuh) against the reference-shaped
    oracle, so the re-association is pinned on CPU too;
  * `RowShardedFit` — the product's CpG-row-sharding orchestration — with world_size 2 over gloo: both ranks end with the
    same alpha, the same outer-iteration count and the rows of u they own, equal to the unsharded run;
  * fit sharding helpers (seed lists, row ranges)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import bssmf_numpy as orc
from oracle.gram_numpy import NumpyShardBackend, momentum_table
from demethify_b200.sharded import RowShardedFit, row_range


def synth(seed, M, N, K, n_true, depth=50):
    rs = np.random.RandomState(seed)
    a = rs.uniform(0.2, 1.0, size=K + n_true)
    Rf = rs.beta(a, a, size=(M, K + n_true))
    unk = rs.uniform(0, 0.9, size=N)
    Ak = rs.dirichlet(np.ones(max(K, 1)), N).T[:K] * (1 - unk)
    Au = rs.dirichlet(np.ones(n_true), N).T * unk
    D = rs.poisson(depth, size=(M, N)) + 1
    cnt = rs.binomial(D, np.clip(Rf @ np.vstack([Ak, Au]), 0, 1))
    return cnt / D, D.astype(np.float64), np.ascontiguousarray(Rf[:, :K])


def test_row_ranges_partition_the_rows():
    for M in (1, 7, 1000, 1_000_003):
        for world in (1, 2, 3, 8):
            edges = [row_range(M, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == M
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            assert max(hi - lo for lo, hi in edges) - min(hi - lo for lo, hi in edges) <= 1


def test_momentum_table_matches_recurrence():
    a, m = momentum_table(50)
    a1 = 1.0
    for t in range(50):
        a0, (a1, beta) = a1, orc._extrapolation(a1, 1.0, 1.0)
        assert a[t] == a0 and a[t + 1] == a1 and min(m[t], 0.9999) == beta


@pytest.mark.parametrize("n_u,it1,it2,tol", [(1, 60, 20, 1e-2), (2, 6, 10, 1e-9)])
def test_gram_form_equals_reference_shape(n_u, it1, it2, tol):
    X, D, Rk = synth(3, 700, 9, 4, max(n_u, 1))
    u0, R0, a0 = orc.draw_init("uniform_", X, D, Rk, n_u, seed=1)
    tr = {}
    uo, ao = orc.solve_partial_reference(u0.copy(), R0, a0.copy(), X, D, Rk, n_u, it1, it2, tol, trace=tr)
    be = NumpyShardBackend(X, D, Rk, n_u, u0, a0)
    (u, a, n_outer, cost), = RowShardedFit(be).fit(it1, it2, tol)
    assert n_outer == tr["n_outer"]
    assert abs(cost - tr["costs"][-1]) <= 1e-10 * cost
    assert np.abs(a - ao).max() <= 1e-10 and np.abs(u - uo).max() <= 1e-10


def test_gram_form_purity_and_unsupervised():
    X, D, Rk = synth(5, 500, 8, 4, 1)
    pur = np.random.RandomState(0).uniform(0.2, 0.9, size=8)
    u0, R0, a0 = orc.draw_init_purity("uniform_", X, D, Rk, 1, pur, seed=2)
    tr = {}
    uo, ao = orc.solve_purity(u0.copy(), R0, a0.copy(), X, D, Rk, 1, pur, 5, 30, 1e-9, trace=tr)
    (u, a, n_outer, _), = RowShardedFit(NumpyShardBackend(X, D, Rk, 1, u0, a0, mode="purity", purity=pur)).fit(5, 30, 1e-9)
    assert n_outer == tr["n_outer"] and np.abs(a - ao).max() <= 1e-10 and np.abs(u - uo).max() <= 1e-10
    # unsupervised: the U gradient is taken at u, not at the extrapolated point (deconvolution.py:163)
    tr = {}
    uo, ao = orc.solve_unsupervised(X, 2, D, "uniform_", 4, 10, 1e-9, seed=3, trace=tr)
    rs = orc.legacy_stream(3)
    u0 = rs.uniform(size=(X.shape[0], 2)); a0 = rs.dirichlet(np.ones(2), X.shape[1]).T
    (u, a, n_outer, _), = RowShardedFit(NumpyShardBackend(X, D, None, 2, u0, a0, mode="unsupervised")).fit(4, 10, 1e-9)
    assert n_outer == tr["n_outer"] and np.abs(a - ao).max() <= 1e-10 and np.abs(u - uo).max() <= 1e-10


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, case, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        X, D, Rk, n_u, u0, a0, mode, pur, it1, it2, tol = case
        lo, hi = row_range(X.shape[0], rank, world)
        be = NumpyShardBackend(X[lo:hi], D[lo:hi], None if Rk is None else Rk[lo:hi], n_u, u0[lo:hi], a0, mode=mode, purity=pur)
        fit = RowShardedFit(be)
        (u, a, n_outer, cost), = fit.fit(it1, it2, tol)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), u=u, a=a, n_outer=n_outer, cost=cost, lo=lo, hi=hi, coll=fit.collectives)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("mode", ["partial", "purity"])
def test_row_sharded_fit_world2_gloo(tmp_path, mode):
    X, D, Rk = synth(11, 901, 6, 3, 2)
    n_u = 2
    pur = np.random.RandomState(1).uniform(0.2, 0.9, size=6) if mode == "purity" else None
    u0, R0, a0 = orc.draw_init("uniform_", X, D, Rk, n_u, seed=4)
    it1, it2, tol = (40, 10, 1e-1) if mode == "partial" else (4, 25, 1e-9)
    (u1, a1, n1, c1), = RowShardedFit(NumpyShardBackend(X, D, Rk, n_u, u0, a0, mode=mode, purity=pur)).fit(it1, it2, tol)
    case = (X, D, Rk, n_u, u0, a0, mode, pur, it1, it2, tol)
    mp.spawn(_worker, args=(2, _free_port(), case, str(tmp_path)), nprocs=2, join=True)
    r = [np.load(tmp_path / f"rank{k}.npz") for k in range(2)]
    assert int(r[0]["n_outer"]) == int(r[1]["n_outer"]) == n1                      # identical termination decision on every rank
    assert np.array_equal(r[0]["a"], r[1]["a"])                                    # alpha is replicated bit for bit
    assert np.abs(r[0]["a"] - a1).max() <= 1e-10 and abs(float(r[0]["cost"]) - c1) <= 1e-10 * c1
    u = np.vstack([r[0]["u"], r[1]["u"]])
    assert (int(r[0]["lo"]), int(r[1]["hi"])) == (0, X.shape[0]) and np.abs(u - u1).max() <= 1e-10
    # communication volume: set-up = 2 collectives (max, sum), every ISSUED outer iteration = 2 sum all-reduces (outer iterations are
    # enqueued in chunks between polls of the done flag, so a few more than n1 can be issued)
    coll = int(r[0]["coll"])
    assert coll == int(r[1]["coll"]) and coll >= 2 + 2 * n1 and (coll - 2) % 2 == 0 and coll <= 2 + 2 * (n1 + 16)


def test_bootstrap_fit_sharding_seed_lists():
    """Fit sharding (bootstrap.py:26-27): rank r takes resamples r, r + world, ...; the union is the reference's seed list."""
    seeds = orc.bootstrap_seed_list(1, 10)
    assert seeds[:5] == [1, 2, 4, 7, 11]
    world = 4
    parts = [seeds[r::world] for r in range(world)]
    assert sorted(s for p in parts for s in p) == sorted(seeds)


def test_percentile_bounds_match_numpy():
    """bt_ci's confidence bounds (bootstrap.py:53-54, :77-78) are taken with torch.quantile on the device; its linear rule must be
    np.percentile's for every stack height, including B = 1 (the shipped test/ci fixture) and ties."""
    from demethify_b200.bootstrap import percentile_bounds_device
    rs = np.random.RandomState(0)
    for B in (1, 2, 3, 4, 7, 100, 2500):
        st = rs.uniform(size=(B, 11, 3))
        st[:, 0, 0] = 0.25                      # ties
        for lo_p, hi_p in ((2.5, 97.5), (5.0, 95.0), (10.0, 90.0)):
            lo, hi = percentile_bounds_device(torch.from_numpy(st), lo_p, hi_p)
            assert np.abs(lo - np.percentile(st, lo_p, axis=0)).max() <= 1e-15
            assert np.abs(hi - np.percentile(st, hi_p, axis=0)).max() <= 1e-15


def _merge_worker(rank, world, port, n_total, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from demethify_b200.bootstrap import bootstrap_seeds, merge_resample_stacks, shard_of
        seeds = bootstrap_seeds(1, n_total)
        mine = shard_of(seeds, rank, world)
        # a fake "fit": the stack entry of resample seed s is a deterministic function of s
        local = torch.tensor([[s * 1.0, s * 0.5 + 1] for s in mine], dtype=torch.float64).reshape(len(mine), 2)
        full = merge_resample_stacks(local, n_total)
        np.save(os.path.join(out_dir, f"merge{rank}.npy"), full.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("n_total,world", [(7, 2), (1, 2), (8, 3)])
def test_fit_sharded_bootstrap_merge_gloo(tmp_path, n_total, world):
    """Fit sharding of the bootstrap (bootstrap.py:26): rank r fits resamples r, r + world, ...; the all-gathered stack holds every
    resample exactly once on every rank (uneven shares and empty shares included), so the percentiles equal the unsharded ones."""
    from demethify_b200.bootstrap import bootstrap_seeds, percentile_bounds_device
    mp.spawn(_merge_worker, args=(world, _free_port(), n_total, str(tmp_path)), nprocs=world, join=True)
    seeds = bootstrap_seeds(1, n_total)
    want = np.array([[s * 1.0, s * 0.5 + 1] for s in seeds])
    got = [np.load(tmp_path / f"merge{r}.npy") for r in range(world)]
    for g in got:
        assert g.shape == want.shape and np.array_equal(np.sort(g, axis=0), np.sort(want, axis=0))
        lo, hi = percentile_bounds_device(torch.from_numpy(g), 5.0, 95.0)
        assert np.allclose(lo, np.percentile(want, 5.0, axis=0), atol=1e-15) and np.allclose(hi, np.percentile(want, 95.0, axis=0), atol=1e-15)
    assert all(np.array_equal(got[0], g) for g in got)


def _sweep_worker(rank, world, port, values, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from demethify_b200.ic import merge_sweep
        local_results, payload, best = {}, {}, float("inf")
        for pos in list(range(len(values)))[rank::world]:
            local_results[pos] = values[pos]
            if values[pos] < best:
                best, payload = values[pos], {pos: (np.full(3, pos), np.full((2, 2), values[pos]))}
        vals, best_pos, pl = merge_sweep(local_results, payload, len(values))
        np.savez(os.path.join(out_dir, f"sweep{rank}.npz"), vals=np.array(vals), best=best_pos, u=pl[0], a=pl[1])
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("values,world", [([5.0, 3.0, 4.0, 3.0, 9.0], 2), ([2.0, 2.0, 1.0, 1.0, 1.0, 7.0, 0.5], 3), ([1.0], 2)])
def test_ic_sweep_sharding_merge_gloo(tmp_path, values, world):
    """evaluate_best_ic shards the n_u sweep over the ranks (ic.py:192-216); the merged criteria are in sweep order and the winner is
    the reference's: the first position whose criterion is strictly below everything before it (ties -> the earlier n_u)."""
    mp.spawn(_sweep_worker, args=(world, _free_port(), values, str(tmp_path)), nprocs=world, join=True)
    want = int(np.argmin(values))                      # first occurrence of the minimum
    for r in range(world):
        z = np.load(tmp_path / f"sweep{r}.npz")
        assert list(z["vals"]) == values and int(z["best"]) == want
        assert np.array_equal(z["u"], np.full(3, want)) and np.array_equal(z["a"], np.full((2, 2), values[want]))


def test_resample_layout_matches_per_fit_numpy():
    """The stacked device-side layout of a wave of resamples (order, sorted rows, multiplicities, CSR offsets) against the
    obvious per-fit numpy construction; positions of one source row keep the reference's order (stable sort)."""
    from demethify_b200.bootstrap import bootstrap_seeds, resample_indices, resample_layout
    M = 1013
    seeds = bootstrap_seeds(1, 6)
    idx = np.stack([resample_indices(s, M) for s in seeds])
    order, rows, mult, offs = resample_layout(torch.from_numpy(idx), M)
    for b in range(len(seeds)):
        o = np.argsort(idx[b], kind="stable")
        assert np.array_equal(order[b].numpy(), o) and np.array_equal(rows[b].numpy(), idx[b][o])
        cnt = np.bincount(idx[b], minlength=M)
        assert np.array_equal(mult[b].numpy(), cnt) and np.array_equal(offs[b].numpy(), np.concatenate([[0], np.cumsum(cnt)]))
        assert int(offs[b, -1]) == M and mult[b].data_ptr() % 16 == 0
        for m in np.flatnonzero(cnt > 1)[:20]:            # the positions offs[m] .. offs[m+1] are exactly the draws of source row m
            p0, p1 = int(offs[b, m]), int(offs[b, m + 1])
            assert np.all(rows[b, p0:p1].numpy() == m) and np.array_equal(np.sort(o[p0:p1]), np.flatnonzero(idx[b] == m))
    order2, rows2, none1, none2 = resample_layout(torch.from_numpy(idx), M, with_csr=False)
    assert none1 is None and none2 is None and torch.equal(order2, order) and torch.equal(rows2, rows)


@pytest.mark.parametrize("n_u", [1, 2])
def test_multiplicity_form_algebra_equals_materialised_resample(n_u):
    """The multiplicity form of a bootstrap resample (shared source matrices + row multiplicities + per-position u; the library's
    csrc/dmf_gram.cuh MULT kernels) against the reference-shaped oracle on the materialised resample X[idx] (bootstrap.py:28)."""
    from oracle.gram_numpy import fit_multiplicity_form
    X, D, Rk = synth(21, 600, 7, 4, n_u)
    idx = np.random.RandomState(9).randint(0, X.shape[0], size=(X.shape[0],))
    u0 = np.random.RandomState(10).uniform(size=(X.shape[0], n_u))
    a0 = np.random.RandomState(11).dirichlet(np.ones(4 + n_u), 7).T
    tr = {}
    uo, ao = orc.solve_partial_reference(u0.copy(), np.c_[Rk[idx], u0], a0.copy(), X[idx], D[idx], Rk[idx], n_u, 12, 10, 1e-3, trace=tr)
    u, a, n_outer, cost = fit_multiplicity_form(X, D, Rk, idx, u0, a0, 12, 10, 1e-3)
    assert n_outer == tr["n_outer"] and abs(cost - tr["costs"][-1]) <= 1e-9 * cost
    assert np.abs(a - ao).max() <= 1e-10 and np.abs(u - uo).max() <= 1e-10
