"""Pins the CPU oracle (oracle/bssmf_numpy.py) to the reference: shipped golden outputs and
live-reference vectors.  Tolerances are rounding-level (the gemm inner order is BLAS-internal)."""
import numpy as np
import pytest

from oracle import bssmf_numpy as orc

TIGHT = 5e-10


def _fit(X, D, Rk, n_u, seed, opt="uniform_", it1=10000, it2=20, tol=1e-2):
    U, R, A = orc.draw_init(opt, X, D, Rk, n_u, seed=seed)
    tr = {}
    U1, A1 = orc.solve_partial_reference(U.copy(), R, A.copy(), X, D, Rk, n_u, it1, it2, tol, trace=tr)
    return U, A, U1, A1, tr


def test_reference_based_matches_shipped(shipped):
    A = orc.reference_based_fit(shipped["X"], shipped["D"], shipped["Rk"])
    assert np.abs(A - shipped["ref_based_alpha"]).max() < 1e-12


@pytest.mark.parametrize("n_u", [1, 2, 4])
def test_partial_reference_matches_live(shipped, live, n_u):
    U0, A0, U, A, tr = _fit(shipped["X"], shipped["D"], shipped["Rk"], n_u, 1)
    assert np.array_equal(U0, live[f"pr{n_u}_u0"]) and np.array_equal(A0, live[f"pr{n_u}_a0"])  # RNG bit-exact
    costs = live[f"pr{n_u}_costs"]
    assert tr["n_outer"] == len(costs) - 1                     # identical iteration count
    assert np.allclose(tr["costs"], costs, rtol=1e-11, atol=0)
    assert np.abs(A - live[f"pr{n_u}_a"]).max() < TIGHT and np.abs(U - live[f"pr{n_u}_u"]).max() < TIGHT


def test_partial_reference_matches_shipped(shipped):
    _, _, U, A, tr = _fit(shipped["X"], shipped["D"], shipped["Rk"], 1, 1)
    assert tr["n_outer"] == 54                                 # SURVEY 2.2 (VERIFIED on the fixture)
    assert np.abs(A - shipped["partial_alpha"]).max() < TIGHT
    assert np.abs(U - shipped["partial_u"]).max() < TIGHT


@pytest.mark.parametrize("tag,opt,seed", [("listseed", "uniform_", [5]), ("beta", "beta", 3), ("svd", "SVD", 1), ("uniform", "uniform", 2)])
def test_init_variants_match_live(shipped, live, tag, opt, seed):
    U0, A0, U, A, tr = _fit(shipped["X"], shipped["D"], shipped["Rk"], 2, seed, opt)
    assert np.abs(U0 - live[f"{tag}_u0"]).max() < 1e-12 and np.abs(A0 - live[f"{tag}_a0"]).max() < 1e-12
    assert tr["n_outer"] == len(live[f"{tag}_costs"]) - 1
    assert np.abs(A - live[f"{tag}_a"]).max() < 1e-8 and np.abs(U - live[f"{tag}_u"]).max() < 1e-8


def test_purity_matches_shipped_and_live(shipped, live):
    X, D, Rk = shipped["X"], shipped["D"], shipped["Rk"]
    pur = 1 - shipped["purity_pct"] / 100.0
    U0, R0, A0 = orc.draw_init_purity("uniform_", X, D, Rk, 1, pur, seed=1)
    assert np.array_equal(U0, live["pur_u0"]) and np.array_equal(A0, live["pur_a0"])
    tr = {}
    U, A = orc.solve_purity(U0, R0, A0, X, D, Rk, 1, pur, 100, 500, 1e-2, trace=tr)
    assert tr["n_outer"] == len(live["pur_costs"]) - 1
    assert np.abs(A - shipped["purity_alpha"]).max() < TIGHT and np.abs(U - shipped["purity_u"]).max() < TIGHT
    assert np.abs(A - live["pur_a"]).max() < TIGHT


def test_purity_two_unknowns_matches_live(live):
    X, D, Rk, pur = live["pur2_X"], live["pur2_D"], live["pur2_Rk"], live["pur2_purity"]
    U0, R0, A0 = orc.draw_init_purity("uniform_", X, D, Rk, 2, pur, seed=4)
    assert np.array_equal(U0, live["pur2_u0"])
    tr = {}
    U, A = orc.solve_purity(U0, R0, A0, X, D, Rk, 2, pur, 30, 40, 1e-3, trace=tr)
    assert tr["n_outer"] == len(live["pur2_costs"]) - 1
    assert np.abs(A - live["pur2_a"]).max() < 1e-9 and np.abs(U - live["pur2_u"]).max() < 1e-9
    assert np.allclose(A[:-2].sum(0), pur) and np.allclose(A[-2:].sum(0), 1 - pur)


def test_unsupervised_matches_shipped_and_live(shipped, live):
    tr = {}
    U, A = orc.solve_unsupervised(shipped["X"], 4, shipped["D"], "uniform_", 10000, 20, 1e-2, seed=1, trace=tr)
    assert tr["n_outer"] == len(live["unsup_costs"]) - 1
    assert np.abs(A - shipped["unsup_alpha"]).max() < 1e-8 and np.abs(U - shipped["unsup_u"]).max() < 1e-8


@pytest.mark.parametrize("tag", ["syn_a", "syn_b", "syn_c", "syn_d"])
def test_synthetic_ragged_match_live(live, tag):
    X, D, Rk = live[f"{tag}_X"], live[f"{tag}_D"], live[f"{tag}_Rk"]
    n_u, it1, it2 = (int(v) for v in live[f"{tag}_cfg"])
    sd = {"syn_a": 21, "syn_b": 22, "syn_c": 23, "syn_d": 24}[tag]
    U0, A0, U, A, tr = _fit(X, D, Rk, n_u, sd, it1=it1, it2=it2, tol=1e-6)
    assert np.array_equal(U0, live[f"{tag}_u0"]) and np.array_equal(A0, live[f"{tag}_a0"])
    assert tr["n_outer"] == len(live[f"{tag}_costs"]) - 1
    assert np.abs(A - live[f"{tag}_a"]).max() < 1e-10 and np.abs(U - live[f"{tag}_u"]).max() < 1e-10


def test_bootstrap_matches_shipped_and_live(shipped, live_drivers):
    X, D, Rk = shipped["X"], shipped["D"], shipped["Rk"]
    # shipped test/ci was produced with B = 1 (lower == upper)
    al, us = orc.bootstrap_fits(1, 1, X, D, Rk, "uniform_", 10000, 20, 1e-2, None, 1)
    lo, hi = orc.percentile_bounds(al, 95)
    assert np.abs(lo - shipped["ci_alpha_lo"]).max() < TIGHT and np.abs(hi - shipped["ci_alpha_hi"]).max() < TIGHT
    ulo, _ = orc.percentile_bounds(us, 95)
    assert np.abs(ulo - shipped["ci_u_lo"]).max() < TIGHT
    # live: B = 4 (triangular seed list 1,2,4,7) and percentile interpolation
    assert orc.bootstrap_seed_list(1, 4) == [1, 2, 4, 7]
    al, us = orc.bootstrap_fits(4, 1, X, D, Rk, "uniform_", 10000, 20, 1e-2, None, 1)
    lo, hi = orc.percentile_bounds(al, 90)
    assert np.abs(lo - live_drivers["bt_alpha_lo"]).max() < 1e-8 and np.abs(hi - live_drivers["bt_alpha_hi"]).max() < 1e-8
    ulo, uhi = orc.percentile_bounds(us, 90)
    assert np.abs(ulo - live_drivers["bt_u_lo"]).max() < 1e-8 and np.abs(uhi - live_drivers["bt_u_hi"]).max() < 1e-8
    # n_u = 0 bootstrap (per-sample NNLS) and purity bootstrap (purity/100 semantics, Q4)
    al, _ = orc.bootstrap_fits(3, 0, X, D, Rk, "uniform_", 10000, 20, 1e-2, None, 1)
    lo, hi = orc.percentile_bounds(al, 80)
    assert np.abs(lo - live_drivers["bt0_alpha_lo"]).max() < 1e-10
    al, _ = orc.bootstrap_fits(3, 1, X, D, Rk, "uniform_", 20, 50, 1e-2, list(shipped["purity_pct"]), 1)
    lo, hi = orc.percentile_bounds(al, 90)
    assert np.abs(lo - live_drivers["btp_alpha_lo"]).max() < 1e-8 and np.abs(hi - live_drivers["btp_alpha_hi"]).max() < 1e-8


@pytest.mark.parametrize("crit,it1", [("AIC", 10000), ("BIC", 10000), ("CCC", 40), ("BCV", 40)])
def test_ic_sweep_matches_live(shipped, live_drivers, crit, it1):
    X, D, Rk = shipped["X"], shipped["D"], shipped["Rk"]
    U, A, best, vals = orc.ic_sweep(X, Rk, D, "uniform_", crit, 1, it1, 20, 1e-2, n_restarts=3 if crit in ("CCC", "BCV") else 5)
    assert best == int(live_drivers[f"ic_{crit}_best"])
    assert np.allclose(vals, live_drivers[f"ic_{crit}_vals"], rtol=1e-8, atol=1e-10)
    assert np.abs(A - live_drivers[f"ic_{crit}_a"]).max() < 1e-7
    if crit == "AIC":
        assert best == int(shipped["ic_best_n_u"])
        assert np.abs(A - shipped["ic_alpha"]).max() < 1e-7 and np.abs(U - shipped["ic_u"]).max() < 1e-7
