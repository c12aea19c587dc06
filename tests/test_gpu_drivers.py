"""GPU tests of the rows around the solver: wls_intercept, uniform / SVD / beta inits, the batched
bootstrap driver, the ic sweep drivers and the `demethify` CLI, against the reference's shipped fixtures and the
frozen live-reference vectors (tests/golden)."""
import os

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-6


@pytest.fixture(scope="module")
def pkg():
    import torch
    assert torch.cuda.is_available()
    import __graft_entry__ as g
    g.build()
    import demethify_b200
    return demethify_b200


def test_reference_based_fit_matches_shipped(pkg, shipped):
    from demethify_b200.init_func import wls_all_samples, wls_intercept
    X, D, Rk = shipped["X"], shipped["D"], shipped["Rk"]
    A = wls_all_samples(X, D, Rk, y_is_dx=True)                         # demethify.py:209-213
    assert np.abs(A - shipped["ref_based_alpha"]).max() <= 1e-9
    one = wls_intercept(D[:, 2:3] * X[:, 2:3], D[:, 2:3], Rk)           # single-sample reference signature
    assert one.shape == (5, 1) and np.abs(one[:, 0] - shipped["ref_based_alpha"][:, 2]).max() <= 1e-9


def test_wls_many_regressors_and_samples_vs_oracle(pkg):
    """K = 13 spans two 8 x 8 moment blocks, N = 300 spans two sample slabs, odd N exercises the zero padding."""
    from demethify_b200.init_func import wls_all_samples
    from oracle import bssmf_numpy as orc
    rs = np.random.RandomState(3)
    M, N, K = 4000, 301, 13
    R = rs.beta(0.5, 0.5, size=(M, K))
    A = rs.dirichlet(np.ones(K) * 0.3, N).T
    D = rs.poisson(30, size=(M, N)) + 1
    X = rs.binomial(D, np.clip(R @ A, 0, 1)) / D
    got = wls_all_samples(X, D, R)
    want = np.stack([orc.wls_simplex_fit(X[:, j], D[:, j], R) for j in range(0, N, 37)], axis=1)
    assert np.abs(got[:, ::37] - want).max() <= 1e-8


@pytest.mark.parametrize("tag,opt,seed", [("beta", "beta", 3), ("svd", "SVD", 1), ("uniform", "uniform", 2), ("listseed", "uniform_", [5])])
def test_init_variants(pkg, shipped, live, tag, opt, seed):
    from demethify_b200 import deconvolution as dec
    X, D, Rk = shipped["X"], shipped["D"], shipped["Rk"]
    u0, R0, a0 = dec.init_BSSMF_md(opt, X, D, Rk, 2, seed=seed)
    assert np.abs(u0 - live[f"{tag}_u0"]).max() <= 1e-8 and np.abs(a0 - live[f"{tag}_a0"]).max() <= 1e-8
    u, a = dec.mdwbssmf_deconv(u0, R0, a0, X, D, Rk, 2, n_iter1=10000, n_iter2=20, tol=1e-2)
    assert dec.last_fit_info()["n_outer"] == len(live[f"{tag}_costs"]) - 1
    assert np.abs(a - live[f"{tag}_a"]).max() <= TOL and np.abs(u - live[f"{tag}_u"]).max() <= TOL


def test_bootstrap_matches_reference(pkg, shipped, live_drivers, tmp_path):
    from demethify_b200.bootstrap import bt_ci, bootstrap_seeds
    X, D, Rk = shipped["X"], shipped["D"], shipped["Rk"]
    header = [str(h) for h in shipped["header"]]
    names = [f"s{i}" for i in range(10)]
    assert bootstrap_seeds(1, 5) == [1, 2, 4, 7, 11]
    # shipped test/ci: B = 1
    d1 = tmp_path / "ci1"; d1.mkdir()
    res = bt_ci(95, 1, 1, X, D, Rk, "uniform_", 10000, 20, 1e-2, header, str(d1), names, None, 1)
    lo = np.array([[c[0] for c in res[0][n]] for n in names]).T
    assert np.abs(lo - shipped["ci_alpha_lo"]).max() <= TOL
    ulo = np.array([c[0] for c in res[1]["unknown_cell_1"]])
    assert np.abs(ulo - shipped["ci_u_lo"][:, 0]).max() <= TOL
    # the CSV cells are plain "(lo, hi)" tuples like the shipped file
    txt = open(d1 / "confidence_interval_celltypes_proportions.csv").read()
    assert "np.float64" not in txt and txt.splitlines()[0].startswith("Cell Type,")
    # live: B = 4, 90 % (percentile interpolation across resamples, batched gather fits)
    d4 = tmp_path / "ci4"; d4.mkdir()
    res = bt_ci(90, 4, 1, X, D, Rk, "uniform_", 10000, 20, 1e-2, header, str(d4), names, None, 1)
    lo = np.array([[c[0] for c in res[0][n]] for n in names]).T
    hi = np.array([[c[1] for c in res[0][n]] for n in names]).T
    assert np.abs(lo - live_drivers["bt_alpha_lo"]).max() <= TOL and np.abs(hi - live_drivers["bt_alpha_hi"]).max() <= TOL
    ulo = np.array([c[0] for c in res[1]["unknown_cell_1"]]); uhi = np.array([c[1] for c in res[1]["unknown_cell_1"]])
    assert np.abs(ulo - live_drivers["bt_u_lo"][:, 0]).max() <= TOL and np.abs(uhi - live_drivers["bt_u_hi"][:, 0]).max() <= TOL
    # supervised (n_u = 0) and purity bootstraps
    d0 = tmp_path / "ci0"; d0.mkdir()
    res = bt_ci(80, 3, 0, X, D, Rk, "uniform_", 10000, 20, 1e-2, header, str(d0), names, None, 1)
    lo = np.array([[c[0] for c in res[0][n]] for n in names]).T
    assert np.abs(lo - live_drivers["bt0_alpha_lo"]).max() <= 1e-8
    dp = tmp_path / "cip"; dp.mkdir()
    res = bt_ci(90, 3, 1, X, D, Rk, "uniform_", 20, 50, 1e-2, header, str(dp), names, list(shipped["purity_pct"]), 1)
    lo = np.array([[c[0] for c in res[0][n]] for n in names]).T
    hi = np.array([[c[1] for c in res[0][n]] for n in names]).T
    assert np.abs(lo - live_drivers["btp_alpha_lo"]).max() <= TOL and np.abs(hi - live_drivers["btp_alpha_hi"]).max() <= TOL
    with pytest.raises(TypeError):
        bt_ci(90, 2, 1, X, D, Rk, "uniform_", 5, 5, 1e-2, header, str(dp), names, None, [5])      # `--seed 5` quirk (Q1)


@pytest.mark.parametrize("n_u,N,purity", [(1, 10, False), (2, 70, False), (1, 33, True)])
def test_multiplicity_form_equals_gathered_fits(pkg, n_u, N, purity):
    """A bootstrap resample as (shared source matrix + multiplicities + CSR over u) must give what the row-gathering form and
    the oracle on the materialised resample give (bootstrap.py:28)."""
    import torch
    from demethify_b200 import _lib
    from demethify_b200.engine import DeviceProblem, FitBatch
    from oracle import bssmf_numpy as orc
    rs = np.random.RandomState(17 + N)
    M, K = 3001, 5
    Rf = rs.beta(0.5, 0.5, size=(M, K + n_u))
    A = rs.dirichlet(np.ones(K + n_u), N).T
    D = rs.poisson(40, size=(M, N)) + 1
    X = rs.binomial(D, np.clip(Rf @ A, 0, 1)) / D
    Rk = np.ascontiguousarray(Rf[:, :K])
    pur = rs.uniform(0.3, 0.9, size=N) if purity else None
    prob = DeviceProblem(X, D, Rk)
    dev = prob.device
    fits = []
    for seed in (3, 4, 9):
        idx = np.random.RandomState(seed).randint(0, M, size=(M,))
        order = np.argsort(idx, kind="stable")
        u0 = np.random.RandomState(seed + 100).uniform(size=(M, n_u))
        a0 = np.random.RandomState(seed + 200).dirichlet(np.ones(K + n_u), N).T
        if purity:
            a0 = np.vstack([a0[:K] / a0[:K].sum(0) * pur, a0[K:] / a0[K:].sum(0) * (1 - pur)])
        fits.append((idx, order, u0, a0))
    mode = _lib.DMF_MODE_PURITY if purity else _lib.DMF_MODE_PARTIAL
    it1, it2, tol = (4, 30, 1e-9) if purity else (6, 10, 1e-9)
    cnts = [torch.bincount(torch.from_numpy(f[0]).to(dev), minlength=M) for f in fits]
    bm = FitBatch(prob, n_u, [f[2][f[1]] for f in fits], [f[3] for f in fits], mode=mode, purity=pur, rows=[f[0][f[1]] for f in fits],
                  mult=[c.to(torch.int32) for c in cnts],
                  offs=[torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(c, 0)]).to(torch.int32) for c in cnts])
    assert bm.engine == "gram"
    rm = bm.results(bm.fit(it1, it2, tol))
    bg = FitBatch(prob, n_u, [f[2][f[1]] for f in fits], [f[3] for f in fits], mode=mode, purity=pur, rows=[f[0][f[1]] for f in fits])
    rg = bg.results(bg.fit(it1, it2, tol))
    for (idx, order, u0, a0), (um, am, nm, cm), (ug, ag, ng_, cg) in zip(fits, rm, rg):
        Xb, Db, Rb = X[idx], D[idx].astype(float), Rk[idx]
        tr = {}
        if purity:
            uo, ao = orc.solve_purity(u0.copy(), np.c_[Rb, u0], a0.copy(), Xb, Db, Rb, n_u, pur, it1, it2, tol, trace=tr)
        else:
            uo, ao = orc.solve_partial_reference(u0.copy(), np.c_[Rb, u0], a0.copy(), Xb, Db, Rb, n_u, it1, it2, tol, trace=tr)
        assert nm == ng_ == tr["n_outer"] and abs(cm - tr["costs"][-1]) <= 1e-9 * cm
        assert np.abs(am - ao).max() <= TOL and np.abs(ag - ao).max() <= TOL
        back = np.empty_like(um); back[order] = um
        assert np.abs(back - uo).max() <= TOL


@pytest.mark.parametrize("crit,it1,r", [("AIC", 10000, 5), ("BIC", 10000, 5), ("CCC", 40, 3), ("BCV", 40, 3)])
def test_ic_sweep_matches_reference(pkg, shipped, live_drivers, crit, it1, r):
    from demethify_b200.ic import evaluate_best_ic
    X, D, Rk = shipped["X"], shipped["D"], shipped["Rk"]
    u, a, best, vals = evaluate_best_ic(X, Rk, D, "uniform_", crit, 1, iter1=it1, iter2=20, tol=1e-2, n_restarts=r)
    assert best == int(live_drivers[f"ic_{crit}_best"])
    assert np.allclose(vals, live_drivers[f"ic_{crit}_vals"], rtol=1e-6, atol=1e-9)
    assert np.abs(a - live_drivers[f"ic_{crit}_a"]).max() <= TOL and np.abs(u - live_drivers[f"ic_{crit}_u"]).max() <= TOL
    if crit == "AIC":
        assert best == int(shipped["ic_best_n_u"]) and np.abs(a - shipped["ic_alpha"]).max() <= TOL


def _write_bed_inputs(shipped, root):
    X, D, Rk = shipped["X"], shipped["D"], shipped["Rk"]
    header = [str(h) for h in shipped["header"]]
    M = X.shape[0]
    pos = pd.DataFrame({"chrom": ["chr1"] * M, "start": np.arange(M), "end": np.arange(M) + 1})
    ref = pd.concat([pos, pd.DataFrame(Rk, columns=header)], axis=1)
    ref.to_csv(root / "ref_matrix.bed", sep="\t", index=False, float_format="%.17g")
    files = []
    for j in range(X.shape[1]):
        t = pos.copy()
        t["valid_coverage"] = D[:, j]
        t["count_modified"] = np.rint(X[:, j] * D[:, j]).astype(int)
        t["percent_modified"] = X[:, j] * 100
        path = root / f"sample{j + 1}.bed"
        t.to_csv(path, sep="\t", index=False, float_format="%.17g")
        files.append(str(path))
    return str(root / "ref_matrix.bed"), files


def test_cli_drop_in(pkg, shipped, tmp_path):
    """The reference's documented invocations (README.md:118-218) through demethify_b200.demethify.main."""
    from demethify_b200.demethify import main
    ref, files = _write_bed_inputs(shipped, tmp_path)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        main(["--ref", ref, "--methfreq", *files, "--bedmethyl", "--outdir", "out_ref", "--noprint"])
        got = pd.read_csv(tmp_path / "out_ref" / "celltypes_proportions.csv", index_col=0)
        assert list(got.index) == [str(h) for h in shipped["header"]] and np.abs(got.values - shipped["ref_based_alpha"]).max() <= 1e-8
        main(["--ref", ref, "--methfreq", *files, "--bedmethyl", "--nbunknown", "1", "--outdir", "out_partial", "--noprint"])
        got = pd.read_csv(tmp_path / "out_partial" / "celltypes_proportions.csv", index_col=0)
        prof = pd.read_csv(tmp_path / "out_partial" / "methylation_profile_estimate.csv")
        assert got.index[-1] == "unknown_cell_1" and list(prof.columns) == ["unknown_cell_1"]
        assert np.abs(got.values - shipped["partial_alpha"]).max() <= 1e-5 and np.abs(prof.values - shipped["partial_u"]).max() <= 1e-5
        assert open(tmp_path / "out_partial" / "log.log").read().startswith("Total execution time = ")
        main(["--ref", ref, "--methfreq", *files, "--bedmethyl", "--nbunknown", "1", "--outdir", "out_purity", "--noprint", "--purity",
              *[str(int(p)) for p in shipped["purity_pct"]]])
        got = pd.read_csv(tmp_path / "out_purity" / "celltypes_proportions.csv", index_col=0)
        assert np.abs(got.values - shipped["purity_alpha"]).max() <= 1e-5
        main(["--methfreq", *files, "--bedmethyl", "--nbunknown", "4", "--outdir", "out_unsup", "--noprint"])
        got = pd.read_csv(tmp_path / "out_unsup" / "celltypes_proportions.csv", index_col=0)
        assert np.abs(got.values - shipped["unsup_alpha"]).max() <= 1e-5
    finally:
        os.chdir(cwd)


def test_fp32_mode_multiplicity_form_and_many_unknowns(pkg):
    """fp32 storage / arithmetic (north-star bar: max |d alpha| <= 1e-4) on the two Gram-engine variants that the fp64 tests do not
    reach in fp32: a bootstrap resample in multiplicity form and a 5-unknown fit (one register row per batch)."""
    import torch
    import demethify_b200
    from demethify_b200 import deconvolution as dec
    from demethify_b200.bootstrap import resample_layout
    from demethify_b200.engine import DeviceProblem, FitBatch
    from oracle import bssmf_numpy as orc
    rs = np.random.RandomState(23)
    M, N, K = 4001, 24, 5
    Rf = rs.beta(0.5, 0.5, size=(M, K + 5))
    A = rs.dirichlet(np.ones(K + 5), N).T
    D = rs.poisson(40, size=(M, N)) + 1
    X = rs.binomial(D, np.clip(Rf @ A, 0, 1)) / D
    Rk = np.ascontiguousarray(Rf[:, :K])
    demethify_b200.set_precision("fp32")
    try:
        # (a) multiplicity form, n_u = 1
        idx = np.random.RandomState(5).randint(0, M, size=(M,))
        u0 = np.random.RandomState(6).uniform(size=(M, 1)); a0 = np.random.RandomState(7).dirichlet(np.ones(K + 1), N).T
        prob = DeviceProblem(X, D, Rk)
        order, rows, mult, offs = resample_layout(torch.from_numpy(idx[None]).to(prob.device), M)
        U0 = torch.from_numpy(u0[None]).to(prob.device).gather(1, order.unsqueeze(-1))
        b = FitBatch(prob, 1, U0, a0[None], rows=rows, mult=mult, offs=offs)
        (u, a, n_o, _), = b.results(b.fit(5, 20, 0.0))
        uo, ao = orc.solve_partial_reference(u0.copy(), np.c_[Rk[idx], u0], a0.copy(), X[idx], D[idx].astype(float), Rk[idx], 1, 5, 20, 0.0)
        back = np.empty_like(u); back[order[0].cpu().numpy()] = u
        assert n_o == 5 and np.abs(a - ao).max() <= 1e-4 and np.abs(back - uo).max() <= 1e-3
        # (b) 5 unknown types
        u0, R0, a0 = orc.draw_init("uniform_", X, D, Rk, 5, seed=3)
        uo, ao = orc.solve_partial_reference(u0.copy(), R0, a0.copy(), X, D.astype(float), Rk, 5, 4, 10, 0.0)
        u, a = dec.mdwbssmf_deconv(u0, R0, a0, X, D, Rk, 5, n_iter1=4, n_iter2=10, tol=0.0)
        assert dec.last_fit_info()["engine"] == "gram" and np.abs(a - ao).max() <= 1e-4 and np.abs(u - uo).max() <= 1e-3
    finally:
        demethify_b200.set_precision("fp64")
