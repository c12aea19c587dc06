"""Generate tests/golden/*.npz by running the LIVE reference (build container only).

    python tests/golden/make_golden.py            # needs /root/reference

The GPU box has no /root/reference: tests read only the committed .npz files.
Two kinds of vectors are frozen:
  * `fixture_*.npz`  — the reference's own shipped inputs (test/output_gen) and golden
    outputs (test/{output_ref_based,output_partial_ref,purity,unsupervised,ci,model_selection})
    parsed into arrays;
  * `live_*.npz`     — outputs of the reference's functions imported from /root/reference,
    on the fixture inputs and on small seeded synthetic inputs, with outer-iteration
    counts and per-outer cost traces (captured by wrapping cost_f_w).
"""
import os
import re
import sys
import types
import warnings

import numpy as np
import pandas as pd

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
warnings.filterwarnings("ignore")


def _import_reference():
    sys.path.insert(0, REF)
    for name in ("colorcet", "seaborn", "matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    from demethify import deconvolution, bootstrap, ic, init_func  # noqa
    return deconvolution, bootstrap, ic, init_func


def load_fixture_inputs():
    g = os.path.join(REF, "test", "output_gen")
    ref_df = pd.read_csv(os.path.join(g, "ref_matrix.bed"), sep="\t").iloc[:, 3:]
    xs, ds = [], []
    for i in range(1, 11):
        t = pd.read_csv(os.path.join(g, f"sample{i}.bed"), sep="\t")
        xs.append(t["percent_modified"].values / 100)
        ds.append(t["valid_coverage"].values)
    return np.column_stack(xs), np.column_stack(ds), np.ascontiguousarray(ref_df.values), list(ref_df.columns)


def parse_tuple_csv(path, index_col):
    df = pd.read_csv(path, index_col=index_col)
    lo = np.zeros(df.shape)
    hi = np.zeros(df.shape)
    for i in range(df.shape[0]):
        for j in range(df.shape[1]):
            # numpy 2 writes the cells as "(np.float64(a), np.float64(b))" 
            a, b = re.findall(r"[-+]?\d[\d.]*(?:[eE][-+]?\d+)?", df.iloc[i, j].replace("float64", ""))
            lo[i, j], hi[i, j] = float(a), float(b)
    return lo, hi


class CostSpy:
    """Wraps deconvolution.cost_f_w to record every evaluation (outer-iteration count + trace)."""

    def __init__(self, mod):
        self.mod, self.orig, self.vals = mod, mod.cost_f_w, []

    def __enter__(self):
        def spy(*a):
            v = self.orig(*a)
            self.vals.append(float(v))
            return v
        self.mod.cost_f_w = spy
        return self

    def __exit__(self, *exc):
        self.mod.cost_f_w = self.orig


def synth(seed, M, N, K, n_true, depth=50):
    """Small in-silico mixture following test/gen_data.ipynb cell 5 (own RandomState)."""
    rs = np.random.RandomState(seed)
    a = rs.uniform(0.2, 1.0, size=K + n_true)
    Rfull = rs.beta(a, a, size=(M, K + n_true))
    unk = rs.uniform(0, 0.9, size=N)
    Ak = rs.dirichlet(np.ones(K), N).T * (1 - unk)
    Au = rs.dirichlet(np.ones(n_true), N).T * unk
    D = rs.poisson(depth, size=(M, N)) + 1
    cnt = rs.binomial(D, np.clip(Rfull @ np.vstack([Ak, Au]), 0, 1))
    return cnt / D, D.astype(np.int64), np.ascontiguousarray(Rfull[:, :K])


def main():
    dec, boot, ic, init_func = _import_reference()
    X, D, Rk, header = load_fixture_inputs()
    t = os.path.join(REF, "test")
    rd_csv = lambda p, **k: pd.read_csv(os.path.join(t, p), **k).values
    ci_lo, ci_hi = parse_tuple_csv(os.path.join(t, "ci", "confidence_interval_celltypes_proportions.csv"), 0)
    ciu_lo, ciu_hi = parse_tuple_csv(os.path.join(t, "ci", "confidence_interval_methylation_estimate.csv"), None)
    np.savez_compressed(
        os.path.join(OUT, "fixture_shipped.npz"),
        X=X, D=D, Rk=Rk, header=np.array(header),
        ref_based_alpha=rd_csv("output_ref_based/celltypes_proportions.csv", index_col=0),
        partial_alpha=rd_csv("output_partial_ref/celltypes_proportions.csv", index_col=0),
        partial_u=rd_csv("output_partial_ref/methylation_profile_estimate.csv"),
        purity_alpha=rd_csv("purity/celltypes_proportions.csv", index_col=0),
        purity_u=rd_csv("purity/methylation_profile_estimate.csv"),
        purity_pct=np.array([60, 80, 90, 20, 50, 90, 100, 30, 50, 10], dtype=float),
        unsup_alpha=rd_csv("unsupervised/celltypes_proportions.csv", index_col=0),
        unsup_u=rd_csv("unsupervised/methylation_profile_estimate.csv"),
        ic_alpha=rd_csv("model_selection/celltypes_proportions.csv", index_col=0),
        ic_u=rd_csv("model_selection/methylation_profile_estimate.csv"),
        ic_best_n_u=np.array(10),
        ci_alpha_lo=ci_lo, ci_alpha_hi=ci_hi, ci_u_lo=ciu_lo, ci_u_hi=ciu_hi,
    )

    live = {}
    # --- partial reference on the fixture, CLI defaults (10000 x 20, tol 1e-2, seed 1)
    for n_u in (1, 2, 4):
        with CostSpy(dec) as spy:
            u0, R0, a0 = dec.init_BSSMF_md("uniform_", X, D, Rk, n_u, seed=1)
            u, a = dec.mdwbssmf_deconv(u0.copy(), R0.copy(), a0.copy(), X, D, Rk, n_u, n_iter1=10000, n_iter2=20, tol=1e-2)
        live[f"pr{n_u}_u0"], live[f"pr{n_u}_a0"] = u0, np.ascontiguousarray(a0)
        live[f"pr{n_u}_u"], live[f"pr{n_u}_a"] = u, np.ascontiguousarray(a)
        live[f"pr{n_u}_costs"] = np.array(spy.vals)
    # --- list seed (what `--seed 5` produces) and beta / SVD / uniform inits
    for tag, opt, seed in (("listseed", "uniform_", [5]), ("beta", "beta", 3), ("svd", "SVD", 1), ("uniform", "uniform", 2)):
        with CostSpy(dec) as spy:
            u0, R0, a0 = dec.init_BSSMF_md(opt, X, D, Rk, 2, seed=seed)
            u, a = dec.mdwbssmf_deconv(u0.copy(), R0.copy(), a0.copy(), X, D, Rk, 2, n_iter1=10000, n_iter2=20, tol=1e-2)
        live[f"{tag}_u0"], live[f"{tag}_a0"] = u0, np.ascontiguousarray(a0)
        live[f"{tag}_u"], live[f"{tag}_a"] = u, np.ascontiguousarray(a)
        live[f"{tag}_costs"] = np.array(spy.vals)
    # --- purity (CLI semantics: internal purity = 1 - pct/100), python body of the njit'd solver to spy on cost
    pct = np.array([60, 80, 90, 20, 50, 90, 100, 30, 50, 10], dtype=float)
    pur = 1 - pct / 100.0
    u0, R0, a0 = dec.init_BSSMF_md_p("uniform_", X, D, Rk, 1, pur, seed=1)
    u, a = dec.mdwbssmf_deconv_p(u0.copy(), R0.copy(), np.ascontiguousarray(a0), X, D.astype(float), Rk, 1, pur, n_iter1=100, n_iter2=500, tol=1e-2)
    with CostSpy(dec) as spy:
        u_py, a_py = dec.mdwbssmf_deconv_p.py_func(u0.copy(), R0.copy(), np.ascontiguousarray(a0), X, D.astype(float), Rk, 1, pur, n_iter1=100, n_iter2=500, tol=1e-2)
    assert np.abs(u_py - u).max() < 1e-12 and np.abs(a_py - a).max() < 1e-12
    live["pur_u0"], live["pur_a0"], live["pur_u"], live["pur_a"] = u0, np.ascontiguousarray(a0), u, a
    live["pur_costs"], live["pur_purity"] = np.array(spy.vals), pur
    # two-unknown purity on a synthetic problem with short loops
    Xs, Ds, Rs = synth(11, 600, 7, 4, 2)
    purs = np.linspace(0.3, 0.9, 7)
    u0, R0, a0 = dec.init_BSSMF_md_p("uniform_", Xs, Ds, Rs, 2, purs, seed=4)
    with CostSpy(dec) as spy:
        u, a = dec.mdwbssmf_deconv_p.py_func(u0.copy(), R0.copy(), np.ascontiguousarray(a0), Xs, Ds.astype(float), Rs, 2, purs, n_iter1=30, n_iter2=40, tol=1e-3)
    live.update(pur2_X=Xs, pur2_D=Ds, pur2_Rk=Rs, pur2_purity=purs, pur2_u0=u0, pur2_a0=np.ascontiguousarray(a0),
                pur2_u=u, pur2_a=a, pur2_costs=np.array(spy.vals))
    # --- unsupervised on the fixture (CLI defaults, n_u = 4)
    with CostSpy(dec) as spy:
        u, a = dec.unsupervised_deconv(X, 4, D, "uniform_", n_iter1=10000, n_iter2=20, tol=1e-2, seed=1)
    live["unsup_u"], live["unsup_a"], live["unsup_costs"] = u, np.ascontiguousarray(a), np.array(spy.vals)
    # --- reference-based (nbunknown = 0)
    live["refbased_a"] = np.concatenate(
        [init_func.wls_intercept(D[:, k:k + 1] * X[:, k:k + 1], D[:, k:k + 1], Rk) for k in range(X.shape[1])], axis=1)
    # --- synthetic ragged shapes (N not a multiple of anything, K odd), fixed short loops
    for tag, (sd, M, N, K, n_u, it1, it2) in {
        "syn_a": (21, 1000, 3, 3, 1, 15, 10),
        "syn_b": (22, 777, 13, 7, 3, 12, 8),
        "syn_c": (23, 2048, 33, 6, 2, 10, 20),
        "syn_d": (24, 500, 1, 2, 1, 8, 5),
    }.items():
        Xs, Ds, Rs = synth(sd, M, N, K, max(n_u, 1))
        with CostSpy(dec) as spy:
            u0, R0, a0 = dec.init_BSSMF_md("uniform_", Xs, Ds, Rs, n_u, seed=sd)
            u, a = dec.mdwbssmf_deconv(u0.copy(), R0.copy(), a0.copy(), Xs, Ds, Rs, n_u, n_iter1=it1, n_iter2=it2, tol=1e-6)
        live.update({f"{tag}_X": Xs, f"{tag}_D": Ds, f"{tag}_Rk": Rs, f"{tag}_u0": u0, f"{tag}_a0": np.ascontiguousarray(a0),
                     f"{tag}_u": u, f"{tag}_a": np.ascontiguousarray(a), f"{tag}_costs": np.array(spy.vals),
                     f"{tag}_cfg": np.array([n_u, it1, it2])})
    np.savez_compressed(os.path.join(OUT, "live_solver.npz"), **live)

    # --- bootstrap: B = 4 resamples through bt_ci itself (files written to a temp dir and parsed back)
    import tempfile
    drv = {}
    with tempfile.TemporaryDirectory() as td:
        boot.bt_ci(90, 4, 1, X, D, Rk, "uniform_", 10000, 20, 1e-2, list(header), td, [f"s{i}" for i in range(10)], None, 1)
        lo, hi = parse_tuple_csv(os.path.join(td, "confidence_interval_celltypes_proportions.csv"), 0)
        ulo, uhi = parse_tuple_csv(os.path.join(td, "confidence_interval_methylation_estimate.csv"), None)
        drv.update(bt_alpha_lo=lo, bt_alpha_hi=hi, bt_u_lo=ulo, bt_u_hi=uhi)
    with tempfile.TemporaryDirectory() as td:
        boot.bt_ci(80, 3, 0, X, D, Rk, "uniform_", 10000, 20, 1e-2, list(header), td, [f"s{i}" for i in range(10)], None, 1)
        lo, hi = parse_tuple_csv(os.path.join(td, "confidence_interval_celltypes_proportions.csv"), 0)
        drv.update(bt0_alpha_lo=lo, bt0_alpha_hi=hi)
    with tempfile.TemporaryDirectory() as td:
        boot.bt_ci(90, 3, 1, X, D, Rk, "uniform_", 20, 50, 1e-2, list(header), td, [f"s{i}" for i in range(10)], list(pct), 1)
        lo, hi = parse_tuple_csv(os.path.join(td, "confidence_interval_celltypes_proportions.csv"), 0)
        drv.update(btp_alpha_lo=lo, btp_alpha_hi=hi)
    # --- ic sweeps on the fixture (n_u = 1..25 hard-coded in the reference)
    for crit in ("AIC", "BIC"):
        u, a, best, vals = ic.evaluate_best_ic(X, Rk, D, "uniform_", crit, 1, iter1=10000, iter2=20, tol=1e-2)
        drv[f"ic_{crit}_vals"], drv[f"ic_{crit}_best"] = np.array(vals), np.array(best)
        drv[f"ic_{crit}_u"], drv[f"ic_{crit}_a"] = u, np.ascontiguousarray(a)
    # CCC / BCV are 25 x r fits: short loops keep generation fast; r = 3
    for crit in ("CCC", "BCV"):
        u, a, best, vals = ic.evaluate_best_ic(X, Rk, D, "uniform_", crit, 1, iter1=40, iter2=20, tol=1e-2, n_restarts=3)
        drv[f"ic_{crit}_vals"], drv[f"ic_{crit}_best"] = np.array(vals), np.array(best)
        drv[f"ic_{crit}_u"], drv[f"ic_{crit}_a"] = u, np.ascontiguousarray(a)
    np.savez_compressed(os.path.join(OUT, "live_drivers.npz"), **drv)
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
