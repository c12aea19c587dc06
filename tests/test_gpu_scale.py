"""GPU parity at BASELINE sizes (run with -m gpu): the CUDA path against the CPU oracle on the full config-2 shape, a
config-3-sized purity fit, and fp32 mode on a converging fit; plus the kernels of the steps around the path (percentiles,
consensus matrix, NNDSVD split) against numpy."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    import torch
    assert torch.cuda.is_available()
    import __graft_entry__ as g
    g.build()
    import demethify_b200
    demethify_b200.set_engine("auto")
    demethify_b200.set_precision("fp64")
    return demethify_b200


def synth(seed, M, N, K, n_true, depth=50):
    rs = np.random.RandomState(seed)
    a = rs.uniform(0.2, 1.0, size=K + n_true)
    Rf = rs.beta(a, a, size=(M, K + n_true))
    unk = rs.uniform(0, 0.9, size=N)
    Ak = rs.dirichlet(np.ones(max(K, 1)), N).T[:K] * (1 - unk)
    Au = rs.dirichlet(np.ones(n_true), N).T * unk
    D = rs.poisson(depth, size=(M, N)) + 1
    cnt = rs.binomial(D, np.clip(Rf @ np.vstack([Ak, Au]), 0, 1))
    return cnt / D, D.astype(np.int64), np.ascontiguousarray(Rf[:, :K]), unk


@pytest.mark.parametrize("engine", ["auto", "gram"])
def test_config2_full_shape_vs_oracle(pkg, engine):
    """BASELINE config 2: 100k CpG x 16 samples, K = 6, n_u = 2, 20 inner iterations; 4 outer iterations against the oracle."""
    from demethify_b200 import deconvolution as dec
    from demethify_b200 import engine as eng
    from oracle import bssmf_numpy as orc
    X, D, Rk, _ = synth(202, 100_000, 16, 6, 2)
    u0, R0, a0 = orc.draw_init("uniform_", X, D, Rk, 2, seed=1)
    tr = {}
    uo, ao = orc.solve_partial_reference(u0.copy(), R0, a0.copy(), X, D.astype(float), Rk, 2, 4, 20, 1e-9, trace=tr)
    eng.FUSED_SLOTS = engine == "auto"
    try:
        u, a = dec.mdwbssmf_deconv(u0, R0, a0, X, D, Rk, 2, n_iter1=4, n_iter2=20, tol=1e-9)
    finally:
        eng.FUSED_SLOTS = True
    info = dec.last_fit_info()
    assert info["engine"] == ("fused" if engine == "auto" else "gram")
    assert info["n_outer"] == tr["n_outer"] == 4
    assert abs(info["cost"] - tr["costs"][-1]) <= 1e-10 * tr["costs"][-1]
    assert np.abs(a - ao).max() <= 1e-6 and np.abs(u - uo).max() <= 1e-6


def test_config3_purity_vs_oracle(pkg):
    """BASELINE config 3 shape, rows reduced to keep the oracle affordable: 50k CpG x 64 samples, K = 6, n_u = 1, purity known,
    2 outer x 50 inner iterations (update_u + Frank-Wolfe)."""
    from demethify_b200 import deconvolution as dec
    from oracle import bssmf_numpy as orc
    X, D, Rk, unk = synth(303, 50_000, 64, 6, 1)
    pur = 1.0 - unk                                     # internal purity vector: known block sums to purity (deconvolution.py:292-294)
    u0, R0, a0 = dec.init_BSSMF_md_p("uniform_", X, D, Rk, 1, pur, seed=1)
    tr = {}
    uo, ao = orc.solve_purity(u0.copy(), R0, a0.copy(), X, D.astype(float), Rk, 1, pur, 2, 50, 1e-9, trace=tr)
    u, a = dec.mdwbssmf_deconv_p(u0, R0, a0, X, D, Rk, 1, pur, n_iter1=2, n_iter2=50, tol=1e-9)
    assert dec.last_fit_info()["n_outer"] == tr["n_outer"] == 2
    assert np.abs(a - ao).max() <= 1e-6 and np.abs(u - uo).max() <= 1e-6


def test_fp32_converging_fit_same_stopping_iteration(pkg):
    """fp32 storage / arithmetic with fp64 reductions: a fit run to termination must stop at the oracle's outer iteration and
    agree to the north-star bar max |d alpha| <= 1e-4."""
    import demethify_b200
    from demethify_b200 import deconvolution as dec
    from oracle import bssmf_numpy as orc
    X, D, Rk, _ = synth(3, 800, 8, 4, 1)
    u0, R0, a0 = orc.draw_init("uniform_", X, D, Rk, 1, seed=1)
    tr = {}
    uo, ao = orc.solve_partial_reference(u0.copy(), R0, a0.copy(), X, D.astype(float), Rk, 1, 500, 20, 1e-2, trace=tr)
    demethify_b200.set_precision("fp32")
    try:
        u, a = dec.mdwbssmf_deconv(u0, R0, a0, X, D, Rk, 1, n_iter1=500, n_iter2=20, tol=1e-2)
        n32 = dec.last_fit_info()["n_outer"]
    finally:
        demethify_b200.set_precision("fp64")
    assert 1 < tr["n_outer"] < 500, "the oracle fit must converge for this test to mean anything"
    assert n32 == tr["n_outer"]
    assert np.abs(a - ao).max() <= 1e-4 and np.abs(u - uo).max() <= 1e-3


@pytest.mark.parametrize("B,level", [(1, 95), (7, 95), (200, 95), (1000, 95), (1000, 50), (2500, 95)])
def test_percentile_kernel_matches_numpy(pkg, B, level):
    """dmf_percentile_bounds == np.percentile(axis=0) with the default linear rule, bit for bit (bootstrap.py:53-54, :77-78)."""
    import torch
    from demethify_b200.bootstrap import percentile_bounds_device
    rs = np.random.RandomState(B + level)
    stack = rs.uniform(size=(B, 37, 3))
    stack[:, 0, 0] = 0.25                                   # ties
    a = 1 - level / 100
    lo_q, hi_q = 100 * (a / 2), 100 * (1 - a / 2)
    lo, hi = percentile_bounds_device(torch.from_numpy(stack).cuda(), lo_q, hi_q)
    assert np.array_equal(lo, np.percentile(stack, lo_q, axis=0)) and np.array_equal(hi, np.percentile(stack, hi_q, axis=0))


def test_consensus_kernel_matches_numpy(pkg):
    from demethify_b200.ic import compute_consensus_matrix
    rs = np.random.RandomState(5)
    runs = [rs.dirichlet(np.ones(7), 23).T for _ in range(6)]
    runs[2][3, 4] = runs[2][5, 4] = runs[2][:, 4].max() + 1.0         # a tie: np.argmax takes the first
    want = np.zeros((23, 23))
    for al in runs:
        lab = np.argmax(al, axis=0)
        want += (lab[:, None] == lab[None, :])
    assert np.array_equal(compute_consensus_matrix(runs), want / len(runs))


def test_nndsvd_split_matches_formula(pkg):
    """dmf_nndsvd_split against the NNDSVD formulas of init_func.py:46-69 evaluated with numpy on the same SVD factors."""
    import torch
    from demethify_b200.init_func import nndsvd_initialize
    rs = np.random.RandomState(9)
    V = rs.uniform(size=(500, 12))
    W, H = nndsvd_initialize(V, rank=4)
    Ut, St, Vh = torch.linalg.svd(torch.from_numpy(V).cuda(), full_matrices=False)
    U, S, E = Ut.cpu().numpy(), St.cpu().numpy(), Vh.cpu().numpy().T
    Ww, Hw = np.zeros((500, 4)), np.zeros((4, 12))
    Ww[:, 0], Hw[0] = np.sqrt(S[0]) * np.abs(U[:, 0]), np.sqrt(S[0]) * np.abs(E[:, 0])
    for i in range(1, 4):
        up, un, vp, vn = np.maximum(U[:, i], 0), np.maximum(-U[:, i], 0), np.maximum(E[:, i], 0), np.maximum(-E[:, i], 0)
        tp, tn = np.linalg.norm(up) * np.linalg.norm(vp), np.linalg.norm(un) * np.linalg.norm(vn)
        (uu, vv, t) = (up, vp, tp) if tp >= tn else (un, vn, tn)
        Ww[:, i], Hw[i] = np.sqrt(S[i] * t) / np.linalg.norm(uu) * uu, np.sqrt(S[i] * t) / np.linalg.norm(vv) * vv
    Ww[Ww < 1e-11] = 0
    Hw[Hw < 1e-11] = 0
    assert np.abs(W - Ww).max() <= 1e-12 and np.abs(H - Hw).max() <= 1e-12
    assert W.min() >= 0 and H.min() >= 0 and np.abs(W @ H - V).mean() < 0.3


def test_best_of_restarts_and_bootstrap_restarts(pkg):
    """Distinct-seed restarts (extension): the batched best-of-R equals the best of R separate fits, and a
    bootstrap with restarts keeps the lower-cost fit of every resample."""
    from demethify_b200 import deconvolution as dec
    from demethify_b200.bootstrap import bootstrap_fits
    X, D, Rk, _ = synth(41, 1500, 12, 5, 1)
    u, a, best, costs = dec.best_of_restarts(X, D, Rk, 1, "uniform_", 3, 4, 30, 10, 1e-9)
    single = []
    for r in range(4):
        u0, R0, a0 = dec.init_BSSMF_md("uniform_", X, D, Rk, 1, seed=3 + r)
        ur, ar = dec.mdwbssmf_deconv(u0, R0, a0, X, D, Rk, 1, n_iter1=30, n_iter2=10, tol=1e-9)
        single.append((dec.last_fit_info()["cost"], ur, ar))
    assert best == int(np.argmin([s[0] for s in single]))
    assert np.allclose(costs, [s[0] for s in single], rtol=1e-12, atol=0)
    assert np.abs(a - single[best][2]).max() <= 1e-12 and np.abs(u - single[best][1]).max() <= 1e-12
    a1, _, _ = bootstrap_fits(3, 1, X, D, Rk, "uniform_", 20, 10, 1e-9, None, 1, keep_u=False)
    a2, _, _ = bootstrap_fits(3, 1, X, D, Rk, "uniform_", 20, 10, 1e-9, None, 1, keep_u=False, restarts=2)
    assert a1.shape == a2.shape == (3, 6, 12) and np.isfinite(a2).all()
