"""GPU parity of the fused pass across its width classes (run with -m gpu).  dmf_fused.cuh picks the tile geometry from N
(N <= 64 / 128 / 256: 32-, 32- and 16-row tiles, different warp maps, ring depths and U-warp counts) and from n_u; every class is
checked against the CPU oracle with row counts that leave a short last tile, fewer rows than one tile, sample counts that leave
empty sample groups, and as a multi-fit batch (the layout of a bootstrap wave in materialised form)."""
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


def synth(seed, M, N, K, n_true, depth=40):
    rs = np.random.RandomState(seed)
    a = rs.uniform(0.2, 1.0, size=K + n_true)
    Rf = rs.beta(a, a, size=(M, K + n_true))
    A = rs.dirichlet(np.ones(K + n_true), N).T
    D = rs.poisson(depth, size=(M, N)) + 1
    X = rs.binomial(D, np.clip(Rf @ A, 0, 1)) / D
    return X, D.astype(np.int64), np.ascontiguousarray(Rf[:, :K])


# (M, N, K, n_u): every (width class, n_u) pair, short / single / partial tiles, odd K, samples that leave sample groups empty
SHAPES = [(4133, 64, 6, 1), (4133, 64, 6, 2), (2999, 33, 5, 1), (3001, 17, 6, 2), (19, 8, 3, 1), (31, 64, 6, 2), (33, 40, 2, 2),
          (3500, 128, 6, 1), (3500, 128, 6, 2), (2050, 65, 7, 2), (1111, 100, 8, 1),
          (2001, 256, 6, 2), (2001, 129, 6, 1), (1500, 200, 4, 2)]


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_width_class_vs_oracle(pkg, shape):
    """One fit per shape on the fused engine: same outer-iteration count as the oracle, |d alpha|, |d u| <= 1e-6 (fp64 bar)."""
    from demethify_b200 import deconvolution as dec
    from oracle import bssmf_numpy as orc
    M, N, K, n_u = shape
    X, D, Rk = synth(M + 7 * N + n_u, M, N, K, n_u)
    u0, R0, a0 = orc.draw_init("uniform_", X, D, Rk, n_u, seed=11)
    tr = {}
    uo, ao = orc.solve_partial_reference(u0.copy(), R0, a0.copy(), X, D.astype(float), Rk, n_u, 5, 20, 1e-9, trace=tr)
    u, a = dec.mdwbssmf_deconv(u0, R0, a0, X, D, Rk, n_u, n_iter1=5, n_iter2=20, tol=1e-9)
    info = dec.last_fit_info()
    assert info["engine"] == "fused"
    assert info["n_outer"] == tr["n_outer"]
    assert abs(info["cost"] - tr["costs"][-1]) <= 1e-10 * tr["costs"][-1]
    assert np.abs(a - ao).max() <= 1e-6 and np.abs(u - uo).max() <= 1e-6


@pytest.mark.parametrize("n_u", [1, 2])
def test_narrow_batch_matches_single_fits(pkg, n_u):
    """A batch of fits on their own gathered copies (a bootstrap wave in materialised form, 32 CTAs per fit) returns, fit by fit,
    what the same fit returns alone (148 CTAs; only the partition of the rows over CTAs, i.e. the order of the sums, differs)."""
    import torch
    from demethify_b200.engine import DeviceProblem, FitBatch
    from oracle import bssmf_numpy as orc
    M, N, K, B = 6007, 48, 6, 5
    X, D, Rk = synth(77 + n_u, M, N, K, n_u)
    prob = DeviceProblem(X, D, Rk)
    rs = np.random.RandomState(5)
    idx = torch.from_numpy(rs.randint(0, M, size=(B, M)).astype(np.int32)).to(prob.device)
    probs = prob.gathered_many(idx)
    U0 = rs.uniform(size=(B, M, n_u))
    A0 = np.stack([rs.dirichlet(np.ones(K + n_u), N).T for _ in range(B)])
    batch = FitBatch(probs, n_u, torch.from_numpy(U0).to(prob.device), A0)
    assert batch.engine == "fused"
    states = batch.fit(40, 20, 1e-3)
    res = batch.results(states)
    batch.close()
    ih = idx.cpu().numpy()
    for b in range(B):
        single = FitBatch(DeviceProblem(X[ih[b]], D[ih[b]], Rk[ih[b]]), n_u, [U0[b]], [A0[b]])
        st = single.fit(40, 20, 1e-3)
        u1, a1, n1, c1 = single.results(st)[0]
        single.close()
        ub, ab, nb, cb = res[b]
        assert nb == n1
        assert np.abs(ab - a1).max() <= 1e-9 and np.abs(ub - u1).max() <= 1e-9
    # and one of them against the oracle
    tr = {}
    R0 = np.hstack([Rk[ih[0]], U0[0]])
    uo, ao = orc.solve_partial_reference(U0[0].copy(), R0, A0[0].copy(), X[ih[0]], D[ih[0]].astype(float), Rk[ih[0]], n_u, 40, 20, 1e-3, trace=tr)
    assert res[0][2] == tr["n_outer"]
    assert np.abs(res[0][1] - ao).max() <= 1e-6 and np.abs(res[0][0] - uo).max() <= 1e-6


@pytest.mark.parametrize("shape", [(1500, 24, 20, 2), (1200, 40, 17, 4), (900, 33, 25, 1)], ids=lambda s: "x".join(map(str, s)))
def test_wide_alpha_kernel_vs_oracle(pkg, shape):
    """17 .. 32 cell types in total: the Gram-form engine with the 32-lanes-per-sample alpha kernel (a whole warp per sample)."""
    from demethify_b200 import deconvolution as dec
    from oracle import bssmf_numpy as orc
    M, N, K, n_u = shape
    X, D, Rk = synth(3 * M + N + K, M, N, K, n_u)
    u0, R0, a0 = orc.draw_init("uniform_", X, D, Rk, n_u, seed=5)
    tr = {}
    uo, ao = orc.solve_partial_reference(u0.copy(), R0, a0.copy(), X, D.astype(float), Rk, n_u, 4, 12, 1e-9, trace=tr)
    u, a = dec.mdwbssmf_deconv(u0, R0, a0, X, D, Rk, n_u, n_iter1=4, n_iter2=12, tol=1e-9)
    info = dec.last_fit_info()
    assert info["engine"] == "gram"
    assert info["n_outer"] == tr["n_outer"]
    assert np.abs(a - ao).max() <= 1e-6 and np.abs(u - uo).max() <= 1e-6


def test_wide_purity_alpha_kernel_vs_oracle(pkg):
    """Purity (Frank-Wolfe) alpha step with 18 known types: the first-argmin reductions of the lane-parallel kernel over 32 lanes."""
    from demethify_b200 import deconvolution as dec
    from oracle import bssmf_numpy as orc
    M, N, K, n_u = 1100, 20, 18, 1
    X, D, Rk = synth(4242, M, N, K, n_u)
    pur = np.random.RandomState(3).uniform(0.2, 0.9, size=N)
    u0, R0, a0 = orc.draw_init_purity("uniform_", X, D, Rk, n_u, pur, seed=2)
    tr = {}
    uo, ao = orc.solve_purity(u0.copy(), R0, a0.copy(), X, D.astype(float), Rk, n_u, pur, 3, 25, 1e-9, trace=tr)
    u, a = dec.mdwbssmf_deconv_p(u0, R0, a0, X, D, Rk, n_u, pur, n_iter1=3, n_iter2=25, tol=1e-9)
    info = dec.last_fit_info()
    assert info["n_outer"] == tr["n_outer"]
    assert np.abs(a - ao).max() <= 1e-6 and np.abs(u - uo).max() <= 1e-6
