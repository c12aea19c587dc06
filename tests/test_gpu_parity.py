"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the reference-shaped
Python API / C ABI, against (a) the committed golden vectors of the live reference and (b) the CPU oracle
on seeded inputs.  Bars (BASELINE.json north_star): fp64 mode max|d alpha| <= 1e-6 with identical
outer-iteration count; fp32 mode max|d alpha| <= 1e-4."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL64 = 1e-6
TOL32 = 1e-4


@pytest.fixture(scope="module", params=["stream", "gram", "auto"])
def dec(request):
    """The reference-shaped API with every device engine: 'stream' = one pass per inner iteration, 'gram' = two passes per
    outer iteration (stream where the Gram form has no instantiation), 'auto' = the fused one-pass engine where the library
    supports the shape (FP64, n_u <= 2, K <= 8, N <= 256), else gram, else stream."""
    import torch
    assert torch.cuda.is_available()
    import __graft_entry__ as g
    g.build()
    import demethify_b200
    from demethify_b200 import deconvolution
    demethify_b200.set_engine("auto" if request.param == "gram" else request.param)
    deconvolution.engine_mode = request.param
    if request.param == "gram":                     # 'auto' minus the fused engine: hide the 4-slot layout it needs
        from demethify_b200 import engine as eng
        deconvolution._saved_fused = eng.FUSED_SLOTS
        eng.FUSED_SLOTS = False
    yield deconvolution
    demethify_b200.set_engine("auto")
    if request.param == "gram":
        from demethify_b200 import engine as eng
        eng.FUSED_SLOTS = True


def expected_engine(mode, n_u, K=5, N=10, fp64=True):
    if mode == "stream":
        return "stream"
    if mode == "auto" and fp64 and n_u <= 2 and K <= 8 and N <= 256:
        return "fused"
    return "gram" if (n_u <= 4 or (n_u <= 8 and K <= 6)) else "stream"


def check_engine(dec, n_u, K=5, N=10):
    assert dec.last_fit_info()["engine"] == expected_engine(dec.engine_mode, n_u, K, N)


@pytest.fixture(scope="module")
def orc():
    from oracle import bssmf_numpy
    return bssmf_numpy


def synth(seed, M, N, K, n_true, depth=50):
    rs = np.random.RandomState(seed)
    a = rs.uniform(0.2, 1.0, size=K + n_true)
    Rf = rs.beta(a, a, size=(M, K + n_true))
    unk = rs.uniform(0, 0.9, size=N)
    Ak = rs.dirichlet(np.ones(max(K, 1)), N).T[:K] * (1 - unk)
    Au = rs.dirichlet(np.ones(n_true), N).T * unk
    D = rs.poisson(depth, size=(M, N)) + 1
    cnt = rs.binomial(D, np.clip(Rf @ np.vstack([Ak, Au]), 0, 1))
    return cnt / D, D.astype(np.int64), np.ascontiguousarray(Rf[:, :K])


# ------------------------------------------------------------------------------- golden: shipped + live reference
@pytest.mark.parametrize("n_u", [1, 2, 4])
def test_partial_reference_fixture(dec, shipped, live, n_u):
    X, D, Rk = shipped["X"], shipped["D"], shipped["Rk"]
    u0, R0, a0 = dec.init_BSSMF_md("uniform_", X, D, Rk, n_u, seed=1)
    assert np.array_equal(u0, live[f"pr{n_u}_u0"]) and np.array_equal(a0, live[f"pr{n_u}_a0"])
    u, a = dec.mdwbssmf_deconv(u0, R0, a0, X, D, Rk, n_u, n_iter1=10000, n_iter2=20, tol=1e-2)
    info = dec.last_fit_info()
    check_engine(dec, n_u)
    assert info["n_outer"] == len(live[f"pr{n_u}_costs"]) - 1, "outer-iteration count differs from the reference"
    assert abs(info["cost"] - live[f"pr{n_u}_costs"][-1]) <= 1e-9 * live[f"pr{n_u}_costs"][-1]
    assert np.abs(a - live[f"pr{n_u}_a"]).max() <= TOL64 and np.abs(u - live[f"pr{n_u}_u"]).max() <= TOL64
    if n_u == 1:
        assert np.abs(a - shipped["partial_alpha"]).max() <= TOL64 and np.abs(u - shipped["partial_u"]).max() <= TOL64


def test_purity_fixture(dec, shipped, live):
    X, D, Rk = shipped["X"], shipped["D"], shipped["Rk"]
    pur = 1 - shipped["purity_pct"] / 100.0
    u0, R0, a0 = dec.init_BSSMF_md_p("uniform_", X, D, Rk, 1, pur, seed=1)
    u, a = dec.mdwbssmf_deconv_p(u0, R0, a0, X, D, Rk, 1, pur, n_iter1=100, n_iter2=500, tol=1e-2)
    check_engine(dec, 1)
    assert dec.last_fit_info()["n_outer"] == len(live["pur_costs"]) - 1
    assert np.abs(a - shipped["purity_alpha"]).max() <= TOL64 and np.abs(u - shipped["purity_u"]).max() <= TOL64


def test_purity_two_unknowns(dec, live):
    X, D, Rk, pur = live["pur2_X"], live["pur2_D"], live["pur2_Rk"], live["pur2_purity"]
    u, a = dec.mdwbssmf_deconv_p(live["pur2_u0"], None, live["pur2_a0"], X, D, Rk, 2, pur, n_iter1=30, n_iter2=40, tol=1e-3)
    assert dec.last_fit_info()["n_outer"] == len(live["pur2_costs"]) - 1
    assert np.abs(a - live["pur2_a"]).max() <= TOL64 and np.abs(u - live["pur2_u"]).max() <= TOL64


def test_unsupervised_fixture(dec, shipped, live):
    u, a = dec.unsupervised_deconv(shipped["X"], 4, shipped["D"], "uniform_", n_iter1=10000, n_iter2=20, tol=1e-2, seed=1)
    check_engine(dec, 4, 0)
    assert dec.last_fit_info()["n_outer"] == len(live["unsup_costs"]) - 1
    assert np.abs(a - shipped["unsup_alpha"]).max() <= TOL64 and np.abs(u - shipped["unsup_u"]).max() <= TOL64


@pytest.mark.parametrize("tag", ["syn_a", "syn_b", "syn_c", "syn_d"])
def test_ragged_shapes_live(dec, live, tag):
    X, D, Rk = live[f"{tag}_X"], live[f"{tag}_D"], live[f"{tag}_Rk"]
    n_u, it1, it2 = (int(v) for v in live[f"{tag}_cfg"])
    u, a = dec.mdwbssmf_deconv(live[f"{tag}_u0"], None, live[f"{tag}_a0"], X, D, Rk, n_u, n_iter1=it1, n_iter2=it2, tol=1e-6)
    assert dec.last_fit_info()["n_outer"] == len(live[f"{tag}_costs"]) - 1
    assert np.abs(a - live[f"{tag}_a"]).max() <= TOL64 and np.abs(u - live[f"{tag}_u"]).max() <= TOL64


# ------------------------------------------------------------------------------- oracle on seeded inputs
@pytest.mark.parametrize("M,N,K,n_u,it1,it2", [
    (5000, 16, 6, 2, 5, 20),       # BASELINE config 2 shape, fewer rows
    (4096, 64, 6, 1, 4, 10),       # config 3/4 sample count
    (3000, 256, 6, 2, 3, 5),       # config 5 sample count (one row spans several warps)
    (2500, 40, 12, 3, 3, 6),       # Kt = 15 -> 16-wide register tile
    (1200, 24, 20, 9, 2, 4),       # Kt = 29 -> 32-wide register tile
    (2000, 256, 6, 7, 3, 6),       # 7 unknown types: Gram engine with one register row per batch, 16-wide alpha panels
    (1500, 20, 4, 5, 3, 5),        # 5 unknown types, short rows
    (1800, 64, 0 + 6, 8, 2, 4),    # 8 unknown types
    (999, 7, 1, 1, 4, 7),          # odd everything: rows not 16-byte aligned
    (33, 2, 2, 1, 3, 3),           # fewer rows than one tile
])
def test_partial_reference_vs_oracle(dec, orc, M, N, K, n_u, it1, it2):
    X, D, Rk = synth(M + N, M, N, K, max(n_u, 1))
    u0, R0, a0 = orc.draw_init("uniform_", X, D, Rk, n_u, seed=7)
    tr = {}
    uo, ao = orc.solve_partial_reference(u0.copy(), R0, a0.copy(), X, D.astype(float), Rk, n_u, it1, it2, 1e-9, trace=tr)
    u, a = dec.mdwbssmf_deconv(u0, R0, a0, X, D, Rk, n_u, n_iter1=it1, n_iter2=it2, tol=1e-9)
    info = dec.last_fit_info()
    check_engine(dec, n_u, K, N)
    assert info["n_outer"] == tr["n_outer"]
    assert abs(info["cost"] - tr["costs"][-1]) <= 1e-9 * tr["costs"][-1]
    assert np.abs(a - ao).max() <= TOL64 and np.abs(u - uo).max() <= TOL64


def test_float_weights_and_cost(dec, orc):
    """Non-integer weights take the float-weight kernels; cost_f_w standalone."""
    X, D, Rk = synth(5, 2000, 10, 4, 2)
    Dw = D * 0.37
    u0, R0, a0 = orc.draw_init("uniform_", X, Dw, Rk, 2, seed=3)
    assert abs(dec.cost_f_w(X, R0, a0, Dw) - orc.weighted_cost(X, R0, a0, Dw)) <= 1e-10 * orc.weighted_cost(X, R0, a0, Dw)
    uo, ao = orc.solve_partial_reference(u0.copy(), R0, a0.copy(), X, Dw, Rk, 2, 3, 8, 1e-9)
    u, a = dec.mdwbssmf_deconv(u0, R0, a0, X, Dw, Rk, 2, n_iter1=3, n_iter2=8, tol=1e-9)
    assert np.abs(a - ao).max() <= TOL64 and np.abs(u - uo).max() <= TOL64


def test_zero_weights_masked_entries(dec, orc):
    """BCV passes X*mask, d*mask (ic.py:75): zero weights must be handled exactly."""
    X, D, Rk = synth(9, 1500, 9, 5, 1)
    mask = np.random.RandomState(1).rand(*X.shape) < 0.3
    Xm, Dm = X * mask, D * mask
    u0, R0, a0 = orc.draw_init("uniform_", Xm, Dm, Rk, 1, seed=2)
    uo, ao = orc.solve_partial_reference(u0.copy(), R0, a0.copy(), Xm, Dm.astype(float), Rk, 1, 4, 10, 1e-9)
    u, a = dec.mdwbssmf_deconv(u0, R0, a0, Xm, Dm, Rk, 1, n_iter1=4, n_iter2=10, tol=1e-9)
    assert np.abs(a - ao).max() <= TOL64 and np.abs(u - uo).max() <= TOL64


def test_fp32_mode(dec, orc):
    import demethify_b200
    X, D, Rk = synth(11, 6000, 32, 6, 2)
    u0, R0, a0 = orc.draw_init("uniform_", X, D, Rk, 2, seed=5)
    uo, ao = orc.solve_partial_reference(u0.copy(), R0, a0.copy(), X, D.astype(float), Rk, 2, 5, 20, 0.0)
    demethify_b200.set_precision("fp32")
    try:
        u, a = dec.mdwbssmf_deconv(u0, R0, a0, X, D, Rk, 2, n_iter1=5, n_iter2=20, tol=0.0)
    finally:
        demethify_b200.set_precision("fp64")
    assert dec.last_fit_info()["n_outer"] == 5
    assert np.abs(a - ao).max() <= TOL32 and np.abs(u - uo).max() <= 1e-3


def test_termination_and_determinism(dec, orc):
    """Converging run: same stopping iteration as the oracle, and bit-identical results run to run."""
    X, D, Rk = synth(3, 800, 8, 4, 1)
    u0, R0, a0 = orc.draw_init("uniform_", X, D, Rk, 1, seed=1)
    tr = {}
    orc.solve_partial_reference(u0.copy(), R0, a0.copy(), X, D.astype(float), Rk, 1, 500, 20, 1e-2, trace=tr)
    u1, a1 = dec.mdwbssmf_deconv(u0, R0, a0, X, D, Rk, 1, n_iter1=500, n_iter2=20, tol=1e-2)
    n1 = dec.last_fit_info()["n_outer"]
    u2, a2 = dec.mdwbssmf_deconv(u0, R0, a0, X, D, Rk, 1, n_iter1=500, n_iter2=20, tol=1e-2)
    assert n1 == tr["n_outer"] == dec.last_fit_info()["n_outer"]
    assert np.array_equal(u1, u2) and np.array_equal(a1, a2)


def test_size_independent_properties_full_scale(dec):
    """At a BASELINE-sized shape (100k x 16, config 2) the oracle is too slow for the default suite; check
    invariants instead: alpha columns on the simplex, u in [0,1], cost non-increasing over outer iterations,
    and row-permutation equivariance (permuting CpG rows permutes u and leaves alpha unchanged to rounding)."""
    import torch
    from demethify_b200 import _lib
    from demethify_b200.engine import DeviceProblem, FitBatch
    X, D, Rk = synth(101, 100_000, 16, 6, 2)
    rs = np.random.RandomState(4)
    u0 = rs.uniform(size=(X.shape[0], 2)); a0 = rs.dirichlet(np.ones(8), 16).T
    prob = DeviceProblem(X, D, Rk)
    b = FitBatch(prob, 2, [u0], [a0], trace_cap=16)
    assert b.engine == expected_engine(dec.engine_mode, 2, 6, 16)
    st = b.fit(8, 20, 0.0)
    (u, a, n_outer, cost), = b.results(st)
    tr = b.trace[0, :9].cpu().numpy()
    assert n_outer == 8 and np.all(np.diff(tr) <= 1e-9 * tr[0])
    assert np.allclose(a.sum(0), 1.0, atol=1e-12) and a.min() >= 0 and u.min() >= 0 and u.max() <= 1
    perm = rs.permutation(X.shape[0])
    prob2 = DeviceProblem(X[perm], D[perm], Rk[perm])
    b2 = FitBatch(prob2, 2, [u0[perm]], [a0])
    (u2, a2, _, cost2), = b2.results(b2.fit(8, 20, 0.0))
    assert np.abs(a2 - a).max() <= 1e-9 and np.abs(u2 - u[perm]).max() <= 1e-9 and abs(cost2 - cost) <= 1e-9 * cost
