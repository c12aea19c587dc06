"""GPU edge cases of the solver path: degenerate shapes, error behaviour, engine fallbacks, non-finite inputs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dec():
    import torch
    assert torch.cuda.is_available()
    import __graft_entry__ as g
    g.build()
    from demethify_b200 import deconvolution
    return deconvolution


def synth(seed, M, N, K, n_true, depth=50):
    rs = np.random.RandomState(seed)
    a = rs.uniform(0.2, 1.0, size=K + n_true)
    Rf = rs.beta(a, a, size=(M, K + n_true))
    A = rs.dirichlet(np.ones(K + n_true), N).T
    D = rs.poisson(depth, size=(M, N)) + 1
    X = rs.binomial(D, np.clip(Rf @ A, 0, 1)) / D
    return X, D.astype(np.int64), np.ascontiguousarray(Rf[:, :K])


@pytest.mark.parametrize("M,N,K,n_u", [(1, 1, 1, 1), (2, 1, 3, 2), (17, 1, 6, 1), (5, 3, 6, 4), (70, 513, 2, 1), (64, 300, 6, 2)])
def test_degenerate_shapes_vs_oracle(dec, M, N, K, n_u):
    """One row, one sample, fewer rows than a tile, more samples than threads of a row group (N = 513 -> gram only on C = 4)."""
    from oracle import bssmf_numpy as orc
    X, D, Rk = synth(M * 7 + N, M, N, K, n_u)
    u0, R0, a0 = orc.draw_init("uniform_", X, D, Rk, n_u, seed=2)
    tr = {}
    uo, ao = orc.solve_partial_reference(u0.copy(), R0, a0.copy(), X, D.astype(float), Rk, n_u, 3, 4, 1e-12, trace=tr)
    if N > 512:
        with pytest.raises(Exception):      # the stream geometry bounds N; documented limit of this build
            dec.mdwbssmf_deconv(u0, R0, a0, X, D, Rk, n_u, n_iter1=3, n_iter2=4, tol=1e-12)
        return
    u, a = dec.mdwbssmf_deconv(u0, R0, a0, X, D, Rk, n_u, n_iter1=3, n_iter2=4, tol=1e-12)
    assert dec.last_fit_info()["n_outer"] == tr["n_outer"]
    assert np.abs(a - ao).max() <= 1e-6 and np.abs(u - uo).max() <= 1e-6


def test_zero_inner_iterations_and_zero_outer(dec):
    from oracle import bssmf_numpy as orc
    X, D, Rk = synth(3, 400, 6, 4, 1)
    u0, R0, a0 = orc.draw_init("uniform_", X, D, Rk, 1, seed=1)
    u, a = dec.mdwbssmf_deconv(u0, R0, a0, X, D, Rk, 1, n_iter1=0, n_iter2=20, tol=1e-2)      # returns the initial iterate
    assert np.array_equal(u, u0) and np.array_equal(a, a0) and dec.last_fit_info()["n_outer"] == 0
    u, a = dec.mdwbssmf_deconv(u0, R0, a0, X, D, Rk, 1, n_iter1=5, n_iter2=0, tol=1e-2)       # cost never changes -> stops after 1
    assert np.array_equal(u, u0) and np.array_equal(a, a0) and dec.last_fit_info()["n_outer"] == 1


def test_engine_fallback_for_many_unknowns(dec):
    """n_u > 8 has no Gram-form instantiation: 'auto' falls back to the streaming engine, 'gram' refuses."""
    import demethify_b200
    from demethify_b200 import _lib
    from oracle import bssmf_numpy as orc
    X, D, Rk = synth(5, 600, 12, 3, 9)
    u0, R0, a0 = orc.draw_init("uniform_", X, D, Rk, 9, seed=1)
    uo, ao = orc.solve_partial_reference(u0.copy(), R0, a0.copy(), X, D.astype(float), Rk, 9, 2, 5, 1e-12)
    u, a = dec.mdwbssmf_deconv(u0, R0, a0, X, D, Rk, 9, n_iter1=2, n_iter2=5, tol=1e-12)
    assert dec.last_fit_info()["engine"] == "stream" and np.abs(a - ao).max() <= 1e-6
    demethify_b200.set_engine("gram")
    try:
        with pytest.raises(_lib.DmfError):
            dec.mdwbssmf_deconv(u0, R0, a0, X, D, Rk, 9, n_iter1=2, n_iter2=5, tol=1e-12)
    finally:
        demethify_b200.set_engine("auto")


def test_non_finite_input_is_reported(dec):
    """A NaN in X reaches the simplex projection; the reference dies there with ZeroDivisionError (rho = -1), the library
    reports a DmfError instead of returning garbage."""
    from demethify_b200 import _lib
    from oracle import bssmf_numpy as orc
    X, D, Rk = synth(7, 300, 5, 3, 1)
    u0, R0, a0 = orc.draw_init("uniform_", X, D, Rk, 1, seed=1)
    X = X.copy(); X[10, 2] = np.nan
    with pytest.raises(_lib.DmfError):
        dec.mdwbssmf_deconv(u0, R0, a0, X, D, Rk, 1, n_iter1=3, n_iter2=5, tol=1e-2)


def test_shape_errors(dec):
    X, D, Rk = synth(9, 100, 4, 3, 1)
    with pytest.raises(Exception):
        dec.mdwbssmf_deconv(np.zeros((100, 1)), None, np.ones((4, 4)) / 4, X, D[:50], Rk, 1, n_iter1=1, n_iter2=1)     # d_x shape
    with pytest.raises(Exception):
        dec.mdwbssmf_deconv(np.zeros((100, 1)), None, np.ones((4, 4)) / 4, X, D, Rk[:50], 1, n_iter1=1, n_iter2=1)     # R_trunc rows


def test_all_zero_weights_column(dec):
    """A sample whose coverage is zero everywhere (masked out completely, ic.py:75) must not poison the others."""
    from oracle import bssmf_numpy as orc
    X, D, Rk = synth(13, 500, 6, 4, 1)
    D = D.copy(); D[:, 3] = 0
    u0, R0, a0 = orc.draw_init("uniform_", X, D, Rk, 1, seed=1)
    uo, ao = orc.solve_partial_reference(u0.copy(), R0, a0.copy(), X, D.astype(float), Rk, 1, 4, 10, 1e-12)
    u, a = dec.mdwbssmf_deconv(u0, R0, a0, X, D, Rk, 1, n_iter1=4, n_iter2=10, tol=1e-12)
    assert np.abs(a - ao).max() <= 1e-6 and np.abs(u - uo).max() <= 1e-6
