"""Device MT19937 streams against numpy's legacy RandomState, bit for bit, through the C ABI."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _streams(seeds, M, n_dbl, want_idx=True, want_u=True):
    import torch
    from demethify_b200 import _lib
    from demethify_b200.engine import _stream_ptr
    lib = _lib.lib()
    dev = torch.device("cuda:0")
    B = len(seeds)
    sd = torch.from_numpy(np.array(seeds, dtype=np.uint32).view(np.int32)).to(dev)
    idx = torch.full((B, max(M, 1)), -1, dtype=torch.int32, device=dev) if want_idx else None
    u = torch.full((B, max(n_dbl, 1)), -1.0, dtype=torch.float64, device=dev) if want_u else None
    st = torch.zeros((B, 625), dtype=torch.int32, device=dev)
    _lib.check(lib.dmf_rng_legacy_streams(C.c_void_p(sd.data_ptr()), B, M, C.c_void_p(idx.data_ptr()) if want_idx else None, max(M, 1), n_dbl,
                                          C.c_void_p(u.data_ptr()) if want_u else None, max(n_dbl, 1), C.c_void_p(st.data_ptr()), _stream_ptr()))
    torch.cuda.synchronize()
    return (idx.cpu().numpy() if want_idx else None, u.cpu().numpy() if want_u else None, st.cpu().numpy().view(np.uint32))


@pytest.mark.parametrize("M", [1, 2, 3, 350, 1000, 4096, 65537, 500_000])
def test_randint_matches_numpy(M):
    seeds = [0, 1, 2, 4, 7, 11, 12345, 2 ** 31, 2 ** 32 - 1]
    idx, _, _ = _streams(seeds, M, 0, want_u=False)
    for b, s in enumerate(seeds):
        ref = np.random.RandomState(s).randint(0, M, size=(M,))        # what sklearn.utils.resample draws (bootstrap.py:28)
        assert np.array_equal(idx[b, :M].astype(np.int64), ref), (M, s)


@pytest.mark.parametrize("n_dbl", [1, 2, 311, 312, 313, 700, 624 * 5, 1_000_001])
def test_uniform_and_state_match_numpy(n_dbl):
    seeds = [1, 2, 4, 7, 99, 2 ** 32 - 1]
    _, u, st = _streams(seeds, 0, n_dbl, want_idx=False)
    for b, s in enumerate(seeds):
        rs = np.random.RandomState(s)
        ref = rs.uniform(size=n_dbl)
        assert np.array_equal(u[b, :n_dbl], ref), (n_dbl, s)
        # the state handed back continues the stream exactly where numpy is (the dirichlet draw of alpha_0 follows, deconvolution.py:56)
        cont = np.random.RandomState(0)
        cont.set_state(("MT19937", st[b, :624], int(st[b, 624]), 0, 0.0))
        assert np.array_equal(cont.dirichlet(np.ones(7), 5), rs.dirichlet(np.ones(7), 5))


def test_sklearn_resample_indices():
    from sklearn.utils import resample
    M = 5000
    base = np.arange(M)
    seeds = [3, 5, 8]
    idx, _, _ = _streams(seeds, M, 0, want_u=False)
    for b, s in enumerate(seeds):
        assert np.array_equal(idx[b], resample(base, random_state=s))


def test_bootstrap_device_draws_equal_host_draws(monkeypatch):
    """bootstrap_fits with the device streams returns exactly (bit for bit) what it returns with numpy's host streams."""
    from demethify_b200 import bootstrap as bt
    rs = np.random.RandomState(5)
    M, N, K, n_u = 6000, 12, 5, 1
    Rf = rs.beta(0.5, 0.5, size=(M, K + n_u))
    A = rs.dirichlet(np.ones(K + n_u), N).T
    D = rs.poisson(40, size=(M, N)) + 1
    X = rs.binomial(D, np.clip(Rf @ A, 0, 1)) / D
    Rk = np.ascontiguousarray(Rf[:, :K])
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("DMF_HOST_RNG", mode)
        a, u, n = bt.bootstrap_fits(5, n_u, X, D, Rk, "uniform_", 8, 20, 1e-2, None, 3)
        out[mode] = (np.asarray(a), np.asarray(u), n)
    assert out["0"][2] == out["1"][2]
    assert np.array_equal(out["0"][0], out["1"][0]) and np.array_equal(out["0"][1], out["1"][1])


@pytest.mark.parametrize("engine", ["fused", "gram", "stream"])
def test_fits_are_bit_reproducible_run_to_run(engine):
    """Deterministic reductions everywhere (no floating-point atomics, single writer per statistic): the same fit twice gives the
    same bits, whatever addresses the allocator hands out."""
    import torch
    import demethify_b200
    from demethify_b200 import deconvolution as dec
    rs = np.random.RandomState(11)
    M, N, K, n_u = 5000, 12, 5, 2
    Rf = rs.beta(0.5, 0.5, size=(M, K + n_u))
    A = rs.dirichlet(np.ones(K + n_u), N).T
    D = rs.poisson(40, size=(M, N)) + 1
    X = rs.binomial(D, np.clip(Rf @ A, 0, 1)) / D
    Rk = np.ascontiguousarray(Rf[:, :K])
    u0, R0, a0 = dec.init_BSSMF_md("uniform_", X, D, Rk, n_u, seed=1)
    demethify_b200.set_engine(engine)
    try:
        outs = []
        for t in range(3):
            u, a = dec.mdwbssmf_deconv(u0, R0, a0, X, D, Rk, n_u, n_iter1=6, n_iter2=20, tol=1e-3)
            outs.append((u.copy(), a.copy()))
            junk = torch.randn(1 << (20 + t), device="cuda")      # move the next run's allocations
            del junk
    finally:
        demethify_b200.set_engine("auto")
    for u, a in outs[1:]:
        assert np.array_equal(u, outs[0][0]) and np.array_equal(a, outs[0][1])
