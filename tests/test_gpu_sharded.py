"""GPU tests of the row-sharded path: GpuShardBackend puts the library into 'publish partial sums / finalize on reduced sums' mode.
With one process the all-reduce is the identity, so the results must equal the oracle like the unsharded engine; the
multi-process behaviour of the same orchestration is covered on CPU (tests/test_sharded_cpu.py, gloo) and on 2 GPUs by
tools/sharded_check.py (torchrun)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sharded():
    import torch
    assert torch.cuda.is_available()
    import __graft_entry__ as g
    g.build()
    from demethify_b200 import sharded
    return sharded


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


@pytest.mark.parametrize("M,N,K,n_u,it1,it2,tol", [(5000, 16, 6, 2, 5, 20, 1e-9), (3000, 256, 6, 2, 3, 5, 1e-9), (800, 8, 4, 1, 500, 20, 1e-2),
                                                   (2500, 40, 12, 3, 3, 6, 1e-9)])
def test_sharded_mode_single_rank_vs_oracle(sharded, M, N, K, n_u, it1, it2, tol):
    from oracle import bssmf_numpy as orc
    X, D, Rk = synth(M + N, M, N, K, max(n_u, 1))
    u0, R0, a0 = orc.draw_init("uniform_", X, D, Rk, n_u, seed=7)
    tr = {}
    uo, ao = orc.solve_partial_reference(u0.copy(), R0, a0.copy(), X, D.astype(float), Rk, n_u, it1, it2, tol, trace=tr)
    u, a, n_outer, cost = sharded.mdwbssmf_deconv_sharded(u0, a0, X, D, Rk, n_u, n_iter1=it1, n_iter2=it2, tol=tol)
    assert n_outer == tr["n_outer"] and abs(cost - tr["costs"][-1]) <= 1e-9 * cost
    assert np.abs(a - ao).max() <= 1e-6 and np.abs(u - uo).max() <= 1e-6


def test_sharded_mode_purity(sharded, live):
    X, D, Rk, pur = live["pur2_X"], live["pur2_D"], live["pur2_Rk"], live["pur2_purity"]
    u, a, n_outer, _ = sharded.mdwbssmf_deconv_sharded(live["pur2_u0"], live["pur2_a0"], X, D, Rk, 2, n_iter1=30, n_iter2=40, tol=1e-3, purity=pur)
    assert n_outer == len(live["pur2_costs"]) - 1
    assert np.abs(a - live["pur2_a"]).max() <= 1e-6 and np.abs(u - live["pur2_u"]).max() <= 1e-6
