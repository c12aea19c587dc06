"""Drop-in mirror of the reference's demethify/deconvolution.py on top of the sm_100a kernel library.

Same function names, argument order, defaults and numpy-in / numpy-out contract as the reference
(file:line cited per function); the arithmetic of every solver step runs in libdemethify_sm100.so.
Initial draws use numpy's legacy MT19937 stream exactly as the reference does (set_seed -> numpy.random),
so that u0 / alpha0 are bit-identical for a given seed.
"""
import numpy as np
import numpy.random as rd

from . import _lib
from .engine import DeviceProblem, FitBatch
from .init_func import wls_intercept, wls_all_samples, constrained_nndsvd, nndsvd_initialize

__all__ = ["set_seed", "cost_f_w", "projection_simplex_sort_2d", "init_BSSMF_md", "init_BSSMF_md_p",
           "mdwbssmf_deconv", "mdwbssmf_deconv_p", "unsupervised_deconv", "last_fit_info", "best_of_restarts"]

_last = {}


def last_fit_info():
    """{'n_outer', 'cost', 'launches'} of the most recent solver call (the reference exposes no such hook;
    needed for the identical-iteration-count parity check)."""
    return dict(_last)


def set_seed(seed=None):
    """deconvolution.py:9-11 — seeds numpy's global legacy stream (int -> init_genrand, list -> init_by_array)."""
    if seed is not None:
        rd.seed(seed)


def cost_f_w(y, R, alpha, d_x):
    """deconvolution.py:15-17 — sum d_x * (y - R @ alpha)^2, one streaming pass on the GPU."""
    R = np.asarray(R)
    alpha = np.asarray(alpha)
    Kt = R.shape[1]
    # the kernels take [R_trunc | u]; any split is equivalent for the cost.  The library pads both blocks to even widths and needs
    # their sum <= 32: the last TWO columns are "u" when Kt is even, the last one when it is odd (works up to Kt = 32)
    nu = 2 if (Kt % 2 == 0 and Kt >= 2) else 1
    prob = DeviceProblem(y, d_x, R[:, :Kt - nu] if Kt > nu else None)
    batch = FitBatch(prob, nu, [R[:, Kt - nu:]], [alpha], mode=_lib.DMF_MODE_PARTIAL if Kt > nu else _lib.DMF_MODE_UNSUPERVISED,
                     engine="stream")
    batch.pass_init()
    cost = batch.states()[0].cost
    batch.close()
    return cost


def projection_simplex_sort_2d(v, z=1):
    """deconvolution.py:21-37 — column-wise Euclidean projection onto the simplex.  On the solver path this
    runs inside the alpha-step kernel; this standalone form (used only by the SVD init on a Kt x N matrix)
    is host-side glue."""
    v = np.asarray(v, dtype=np.float64)
    p, n = v.shape
    srt = -np.sort(-v, axis=0)
    pi = np.cumsum(srt, axis=0) - z
    ok = (srt - pi / np.arange(1, p + 1)[:, None]) > 0
    if not ok.any(axis=0).all():
        raise ZeroDivisionError("division by zero")
    rho = p - 1 - np.argmax(ok[::-1], axis=0)
    theta = pi[rho, np.arange(n)] / (rho + 1)
    return np.maximum(v - theta, 0)


def _draw(init_option, meth_frequency, d_x, R_trunc, n_u):
    M, K = R_trunc.shape
    nb = meth_frequency.shape[1]
    if init_option == "uniform":
        u = rd.uniform(size=(M, n_u))
        alpha = wls_all_samples(meth_frequency, d_x, R_trunc, extra=u)      # per-sample wls_intercept on [R_trunc | u], :49-52
    elif init_option == "uniform_":
        u = rd.uniform(size=(M, n_u))
        alpha = rd.dirichlet(np.ones(K + n_u), nb).T
    elif init_option == "beta":
        temp = np.ones((M, n_u))
        u = rd.beta(temp * 0.5, temp * 0.5)
        alpha = rd.dirichlet(np.ones(K + n_u), nb).T
    elif init_option == "SVD":
        W, alpha = constrained_nndsvd(meth_frequency, R_trunc, d_x, rank=n_u, flag=0)
        u = W[:, K:]
    elif init_option == "ICA":
        raise NotImplementedError("--init ICA forms an M x M covariance (init_func.py:120) and is out of scope of the "
                                  "B200 path (SURVEY.md 2.1 row 4); use uniform_, uniform, beta or SVD")
    else:
        raise ValueError(f"unknown init option {init_option!r}")
    return u, alpha


def init_BSSMF_md(init_option, meth_frequency, d_x, R_trunc, n_u, seed=None, rb_alg=wls_intercept):
    """deconvolution.py:40-78."""
    set_seed(seed)
    nb = meth_frequency.shape[1]
    if init_option != "uniform_" and n_u > nb:
        init_option = "uniform_"
    u, alpha = _draw(init_option, meth_frequency, d_x, R_trunc, n_u)
    if init_option == "SVD":
        alpha = projection_simplex_sort_2d(alpha)
    R = np.c_[R_trunc, u]
    if alpha[-n_u:][0].all() == 0.0:                      # :74-76
        alpha[-n_u:][0] = 1e-10
        alpha[:-n_u] = (1 - 1e-10) * alpha[:-n_u]
    return u, R, alpha


def init_BSSMF_md_p(init_option, meth_frequency, d_x, R_trunc, n_u, purity, rb_alg=wls_intercept, seed=None):
    """deconvolution.py:228-267."""
    set_seed(seed)
    nb = meth_frequency.shape[1]
    if init_option != "uniform" and n_u > nb:
        print("The number of unknowns is greater than the number of samples, we'll go with a uniform initialisation. ")
        init_option = "uniform"
    if init_option != "uniform_" and n_u > nb:
        init_option = "uniform_"
    u, alpha = _draw(init_option, meth_frequency, d_x, R_trunc, n_u)
    if init_option == "SVD":                               # :262
        alpha = np.vstack((purity * projection_simplex_sort_2d(alpha[:-n_u]), projection_simplex_sort_2d(alpha[-n_u:])))
    R = np.c_[R_trunc, u]
    return u, R, alpha


def _solve(mode, u, alpha, meth_frequency, d_x, R_trunc, n_u, n_iter1, n_iter2, tol, purity=None):
    # `meth_frequency` may be a DeviceProblem that is already resident in HBM (the n_u sweep of ic.py fits the same data many times)
    prob = meth_frequency if isinstance(meth_frequency, DeviceProblem) else DeviceProblem(meth_frequency, d_x, R_trunc)
    batch = FitBatch(prob, n_u, [np.asarray(u).reshape(-1, n_u)], [np.asarray(alpha)], mode=mode, purity=purity)
    states = batch.fit(n_iter1, n_iter2, tol)
    (u_out, a_out, n_outer, cost), = batch.results(states)
    _last.update(n_outer=n_outer, cost=cost, launches=batch.launch_count(), geometry=batch.geometry(), engine=batch.engine)
    batch.close()
    return u_out, a_out


def mdwbssmf_deconv(u, R, alpha, meth_frequency, d_x, R_trunc, n_u, n_iter1=100000, n_iter2=50, tol=1e-3):
    """deconvolution.py:190-223 — partial-reference accelerated projected gradient; returns (u, alpha)."""
    return _solve(_lib.DMF_MODE_PARTIAL, u, alpha, meth_frequency, d_x, R_trunc, n_u, n_iter1, n_iter2, tol)


def mdwbssmf_deconv_p(u, R, alpha, meth_frequency, d_x, R_trunc, n_u, purity, n_iter1=100, n_iter2=500, tol=1e-3):
    """deconvolution.py:305-337 — purity-constrained variant (U step + Frank-Wolfe alpha step)."""
    return _solve(_lib.DMF_MODE_PURITY, u, alpha, meth_frequency, d_x, R_trunc, n_u, n_iter1, n_iter2, tol, purity=purity)


def unsupervised_deconv(meth_frequency, n_u, d_x, init_option, n_iter1=100000, n_iter2=20, tol=1e-3, seed=None):
    """deconvolution.py:107-184 — reference-free variant (no R_trunc)."""
    set_seed(seed)
    M, nb = meth_frequency.shape
    if init_option != "uniform_" and n_u > nb:
        init_option = "uniform_"
    if init_option == "uniform_":
        u = rd.uniform(size=(M, n_u))
        alpha = rd.dirichlet(np.ones(n_u), nb).T
    elif init_option == "beta":
        temp = np.ones((M, n_u))
        u = rd.beta(temp * 0.5, temp * 0.5)
        alpha = rd.dirichlet(np.ones(n_u), nb).T
    elif init_option == "SVD":
        u, alpha = nndsvd_initialize(meth_frequency, rank=n_u)
        u = u.clip(0, 1)
        alpha = projection_simplex_sort_2d(alpha)
    elif init_option == "uniform":
        raise NameError("name 'R_trunc' is not defined")   # what the reference does (deconvolution.py:117)
    else:
        raise NotImplementedError(f"init option {init_option!r} is not available on the B200 path")
    return _solve(_lib.DMF_MODE_UNSUPERVISED, u, alpha, meth_frequency, d_x, None, n_u, n_iter1, n_iter2, tol)


def best_of_restarts(meth_frequency, d_x, R_trunc, n_u, init_option, seed, n_restarts, n_iter1, n_iter2, tol, purity=None,
                     distinct_seeds=True):
    """The restart loops of demethify.py:167-203 as ONE batched launch set: n_restarts fits, the one with the lowest final cost
    wins (first on ties, `cost < best_cost` at :199-203).  The reference re-seeds every restart with the same seed, so its
    restarts are one and the same fit: distinct_seeds=False reproduces that (a single fit is run);
    distinct_seeds=True draws restart r from seed + r, the convention ic.py:196 uses.  -> (u, alpha, best restart, costs)."""
    if not distinct_seeds or n_restarts <= 1:
        seeds = [seed]
    else:
        base = seed[0] if isinstance(seed, (list, tuple)) else seed
        seeds = [base + r for r in range(n_restarts)]
    prob = meth_frequency if isinstance(meth_frequency, DeviceProblem) else DeviceProblem(meth_frequency, d_x, R_trunc)
    src = prob if init_option in ("uniform_", "uniform", "beta") else meth_frequency
    inits = []
    for s in seeds:
        if R_trunc is None:
            raise NotImplementedError("best_of_restarts: reference-free fits go through unsupervised_deconv one seed at a time")
        if purity is not None:
            u0, _, a0 = init_BSSMF_md_p(init_option, src, d_x, R_trunc, n_u, purity, seed=s)
        else:
            u0, _, a0 = init_BSSMF_md(init_option, src, d_x, R_trunc, n_u, seed=s)
        inits.append((np.asarray(u0).reshape(-1, n_u), np.asarray(a0)))
    mode = _lib.DMF_MODE_PURITY if purity is not None else _lib.DMF_MODE_PARTIAL
    batch = FitBatch(prob, n_u, [i[0] for i in inits], [i[1] for i in inits], mode=mode, purity=purity)
    res = batch.results(batch.fit(n_iter1, n_iter2, tol))
    costs = [r[3] for r in res]
    best = int(np.argmin(costs))
    _last.update(n_outer=res[best][2], cost=costs[best], launches=batch.launch_count(), geometry=batch.geometry(), engine=batch.engine)
    batch.close()
    return res[best][0], res[best][1], best, costs
