"""numpy (fp64) restatement of the reference's BSSMF deconvolution path.

TEST INFRASTRUCTURE — see oracle/__init__.py.  Every function names the
reference lines (`/root/reference/demethify/...`) whose arithmetic it follows,
in the same operation order, so that results agree with the live reference to
rounding (the gemm/nrm2 inner orders are BLAS-internal and cannot be fixed).

Notation: M CpG rows, N samples, K known cell types, n_u unknown types,
Kt = K + n_u.  X: MxN methylation frequencies, D: MxN coverage weights,
Rk: MxK known profiles, U: Mxn_u unknown profiles, A: KtxN proportions.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "weighted_cost", "simplex_project_columns", "lawson_hanson_nnls", "wls_simplex_fit",
    "legacy_stream", "draw_init", "draw_init_purity", "nndsvd_factors", "nndsvd_with_reference",
    "u_inner_loop", "alpha_inner_loop", "solve_partial_reference", "frank_wolfe_alpha",
    "solve_purity", "solve_unsupervised", "bootstrap_seed_list", "bootstrap_row_indices",
    "bootstrap_fits", "percentile_bounds", "aic_value", "bic_value", "consensus_ccc",
    "fit_for_ic", "bicross_validation_press", "ic_sweep", "reference_based_fit",
]


# --------------------------------------------------------------------------- cost / projection
def weighted_cost(X, R, A, D):
    """sum D*(X - R@A)^2, formed as ||sqrt(D)*(X-RA)||_F^2  (deconvolution.py:15-17)."""
    resid = np.sqrt(D) * (X - R @ A)
    return float(np.linalg.norm(resid) ** 2)


def simplex_project_columns(V, z=1.0):
    """Sort-based Euclidean projection of every column of V onto {w>=0, sum w = z}
    (deconvolution.py:21-37).  rho is the LAST index passing the test, as in the
    reference loop; a column with no passing index raises like the reference does."""
    p, n = V.shape
    srt = np.sort(V, axis=0)[::-1]
    pi = np.cumsum(srt, axis=0) - z
    ranks = np.arange(1, p + 1, dtype=V.dtype)[:, None]
    ok = (srt - pi / ranks) > 0
    if not ok.any(axis=0).all():
        raise ZeroDivisionError("simplex projection: no admissible rho (NaN input?)")
    rho = p - 1 - np.argmax(ok[::-1], axis=0)
    theta = pi[rho, np.arange(n)] / (rho + 1)
    return np.maximum(V - theta[None, :], 0.0)


# --------------------------------------------------------------------------- NNLS / reference-based fit
def lawson_hanson_nnls(Amat, b, max_iter=None, tol=None):
    """Lawson-Hanson active-set NNLS: argmin_{c>=0} ||Amat c - b||_2.
    Third-party on the reference path: scipy.optimize.nnls (scipy 1.11.4 pinned,
    requirements.txt:3) reached through sklearn LinearRegression(positive=True)
    (init_func.py:9).  The minimiser is unique for full-column-rank Amat, so any
    exact active-set method reproduces it to rounding."""
    m, n = Amat.shape
    if max_iter is None:
        max_iter = 3 * n
    AtA = Amat.T @ Amat
    Atb = Amat.T @ b
    if tol is None:
        tol = 10 * max(m, n) * np.spacing(1.0) * np.linalg.norm(Atb, 1)
    x = np.zeros(n)
    passive = np.zeros(n, dtype=bool)
    w = Atb - AtA @ x
    it = 0
    while (not passive.all()) and (w[~passive] > tol).any():
        cand = np.where(~passive, w, -np.inf)
        passive[int(np.argmax(cand))] = True
        s = np.zeros(n)
        s[passive] = np.linalg.solve(AtA[np.ix_(passive, passive)], Atb[passive])
        while it < max_iter and s[passive].min() <= 0:
            it += 1
            bad = passive & (s <= 0)
            step = (x[bad] / (x[bad] - s[bad])).min()
            x = x * (1 - step) + step * s
            passive[x <= tol] = False
            x[~passive] = 0.0
            s = np.zeros(n)
            if passive.any():
                s[passive] = np.linalg.solve(AtA[np.ix_(passive, passive)], Atb[passive])
        x = s
        w = Atb - AtA @ x
    return x


def wls_simplex_fit(y, d, Rfull):
    """init_func.py:8-14: weighted least squares with intercept and c>=0, then
    c / max(sum c, 1e-10).  sklearn's fit = weighted centring of R and y, sqrt(weight)
    row scaling, NNLS on the centred system.  Returns shape (K,1) for 2-D y, (K,) for 1-D."""
    w = np.asarray(d, dtype=float).ravel()
    yy = np.asarray(y, dtype=float).reshape(len(w))
    r_off = np.average(Rfull, axis=0, weights=w)
    y_off = np.average(yy, weights=w)
    sw = np.sqrt(w)
    coef = lawson_hanson_nnls((Rfull - r_off) * sw[:, None], (yy - y_off) * sw)
    coef = coef / max(coef.sum(), 1e-10)
    return coef.reshape(-1, 1) if np.ndim(y) == 2 else coef


def reference_based_fit(X, D, Rk):
    """nbunknown=0 path, demethify.py:209-213: regress the methylated counts D*X on Rk
    with weights D, one sample at a time."""
    cols = [wls_simplex_fit(D[:, j:j + 1] * X[:, j:j + 1], D[:, j:j + 1], Rk) for j in range(X.shape[1])]
    return np.concatenate(cols, axis=1)


# --------------------------------------------------------------------------- initialisation
def legacy_stream(seed):
    """The reference seeds numpy's GLOBAL legacy MT19937 stream (deconvolution.py:9-11).
    RandomState(seed) yields the same sequence; a list seed selects init_by_array."""
    return np.random.RandomState(seed)


def _pos(v):
    return np.maximum(v, 0.0)


def nndsvd_factors(V, rank):
    """init_func.py:40-82 with flag=0: NNDSVD from the thin SVD of V."""
    if np.any(V < 0):
        raise ValueError("The input matrix contains negative elements.")
    from scipy.linalg import svd
    Us, S, Vt = svd(V, full_matrices=False)
    W = np.zeros((V.shape[0], rank))
    H = np.zeros((rank, V.shape[1]))
    W[:, 0] = np.sqrt(S[0]) * np.abs(Us[:, 0])
    H[0, :] = np.sqrt(S[0]) * np.abs(Vt[0, :])
    for i in range(1, rank):
        a, b = Us[:, i], Vt[i, :]
        ap, an, bp, bn = _pos(a), _pos(-a), _pos(b), _pos(-b)
        nap, nbp = np.linalg.norm(ap, 2), np.linalg.norm(bp, 2)
        nan_, nbn = np.linalg.norm(an, 2), np.linalg.norm(bn, 2)
        tp, tn = nap * nbp, nan_ * nbn
        if tp >= tn:
            W[:, i] = np.sqrt(S[i] * tp) / nap * ap
            H[i, :] = np.sqrt(S[i] * tp) / nbp * bp
        else:
            W[:, i] = np.sqrt(S[i] * tn) / nan_ * an
            H[i, :] = np.sqrt(S[i] * tn) / nbn * bn
    W[W < 1e-11] = 0
    H[H < 1e-11] = 0
    return W, H


def nndsvd_with_reference(X, Rk, D, rank):
    """init_func.py:17-37: per-sample reference-based fit, NNDSVD of the floored residual."""
    H1 = np.zeros((Rk.shape[1], X.shape[1]))
    for j in range(X.shape[1]):
        H1[:, j] = wls_simplex_fit(X[:, j], D[:, j], Rk)
    W2, H2 = nndsvd_factors(np.maximum(X - Rk @ H1, 1e-8), rank)
    return np.clip(W2, 0, 1), np.vstack([H1, H2])


def _draw_u_alpha(option, X, D, Rk, n_u, seed):
    rs = legacy_stream(seed)
    M, K = Rk.shape
    N = X.shape[1]
    if option != "uniform_" and n_u > N:
        option = "uniform_"                                   # deconvolution.py:44-45
    if option == "uniform":                                     # :47-52
        U = rs.uniform(size=(M, n_u))
        Rfull = np.c_[Rk, U]
        A = np.concatenate([wls_simplex_fit(X[:, j:j + 1], D[:, j:j + 1], Rfull) for j in range(N)], axis=1)
    elif option == "uniform_":                                  # :54-56
        U = rs.uniform(size=(M, n_u))
        A = rs.dirichlet(np.ones(K + n_u), N).T
    elif option == "beta":                                      # :58-61
        half = np.ones((M, n_u)) * 0.5
        U = rs.beta(half, half)
        A = rs.dirichlet(np.ones(K + n_u), N).T
    elif option == "SVD":                                       # :68-71 (projection applied by caller)
        U, A = nndsvd_with_reference(X, Rk, D, n_u)
    else:
        raise ValueError(f"oracle does not restate init option {option!r}")
    return option, U, A


def draw_init(option, X, D, Rk, n_u, seed=None):
    """init_BSSMF_md, deconvolution.py:40-78 (ICA not restated: out of scope)."""
    option, U, A = _draw_u_alpha(option, X, D, Rk, n_u, seed)
    if option == "SVD":
        A = simplex_project_columns(A)
    A = np.array(A)
    if not A[-n_u:][0].all():                                   # :74-76 zero guard on first unknown row
        A[-n_u:][0] = 1e-10
        A[:-n_u] = (1 - 1e-10) * A[:-n_u]
    return U, np.c_[Rk, U], A


def draw_init_purity(option, X, D, Rk, n_u, purity, seed=None):
    """init_BSSMF_md_p, deconvolution.py:228-267 (no zero guard; purity-scaled projection for SVD)."""
    N = X.shape[1]
    if option != "uniform" and n_u > N:                         # :232-234
        option = "uniform"
    option, U, A = _draw_u_alpha(option, X, D, Rk, n_u, seed)
    if option == "SVD":                                         # :262 (unknown block NOT scaled)
        A = np.vstack((purity * simplex_project_columns(A[:-n_u]), simplex_project_columns(A[-n_u:])))
    return U, np.c_[Rk, U], np.array(A)


# --------------------------------------------------------------------------- inner loops
def _extrapolation(a_prev, l_old, l_new):
    a_next = (1 + np.sqrt(1 + 4 * a_prev * a_prev)) / 2
    beta = min((a_prev - 1) / a_next, 0.9999 * np.sqrt(l_old / l_new))
    return a_next, beta


def u_inner_loop(U, A, n_inner, a1, l_w_old, l_w, U_prev, X, Rk, n_u, D):
    """update_u, deconvolution.py:81-90."""
    A_known, A_unk = A[:-n_u], A[-n_u:]
    for _ in range(n_inner):
        a1, beta = _extrapolation(a1, l_w_old, l_w)
        U_t = U + beta * (U - U_prev)
        U_prev = U
        U = np.clip(U_t + (D * (X - Rk @ A_known - U_t @ A_unk)) @ A_unk.T / l_w, 0, 1)
        l_w_old = l_w
    return U, U_prev, a1, l_w_old


def alpha_inner_loop(n_inner, A, a2, l_h_old, l_h, A_prev, R, D, X):
    """update_alpha, deconvolution.py:93-102."""
    for _ in range(n_inner):
        a2, beta = _extrapolation(a2, l_h_old, l_h)
        A_t = A + beta * (A - A_prev)
        A_prev = A
        A = simplex_project_columns(A_t + (R.T @ (D * (X - R @ A_t))) / l_h)
        l_h_old = l_h
    return A, A_prev, a2, l_h_old


# --------------------------------------------------------------------------- outer solvers
def solve_partial_reference(U, R, A, X, D, Rk, n_u, n_iter1=100000, n_iter2=50, tol=1e-3, trace=None):
    """mdwbssmf_deconv, deconvolution.py:190-223.  `trace` (dict) receives n_outer and the cost list."""
    a1 = a2 = 1.0
    U_prev, A_prev = U.copy(), A.copy()
    dmax2 = D.max() ** 2
    l_w = np.linalg.norm(A[-n_u:]) ** 2 * dmax2
    l_w_old = l_w
    l_h = np.linalg.norm(R) ** 2 * dmax2
    l_h_old = l_h
    cf = weighted_cost(X, R, A, D)
    costs = [cf]
    n_outer = 0
    for _ in range(n_iter1):
        cf0 = cf
        U, U_prev, a1, l_w_old = u_inner_loop(U, A, n_iter2, a1, l_w_old, l_w, U_prev, X, Rk, n_u, D)
        R = np.hstack((Rk, U.reshape(-1, n_u)))
        l_h = np.linalg.norm(R) ** 2 * dmax2
        A, A_prev, a2, l_h_old = alpha_inner_loop(n_iter2, A, a2, l_h_old, l_h, A_prev, R, D, X)
        l_w = np.linalg.norm(A[-n_u:]) ** 2 * dmax2
        cf = weighted_cost(X, R, A, D)
        costs.append(cf)
        n_outer += 1
        if abs(cf - cf0) < tol:
            break
    if trace is not None:
        trace["n_outer"] = n_outer
        trace["costs"] = np.array(costs)
    return U, A


def frank_wolfe_alpha(Rk, U, X, A1, A2, purity, n_inner, D):
    """frank_wolfe_nmf, deconvolution.py:280-302.  Vertex = first argmin of the gradient block,
    scaled by purity (known block) and 1-purity (unknown block); gamma_k = 2/(k+2)."""
    A1, A2 = A1.copy(), A2.copy()
    cols = np.arange(X.shape[1])
    for k in range(n_inner):
        g1 = -Rk.T @ (D * (X - Rk @ A1 - U @ A2))
        g2 = -U.T @ (D * (X - Rk @ A1 - U @ A2))
        S1 = np.zeros_like(A1)
        S2 = np.zeros_like(A2)
        S1[np.argmin(g1, axis=0), cols] = purity
        S2[np.argmin(g2, axis=0), cols] = 1 - purity
        gamma = 2 / (k + 2)
        A1 = (1 - gamma) * A1 + gamma * S1
        A2 = (1 - gamma) * A2 + gamma * S2
    return A1, A2


def solve_purity(U, R, A, X, D, Rk, n_u, purity, n_iter1=100, n_iter2=500, tol=1e-3, trace=None):
    """mdwbssmf_deconv_p, deconvolution.py:305-337."""
    a1 = 1.0
    U_prev = U.copy()
    A1, A2 = A[:-n_u], A[-n_u:]
    dmax2 = D.max() ** 2
    l_w = np.linalg.norm(A2) ** 2 * dmax2
    l_w_old = l_w
    cf = weighted_cost(X, R, A, D)
    costs = [cf]
    n_outer = 0
    for _ in range(n_iter1):
        cf0 = cf
        U, U_prev, a1, l_w_old = u_inner_loop(U, A, n_iter2, a1, l_w_old, l_w, U_prev, X, Rk, n_u, D)
        R = np.hstack((Rk, U.reshape(-1, n_u)))
        A1, A2 = frank_wolfe_alpha(Rk, U, X, A1, A2, purity, n_iter2, D)
        l_w = np.linalg.norm(A2) ** 2 * dmax2
        A = np.vstack((A1, A2))
        cf = weighted_cost(X, R, A, D)
        costs.append(cf)
        n_outer += 1
        if abs(cf - cf0) < tol:
            break
    if trace is not None:
        trace["n_outer"] = n_outer
        trace["costs"] = np.array(costs)
    return U, A


def solve_unsupervised(X, n_u, D, option, n_iter1=100000, n_iter2=20, tol=1e-3, seed=None, trace=None):
    """unsupervised_deconv, deconvolution.py:107-184 (uniform_/beta/SVD inits).  Quirk kept: the
    U gradient is evaluated at the post-swap U, not at the extrapolated point (:163)."""
    rs = legacy_stream(seed)
    M, N = X.shape
    if option != "uniform_" and n_u > N:
        option = "uniform_"
    if option == "uniform_":
        U = rs.uniform(size=(M, n_u))
        A = rs.dirichlet(np.ones(n_u), N).T
    elif option == "beta":
        half = np.ones((M, n_u)) * 0.5
        U = rs.beta(half, half)
        A = rs.dirichlet(np.ones(n_u), N).T
    elif option == "SVD":
        U, A = nndsvd_factors(X, n_u)
        U = U.clip(0, 1)
        A = simplex_project_columns(A)
    else:
        raise ValueError(f"oracle does not restate unsupervised init {option!r}")
    a1 = a2 = 1.0
    U_prev, A_prev = U.copy(), A.copy()
    dmax2 = D.max() ** 2
    l_w = np.linalg.norm(A[-n_u:]) ** 2 * dmax2
    l_w_old = l_w
    l_h = np.linalg.norm(U) ** 2 * dmax2
    l_h_old = l_h
    cf = weighted_cost(X, U, A, D)
    costs = [cf]
    n_outer = 0
    for _ in range(n_iter1):
        cf0 = cf
        for _i in range(n_iter2):
            a1, beta = _extrapolation(a1, l_w_old, l_w)
            U_t = U + beta * (U - U_prev)
            U_prev = U
            U = np.clip(U_t + (D * (X - U @ A)) @ A.T / l_w, 0, 1)
            l_w_old = l_w
        l_h = np.linalg.norm(U) ** 2 * dmax2
        A, A_prev, a2, l_h_old = alpha_inner_loop(n_iter2, A, a2, l_h_old, l_h, A_prev, U, D, X)
        l_w = np.linalg.norm(A[-n_u:]) ** 2 * dmax2
        cf = weighted_cost(X, U, A, D)
        costs.append(cf)
        n_outer += 1
        if abs(cf - cf0) < tol:
            break
    if trace is not None:
        trace["n_outer"] = n_outer
        trace["costs"] = np.array(costs)
    return U, A


# --------------------------------------------------------------------------- bootstrap driver
def bootstrap_seed_list(seed, n_bootstrap):
    """bootstrap.py:26-27: the seed ACCUMULATES, seed_i = seed_{i-1} + i."""
    out, s = [], seed
    for i in range(n_bootstrap):
        s = s + i
        out.append(s)
    return out


def bootstrap_row_indices(seed, M):
    """sklearn.utils.resample(..., random_state=seed) (bootstrap.py:28; sklearn 1.2.2 pinned):
    indices = RandomState(seed).randint(0, M, size=(M,)), applied to every array alike."""
    return np.random.RandomState(seed).randint(0, M, size=(M,))


def bootstrap_fits(n_bootstrap, n_u, X, D, Rk, option, n_iter1, n_iter2, tol, purity_pct, seed):
    """The per-resample loop of bt_ci (bootstrap.py:26-46).  purity_pct is the CLI list; bt_ci
    uses purity/100 (NOT 1 - purity/100).  Returns alphas (B,Kt,N) and Us (B,M,n_u)."""
    purity = None if not purity_pct else np.array(purity_pct) / 100.0
    alphas, us = [], []
    for s in bootstrap_seed_list(seed, n_bootstrap):
        idx = bootstrap_row_indices(s, X.shape[0])
        Xb, Db, Rb = X[idx], D[idx], Rk[idx]
        if n_u == 0:
            alphas.append(reference_based_fit(Xb, Db, Rb))
            continue
        if purity is not None:
            U, R, A = draw_init_purity(option, Xb, Db, Rb, n_u, purity, seed=s)
            U, A = solve_purity(U, R, A, Xb, Db, Rb, n_u, purity, n_iter1, n_iter2, tol)
        else:
            U, R, A = draw_init(option, Xb, Db, Rb, n_u, seed=s)
            U, A = solve_partial_reference(U, R, A, Xb, Db, Rb, n_u, n_iter1, n_iter2, tol)
        alphas.append(A)
        us.append(U)
    return np.array(alphas), (np.array(us) if us else None)


def percentile_bounds(stack, confidence_level):
    """bootstrap.py:12-14,53-54,77-78: np.percentile (linear interpolation) over the fit axis."""
    a = 1 - confidence_level / 100
    lo, hi = 100 * (a / 2), 100 * (1 - a / 2)
    return np.percentile(stack, lo, axis=0), np.percentile(stack, hi, axis=0)


# --------------------------------------------------------------------------- ic driver
def _n_params(n_u, n_cpg, n_ct, n_samples):
    return n_u * n_cpg + (n_ct + n_u - 1) * n_samples


def bic_value(cost, n_u, n_cpg, n_ct, n_samples):
    """ic.py:11-15 (product form kept)."""
    l = n_samples * n_cpg
    k = _n_params(n_u, n_cpg, n_ct, n_samples)
    return 2 * np.log(cost) * k * np.log(l) + (k * np.log(l) * (k + 1)) / (l - k - 1)


def aic_value(cost, n_u, n_cpg, n_ct, n_samples):
    """ic.py:18-22."""
    l = n_samples * n_cpg
    k = _n_params(n_u, n_cpg, n_ct, n_samples)
    return l * np.log(cost / l) + 2 * k + (2 * k * (k + 1)) / (l - k - 1)


def consensus_ccc(alpha_runs):
    """ic.py:24-45: argmax cluster assignment per sample, co-clustering frequency, cophenetic
    correlation of average linkage on euclidean distances between consensus rows."""
    from scipy.cluster.hierarchy import linkage, cophenet
    from scipy.spatial.distance import pdist
    n = alpha_runs[0].shape[1]
    cons = np.zeros((n, n))
    for A in alpha_runs:
        lab = np.argmax(A, axis=0)
        cons += (lab[:, None] == lab[None, :])
    cons /= len(alpha_runs)
    dist = pdist(cons, metric="euclidean")
    ccc, _ = cophenet(linkage(dist, method="average"), dist)
    return ccc


def fit_for_ic(X, D, Rk, n_u, option, seed, n_iter1, n_iter2, tol):
    """run_deconvolution, ic.py:47-55."""
    if Rk is not None:
        U, R, A = draw_init(option, X, D, Rk, n_u, seed=seed)
        U, A = solve_partial_reference(U, R, A, X, D, Rk, n_u, n_iter1, n_iter2, tol)
        return U, np.hstack((Rk, U.reshape(-1, n_u))), A
    U, A = solve_unsupervised(X, n_u, D, option, n_iter1, n_iter2, tol, seed=seed)
    return U, U, A


def bicross_validation_press(X, n_u, D, n_iter1, n_iter2, tol, n_folds, seed, Rk, option, fraction=0.3):
    """bicross_validation, ic.py:58-89.  Masks come from the GLOBAL stream which each fold's init
    re-seeds: replayed here with the global np.random exactly as the reference does."""
    np.random.seed(seed)
    total, best = 0.0, (np.inf, None, None)
    for _ in range(n_folds):
        train = np.random.rand(*X.shape) < fraction
        test = ~train
        if test.sum() == 0 or train.sum() == 0:
            continue
        U, R, A = fit_for_ic(X * train, D * train, Rk, n_u, option, seed, n_iter1, n_iter2, tol)
        np.random.seed(seed)                     # what set_seed inside the init did to the global stream ...
        _replay_init_draws(option, X.shape, Rk, n_u)   # ... followed by the init's own draws
        err = np.linalg.norm((X - R @ A) * test, "fro") ** 2 / test.sum()
        total += err
        if err < best[0]:
            best = (err, U, A)
    return total, best[1], best[2]


def _replay_init_draws(option, shape, Rk, n_u):
    """Advance the global stream the way init (uniform_/beta) would have, so the next fold's mask
    matches the reference (our oracle inits use a private RandomState)."""
    M, N = shape
    K = 0 if Rk is None else Rk.shape[1]
    if option != "uniform_" and n_u > N:
        option = "uniform_"
    if option in ("uniform_", "uniform"):
        np.random.uniform(size=(M, n_u))
    elif option == "beta":
        half = np.ones((M, n_u)) * 0.5
        np.random.beta(half, half)
    if option in ("uniform_", "beta"):
        np.random.dirichlet(np.ones(K + n_u), N)


def ic_sweep(X, Rk, D, option, ic, seed, n_iter1, n_iter2, tol, n_restarts=5, n_u_values=range(1, 26)):
    """evaluate_best_ic, ic.py:169-218 (AIC/BIC/CCC/BCV; minka crashes in the reference, Q7)."""
    n_cpg, n_samples = X.shape
    n_ct = 0 if Rk is None else Rk.shape[1]
    best = (np.inf, None, None, None)
    values = []
    for n_u in n_u_values:
        if ic == "CCC":
            runs = []
            for r in range(n_restarts):
                U, R, A = fit_for_ic(X, D, Rk, n_u, option, seed + r, n_iter1, n_iter2, tol)
                runs.append(A)
            val = -consensus_ccc(runs)
        elif ic == "BCV":
            val, U, A = bicross_validation_press(X, n_u, D, n_iter1, n_iter2, tol, n_restarts, seed, Rk, option)
        else:
            U, R, A = fit_for_ic(X, D, Rk, n_u, option, seed, n_iter1, n_iter2, tol)
            cost = weighted_cost(X, R, A, D)
            val = bic_value(cost, n_u, n_cpg, n_ct, n_samples) if ic == "BIC" else aic_value(cost, n_u, n_cpg, n_ct, n_samples)
        values.append(val)
        if val < best[0]:
            best = (val, n_u, U, A)
    return best[2], best[3], best[1], values
