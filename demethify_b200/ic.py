"""Mirror of the reference's demethify/ic.py: choosing the number of unknown cell types.

run_deconvolution / evaluate_best_ic / bicross_validation keep the reference signatures.  The fits of one n_u
(CCC restarts, BCV folds) are one batched launch set; the scalar criteria and the consensus clustering are the
reference's formulas on the host.
"""
import numpy as np
import torch.distributed as dist
import tqdm

from . import _lib
from .deconvolution import init_BSSMF_md, mdwbssmf_deconv, unsupervised_deconv, cost_f_w, last_fit_info
from .engine import DeviceProblem, FitBatch

__all__ = ["compute_bic", "compute_aic", "compute_consensus_matrix", "compute_ccc", "run_deconvolution", "bicross_validation",
           "evaluate_best_ic", "merge_sweep", "sweep_member_supported"]


def _world(group=None):
    return (dist.get_rank(group), dist.get_world_size(group)) if dist.is_available() and dist.is_initialized() else (0, 1)


def merge_sweep(local_results, local_payload, n_values, group=None):
    """Fit sharding of the n_u sweep (ic.py:192-216; every n_u re-seeds and is independent): rank r evaluated the positions
    r, r + world, ... of the sweep.  `local_results` maps position -> criterion, `local_payload` position -> (u, alpha) of this
    rank's candidates for the overall best.  Returns (criteria in sweep order, best position, (u, alpha) of the best) on every
    rank; the best is the reference's: the FIRST position whose criterion is strictly below everything before it."""
    rank, world = _world(group)
    gathered = [local_results]
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, local_results, group=group)
    merged = {}
    for g in gathered:
        merged.update(g)
    values = [merged[p] for p in range(n_values)]
    best, best_pos = float("inf"), None
    for p, v in enumerate(values):
        if v < best:
            best, best_pos = v, p
    payload = local_payload.get(best_pos)
    if world > 1 and best_pos is not None:
        box = [payload if best_pos % world == rank else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, best_pos % world) if group is not None else best_pos % world, group=group)
        payload = box[0]
    return values, best_pos, payload


def compute_bic(cost, n_u, n_cpg, n_ct, n_samples):
    """ic.py:11-15 (the product form is the reference's)."""
    l = n_samples * n_cpg
    k = n_u * n_cpg + (n_ct + n_u - 1) * n_samples
    return 2 * np.log(cost) * k * np.log(l) + (k * np.log(l) * (k + 1)) / (l - k - 1)


def compute_aic(cost, n_u, n_cpg, n_ct, n_samples):
    """ic.py:18-22."""
    l = n_samples * n_cpg
    k = n_u * n_cpg + (n_ct + n_u - 1) * n_samples
    return l * np.log(cost / l) + 2 * k + (2 * k * (k + 1)) / (l - k - 1)


def compute_consensus_matrix(alpha_runs):
    """ic.py:24-37: fraction of the runs in which two samples share their dominant cell type (argmax of the alpha column) - one
    kernel pair of the library (dmf_consensus) on the stacked runs."""
    import ctypes as C
    import torch
    from .engine import _stream_ptr, to_device
    stack = to_device(np.ascontiguousarray(np.stack([np.asarray(a, dtype=np.float64) for a in alpha_runs])), torch.float64)
    n_runs, kt, n = stack.shape
    labels = torch.empty((n_runs, n), dtype=torch.int32, device=stack.device)
    out = torch.empty((n, n), dtype=torch.float64, device=stack.device)
    _lib.check(_lib.lib().dmf_consensus(C.c_void_p(stack.data_ptr()), n_runs, kt, n, C.c_void_p(labels.data_ptr()),
                                        C.c_void_p(out.data_ptr()), _stream_ptr()))
    return out.cpu().numpy()


def compute_ccc(alpha_runs):
    """ic.py:40-45."""
    from scipy.cluster.hierarchy import linkage, cophenet
    from scipy.spatial.distance import pdist
    distance = pdist(compute_consensus_matrix(alpha_runs), metric="euclidean")
    ccc, _ = cophenet(linkage(distance, method="average"), distance)
    return ccc


def run_deconvolution(meth_f, counts, ref, n_u, init_option, seed, iter1, iter2, tol):
    """ic.py:47-55."""
    if ref is not None:
        u, R, alpha = init_BSSMF_md(init_option, meth_f, counts, ref, n_u, seed=seed)
        u, alpha = mdwbssmf_deconv(u, R, alpha, meth_f, counts, ref, n_u, n_iter1=iter1, n_iter2=iter2, tol=tol)
        R = np.hstack((ref, u.reshape(-1, n_u)))
    else:
        u, alpha = unsupervised_deconv(meth_f, n_u, counts, init_option, n_iter1=iter1, n_iter2=iter2, tol=tol, seed=seed)
        R = u
    return u, R, alpha


def _batched_fits(prob_list, ref, n_u, inits, iter1, iter2, tol):
    """Fits sharing one shape -> [(u, R, alpha, cost)]; partial-reference only (ref is not None)."""
    batch = FitBatch(prob_list, n_u, [i[0] for i in inits], [i[1] for i in inits], mode=_lib.DMF_MODE_PARTIAL)
    res = batch.results(batch.fit(iter1, iter2, tol))
    batch.close()
    return [(u, np.hstack((ref, u.reshape(-1, n_u))), a, c) for (u, a, _n, c) in res]


def bicross_validation(meth_f, n_u, counts, iter1, iter2, tol, n_folds=10, seed=None, ref=None, init_option="uniform_", fraction=0.3,
                       base=None):
    """ic.py:58-89.  The fold masks come from numpy's GLOBAL stream, which every fold's init re-seeds:
    the same numpy calls are issued in the same order, then all folds are fitted as one batch."""
    np.random.seed(seed)
    total_press, best_u, best_alpha, min_error = 0, None, None, float("inf")
    meth_f = np.asarray(meth_f)
    counts = np.asarray(counts)
    folds = []
    for _ in range(n_folds):
        train_mask = np.random.rand(*meth_f.shape) < fraction
        test_mask = ~train_mask
        if np.sum(test_mask) == 0 or np.sum(train_mask) == 0:
            continue
        if ref is not None:
            u0, _, a0 = init_BSSMF_md(init_option, meth_f * train_mask, counts * train_mask, ref, n_u, seed=seed)
            folds.append((train_mask, test_mask, u0, a0))
        else:      # the reference-free solver draws its own init from the same re-seeded stream
            u, alpha = unsupervised_deconv(meth_f * train_mask, n_u, counts * train_mask, init_option, n_iter1=iter1, n_iter2=iter2,
                                           tol=tol, seed=seed)
            folds.append((train_mask, test_mask, u, alpha))
    if ref is not None and folds:
        base = base if base is not None else DeviceProblem(meth_f, counts, ref)
        probs = [base.masked(f[0]) for f in folds]
        fits = _batched_fits(probs, np.asarray(ref), n_u, [(f[2], f[3]) for f in folds], iter1, iter2, tol)
    else:
        fits = [(f[2], f[2], f[3], None) for f in folds]
    for (train_mask, test_mask, _u0, _a0), (u, R, alpha, _c) in zip(folds, fits):
        # ||(X - R alpha) o test_mask||_F^2 == weighted cost with 0/1 weights: one more streaming pass
        test_error = cost_f_w(meth_f, R, alpha, test_mask.astype(np.float64)) / np.sum(test_mask)
        total_press += test_error
        if test_error < min_error:
            min_error, best_u, best_alpha = test_error, u, alpha
    return total_press, best_u, best_alpha       # the reference returns total, not mean (Q15)


def sweep_member_supported(K, n_u):
    """The library pads the known and the unknown block to even widths and needs their sum <= 32 (include/demethify_b200.h)."""
    return n_u >= 0 and (K + (K & 1)) + (n_u + (n_u & 1)) <= 32


def _row_sharded_member(prob_local, lo, hi, M, K, N, n_u, init_option, seed, iter1, iter2, tol, group):
    """One sweep member with the CpG rows sharded over the ranks of `group` (BASELINE config 5): every rank draws the same
    initial iterate from the same seed, keeps its rows of u and runs the row-sharded solver -> (u_local, alpha, cost)."""
    from .bootstrap import _shape_only_init
    from .sharded import mdwbssmf_deconv_sharded
    if init_option not in ("uniform_", "beta"):
        raise NotImplementedError("row-sharded sweeps draw the initial iterate from shapes only: --init uniform_ or beta")
    # RandomState(seed) == the reference's set_seed(seed) + global stream (int -> init_genrand, list -> init_by_array, Q1)
    u0, a0 = _shape_only_init(seed, init_option if n_u <= N else "uniform_", M, K, N, n_u, with_zero_guard=True)
    u_loc, alpha, _n, cost = mdwbssmf_deconv_sharded(u0[lo:hi], a0, prob_local, None, None, n_u, n_iter1=iter1, n_iter2=iter2, tol=tol,
                                                     group=group)
    return u_loc, alpha, cost


def evaluate_best_ic(meth_f, ref, counts, init_option, ic, seed, iter1, iter2, tol, n_restarts=5, n_u_values=None, shard="fits",
                     group=None):
    """ic.py:169-218.  n_u_values defaults to the reference's hard-coded range(1, 26) (Q9); 0 may be listed explicitly (AIC / BIC):
    it is the reference-based fit of demethify.py:209-213 (the reference's own loop breaks at 0).

    X, d_x and R_trunc are uploaded ONCE and stay resident in HBM for the whole sweep; the criterion of a member comes from the
    cost the solver finished with (the reference recomputes the same number with cost_f_w, ic.py:206).  Members whose shape the
    library cannot run (known + unknown types, each padded to even, > 32) are reported up front, skipped and recorded as inf.

    Under torch.distributed: shard="fits" (default) gives position p of the sweep to rank p mod world (every member re-seeds and
    is independent); shard="rows" runs EVERY member with the CpG rows sharded over the ranks (one all-reduce per outer
    iteration, sharded.py) - the partitioning for matrices that are large per member (AIC / BIC, shape-only inits)."""
    n_u_values = range(1, 25 + 1) if n_u_values is None else n_u_values
    meth_f = np.asarray(meth_f)
    n_cpg, n_samples = meth_f.shape
    n_ct = ref.shape[1] if ref is not None else 0
    best_n_u, best_u_overall, best_alpha_overall = None, None, None
    if ic == "minka":
        # the reference raises here: run_deconvolution() is called with three arguments missing (ic.py:189, Q7)
        raise TypeError("run_deconvolution() missing 3 required positional arguments: 'iter1', 'iter2', and 'tol'")
    if isinstance(seed, (list, tuple)) and ic == "CCC":
        raise TypeError("can only concatenate list (not \"int\") to list")      # ic.py:196 with `--seed S` (Q1)
    n_u_values = list(n_u_values)
    unsupported = [n for n in n_u_values if not sweep_member_supported(n_ct, n)]
    if unsupported:
        import warnings
        warnings.warn(f"n_u = {unsupported} with {n_ct} known cell types exceed the library limit (known + unknown types, each padded "
                      "to even, <= 32): these members are skipped and recorded as inf", RuntimeWarning, stacklevel=2)
    if 0 in n_u_values and (ic not in ("AIC", "BIC") or ref is None):
        raise ValueError("n_u = 0 (reference-based fit) is available for AIC / BIC with a reference matrix only")
    rank, world = _world(group)
    rows_mode = shard == "rows" and world > 1
    if rows_mode and (ic not in ("AIC", "BIC") or ref is None):
        raise NotImplementedError("row-sharded sweeps: AIC / BIC with a reference matrix")
    if rows_mode:
        from .sharded import row_range
        lo, hi = row_range(n_cpg, rank, world)
        prob = DeviceProblem(meth_f[lo:hi], np.asarray(counts)[lo:hi], np.asarray(ref)[lo:hi])
    else:
        lo, hi = 0, n_cpg
        prob = DeviceProblem(meth_f, counts, ref)                       # resident for the whole sweep
    src = prob if init_option in ("uniform_", "uniform", "beta") else meth_f      # the SVD init reads the host matrix
    local_results, local_payload, local_best = {}, {}, float("inf")
    positions = list(range(len(n_u_values))) if rows_mode else list(range(len(n_u_values)))[rank::world]
    for pos in tqdm.tqdm(positions, disable=rank != 0):
        n_u = n_u_values[pos]
        if n_u in unsupported:
            local_results[pos] = float("inf")
            continue
        if n_u == 0:
            from .init_func import wls_all_samples
            if rows_mode:
                raise NotImplementedError("n_u = 0 in a row-sharded sweep")
            alpha = wls_all_samples(prob, None, None, y_is_dx=True)          # demethify.py:209-213
            u = np.zeros((n_cpg, 0))
            cost = cost_f_w(meth_f, np.asarray(ref), alpha, counts)
            ic_result = compute_bic(cost, 0, n_cpg, n_ct, n_samples) if ic == "BIC" else compute_aic(cost, 0, n_cpg, n_ct, n_samples)
        elif ic == "CCC":
            if ref is not None:
                inits = [init_BSSMF_md(init_option, src, counts, ref, n_u, seed=seed + r) for r in range(n_restarts)]
                fits = _batched_fits(prob, np.asarray(ref), n_u, [(i[0], i[2]) for i in inits], iter1, iter2, tol)
                alpha_runs = [f[2] for f in fits]
                u, alpha = fits[-1][0], fits[-1][2]
            else:
                alpha_runs = []
                for r in range(n_restarts):
                    u, alpha = unsupervised_deconv(src, n_u, counts, init_option, n_iter1=iter1, n_iter2=iter2, tol=tol, seed=seed + r)
                    alpha_runs.append(alpha)
            ic_result = -compute_ccc(alpha_runs)
        elif ic == "BCV":
            ic_result, u, alpha = bicross_validation(meth_f, n_u, counts, iter1, iter2, tol, fraction=0.3, n_folds=n_restarts, seed=seed,
                                                     ref=ref, init_option=init_option, base=prob)
        else:
            if rows_mode:
                u, alpha, cost = _row_sharded_member(prob, lo, hi, n_cpg, n_ct, n_samples, n_u, init_option, seed, iter1, iter2, tol, group)
            elif ref is not None:
                u0, R0, a0 = init_BSSMF_md(init_option, src, counts, ref, n_u, seed=seed)
                u, alpha = mdwbssmf_deconv(u0, R0, a0, prob, None, None, n_u, n_iter1=iter1, n_iter2=iter2, tol=tol)
                cost = last_fit_info()["cost"]
            else:
                u, alpha = unsupervised_deconv(src, n_u, counts, init_option, n_iter1=iter1, n_iter2=iter2, tol=tol, seed=seed)
                cost = last_fit_info()["cost"]
            ic_result = compute_bic(cost, n_u, n_cpg, n_ct, n_samples) if ic == "BIC" else compute_aic(cost, n_u, n_cpg, n_ct, n_samples)
        local_results[pos] = ic_result
        if ic_result < local_best:                 # the overall best (first position of the minimum) is the LAST strict improvement
            local_best = ic_result                 # inside its rank's share: keep only that one
            local_payload = {pos: (u, alpha)}
    if rows_mode:
        # every rank evaluated every member on its rows: criteria are replicated, u of the winner is gathered row range by row range
        values = [local_results[p] for p in range(len(n_u_values))]
        best_pos = min(local_payload) if local_payload else None
        payload = None
        if best_pos is not None:
            u_loc, alpha = local_payload[best_pos]
            parts = [None] * world
            dist.all_gather_object(parts, np.asarray(u_loc), group=group)
            payload = (np.concatenate(parts, axis=0), alpha)
        list_result = values
    else:
        list_result, best_pos, payload = merge_sweep(local_results, local_payload, len(n_u_values), group)
    if best_pos is not None:
        best_n_u, (best_u_overall, best_alpha_overall) = n_u_values[best_pos], payload
    return best_u_overall, best_alpha_overall, best_n_u, list_result
