"""Mirror of the reference's demethify/ic.py: choosing the number of unknown cell types.

run_deconvolution / evaluate_best_ic / bicross_validation keep the reference signatures.  The fits of one n_u
(CCC restarts, BCV folds) are one batched launch set; the scalar criteria and the consensus clustering are the
reference's formulas on the host.
"""
import numpy as np
import torch.distributed as dist
import tqdm

from . import _lib
from .deconvolution import init_BSSMF_md, mdwbssmf_deconv, unsupervised_deconv, cost_f_w
from .engine import DeviceProblem, FitBatch

__all__ = ["compute_bic", "compute_aic", "compute_consensus_matrix", "compute_ccc", "run_deconvolution", "bicross_validation",
           "evaluate_best_ic", "merge_sweep"]


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
    """ic.py:24-37."""
    n_samples = alpha_runs[0].shape[1]
    consensus = np.zeros((n_samples, n_samples))
    for alpha in alpha_runs:
        lab = np.argmax(alpha, axis=0)
        consensus += (lab[:, None] == lab[None, :])
    return consensus / len(alpha_runs)


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


def bicross_validation(meth_f, n_u, counts, iter1, iter2, tol, n_folds=10, seed=None, ref=None, init_option="uniform_", fraction=0.3):
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
        base = DeviceProblem(meth_f, counts, ref)
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


def evaluate_best_ic(meth_f, ref, counts, init_option, ic, seed, iter1, iter2, tol, n_restarts=5, n_u_values=None):
    """ic.py:169-218.  n_u_values defaults to the reference's hard-coded range(1, 26) (Q9)."""
    n_u_values = range(1, 25 + 1) if n_u_values is None else n_u_values
    n_cpg, n_samples = np.asarray(meth_f).shape
    n_ct = ref.shape[1] if ref is not None else 0
    best_ic, best_n_u, best_u_overall, best_alpha_overall = float("inf"), None, None, None
    list_result = []
    if ic == "minka":
        # the reference raises here: run_deconvolution() is called with three arguments missing (ic.py:189, Q7)
        raise TypeError("run_deconvolution() missing 3 required positional arguments: 'iter1', 'iter2', and 'tol'")
    if isinstance(seed, (list, tuple)) and ic == "CCC":
        raise TypeError("can only concatenate list (not \"int\") to list")      # ic.py:196 with `--seed S` (Q1)
    n_u_values = list(n_u_values)
    rank, world = _world()
    local_results, local_payload, local_best = {}, {}, float("inf")
    positions = list(range(len(n_u_values)))[rank::world]          # under torch.distributed the sweep is sharded over the ranks
    for pos in tqdm.tqdm(positions, disable=rank != 0):
        n_u = n_u_values[pos]
        if ic == "CCC":
            if ref is not None:
                prob = DeviceProblem(meth_f, counts, ref)
                inits = [init_BSSMF_md(init_option, meth_f, counts, ref, n_u, seed=seed + r) for r in range(n_restarts)]
                fits = _batched_fits(prob, np.asarray(ref), n_u, [(i[0], i[2]) for i in inits], iter1, iter2, tol)
                alpha_runs = [f[2] for f in fits]
                u, alpha = fits[-1][0], fits[-1][2]
            else:
                alpha_runs = []
                for r in range(n_restarts):
                    u, R, alpha = run_deconvolution(meth_f, counts, ref, n_u, init_option, seed + r, iter1, iter2, tol)
                    alpha_runs.append(alpha)
            ic_result = -compute_ccc(alpha_runs)
        elif ic == "BCV":
            ic_result, u, alpha = bicross_validation(meth_f, n_u, counts, iter1, iter2, tol, fraction=0.3, n_folds=n_restarts, seed=seed,
                                                     ref=ref, init_option=init_option)
        else:
            u, R, alpha = run_deconvolution(meth_f, counts, ref, n_u, init_option, seed, iter1, iter2, tol)
            cost = cost_f_w(meth_f, R, alpha, counts)
            ic_result = compute_bic(cost, n_u, n_cpg, n_ct, n_samples) if ic == "BIC" else compute_aic(cost, n_u, n_cpg, n_ct, n_samples)
        local_results[pos] = ic_result
        if ic_result < local_best:                 # the overall best (first position of the minimum) is the LAST strict improvement
            local_best = ic_result                 # inside its rank's share: keep only that one
            local_payload = {pos: (u, alpha)}
    list_result, best_pos, payload = merge_sweep(local_results, local_payload, len(n_u_values))
    if best_pos is not None:
        best_n_u, (best_u_overall, best_alpha_overall) = n_u_values[best_pos], payload
    return best_u_overall, best_alpha_overall, best_n_u, list_result
