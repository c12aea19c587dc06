"""Mirror of the reference's demethify/init_func.py for the B200 path.

  wls_intercept        init_func.py:8-14   — weighted NNLS with intercept; one streaming moment pass + per-sample
                                            Lawson-Hanson on the GPU (dmf_wls_fit).  `wls_all_samples` is the batched
                                            form every internal caller uses (all sample columns in one launch set).
  nndsvd_initialize    init_func.py:40-82  — NNDSVD; the thin SVD runs in cuSOLVER through torch.linalg.svd (a library
                                            call made once per fit), the rest is elementwise glue.
  constrained_nndsvd   init_func.py:17-37
ICA (init_func.py:99-176) forms an M x M covariance and is out of scope.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .engine import DeviceProblem, _handle, _stream_ptr, _even, _pad_cols, to_device


def wls_all_samples(X, d_x, R_full, extra=None, y_is_dx=False, precision=None):
    """wls_intercept for every column of X at once -> (K [+ K2], N) float64 ndarray.
    y = X, or d_x * X when y_is_dx (demethify.py:212).  `extra` is an optional second block of regressor columns
    (the unknown profiles u of deconvolution.py:51)."""
    prob = X if isinstance(X, DeviceProblem) else DeviceProblem(X, d_x, R_full, precision=precision)
    dev, dt = prob.device, prob.dtype
    K2, R2 = 0, None
    if extra is not None:
        e = to_device(extra, dt, dev)
        K2 = e.shape[1]
        R2 = _pad_cols(e, _even(K2))
    out = torch.zeros((prob.K + K2, prob.N), dtype=torch.float64, device=dev)
    desc = _lib.WlsDesc(M=prob.M, N=prob.N, K=prob.K, K2=K2, dtype=_lib.DMF_F64 if dt == torch.float64 else _lib.DMF_F32,
                        wtype=prob.wtype, y_is_dx=int(bool(y_is_dx)), ldx=prob.ldx, ldd=prob.ldx, ldr=_even(prob.K),
                        ldr2=_even(K2), X=prob.X.data_ptr(), D=prob.D.data_ptr(),
                        R1=prob.Rk.data_ptr() if prob.Rk is not None else None, R2=R2.data_ptr() if R2 is not None else None,
                        out=out.data_ptr())
    lib = _lib.lib()
    h = _handle(dev.index if dev.index is not None else torch.cuda.current_device())
    nbytes = C.c_size_t()
    _lib.check(lib.dmf_wls_workspace_bytes(h, C.byref(desc), C.byref(nbytes)))
    ws = torch.empty(nbytes.value + 256, dtype=torch.uint8, device=dev)
    ws_ptr = (ws.data_ptr() + 255) // 256 * 256
    _lib.check(lib.dmf_wls_fit(h, C.byref(desc), C.c_void_p(ws_ptr), nbytes.value, _stream_ptr()))
    return out.cpu().numpy()


def wls_intercept(x, d_x, R_full):
    """init_func.py:8-14.  x: (M,1) or (M,), d_x likewise, R_full: (M,K).  Returns (K,1) for 2-D x, (K,) for 1-D."""
    x = np.asarray(x)
    col = wls_all_samples(x.reshape(-1, 1), np.asarray(d_x).reshape(-1, 1), R_full)
    return col if x.ndim == 2 else col[:, 0]


def _pos_neg(v):
    return np.maximum(v, 0), np.maximum(-v, 0)


def nndsvd_initialize(V, rank, flag=0):
    """init_func.py:40-82 (flag 0/1/2 as in the reference)."""
    V = np.asarray(V, dtype=np.float64)
    if np.any(V < 0):
        raise ValueError("The input matrix contains negative elements.")
    Ut, St, Et = torch.linalg.svd(to_device(V, torch.float64), full_matrices=False)     # cuSOLVER gesvd, once per fit
    U, S, E = Ut.cpu().numpy(), St.cpu().numpy(), Et.cpu().numpy().T
    W = np.zeros((V.shape[0], rank))
    H = np.zeros((rank, V.shape[1]))
    W[:, 0] = np.sqrt(S[0]) * np.abs(U[:, 0])
    H[0, :] = np.sqrt(S[0]) * np.abs(E[:, 0].T)
    for i in range(1, rank):
        uup, uun = _pos_neg(U[:, i])
        vvp, vvn = _pos_neg(E[:, i])
        n_uup, n_vvp = np.linalg.norm(uup, 2), np.linalg.norm(vvp, 2)
        n_uun, n_vvn = np.linalg.norm(uun, 2), np.linalg.norm(vvn, 2)
        termp, termn = n_uup * n_vvp, n_uun * n_vvn
        if termp >= termn:
            W[:, i] = np.sqrt(S[i] * termp) / n_uup * uup
            H[i, :] = np.sqrt(S[i] * termp) / n_vvp * vvp.T
        else:
            W[:, i] = np.sqrt(S[i] * termn) / n_uun * uun
            H[i, :] = np.sqrt(S[i] * termn) / n_vvn * vvn.T
    W[W < 1e-11] = 0
    H[H < 1e-11] = 0
    if flag == 1:
        avg = np.mean(V)
        W[W == 0] = avg
        H[H == 0] = avg
    elif flag == 2:
        avg = np.mean(V)
        W[W == 0] = avg * np.random.uniform(0, 1, size=W[W == 0].shape) / 100
        H[H == 0] = avg * np.random.uniform(0, 1, size=H[H == 0].shape) / 100
    return W, H


def constrained_nndsvd(Y, W1, counts, rank, flag=0):
    """init_func.py:17-37: reference-based fit of every sample, NNDSVD of the floored residual."""
    Y = np.asarray(Y, dtype=np.float64)
    W1 = np.asarray(W1, dtype=np.float64)
    H1 = wls_all_samples(Y, counts, W1)
    Y_residual = np.maximum(Y - W1 @ H1, 1e-8)
    W2, H2 = nndsvd_initialize(Y_residual, rank=rank, flag=flag)
    W2 = np.clip(W2, 0, 1)
    return np.hstack([W1, W2]), np.vstack([H1, H2])
