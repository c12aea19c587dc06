"""demethify_b200 — B200-native (sm_100a) implementation of DeMethify's NMF deconvolution hot path.

Python entry points mirror the reference's modules (`deconvolution`, `bootstrap`, `ic`, `init_func`,
`demethify` CLI); the arithmetic runs in libdemethify_sm100.so (hand-written CUDA, C ABI declared in
include/demethify_b200.h).  There is no CPU fallback: importing works anywhere, computing needs a B200.
"""
__version__ = "0.1.0"

_PRECISION = "fp64"


def set_precision(mode):
    """'fp64' (default; max|d alpha| <= 1e-6 vs the reference) or 'fp32' (<= 1e-4)."""
    global _PRECISION
    if mode not in ("fp64", "fp32"):
        raise ValueError("precision must be 'fp64' or 'fp32'")
    _PRECISION = mode


def get_precision():
    return _PRECISION


_ENGINE = "auto"


def set_engine(mode):
    """How the inner iterations run on the device (include/demethify_b200.h, DMF_ENGINE_*):
    'stream' = one streaming pass over X, d_x per reference inner iteration;
    'gram'   = two streaming passes per OUTER iteration build per-row / per-sample sufficient statistics and the
               n_iter2 inner iterations run on those (n_u <= 4, or n_u <= 8 with K <= 6);
    'fused'  = ONE streaming pass per outer iteration: row statistics -> U iterations -> Gram panel on a single visit of every
               row tile (FP64, n_u <= 2, K <= 8, N <= 256, n_iter2 <= 64);
    'auto'   = 'fused' where the library supports the shape, else 'gram', else 'stream' (default)."""
    global _ENGINE
    if mode not in ("auto", "fused", "gram", "stream"):
        raise ValueError("engine must be 'auto', 'fused', 'gram' or 'stream'")
    _ENGINE = mode


def get_engine():
    return _ENGINE
