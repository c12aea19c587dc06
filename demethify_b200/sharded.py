"""CpG-row sharding of ONE fit over the GPUs of a box (BASELINE config 5).

One process per GPU (torch.distributed, NCCL over NVLink).  Rank r holds a contiguous row range of X, d_x, R_trunc
and u plus a replica of alpha and of the fit state.  Per outer iteration (deconvolution.py:206-221 / :320-335):

  U step      row-local, no communication          (dmf_gram_u_inner on the rows' statistics)
  alpha step  ONE sum all-reduce of the per-sample statistics [G_j | bx_j | ||u||^2]  (Kt (Kt + 1) N + 8 doubles per fit),
              then every rank runs the identical n_iter2 inner iterations on the reduced copy (dmf_gram_alpha_inner)
  cost        ONE sum all-reduce of 8 doubles per fit, then the identical termination test on every rank

plus one max all-reduce (max d_x) at set-up.  Every rank therefore holds the same alpha and makes the same termination
decision; u stays distributed.  `RowShardedFit` is the orchestration; the arithmetic sits behind a small backend
interface whose product implementation (`GpuShardBackend`) drives libdemethify_sm100 through `FitBatch`.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .engine import DeviceProblem, FitBatch, _stream_ptr

__all__ = ["row_range", "GpuShardBackend", "RowShardedFit", "mdwbssmf_deconv_sharded", "last_info"]

_peer_cache = {}    # (group, world, device) -> (symmetric buffer, rendezvous handle, bytes): see enable_peer_exchange
last_info = {}      # of the most recent mdwbssmf_deconv_sharded call: {"peer_exchange": bool, "peer_error": str | None, "collectives": int}


def row_range(M, rank, world):
    """Rows [lo, hi) of rank `rank`: row m belongs to rank floor(m * world / M)."""
    lo = (M * rank + world - 1) // world
    hi = (M * (rank + 1) + world - 1) // world
    return lo, hi


class GpuShardBackend:
    """This rank's rows on its GPU: a one-fit (or few-fit) FitBatch in sharded Gram-engine mode."""

    def __init__(self, X_local, D_local, Rk_local, n_u, U0_local, A0, mode=_lib.DMF_MODE_PARTIAL, purity=None, precision=None, engine=None):
        self.prob = X_local if isinstance(X_local, DeviceProblem) else DeviceProblem(X_local, D_local, Rk_local, precision=precision)
        self.batch = FitBatch(self.prob, n_u, [U0_local], [A0], mode=mode, purity=purity, engine=engine)
        lib, b = _lib.lib(), self.batch
        _lib.check(lib.dmf_batch_set_sharded(b.b, 1, _stream_ptr()))
        # fused engine where the shape allows it: ONE pass and ONE all-reduce per outer iteration (RowShardedFit.outer)
        self.fused = b.engine == "fused"
        loc, glo, per, so = C.c_void_p(), C.c_void_p(), C.c_int64(), C.c_int64()
        _lib.check(lib.dmf_batch_stats_buffers(b.b, C.byref(loc), C.byref(glo), C.byref(per), C.byref(so)))
        base = b.ws.data_ptr()
        nbytes = per.value * 8 * b.n_fits

        def view(ptr):
            off = ptr - base
            return b.ws[off:off + nbytes].view(torch.float64).view(b.n_fits, per.value)
        self.local, self.glob = view(loc.value), view(glo.value)
        self.scal_off = so.value
        self.has_known = b.K > 0
        self.device = b.device
        self.supports_graphs = True
        self.peer = None

    def enable_peer_exchange(self, group=None):
        """All-reduce over NVLink peer memory inside one kernel of the library (dmf_gram_exchange) instead of NCCL: allocates a
        symmetric buffer that every rank of `group` maps (torch symmetric memory) and hands the peer pointers to the library.
        Returns False (and leaves the NCCL path in place) when symmetric memory is not available."""
        if not (dist.is_initialized() and dist.get_world_size(group) > 1):
            return False
        try:
            import torch.distributed._symmetric_memory as symm_mem
            lib, b = _lib.lib(), self.batch
            world, rank = dist.get_world_size(group), dist.get_rank(group)
            nbytes = C.c_size_t()
            _lib.check(lib.dmf_batch_peer_bytes(b.b, world, C.byref(nbytes)))
            # the symmetric buffer and its rendezvous (an exchange of IPC handles through the store: tens of ms) are kept for the
            # process and reused by later fits of the same or a smaller statistics block; every use starts from zeroed flags
            key = (id(group) if group is not None else 0, world, self.device.index)
            cached = _peer_cache.get(key)
            if cached is not None and cached[2] >= nbytes.value:
                buf, hdl = cached[0], cached[1]
            else:
                buf = symm_mem.empty((nbytes.value + 7) // 8, dtype=torch.float64, device=self.device)
                hdl = symm_mem.rendezvous(buf, group if group is not None else dist.group.WORLD)
                _peer_cache[key] = (buf, hdl, nbytes.value)
            buf.zero_()
            ptrs = (C.c_void_p * world)(*[int(p) for p in hdl.buffer_ptrs])
            torch.cuda.synchronize()
            dist.barrier(group)                      # every rank's buffer (flags) is zero before anybody pushes
            _lib.check(lib.dmf_batch_set_peers(b.b, rank, world, ptrs, nbytes.value, _stream_ptr()))
            self.peer = (buf, hdl)
        except Exception as e:                       # no symmetric memory on this system: keep NCCL
            self.peer_error = repr(e)
            self.peer = None
        # the ranks must take the same path: one rank on NCCL while the others wait in the exchange kernel would hang the job
        ok = torch.tensor([1 if self.peer is not None else 0], dtype=torch.int32, device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            self.peer = None
        return self.peer is not None

    def exchange(self, which):
        _lib.check(_lib.lib().dmf_gram_exchange(self.batch.b, int(which), _stream_ptr()))

    def reserve(self, n_inner_total):
        """Make the following steps allocation- and sync-free (CUDA-graph capture)."""
        _lib.check(_lib.lib().dmf_batch_reserve_momentum(self.batch.b, int(n_inner_total), _stream_ptr()))

    # ---- statistics blocks: (n_fits, doubles_per_fit) tensors [G | bx | scal(8)]
    def stats_local(self):
        return self.local

    def stats_global(self):
        return self.glob

    # ---- this rank's arithmetic
    def rowgram(self, initial, tol):
        self.batch.gram_rowgram(initial, tol)

    def u_inner(self, n_iter2):
        self.batch.gram_u_inner(n_iter2)

    def panels(self, known_block):
        self.batch.gram_panels(known_block)

    def alpha_inner(self, n_iter2):
        self.batch.gram_alpha_inner(n_iter2)

    def fused_pass(self, n_iter2, tol):
        self.batch.fused_pass(n_iter2, tol)

    def fused_alpha_commit(self, n_iter2, tol):
        _lib.check(_lib.lib().dmf_fused_alpha_commit(self.batch.b, int(n_iter2), float(tol), _stream_ptr()))

    def finalize_cost(self, initial, tol):
        _lib.check(_lib.lib().dmf_gram_finalize_cost(self.batch.b, int(bool(initial)), float(tol), _stream_ptr()))

    def all_done(self):
        sts = self.batch.states()
        bad = [i for i, s in enumerate(sts) if s.done == 3]
        if bad:
            raise _lib.DmfError(f"non-finite values reached the simplex projection (fit {bad[0]})")
        return all(s.done != 0 for s in sts)

    def results(self):
        """[(u_local, alpha, n_outer, cost)] — u holds only this rank's rows."""
        return self.batch.results()

    def close(self):
        self.batch.close()


class RowShardedFit:
    """Outer loop of one row-sharded fit; `backend` does this rank's arithmetic, `group` is the torch.distributed group."""

    def __init__(self, backend, group=None):
        self.be, self.group = backend, group
        self.so = backend.scal_off
        self.collectives = 0

    def _allreduce(self, t, op=dist.ReduceOp.SUM):
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(t, op=op, group=self.group)
        self.collectives += 1

    def _reduce_scal(self, with_max):
        if getattr(self.be, "peer", None) is not None:          # one kernel: push to the peers, wait, rank-ordered sum (max for max d_x)
            self.be.exchange(2 if with_max else 1)
            self.collectives += 1
            return
        loc, glo, so = self.be.stats_local(), self.be.stats_global(), self.so
        sc = loc[:, so:so + 8].contiguous()
        if with_max:
            mx = sc[:, 3].clone()
            self._allreduce(mx, dist.ReduceOp.MAX)
        self._allreduce(sc)
        if with_max:
            sc[:, 3] = mx
        glo[:, so:so + 8] = sc

    def init(self):
        """deconvolution.py:192-204: cost, ||R||^2, max d_x over ALL rows; then this rank's known x known statistics."""
        be = self.be
        be.rowgram(True, 0.0)
        self._reduce_scal(with_max=True)
        be.finalize_cost(True, 0.0)
        if be.has_known:
            be.panels(True)

    def _reduce_blocks(self):
        be = self.be
        if getattr(be, "peer", None) is not None:
            be.exchange(0)                            # G_j, bx_j, scalars over all rows, over NVLink peer memory
            self.collectives += 1
        else:
            glo = be.stats_global()
            glo.copy_(be.stats_local())
            self._allreduce(glo)                      # the same through NCCL

    def use_fused(self, n_iter2):
        return getattr(self.be, "fused", False) and 1 <= n_iter2 <= 64

    def outer(self, n_iter2, tol):
        be = self.be
        if self.use_fused(n_iter2):
            # fused engine: cost of the incoming iterate + U step + panel in ONE pass over this rank's rows, ONE all-reduce of
            # [G_j | bx_j | cost, ||u||^2], then the test / commit / alpha iterations identically on every rank
            be.fused_pass(n_iter2, tol)
            self._reduce_blocks()
            be.fused_alpha_commit(n_iter2, tol)
            self.pending = True
            return
        if n_iter2 > 0:
            be.u_inner(n_iter2)                       # row-local
            be.panels(False)
            if getattr(be, "peer", None) is not None:
                be.exchange(0)                        # G_j, bx_j, ||u||^2 over all rows, over NVLink peer memory
                self.collectives += 1
            else:
                glo = be.stats_global()
                glo.copy_(be.stats_local())
                self._allreduce(glo)                  # the same through NCCL
            be.alpha_inner(n_iter2)                   # identical on every rank
        be.rowgram(False, tol)
        self._reduce_scal(with_max=False)
        be.finalize_cost(False, tol)                  # identical termination decision on every rank

    def finish(self, tol):
        """Fused engine: the cost of the last iterate is still pending after the last outer() - one cost-only pass."""
        if getattr(self, "pending", False):
            self.pending = False
            if not self.be.all_done():
                self.be.rowgram(False, tol)
                self._reduce_scal(with_max=False)
                self.be.finalize_cost(False, tol)

    def capture_outer(self, n_iter2, tol):
        """One outer iteration (kernels, D2D copies and both NCCL all-reduces) as a CUDA graph: a replay costs one launch
        instead of ~12 host calls, which is what bounds strong scaling once a rank's share of the rows takes < 0.2 ms."""
        self.outer(n_iter2, tol)                      # eager once: NCCL and the allocator are warm before the capture
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.outer(n_iter2, tol)
        return graph

    def fit(self, n_iter1, n_iter2, tol, use_graph=None):
        self.init()
        issued, chunk = 0, 2
        graph = None
        if use_graph is None:      # opt-in: DMF_SHARDED_GRAPH=1 (see DESIGN.md 6 for what was measured)
            import os
            use_graph = getattr(self.be, "supports_graphs", False) and n_iter1 > 4 and os.environ.get("DMF_SHARDED_GRAPH", "0") == "1"
        if use_graph:
            self.be.reserve((n_iter1 + 2) * max(n_iter2, 1))
            graph = self.capture_outer(n_iter2, tol)
            issued = 1
            if self.be.all_done():
                return self.be.results()
        while issued < n_iter1:
            todo = min(chunk, n_iter1 - issued)
            for _ in range(todo):
                if graph is not None:
                    graph.replay()
                else:
                    self.outer(n_iter2, tol)
            issued += todo
            if self.be.all_done():                    # the state is replicated: every rank leaves the loop together
                self.pending = False
                break
            chunk = min(chunk * 2, 16)
        self.finish(tol)
        return self.be.results()


def mdwbssmf_deconv_sharded(u_local, alpha, X_local, d_local, R_local, n_u, n_iter1=100000, n_iter2=50, tol=1e-3, purity=None,
                            group=None):
    """mdwbssmf_deconv / mdwbssmf_deconv_p (deconvolution.py:190-223, :305-337) on a row-sharded problem: every rank passes
    ITS rows of u, X, d_x, R_trunc (see `row_range`) and the full alpha; returns (u_local, alpha, n_outer, cost)."""
    mode = _lib.DMF_MODE_PURITY if purity is not None else (_lib.DMF_MODE_PARTIAL if R_local is not None else _lib.DMF_MODE_UNSUPERVISED)
    be = GpuShardBackend(X_local, d_local, R_local, n_u, np.asarray(u_local).reshape(-1, n_u), np.asarray(alpha), mode=mode, purity=purity)
    import os
    if os.environ.get("DMF_PEER_XCHG", "1") != "0":          # the in-kernel NVLink all-reduce is the default; NCCL when symmetric memory is unavailable
        be.enable_peer_exchange(group)
    fit = RowShardedFit(be, group)
    try:
        (u, a, n_outer, cost), = fit.fit(n_iter1, n_iter2, tol)
    finally:
        last_info.update(peer_exchange=be.peer is not None, peer_error=getattr(be, "peer_error", None), collectives=fit.collectives)
        torch.cuda.synchronize()
        be.close()
    return u, a, n_outer, cost
