"""numpy (fp64) restatement of the Gram-form outer iteration (demethify_b200/csrc/dmf_gram.cuh) as a shard backend.

TEST INFRASTRUCTURE — see oracle/__init__.py.  Same interface as demethify_b200.sharded.GpuShardBackend, so that
`RowShardedFit` (the product's multi-GPU orchestration) can be exercised on CPU with the gloo backend, and the
re-association the Gram form introduces can be checked against the reference-shaped oracle (bssmf_numpy) without a GPU.

Reference lines restated: set-up deconvolution.py:192-204, update_u :82-89 (unsupervised variant :157-164),
update_alpha :94-101 with projection :21-37, frank_wolfe_nmf :285-299, cost/termination :218-221.
"""
import numpy as np
import torch

from .bssmf_numpy import simplex_project_columns


def momentum_table(n):
    """a_t (a_0 = 1, a_{t+1} = (1 + sqrt(1 + 4 a_t^2)) / 2, :84) and (a_t - 1) / a_{t+1} for t < n."""
    a = np.empty(n + 1)
    a[0] = 1.0
    for t in range(n):
        a[t + 1] = (1 + np.sqrt(1 + 4 * a[t] * a[t])) / 2
    return a, (a[:-1] - 1) / a[1:]


class NumpyShardBackend:
    def __init__(self, X, D, Rk, n_u, U0, A0, mode="partial", purity=None):
        self.X = np.asarray(X, dtype=np.float64)
        self.D = np.asarray(D, dtype=np.float64)
        self.K = 0 if Rk is None else Rk.shape[1]
        self.Rk = np.zeros((self.X.shape[0], 0)) if Rk is None else np.asarray(Rk, dtype=np.float64)
        self.n_u, self.mode = n_u, mode
        self.Kt = self.K + n_u
        self.N = self.X.shape[1]
        self.u, self.u_prev = np.array(U0, dtype=np.float64).reshape(-1, n_u), np.array(U0, dtype=np.float64).reshape(-1, n_u)
        self.A, self.A_prev = np.array(A0, dtype=np.float64), np.array(A0, dtype=np.float64)
        self.purity = None if purity is None else np.asarray(purity, dtype=np.float64)
        self.n_gram, self.n_gbx = self.Kt * self.Kt * self.N, self.Kt * self.N
        self.scal_off = self.n_gram + self.n_gbx
        per = self.scal_off + 8
        self.local = torch.zeros((1, per), dtype=torch.float64)
        self.glob = torch.zeros((1, per), dtype=torch.float64)
        self.has_known = self.K > 0
        self.st = dict(done=0, n_outer=0, t_u=0, t_a=0)
        self.costs = []
        self.mom_a, self.mom_m = momentum_table(4096)

    def _mom(self, t_hi):
        if t_hi + 1 > len(self.mom_m):
            self.mom_a, self.mom_m = momentum_table(2 * (t_hi + 1))

    def stats_local(self):
        return self.local

    def stats_global(self):
        return self.glob

    # ---- streaming pass 1: row statistics + this shard's cost
    def rowgram(self, initial, tol):
        if self.st["done"]:
            return
        Ak, Au = self.A[:self.K], self.A[self.K:]
        c = self.X - self.Rk @ Ak
        self.b = (self.D * c) @ Au.T
        self.H = np.einsum("mj,qj,rj->mqr", self.D, Au, Au)
        sc = self.local[0, self.scal_off:].numpy()
        sc[0] = float(np.sum(self.D * (c - self.u @ Au) ** 2))
        if initial:
            sc[1], sc[2], sc[3] = float(np.sum(self.Rk ** 2)), float(np.sum(self.u ** 2)), float(self.D.max()) if self.D.size else 0.0

    def finalize_cost(self, initial, tol):
        st = self.st
        if st["done"]:
            return
        sc = self.glob[0, self.scal_off:].numpy()
        cf = float(sc[0])
        if initial:
            st["dmax2"] = float(sc[3]) ** 2
            st["ssq_rk"], st["ssq_u"] = float(sc[1]), float(sc[2])
            st["l_w"] = st["l_w_old"] = np.linalg.norm(self.A[self.K:]) ** 2 * st["dmax2"]
            st["l_h"] = st["l_h_old"] = np.sqrt(sc[1] + sc[2]) ** 2 * st["dmax2"]
            st.update(cf=cf, n_outer=0, t_u=0, t_a=0)
            self.costs = [cf]
        else:
            prev = st["cf"]
            st["cf"] = cf
            st["n_outer"] += 1
            self.costs.append(cf)
            if abs(cf - prev) < tol:
                st["done"] = 1

    # ---- n_iter2 update_u iterations on (b, H), row-local
    def u_inner(self, n2):
        st = self.st
        if st["done"]:
            return
        self._mom(st["t_u"] + n2)
        l_w = st["l_w"]
        caps = [0.9999 * np.sqrt(st["l_w_old"] / l_w), 0.9999 * np.sqrt(l_w / l_w)]
        u, up = self.u, self.u_prev
        for it in range(n2):
            beta = min(self.mom_m[st["t_u"] + it], caps[0 if it == 0 else 1])
            ut = u + beta * (u - up)
            ug = u if self.mode == "unsupervised" else ut
            g = self.b - np.einsum("mqr,mr->mq", self.H, ug)
            up = u
            u = np.clip(ut + g / l_w, 0, 1)
        self.u, self.u_prev = u, up
        st["t_u"] += n2
        if n2 > 0:
            st["l_w_old"] = l_w
        self.local[0, self.scal_off + 4] = float(np.sum(u ** 2))

    # ---- streaming pass 2: per-sample statistics of this shard
    def panels(self, known_block):
        if self.st["done"]:
            return
        R = np.hstack([self.Rk, self.u])
        G = np.einsum("mj,mk,ml->klj", self.D, R, R)
        bx = np.einsum("mj,mk->kj", self.D * self.X, R)
        self.local[0, :self.n_gram] = torch.from_numpy(G.reshape(-1))
        self.local[0, self.n_gram:self.scal_off] = torch.from_numpy(bx.reshape(-1))

    # ---- n_iter2 update_alpha / Frank-Wolfe iterations on the all-reduced (G, bx)
    def alpha_inner(self, n2):
        st = self.st
        if st["done"]:
            return
        g = self.glob[0].numpy()
        G = g[:self.n_gram].reshape(self.Kt, self.Kt, self.N)
        bx = g[self.n_gram:self.scal_off].reshape(self.Kt, self.N)
        st["ssq_u"] = float(g[self.scal_off + 4])
        l_h = st["l_h"] = np.sqrt(st["ssq_rk"] + st["ssq_u"]) ** 2 * st["dmax2"]
        A, Ap = self.A, self.A_prev
        if self.mode != "purity":
            self._mom(st["t_a"] + n2)
            caps = [0.9999 * np.sqrt(st["l_h_old"] / l_h), 0.9999 * np.sqrt(l_h / l_h)]
            for it in range(n2):
                beta = min(self.mom_m[st["t_a"] + it], caps[0 if it == 0 else 1])
                At = A + beta * (A - Ap)
                grad = bx - np.einsum("klj,lj->kj", G, At)
                Ap = A
                A = simplex_project_columns(At + grad / l_h)
            st["t_a"] += n2
            if n2 > 0:
                st["l_h_old"] = l_h
        else:
            K, cols = self.K, np.arange(self.N)
            for it in range(n2):
                grad = -(bx - np.einsum("klj,lj->kj", G, A))
                S = np.zeros_like(A)
                S[np.argmin(grad[:K], axis=0), cols] = self.purity
                S[K + np.argmin(grad[K:], axis=0), cols] = 1 - self.purity
                gamma = 2 / (it + 2)
                A = (1 - gamma) * A + gamma * S
        self.A, self.A_prev = A, Ap
        st["l_w"] = np.linalg.norm(A[self.K:]) ** 2 * st["dmax2"]

    def all_done(self):
        return self.st["done"] != 0

    def results(self):
        return [(self.u.copy(), self.A.copy(), self.st["n_outer"], self.st["cf"])]

    def close(self):
        pass


def fit_multiplicity_form(X, D, Rk, idx, U0, A0, n_iter1, n_iter2, tol):
    """A bootstrap resample (rows idx of X, D, Rk; bootstrap.py:28) fitted WITHOUT materialising it — the algebra of the library's
    multiplicity form (csrc/dmf_gram.cuh: rowgram4 with row multiplicities, u_inner_mult, usum, MULT panel, cost_cross), in numpy.
    U0 is position-indexed (one row per resampled position, reference order).  Returns (U (positions), A, n_outer, cost)."""
    X = np.asarray(X, dtype=np.float64)
    D = np.asarray(D, dtype=np.float64)
    M, N = X.shape
    K = Rk.shape[1]
    n_u = U0.shape[1]
    mult = np.bincount(idx, minlength=M).astype(np.float64)
    row_of = np.asarray(idx)                                   # source row of every position
    u, up = np.array(U0, dtype=np.float64), np.array(U0, dtype=np.float64)
    A, Ap = np.array(A0, dtype=np.float64), np.array(A0, dtype=np.float64)
    mom_a, mom_m = momentum_table((n_iter1 + 1) * max(n_iter2, 1) + 1)
    live = mult > 0
    dmax2 = float(D[live].max()) ** 2
    ssq_rk = float(np.sum(mult[:, None] * Rk ** 2))
    # known x known statistics, once: sum_p d R R^T = sum_m mult_m d_m R_m R_m^T
    Gkk = np.einsum("m,mj,mk,ml->klj", mult, D, Rk, Rk)
    bk = np.einsum("m,mj,mk->kj", mult, D * X, Rk)

    def row_stats():
        c = X - Rk @ A[:K]
        b = (D * c) @ A[K:].T                                  # per SOURCE row
        H = np.einsum("mj,qj,rj->mqr", D, A[K:], A[K:])
        cc = np.sum(D * c * c, axis=1)
        cross = -2 * np.einsum("pq,pq->p", u, b[row_of]) + np.einsum("pq,pqr,pr->p", u, H[row_of], u)
        return b, H, float(np.sum(mult * cc) + np.sum(cross))

    b, H, cf = row_stats()
    l_w = l_w_old = np.linalg.norm(A[K:]) ** 2 * dmax2
    l_h = l_h_old = np.sqrt(ssq_rk + np.sum(u ** 2)) ** 2 * dmax2
    t_u = t_a = n_outer = 0
    for _ in range(n_iter1):
        cf0 = cf
        caps = [0.9999 * np.sqrt(l_w_old / l_w), 0.9999]
        for it in range(n_iter2):                              # update_u per position on the statistics of its source row
            beta = min(mom_m[t_u + it], caps[0 if it == 0 else 1])
            ut = u + beta * (u - up)
            g = b[row_of] - np.einsum("pqr,pr->pq", H[row_of], ut)
            up, u = u, np.clip(ut + g / l_w, 0, 1)
        t_u += n_iter2
        if n_iter2 > 0:
            l_w_old = l_w
        l_h = np.sqrt(ssq_rk + np.sum(u ** 2)) ** 2 * dmax2
        # per-source-row sums of the positions: s_m = sum u_p, S_m = sum u_p u_p^T; then the alpha statistics from the SHARED matrices
        s = np.zeros((M, n_u)); np.add.at(s, row_of, u)
        S = np.zeros((M, n_u, n_u)); np.add.at(S, row_of, np.einsum("pq,pr->pqr", u, u))
        Kt = K + n_u
        G = np.zeros((Kt, Kt, N)); bx = np.zeros((Kt, N))
        G[:K, :K], bx[:K] = Gkk, bk
        G[K:, :K] = np.einsum("mj,mq,mk->qkj", D, s, Rk)
        G[:K, K:] = np.transpose(G[K:, :K], (1, 0, 2))
        G[K:, K:] = np.einsum("mj,mqr->qrj", D, S)
        bx[K:] = np.einsum("mj,mq->qj", D * X, s)
        caps = [0.9999 * np.sqrt(l_h_old / l_h), 0.9999]
        for it in range(n_iter2):
            beta = min(mom_m[t_a + it], caps[0 if it == 0 else 1])
            At = A + beta * (A - Ap)
            grad = bx - np.einsum("klj,lj->kj", G, At)
            Ap, A = A, simplex_project_columns(At + grad / l_h)
        t_a += n_iter2
        if n_iter2 > 0:
            l_h_old = l_h
        l_w = np.linalg.norm(A[K:]) ** 2 * dmax2
        b, H, cf = row_stats()
        n_outer += 1
        if abs(cf - cf0) < tol:
            break
    return u, A, n_outer, cf
