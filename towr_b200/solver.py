"""Device-resident multi-start solver loop on top of the batched evaluation (BASELINE configs[0] / [4] name IPOPT solves;
neither IPOPT nor ifopt exists in this image).

The stand-in is a deterministic Levenberg-Marquardt FEASIBILITY iteration on the NLP's constraints — what IPOPT's
restoration phase minimises — for every instance of a batch in lock step:

    r  = violation of g(x) against [g_lower, g_upper]              (twb_batch_eval_device: g and the CSR Jacobian values)
    Js = diag(s) J,  rs = s r,   s_i = 1 / max(1, max_k |J_ik|)    (row scaling)
    (Js^T Js + mu I) dx = -Js^T rs                                 (`cg_iters` conjugate-gradient iterations, matrix-free)
    x <- clip(x + dx * min(1, cap / max|dx|), x_lower, x_upper)

The iterates, g, the Jacobian values and all solver vectors stay on the GPU.  An iteration is two library calls:
`twb_batch_eval_device` (g and the CSR Jacobian values of all instances) and `twb_batch_lm_step_device`
(towr_b200/csrc/lm_kernel.cu: CTA = instance, the conjugate gradients run in shared memory on the problem's ONE shared
CSR pattern and its transpose, fixed summation orders, no atomics).  `native=False` keeps the first implementation of the
same step in PyTorch tensor algebra (padded ELL gathers) — 30 x slower, used by the tests as a second opinion.
tests/test_solver_loop.py holds the same algorithm in numpy, driven one instance at a time like ifopt by the CPU
restatement of the reference; all three must walk the same iterates.
"""
import numpy as np


def ell_maps(row_ptr, col_idx, n):
    """Padded index maps of a CSR pattern: (rows, cols_of, rowsT, rows_of) with
    rows[i, k]   = slot of the k-th entry of row i            (pad: nnz)
    cols_of[i,k] = its column                                  (pad: 0)
    rowsT[j, k]  = slot of the k-th entry of column j          (pad: nnz)
    rows_of[j,k] = its row                                     (pad: 0)
    Slots index the CSR value array extended by one zero."""
    row_ptr = np.asarray(row_ptr, np.int64); col_idx = np.asarray(col_idx, np.int64)
    m, nnz = len(row_ptr) - 1, len(col_idx)
    lens = np.diff(row_ptr)
    W = int(lens.max()) if m else 0
    rows = np.full((m, W), nnz, np.int64); cols_of = np.zeros((m, W), np.int64)
    for i in range(m):
        k = np.arange(row_ptr[i], row_ptr[i + 1])
        rows[i, :len(k)] = k; cols_of[i, :len(k)] = col_idx[k]
    row_of_slot = np.repeat(np.arange(m), lens)
    order = np.argsort(col_idx, kind="stable")                 # slots grouped by column, ascending row inside a column
    counts = np.bincount(col_idx, minlength=n)
    Wc = int(counts.max()) if nnz else 0
    rowsT = np.full((n, Wc), nnz, np.int64); rows_of = np.zeros((n, Wc), np.int64)
    start = np.concatenate([[0], np.cumsum(counts)])
    for j in range(n):
        k = order[start[j]:start[j + 1]]
        rowsT[j, :len(k)] = k; rows_of[j, :len(k)] = row_of_slot[k]
    return rows, cols_of, rowsT, rows_of


class BatchedLevenbergMarquardt:
    """All instances of a `Batch` in lock step, everything on the batch's GPU."""

    def __init__(self, batch, x_lower=None, x_upper=None, mu=1e-2, cap=0.1, cg_iters=25, native=True):
        import torch
        self.torch = torch
        self.native = native
        self.batch, p = batch, batch.problem
        self.p, self.B = p, batch.B
        self.dev = torch.device("cuda", batch.device)
        xl, xu, gl, gu = p.bounds()
        to = lambda a: a.to(self.dev, torch.float64) if torch.is_tensor(a) else torch.as_tensor(np.asarray(a, np.float64), device=self.dev)
        self.xl = to(xl if x_lower is None else x_lower); self.xu = to(xu if x_upper is None else x_upper)
        self.gl, self.gu = to(gl), to(gu)
        rows, cols_of, rowsT, rows_of = ell_maps(p.row_ptr(), p.structure()[1], p.n)
        ti = lambda a: torch.as_tensor(a, device=self.dev)
        self.rows, self.cols_of, self.rowsT, self.rows_of = ti(rows), ti(cols_of), ti(rowsT), ti(rows_of)
        self.mu, self.cap, self.cg_iters = mu, cap, cg_iters
        B = self.B
        self.g = torch.empty((B, p.m), dtype=torch.float64, device=self.dev)
        self.jac = None if native else torch.empty((B, p.nnz + 1), dtype=torch.float64, device=self.dev)   # one extra zero: the pad slot
        self.jac_vals = torch.empty((B, p.nnz), dtype=torch.float64, device=self.dev)
        self.status = torch.zeros(B, dtype=torch.int32, device=self.dev)
        self.viol = torch.zeros(B, dtype=torch.float64, device=self.dev)
        self.xl_full = self.xl.expand(B, p.n).contiguous(); self.xu_full = self.xu.expand(B, p.n).contiguous()

    def _violation(self, g):
        t = self.torch
        return t.where(g < self.gl, g - self.gl, t.where(g > self.gu, g - self.gu, t.zeros_like(g)))

    def step(self, X):
        """One iteration for all instances.  X: (B, n) CUDA float64, updated in place.  Returns the largest scaled
        violation per instance BEFORE the step, (B,)."""
        t = self.torch
        from . import capi
        self.batch.eval_device(X, g=self.g, jac=self.jac_vals, status=self.status, flags=capi.EVAL_G | capi.EVAL_JAC)
        if self.native:
            self.batch.lm_step_device(X, self.g, self.jac_vals, self.xl_full, self.xu_full, self.mu, self.cap, self.cg_iters, self.viol)
            return self.viol.clone()
        self.jac[:, :-1] = self.jac_vals; self.jac[:, -1] = 0.0
        A = self.jac[:, self.rows]                                   # (B, m, W) rows of J
        s = 1.0 / t.clamp(A.abs().amax(dim=2), min=1.0)              # (B, m)
        A = A * s[:, :, None]
        AT = self.jac[:, self.rowsT] * s[:, self.rows_of]            # (B, n, Wc) columns of Js
        rs = s * self._violation(self.g)
        Jv = lambda v: (A * v[:, self.cols_of]).sum(dim=2)           # Js v
        JTu = lambda u: (AT * u[:, self.rows_of]).sum(dim=2)         # Js^T u
        b = -JTu(rs)
        dx = t.zeros_like(b); res = b.clone(); pdir = b.clone()
        rr = (res * res).sum(dim=1)
        for _ in range(self.cg_iters):
            Ap = JTu(Jv(pdir)) + self.mu * pdir
            pAp = (pdir * Ap).sum(dim=1)
            alpha = t.where(pAp > 0, rr / pAp, t.zeros_like(rr))
            dx += alpha[:, None] * pdir
            res -= alpha[:, None] * Ap
            rr_new = (res * res).sum(dim=1)
            beta = t.where(rr > 0, rr_new / rr, t.zeros_like(rr))
            pdir = res + beta[:, None] * pdir
            rr = rr_new
        big = dx.abs().amax(dim=1)
        dx *= t.where(big > self.cap, self.cap / big, t.ones_like(big))[:, None]
        X.copy_(t.minimum(t.maximum(X + dx, self.xl), self.xu))
        return rs.abs().amax(dim=1)

    def run(self, X, iters):
        """`iters` iterations; returns the violation history (iters, B) as a CUDA tensor."""
        hist = self.torch.empty((iters, self.B), dtype=self.torch.float64, device=self.dev)
        for it in range(iters):
            hist[it] = self.step(X)
        return hist
