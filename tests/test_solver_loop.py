"""Solver-in-the-loop stand-in (BASELINE configs[0] / [4] name IPOPT solves; there is no IPOPT in this image):
a deterministic Levenberg-Marquardt feasibility iteration on the NLP's constraints, driven once by the CPU oracle
(one instance at a time, like ifopt) and once by the batched CUDA evaluation (all multi-start instances in lock
step, one twb_batch_eval per iteration).  Both runs must walk the same iterates."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import towr_b200 as tb
import oracle_lib

pytestmark = pytest.mark.gpu


def _violation(g, gl, gu):
    return np.where(g < gl, g - gl, np.where(g > gu, g - gu, 0.0))


def _lm_step(x, g, jac_vals, iRow, jCol, shape, bounds, mu=1e-2, cap=0.1):
    """One damped Gauss-Newton step on the row-scaled constraint violation, step length capped, bounds by projection."""
    xl, xu, gl, gu = bounds
    r = _violation(g, gl, gu)
    J = sp.csr_matrix((jac_vals, (iRow, jCol)), shape=shape)
    scale = 1.0 / np.maximum(1.0, np.abs(J).max(axis=1).toarray().ravel())
    Js, rs = sp.diags(scale) @ J, scale * r
    dx = spla.spsolve((Js.T @ Js + mu * sp.identity(shape[1])).tocsc(), -(Js.T @ rs))
    dx *= min(1.0, cap / np.abs(dx).max())
    return np.clip(x + dx, xl, xu), float(np.abs(rs).max())


def test_lm_feasibility_loop_gpu_matches_oracle():
    spec = tb.make_formulation("hopper").to_spec()
    p = tb.Problem(spec)
    o = oracle_lib.Oracle(spec)
    bounds = p.bounds()
    iRow, jCol = p.structure()
    B, iters = 4, 20
    rng = np.random.default_rng(42)
    x0 = p.GetVariableValues()
    starts = np.clip(x0 + 0.01 * rng.standard_normal((B, p.n)), bounds[0], bounds[1])    # multi-start initial guesses
    # CPU: one instance at a time through the oracle
    X_cpu, viol_cpu = starts.copy(), np.zeros((iters, B))
    for b in range(B):
        x = X_cpu[b]
        for it in range(iters):
            r = o.eval(x)
            x, viol_cpu[it, b] = _lm_step(x, r["g"], r["jac"], iRow, jCol, (p.m, p.n), bounds)
        X_cpu[b] = x
    # GPU: all instances in lock step, one batched evaluation per iteration
    bt = p.batch(B)
    X_gpu, viol_gpu = starts.copy(), np.zeros((iters, B))
    for it in range(iters):
        out = bt.eval_host(X_gpu)
        assert not out["status"].any()
        for b in range(B):
            X_gpu[b], viol_gpu[it, b] = _lm_step(X_gpu[b], out["g"][b], out["jac"][b], iRow, jCol, (p.m, p.n), bounds)
    assert np.all(viol_cpu[-1] < 0.5 * viol_cpu[0])                     # the loop makes progress towards feasibility
    assert np.allclose(viol_gpu, viol_cpu, rtol=1e-6, atol=1e-9)         # same residual history
    assert np.allclose(X_gpu, X_cpu, rtol=1e-6, atol=1e-8)               # same iterates, to the linear solver's conditioning
