"""Solver-in-the-loop stand-in (BASELINE configs[0] / [4] name IPOPT solves; there is no IPOPT in this image):
a deterministic Levenberg-Marquardt feasibility iteration on the NLP's constraints, driven once by the CPU oracle
(one instance at a time, like ifopt) and once by the batched CUDA evaluation (all multi-start instances in lock
step, one twb_batch_eval per iteration).  Both runs must walk the same iterates."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import towr_b200 as tb
from towr_b200.configs import synthetic_iterates_fast
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


class LevenbergMarquardtReference:
    """towr_b200.solver.BatchedLevenbergMarquardt restated in numpy for ONE instance, driven by the CPU oracle (one
    instance at a time, like ifopt::Problem under IPOPT): the same row scaling, the same fixed number of conjugate-gradient
    iterations on (Js^T Js + mu I) dx = -Js^T rs through the same padded index maps, the same step cap and projection."""

    def __init__(self, oracle, problem, terrain, x_lower, x_upper, mu=1e-2, cap=0.1, cg_iters=25):
        from towr_b200.solver import ell_maps
        self.o, self.p = oracle, problem
        self.terrain = terrain
        _, _, self.gl, self.gu = problem.bounds()
        self.xl, self.xu = x_lower, x_upper
        self.rows, self.cols_of, self.rowsT, self.rows_of = ell_maps(problem.row_ptr(), problem.structure()[1], problem.n)
        self.mu, self.cap, self.cg_iters = mu, cap, cg_iters

    def step(self, x):
        self.o.set_terrain(self.terrain)
        r = self.o.eval(x)
        assert r["rc"] == 0
        jac = np.append(r["jac"], 0.0)
        A = jac[self.rows]
        s = 1.0 / np.maximum(np.abs(A).max(axis=1), 1.0)
        A = A * s[:, None]
        AT = jac[self.rowsT] * s[self.rows_of]
        rs = s * _violation(r["g"], self.gl, self.gu)
        Jv = lambda v: (A * v[self.cols_of]).sum(axis=1)
        JTu = lambda u: (AT * u[self.rows_of]).sum(axis=1)
        b = -JTu(rs)
        dx = np.zeros_like(b); res = b.copy(); pdir = b.copy(); rr = res @ res
        for _ in range(self.cg_iters):
            Ap = JTu(Jv(pdir)) + self.mu * pdir
            pAp = pdir @ Ap
            alpha = rr / pAp if pAp > 0 else 0.0
            dx += alpha * pdir; res -= alpha * Ap
            rr_new = res @ res
            pdir = res + (rr_new / rr if rr > 0 else 0.0) * pdir
            rr = rr_new
        big = np.abs(dx).max()
        if big > self.cap:
            dx *= self.cap / big
        return np.minimum(np.maximum(x + dx, self.xl), self.xu), float(np.abs(rs).max())


def test_native_solver_step_matches_the_tensor_algebra_step():
    """twb_batch_lm_step_device (CTA = instance, conjugate gradients in shared memory) against the same step written in PyTorch
    tensor algebra: the two walk the same iterates (different summation orders: agreement to rounding, not bit for bit)."""
    import torch
    from towr_b200.solver import BatchedLevenbergMarquardt
    for name, B in (("anymal_trot_block", 70), ("hopper", 5)):
        p = tb.Problem(tb.make_formulation(name).to_spec())
        bt = p.batch(B)
        X0 = torch.from_numpy(synthetic_iterates_fast(p, B, seed=21)).cuda()
        Xa, Xb = X0.clone(), X0.clone()
        a = BatchedLevenbergMarquardt(bt, native=True); b = BatchedLevenbergMarquardt(bt, native=False)
        for it in range(4):
            va, vb = a.step(Xa), b.step(Xb)
            torch.cuda.synchronize()
            assert torch.allclose(va, vb, rtol=1e-12, atol=1e-14)
            assert torch.allclose(Xa, Xb, rtol=1e-8, atol=1e-10), (name, it, float((Xa - Xb).abs().max()))
        assert not torch.equal(Xa, X0)


def test_device_resident_multistart_loop_config5_matches_oracle_driven_loop():
    """BASELINE configs[4] at one GPU's size: 4096 Anymal multi-start instances on mixed Slope / Chimney / Gap terrains with
    goal-randomised initial guesses and bounds (set up by the device kernel), 12 Levenberg-Marquardt iterations with the
    iterates never leaving the GPU; a subsample of the instances is re-run by the oracle-driven numpy twin and must walk
    the same iterates."""
    import torch
    from towr_b200.solver import BatchedLevenbergMarquardt
    spec = tb.make_formulation("anymal_trot_mixed").to_spec()
    p = tb.Problem(spec)
    B, iters = 4096, 12
    rng = np.random.default_rng(77)
    terr = rng.choice([tb.SLOPE, tb.CHIMNEY, tb.GAP], B).astype(np.int32)
    goals = np.column_stack([rng.uniform(1.0, 2.0, B), rng.uniform(-0.2, 0.2, B), np.full(B, 0.5), np.zeros(B), np.zeros(B), rng.uniform(-0.2, 0.2, B)])
    bt = p.batch(B); bt.set_terrains(terr)
    x0, xl, xu = bt.goal_instances_device(torch.from_numpy(goals).cuda())
    X = torch.minimum(torch.maximum(x0 + 0.01 * torch.from_numpy(rng.standard_normal((B, p.n))).cuda(), xl), xu)   # multi-start perturbations
    starts = X.cpu().numpy().copy()
    lm = BatchedLevenbergMarquardt(bt, x_lower=xl, x_upper=xu)
    lm.run(X.clone(), 1)                         # warm-up on a copy: graph capture, pattern upload, module load
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    hist = lm.run(X, iters)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    hist = hist.cpu().numpy(); X_gpu = X.cpu().numpy()
    assert not lm.status.any().item()
    assert (hist[-1] < hist[0]).mean() > 0.9 and np.median(hist[-1]) < np.median(hist[0])   # the batch moves towards feasibility (steps are capped at 0.1)
    # the oracle-driven twin on a subsample (every terrain present)
    sub = np.concatenate([np.flatnonzero(terr == t)[:6] for t in (tb.SLOPE, tb.CHIMNEY, tb.GAP)])
    o = oracle_lib.Oracle(spec)
    xl_h, xu_h = xl.cpu().numpy(), xu.cpu().numpy()
    worst_x, worst_v = 0.0, 0.0
    for b in sub:
        ref = LevenbergMarquardtReference(o, p, int(terr[b]), xl_h[b], xu_h[b])
        x = starts[b].copy()
        for it in range(iters):
            x, v = ref.step(x)
            worst_v = max(worst_v, abs(v - hist[it, b]) / max(1e-9, abs(v)))
        worst_x = max(worst_x, float(np.abs(x - X_gpu[b]).max()))
        assert np.allclose(x, X_gpu[b], rtol=1e-6, atol=1e-8), b
    assert worst_v < 1e-6
    import json, os
    rec = {"workload": "anymal_trot_mixed (BASELINE configs[4] shard)", "instances": B, "iterations": iters, "cg_iters_per_iteration": lm.cg_iters,
           "seconds": dt, "lm_iterations_per_s": iters / dt, "instance_iterations_per_s": B * iters / dt,
           "violation_median_first_last": [float(np.median(hist[0])), float(np.median(hist[-1]))],
           "subsample": {"instances": int(len(sub)), "max_abs_iterate_difference_vs_oracle_driven_loop": worst_x,
                         "max_rel_violation_difference": worst_v}}
    try:
        out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
        os.makedirs(out_dir, exist_ok=True)
        json.dump(rec, open(os.path.join(out_dir, "solver_loop.json"), "w"), indent=1)
    except OSError:
        pass
