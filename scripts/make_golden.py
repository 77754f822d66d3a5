"""Generates tests/golden/oracle_golden.npz from the CPU oracle (the reference itself cannot
be built in this image: it needs Eigen 3 + ifopt — see DESIGN.md).  Seeded, deterministic."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import towr_b200 as tb
from towr_b200.configs import synthetic_iterates
import oracle_lib

SEED = 4242
out = {"seed": np.int64(SEED)}
for name in ("hopper", "anymal_trot_block", "biped_walk_stairs"):
    spec = tb.make_formulation(name).to_spec()
    o = oracle_lib.Oracle(spec); p = tb.Problem(spec)
    X = synthetic_iterates(p, 2, seed=SEED)
    rp, ci = o.structure()
    res = [o.eval(x) for x in X]
    out[f"{name}_x"] = X
    out[f"{name}_row_ptr"] = rp; out[f"{name}_col_idx"] = ci
    out[f"{name}_g"] = np.stack([r["g"] for r in res]); out[f"{name}_jac"] = np.stack([r["jac"] for r in res])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "oracle_golden.npz"), **out)
print("wrote golden:", {k: v.shape for k, v in out.items() if hasattr(v, "shape")})

# ---- solution post-processing (fpowr): trajectory, initial guesses, footstep plan -> tests/golden/postproc_golden.npz
post = {"seed": np.int64(SEED), "dt": np.float64(0.1), "times": np.array([0.0, 0.013, 0.4, 0.99, 1.0, 1.37, 2.0]),
        "time_horizon": np.float64(2.0)}
for name in ("hopper", "hyq_gallop_gap"):
    spec = tb.make_formulation(name).to_spec()
    o = oracle_lib.Oracle(spec); p = tb.Problem(spec)
    x = synthetic_iterates(p, 1, seed=SEED)[0]
    post[f"{name}_x"] = x
    post[f"{name}_trajectory"] = o.trajectory(x, float(post["dt"]))
    post[f"{name}_initial_guesses"] = o.initial_guesses(x, post["times"])
    post[f"{name}_footstep_plan"] = o.footstep_plan(x, float(post["time_horizon"]))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "postproc_golden.npz"), **post)
print("wrote post-processing golden:", {k: v.shape for k, v in post.items() if hasattr(v, "shape") and v.shape})
