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
