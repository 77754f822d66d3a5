"""Debug helper: per-constraint-set error of the CUDA path vs the CPU oracle."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import towr_b200 as tb
from towr_b200.configs import synthetic_iterates
import oracle_lib


def report(name, terrain=None, B=8, mixed=False):
    f = tb.make_formulation(name, terrain=terrain)
    spec = f.to_spec()
    p = tb.Problem(spec)
    X = synthetic_iterates(p, B)
    terr = None
    if mixed:
        terr = (np.arange(B) % 7).astype(np.int32)
    ref = oracle_lib.batch_eval(spec, X, terrain_ids=terr)
    assert ref["rc"] == 0
    bt = p.batch(B)
    if terr is not None:
        bt.set_terrains(terr)
    out = bt.eval_host(X)
    rp = p.row_ptr()
    print(f"== {name} terrain={terrain} mixed={mixed} n={p.n} m={p.m} nnz={p.nnz} status={out['status'].tolist()}")
    for cname, r0, nr in p.constraint_sets():
        dg = np.abs(out["g"][:, r0:r0 + nr] - ref["g"][:, r0:r0 + nr])
        sg = np.abs(ref["g"][:, r0:r0 + nr])
        s0, s1 = rp[r0], rp[r0 + nr]
        dj = np.abs(out["jac"][:, s0:s1] - ref["jac"][:, s0:s1])
        sj = np.abs(ref["jac"][:, s0:s1])
        relg = (dg / np.maximum(sg, 1e-300))[dg > 1e-14]
        relj = (dj / np.maximum(sj, 1e-300))[dj > 1e-14]
        print(f"  {cname:28s} g: maxabs {dg.max():.2e} maxrel(where abs>1e-14) {relg.max() if relg.size else 0:.2e} | "
              f"jac: maxabs {dj.max():.2e} maxrel {relj.max() if relj.size else 0:.2e} (max|ref| {sj.max():.3g})")


if __name__ == "__main__" and len(sys.argv) == 1:
    report("hopper")
    report("anymal_trot_block")
    report("anymal_trot_block", terrain=tb.GAP)
    report("biped_walk_stairs", mixed=True)
    report("go1_trot_flat", mixed=True, B=16)


def offenders(name, B=8):
    f = tb.make_formulation(name)
    spec = f.to_spec(); p = tb.Problem(spec)
    X = synthetic_iterates(p, B)
    ref = oracle_lib.batch_eval(spec, X)
    out = p.batch(B).eval_host(X)
    iRow, jCol = p.structure()
    d = np.abs(out["jac"] - ref["jac"]); r = np.abs(ref["jac"])
    bad = (d > 1e-14) & (d > 1e-12 * r)
    print("offending jac entries:", bad.sum(), "of", bad.size)
    vs = p.variable_sets(); cs = p.constraint_sets()
    def vname(c):
        for nm, s, k in vs:
            if s <= c < s + k: return f"{nm}[{c - s}]"
    def cname(rw):
        for nm, s, k in cs:
            if s <= rw < s + k: return f"{nm}[{rw - s}]"
    bi, si = np.nonzero(bad)
    for b, s in list(zip(bi, si))[:40]:
        print(f"  b={b} {cname(iRow[s])} x {vname(jCol[s])}: ref={ref['jac'][b, s]:.17g} got={out['jac'][b, s]:.17g}")
    dg = np.abs(out["g"] - ref["g"]); rg = np.abs(ref["g"])
    badg = (dg > 1e-14) & (dg > 1e-12 * rg)
    print("offending g entries:", badg.sum(), "of", badg.size)
    bi, ri = np.nonzero(badg)
    for b, rw in list(zip(bi, ri))[:20]:
        print(f"  b={b} {cname(rw)}: ref={ref['g'][b, rw]:.17g} got={out['g'][b, rw]:.17g}")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "offenders":
    offenders(sys.argv[2] if len(sys.argv) > 2 else "anymal_trot_block", B=int(sys.argv[3]) if len(sys.argv) > 3 else 8)
