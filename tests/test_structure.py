"""CPU tests: the product's structure builder against the oracle, hand-verified
counts, gait tables, and the C ABI surface.  No GPU needed."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import towr_b200 as tb
from towr_b200 import capi
import oracle_lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = [
    ("hopper", None, {}),
    ("anymal_trot_block", None, {}),
    ("anymal_trot_block", tb.GAP, dict(goal_xy=(2.0, 0.3), goal_yaw=0.4)),
    ("biped_walk_stairs", None, {}),
    ("go1_trot_flat", tb.SLOPE, dict(t_total=2.4)),
    ("anymal_trot_mixed", tb.CHIMNEY_LR, dict(goal_xy=(1.2, -0.2))),
    ("anymal_trot_block_base_rom", None, {}),
    ("hopper_base_rom", None, {}),
    ("hyq_gallop_gap", None, {}),                      # BASELINE configs[3]: phase durations optimised
    ("hyq_gallop_gap", tb.STAIRS, dict(t_total=2.4)),
]


def test_duration_optimised_variants_match_oracle():
    """OptimizePhaseDurations on the other robots: ee-schedule sets, dense PhaseSpline pattern, total-duration rows."""
    for name in ("hopper", "biped_walk_stairs", "anymal_trot_block"):
        f = tb.make_formulation(name); f.params_.OptimizePhaseDurations()
        spec = f.to_spec()
        p = tb.Problem(spec); o = oracle_lib.Oracle(spec)
        assert (p.n, p.m, p.nnz) == (o.n, o.m, o.nnz)
        rp, ci = o.structure()
        assert np.array_equal(p.row_ptr(), rp) and np.array_equal(p.structure()[1], ci)
        for a, b in zip(p.bounds(), o.bounds()):
            assert np.array_equal(a, b)
        assert np.array_equal(p.GetVariableValues(), o.x0())
        assert p.variable_sets() == o.variable_sets() and p.constraint_sets() == o.constraint_sets()
    p4 = tb.Problem(tb.make_formulation("hyq_gallop_gap").to_spec())
    assert (p4.n, p4.m, p4.nnz) == (712, 930, 54516)   # SURVEY.md 8d / BASELINE.md config 4


@pytest.mark.parametrize("name,terrain,kw", CASES)
def test_structure_bounds_x0_match_oracle(name, terrain, kw):
    spec = tb.make_formulation(name, terrain=terrain, **kw).to_spec()
    p = tb.Problem(spec)
    o = oracle_lib.Oracle(spec)
    assert (p.n, p.m, p.nnz) == (o.n, o.m, o.nnz)
    rp, ci = o.structure()
    assert np.array_equal(p.row_ptr(), rp)          # bit-exact indices
    iRow, jCol = p.structure()
    assert np.array_equal(jCol, ci)
    assert np.array_equal(iRow, np.repeat(np.arange(o.m), np.diff(rp)))
    assert np.all(np.diff(jCol)[iRow[1:] == iRow[:-1]] > 0)   # ascending columns inside a row
    for a, b in zip(p.bounds(), o.bounds()):
        assert np.array_equal(a, b)
    assert np.array_equal(p.GetVariableValues(), o.x0())
    assert p.variable_sets() == o.variable_sets()
    assert p.constraint_sets() == o.constraint_sets()


def test_hand_verified_counts():
    # hopper n = 126+126+27+60, m = 10+132+57+57+81+50+12
    p = tb.Problem(tb.make_formulation("hopper").to_spec())
    assert [k for _, _, k in p.variable_sets()] == [126, 126, 27, 60]
    assert [k for _, _, k in p.constraint_sets()] == [10, 132, 57, 57, 81, 50, 12]
    assert [nm for nm, _, _ in p.constraint_sets()] == [
        "terrain-ee-motion_0", "dynamic", "splineacc-base-lin", "splineacc-base-ang", "rangeofmotion-0",
        "force-ee-force_0", "swing-ee-motion_0"]
    assert (p.n, p.m, p.nnz) == (339, 399, 5392)
    p2 = tb.Problem(tb.make_formulation("anymal_trot_block").to_spec())
    assert (p2.n, p2.m, p2.nnz) == (640, 892, 15096)
    p3 = tb.Problem(tb.make_formulation("biped_walk_stairs").to_spec())
    assert (p3.n, p3.m, p3.nnz) == (466, 586, 8709)


def test_gait_tables():
    # fly trot C1 and biped walk C0 scaled to T = 2.0
    gg = tb.GaitGenerator.MakeGaitGenerator(4); gg.SetCombo(1)
    lf = gg.GetPhaseDurations(2.0, 0); rf = gg.GetPhaseDurations(2.0, 1)
    assert np.allclose(lf, [0.35, 0.3, 0.2, 0.3, 0.2, 0.3, 0.35], atol=1e-15)
    assert np.allclose(rf, [0.15, 0.25, 0.2, 0.3, 0.2, 0.3, 0.2, 0.25, 0.15], atol=1e-15)
    assert np.allclose(gg.GetPhaseDurations(2.0, 3), lf, atol=0) and np.allclose(gg.GetPhaseDurations(2.0, 2), rf, atol=0)
    assert all(gg.IsInContactAtStart(e) for e in range(4))
    bg = tb.GaitGenerator.MakeGaitGenerator(2); bg.SetCombo(0)
    assert np.allclose(bg.GetPhaseDurations(2.0, 0), [0.125, 0.1875, 0.25, 0.1875, 0.25, 0.1875, 0.25, 0.1875, 0.375], atol=1e-15)
    assert np.allclose(bg.GetPhaseDurations(2.0, 1), [0.34375, 0.1875, 0.25, 0.1875, 0.25, 0.1875, 0.25, 0.1875, 0.15625], atol=1e-15)
    g4 = tb.GaitGenerator.MakeGaitGenerator(4); g4.SetCombo(4)
    assert all(len(g4.GetPhaseDurations(2.0, e)) == 9 for e in range(4))
    for e in range(4):
        assert abs(sum(g4.GetPhaseDurations(2.0, e)) - 2.0) < 1e-12


def test_sample_grid_quirks():
    # duplicated final sample at T = 2.0; 22 dynamic and 27 RoM samples
    p = tb.Problem(tb.make_formulation("anymal_trot_block").to_spec())
    rows = dict((nm, k) for nm, _, k in p.constraint_sets())
    assert rows["dynamic"] == 22 * 6 and rows["rangeofmotion-0"] == 27 * 3
    p24 = tb.Problem(tb.make_formulation("anymal_trot_block", t_total=2.4).to_spec())
    rows = dict((nm, k) for nm, _, k in p24.constraint_sets())
    assert rows["dynamic"] == 25 * 6
    assert (p24.n, p24.m, p24.nnz) == (688, 994, 17364)


def test_terrain_heights_match_oracle():
    rng = np.random.default_rng(0)
    pts = np.concatenate([rng.uniform(-1, 5, (200, 2)),
                          [[0.7, 0], [0.73, 0], [4.2, 0], [1.0, 0.1], [1.4, 0], [2.4, 0], [1.5, 0.2], [2.0, 0], [3.0, 0], [2.5, 1], [0.5, 0], [1.5, -1]]])
    for t in range(7):
        for x, y in pts:
            assert tb.terrain_height(t, x, y) == oracle_lib.lib().oracle_terrain_height(t, x, y)


def test_unsupported_and_invalid_specs():
    s = tb.make_formulation("hopper").to_spec(); s.n_ee = 2
    with pytest.raises(tb.TowrB200Error) as e:
        tb.Problem(s)
    assert e.value.code == capi.ERR_INVALID
    s = tb.make_formulation("hopper").to_spec(); s.constraints[0] = 99
    with pytest.raises(tb.TowrB200Error):
        tb.Problem(s)
    # numeric fields that drive loops / allocations: a zeroed or corrupted spec fails with TWB_ERR_INVALID (no hang, no crash)
    for field, bad in (("duration_base_polynomial", 0.0), ("dt_constraint_dynamic", 0.0), ("dt_constraint_range_of_motion", -0.08),
                       ("dt_constraint_base_motion", float("nan")), ("ee_polynomials_per_swing_phase", 0),
                       ("force_polynomials_per_stance_phase", 0), ("dt_constraint_dynamic", 1e-9), ("duration_base_polynomial", float("inf"))):
        s = tb.make_formulation("hopper").to_spec(); setattr(s, field, bad)
        with pytest.raises(tb.TowrB200Error) as e:
            tb.Problem(s)
        assert e.value.code == capi.ERR_INVALID, field
    s = tb.make_formulation("hopper").to_spec(); s.phase_durations[0][1] = 0.0
    with pytest.raises(tb.TowrB200Error):
        tb.Problem(s)
    zeroed = capi.Spec()                      # never passed through twb_spec_default
    with pytest.raises(tb.TowrB200Error):
        tb.Problem(zeroed)


def test_c_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "towr_b200.h")).read()
    declared = set(re.findall(r"\b(twb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = C.CDLL(capi.LIB_PATH)
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert declared == set(capi.EXPORTS)


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    p = tb.Problem(tb.make_formulation("hopper").to_spec())
    with pytest.raises(tb.TowrB200Error) as e:
        p.batch(4)
    assert e.value.code == capi.ERR_NO_DEVICE


def test_product_does_not_reference_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "towr_b200")):
        for fn in files:
            if fn.endswith((".py", ".cc", ".cu", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "oracle" not in txt.lower(), os.path.join(dirpath, fn)


def test_goal_randomised_instances_match_oracle():
    """BASELINE configs[2]: x0 / bounds of goal-randomised instances of one structure class, without re-building it."""
    rng = np.random.default_rng(8)
    for name in ("biped_walk_stairs", "anymal_trot_block", "hyq_gallop_gap"):
        base = tb.make_formulation(name)
        p = tb.Problem(base.to_spec())
        G = 5
        goals = np.zeros((G, 6))
        goals[:, 0] = rng.uniform(0.5, 2.5, G); goals[:, 1] = rng.uniform(-0.3, 0.3, G)
        goals[:, 2] = base.final_base_.lin.p[2]; goals[:, 5] = rng.uniform(-0.3, 0.3, G)
        x0, xl, xu = p.goal_instances(goals)
        for i in range(G):
            f = tb.make_formulation(name, goal_xy=(goals[i, 0], goals[i, 1]), goal_yaw=goals[i, 5])
            o = oracle_lib.Oracle(f.to_spec())
            oxl, oxu, _, _ = o.bounds()
            assert np.array_equal(x0[i], o.x0()) and np.array_equal(xl[i], oxl) and np.array_equal(xu[i], oxu)
    # the spec's own goal reproduces the problem's x0 / bounds
    fb = base.final_base_
    own = np.array([[*fb.lin.p, *fb.ang.p]])
    x0, xl, xu = p.goal_instances(own)
    assert np.array_equal(x0[0], p.GetVariableValues()) and np.array_equal(xl[0], p.bounds()[0])


def test_bench_spec_fixture_is_current():
    """tests/golden/bench_specs.json (what bench.py's reference arm evaluates without loading the product library) equals
    the recipes of towr_b200.configs."""
    import json
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "bench_specs.json")))
    for name, hexbytes in fx.items():
        assert bytes(tb.make_formulation(name).to_spec()).hex() == hexbytes, name


def test_trajectory_dims_match_oracle():
    for name, dt in (("hopper", 0.05), ("anymal_trot_block", 0.01), ("hyq_gallop_gap", 0.1)):
        spec = tb.make_formulation(name).to_spec(); p = tb.Problem(spec)
        ns, nv = C.c_int(), C.c_int()
        capi.check(capi.lib.twb_problem_trajectory_dims(p._h, dt, C.byref(ns), C.byref(nv)))
        ref = oracle_lib.Oracle(spec).trajectory(p.GetVariableValues(), dt)
        assert (ns.value, nv.value) == ref.shape
