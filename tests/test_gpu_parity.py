"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the
CPU oracle on the same seeded inputs; plus size-independent properties at the
full BASELINE batch sizes."""
import os

import numpy as np
import pytest

import towr_b200 as tb
from towr_b200 import capi
from towr_b200.configs import synthetic_iterates, synthetic_iterates_fast
import oracle_lib
from tolerance import check_rows, check_sets

pytestmark = pytest.mark.gpu

# every comparison against the oracle is recorded (worst error / tolerance, fraction of entries that miss the BARE
# 1e-12 relative / 1e-14 absolute criterion) and written to gpurun_out/parity.json at the end of the module; the copy of
# the last GPU run of a round is committed as profiles/parity_r<round>.json
PARITY_LOG = []


@pytest.fixture(scope="module", autouse=True)
def _write_parity_log():
    yield
    import json
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, "parity.json"), "w") as fh:
            json.dump({"tolerance": {"rel": 1e-12, "abs": 1e-14, "abs_scaled_by": "largest |ref| of the Jacobian row / constraint set"},
                       "cases": PARITY_LOG}, fh, indent=1)
    except OSError:
        pass


def _compare(name, B, terrain=None, terrains=None, costs=None, strict=1e-4, **kw):
    f = tb.make_formulation(name, terrain=terrain, **kw)
    if costs:
        f.params_.costs_ = costs
    spec = f.to_spec()
    p = tb.Problem(spec)
    X = synthetic_iterates(p, B)
    bt = p.batch(B)
    if terrains is not None:
        bt.set_terrains(terrains)
    flags = capi.EVAL_ALL if costs else capi.EVAL_G | capi.EVAL_JAC
    out = bt.eval_host(X, flags=flags)
    ref = oracle_lib.batch_eval(spec, X, terrain_ids=terrains, want_cost=bool(costs))
    assert ref["rc"] == 0                      # oracle pattern at x == pattern at x0
    bad_j, strict_j, worst_j = check_rows(out["jac"], ref["jac"], p.row_ptr())
    bad_g, strict_g, worst_g = check_sets(out["g"], ref["g"], p.constraint_sets())
    PARITY_LOG.append({"case": name, "B": B, "n": p.n, "m": p.m, "nnz": p.nnz, "kw": {k: str(v) for k, v in kw.items()},
                       "terrain": None if terrain is None else int(terrain), "mixed_terrains": terrains is not None,
                       "jac": {"violations": bad_j, "worst_err_over_tol": worst_j, "strict_miss_fraction": strict_j},
                       "g": {"violations": bad_g, "worst_err_over_tol": worst_g, "strict_miss_fraction": strict_g},
                       "strict_miss_limit": strict})
    assert bad_j == 0, (bad_j, worst_j)
    assert bad_g == 0, (bad_g, worst_g)
    assert strict_j < strict and strict_g < strict   # entries missing the bare (unscaled) 1e-12/1e-14 criterion are rare
    assert not out["status"].any()
    return p, out, ref


def test_hopper_config1():
    _compare("hopper", 32)


def test_anymal_trot_block_config2():
    _compare("anymal_trot_block", 64)


def test_biped_walk_stairs_config3():
    _compare("biped_walk_stairs", 64, goal_xy=(2.0, 0.2), goal_yaw=0.3)


def test_anymal_mixed_terrains_config5():
    B = 63                                      # ragged: not a multiple of the CTA group size
    terr = np.array([tb.SLOPE, tb.CHIMNEY, tb.GAP] * 21, np.int32)
    _compare("anymal_trot_mixed", B, terrains=terr)


def test_hyq_gallop_gap_durations_config4():
    """Phase durations optimised: PhaseSpline / PhaseDurations Jacobians, TotalDurationConstraint."""
    # optimised durations enter as T^-2 .. T^-4 in front of cancelling sums; the device computes the powers correctly
    # rounded (Pow3 / Pow4 in kernels.cu) like std::pow, so the bare criterion holds as for the other configs
    p, out, ref = _compare("hyq_gallop_gap", 48, strict=1e-4)
    (_, r0, nr), = [c for c in p.constraint_sets() if c[0] == "totalduration-2"]
    assert np.array_equal(out["g"][:, r0], ref["g"][:, r0])


def test_durations_spread_widely_over_a_tile():
    """The instances of one tile of 32 fall into polynomials / phases far apart (durations x U[0.35, 1.65]): the phase
    elements then take the table form and the per-instance duration columns instead of the window form (device_tables.h)."""
    from towr_b200.configs import _renormalise
    for name, B in (("hyq_gallop_gap", 64), ("biped_walk_stairs", 37)):
        f = tb.make_formulation(name)
        f.params_.OptimizePhaseDurations()
        spec = f.to_spec(); p = tb.Problem(spec)
        X = synthetic_iterates(p, B)
        x0 = p.GetVariableValues()
        rng = np.random.default_rng(99)
        for sname, start, count in p.variable_sets():
            if sname.startswith("ee-schedule"):
                X[:, start:start + count] = _renormalise(p, sname, x0[start:start + count] * rng.uniform(0.35, 1.65, (B, count)))
        out = p.batch(B).eval_host(X)
        ref = oracle_lib.batch_eval(spec, X)
        assert ref["rc"] == 0
        bad_j, strict_j, worst_j = check_rows(out["jac"], ref["jac"], p.row_ptr())
        bad_g, strict_g, worst_g = check_sets(out["g"], ref["g"], p.constraint_sets())
        PARITY_LOG.append({"case": name + " (durations x U[0.35, 1.65])", "B": B, "n": p.n, "m": p.m, "nnz": p.nnz,
                           "jac": {"violations": bad_j, "worst_err_over_tol": worst_j, "strict_miss_fraction": strict_j},
                           "g": {"violations": bad_g, "worst_err_over_tol": worst_g, "strict_miss_fraction": strict_g}})
        assert bad_j == 0 and bad_g == 0, (name, bad_j, worst_j, bad_g, worst_g)
        assert not out["status"].any()


def test_durations_other_robots_and_terrains():
    f = tb.make_formulation("hopper"); f.params_.OptimizePhaseDurations()
    spec = f.to_spec(); p = tb.Problem(spec)
    X = synthetic_iterates(p, 33)
    out = p.batch(33).eval_host(X)
    ref = oracle_lib.batch_eval(spec, X)
    assert ref["rc"] == 0
    assert check_rows(out["jac"], ref["jac"], p.row_ptr())[0] == 0
    assert check_sets(out["g"], ref["g"], p.constraint_sets())[0] == 0
    f = tb.make_formulation("biped_walk_stairs", terrain=tb.SLOPE); f.params_.OptimizePhaseDurations()
    spec = f.to_spec(); p = tb.Problem(spec)
    X = synthetic_iterates(p, 40)
    X[3] = p.GetVariableValues()                 # nominal durations: samples sit exactly on phase junctions
    out = p.batch(40).eval_host(X)
    ref = oracle_lib.batch_eval(spec, X)
    assert check_rows(out["jac"], ref["jac"], p.row_ptr())[0] == 0
    assert check_sets(out["g"], ref["g"], p.constraint_sets())[0] == 0
    X[5, p.n - 3:] = 5.0                         # durations summing to more than T: flagged, not fatal
    out = p.batch(40).eval_host(X)
    assert out["status"][5] & 2 and not (out["status"][4] & 2)


def test_every_terrain_go1():
    B = 28
    _compare("go1_trot_flat", B, terrains=(np.arange(B) % 7).astype(np.int32))


def test_base_motion_constraint():
    _compare("anymal_trot_block_base_rom", 40)
    _compare("hopper_base_rom", 5)


def test_longer_horizon():
    _compare("anymal_trot_block", 8, t_total=2.4)


def test_node_costs():
    p, out, ref = _compare("anymal_trot_block", 16, costs=[(capi.COST_FORCES, 1.0), (capi.COST_EE_MOTION, 0.5)])
    assert np.allclose(out["cost"], ref["cost"], rtol=1e-12, atol=0)
    assert np.allclose(out["grad"], ref["grad"], rtol=1e-12, atol=1e-14)
    assert np.abs(out["grad"]).max() > 0


def test_batch_of_one_and_flags():
    f = tb.make_formulation("hopper"); spec = f.to_spec(); p = tb.Problem(spec)
    X = synthetic_iterates(p, 1)
    bt = p.batch(1)
    full = bt.eval_host(X)
    only_g = bt.eval_host(X, flags=capi.EVAL_G)
    only_j = bt.eval_host(X, flags=capi.EVAL_JAC)
    assert only_g["jac"] is None and np.array_equal(only_g["g"], full["g"])
    assert only_j["g"] is None and np.array_equal(only_j["jac"], full["jac"])
    zero = bt.eval_host(X, flags=capi.EVAL_ALL)       # no cost terms: cost 0, gradient 0
    assert zero["cost"][0] == 0.0 and not zero["grad"].any()


def test_x0_iterate_and_status_flag():
    f = tb.make_formulation("anymal_trot_block"); spec = f.to_spec(); p = tb.Problem(spec)
    X = np.tile(p.GetVariableValues(), (4, 1))
    out = p.batch(4).eval_host(X)
    ref = oracle_lib.batch_eval(spec, X)
    assert check_rows(out["jac"], ref["jac"], p.row_ptr())[0] == 0
    assert check_sets(out["g"], ref["g"], p.constraint_sets())[0] == 0
    X[2, 5] = np.nan
    out = p.batch(4).eval_host(X)
    assert out["status"][2] & 1


def test_device_pointer_variant_matches_host_variant():
    import torch
    f = tb.make_formulation("anymal_trot_block"); p = tb.Problem(f.to_spec())
    B = 33
    X = synthetic_iterates(p, B)
    bt = p.batch(B)
    host = bt.eval_host(X)
    xd = torch.from_numpy(X).cuda()
    g = torch.empty((B, p.m), dtype=torch.float64, device="cuda")
    jac = torch.empty((B, p.nnz), dtype=torch.float64, device="cuda")
    st = torch.empty(B, dtype=torch.int32, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        bt.eval_device(xd, g=g, jac=jac, status=st, stream=s)
    s.synchronize()
    assert np.array_equal(g.cpu().numpy(), host["g"]) and np.array_equal(jac.cpu().numpy(), host["jac"])


def test_device_variant_graph_cache_and_chunks():
    """twb_batch_eval_device replays a CUDA graph per argument set (16 cached, least recently used replaced, dropped when the
    terrains change) and walks batches of 1.5 - 3.5 x 4096 instances in chunks: 20 argument sets in rotation, a terrain change
    in between, the legacy default stream, a 6144-instance batch — every result equals the host-pointer variant's bit for bit."""
    import torch
    f = tb.make_formulation("anymal_trot_block"); p = tb.Problem(f.to_spec())
    B = 70
    bt = p.batch(B)
    Xs = [synthetic_iterates_fast(p, B, seed=500 + i) for i in range(20)]
    want = [bt.eval_host(X) for X in Xs]
    xd = [torch.from_numpy(X).cuda() for X in Xs]
    outs = [(torch.empty((B, p.m), dtype=torch.float64, device="cuda"), torch.empty((B, p.nnz), dtype=torch.float64, device="cuda"),
             torch.empty(B, dtype=torch.int32, device="cuda")) for _ in range(3)]
    for rounds in range(2):                      # second round: sets 0 .. 3 were evicted from the cache and are captured again
        for i in range(20):
            g, jac, st = outs[i % 3]
            bt.eval_device(xd[i], g=g, jac=jac, status=st)          # current (legacy default) stream
            torch.cuda.synchronize()
            assert np.array_equal(g.cpu().numpy(), want[i]["g"]) and np.array_equal(jac.cpu().numpy(), want[i]["jac"])
    terr = np.full(B, tb.SLOPE, dtype=np.int32); terr[::3] = tb.GAP
    bt.set_terrains(terr)                        # cached graphs hold the old terrain pointer: must be dropped
    ref = bt.eval_host(Xs[0])
    g, jac, st = outs[0]
    bt.eval_device(xd[0], g=g, jac=jac, status=st)
    torch.cuda.synchronize()
    assert np.array_equal(jac.cpu().numpy(), ref["jac"]) and not np.array_equal(ref["jac"], want[0]["jac"])
    g_only = torch.zeros_like(g)
    bt.eval_device(xd[0], g=g_only, status=st, flags=capi.EVAL_G)   # same pointers, other flags: another graph
    torch.cuda.synchronize()
    assert np.array_equal(g_only.cpu().numpy(), ref["g"])
    # chunked evaluation: 6144 instances = 1.5 chunks
    B2 = 6144
    X2 = synthetic_iterates_fast(p, B2, seed=9)
    b2 = p.batch(B2)
    h2 = b2.eval_host(X2)
    x2 = torch.from_numpy(X2).cuda()
    g2 = torch.empty((B2, p.m), dtype=torch.float64, device="cuda"); j2 = torch.empty((B2, p.nnz), dtype=torch.float64, device="cuda")
    s2 = torch.empty(B2, dtype=torch.int32, device="cuda")
    b2.eval_device(x2, g=g2, jac=j2, status=s2)
    torch.cuda.synchronize()
    assert np.array_equal(g2.cpu().numpy(), h2["g"]) and np.array_equal(j2.cpu().numpy(), h2["jac"]) and not s2.any().item()
    assert b2.launches_per_eval() == 2 * bt.launches_per_eval()


@pytest.mark.parametrize("name,B,durations", [("anymal_trot_block", 45, False), ("biped_walk_stairs", 101, False), ("hyq_gallop_gap", 37, True),
                                               ("biped_walk_stairs", 70, True), ("hopper", 3, False)])
def test_no_write_outside_the_output_arrays(name, B, durations):
    """compute-sanitizer is not available on this GPU pool (profiles/r2_compute_sanitizer_refused.txt): every output array of the
    device-pointer variant sits between two guard zones filled with a sentinel; after the evaluation the guards are intact and
    every element between them has been written (ragged batches, odd row lengths = interleaved tiles, optimised durations)."""
    import torch
    f = tb.make_formulation(name)
    if durations:
        f.params_.OptimizePhaseDurations()
    p = tb.Problem(f.to_spec())
    X = torch.from_numpy(synthetic_iterates_fast(p, B, seed=5)).cuda()
    bt = p.batch(B)
    guard, sentinel = 4096, -7.25e300

    def guarded(count, dtype, fill):
        buf = torch.full((count + 2 * guard,), fill, dtype=dtype, device="cuda")
        return buf, buf[guard:guard + count]
    gb, g = guarded(B * p.m, torch.float64, sentinel)
    jb, jac = guarded(B * p.nnz, torch.float64, sentinel)
    sb, st = guarded(B, torch.int32, 12345)
    bt.eval_device(X, g=g, jac=jac, status=st)
    torch.cuda.synchronize()
    for buf, count, fill in ((gb, B * p.m, sentinel), (jb, B * p.nnz, sentinel), (sb, B, 12345)):
        assert bool((buf[:guard] == fill).all()) and bool((buf[guard + count:] == fill).all()), "write outside the array"
        assert not bool((buf[guard:guard + count] == fill).any()), "element left unwritten"
    host = bt.eval_host(X.cpu().numpy())
    assert np.array_equal(jac.cpu().numpy().reshape(B, p.nnz), host["jac"]) and np.array_equal(g.cpu().numpy().reshape(B, p.m), host["g"])


def test_full_size_properties_config2():
    """BASELINE configs[1] at full size (4096): determinism, instance independence (permutation
    equivariance), iterate-independent entries constant across the batch, and oracle parity on a
    seeded subsample."""
    f = tb.make_formulation("anymal_trot_block"); spec = f.to_spec(); p = tb.Problem(spec)
    B = 4096
    X = synthetic_iterates_fast(p, B)
    bt = p.batch(B)
    a = bt.eval_host(X)
    b = bt.eval_host(X)
    assert np.array_equal(a["jac"], b["jac"]) and np.array_equal(a["g"], b["g"])
    perm = np.random.default_rng(7).permutation(B)
    c = bt.eval_host(X[perm])
    assert np.array_equal(c["jac"], a["jac"][perm]) and np.array_equal(c["g"], a["g"][perm])
    # SplineAcc / Swing Jacobians and the "1.0" terrain entries do not depend on x
    rp = p.row_ptr()
    for name, r0, nr in p.constraint_sets():
        if name.startswith(("splineacc", "swing")):
            blk = a["jac"][:, rp[r0]:rp[r0 + nr]]
            assert np.all(blk == blk[0])
    idx = np.random.default_rng(11).choice(B, 48, replace=False)
    ref = oracle_lib.batch_eval(spec, X[idx])
    assert check_rows(a["jac"][idx], ref["jac"], rp)[0] == 0
    assert check_sets(a["g"][idx], ref["g"], p.constraint_sets())[0] == 0
    assert not a["status"].any()


def test_linearity_in_forces():
    """g_dynamic is affine in the force variables: g(x + 2d) - g(x + d) == g(x + d) - g(x) for force-only d,
    and the Jacobian reproduces the directional derivative."""
    f = tb.make_formulation("anymal_trot_block"); p = tb.Problem(f.to_spec())
    X = synthetic_iterates(p, 4)
    d = np.zeros_like(X)
    for name, s, k in p.variable_sets():
        if name.startswith("ee-force"):
            d[:, s:s + k] = np.random.default_rng(3).standard_normal((4, k))
    bt = p.batch(4)
    g0, g1, g2 = (bt.eval_host(X + a * d) for a in (0.0, 1.0, 2.0))
    (_, r0, nr), = [c for c in p.constraint_sets() if c[0] == "dynamic"]
    lhs = g2["g"][:, r0:r0 + nr] - g1["g"][:, r0:r0 + nr]
    rhs = g1["g"][:, r0:r0 + nr] - g0["g"][:, r0:r0 + nr]
    assert np.allclose(lhs, rhs, rtol=0, atol=1e-9)
    iRow, jCol = p.structure()
    for b in range(4):
        J = np.zeros((p.m, p.n)); J[iRow, jCol] = g0["jac"][b]
        assert np.allclose((J @ d[b])[r0:r0 + nr], rhs[b], rtol=0, atol=1e-9)


def _device_eval(p, bt, X, terrains=None):
    """Device-resident evaluation (no 10+ GB host copies): returns torch tensors g, jac, status."""
    import torch
    B = X.shape[0]
    xd = torch.from_numpy(X).cuda()
    g = torch.empty((B, p.m), dtype=torch.float64, device="cuda")
    jac = torch.empty((B, p.nnz), dtype=torch.float64, device="cuda")
    st = torch.empty(B, dtype=torch.int32, device="cuda")
    bt.eval_device(xd, g=g, jac=jac, status=st)
    torch.cuda.synchronize()
    return g, jac, st


def test_full_size_properties_config3_biped_16384():
    """BASELINE configs[2] at full size: determinism, instance independence, oracle parity on a seeded subsample
    (odd nnz: four alignment classes of the output rows)."""
    f = tb.make_formulation("biped_walk_stairs"); spec = f.to_spec(); p = tb.Problem(spec)
    B = 16384
    rng = np.random.default_rng(16384)                    # goal x in U[0.5, 2.5], y, yaw in U[-0.3, 0.3]
    goals = np.zeros((B, 6))
    goals[:, 0] = rng.uniform(0.5, 2.5, B); goals[:, 1] = rng.uniform(-0.3, 0.3, B)
    goals[:, 2] = f.final_base_.lin.p[2]; goals[:, 5] = rng.uniform(-0.3, 0.3, B)
    x0s, xl, xu = p.goal_instances(goals)                 # per-instance initial guess / bounds, one structure class
    assert np.all(xl <= xu) and np.ptp(x0s[:, 120]) > 0.5    # (the reference's interpolated guess need not satisfy the bounds)
    X = synthetic_iterates_fast(p, B, x0=x0s)
    bt = p.batch(B)
    g1, j1, s1 = _device_eval(p, bt, X)
    g2, j2, s2 = _device_eval(p, bt, X)
    assert bool((j1 == j2).all()) and bool((g1 == g2).all()) and int(s1.sum()) == 0
    perm = np.random.default_rng(3).permutation(B)
    g3, j3, _ = _device_eval(p, bt, X[perm])
    import torch
    pt = torch.from_numpy(perm).cuda()
    assert bool((j3 == j1[pt]).all()) and bool((g3 == g1[pt]).all())
    idx = np.sort(np.random.default_rng(5).choice(B, 40, replace=False))
    ref = oracle_lib.batch_eval(spec, X[idx])
    it = torch.from_numpy(idx).cuda()
    assert check_rows(j1[it].cpu().numpy(), ref["jac"], p.row_ptr())[0] == 0
    assert check_sets(g1[it].cpu().numpy(), ref["g"], p.constraint_sets())[0] == 0


def test_full_size_properties_config4_hyq_32768():
    """BASELINE configs[3] at full size (14 GB of Jacobian values, device-resident): determinism, oracle parity on a
    seeded subsample, total-duration rows = sum of the duration variables, no status flags."""
    import torch
    f = tb.make_formulation("hyq_gallop_gap"); spec = f.to_spec(); p = tb.Problem(spec)
    B = 32768
    X = synthetic_iterates_fast(p, B)
    bt = p.batch(B)
    g1, j1, s1 = _device_eval(p, bt, X)
    assert int(s1.sum()) == 0
    chk1 = (float(j1.sum()), float(g1.sum()), float(j1.abs().max()))
    g1c = g1.clone(); del g1
    idx = np.sort(np.random.default_rng(9).choice(B, 24, replace=False))
    it = torch.from_numpy(idx).cuda()
    jsub, gsub = j1[it].cpu().numpy(), g1c[it].cpu().numpy()
    g2, j2, _ = _device_eval(p, bt, X)                 # same buffers' worth of work again
    assert chk1 == (float(j2.sum()), float(g2.sum()), float(j2.abs().max()))
    ref = oracle_lib.batch_eval(spec, X[idx])
    assert ref["rc"] == 0
    assert check_rows(jsub, ref["jac"], p.row_ptr())[0] == 0
    assert check_sets(gsub, ref["g"], p.constraint_sets())[0] == 0
    sets = dict((nm, (s, k)) for nm, s, k in p.variable_sets())
    for ee in range(4):
        (_, r0, _), = [c for c in p.constraint_sets() if c[0] == f"totalduration-{ee}"]
        s, k = sets[f"ee-schedule{ee}"]
        assert np.allclose(g1c[:, r0].cpu().numpy(), X[:, s:s + k].sum(axis=1), rtol=1e-15, atol=0)


def test_full_size_properties_config5_shard_mixed_terrains():
    """BASELINE configs[4]: one GPU's shard (65536 / 8 instances) with terrains drawn per instance."""
    import torch
    f = tb.make_formulation("anymal_trot_mixed"); spec = f.to_spec(); p = tb.Problem(spec)
    B = 65536 // 8
    X = synthetic_iterates_fast(p, B, seed=77)
    terr = np.random.default_rng(7).choice([tb.SLOPE, tb.CHIMNEY, tb.GAP], B).astype(np.int32)
    bt = p.batch(B); bt.set_terrains(terr)
    g1, j1, s1 = _device_eval(p, bt, X)
    assert int(s1.sum()) == 0
    idx = np.sort(np.random.default_rng(13).choice(B, 48, replace=False))
    ref = oracle_lib.batch_eval(spec, X[idx], terrain_ids=terr[idx])
    it = torch.from_numpy(idx).cuda()
    assert check_rows(j1[it].cpu().numpy(), ref["jac"], p.row_ptr())[0] == 0
    assert check_sets(g1[it].cpu().numpy(), ref["g"], p.constraint_sets())[0] == 0
    # the terrain only enters the terrain / force rows: everything else equals the single-terrain evaluation
    bt2 = p.batch(B)
    g0, j0, _ = _device_eval(p, bt2, X)
    rp = p.row_ptr()
    for name, r0, nr in p.constraint_sets():
        if not name.startswith(("terrain", "force")):
            assert bool((j1[:, rp[r0]:rp[r0 + nr]] == j0[:, rp[r0]:rp[r0 + nr]]).all()), name


def test_csv_grid_terrain():
    """towr::HeightMapFromCSV as per-batch terrain data (TWB_GRID_CSV), mixed with analytic terrains per instance."""
    rng = np.random.default_rng(21)
    grid = rng.integers(0, 4, (12, 20)) * 0.04
    B = 45
    terr = np.where(np.arange(B) % 3 == 2, tb.BLOCK, tb.GRID_CSV).astype(np.int32)
    f = tb.make_formulation("anymal_trot_block"); spec = f.to_spec(); p = tb.Problem(spec)
    X = synthetic_iterates(p, B)
    # put some feet right into the edge bands of the grid (one-sided slopes)
    for name, s, k in p.variable_sets():
        if name == "ee-motion_0":
            X[0:8, s] = np.arange(1, 9) * 0.17 - 0.001
            X[8:16, s] = np.arange(1, 9) * 0.17 + 0.001
        if name == "ee-motion_1":                # coordinates in (-res, 0): cell 0 of the reference (size_t truncation), not "outside"
            X[16:24, s] = -0.05 - 0.01 * np.arange(8); X[20:28, s + 1] = -0.02 - 0.015 * np.arange(8)
    grid[0, :] += 0.04; grid[:, 0] += 0.04
    bt = p.batch(B); bt.set_terrains(terr); bt.set_grid_terrain(grid)
    out = bt.eval_host(X)
    oracle_lib.set_grid(grid)
    ref = oracle_lib.batch_eval(spec, X, terrain_ids=terr)
    assert ref["rc"] == 0
    assert check_rows(out["jac"], ref["jac"], p.row_ptr())[0] == 0
    assert check_sets(out["g"], ref["g"], p.constraint_sets())[0] == 0
    assert not out["status"].any()
    rp = p.row_ptr()
    (_, r0, nr), = [c for c in p.constraint_sets() if c[0] == "terrain-ee-motion_0"]
    blk = out["jac"][:16, rp[r0]:rp[r0 + nr]]
    assert np.abs(blk).max() > 1.5               # an edge slope (0.04 / (0.17 / 50) = 11.8) showed up in the terrain rows
    bt.set_grid_terrain(None)                    # without a grid every height is 0
    flat = bt.eval_host(X)
    oracle_lib.set_grid(np.zeros((1, 1)))
    ref0 = oracle_lib.batch_eval(spec, X, terrain_ids=terr)
    assert check_sets(flat["g"], ref0["g"], p.constraint_sets())[0] == 0


def test_grid_map_terrain():
    """towr `Grid` (grid_height_map.h:16-59: grid_map elevation layer, bilinear float heights, eps = res / 6 slopes) as
    per-batch terrain data (TWB_GRID_MAP), mixed per instance with analytic terrains; terrain and force rows against the oracle."""
    rng = np.random.default_rng(23)
    sx, sy, res, pos = 60, 40, 0.1, (1.0, 0.0)
    xs = pos[0] + sx * res / 2 - res / 2 - res * np.arange(sx)
    ys = pos[1] + sy * res / 2 - res / 2 - res * np.arange(sy)
    H = (0.15 * np.sin(2.0 * xs)[:, None] * np.cos(1.5 * ys)[None, :] + 0.02 * rng.standard_normal((sx, sy))).astype(np.float32)
    B = 40
    terr = np.where(np.arange(B) % 4 == 3, tb.SLOPE, tb.GRID_MAP).astype(np.int32)
    f = tb.make_formulation("anymal_trot_block"); spec = f.to_spec(); p = tb.Problem(spec)
    X = synthetic_iterates(p, B)
    bt = p.batch(B); bt.set_terrains(terr); bt.set_grid_map(H, res, pos)
    out = bt.eval_host(X)
    oracle_lib.set_grid_map(H, res, pos)
    ref = oracle_lib.batch_eval(spec, X, terrain_ids=terr)
    assert ref["rc"] == 0
    bad_j, strict_j, worst_j = check_rows(out["jac"], ref["jac"], p.row_ptr())
    bad_g, strict_g, worst_g = check_sets(out["g"], ref["g"], p.constraint_sets())
    PARITY_LOG.append({"case": "anymal_trot_block+grid_map", "B": B, "jac": {"violations": bad_j, "worst_err_over_tol": worst_j, "strict_miss_fraction": strict_j},
                       "g": {"violations": bad_g, "worst_err_over_tol": worst_g, "strict_miss_fraction": strict_g}})
    assert bad_j == 0 and bad_g == 0, (bad_j, bad_g, worst_j, worst_g)
    assert not out["status"].any()
    rp = p.row_ptr()
    (_, r0, nr), = [c for c in p.constraint_sets() if c[0] == "terrain-ee-motion_0"]
    assert np.abs(out["jac"][:3, rp[r0]:rp[r0 + nr]]).max() > 0.05      # bilinear slopes showed up in the terrain rows
    flat = p.batch(B); flat.set_terrains(terr)                          # without a layer every grid-map height is FLT_MAX: flagged, not silently flat
    res0 = flat.eval_host(X)
    assert (res0["status"][terr == tb.GRID_MAP] & 1).all() or np.abs(res0["g"][terr == tb.GRID_MAP]).max() > 1e30


def test_trajectory_sampling_matches_oracle():
    """fpowr::GetTrajectory (footstep_plan_extractor.h:19-53) batched on the device: splines, quaternion, angular
    velocity / acceleration, contact flags — against the oracle, for fixed and for optimised phase durations."""
    for name, B, dt in (("hopper", 5, 0.05), ("anymal_trot_block", 33, 0.01), ("hyq_gallop_gap", 12, 0.02)):
        spec = tb.make_formulation(name).to_spec(); p = tb.Problem(spec)
        X = synthetic_iterates(p, B)
        got = p.batch(B).sample_trajectory(X, dt)
        o = oracle_lib.Oracle(spec)
        for b in range(B):
            ref = o.trajectory(X[b], dt)
            assert got[b].shape == ref.shape
            scale = np.maximum(1.0, np.abs(ref).max(axis=0, keepdims=True))
            assert np.all(np.abs(got[b] - ref) <= 1e-12 * np.abs(ref) + 1e-13 * scale), (name, b, np.abs(got[b] - ref).max())
        n_ee = (got.shape[2] - 19) // 13
        assert np.allclose((got[..., 9:13] ** 2).sum(-1), 1.0, atol=1e-12)          # unit quaternions
        assert set(np.unique(got[..., [19 + 13 * e for e in range(n_ee)]])) <= {0.0, 1.0}


def test_initial_guess_extraction_matches_oracle():
    """fpowr::ExtractInitialGuesses (initial_guess_extractor.h:17-48) batched on the device — time | state[12] |
    controls[36] at caller-given sample times — against the oracle, for fixed and for optimised phase durations."""
    times = np.array([0.0, 0.013, 0.4, 0.99, 1.0, 1.37, 2.0])
    for name, B in (("hopper", 5), ("anymal_trot_block", 33), ("biped_walk_stairs", 7), ("hyq_gallop_gap", 12)):
        spec = tb.make_formulation(name).to_spec(); p = tb.Problem(spec)
        X = synthetic_iterates(p, B)
        got = p.batch(B).initial_guesses(X, times)
        o = oracle_lib.Oracle(spec)
        for b in range(B):
            ref = o.initial_guesses(X[b], times)
            scale = np.maximum(1.0, np.abs(ref).max(axis=0, keepdims=True))
            assert np.all(np.abs(got[b] - ref) <= 1e-12 * np.abs(ref) + 1e-13 * scale), (name, b, np.abs(got[b] - ref).max())
        assert np.all(got[..., 0] == times) and np.all(got[..., 25:37] == 0.0)      # joint torques are zero


def test_footstep_plan_extraction_matches_oracle():
    """fpowr::ExtractFootstepPlan (footstep_plan_extractor.h:55-133, without the nearest-plane lookup): contact-change
    scan of the 0.01 s trajectory — footstep times and contact sets identical, durations and positions within 1e-12."""
    for name, B in (("hopper", 5), ("anymal_trot_block", 33), ("hyq_gallop_gap", 12)):
        spec = tb.make_formulation(name).to_spec(); p = tb.Problem(spec)
        X = synthetic_iterates(p, B)
        T = 2.0
        plans = p.batch(B).footstep_plans(X, T)
        o = oracle_lib.Oracle(spec)
        for b in range(B):
            ref = o.footstep_plan(X[b], T)
            got = plans[b]
            assert got.shape == ref.shape and got.shape[0] >= 2, (name, b, got.shape, ref.shape)
            n_ee = (got.shape[1] - 2) // 4
            flags = [2 + 4 * e for e in range(n_ee)]
            assert np.array_equal(got[:, 0], ref[:, 0]) and np.array_equal(got[:, flags], ref[:, flags])
            assert np.all(np.abs(got - ref) <= 1e-12 * np.abs(ref) + 1e-13), (name, b, np.abs(got - ref).max())
            assert abs(got[:, 1].sum() - T) < 1e-9                                     # durations tile the horizon


def test_goal_instances_on_the_device():
    """Batched setup: x0 / variable bounds of goal-randomised instances from a kernel, against the host
    builder (twb_problem_goal_instances, itself bit-equal to the oracle in tests/test_structure.py) — every robot class,
    optimised durations included; then with a height grid as the instances' terrain, which only the device path serves."""
    import torch
    rng = np.random.default_rng(31)
    for name in ("hopper", "anymal_trot_block", "biped_walk_stairs", "hyq_gallop_gap", "anymal_trot_block_base_rom"):
        p = tb.Problem(tb.make_formulation(name).to_spec())
        B = 70
        goals = np.column_stack([rng.uniform(0.3, 2.5, B), rng.uniform(-0.4, 0.4, B), np.full(B, 0.5), np.zeros(B), np.zeros(B), rng.uniform(-0.5, 0.5, B)])
        x0h, xlh, xuh = p.goal_instances(goals)
        bt = p.batch(B)
        x0, lo, up = bt.goal_instances_device(torch.from_numpy(goals).cuda())
        assert np.allclose(x0.cpu().numpy(), x0h, rtol=1e-14, atol=1e-15), name    # sincos(yaw) may differ from the host's libm in the last place
        assert np.allclose(lo.cpu().numpy(), xlh, rtol=1e-14, atol=0) and np.allclose(up.cpu().numpy(), xuh, rtol=1e-14, atol=0), name
    # the CSV height grid under the goals: final base z and footholds follow the grid
    p = tb.Problem(tb.make_formulation("anymal_trot_block").to_spec())
    B = 33
    grid = rng.integers(0, 4, (12, 20)) * 0.04
    bt = p.batch(B); bt.set_terrains(np.full(B, tb.GRID_CSV, np.int32)); bt.set_grid_terrain(grid)
    goals = np.column_stack([rng.uniform(0.5, 2.5, B), rng.uniform(0.3, 1.0, B), np.full(B, 0.5), np.zeros(B), np.zeros(B), np.zeros(B)])
    x0, _, _ = bt.goal_instances_device(torch.from_numpy(goals).cuda(), want_bounds=False)
    x0 = x0.cpu().numpy()
    oracle_lib.set_grid(grid)
    (_, s_lin, n_lin), = [v for v in p.variable_sets() if v[0] == "base-lin"]
    nominal_z = tb.robot_info(tb.ANYMAL)["nominal_stance"][0][2]
    for b in range(B):
        h = oracle_lib.terrain_point(tb.GRID_CSV, goals[b, 0], goals[b, 1])[0]
        assert abs(x0[b, s_lin + n_lin - 6 + 2] - (h - nominal_z)) < 1e-14          # last base-lin node, z


def test_linear_equality_and_soft_constraint():
    """towr::LinearEqualityConstraint and towr::SoftConstraint (the two ifopt components no Parameters enum reaches) on
    the device against the oracle."""
    rng = np.random.default_rng(10)
    spec = tb.make_formulation("anymal_trot_block").to_spec(); p = tb.Problem(spec)
    B = 45
    X = synthetic_iterates(p, B)
    bt = p.batch(B)
    (_, c0, nc), = [v for v in p.variable_sets() if v[0] == "ee-motion_1"]
    M = rng.standard_normal((9, nc)); M[rng.random(M.shape) < 0.5] = 0.0
    g = bt.linear_equality(X, "ee-motion_1", M)
    for b in range(B):
        assert np.allclose(g[b], oracle_lib.linear_equality(M, X[b, c0:c0 + nc]), rtol=1e-12, atol=1e-14)
    out = bt.eval_host(X)
    o = oracle_lib.Oracle(spec)
    for name, weights in (("dynamic", None), ("rangeofmotion-2", rng.uniform(0.5, 2.0, 81)), ("terrain-ee-motion_0", None)):
        (_, r0, nr), = [c for c in p.constraint_sets() if c[0] == name]
        cost, grad = bt.soft_constraint(name, weights)
        for b in (0, 7, B - 1):
            r = o.eval(X[b])
            c_ref, g_ref = oracle_lib.soft_constraint(o, r["g"], r["jac"], r0, nr, weights)
            assert np.isclose(cost[b], c_ref, rtol=1e-12), (name, b)
            assert np.allclose(grad[b], g_ref, rtol=1e-11, atol=1e-12 * np.abs(g_ref).max()), (name, b)


def test_footstep_contact_sets_nearest_plane_lookup():
    """fpowr::ExtractFootstepPlan's contact_set (footstep_plan_extractor.h:106-116): the polygon under every foot in contact
    (NearestPlaneLookup, nearest_plane_lookup.h:62-84), -1 in the air — device kernel against the oracle's lookup."""
    from test_oracle import _polys_for_lookup
    polys = _polys_for_lookup()
    p = tb.Problem(tb.make_formulation("anymal_trot_block").to_spec())
    B = 37
    X = synthetic_iterates(p, B)
    plans, sets = p.batch(B).footstep_contact_sets(X, 2.0, polys)
    seen = set()
    for b in range(B):
        assert sets[b].shape == (len(plans[b]), 4)
        for s_i, st in enumerate(plans[b]):
            for e in range(4):
                flag, x, y = st[2 + 4 * e], st[3 + 4 * e], st[4 + 4 * e]
                want = oracle_lib.nearest_plane(polys, x, y) if flag else -1
                assert sets[b][s_i, e] == want, (b, s_i, e)
                seen.add(int(want))
    assert -1 in seen and len(seen) >= 3


def test_postprocessing_matches_golden_fixtures():
    """The batched CUDA post-processing against the committed fixtures (tests/golden/postproc_golden.npz): trajectory,
    initial guesses, footstep plan of one seeded iterate, replicated over a batch with a ragged last tile."""
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "postproc_golden.npz"))
    for name in ("hopper", "hyq_gallop_gap"):
        p = tb.Problem(tb.make_formulation(name).to_spec())
        B = 35
        X = np.tile(z[f"{name}_x"], (B, 1))
        bt = p.batch(B)
        traj = bt.sample_trajectory(X, float(z["dt"])); ig = bt.initial_guesses(X, z["times"]); plans = bt.footstep_plans(X, float(z["time_horizon"]))
        for b in (0, 31, 34):
            for got, ref in ((traj[b], z[f"{name}_trajectory"]), (ig[b], z[f"{name}_initial_guesses"]), (plans[b], z[f"{name}_footstep_plan"])):
                assert got.shape == ref.shape
                scale = np.maximum(1.0, np.abs(ref).max(axis=0, keepdims=True))
                assert np.all(np.abs(got - ref) <= 1e-12 * np.abs(ref) + 1e-13 * scale), (name, b, np.abs(got - ref).max())
