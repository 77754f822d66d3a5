"""CPU tests that pin the oracle itself (the reference holds no golden vectors
for this path — SURVEY.md §8c): finite differences of its own g, closed-form
checks with mpmath, and the committed golden fixtures."""
import os

import numpy as np
import pytest

import towr_b200 as tb
from towr_b200.configs import synthetic_iterates
import oracle_lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _dense(o, vals):
    rp, ci = o.structure()
    J = np.zeros((o.m, o.n))
    J[np.repeat(np.arange(o.m), np.diff(rp)), ci] = vals
    return J


@pytest.mark.parametrize("name,terrain", [("hopper", None), ("hopper", tb.SLOPE), ("biped_walk_stairs", tb.CHIMNEY),
                                          ("hyq_gallop_gap", tb.FLAT)])   # last: PhaseSpline / PhaseDurations derivatives
def test_oracle_jacobian_matches_finite_differences(name, terrain):
    """What IPOPT's derivative_test would do (hopper_example.cc:86), on random iterates."""
    spec = tb.make_formulation(name, terrain=terrain).to_spec()
    o = oracle_lib.Oracle(spec)
    p = tb.Problem(spec)
    x = synthetic_iterates(p, 1, seed=99)[0]
    r = o.eval(x)
    assert r["rc"] == 0
    J = _dense(o, r["jac"])
    h = 1e-6
    cols = np.random.default_rng(5).choice(o.n, 90, replace=False)
    if name == "hyq_gallop_gap":
        cols = np.concatenate([cols[:40], np.arange(o.n - 32, o.n)])   # every ee-schedule variable
    for j in cols:
        xp, xm = x.copy(), x.copy()
        xp[j] += h; xm[j] -= h
        fd = (o.eval(xp)["g"] - o.eval(xm)["g"]) / (2 * h)
        assert np.allclose(J[:, j], fd, rtol=2e-6, atol=2e-6), j


def test_oracle_pattern_is_value_independent():
    """The reference's pattern comes from Eigen algebra at x0; it must not change with x."""
    for name in ("hopper", "anymal_trot_block"):
        spec = tb.make_formulation(name).to_spec()
        o = oracle_lib.Oracle(spec); p = tb.Problem(spec)
        for x in synthetic_iterates(p, 3, seed=5):
            assert o.eval(x)["rc"] == 0
        assert o.eval(np.zeros(o.n))["rc"] == 0


def test_gap_terrain_quirk_reproduced():
    """the force-row derivative w.r.t. the foot position uses the reference's
    element-wise 'derivative of the normalised basis', which is NOT the true chain rule on Gap."""
    spec = tb.make_formulation("hopper", terrain=tb.GAP).to_spec()
    o = oracle_lib.Oracle(spec); p = tb.Problem(spec)
    x = p.GetVariableValues().copy()
    sets = dict((nm, (s, k)) for nm, s, k in p.variable_sets())
    s, k = sets["ee-motion_0"]
    x[s:s + k] += 1.2      # move the feet into the gap parabola x in [1.0, 1.5]
    r = o.eval(x)
    J = _dense(o, r["jac"])
    (_, r0, nr), = [c for c in p.constraint_sets() if c[0].startswith("force")]
    # closed form with mpmath for the first force node's normal-force row
    import mpmath as mp
    mp.mp.dps = 40
    px = mp.mpf(x[s])      # stance position x of phase 0
    w, hh, gs = mp.mpf("0.5"), mp.mpf("1.5"), mp.mpf(1)
    xc = gs + w / 2
    a = 4 * hh / (w * w); b = -8 * hh * xc / (w * w)
    assert 1.0 <= float(px) <= 1.5
    hx, hxx = 2 * a * px + b, 2 * a
    v = [-hx, mp.mpf(0), mp.mpf(1)]
    sn = sum(c * c for c in v); nrm = mp.sqrt(sn)
    dn = [(1 / sn * (nrm * (1 if i == 0 else 0) - v[0] * v[i] / nrm)) * dv for i, dv in enumerate([-hxx, 0, 0])]
    fs, fk = sets["ee-force_0"]
    fvec = [mp.mpf(x[fs + 0]), mp.mpf(x[fs + 2]), mp.mpf(x[fs + 4])]   # node 0: px,vx,py,vy,pz,vz
    expect = sum(fi * di for fi, di in zip(fvec, dn))
    got = J[r0, s]         # d(f.n)/d(foot x)
    assert abs(got - float(expect)) <= 1e-12 * abs(float(expect))


def test_hermite_basis_closed_form():
    """Spline value and node-Jacobian of the oracle against an independent mpmath Hermite evaluation
    (SURVEY App. A.1), through the RoM constraint of the hopper (g = R^T (p_ee - c) with zero angles)."""
    import mpmath as mp
    mp.mp.dps = 40
    spec = tb.make_formulation("hopper").to_spec()
    o = oracle_lib.Oracle(spec); p = tb.Problem(spec)
    x = synthetic_iterates(p, 1, seed=3)[0]
    sets = dict((nm, (s, k)) for nm, s, k in p.variable_sets())
    s_ang, k_ang = sets["base-ang"]
    x[s_ang:s_ang + k_ang] = 0.0                         # R = I exactly
    r = o.eval(x)
    (_, r0, nr), = [c for c in p.constraint_sets() if c[0] == "rangeofmotion-0"]
    # sample k=3 -> t = 0.24 (3 x 0.08): base poly 2 (t_local 0.04), ee-motion: stance phase 0 (0..0.4) => p = node value
    k = 3
    t = 0.0
    for _ in range(k):
        t += 0.08
    s_lin = sets["base-lin"][0]
    def herm(p0, v0, p1, v1, tl, T):
        p0, v0, p1, v1, tl, T = map(mp.mpf, (p0, v0, p1, v1, tl, T))
        C = -(3 * (p0 - p1) + T * (2 * v0 + v1)) / T**2
        D = (2 * (p0 - p1) + T * (v0 + v1)) / T**3
        return p0 + v0 * tl + C * tl**2 + D * tl**3
    tl = t - 0.1 - 0.1
    for d in range(3):
        c = herm(x[s_lin + 12 + d], x[s_lin + 15 + d], x[s_lin + 18 + d], x[s_lin + 21 + d], tl, 0.1)
        pe = mp.mpf(x[sets["ee-motion_0"][0] + d])
        assert abs(r["g"][r0 + 3 * k + d] - float(pe - c)) < 1e-14


def test_golden_fixtures():
    """Outputs of the oracle committed under tests/golden (made by scripts/make_golden.py)."""
    path = os.path.join(ROOT, "tests", "golden", "oracle_golden.npz")
    z = np.load(path)
    for name in ("hopper", "anymal_trot_block", "biped_walk_stairs"):
        spec = tb.make_formulation(name).to_spec()
        o = oracle_lib.Oracle(spec); p = tb.Problem(spec)
        X = synthetic_iterates(p, 2, seed=int(z["seed"]))
        assert np.array_equal(X, z[f"{name}_x"])
        rp, ci = o.structure()
        assert np.array_equal(rp, z[f"{name}_row_ptr"]) and np.array_equal(ci, z[f"{name}_col_idx"])
        for b in range(2):
            r = o.eval(X[b])
            assert np.allclose(r["g"], z[f"{name}_g"][b], rtol=1e-13, atol=1e-13)
            assert np.allclose(r["jac"], z[f"{name}_jac"][b], rtol=1e-13, atol=1e-13)


def _csv_grid_reference(grid, x, y):
    """Independent restatement of towr::HeightMapFromCSV (height_map_from_csv.h:29-111) in plain Python."""
    res, eps = 0.17, 0.17 / 50
    rows, cols = grid.shape
    if x / res <= -1 or y / res <= -1:          # static_cast<size_t> truncates toward zero: (-1, 0) is cell 0, :31-32
        return 0.0, 0.0, 0.0
    xc, yc = int(x / res), int(y / res)
    if xc >= cols or yc >= rows:
        return 0.0, 0.0, 0.0
    def slope(c, n, at, coord):
        if c + 1 < n:
            d = at(c + 1) - at(c); end = (c + 1) * res
            if d > 0 and end - eps <= coord <= end:
                return d / eps
        if c - 1 >= 0:
            d = at(c) - at(c - 1); start = c * res
            if d < 0 and start <= coord <= start + eps:
                return d / eps
        return 0.0
    return (grid[yc, xc], slope(xc, cols, lambda k: grid[yc, k], x), slope(yc, rows, lambda k: grid[k, xc], y))


def test_csv_grid_terrain_of_the_oracle():
    rng = np.random.default_rng(4)
    grid = rng.integers(0, 4, (9, 14)) * 0.05
    oracle_lib.set_grid(grid)
    res, eps = 0.17, 0.17 / 50
    pts = list(rng.uniform(-0.3, 2.6, (300, 2)))
    for k in range(1, 12):                       # points inside the eps bands at the cell edges, both directions
        pts += [(k * res - 0.4 * eps, 0.5), (k * res + 0.4 * eps, 0.5), (0.6, (k % 8 + 1) * res - 0.3 * eps), (0.6, (k % 8 + 1) * res + 0.3 * eps)]
    pts += [(-0.05, 0.4), (0.4, -0.05), (-0.05, -0.05), (-0.169, 0.2), (-0.17, 0.2), (-0.171, 0.2), (0.2, -0.1701)]   # the (-res, 0) band is cell 0
    assert grid[:, 0].any() and grid[0, :].any()
    assert oracle_lib.terrain_point(tb.GRID_CSV, -0.05, 0.4)[0] == grid[2, 0]
    hit = 0
    for x, y in pts:
        got = oracle_lib.terrain_point(tb.GRID_CSV, x, y)
        want = _csv_grid_reference(grid, x, y)
        assert tuple(got) == tuple(float(v) for v in want), (x, y, got, want)
        hit += (want[1] != 0.0) or (want[2] != 0.0)
    assert hit > 5                               # the edge-slope branches were exercised


def test_grid_map_terrain_of_the_oracle():
    """towr `Grid` (grid_height_map.h) over grid_map's bilinear atPosition: the oracle against an independent numpy
    statement (continuous cell coordinate, floor, four-cell blend) inside the map, FLT_MAX outside, eps = res / 6 slopes."""
    rng = np.random.default_rng(11)
    sx, sy, res, pos = 40, 30, 0.1, (1.0, 0.0)
    H = (0.2 * rng.standard_normal((sx, sy))).astype(np.float32)
    oracle_lib.set_grid_map(H, res, pos)
    Lx, Ly = sx * res, sy * res

    def bilinear(x, y):
        u = (pos[0] + Lx / 2 - res / 2 - x) / res          # continuous index along x (index 0 = largest coordinate)
        v = (pos[1] + Ly / 2 - res / 2 - y) / res
        i0, j0 = int(np.floor(u)), int(np.floor(v))
        a, b = u - i0, v - j0
        return ((1 - a) * (1 - b) * H[i0, j0] + a * (1 - b) * H[i0 + 1, j0] + (1 - a) * b * H[i0, j0 + 1] + a * b * H[i0 + 1, j0 + 1])

    for x, y in zip(rng.uniform(-0.8, 2.8, 400), rng.uniform(-1.3, 1.3, 400)):
        h, hx, hy = oracle_lib.terrain_point(tb.GRID_MAP, x, y)
        assert abs(h - bilinear(x, y)) <= 3e-7 * max(1.0, abs(h)), (x, y, h, bilinear(x, y))
        eps = res / 6
        assert abs(hx - (bilinear(x + eps, y) - bilinear(x - eps, y)) / (2 * eps)) <= 2e-5
        assert abs(hy - (bilinear(x, y + eps) - bilinear(x, y - eps)) / (2 * eps)) <= 2e-5
        assert h == float(np.float32(h))                     # heights are floats (grid_height_map.h:38-45)
    fmax = float(np.finfo(np.float32).max)
    for x, y in ((1.0, 1.7), (1.0, -1.6), (-1.3, -1.6), (3.3, 1.7)):
        assert oracle_lib.terrain_point(tb.GRID_MAP, x, y)[0] == fmax      # std::out_of_range -> numeric_limits<float>::max()
    # grid_map checks the four cells by their LINEAR index only: an x index beyond the map wraps into the neighbouring
    # column, so a position outside along x (with y inside) still "interpolates" — restated as is
    assert oracle_lib.terrain_point(tb.GRID_MAP, -1.2, 0.0)[0] != fmax
    # inside the map but beyond the outermost cell centres at a corner: the linear index of a neighbour is out of range ->
    # nearest cell (INTER_NEAREST)
    assert oracle_lib.terrain_point(tb.GRID_MAP, 2.98, 1.48)[0] == float(H[0, 0])
    assert oracle_lib.terrain_point(tb.GRID_MAP, -0.98, -1.48)[0] == float(H[39, 29])


def _polys_for_lookup():
    sq = lambda x0, y0, w, h: np.array([[x0, y0], [x0, y0 + h], [x0 + w, y0 + h], [x0 + w, y0], [x0, y0]])   # closed, clockwise
    return [sq(0.0, -0.5, 1.0, 1.0), sq(1.2, -0.5, 0.8, 1.0), np.array([[2.5, 0.0], [3.0, 0.8], [3.5, 0.0], [2.5, 0.0]]), sq(1.2, -0.5, 0.8, 1.0)]


def _nearest_plane_reference(polys, x, y):
    """Independent statement for CLOSED simple polygons: 0 inside (even-odd ray test) or on the boundary, else the
    smallest vertex / edge distance; first smallest wins."""
    best, idx = np.inf, -1
    for k, q in enumerate(polys):
        inside = False; d = np.inf
        for (ax, ay), (bx, by) in zip(q[:-1], q[1:]):
            if (ay > y) != (by > y) and x < ax + (y - ay) * (bx - ax) / (by - ay):
                inside = not inside
            t = np.clip(((x - ax) * (bx - ax) + (y - ay) * (by - ay)) / ((bx - ax) ** 2 + (by - ay) ** 2), 0.0, 1.0)
            d = min(d, np.hypot(x - (ax + t * (bx - ax)), y - (ay + t * (by - ay))))
        d = 0.0 if inside else d
        if d < best:
            best, idx = d, k
    return idx


def test_nearest_plane_lookup_of_the_oracle():
    polys = _polys_for_lookup()
    rng = np.random.default_rng(5)
    pts = list(rng.uniform((-0.5, -1.0), (4.0, 1.2), (500, 2))) + [(0.5, 0.0), (1.1, 0.0), (1.6, 0.2), (3.0, 0.3), (1.1, 0.6), (1.0, 0.0), (5.0, 5.0)]
    hits = set()
    for x, y in pts:
        got = oracle_lib.nearest_plane(polys, x, y)
        assert got == _nearest_plane_reference(polys, x, y), (x, y)
        hits.add(got)
    assert hits == {0, 1, 2}                      # polygon 3 duplicates polygon 1: the first of two equal distances wins (strict <)
    assert oracle_lib.nearest_plane(polys, 1.0, 0.0) == 0      # on the boundary of polygon 0: distance 0


def test_linear_equality_and_soft_constraint_of_the_oracle():
    """towr::LinearEqualityConstraint (linear_constraint.cc) and towr::SoftConstraint (soft_constraint.cc): the oracle's
    restatements against numpy (M x; 0.5 (g-b)^T W (g-b) and J^T W (g-b) with the dense Jacobian of the set)."""
    rng = np.random.default_rng(9)
    spec = tb.make_formulation("hopper").to_spec()
    o = oracle_lib.Oracle(spec); p = tb.Problem(spec)
    x = synthetic_iterates(p, 1)[0]
    (_, c0, nc), = [v for v in p.variable_sets() if v[0] == "base-lin"]
    M = rng.standard_normal((7, nc)); M[rng.random(M.shape) < 0.6] = 0.0
    assert np.allclose(oracle_lib.linear_equality(M, x[c0:c0 + nc]), M @ x[c0:c0 + nc], rtol=1e-13, atol=1e-13)
    r = o.eval(x)
    rp, ci = o.structure()
    _, _, gl, gu = p.bounds()
    for name in ("dynamic", "rangeofmotion-0", "force-ee-force_0"):
        (_, r0, nr), = [c for c in p.constraint_sets() if c[0] == name]
        w = rng.uniform(0.5, 2.0, nr)
        cost, grad = oracle_lib.soft_constraint(o, r["g"], r["jac"], r0, nr, w)
        J = np.zeros((nr, p.n))
        for i in range(nr):
            J[i, ci[rp[r0 + i]:rp[r0 + i + 1]]] = r["jac"][rp[r0 + i]:rp[r0 + i + 1]]
        d = r["g"][r0:r0 + nr] - 0.5 * (np.clip(gu[r0:r0 + nr], -1e20, 1e20) + np.clip(gl[r0:r0 + nr], -1e20, 1e20))
        assert np.isclose(cost, 0.5 * d @ (w * d), rtol=1e-12)
        assert np.allclose(grad, J.T @ (w * d), rtol=1e-11, atol=1e-9 * np.abs(grad).max())


def _extended_precision_oracle():
    """oracle/_build/libtowr_oracle_ld.so: the oracle source compiled with every double as x87 long double (64-bit mantissa)."""
    import ctypes as C
    oracle_lib.build()
    lib = C.CDLL(os.path.join(os.path.dirname(oracle_lib.OUT), "libtowr_oracle_ld.so"))
    lib.oracle_create.restype = C.c_void_p; lib.oracle_create.argtypes = [C.c_void_p]
    lib.oracle_destroy.argtypes = [C.c_void_p]
    lib.oracle_eval.restype = C.c_int
    return lib


@pytest.mark.parametrize("name", ["hopper", "anymal_trot_block", "biped_walk_stairs", "hyq_gallop_gap"])
def test_jacobian_pinned_to_1e10_by_extended_precision_differentiation(name):
    """The analytic Jacobian of the (double) oracle against a numerical derivative of g that is accurate to ~1e-12:
    Richardson-extrapolated central differences of the SAME source compiled in 80-bit extended precision, for EVERY
    variable — node values of all sets and, for the duration-optimised config, every ee-schedule column.  A double-precision
    finite difference (the derivative_test of hopper_example.cc:86) resolves ~1e-6; this pins each entry to 1e-10 of its row."""
    import ctypes as C
    assert np.finfo(np.longdouble).nmant >= 63, "needs x87 extended precision"
    spec = tb.make_formulation(name).to_spec()
    o = oracle_lib.Oracle(spec); p = tb.Problem(spec)
    x = synthetic_iterates(p, 1, seed=4321)[0]
    r = o.eval(x)
    assert r["rc"] == 0
    rp, ci = o.structure()
    J = np.zeros((p.m, p.n))
    for i in range(p.m):
        J[i, ci[rp[i]:rp[i + 1]]] = r["jac"][rp[i]:rp[i + 1]]
    structural = np.zeros((p.m, p.n), bool)
    for i in range(p.m):
        structural[i, ci[rp[i]:rp[i + 1]]] = True
    lib = _extended_precision_oracle()
    h_ld = C.c_void_p(lib.oracle_create(C.byref(spec)))
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)

    def g_of(xq):
        g = np.empty(p.m, np.longdouble)
        lib.oracle_eval(h_ld, ptr(np.ascontiguousarray(xq, np.longdouble)), ptr(g), None, None, None)
        return g

    # The one place where the reference's Jacobian is NOT the derivative of its g: ForceConstraint
    # differentiates the normalised terrain basis element-wise (height_map.cc:62-91 as used by force_constraint.cc:131-171),
    # which is only exact where the terrain has no curvature.  On Gap (d2h/dx2 != 0) the force rows' entries w.r.t. the
    # foot position follow the reference, not the finite difference: excluded from the pin (and counted).
    quirk = np.zeros((p.m, p.n), bool)
    if spec.terrain == tb.GAP:
        for cn, r0, nr in p.constraint_sets():
            if cn.startswith("force-"):
                (_, c0, nc), = [v for v in p.variable_sets() if v[0] == "ee-motion_" + cn[-1]]
                quirk[r0:r0 + nr, c0:c0 + nc] = True
    xq = x.astype(np.longdouble)
    row_scale = np.maximum(1.0, np.abs(J).max(axis=1))
    worst, skipped, checked = 0.0, 0, 0
    for j in range(p.n):
        h = np.longdouble(1e-3) * max(1.0, abs(x[j]))
        def central(step):
            a, b = xq.copy(), xq.copy(); a[j] += step; b[j] -= step
            return (g_of(a) - g_of(b)) / (2 * step)
        d1, d2, d3 = central(h), central(h / 2), central(h / 4)
        rich = (64 * d3 - 20 * d2 + d1) / 45                       # two Richardson steps: error O(h^6)
        rough = np.abs(np.asarray((4 * d2 - d1) / 3 - rich, np.float64))
        col = np.asarray(rich, np.float64)
        # a kink of a piecewise terrain (Block / Stairs edges) inside the stencil shows as disagreeing extrapolations: skip
        ok = (rough <= 1e-9 * row_scale) & ~quirk[:, j]
        skipped += int((~ok).sum()); checked += int(ok.sum())
        err = np.abs(col - J[:, j]) / row_scale
        worst = max(worst, float(err[ok].max()))
        assert np.all(err[ok] <= 1e-10), (name, j, float(err[ok].max()))
        assert np.all(np.abs(col[ok & ~structural[:, j]]) <= 1e-10 * row_scale[ok & ~structural[:, j]]), (name, j)   # nothing outside the pattern
    lib.oracle_destroy(h_ld)
    assert skipped - int(quirk.sum()) <= 0.002 * (checked + skipped), (skipped, checked)
    print(f"{name}: {checked} derivative entries pinned, worst |dJ| / row scale {worst:.2e}, {skipped} skipped at terrain kinks")


def _quat_from_euler_zyx(roll, pitch, yaw):
    """Independent check: quaternion (w, x, y, z) of R = Rz(yaw) Ry(pitch) Rx(roll)."""
    cr, sr, cp, sp, cy, sy = np.cos(roll / 2), np.sin(roll / 2), np.cos(pitch / 2), np.sin(pitch / 2), np.cos(yaw / 2), np.sin(yaw / 2)
    return np.array([cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy])


@pytest.mark.parametrize("name", ["hopper", "anymal_trot_block", "hyq_gallop_gap"])
def test_oracle_initial_guess_is_consistent_with_trajectory(name):
    """fpowr::ExtractInitialGuess (initial_guess_extractor.h:17-34) restated by the oracle: the state / controls entries
    are the same spline points fpowr::GetTrajectory samples; the Euler angles reproduce the trajectory's quaternion."""
    spec = tb.make_formulation(name).to_spec(); p = tb.Problem(spec); o = oracle_lib.Oracle(spec)
    x = synthetic_iterates(p, 1, seed=3)[0]
    dt = 0.25
    tr = o.trajectory(x, dt)
    times = np.arange(tr.shape[0]) * dt
    ig = o.initial_guesses(x, times)
    n_ee = (tr.shape[1] - 19) // 13
    assert np.array_equal(ig[:, 0], times)
    assert np.allclose(ig[:, 1:4], tr[:, 0:3], rtol=0, atol=1e-13) and np.allclose(ig[:, 7:10], tr[:, 3:6], rtol=0, atol=1e-13)
    for k in range(len(times)):
        q = _quat_from_euler_zyx(*ig[k, 4:7])
        assert min(np.abs(q - tr[k, 9:13]).max(), np.abs(q + tr[k, 9:13]).max()) < 1e-12
    for e in range(n_ee):
        assert np.allclose(ig[:, 13 + 3 * e:16 + 3 * e], tr[:, 19 + 13 * e + 7:19 + 13 * e + 10], rtol=0, atol=1e-12)   # ee acceleration
        assert np.allclose(ig[:, 37 + 3 * e:40 + 3 * e], tr[:, 19 + 13 * e + 10:19 + 13 * e + 13], rtol=0, atol=1e-12)  # ee force
    assert np.all(ig[:, 25:37] == 0.0) and np.all(ig[:, 13 + 3 * n_ee:25] == 0.0)


def test_oracle_footstep_plan_of_the_hopper():
    """fpowr::ExtractFootstepPlan (footstep_plan_extractor.h:68-133) on hopper_example's gait (phases 0.4 0.2 0.4 0.2 0.4
    0.2 0.2, in contact at the start): a phase boundary belongs to the ending phase (Spline::GetSegmentID), so every
    contact change shows up at the first 0.01 s sample after it; durations tile the horizon."""
    spec = tb.make_formulation("hopper").to_spec(); p = tb.Problem(spec); o = oracle_lib.Oracle(spec)
    x0 = p.GetVariableValues()
    plan = o.footstep_plan(x0, 2.0)
    assert plan.shape == (7, 6)
    assert np.allclose(plan[:, 0], [0.0, 0.41, 0.61, 1.01, 1.21, 1.61, 1.81], rtol=0, atol=1e-9)
    assert list(plan[:, 2]) == [1.0, 0.0, 1.0, 0.0, 1.0, 0.0, 1.0]
    assert abs(plan[:, 1].sum() - 2.0) < 1e-12 and np.allclose(plan[:-1, 1], np.diff(plan[:, 0]), rtol=0, atol=1e-15)
    tr = o.trajectory(x0, 0.01)
    for row in plan:                                   # positions are those of the trajectory sample at t_global
        k = int(round(row[0] / 0.01))
        assert np.array_equal(row[3:6], tr[k, 20:23])


def test_postprocessing_golden_fixtures():
    """fpowr post-processing outputs of the oracle committed under tests/golden/postproc_golden.npz (scripts/make_golden.py)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "postproc_golden.npz"))
    for name in ("hopper", "hyq_gallop_gap"):
        spec = tb.make_formulation(name).to_spec(); o = oracle_lib.Oracle(spec); p = tb.Problem(spec)
        x = synthetic_iterates(p, 1, seed=int(z["seed"]))[0]
        assert np.array_equal(x, z[f"{name}_x"])
        assert np.allclose(o.trajectory(x, float(z["dt"])), z[f"{name}_trajectory"], rtol=1e-13, atol=1e-13)
        assert np.allclose(o.initial_guesses(x, z["times"]), z[f"{name}_initial_guesses"], rtol=1e-13, atol=1e-13)
        plan = o.footstep_plan(x, float(z["time_horizon"]))
        assert plan.shape == z[f"{name}_footstep_plan"].shape and np.allclose(plan, z[f"{name}_footstep_plan"], rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("name,robot", [("anymal_trot_block", tb.ANYMAL), ("hyq_gallop_gap", tb.HYQ)])
def test_dynamic_and_range_of_motion_values_against_scipy(name, robot):
    """The oracle's constraint VALUES of the two big sets, recomputed with independent library code: scipy's rotations for the
    Euler convention (intrinsic Z-Y'-X'', euler_converter.cc:207-221) and numerical differentiation of those rotations for the
    angular velocity / acceleration in the world frame (euler_converter.cc:58-83), numpy for the single-rigid-body equations
    (single_rigid_body_dynamics.cc:76-101) and the range-of-motion vector (range_of_motion_constraint.cc:58-66).  Pins the row
    order (angular rows first), every sign, the inertia convention (products of inertia enter negated) and gravity."""
    from scipy.spatial.transform import Rotation
    spec = tb.make_formulation(name).to_spec()
    o = oracle_lib.Oracle(spec)
    p = tb.Problem(spec)
    x = synthetic_iterates(p, 1, seed=11)[0]
    g = o.eval(x)["g"]
    info = tb.robot_info(robot)
    n_ee, mass = info["n_ee"], info["mass"]
    ixx, iyy, izz, ixy, ixz, iyz = info["inertia"]
    I_b = np.array([[ixx, -ixy, -ixz], [-ixy, iyy, -iyz], [-ixz, -iyz, izz]])
    sets = {nm: (r0, nr) for nm, r0, nr in o.constraint_sets()}

    def rot(euler):                       # towr stores roll, pitch, yaw; applied Z (yaw), Y' (pitch), X'' (roll)
        return Rotation.from_euler("ZYX", [euler[2], euler[1], euler[0]]).as_matrix()

    def state(t):                         # base position, Euler angles; feet positions, forces; base acceleration
        ig = o.initial_guesses(x, np.array([t]))[0]
        return ig[1:4], ig[4:7], ig

    def vee(S):
        return np.array([S[2, 1] - S[1, 2], S[0, 2] - S[2, 0], S[1, 0] - S[0, 1]]) / 2.0

    # The dynamic samples sit exactly on the junctions of the base polynomials (both 0.1 s), where the acceleration jumps and
    # Spline::GetSegmentID selects the polynomial that ENDS there: derivatives from the left, 4th-order backward stencil.
    def dleft(f, t, h):
        return (25 * f(t) - 48 * f(t - h) + 36 * f(t - 2 * h) - 16 * f(t - 3 * h) + 3 * f(t - 4 * h)) / (12 * h)

    def omega(t, h=1e-4):                 # vee(dR/dt R^T) of scipy rotations (truncation 6e-12 at this step: the error falls as h^4)
        return vee(dleft(lambda u: rot(state(u)[1]), t, h) @ rot(state(t)[1]).T)

    dt = 0.1
    traj = o.trajectory(x, dt)            # samples at 0, 0.1, ...: base lin p v a | quaternion | omega | omega_dot | feet: contact, p, v, a, f
    r0, nr = sets["dynamic"]
    T = 2.0
    worst = 0.0
    for k in range(1, nr // 6 - 1):       # interior samples (the differences need t +- h inside the horizon)
        t = k * dt
        s = traj[k]
        c, cdd = s[0:3], s[6:9]
        R = rot(state(t)[1])
        w = omega(t)
        wd = dleft(omega, t, 5e-4)
        I_w = R @ I_b @ R.T
        fsum, tau = np.zeros(3), np.zeros(3)
        for e in range(n_ee):
            foot = s[19 + 13 * e: 19 + 13 * (e + 1)]
            pe, f = foot[1:4], foot[10:13]
            fsum += f
            tau += np.cross(pe - c, f)
        ang = I_w @ wd + np.cross(w, I_w @ w) - tau
        lin = mass * cdd - fsum - np.array([0.0, 0.0, -mass * 9.80665])
        ref = np.concatenate([ang, lin])
        got = g[r0 + 6 * k: r0 + 6 * k + 6]
        scale = max(1.0, np.abs(ref).max())
        worst = max(worst, np.abs(got - ref).max() / scale)
        # the oracle's own world-frame angular velocity / acceleration (fpowr::GetTrajectory) against the differentiated rotations
        assert np.allclose(s[13:16], w, atol=1e-9, rtol=1e-9) and np.allclose(s[16:19], wd, atol=1e-6, rtol=1e-6)
    assert worst < 1e-6, worst            # limited by the second numerical derivative, not by the formulas
    # range of motion: g_e = R^T (p_e - c) at t = k * 0.08 (+ the horizon end)
    dtr = 0.08
    trr = o.trajectory(x, dtr)
    for e in range(n_ee):
        r0, nr = sets[f"rangeofmotion-{e}"]
        for k in range(nr // 3 - 1):      # (the last sample sits at T, off the 0.08 grid)
            s = trr[k]
            R = rot(state(k * dtr)[1])
            pe = s[19 + 13 * e + 1: 19 + 13 * e + 4]
            ref = R.T @ (pe - s[0:3])
            assert np.allclose(g[r0 + 3 * k: r0 + 3 * k + 3], ref, atol=1e-11, rtol=1e-11), (e, k)
