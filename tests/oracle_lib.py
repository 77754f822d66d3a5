"""Test-side loader of the CPU oracle (oracle/towr_oracle.cc).  Test
infrastructure only — the product package never imports this."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "oracle", "towr_oracle.cc")
OUT = os.path.join(ROOT, "oracle", "_build", "libtowr_oracle.so")


def build(force=False):
    if force or not os.path.exists(OUT) or os.path.getmtime(OUT) < os.path.getmtime(SRC):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    return OUT


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_create.restype = C.c_void_p
        _lib.oracle_create.argtypes = [C.c_void_p]
        _lib.oracle_destroy.argtypes = [C.c_void_p]
        for fn in ("oracle_dims", "oracle_structure", "oracle_bounds", "oracle_x0", "oracle_set_terrain", "oracle_set_grid", "oracle_set_grid_map", "oracle_linear_equality", "oracle_soft_constraint",
                   "oracle_terrain_point"):
            getattr(_lib, fn).restype = None
        _lib.oracle_eval.restype = C.c_int
        _lib.oracle_batch_eval.restype = C.c_int
        _lib.oracle_terrain_height.restype = C.c_double
        _lib.oracle_terrain_height.argtypes = [C.c_int, C.c_double, C.c_double]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """One problem instance evaluated by the CPU restatement."""

    def __init__(self, spec):
        self.spec = spec
        self._h = C.c_void_p(lib().oracle_create(C.byref(spec)))
        n, m, nnz = C.c_int(), C.c_int(), C.c_int()
        lib().oracle_dims(self._h, C.byref(n), C.byref(m), C.byref(nnz))
        self.n, self.m, self.nnz = n.value, m.value, nnz.value

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_destroy(self._h)
            self._h = None

    def structure(self):
        rp = np.empty(self.m + 1, np.int32)
        ci = np.empty(self.nnz, np.int32)
        lib().oracle_structure(self._h, _p(rp), _p(ci))
        return rp, ci

    def bounds(self):
        xl, xu, gl, gu = np.empty(self.n), np.empty(self.n), np.empty(self.m), np.empty(self.m)
        lib().oracle_bounds(self._h, _p(xl), _p(xu), _p(gl), _p(gu))
        return xl, xu, gl, gu

    def x0(self):
        x = np.empty(self.n)
        lib().oracle_x0(self._h, _p(x))
        return x

    def _sets(self, count_fn, get_fn):
        out, buf, start = [], C.create_string_buffer(64), 0
        get_fn.restype = C.c_int
        for i in range(count_fn(self._h)):
            rows = get_fn(self._h, i, buf, 64)
            out.append((buf.value.decode(), start, rows))
            start += rows
        return out

    def variable_sets(self):
        return self._sets(lib().oracle_num_varsets, lib().oracle_varset)

    def constraint_sets(self):
        return self._sets(lib().oracle_num_csets, lib().oracle_cset)

    def set_terrain(self, terrain):
        lib().oracle_set_terrain(self._h, int(terrain))

    def trajectory(self, x, dt):
        """fpowr::GetTrajectory: (n_samples, 19 + 13 n_ee) samples of the solution x every dt."""
        ns, nv = C.c_int(), C.c_int()
        lib().oracle_trajectory_dims(self._h, C.c_double(dt), C.byref(ns), C.byref(nv))
        out = np.empty((ns.value, nv.value))
        x = np.ascontiguousarray(x, np.float64)
        lib().oracle_trajectory(self._h, _p(x), C.c_double(dt), _p(out))
        return out

    def initial_guesses(self, x, times):
        """fpowr::ExtractInitialGuesses: (n_times, 49) = time | state[12] | controls[36]."""
        x = np.ascontiguousarray(x, np.float64); times = np.ascontiguousarray(times, np.float64)
        out = np.empty((times.size, 49))
        rc = lib().oracle_initial_guess(self._h, _p(x), _p(times), C.c_int(times.size), _p(out))
        assert rc == 0
        return out

    def footstep_plan(self, x, time_horizon, max_states=64):
        """fpowr::ExtractFootstepPlan without the plane lookup: (n_states, 2 + 4 n_ee)."""
        x = np.ascontiguousarray(x, np.float64)
        ns, nv = C.c_int(), C.c_int()
        lib().oracle_trajectory_dims(self._h, C.c_double(0.01), C.byref(ns), C.byref(nv))
        n_ee = (nv.value - 19) // 13
        out = np.zeros((max_states, 2 + 4 * n_ee)); count = C.c_int()
        lib().oracle_footstep_plan(self._h, _p(x), C.c_double(time_horizon), C.c_int(max_states), C.byref(count), _p(out))
        assert count.value <= max_states
        return out[:count.value]

    def eval(self, x, want_cost=False):
        x = np.ascontiguousarray(x, np.float64)
        g, vals = np.empty(self.m), np.empty(self.nnz)
        cost = np.zeros(1)
        grad = np.zeros(self.n)
        rc = lib().oracle_eval(self._h, _p(x), _p(g), _p(vals), _p(cost) if want_cost else None,
                               _p(grad) if want_cost else None)
        return dict(rc=rc, g=g, jac=vals, cost=float(cost[0]), grad=grad)


def batch_eval(spec, X, terrain_ids=None, want_cost=False, threads=0, want_jac=True):
    X = np.ascontiguousarray(X, np.float64)
    B, n = X.shape
    o = Oracle(spec)
    g, vals = np.empty((B, o.m)), (np.empty((B, o.nnz)) if want_jac else None)
    cost = np.zeros(B) if want_cost else None
    grad = np.zeros((B, n)) if want_cost else None
    t = None if terrain_ids is None else np.ascontiguousarray(terrain_ids, np.int32)
    rc = lib().oracle_batch_eval(C.byref(spec), B, _p(t), _p(X), _p(g), _p(vals), _p(cost), _p(grad), int(threads))
    return dict(rc=rc, g=g, jac=vals, cost=cost, grad=grad)


def set_grid(heights):
    """Height grid of the GRID_CSV terrain (process-global in the oracle)."""
    h = np.ascontiguousarray(heights, np.float64)
    lib().oracle_set_grid(_p(h), C.c_int(h.shape[0]), C.c_int(h.shape[1]))


def set_grid_map(heights, resolution, position=(0.0, 0.0)):
    """Elevation layer of the GRID_MAP terrain (process-global in the oracle): heights[ix, iy] float32."""
    h = np.ascontiguousarray(heights, np.float32)
    lib().oracle_set_grid_map(_p(h), C.c_int(h.shape[0]), C.c_int(h.shape[1]), C.c_double(resolution), C.c_double(position[0]), C.c_double(position[1]))


def nearest_plane(polygons, x, y):
    """fpowr::NearestPlaneLookup::GetNearestPlaneIndex for one point; polygons: list of (k, 2) arrays."""
    offs = np.zeros(len(polygons) + 1, np.int32)
    offs[1:] = np.cumsum([len(q) for q in polygons])
    verts = np.ascontiguousarray(np.concatenate([np.asarray(q, np.float64).reshape(-1, 2) for q in polygons]))
    lib().oracle_nearest_plane.restype = C.c_int
    return lib().oracle_nearest_plane(_p(offs), C.c_int(len(polygons)), _p(verts), C.c_double(x), C.c_double(y))


def linear_equality(M, x_set):
    M = np.ascontiguousarray(M, np.float64); x_set = np.ascontiguousarray(x_set, np.float64)
    g = np.empty(M.shape[0])
    lib().oracle_linear_equality(_p(M), C.c_int(M.shape[0]), C.c_int(M.shape[1]), _p(x_set), _p(g))
    return g


def soft_constraint(oracle, g, vals, row0, n_rows, weights=None):
    """towr::SoftConstraint of rows row0.. of an evaluated Oracle instance: (cost, grad[n])."""
    cost = C.c_double(); grad = np.empty(oracle.n)
    w = None if weights is None else np.ascontiguousarray(weights, np.float64)
    lib().oracle_soft_constraint(oracle._h, _p(np.ascontiguousarray(g)), _p(np.ascontiguousarray(vals)), C.c_int(row0), C.c_int(n_rows), _p(w),
                                 C.byref(cost), _p(grad))
    return cost.value, grad


def terrain_point(terrain, x, y):
    out = np.empty(3)
    lib().oracle_terrain_point(C.c_int(int(terrain)), C.c_double(x), C.c_double(y), _p(out))
    return out


def max_threads():
    return lib().oracle_max_threads()
