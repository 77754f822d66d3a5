"""ctypes binding of include/towr_b200.h (the C ABI of libtowr_b200.so).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C
towr_b200/csrc``.  There is no Python or CPU fallback for the evaluation path:
if the shared library is missing, importing this module raises.
"""
import ctypes as C
import os

MAX_EE, MAX_PHASES, MAX_CONSTRAINTS, MAX_COSTS = 4, 32, 16, 8

OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_NO_DEVICE = 0, 1, 2, 3, 4
EVAL_G, EVAL_JAC, EVAL_COST, EVAL_ALL = 1, 2, 4, 7

# towr::RobotModel::Robot (robot_model.h:70-75)
MONOPED, BIPED, HYQ, ANYMAL, GO1 = range(5)
# towr::HeightMap::TerrainID (height_map.h:79-86)
FLAT, BLOCK, STAIRS, GAP, SLOPE, CHIMNEY, CHIMNEY_LR = range(7)
GRID_CSV = 7   # towr::HeightMapFromCSV; grid data per batch (Batch.set_grid_terrain)
GRID_MAP = 8   # towr Grid (grid_height_map.h): grid_map elevation layer per batch (Batch.set_grid_map)
# towr::Parameters::ConstraintName (parameters.h:139-147)
C_DYNAMIC, C_EE_ROM, C_TOTAL_TIME, C_TERRAIN, C_FORCE, C_SWING, C_BASE_ROM, C_BASE_ACC = range(8)
# towr::Parameters::CostName
COST_FORCES, COST_EE_MOTION = range(2)


class Spec(C.Structure):
    """struct twb_spec"""
    _fields_ = [
        ("robot", C.c_int), ("terrain", C.c_int), ("n_ee", C.c_int),
        ("n_phases", C.c_int * MAX_EE),
        ("phase_durations", (C.c_double * MAX_PHASES) * MAX_EE),
        ("in_contact_at_start", C.c_int * MAX_EE),
        ("initial_base_lin_pos", C.c_double * 3), ("initial_base_lin_vel", C.c_double * 3),
        ("initial_base_ang_pos", C.c_double * 3), ("initial_base_ang_vel", C.c_double * 3),
        ("final_base_lin_pos", C.c_double * 3), ("final_base_lin_vel", C.c_double * 3),
        ("final_base_ang_pos", C.c_double * 3), ("final_base_ang_vel", C.c_double * 3),
        ("initial_ee_W", (C.c_double * 3) * MAX_EE),
        ("duration_base_polynomial", C.c_double),
        ("force_polynomials_per_stance_phase", C.c_int),
        ("ee_polynomials_per_swing_phase", C.c_int),
        ("force_limit_in_normal_direction", C.c_double),
        ("dt_constraint_range_of_motion", C.c_double),
        ("dt_constraint_dynamic", C.c_double),
        ("dt_constraint_base_motion", C.c_double),
        ("bound_phase_duration_min", C.c_double), ("bound_phase_duration_max", C.c_double),
        ("n_constraints", C.c_int), ("constraints", C.c_int * MAX_CONSTRAINTS),
        ("n_costs", C.c_int), ("cost_ids", C.c_int * MAX_COSTS), ("cost_weights", C.c_double * MAX_COSTS),
        ("bounds_final_lin_pos", C.c_int * 3), ("bounds_final_lin_vel", C.c_int * 3),
        ("bounds_final_ang_pos", C.c_int * 3), ("bounds_final_ang_vel", C.c_int * 3),
    ]


# TWB_LIB: developer knob for tuning sweeps (scripts/build_variants.sh) — another build of the same library
LIB_PATH = os.environ.get("TWB_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libtowr_b200.so")


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(towr_b200 has no fallback path)")
    lib = C.CDLL(LIB_PATH)
    P, D, I = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)
    sig = {
        "twb_spec_default": (C.c_int, [C.POINTER(Spec), C.c_int]),
        "twb_spec_optimize_phase_durations": (C.c_int, [C.POINTER(Spec)]),
        "twb_spec_set_gait": (C.c_int, [C.POINTER(Spec), C.c_int, C.c_int, C.c_double]),
        "twb_robot_info": (C.c_int, [C.c_int, I, D, D, D, D]),
        "twb_terrain_height": (C.c_double, [C.c_int, C.c_double, C.c_double]),
        "twb_problem_create": (C.c_int, [C.POINTER(Spec), C.POINTER(P)]),
        "twb_problem_destroy": (None, [P]),
        "twb_problem_dims": (C.c_int, [P, I, I, I]),
        "twb_problem_structure": (C.c_int, [P, I, I]),
        "twb_problem_row_ptr": (C.c_int, [P, I]),
        "twb_problem_bounds": (C.c_int, [P, D, D, D, D]),
        "twb_problem_x0": (C.c_int, [P, D]),
        "twb_problem_has_cost": (C.c_int, [P]),
        "twb_problem_goal_instances": (C.c_int, [P, C.c_int, D, D, D, D]),
        "twb_layout_num_variable_sets": (C.c_int, [P]),
        "twb_layout_variable_set": (C.c_int, [P, C.c_int, C.c_char_p, C.c_int, I, I]),
        "twb_layout_num_constraint_sets": (C.c_int, [P]),
        "twb_layout_constraint_set": (C.c_int, [P, C.c_int, C.c_char_p, C.c_int, I, I]),
        "twb_batch_create": (C.c_int, [P, C.c_int, C.c_int, C.POINTER(P)]),
        "twb_batch_destroy": (None, [P]),
        "twb_batch_set_terrains": (C.c_int, [P, I]),
        "twb_batch_set_grid_terrain": (C.c_int, [P, D, C.c_int, C.c_int]),
        "twb_batch_set_grid_map": (C.c_int, [P, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]),
        "twb_batch_eval_device": (C.c_int, [P, P, P, P, P, P, P, C.c_uint, P]),
        "twb_batch_eval_host": (C.c_int, [P, P, P, P, P, P, P, C.c_uint]),
        "twb_problem_trajectory_dims": (C.c_int, [P, C.c_double, I, I]),
        "twb_batch_sample_trajectory_host": (C.c_int, [P, P, C.c_double, P]),
        "twb_batch_initial_guess_host": (C.c_int, [P, P, P, C.c_int, P]),
        "twb_problem_footstep_plan_dims": (C.c_int, [P, I, I]),
        "twb_batch_footstep_plan_host": (C.c_int, [P, P, C.c_double, P, P]),
        "twb_batch_nearest_planes_host": (C.c_int, [P, P, P, P, C.c_int, P, P]),
        "twb_batch_linear_equality_host": (C.c_int, [P, P, C.c_int, P, C.c_int, P]),
        "twb_batch_soft_constraint_host": (C.c_int, [P, C.c_int, P, P, P]),
        "twb_batch_launches_per_eval": (C.c_int, [P, C.c_uint]),
        "twb_last_error": (C.c_char_p, []),
        "twb_version": (C.c_char_p, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib, list(sig)


lib, EXPORTS = _load()


class TowrB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"towr_b200 error {code}: {msg}")
        self.code = code


def check(rc):
    if rc != OK:
        raise TowrB200Error(rc, lib.twb_last_error().decode())
