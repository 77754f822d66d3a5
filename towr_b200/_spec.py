"""struct twb_spec and the enums of include/towr_b200.h as ctypes definitions.  Pure data: importing this module does NOT
load libtowr_b200.so (bench.py's reference arm and the fixture scripts load it by file path, without the package)."""
import ctypes as C

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


