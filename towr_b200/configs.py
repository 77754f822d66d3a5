"""The named workloads of BASELINE.json (`configs`) as NlpFormulation recipes,
plus the synthetic-iterate generator shared by tests and bench.py
(x_b = x0 + sigma*N(0,1), sigma by variable kind).
"""
import numpy as np

from . import _capi as capi
from ._spec import motion_velocity_mask
from .formulation import GaitGenerator, NlpFormulation, robot_info


def _standing(f, robot, goal_xy=(1.0, 0.0), goal_yaw=0.0):
    """Robot standing at the origin in its nominal stance; goal pose ahead."""
    info = robot_info(robot)
    z0 = -info["nominal_stance"][0][2]
    f.initial_base_.lin.p[:] = (0.0, 0.0, z0)
    f.initial_ee_W_ = [np.array([p[0], p[1], 0.0]) for p in info["nominal_stance"]]
    f.final_base_.lin.p[:] = (goal_xy[0], goal_xy[1], z0)
    f.final_base_.ang.p[:] = (0.0, 0.0, goal_yaw)
    return f


def make_formulation(name, goal_xy=None, goal_yaw=0.0, terrain=None, t_total=2.0):
    if name.endswith("_base_rom"):
        # every Parameters::ConstraintName at once: the default list plus BaseRom (BaseMotionConstraint)
        f = make_formulation(name[:-len("_base_rom")], goal_xy, goal_yaw, terrain, t_total)
        f.params_.constraints_.insert(2, capi.C_BASE_ROM)
        return f
    if name == "hopper":
        # towr/test/hopper_example.cc:47-68
        f = NlpFormulation(capi.MONOPED, capi.FLAT)
        f.initial_base_.lin.p[2] = 0.5
        f.initial_ee_W_ = [np.zeros(3)]
        f.final_base_.lin.p[:] = (1.0, 0.0, 0.5)
        f.params_.ee_phase_durations_ = [[0.4, 0.2, 0.4, 0.2, 0.4, 0.2, 0.2]]
        f.params_.ee_in_contact_at_start_ = [True]
        return f
    recipes = {
        # name: (robot, terrain, gait combo, default goal, optimise durations)
        "anymal_trot_block": (capi.ANYMAL, capi.BLOCK, 1, (1.5, 0.0), False),
        "biped_walk_stairs": (capi.BIPED, capi.STAIRS, 0, (1.5, 0.0), False),
        "hyq_gallop_gap": (capi.HYQ, capi.GAP, 4, (2.0, 0.0), True),
        "anymal_trot_mixed": (capi.ANYMAL, capi.SLOPE, 1, (1.5, 0.0), False),
        "go1_trot_flat": (capi.GO1, capi.FLAT, 1, (0.5, 0.0), False),   # fpowr recipe, footstep_plan_server.cc:152-220
    }
    robot, terr, combo, goal, opt = recipes[name]
    f = NlpFormulation(robot, terr if terrain is None else terrain)
    _standing(f, robot, goal if goal_xy is None else goal_xy, goal_yaw)
    n_ee = robot_info(robot)["n_ee"]
    gg = GaitGenerator.MakeGaitGenerator(n_ee)
    gg.SetCombo(combo)
    for ee in range(n_ee):
        f.params_.ee_phase_durations_.append(gg.GetPhaseDurations(t_total, ee))
        f.params_.ee_in_contact_at_start_.append(gg.IsInContactAtStart(ee))
    if opt:
        f.params_.OptimizePhaseDurations()
    return f


CONFIGS = {
    # BASELINE.json configs[i] -> (recipe, batch size)
    "config1_hopper": ("hopper", 1),
    "config2_anymal_trot_block_4096": ("anymal_trot_block", 4096),
    "config3_biped_walk_stairs_16384": ("biped_walk_stairs", 16384),
    "config4_hyq_gallop_gap_durations_32768": ("hyq_gallop_gap", 32768),
    "config5_anymal_trot_mixed_65536": ("anymal_trot_mixed", 65536),
}


def variable_sigma(problem):
    """Per-variable noise scale: 0.05 (positions m / angles rad), 0.2 (velocities),
    10 N (forces), 50 N/s (force rates)."""
    sig = np.empty(problem.n)
    spec = problem.spec
    for name, start, count in problem.variable_sets():
        idx = np.arange(count)
        if name.startswith("base-"):
            s = np.where((idx % 6) < 3, 0.05, 0.2)
        elif name.startswith("ee-motion"):
            vel = np.array(motion_velocity_mask(spec, int(name[len("ee-motion_"):])))
            assert vel.size == count
            s = np.where(vel, 0.2, 0.05)
        elif name.startswith("ee-force"):
            s = np.where((idx % 2) == 0, 10.0, 50.0)
        else:                            # ee-schedule durations are perturbed multiplicatively by the caller
            s = np.zeros(count)
        sig[start:start + count] = s
    return sig


def _renormalise(problem, set_name, d):
    """Scale the optimised phase durations of one foot (last axis) so that the remaining last phase keeps at least
    2 % of the horizon (phase_durations.cc:92 needs sum < T)."""
    ee = int(set_name[len("ee-schedule"):])
    spec = problem.spec
    t_total = sum(spec.phase_durations[ee][i] for i in range(spec.n_phases[ee]))
    tot = d.sum(axis=-1, keepdims=True)
    return d * np.minimum(1.0, 0.98 * t_total / tot)


def synthetic_iterates(problem, batch, seed=1234, first=0):
    """(batch, n) float64 iterates; instance b uses numpy default_rng(seed + first + b)."""
    x0 = problem.GetVariableValues()
    sig = variable_sigma(problem)
    X = np.empty((batch, problem.n))
    for b in range(batch):
        rng = np.random.default_rng(seed + first + b)
        X[b] = x0 + sig * rng.standard_normal(problem.n)
        for name, start, count in problem.variable_sets():
            if name.startswith("ee-schedule"):   # durations: nominal x U[0.9, 1.1], re-normalised so that their sum stays below T
                X[b, start:start + count] = _renormalise(problem, name, x0[start:start + count] * rng.uniform(0.9, 1.1, count))
    for name, start, count in problem.variable_sets():
        if name == "base-ang":           # keep pitch/yaw nodes within +-1 rad (far from gimbal lock)
            blk = X[:, start:start + count].reshape(batch, -1, 6)
            blk[:, :, :3] = np.clip(blk[:, :, :3], -1.0, 1.0)
    return X


def synthetic_iterates_fast(problem, batch, seed=1234, x0=None):
    """Same distribution, one generator for the whole batch (for the large bench batches).  `x0`: optional (batch, n)
    per-instance initial guesses (goal-randomised instances, Problem.goal_instances)."""
    x0 = problem.GetVariableValues() if x0 is None else np.asarray(x0)
    sig = variable_sigma(problem)
    rng = np.random.default_rng(seed)
    X = x0 + sig * rng.standard_normal((batch, problem.n))
    for name, start, count in problem.variable_sets():
        if name.startswith("ee-schedule"):
            X[:, start:start + count] = _renormalise(problem, name, x0[..., start:start + count] * rng.uniform(0.9, 1.1, (batch, count)))
        if name == "base-ang":
            blk = X[:, start:start + count].reshape(batch, -1, 6)
            blk[:, :, :3] = np.clip(blk[:, :, :3], -1.0, 1.0)
    return X
