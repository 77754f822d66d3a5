// device_tables.h — plain-old-data tables the host builder (formulation.cc)
// uploads once per structure class and the CUDA kernels (kernels.cu) read on
// every evaluation.  Nothing in here depends on the iterate x.
#ifndef TOWR_B200_DEVICE_TABLES_H_
#define TOWR_B200_DEVICE_TABLES_H_

#include <cstdint>

namespace twb {

constexpr int kMaxEE = 4;

// One spline evaluated at one constraint sample when phase durations are
// fixed: the active polynomial (Spline::GetSegmentID, spline.cc:48-63), its
// local time (Spline::GetLocalTime, spline.cc:66-78) and where its two
// boundary nodes live in x.  `xi` holds x indices of p0[3], v0[3], p1[3],
// v1[3]; node values that are not optimised (always 0 in towr) point at the
// zero slot, index n.
struct alignas(16) SplineSample {   // 96 bytes = six 16-byte loads
  double T, T2, T3;   // polynomial duration and std::pow(T,2), std::pow(T,3)
  double rT2, rT3;    // correctly rounded 1/T2, 1/T3 (seed of the exact division on the device)
  double t, t2, t3;   // local time and std::pow(t,2), std::pow(t,3)
  int16_t xi[12];
  int16_t pad[4];
};

// One spline-kernel work item: evaluate spline sample `sample` into state rows scratch..
// kind 0: position (3 doubles); 1: position + acceleration (6); 2: position + velocity + acceleration (9)
struct EvalItem {
  int32_t sample;
  int16_t scratch;
  int16_t kind;
};

// TerrainConstraint row (terrain_constraint.cc:59-108): one ee-motion node
struct TerrainUnit {
  int16_t xi[3];      // x index of node position x,y,z
  int16_t pad;
  int32_t g_row;      // constraint row
  int32_t s0;         // first of the row's 3 CSR slots
};

// ForceConstraint node (force_constraint.cc:64-171): 5 rows
struct ForceUnit {
  int16_t xf[3];      // x index of the force node value
  int16_t xp[3];      // x index of the stance-foot position (phase start node); [2] unused
  int16_t pad[2];
  int32_t g_row;      // first of the 5 rows
  int32_t s0;         // first of the 25 CSR slots of these rows
};

// SwingConstraint node (swing_constraint.cc:57-83): 4 rows
struct SwingUnit {
  int16_t xc_p[2], xc_v[2];  // current node pos/vel x,y
  int16_t xprev[2], xnext[2];
  int32_t g_row;
  int32_t pad;
};

// SplineAccConstraint junction (spline_acc_constraint.cc:49-65), fixed durations
struct AccUnit {
  double Tp, Tp2, Tp3, rTp2, rTp3;  // previous polynomial duration, pow 2, pow 3, reciprocals
  double Tn, Tn2, rTn2;             // next polynomial
  int32_t x0;           // x index of node j, dim 0 position (NodesVariablesAll layout)
  int32_t g_row;        // first of 3 rows
};

// CSR slot range and first constraint row of one dynamic sample (6 rows) / one RoM sample (3 rows per foot)
struct DynInfo { int32_t s0, s1, g_row, pad; };
struct RomInfo { int32_t s0[kMaxEE], s1[kMaxEE], g_row[kMaxEE]; };
// run of CSR slots whose values do not depend on the iterate (SplineAcc, Swing rows): value = coef
struct ConstSeg { int32_t s0, s1; };

// NodeCost term (node_cost.cc:53-76) flattened: one entry per node value that
// enters the cost; var >= 0 when the value is an optimisation variable.
struct CostEntry {
  int16_t xi;        // x index holding the node value (zero slot if fixed)
  int16_t grad_col;  // column of the gradient this node contributes to, or -1
  int32_t pad;
  double weight;
};

// Jacobian slot descriptor: value = S[desc] * coef

struct Plan {
  int n, m, nnz, n_ee;
  // sizes
  int n_dyn, n_rom, n_terr, n_force, n_swing, n_acc, n_totdur, n_cost, n_eval_items, n_const_seg, n_const_runs;
  int max_dyn_slots, max_rom_slots;   // largest number of CSR slots one dynamic / RoM sample owns
  // g rows
  int dyn_row0;
  int rom_row0[kMaxEE];
  int totdur_row0;
  // rows of the spline-value matrix ST: per dynamic sample [c, c_dd, th, th_d, th_dd, p_e.., f_e..],
  // per RoM sample [c, th, p_e..]
  int S_size, S_dyn0, S_dyn_stride, S_rom0, S_rom_stride;
  // robot
  double mass, gravity;
  double I_b[9];
  double mu;
  // tables (device pointers)
  const SplineSample* samples;      // every (constraint sample, spline) pair of the dynamic and RoM sets
  const EvalItem* eval_items;       // [n_eval_items]
  const TerrainUnit* terr;
  const ForceUnit* force;
  const SwingUnit* swing;
  const AccUnit* acc;
  const CostEntry* cost;
  const double* dyn_ang_basis;  // [n_dyn][12]: base-ang basis of the active polynomial, {pos, vel, acc} x {p0, v0, p1, v1}
  const uint32_t* desc;   // [nnz padded to even]  row of the slot's value in its unit's local state block (0 = the constant 1)
  const double* coef;     // [nnz padded to even]  constant of every CSR slot
  const DynInfo* dyn_info;   // [n_dyn]
  const RomInfo* rom_info;   // [n_rom]
  const ConstSeg* const_seg; // [n_const_seg]
};

}  // namespace twb
#endif
