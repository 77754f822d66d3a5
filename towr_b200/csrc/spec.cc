// spec.cc — setup-time helpers of the C ABI: towr::Parameters defaults, the
// predefined gait tables, robot constants and analytic terrain heights.
// Host only; nothing here runs per iterate.
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/towr_b200.h"
#include "formulation.h"

namespace twb {

// ---- robots: towr/src/robot_model.cc:41-68, models/examples/*_model.h, models/go1/go1_model.h
bool GetRobot(int id, RobotConst* r) {
  struct Row { int n_ee; double mass; double I[6]; double nx, ny, nz; double dev[3]; };
  static const Row rows[] = {
      /* Monoped monoped_model.h:45-58 */ {1, 20.0, {1.2, 5.5, 6.0, 0.0, -0.2, -0.01}, 0.0, 0.0, -0.58, {0.25, 0.15, 0.2}},
      /* Biped   biped_model.h:46-64   */ {2, 20.0, {1.209, 5.583, 6.056, 0.005, -0.190, -0.012}, 0.0, 0.20, -0.65, {0.25, 0.15, 0.15}},
      /* HyQ     hyq_model.h:46-66     */ {4, 83.0, {4.26, 8.97, 9.88, -0.0063, 0.193, 0.0126}, 0.31, 0.29, -0.58, {0.25, 0.20, 0.10}},
      /* Anymal  anymal_model.h:46-67  */ {4, 29.5, {0.946438, 1.94478, 2.01835, 0.000938112, -0.00595386, -0.00146328}, 0.34, 0.19, -0.42, {0.15, 0.1, 0.10}},
      /* Go1     go1_model.h:21-52     */ {4, 12.84, {0.0168128557, 0.063009565, 0.0716547275, -0.0002296769, -0.0002945293, -0.0000418731}, 0.1881, 0.04675 + 0.08, -0.3, {0.16, 0.12, 0.06}},
  };
  if (id < 0 || id > TWB_GO1) return false;
  const Row& w = rows[id];
  r->n_ee = w.n_ee; r->mass = w.mass;
  for (int i = 0; i < 6; ++i) r->inertia[i] = w.I[i];
  for (int i = 0; i < 3; ++i) r->max_dev[i] = w.dev[i];
  // endeffector_mappings.h:43-44: biped L,R ; quadruped LF,RF,LH,RH
  const double sx[4] = {+1, +1, -1, -1}, sy[4] = {+1, -1, +1, -1};
  for (int e = 0; e < w.n_ee; ++e) {
    double x = w.nx, y = w.ny;
    if (w.n_ee == 4) { x = sx[e] * w.nx; y = sy[e] * w.ny; }
    if (w.n_ee == 2) { y = (e == 0 ? +1 : -1) * w.ny; }
    r->nominal[e][0] = x; r->nominal[e][1] = y; r->nominal[e][2] = w.nz;
  }
  return true;
}

// ---- terrains: height_map_examples.cc:35-211 (+ constants in height_map_examples.h)
double TerrainHeight(int id, double x, double y) {
  double h = 0.0;
  switch (id) {
    case TWB_BLOCK: {
      const double start = 0.7, eps = 0.03, len = 3.5, height = 0.5;
      if (start <= x && x <= start + eps) h = (height / eps) * (x - start);
      if (start + eps <= x && x <= start + len) h = height;
      break; }
    case TWB_STAIRS:
      if (x >= 1.0) h = 0.2;
      if (x >= 1.0 + 0.4) h = 0.4;
      if (x >= 1.0 + 0.4 + 1.0) h = 0.0;
      break;
    case TWB_GAP: {
      const double gs = 1.0, w = 0.5, hh = 1.5, xc = gs + w / 2.0;
      const double a = (4 * hh) / (w * w), b = -(8 * hh * xc) / (w * w), c = -(hh * (w - 2 * xc) * (w + 2 * xc)) / (w * w);
      if (gs <= x && x <= gs + w) h = a * x * x + b * x + c;
      break; }
    case TWB_SLOPE: {
      const double s0 = 1.0, up = 1.0, down = 1.0, hc = 0.7, slope = hc / up;
      if (x >= s0) h = slope * (x - s0);
      if (x >= s0 + up) h = hc - slope * (x - (s0 + up));
      if (x >= (s0 + up) + down) h = 0.0;
      break; }
    case TWB_CHIMNEY:
      if (1.0 <= x && x <= 1.0 + 1.5) h = 3.0 * (y - 0.5);
      break;
    case TWB_CHIMNEY_LR:
      if (0.5 <= x && x <= 0.5 + 1.0) h = 2 * (y - 0.5);
      if (0.5 + 1.0 <= x && x <= 0.5 + 2 * 1.0) h = -2 * (y + 0.5);
      break;
    default: break;  // FlatGround(0.0)
  }
  return h;
}

// ---- gaits: gait_generator.cc:54-144 and the three generators' stride tables
namespace {
struct Stride { std::vector<double> t; std::vector<unsigned> c; };  // contact bitmask, bit ee

Stride DropTransition(Stride s) {  // GaitGenerator::RemoveTransition, gait_generator.cc:131-144
  double last = s.t.back();
  s.t.pop_back(); s.t.back() += last; s.c.pop_back();
  return s;
}

// quadruped: bit0 LF, bit1 RF, bit2 LH, bit3 RH (quadruped_gait_generator.cc:38-75)
constexpr unsigned LF = 1, RF = 2, LH = 4, RH = 8;
constexpr unsigned II = 0, IP = LF, Pb = LH | RF, bP = RH | LF, BI = LH | RH, IB = LF | RF, PP = LH | LF,
                   bb = RH | RF, Bb = LH | RH | RF, BP = LH | RH | LF, bB = RH | LF | RF, PB = LH | LF | RF, BB = 15;

enum Gait { Stand = 0, Flight, Walk1, Walk2, Walk2E, Run2, Run2E, Run1, Run1E, Run3, Run3E,
            Hop1, Hop1E, Hop2, Hop3, Hop3E, Hop5, Hop5E };

bool QuadStride(int g, Stride* s) {  // quadruped_gait_generator.cc:89-366
  switch (g) {
    case Stand:  *s = {{0.3}, {BB}}; return true;
    case Flight: *s = {{0.3}, {Bb}}; return true;
    case Walk1:  *s = {{0.3, 0.2, 0.3, 0.2, 0.3, 0.2, 0.3, 0.2}, {bB, BB, Bb, BB, PB, BB, BP, BB}}; return true;
    case Walk2:
    case Walk2E: *s = {{0.25, 0.13, 0.25, 0.13, 0.25, 0.13, 0.25, 0.13}, {bB, bb, Bb, Pb, PB, PP, BP, bP}};
                 if (g == Walk2E) *s = DropTransition(*s); return true;
    case Run1:   *s = {{0.3, 0.2, 0.3, 0.2}, {bP, BB, Pb, BB}}; return true;
    case Run2:   *s = {{0.4, 0.1, 0.4, 0.1}, {bP, II, Pb, II}}; return true;
    case Run2E:  *s = {{0.4}, {bP}}; return true;
    case Run3:   *s = {{0.3, 0.1, 0.3, 0.1}, {PP, II, bb, II}}; return true;
    case Run3E:  *s = {{0.3}, {PP}}; return true;
    case Hop1:   *s = {{0.3, 0.1, 0.3, 0.1}, {BI, II, IB, II}}; return true;
    case Hop1E:  *s = {{0.3}, {BI}}; return true;
    case Hop2:   *s = {{0.3, 0.4, 0.3}, {BB, II, BB}}; return true;
    case Hop3:
    case Hop3E:  *s = {{0.2, 0.3, 0.2, 0.2, 0.2, 0.3, 0.2, 0.2}, {Bb, BI, BP, bP, bB, IB, PB, Pb}};
                 if (g == Hop3E) *s = DropTransition(*s); return true;
    case Hop5:   *s = {{0.1, 0.2, 0.1, 0.1, 0.2, 0.1}, {Bb, BB, IP, Bb, BB, IP}}; return true;
    default: return false;
  }
}
// biped: bit0 L, bit1 R (biped_gait_generator.cc:39-48): P_ = L only, b_ = R only
bool BipedStride(int g, Stride* s) {  // biped_gait_generator.cc:63-226
  const unsigned I = 0, P = 1, b = 2, B = 3;
  switch (g) {
    case Stand:  *s = {{0.2}, {B}}; return true;
    case Flight: *s = {{0.5}, {I}}; return true;
    case Walk1: case Walk2: *s = {{0.3, 0.05, 0.3, 0.05}, {b, B, P, B}}; return true;
    case Run1: case Run3:   *s = {{0.15, 0.4, 0.15 + 0.15, 0.4, 0.15}, {b, I, P, I, b}}; return true;
    case Hop1:   *s = {{0.15, 0.5, 0.15}, {B, I, B}}; return true;
    case Hop2:   *s = {{0.15, 0.4, 0.15}, {b, I, b}}; return true;
    case Hop3:   *s = {{0.2, 0.2, 0.2}, {P, I, P}}; return true;
    case Hop5:   *s = {{0.2, 0.3, 0.2, 0.2}, {P, I, b, B}}; return true;
    default: return false;
  }
}
bool MonopedStride(int g, Stride* s) {  // monoped_gait_generator.cc:52-120
  switch (g) {
    case Stand:  *s = {{0.5}, {1}}; return true;
    case Flight: *s = {{0.5}, {0}}; return true;
    case Hop1:   *s = {{0.3, 0.3}, {1, 0}}; return true;
    case Hop2:   *s = {{0.2, 0.3}, {1, 0}}; return true;
    default: return false;
  }
}
bool ComboGaits(int n_ee, int combo, std::vector<int>* gaits) {
  if (n_ee == 4) {  // quadruped_gait_generator.cc:77-87
    switch (combo) {
      case 0: *gaits = {Stand, Walk2, Walk2, Walk2, Walk2E, Stand}; return true;
      case 1: *gaits = {Stand, Run2, Run2, Run2, Run2E, Stand}; return true;
      case 2: *gaits = {Stand, Run3, Run3, Run3, Run3E, Stand}; return true;
      case 3: *gaits = {Stand, Hop1, Hop1, Hop1, Hop1E, Stand}; return true;
      case 4: *gaits = {Stand, Hop3, Hop3, Hop3, Hop3E, Stand}; return true;
    }
  } else if (n_ee == 2) {  // biped_gait_generator.cc:50-61
    switch (combo) {
      case 0: *gaits = {Stand, Walk1, Walk1, Walk1, Walk1, Stand}; return true;
      case 1: *gaits = {Stand, Run1, Run1, Run1, Run1, Stand}; return true;
      case 2: *gaits = {Stand, Hop1, Hop1, Hop1, Stand}; return true;
      case 3: *gaits = {Stand, Hop1, Hop2, Hop2, Stand}; return true;
      case 4: *gaits = {Stand, Hop5, Hop5, Hop5, Stand}; return true;
    }
  } else if (n_ee == 1) {  // monoped_gait_generator.cc:38-50
    switch (combo) {
      case 0: *gaits = {Stand, Hop1, Hop1, Hop1, Hop1, Stand}; return true;
      case 1: *gaits = {Stand, Hop1, Hop1, Hop1, Stand}; return true;
      case 2: *gaits = {Stand, Hop1, Hop1, Hop1, Hop1, Stand}; return true;
      case 3: *gaits = {Stand, Hop2, Hop2, Hop2, Stand}; return true;
      case 4: *gaits = {Stand, Hop2, Hop2, Hop2, Hop2, Hop2, Stand}; return true;
    }
  }
  return false;
}
}  // namespace

// GaitGenerator::SetGaits + GetPhaseDurations(T, ee) + IsInContactAtStart
bool GaitPhases(int n_ee, int combo, double t_total, std::vector<std::vector<double>>* durations,
                std::vector<bool>* contact_at_start) {
  std::vector<int> gaits;
  if (!ComboGaits(n_ee, combo, &gaits)) return false;
  std::vector<double> times; std::vector<unsigned> contacts;
  for (int g : gaits) {
    Stride s;
    bool ok = n_ee == 4 ? QuadStride(g, &s) : n_ee == 2 ? BipedStride(g, &s) : MonopedStride(g, &s);
    if (!ok) return false;
    times.insert(times.end(), s.t.begin(), s.t.end());
    contacts.insert(contacts.end(), s.c.begin(), s.c.end());
  }
  durations->assign(n_ee, {});
  contact_at_start->assign(n_ee, false);
  for (int ee = 0; ee < n_ee; ++ee) {
    // gait_generator.cc:76-105: merge consecutive global phases with equal contact flag
    std::vector<double> d; double acc = 0.0;
    for (size_t ph = 0; ph + 1 < contacts.size(); ++ph) {
      acc += times[ph];
      bool cur = (contacts[ph] >> ee) & 1u, nxt = (contacts[ph + 1] >> ee) & 1u;
      if (cur != nxt) { d.push_back(acc); acc = 0.0; }
    }
    d.push_back(acc + times.back());
    // :64-74 normalise by the foot's own total, :54-62 scale to t_total
    double total = 0.0; for (double v : d) total += v;
    for (double& v : d) v = v / total;
    for (double& v : d) v = v * t_total;
    (*durations)[ee] = d;
    (*contact_at_start)[ee] = (contacts.front() >> ee) & 1u;
  }
  return true;
}

}  // namespace twb

extern "C" {

int twb_spec_default(twb_spec* s, int robot) {
  if (!s) return TWB_ERR_INVALID;
  twb::RobotConst rc;
  if (!twb::GetRobot(robot, &rc)) return TWB_ERR_INVALID;
  std::memset(s, 0, sizeof(*s));
  s->robot = robot; s->terrain = TWB_FLAT; s->n_ee = 0;
  // parameters.cc:40-73
  s->duration_base_polynomial = 0.1;
  s->force_polynomials_per_stance_phase = 3;
  s->ee_polynomials_per_swing_phase = 2;
  s->force_limit_in_normal_direction = 1000;
  s->dt_constraint_range_of_motion = 0.08;
  s->dt_constraint_dynamic = 0.1;
  s->dt_constraint_base_motion = s->duration_base_polynomial / 4.;
  s->bound_phase_duration_min = 0.2; s->bound_phase_duration_max = 1.0;
  const int order[6] = {TWB_C_TERRAIN, TWB_C_DYNAMIC, TWB_C_BASE_ACC, TWB_C_EE_ROM, TWB_C_FORCE, TWB_C_SWING};
  s->n_constraints = 6; for (int i = 0; i < 6; ++i) s->constraints[i] = order[i];
  s->n_costs = 0;
  s->bounds_final_lin_pos[0] = s->bounds_final_lin_pos[1] = 1; s->bounds_final_lin_pos[2] = 0;
  for (int i = 0; i < 3; ++i) s->bounds_final_lin_vel[i] = s->bounds_final_ang_pos[i] = s->bounds_final_ang_vel[i] = 1;
  return TWB_OK;
}

int twb_spec_optimize_phase_durations(twb_spec* s) {
  if (!s || s->n_constraints >= TWB_MAX_CONSTRAINTS) return TWB_ERR_INVALID;
  s->constraints[s->n_constraints++] = TWB_C_TOTAL_TIME;
  return TWB_OK;
}

int twb_spec_set_gait(twb_spec* s, int n_ee, int combo, double t_total) {
  if (!s) return TWB_ERR_INVALID;
  std::vector<std::vector<double>> d; std::vector<bool> c;
  if (!twb::GaitPhases(n_ee, combo, t_total, &d, &c)) return TWB_ERR_INVALID;
  s->n_ee = n_ee;
  for (int ee = 0; ee < n_ee; ++ee) {
    if ((int)d[ee].size() > TWB_MAX_PHASES) return TWB_ERR_INVALID;
    s->n_phases[ee] = (int)d[ee].size();
    for (size_t i = 0; i < d[ee].size(); ++i) s->phase_durations[ee][i] = d[ee][i];
    s->in_contact_at_start[ee] = c[ee] ? 1 : 0;
  }
  return TWB_OK;
}

int twb_robot_info(int robot, int* n_ee, double* mass, double inertia6[6],
                   double nominal_stance[TWB_MAX_EE][3], double max_dev[3]) {
  twb::RobotConst rc;
  if (!twb::GetRobot(robot, &rc)) return TWB_ERR_INVALID;
  if (n_ee) *n_ee = rc.n_ee;
  if (mass) *mass = rc.mass;
  if (inertia6) for (int i = 0; i < 6; ++i) inertia6[i] = rc.inertia[i];
  if (nominal_stance) for (int e = 0; e < rc.n_ee; ++e) for (int i = 0; i < 3; ++i) nominal_stance[e][i] = rc.nominal[e][i];
  if (max_dev) for (int i = 0; i < 3; ++i) max_dev[i] = rc.max_dev[i];
  return TWB_OK;
}

double twb_terrain_height(int terrain, double x, double y) { return twb::TerrainHeight(terrain, x, y); }

}  // extern "C"
