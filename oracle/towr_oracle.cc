// towr_oracle.cc — CPU restatement of towr's NLP evaluation path.
//
// *** TEST INFRASTRUCTURE ONLY.  Nothing in towr_b200/ (the product) may
// *** include, link, import or execute this file.  It is the checker for
// *** tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// *** --impl reference legs.
//
// PARITY UNPINNED: the reference's own tests hold no golden vectors for this
// path (towr/test/dynamic_constraint_test.cc:40-43 and dynamic_model_test.cc:
// 36-49 are empty) and the reference cannot be compiled in this image (needs
// Eigen 3, ifopt, catkin/roscpp — towr/CMakeLists.txt:5-10).  The oracle is
// pinned instead by (1) central finite differences of its own g against its
// own Jacobian, (2) hand-verified structure counts, (3) an independent
// mpmath restatement of the closed-form math (tests/test_oracle_*.py).
//
// Third-party arithmetic restated here because it is absent from
// /root/reference:
//   * Eigen 3.3.x (libeigen3-dev of Ubuntu 20.04, Dockerfile:27): RowMajor
//     SparseMatrix pattern semantics — coeffRef() inserts a structural entry,
//     sparseView(1.0,-1.0) keeps every entry, default sparseView() drops exact
//     zeros, +/- is a pattern union, scalar*sparse keeps the pattern,
//     sparse*sparse is conservative (no pruning), setFromTriplets sums
//     duplicates and keeps explicit zeros, inner indices ascending.  `Sp` below.
//   * ifopt 2.x (KaiNakamura/ifopt HEAD, un-pinned, Dockerfile:97-106):
//     ConstraintSet::GetJacobian (FillJacobianBlock per variable set on a fresh
//     block -> triplets with column offset -> setFromTriplets),
//     Composite::GetJacobian (row offsets), Composite::GetValues (cost terms are
//     summed into one row), Bounds/inf=1e20.  `Problem` below.
//
// The code follows the reference's operation sequence literally (function by
// function, citing file:line) so that both the sparsity pattern and the
// floating-point association are emergent properties of the same steps, as
// they are in the reference.  It is written for fidelity, not speed.
#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <limits>
#include <cstdio>
#include <cstring>
#include <memory>
#include <numeric>
#include <string>
#include <utility>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/towr_b200.h"  // twb_spec layout + enums only (plain C data)
typedef double twb_f64;   // IEEE double in every build (see below)

// Extended-precision build (oracle/Makefile: libtowr_oracle_ld.so, -DTWB_ORACLE_LONG_DOUBLE): every `double` below this
// line becomes the x87 80-bit long double (64-bit mantissa), math calls resolve to the std:: long double overloads, and the
// extern "C" entry points take numpy.longdouble arrays.  Used by tests/test_oracle.py only, to differentiate g numerically
// at ~1e-12 accuracy (Richardson-extrapolated central differences) and pin the analytic Jacobian of the double build to
// <= 1e-10 — four orders tighter than a double-precision finite difference can.  twb_spec (above) stays plain double.
#ifdef TWB_ORACLE_LONG_DOUBLE
typedef long double twb_ld;
#define double twb_ld
#endif
// the time grids decide the STRUCTURE (sample counts, polynomial counts): their arithmetic stays IEEE double in both builds

namespace {
#ifdef TWB_ORACLE_LONG_DOUBLE   // unqualified calls below must not fall back to the C library's double versions
inline double sin(double x) { return std::sin(x); }
inline double cos(double x) { return std::cos(x); }
inline double sqrt(double x) { return std::sqrt(x); }
inline double fabs(double x) { return std::fabs(x); }
inline double floor(double x) { return std::floor(x); }
inline double pow(double x, double y) { return std::pow(x, y); }
#endif

constexpr double kInf = 1e20;  // ifopt::inf
struct Bound { double lo, up; };
const Bound kNoBound{-kInf, +kInf}, kBoundZero{0.0, 0.0}, kBoundGreaterZero{0.0, +kInf},
    kBoundSmallerZero{-kInf, 0.0};

enum { X = 0, Y = 1, Z = 2 };
enum { kPos = 0, kVel = 1, kAcc = 2 };
enum { AX = 0, AY, AZ, LX, LY, LZ };

using V3 = std::array<double, 3>;
struct M3 { double a[3][3]; };

inline V3 vadd(const V3& a, const V3& b) { return {a[0] + b[0], a[1] + b[1], a[2] + b[2]}; }
inline V3 vsub(const V3& a, const V3& b) { return {a[0] - b[0], a[1] - b[1], a[2] - b[2]}; }
inline V3 vscale(double s, const V3& a) { return {s * a[0], s * a[1], s * a[2]}; }
// Eigen cross3 (Geometry/OrthoMethods.h): (a1 b2 - a2 b1, a2 b0 - a0 b2, a0 b1 - a1 b0)
inline V3 vcross(const V3& a, const V3& b) {
  return {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
}
inline double vdot(const V3& a, const V3& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// ---------------------------------------------------------------------------
// Sp: Eigen::SparseMatrix<double, RowMajor> restated (pattern semantics above)
// ---------------------------------------------------------------------------
using Row = std::vector<std::pair<int, double>>;  // ascending column

Row radd(const Row& a, const Row& b) {  // a + b, union of patterns
  Row o; o.reserve(a.size() + b.size());
  size_t i = 0, j = 0;
  while (i < a.size() || j < b.size()) {
    if (j >= b.size() || (i < a.size() && a[i].first < b[j].first)) { o.push_back(a[i]); ++i; }
    else if (i >= a.size() || b[j].first < a[i].first) { o.emplace_back(b[j].first, 0.0 + b[j].second); ++j; }
    else { o.emplace_back(a[i].first, a[i].second + b[j].second); ++i; ++j; }
  }
  return o;
}
Row rsub(const Row& a, const Row& b) {  // a - b
  Row o; o.reserve(a.size() + b.size());
  size_t i = 0, j = 0;
  while (i < a.size() || j < b.size()) {
    if (j >= b.size() || (i < a.size() && a[i].first < b[j].first)) { o.push_back(a[i]); ++i; }
    else if (i >= a.size() || b[j].first < a[i].first) { o.emplace_back(b[j].first, 0.0 - b[j].second); ++j; }
    else { o.emplace_back(a[i].first, a[i].second - b[j].second); ++i; ++j; }
  }
  return o;
}
Row rscale(double s, const Row& a) { Row o = a; for (auto& e : o) e.second = s * e.second; return o; }
Row rscale_r(const Row& a, double s) { Row o = a; for (auto& e : o) e.second = e.second * s; return o; }

struct Sp {
  int nr = 0, nc = 0;
  std::vector<Row> r;
  Sp() = default;
  Sp(int rows, int cols) : nr(rows), nc(cols), r(rows) {}
  double& ref(int i, int j) {  // coeffRef: inserts a structural zero if absent
    Row& R = r[i];
    auto it = std::lower_bound(R.begin(), R.end(), j, [](const std::pair<int, double>& e, int c) { return e.first < c; });
    if (it == R.end() || it->first != j) it = R.insert(it, {j, 0.0});
    return it->second;
  }
  size_t nnz() const { size_t s = 0; for (auto& R : r) s += R.size(); return s; }
};

Sp sp_add(const Sp& a, const Sp& b) {
  assert(a.nr == b.nr && a.nc == b.nc);
  Sp o(a.nr, a.nc);
  for (int i = 0; i < a.nr; ++i) o.r[i] = radd(a.r[i], b.r[i]);
  return o;
}
Sp sp_sub(const Sp& a, const Sp& b) {
  assert(a.nr == b.nr && a.nc == b.nc);
  Sp o(a.nr, a.nc);
  for (int i = 0; i < a.nr; ++i) o.r[i] = rsub(a.r[i], b.r[i]);
  return o;
}
Sp sp_scale(double s, const Sp& a) { Sp o = a; for (auto& R : o.r) for (auto& e : R) e.second = s * e.second; return o; }
Sp sp_neg(const Sp& a) { Sp o = a; for (auto& R : o.r) for (auto& e : R) e.second = -e.second; return o; }
// conservative sparse*sparse: result(i,c) = sum_k a(i,k) b(k,c), k ascending, no pruning
Sp sp_mul(const Sp& a, const Sp& b) {
  assert(a.nc == b.nr);
  Sp o(a.nr, b.nc);
  std::vector<double> acc(b.nc, 0.0);
  std::vector<char> mark(b.nc, 0);
  std::vector<int> cols;
  for (int i = 0; i < a.nr; ++i) {
    cols.clear();
    for (auto& ak : a.r[i]) {
      for (auto& bc : b.r[ak.first]) {
        if (!mark[bc.first]) { mark[bc.first] = 1; acc[bc.first] = ak.second * bc.second; cols.push_back(bc.first); }
        else acc[bc.first] += ak.second * bc.second;
      }
    }
    std::sort(cols.begin(), cols.end());
    o.r[i].reserve(cols.size());
    for (int c : cols) { o.r[i].emplace_back(c, acc[c]); mark[c] = 0; }
  }
  return o;
}
Sp sp_transpose(const Sp& a) {
  Sp o(a.nc, a.nr);
  for (int i = 0; i < a.nr; ++i) for (auto& e : a.r[i]) o.r[e.first].emplace_back(i, e.second);
  return o;
}
Sp sp_row(const Sp& a, int i) { Sp o(1, a.nc); o.r[0] = a.r[i]; return o; }
// jac.middleRows(start, src.nr) = src
void sp_set_rows(Sp& dst, int start, const Sp& src) {
  assert(dst.nc == src.nc);
  for (int k = 0; k < src.nr; ++k) dst.r[start + k] = src.r[k];
}
// M.sparseView(1.0,-1.0) keeps all; M.sparseView() drops exact zeros
Sp sp_from_dense3(const M3& m, bool keep_zeros) {
  Sp o(3, 3);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
    if (keep_zeros || m.a[i][j] != 0.0) o.r[i].emplace_back(j, m.a[i][j]);
  return o;
}
Sp sp_rowvec_full(const V3& v) { Sp o(1, 3); for (int j = 0; j < 3; ++j) o.r[0].emplace_back(j, v[j]); return o; }
V3 sp_mulvec3(const Sp& a, const V3& v) {
  V3 o{0, 0, 0};
  for (int i = 0; i < 3; ++i) { double s = 0.0; bool first = true; for (auto& e : a.r[i]) { if (first) { s = e.second * v[e.first]; first = false; } else s += e.second * v[e.first]; } o[i] = s; }
  return o;
}
M3 sp_to_dense3(const Sp& a) {
  M3 m{}; for (int i = 0; i < 3; ++i) for (auto& e : a.r[i]) m.a[i][e.first] = e.second; return m;
}
M3 m3_transpose(const M3& m) { M3 t; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) t.a[i][j] = m.a[j][i]; return t; }
V3 m3_mulvec(const M3& m, const V3& v) {
  V3 o; for (int i = 0; i < 3; ++i) o[i] = m.a[i][0] * v[0] + m.a[i][1] * v[1] + m.a[i][2] * v[2]; return o;
}

// ---------------------------------------------------------------------------
// Terrains — towr/src/height_map.cc:52-163, height_map_examples.cc:35-211,
// constants towr/include/towr/terrain/examples/height_map_examples.h:45-166
// ---------------------------------------------------------------------------
// towr::HeightMapFromCSV (towr/include/towr/terrain/height_map_from_csv.h:29-111): the grid the reference reads from
// a CSV file is process-global test data here (oracle_set_grid); grid_(y_cell, x_cell), row-major.
struct CsvGrid {
  std::vector<double> h; int rows = 0, cols = 0;
  const double res = 0.17;        // res_m_p_cell_, :113
  const double eps = 0.17 / 50;   // eps_, :114
  // static_cast<size_t>(x / res), :31-32: the conversion truncates toward zero, so a quotient in (-1, 0) is cell 0 (defined
  // behaviour, INSIDE the grid); a quotient <= -1 (or NaN) converts to a huge cell index in practice (x86-64: cvttsd2si
  // of the negative value reinterpreted as unsigned): "outside the grid"
  bool Cell(double x, double y, long long* xc, long long* yc) const {
    double fx = x / res, fy = y / res;
    if (!(fx > -1.0) || !(fy > -1.0)) return false;
    *xc = (long long)fx; *yc = (long long)fy;
    return *xc < cols && *yc < rows;      // isCellValid, :47-49
  }
  double At(long long yc, long long xc) const { return h[(size_t)yc * cols + xc]; }
  double Height(double x, double y) const {  // :29-37
    long long xc, yc; if (!Cell(x, y, &xc, &yc)) return 0.0;
    return At(yc, xc);
  }
  double Hx(double x, double y) const {  // :40-74
    long long xc, yc; if (!Cell(x, y, &xc, &yc)) return 0.0;
    if (xc + 1 < cols) {
      double diff_end = At(yc, xc + 1) - At(yc, xc), x_end = (xc + 1) * res;
      if ((diff_end > 0) && (x <= x_end) && (x >= x_end - eps)) return diff_end / eps;
    }
    if (xc - 1 >= 0) {
      double diff_start = At(yc, xc) - At(yc, xc - 1), x_start = xc * res;
      if ((diff_start < 0) && (x >= x_start) && (x <= x_start + eps)) return diff_start / eps;
    }
    return 0.0;
  }
  double Hy(double x, double y) const {  // :77-111
    long long xc, yc; if (!Cell(x, y, &xc, &yc)) return 0.0;
    if (yc + 1 < rows) {
      double diff_end = At(yc + 1, xc) - At(yc, xc), y_end = (yc + 1) * res;
      if ((diff_end > 0) && (y <= y_end) && (y >= y_end - eps)) return diff_end / eps;
    }
    if (yc - 1 >= 0) {
      double diff_start = At(yc, xc) - At(yc - 1, xc), y_start = yc * res;
      if ((diff_start < 0) && (y >= y_start) && (y <= y_start + eps)) return diff_start / eps;
    }
    return 0.0;
  }
};
CsvGrid g_grid;

// towr `Grid` (towr/include/towr/terrain/grid_height_map.h:16-59, the terrain fpowr instantiates at
// footstep_plan_server.cc:155): heights come from grid_map::GridMap::atPosition("elevation", p, INTER_LINEAR) in FLOAT,
// FLT_MAX outside the map (the caught std::out_of_range, :41-45), slopes are central differences of those floats with
// eps = resolution / 6 (:24, :49-61).  grid_map is an un-vendored ROS dependency (ros-noetic-grid-map 1.6.x); its
// algorithm is restated from grid_map_core/src/{GridMap.cpp, GridMapMath.cpp}: atPosition -> atPositionLinearInterpolated
// (bilinear over the four cell centres around p, weights in double, result stored to float) and, when one of the four
// cells is outside, INTER_NEAREST (getIndex + at), else out_of_range.  Geometry: cell (ix, iy) has its centre at
// pos + 0.5 * length - 0.5 * res - res * (ix, iy) (index 0 = largest coordinate); buffer start index (0, 0) assumed.
// Data: row-major [size_x][size_y] floats (layer(ix, iy)).  Parity unpinned (no grid_map source or build in this image).
struct GridMapTerrain {
  std::vector<float> h; int sx = 0, sy = 0; double res = 1.0, px = 0.0, py = 0.0;
  double Lx() const { return sx * res; }
  double Ly() const { return sy * res; }
  bool WithinMap(double x, double y) const {   // checkIfPositionWithinMap
    const double tx = -(x - px - 0.5 * Lx()), ty = -(y - py - 0.5 * Ly());
    return tx >= 0.0 && ty >= 0.0 && tx < Lx() && ty < Ly();
  }
  void IndexOf(double x, double y, int* ix, int* iy) const {   // getIndexFromPosition without the range checks: (int) truncates toward zero
    const double vx = (x - 0.5 * Lx() - px) / res, vy = (y - 0.5 * Ly() - py) / res;
    *ix = (int)(-vx); *iy = (int)(-vy);
  }
  void CentreOf(int ix, int iy, double* x, double* y) const { *x = px + (0.5 * Lx() - 0.5 * res) + res * (double)(-ix); *y = py + (0.5 * Ly() - 0.5 * res) + res * (double)(-iy); }
  bool Linear(double x, double y, float* value) const {   // atPositionLinearInterpolated
    int ix[4], iy[4]; size_t shift[4]; double cx, cy;
    IndexOf(x, y, &ix[0], &iy[0]);
    CentreOf(ix[0], iy[0], &cx, &cy);
    bool dir;
    if (x >= cx) { ix[1] = ix[0] - 1; iy[1] = iy[0]; dir = true; } else { ix[1] = ix[0] + 1; iy[1] = iy[0]; dir = false; }
    if (y >= cy) {
      ix[2] = ix[0]; iy[2] = iy[0] - 1;
      if (dir) { shift[0] = 0; shift[1] = 1; shift[2] = 2; shift[3] = 3; } else { shift[0] = 1; shift[1] = 0; shift[2] = 3; shift[3] = 2; }
    } else {
      ix[2] = ix[0]; iy[2] = iy[0] + 1;
      if (dir) { shift[0] = 2; shift[1] = 3; shift[2] = 0; shift[3] = 1; } else { shift[0] = 3; shift[1] = 2; shift[2] = 1; shift[3] = 0; }
    }
    ix[3] = ix[1]; iy[3] = iy[2];
    const size_t buffer = (size_t)sx * sy;
    float f[4];
    for (int i = 0; i < 4; ++i) {
      // getLinearIndexFromIndex (column-major Eigen matrix: iy * size_x + ix) in size_t: a negative ix with iy >= 1 wraps into
      // the previous column like in grid_map; linear indices >= the buffer size are rejected here (grid_map's `>` lets
      // index == size through, which reads past the buffer: undefined there)
      const size_t lin = (size_t)((long long)iy[shift[i]] * sx + ix[shift[i]]);
      if (lin >= buffer) return false;
      f[i] = h[(lin % (size_t)sx) * sy + lin / (size_t)sx];
    }
    CentreOf(ix[shift[0]], iy[shift[0]], &cx, &cy);
    const double rx = (x - cx) / res, ry = (y - cy) / res, fx = 1.0 - rx, fy = 1.0 - ry;
    *value = f[0] * fx * fy + f[1] * rx * fy + f[2] * fx * ry + f[3] * rx * ry;
    return true;
  }
  double Height(double x, double y) const {   // grid_height_map.h:29-46
    float height;
    if (Linear(x, y, &height)) return height;
    int ix, iy; IndexOf(x, y, &ix, &iy);
    if (WithinMap(x, y) && ix >= 0 && iy >= 0 && ix < sx && iy < sy) { height = h[(size_t)ix * sy + iy]; return height; }
    height = std::numeric_limits<float>::max();
    return height;
  }
  double Hx(double x, double y) const { const double eps = res / 6.0; float r = (float)Height(x + eps, y), l = (float)Height(x - eps, y); return (r - l) / (2 * eps); }   // :48-53
  double Hy(double x, double y) const { const double eps = res / 6.0; float u = (float)Height(x, y + eps), d = (float)Height(x, y - eps); return (u - d) / (2 * eps); }   // :55-60
};
GridMapTerrain g_gridmap;

struct Terrain {
  int id = TWB_FLAT;
  double flat_height = 0.0;
  static constexpr double mu = 0.5;  // height_map.h:136

  double Height(double x, double y) const {
    switch (id) {
      case TWB_GRID_CSV: return g_grid.Height(x, y);
      case TWB_GRID_MAP: return g_gridmap.Height(x, y);
      case TWB_FLAT: return flat_height;
      case TWB_BLOCK: {  // height_map_examples.cc:40-53
        const double block_start = 0.7, length = 3.5, height = 0.5, eps = 0.03; const double slope = height / eps;
        double h = 0.0;
        if (block_start <= x && x <= block_start + eps) h = slope * (x - block_start);
        if (block_start + eps <= x && x <= block_start + length) h = height;
        return h; }
      case TWB_STAIRS: {  // :69-84
        const double first_step_start = 1.0, first_step_width = 0.4, h1 = 0.2, h2 = 0.4, width_top = 1.0;
        double h = 0.0;
        if (x >= first_step_start) h = h1;
        if (x >= first_step_start + first_step_width) h = h2;
        if (x >= first_step_start + first_step_width + width_top) h = 0.0;
        return h; }
      case TWB_GAP: {  // :88-98
        double h = 0.0; GapC g;
        if (g.gap_start <= x && x <= g.gap_end_x) h = g.a * x * x + g.b * x + g.c;
        return h; }
      case TWB_SLOPE: {  // :124-140
        SlopeC s; double z = 0.0;
        if (x >= s.slope_start) z = s.slope * (x - s.slope_start);
        if (x >= s.x_down_start) z = s.height_center - s.slope * (x - s.x_down_start);
        if (x >= s.x_flat_start) z = 0.0;
        return z; }
      case TWB_CHIMNEY: {  // :161-170
        const double x_start = 1.0, length = 1.5, y_start = 0.5, slope = 3.0; const double x_end = x_start + length;
        double z = 0.0;
        if (x_start <= x && x <= x_end) z = slope * (y - y_start);
        return z; }
      case TWB_CHIMNEY_LR: {  // :185-197
        const double x_start = 0.5, length = 1.0, y_start = 0.5, slope = 2; const double x_end1 = x_start + length, x_end2 = x_start + 2 * length;
        double z = 0.0;
        if (x_start <= x && x <= x_end1) z = slope * (y - y_start);
        if (x_end1 <= x && x <= x_end2) z = -slope * (y + y_start);
        return z; }
    }
    return 0.0;
  }
  double Hx(double x, double y) const {
    switch (id) {
      case TWB_GRID_CSV: return g_grid.Hx(x, y);
      case TWB_GRID_MAP: return g_gridmap.Hx(x, y);
      case TWB_BLOCK: { const double block_start = 0.7, height = 0.5, eps = 0.03; const double slope = height / eps;
        double d = 0.0; if (block_start <= x && x <= block_start + eps) d = slope; return d; }  // :55-65
      case TWB_GAP: { GapC g; double d = 0.0; if (g.gap_start <= x && x <= g.gap_end_x) d = 2 * g.a * x + g.b; return d; }  // :100-109
      case TWB_SLOPE: { SlopeC s; double d = 0.0;  // :142-157
        if (x >= s.slope_start) d = s.slope;
        if (x >= s.x_down_start) d = -s.slope;
        if (x >= s.x_flat_start) d = 0.0;
        return d; }
      default: return 0.0;
    }
  }
  double Hy(double x, double y) const {
    switch (id) {
      case TWB_GRID_CSV: return g_grid.Hy(x, y);
      case TWB_GRID_MAP: return g_gridmap.Hy(x, y);
      case TWB_CHIMNEY: { const double x_start = 1.0, length = 1.5, slope = 3.0; const double x_end = x_start + length;
        double d = 0.0; if (x_start <= x && x <= x_end) d = slope; return d; }  // :172-181
      case TWB_CHIMNEY_LR: { const double x_start = 0.5, length = 1.0, slope = 2; const double x_end1 = x_start + length, x_end2 = x_start + 2 * length;
        double d = 0.0;  // :199-211
        if (x_start <= x && x <= x_end1) d = slope;
        if (x_end1 <= x && x <= x_end2) d = -slope;
        return d; }
      default: return 0.0;
    }
  }
  double Hxx(double x, double) const {  // only Gap overrides, :111-120
    if (id == TWB_GAP) { GapC g; double d = 0.0; if (g.gap_start <= x && x <= g.gap_end_x) d = 2 * g.a; return d; }
    return 0.0;
  }
  // height_map.cc:52-60
  double DerivOfHeightWrt(int dim, double x, double y) const { return dim == X ? Hx(x, y) : Hy(x, y); }
  // height_map.cc:148-163 (XY, YX, YY never overridden -> 0)
  double SecondDeriv(int d1, int d2, double x, double y) const { return (d1 == X && d2 == X) ? Hxx(x, y) : 0.0; }

  // height_map.cc:93-138; deriv<0: the basis itself, else d/d(deriv)
  V3 Basis(int basis, double x, double y, int deriv) const {
    bool req = deriv < 0;
    V3 v;
    if (basis == 0) {  // Normal
      for (int dim : {X, Y}) v[dim] = req ? -DerivOfHeightWrt(dim, x, y) : -SecondDeriv(dim, deriv, x, y);
      v[Z] = req ? 1.0 : 0.0;
    } else if (basis == 1) {  // Tangent1
      v[X] = req ? 1.0 : 0.0; v[Y] = 0.0;
      v[Z] = req ? DerivOfHeightWrt(X, x, y) : SecondDeriv(X, deriv, x, y);
    } else {  // Tangent2
      v[X] = 0.0; v[Y] = req ? 1.0 : 0.0;
      v[Z] = req ? DerivOfHeightWrt(Y, x, y) : SecondDeriv(Y, deriv, x, y);
    }
    return v;
  }
  static V3 Normalized(const V3& v) {  // Eigen MatrixBase::normalized()
    double z = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    if (z > 0.0) { double n = std::sqrt(z); return {v[0] / n, v[1] / n, v[2] / n}; }
    return v;
  }
  V3 NormalizedBasis(int basis, double x, double y) const { return Normalized(Basis(basis, x, y, -1)); }  // :62-66
  // height_map.cc:80-91 — element-wise product of ONE column of the
  // normalisation Jacobian with dv/d(dim), as in the reference
  V3 DerivOfNormalizedBasisWrt(int basis, int dim, double x, double y) const {
    V3 dv = Basis(basis, x, y, dim);
    V3 v = Basis(basis, x, y, -1);
    // :140-146  1/|v|^2 * (|v| e_idx - v(idx) * v.normalized())
    double sn = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    double nrm = std::sqrt(sn);
    V3 vh = Normalized(v);
    V3 o;
    for (int i = 0; i < 3; ++i) {
      double unit = (i == dim) ? 1.0 : 0.0;
      double d = 1 / sn * (nrm * unit - v[dim] * vh[i]);
      o[i] = d * dv[i];
    }
    return o;
  }

 private:
  struct GapC {  // height_map_examples.h:96-111
    const double gap_start = 1.0, w = 0.5, h = 1.5;
    const double dx = w / 2.0; const double xc = gap_start + dx; const double gap_end_x = gap_start + w;
    const double a = (4 * h) / (w * w);
    const double b = -(8 * h * xc) / (w * w);
    const double c = -(h * (w - 2 * xc) * (w + 2 * xc)) / (w * w);
  };
  struct SlopeC {  // height_map_examples.h:123-130
    const double slope_start = 1.0, up_length = 1.0, down_length = 1.0, height_center = 0.7;
    const double x_down_start = slope_start + up_length; const double x_flat_start = x_down_start + down_length;
    const double slope = height_center / up_length;
  };
};

// ---------------------------------------------------------------------------
// Robots — towr/src/robot_model.cc:41-68 + models/examples/*.h, go1/go1_model.h
// ---------------------------------------------------------------------------
struct Robot { int n_ee; double mass; double I[6]; V3 nominal[4]; V3 max_dev; };
Robot MakeRobot(int id) {
  Robot r{};
  auto quad = [&](double x, double y, double z) {
    r.nominal[0] = {x, y, z}; r.nominal[1] = {x, -y, z}; r.nominal[2] = {-x, y, z}; r.nominal[3] = {-x, -y, z}; };
  switch (id) {
    case TWB_MONOPED: r = {1, 20, {1.2, 5.5, 6.0, 0.0, -0.2, -0.01}, {}, {0.25, 0.15, 0.2}}; r.nominal[0] = {0.0, 0.0, -0.58}; break;
    case TWB_BIPED: r = {2, 20, {1.209, 5.583, 6.056, 0.005, -0.190, -0.012}, {}, {0.25, 0.15, 0.15}};
      r.nominal[0] = {0.0, 0.20, -0.65}; r.nominal[1] = {0.0, -0.20, -0.65}; break;
    case TWB_HYQ: r = {4, 83, {4.26, 8.97, 9.88, -0.0063, 0.193, 0.0126}, {}, {0.25, 0.20, 0.10}}; quad(0.31, 0.29, -0.58); break;
    case TWB_ANYMAL: r = {4, 29.5, {0.946438, 1.94478, 2.01835, 0.000938112, -0.00595386, -0.00146328}, {}, {0.15, 0.1, 0.10}}; quad(0.34, 0.19, -0.42); break;
    case TWB_GO1: r = {4, 12.84, {0.0168128557, 0.063009565, 0.0716547275, -0.0002296769, -0.0002945293, -0.0000418731}, {}, {0.16, 0.12, 0.06}};
      quad(0.1881, 0.04675 + 0.08, -0.3); break;
    default: r.n_ee = 0;
  }
  return r;
}

// ---------------------------------------------------------------------------
// Variables
// ---------------------------------------------------------------------------
struct Info { int id, deriv, dim; };  // NodesVariables::NodeValueInfo
struct Spline;                          // observer

struct VarSet {
  std::string name;
  virtual ~VarSet() = default;
  virtual int Rows() const = 0;
  virtual void SetVariables(const double* x) = 0;
  virtual void GetValues(double* x) const = 0;
  virtual void GetBounds(Bound* b) const = 0;
};

// towr/src/nodes_variables.cc, nodes_variables_all.cc, nodes_variables_phase_based.cc
struct NodeVars : VarSet {
  struct PolyInfo { int phase, poly_in_phase, n_polys_in_phase; bool is_constant; };
  std::vector<std::array<V3, 2>> nodes;      // [node][kPos|kVel]
  std::vector<std::vector<Info>> infos;      // index -> node values it sets
  std::vector<Bound> bounds;
  std::vector<PolyInfo> poly;                // phase-based only
  std::vector<Spline*> observers;

  int Rows() const override { return (int)infos.size(); }
  int PolyCount() const { return (int)nodes.size() - 1; }
  // nodes_variables.cc:40-50
  int GetOptIndex(int id, int deriv, int dim) const {
    for (int idx = 0; idx < Rows(); ++idx)
      for (auto& n : infos[idx]) if (n.id == id && n.deriv == deriv && n.dim == dim) return idx;
    return -1;
  }
  void GetValues(double* x) const override {  // :53-61 (last info wins)
    for (int idx = 0; idx < Rows(); ++idx) for (auto& n : infos[idx]) x[idx] = nodes[n.id][n.deriv][n.dim];
  }
  void SetVariables(const double* x) override;  // :64-72
  void GetBounds(Bound* b) const override { for (int i = 0; i < Rows(); ++i) b[i] = bounds[i]; }
  // :126-150
  void SetByLinearInterpolation(const V3& initial_val, const V3& final_val, double t_total) {
    V3 dp = vsub(final_val, initial_val);
    V3 avg = {dp[0] / t_total, dp[1] / t_total, dp[2] / t_total};
    int num_nodes = (int)nodes.size();
    for (int idx = 0; idx < Rows(); ++idx) for (auto& n : infos[idx]) {
      if (n.deriv == kPos) {
        double f = n.id / static_cast<double>(num_nodes - 1);
        V3 pos = {initial_val[0] + f * dp[0], initial_val[1] + f * dp[1], initial_val[2] + f * dp[2]};
        nodes[n.id][kPos][n.dim] = pos[n.dim];
      }
      if (n.deriv == kVel) nodes[n.id][kVel][n.dim] = avg[n.dim];
    }
  }
  void AddBound(int id, int deriv, int dim, double val) {  // :161-168
    for (int idx = 0; idx < Rows(); ++idx) for (auto& n : infos[idx])
      if (n.id == id && n.deriv == deriv && n.dim == dim) bounds[idx] = Bound{val, val};
  }
  void AddStartBound(int deriv, const std::vector<int>& dims, const V3& v) { for (int d : dims) AddBound(0, deriv, d, v[d]); }
  void AddFinalBound(int deriv, const std::vector<int>& dims, const V3& v) { for (int d : dims) AddBound((int)nodes.size() - 1, deriv, d, v[d]); }

  // ---- NodesVariablesAll (nodes_variables_all.cc:34-61)
  static std::unique_ptr<NodeVars> MakeAll(int n_nodes, const std::string& name) {
    auto v = std::make_unique<NodeVars>(); v->name = name;
    v->nodes.assign(n_nodes, {V3{0, 0, 0}, V3{0, 0, 0}});
    int n = n_nodes * 2 * 3;
    v->infos.resize(n); v->bounds.assign(n, kNoBound);
    for (int idx = 0; idx < n; ++idx) {
      int internal = idx % 6;
      v->infos[idx].push_back(Info{(int)std::floor(idx / 6), internal < 3 ? kPos : kVel, internal % 3});
    }
    return v;
  }
  // ---- phase based (nodes_variables_phase_based.cc:38-76)
  void BuildPolyInfos(int phase_count, bool first_phase_constant, int n_polys_in_changing_phase) {
    bool c = first_phase_constant;
    for (int i = 0; i < phase_count; ++i) {
      if (c) poly.push_back({i, 0, 1, true});
      else for (int j = 0; j < n_polys_in_changing_phase; ++j) poly.push_back({i, j, n_polys_in_changing_phase, false});
      c = !c;
    }
    nodes.assign(poly.size() + 1, {V3{0, 0, 0}, V3{0, 0, 0}});
  }
  std::vector<double> ConvertPhaseToPolyDurations(const std::vector<double>& ph) const {  // :78-89
    std::vector<double> d;
    for (int i = 0; i < PolyCount(); ++i) d.push_back(ph.at(poly[i].phase) / poly[i].n_polys_in_phase);
    return d;
  }
  std::vector<int> AdjacentPolyIds(int node_id) const {  // :156-172
    int last = (int)nodes.size() - 1;
    if (node_id == 0) return {0};
    if (node_id == last) return {last - 1};
    return {node_id - 1, node_id};
  }
  bool IsConstantNode(int node_id) const {  // :104-116
    bool c = false; for (int p : AdjacentPolyIds(node_id)) if (poly[p].is_constant) c = true; return c;
  }
  std::vector<int> NonConstantNodes() const {  // :124-134
    std::vector<int> ids; for (int i = 0; i < (int)nodes.size(); ++i) if (!IsConstantNode(i)) ids.push_back(i); return ids;
  }
  int GetPhase(int node_id) const { return poly[AdjacentPolyIds(node_id).front()].phase; }  // :136-143
  int NodeIDAtStartOfPhase(int phase) const {  // :145-168
    for (int i = 0; i < (int)poly.size(); ++i) if (poly[i].phase == phase) return i;  // GetNodeId(poly, Start) = poly
    return 0;
  }
  // ---- NodesVariablesEEMotion (:190-253)
  static std::unique_ptr<NodeVars> MakeEEMotion(int phase_count, bool contact_at_start, const std::string& name, int n_polys) {
    auto v = std::make_unique<NodeVars>(); v->name = name;
    v->BuildPolyInfos(phase_count, contact_at_start, n_polys);
    for (int id = 0; id < (int)v->nodes.size(); ++id) {
      if (!v->IsConstantNode(id)) {
        for (int dim = 0; dim < 3; ++dim) {
          v->infos.push_back({Info{id, kPos, dim}});
          if (dim == Z) v->nodes[id][kVel][Z] = 0.0;
          else v->infos.push_back({Info{id, kVel, dim}});
        }
      } else {
        v->nodes[id][kVel] = {0, 0, 0}; v->nodes[id + 1][kVel] = {0, 0, 0};
        for (int dim = 0; dim < 3; ++dim) v->infos.push_back({Info{id, kPos, dim}, Info{id + 1, kPos, dim}});
        id += 1;
      }
    }
    v->bounds.assign(v->infos.size(), kNoBound);
    return v;
  }
  // ---- NodesVariablesEEForce (:255-298)
  static std::unique_ptr<NodeVars> MakeEEForce(int phase_count, bool contact_at_start, const std::string& name, int n_polys) {
    auto v = std::make_unique<NodeVars>(); v->name = name;
    v->BuildPolyInfos(phase_count, !contact_at_start, n_polys);
    for (int id = 0; id < (int)v->nodes.size(); ++id) {
      if (!v->IsConstantNode(id)) {
        for (int dim = 0; dim < 3; ++dim) { v->infos.push_back({Info{id, kPos, dim}}); v->infos.push_back({Info{id, kVel, dim}}); }
      } else {
        v->nodes[id] = {V3{0, 0, 0}, V3{0, 0, 0}}; v->nodes[id + 1] = {V3{0, 0, 0}, V3{0, 0, 0}};
        id += 1;
      }
    }
    v->bounds.assign(v->infos.size(), kNoBound);
    return v;
  }
};

// towr/src/spline.cc:48-78
int GetSegmentID(double t_global, const std::vector<double>& durations) {
  double eps = 1e-10;
  double t = 0; int i = 0;
  for (double d : durations) { t += d; if (t >= t_global - eps) return i; i++; }
  return (int)durations.size() - 1;  // reference: assert(false); Release falls off the end
}

// towr/src/phase_durations.cc
struct Durations : VarSet {
  std::vector<double> durations; double t_total; Bound bound; bool initial_contact;
  std::vector<Spline*> observers;
  Durations(int ee, const std::vector<double>& timings, bool contact, double mn, double mx) {  // :40-52
    name = "ee-schedule" + std::to_string(ee);
    durations = timings; t_total = std::accumulate(timings.begin(), timings.end(), 0.0);
    bound = Bound{mn, mx}; initial_contact = contact;
  }
  int Rows() const override { return (int)durations.size() - 1; }
  void GetValues(double* x) const override { for (int i = 0; i < Rows(); ++i) x[i] = durations[i]; }
  void SetVariables(const double* x) override;  // :79-100
  void GetBounds(Bound* b) const override { for (int i = 0; i < Rows(); ++i) b[i] = bound; }
  // :122-154 — dense 3 x Rows(), converted with sparseView(1.0,-1.0): every entry structural
  Sp JacobianOfPos(int current_phase, const V3& dx_dT, const V3& xd) const {
    int P = Rows();
    std::vector<V3> col(P, V3{0, 0, 0});
    bool in_last = (current_phase == (int)durations.size() - 1);
    if (!in_last) col[current_phase] = dx_dT;
    for (int ph = 0; ph < current_phase; ++ph) {
      col[ph] = {-1 * xd[0], -1 * xd[1], -1 * xd[2]};
      if (in_last) col[ph] = vsub(col[ph], dx_dT);
    }
    Sp o(3, P);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < P; ++j) o.r[i].emplace_back(j, col[j][i]);
    return o;
  }
};

// towr/src/polynomial.cc + spline.cc + node_spline.cc + phase_spline.cc
struct State3 { V3 p{0, 0, 0}, v{0, 0, 0}, a{0, 0, 0}; };
struct Spline {
  NodeVars* nv = nullptr;
  Durations* pd = nullptr;              // non-null => PhaseSpline
  std::vector<double> T;                // poly durations
  std::vector<std::array<V3, 4>> coeff; // A,B,C,D per poly
  std::vector<std::array<std::array<V3, 2>, 2>> bn;  // boundary nodes per poly [side][deriv]
  Sp structure;                         // jac_wrt_nodes_structure_

  // NodeSpline ctor node_spline.cc:36-43 / PhaseSpline ctor phase_spline.cc:35-52
  Spline(NodeVars* nodes, const std::vector<double>& poly_durations, Durations* phase_durations = nullptr)
      : nv(nodes), pd(phase_durations), T(poly_durations) {
    coeff.assign(T.size(), {}); bn.assign(T.size(), {});
    nv->observers.push_back(this);
    UpdateNodes();
    structure = Sp(3, nv->Rows());
    if (pd) {
      pd->observers.push_back(this);
      UpdatePolynomialDurations();
      for (int i = 0; i < nv->PolyCount(); ++i) Fill(i, 0.0, kPos, structure, true);
    }
  }
  void UpdateCoeff() {  // polynomial.cc:97-104
    for (size_t i = 0; i < T.size(); ++i) {
      const V3 &p0 = bn[i][0][kPos], &v0 = bn[i][0][kVel], &p1 = bn[i][1][kPos], &v1 = bn[i][1][kVel];
      double T_ = T[i]; double T2 = std::pow(T_, 2), T3 = std::pow(T_, 3);
      for (int k = 0; k < 3; ++k) {
        coeff[i][0][k] = p0[k];
        coeff[i][1][k] = v0[k];
        coeff[i][2][k] = -(3 * (p0[k] - p1[k]) + T_ * (2 * v0[k] + v1[k])) / T2;
        coeff[i][3][k] = (2 * (p0[k] - p1[k]) + T_ * (v0[k] + v1[k])) / T3;
      }
    }
  }
  void UpdateNodes() {  // node_spline.cc:45-54
    for (size_t i = 0; i < T.size(); ++i) { bn[i][0] = nv->nodes[i]; bn[i][1] = nv->nodes[i + 1]; }
    UpdateCoeff();
  }
  void UpdatePolynomialDurations() {  // phase_spline.cc:54-65
    T = nv->ConvertPhaseToPolyDurations(pd->durations);
    UpdateCoeff();
  }
  std::pair<int, double> LocalTime(double t_global) const {  // spline.cc:66-78
    int id = GetSegmentID(t_global, T);
    double tl = t_global; for (int i = 0; i < id; ++i) tl -= T[i];
    return {id, tl};
  }
  static double DerivWrtCoeff(double t, int deriv, int c) {  // polynomial.cc:63-72
    switch (deriv) {
      case kPos: return std::pow(t, c);
      case kVel: return c >= 1 ? c * std::pow(t, c - 1) : 0.0;
      default: return c >= 2 ? c * (c - 1) * std::pow(t, c - 2) : 0.0;
    }
  }
  State3 GetPoint(int id, double tl) const {  // polynomial.cc:47-61
    State3 s; V3* out[3] = {&s.p, &s.v, &s.a};
    for (int d = 0; d < 3; ++d) for (int c = 0; c < 4; ++c) {
      double w = DerivWrtCoeff(tl, d, c);
      for (int k = 0; k < 3; ++k) (*out[d])[k] += w * coeff[id][c][k];
    }
    return s;
  }
  State3 GetPoint(double t) const { auto lt = LocalTime(t); return GetPoint(lt.first, lt.second); }
  // polynomial.cc:106-234
  double DerivWrtNode(int poly, int side, int dfdt, int node_deriv, double t) const {
    double t2 = std::pow(t, 2), t3 = std::pow(t, 3);
    double T_ = T[poly], T2 = std::pow(T_, 2), T3 = std::pow(T_, 3);
    if (side == 0) {
      if (dfdt == kPos) return node_deriv == kPos ? (2 * t3) / T3 - (3 * t2) / T2 + 1 : t - (2 * t2) / T_ + t3 / T2;
      if (dfdt == kVel) return node_deriv == kPos ? (6 * t2) / T3 - (6 * t) / T2 : (3 * t2) / T2 - (4 * t) / T_ + 1;
      return node_deriv == kPos ? (12 * t) / T3 - 6 / T2 : (6 * t) / T2 - 4 / T_;
    } else {
      if (dfdt == kPos) return node_deriv == kPos ? (3 * t2) / T2 - (2 * t3) / T3 : t3 / T2 - t2 / T_;
      if (dfdt == kVel) return node_deriv == kPos ? (6 * t) / T2 - (6 * t2) / T3 : (3 * t2) / T2 - (2 * t) / T_;
      return node_deriv == kPos ? 6 / T2 - (12 * t) / T3 : (6 * t) / T2 - 2 / T_;
    }
  }
  // node_spline.cc:85-112
  void Fill(int poly_id, double tl, int dxdt, Sp& jac, bool fill_with_zeros) const {
    for (int idx = 0; idx < jac.nc; ++idx)
      for (auto& nvi : nv->infos[idx])
        for (int side : {0, 1}) {
          int node = poly_id + side;
          if (node == nvi.id) {
            double val = DerivWrtNode(poly_id, side, dxdt, nvi.deriv, tl);
            if (fill_with_zeros) val = 0.0;
            jac.ref(nvi.dim, idx) += val;
          }
        }
  }
  Sp JacWrtNodes(int id, double tl, int dxdt) const { Sp j = structure; Fill(id, tl, dxdt, j, false); return j; }  // :71-83
  Sp JacWrtNodes(double t, int dxdt) const { auto lt = LocalTime(t); return JacWrtNodes(lt.first, lt.second, dxdt); }  // :62-69
  // polynomial.cc:236-257
  V3 DerivOfPosWrtDuration(int poly, double t) const {
    const V3 &x0 = bn[poly][0][kPos], &x1 = bn[poly][1][kPos], &v0 = bn[poly][0][kVel], &v1 = bn[poly][1][kVel];
    double t2 = std::pow(t, 2), t3 = std::pow(t, 3);
    double T_ = T[poly], T2 = std::pow(T_, 2), T3 = std::pow(T_, 3), T4 = std::pow(T_, 4);
    V3 d;
    for (int k = 0; k < 3; ++k)
      d[k] = (t3 * (v0[k] + v1[k])) / T3 - (t2 * (2 * v0[k] + v1[k])) / T2
             - (3 * t3 * (2 * x0[k] - 2 * x1[k] + T_ * v0[k] + T_ * v1[k])) / T4
             + (2 * t2 * (3 * x0[k] - 3 * x1[k] + 2 * T_ * v0[k] + T_ * v1[k])) / T3;
    return d;
  }
  // phase_spline.cc:77-93
  V3 DerivOfPosWrtPhaseDuration(double t_global) const {
    auto lt = LocalTime(t_global);
    V3 vel = GetPoint(t_global).v;
    V3 dxdT = DerivOfPosWrtDuration(lt.first, lt.second);
    double inner = 1. / nv->poly[lt.first].n_polys_in_phase;
    double prev = nv->poly[lt.first].poly_in_phase;
    V3 o; for (int k = 0; k < 3; ++k) o[k] = inner * (dxdT[k] - prev * vel[k]);
    return o;
  }
  // phase_spline.cc:67-75
  Sp JacOfPosWrtDurations(double t_global) const {
    V3 dx_dT = DerivOfPosWrtPhaseDuration(t_global);
    V3 xd = GetPoint(t_global).v;
    int cur = GetSegmentID(t_global, pd->durations);
    return pd->JacobianOfPos(cur, dx_dT, xd);
  }
};

void NodeVars::SetVariables(const double* x) {
  for (int idx = 0; idx < Rows(); ++idx) for (auto& n : infos[idx]) nodes[n.id][n.deriv][n.dim] = x[idx];
  for (auto* o : observers) o->UpdateNodes();
}
void Durations::SetVariables(const double* x) {
  double sum = 0.0;  // Eigen x.sum(); association of the packet reduction is not pinned — sequential here
  for (int i = 0; i < Rows(); ++i) sum += x[i];
  for (int i = 0; i < Rows(); ++i) durations[i] = x[i];
  durations.back() = t_total - sum;
  for (auto* o : observers) o->UpdatePolynomialDurations();
}

// ---------------------------------------------------------------------------
// EulerConverter — towr/src/euler_converter.cc
// ---------------------------------------------------------------------------
struct Euler {
  const Spline* s = nullptr;
  int n() const { return s->nv->Rows(); }

  static Sp GetM(const V3& xyz) {  // :133-148
    double z = xyz[Z], y = xyz[Y];
    Sp M(3, 3);
    M.ref(0, Y) = -sin(z); M.ref(0, X) = cos(y) * cos(z);
    M.ref(1, Y) = cos(z);  M.ref(1, X) = cos(y) * sin(z);
    M.ref(2, Z) = 1.0;     M.ref(2, X) = -sin(y);
    return M;
  }
  static Sp GetMdot(const V3& xyz, const V3& xyz_d) {  // :150-166
    double z = xyz[Z], zd = xyz_d[Z], y = xyz[Y], yd = xyz_d[Y];
    Sp Md(3, 3);
    Md.ref(0, Y) = -cos(z) * zd; Md.ref(0, X) = -cos(z) * sin(y) * yd - cos(y) * sin(z) * zd;
    Md.ref(1, Y) = -sin(z) * zd; Md.ref(1, X) = cos(y) * cos(z) * zd - sin(y) * sin(z) * yd;
    Md.ref(2, X) = -cos(y) * yd;
    return Md;
  }
  static Sp RotBaseToWorld(const V3& xyz) {  // :207-221, sparseView(1.0,-1.0): all 9 structural
    double x = xyz[X], y = xyz[Y], z = xyz[Z];
    M3 M;
    M.a[0][0] = cos(y) * cos(z); M.a[0][1] = cos(z) * sin(x) * sin(y) - cos(x) * sin(z); M.a[0][2] = sin(x) * sin(z) + cos(x) * cos(z) * sin(y);
    M.a[1][0] = cos(y) * sin(z); M.a[1][1] = cos(x) * cos(z) + sin(x) * sin(y) * sin(z); M.a[1][2] = cos(x) * sin(y) * sin(z) - cos(z) * sin(x);
    M.a[2][0] = -sin(y);         M.a[2][1] = cos(y) * sin(x);                            M.a[2][2] = cos(x) * cos(y);
    return sp_from_dense3(M, true);
  }
  Sp RotBaseToWorld(double t) const { return RotBaseToWorld(s->GetPoint(t).p); }  // :200-205
  V3 AngVel(double t) const { State3 o = s->GetPoint(t); return sp_mulvec3(GetM(o.p), o.v); }  // :58-70
  V3 AngAcc(double t) const {  // :72-83
    State3 o = s->GetPoint(t);
    return vadd(sp_mulvec3(GetMdot(o.p, o.v), o.v), sp_mulvec3(GetM(o.p), o.a));
  }
  Row GetJac(double t, int deriv, int dim) const { return s->JacWrtNodes(t, deriv).r[dim]; }  // :306-310

  Sp DerivMwrtNodes(double t, int dim) const {  // :168-198
    State3 o = s->GetPoint(t);
    double z = o.p[Z], y = o.p[Y];
    Row jz = GetJac(t, kPos, Z), jy = GetJac(t, kPos, Y);
    Sp jac(3, n());
    switch (dim) {
      case X: jac.r[Y] = rscale(-cos(z), jz);
              jac.r[X] = rsub(rscale(-cos(z) * sin(y), jy), rscale(cos(y) * sin(z), jz)); break;
      case Y: jac.r[Y] = rscale(-sin(z), jz);
              jac.r[X] = rsub(rscale(cos(y) * cos(z), jz), rscale(sin(y) * sin(z), jy)); break;
      case Z: jac.r[X] = rscale(-cos(y), jy); break;
    }
    return jac;
  }
  Sp DerivMdotwrtNodes(double t, int dim) const {  // :270-304
    State3 o = s->GetPoint(t);
    double z = o.p[Z], zd = o.v[Z], y = o.p[Y], yd = o.v[Y];
    Row jz = GetJac(t, kPos, Z), jy = GetJac(t, kPos, Y), jzd = GetJac(t, kVel, Z), jyd = GetJac(t, kVel, Y);
    Sp jac(3, n());
    switch (dim) {
      case X:
        jac.r[Y] = rsub(rscale(sin(z) * zd, jz), rscale(cos(z), jzd));
        jac.r[X] = radd(rsub(rsub(rsub(rsub(rscale(sin(y) * sin(z) * yd, jz), rscale(cos(y) * sin(z), jzd)),
                                            rscale(cos(y) * cos(z) * yd, jy)), rscale(cos(y) * cos(z) * zd, jz)),
                             rscale(cos(z) * sin(y), jyd)), rscale_r(rscale(sin(y) * sin(z), jy), zd));
        break;
      case Y:
        jac.r[Y] = rsub(rscale(-sin(z), jzd), rscale(cos(z) * zd, jz));
        jac.r[X] = rsub(rsub(rsub(rsub(rsub(rscale(cos(y) * cos(z), jzd), rscale(sin(y) * sin(z), jyd)),
                                            rscale(cos(y) * sin(z) * yd, jy)), rscale(cos(z) * sin(y) * yd, jz)),
                             rscale_r(rscale(cos(z) * sin(y), jy), zd)), rscale(cos(y) * sin(z) * zd, jz));
        break;
      case Z:
        jac.r[X] = rsub(rscale(sin(y) * yd, jy), rscale(cos(y), jyd));
        break;
    }
    return jac;
  }
  Sp DerivOfAngVelWrtNodes(double t) const {  // :85-103
    Sp jac(3, n());
    State3 o = s->GetPoint(t);
    Sp vel = sp_rowvec_full(o.v);
    Sp dVel_du = s->JacWrtNodes(t, kVel);
    for (int dim : {X, Y, Z}) {
      Sp dM_du = DerivMwrtNodes(t, dim);
      jac.r[dim] = radd(sp_mul(vel, dM_du).r[0], sp_mul(sp_row(GetM(o.p), dim), dVel_du).r[0]);
    }
    return jac;
  }
  Sp DerivOfAngAccWrtNodes(double t) const {  // :105-131
    Sp jac(3, n());
    State3 o = s->GetPoint(t);
    Sp vel = sp_rowvec_full(o.v), acc = sp_rowvec_full(o.a);
    Sp dVel_du = s->JacWrtNodes(t, kVel), dAcc_du = s->JacWrtNodes(t, kAcc);
    for (int dim : {X, Y, Z}) {
      Sp dMdot_du = DerivMdotwrtNodes(t, dim);
      Sp dM_du = DerivMwrtNodes(t, dim);
      Row a = sp_mul(vel, dMdot_du).r[0];
      Row b = sp_mul(sp_row(GetMdot(o.p, o.v), dim), dVel_du).r[0];
      Row c = sp_mul(acc, dM_du).r[0];
      Row d = sp_mul(sp_row(GetM(o.p), dim), dAcc_du).r[0];
      jac.r[dim] = radd(radd(radd(a, b), c), d);
    }
    return jac;
  }
  // :241-268 — nine row vectors dR_ij/du
  std::array<std::array<Row, 3>, 3> DerivOfRotWrtNodes(double t) const {
    std::array<std::array<Row, 3>, 3> J;
    State3 o = s->GetPoint(t);
    double x = o.p[X], y = o.p[Y], z = o.p[Z];
    Row jx = GetJac(t, kPos, X), jy = GetJac(t, kPos, Y), jz = GetJac(t, kPos, Z);
    auto S = [](double c, const Row& r) { return rscale(c, r); };
    J[X][X] = rsub(S(-cos(z) * sin(y), jy), S(cos(y) * sin(z), jz));
    J[X][Y] = radd(radd(rsub(rsub(S(sin(x) * sin(z), jx), S(cos(x) * cos(z), jz)), S(sin(x) * sin(y) * sin(z), jz)),
                        S(cos(x) * cos(z) * sin(y), jx)), S(cos(y) * cos(z) * sin(x), jy));
    J[X][Z] = radd(rsub(rsub(radd(S(cos(x) * sin(z), jx), S(cos(z) * sin(x), jz)), S(cos(z) * sin(x) * sin(y), jx)),
                        S(cos(x) * sin(y) * sin(z), jz)), S(cos(x) * cos(y) * cos(z), jy));
    J[Y][X] = rsub(S(cos(y) * cos(z), jz), S(sin(y) * sin(z), jy));
    J[Y][Y] = radd(radd(rsub(rsub(S(cos(x) * sin(y) * sin(z), jx), S(cos(x) * sin(z), jz)), S(cos(z) * sin(x), jx)),
                        S(cos(y) * sin(x) * sin(z), jy)), S(cos(z) * sin(x) * sin(y), jz));
    J[Y][Z] = radd(radd(rsub(rsub(S(sin(x) * sin(z), jz), S(cos(x) * cos(z), jx)), S(sin(x) * sin(y) * sin(z), jx)),
                        S(cos(x) * cos(y) * sin(z), jy)), S(cos(x) * cos(z) * sin(y), jz));
    J[Z][X] = S(-cos(y), jy);
    J[Z][Y] = rsub(S(cos(x) * cos(y), jx), S(sin(x) * sin(y), jy));
    J[Z][Z] = rsub(S(-cos(y) * sin(x), jx), S(cos(x) * sin(y), jy));
    return J;
  }
  Sp DerivOfRotVecMult(double t, const V3& v, bool inverse) const {  // :223-239
    auto Rd = DerivOfRotWrtNodes(t);
    Sp jac(3, n());
    for (int row : {X, Y, Z}) for (int col : {X, Y, Z}) {
      const Row& jr = inverse ? Rd[col][row] : Rd[row][col];
      jac.r[row] = radd(jac.r[row], rscale(v[col], jr));
    }
    return jac;
  }
};

// ---------------------------------------------------------------------------
// SingleRigidBodyDynamics — towr/src/single_rigid_body_dynamics.cc, dynamic_model.cc
// ---------------------------------------------------------------------------
Sp Cross(const V3& in) {  // :46-57
  Sp o(3, 3);
  o.ref(0, 1) = -in[2]; o.ref(0, 2) = in[1];
  o.ref(1, 0) = in[2];  o.ref(1, 2) = -in[0];
  o.ref(2, 0) = -in[1]; o.ref(2, 1) = in[0];
  return o;
}
struct Srbd {
  double m = 0, g = 9.80665;  // dynamic_model.cc:37
  Sp I_b;                      // inertia_b.sparseView(): exact zeros dropped (:73)
  V3 com_pos{0, 0, 0}, com_acc{0, 0, 0}, omega{0, 0, 0}, omega_dot{0, 0, 0};
  M3 R{{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}};
  std::vector<V3> ee_force, ee_pos;

  void Init(const Robot& rb) {
    m = rb.mass;
    M3 I;  // :36-44
    I.a[0][0] = rb.I[0];  I.a[0][1] = -rb.I[3]; I.a[0][2] = -rb.I[4];
    I.a[1][0] = -rb.I[3]; I.a[1][1] = rb.I[1];  I.a[1][2] = -rb.I[5];
    I.a[2][0] = -rb.I[4]; I.a[2][1] = -rb.I[5]; I.a[2][2] = rb.I[2];
    I_b = sp_from_dense3(I, false);
    ee_force.assign(rb.n_ee, V3{0, 0, 0}); ee_pos.assign(rb.n_ee, V3{0, 0, 0});
  }
  Sp Iw() const {  // w_R_b_.sparseView() * I_b * w_R_b_.transpose().sparseView()  (:91, :127)
    return sp_mul(sp_mul(sp_from_dense3(R, false), I_b), sp_from_dense3(m3_transpose(R), false));
  }
  std::array<double, 6> DynamicViolation() const {  // :76-101
    V3 f_sum{0, 0, 0}, tau_sum{0, 0, 0};
    for (size_t ee = 0; ee < ee_pos.size(); ++ee) {
      V3 f = ee_force[ee];
      tau_sum = vadd(tau_sum, vcross(f, vsub(com_pos, ee_pos[ee])));
      f_sum = vadd(f_sum, f);
    }
    Sp I_w = Iw();
    V3 a1 = sp_mulvec3(I_w, omega_dot);
    V3 a2 = sp_mulvec3(Cross(omega), sp_mulvec3(I_w, omega));
    V3 grav{0.0, 0.0, -m * g};
    std::array<double, 6> acc;
    for (int k = 0; k < 3; ++k) {
      acc[AX + k] = a1[k] + a2[k] - tau_sum[k];
      acc[LX + k] = m * com_acc[k] - f_sum[k] - grav[k];
    }
    return acc;
  }
  Sp JacWrtBaseLin(const Sp& jac_pos, const Sp& jac_acc) const {  // :103-121
    int n = jac_pos.nc;
    Sp jac_tau_sum(3, n);
    for (const V3& f : ee_force) jac_tau_sum = sp_add(jac_tau_sum, sp_mul(Cross(f), jac_pos));
    Sp jac(6, n);
    sp_set_rows(jac, AX, sp_neg(jac_tau_sum));
    sp_set_rows(jac, LX, sp_scale(m, jac_acc));
    return jac;
  }
  V3 IbRtv(const V3& v) const {  // I_b*w_R_b_.transpose()*v : (sparse*dense) then *vector
    M3 Rt = m3_transpose(R), P{};
    for (int i = 0; i < 3; ++i) for (int k = 0; k < 3; ++k) {
      double s = 0.0; bool first = true;
      for (auto& e : I_b.r[i]) { if (first) { s = e.second * Rt.a[e.first][k]; first = false; } else s += e.second * Rt.a[e.first][k]; }
      P.a[i][k] = s;
    }
    return m3_mulvec(P, v);
  }
  Sp JacWrtBaseAng(const Euler& eu, double t) const {  // :123-165
    Sp I_w = Iw();
    Sp Rs = sp_from_dense3(R, false);
    V3 v11 = IbRtv(omega_dot);
    Sp jac11 = eu.DerivOfRotVecMult(t, v11, false);
    Sp jac12 = sp_mul(sp_mul(Rs, I_b), eu.DerivOfRotVecMult(t, omega_dot, true));
    Sp jac_ang_acc = eu.DerivOfAngAccWrtNodes(t);
    Sp jac13 = sp_mul(I_w, jac_ang_acc);
    Sp jac1 = sp_add(sp_add(jac11, jac12), jac13);

    V3 v21 = IbRtv(omega);
    Sp jac21 = eu.DerivOfRotVecMult(t, v21, false);
    Sp jac22 = sp_mul(sp_mul(Rs, I_b), eu.DerivOfRotVecMult(t, omega, true));
    Sp jac_ang_vel = eu.DerivOfAngVelWrtNodes(t);
    Sp jac23 = sp_mul(I_w, jac_ang_vel);
    Sp jac2 = sp_sub(sp_mul(Cross(omega), sp_add(sp_add(jac21, jac22), jac23)),
                     sp_mul(Cross(sp_mulvec3(I_w, omega)), jac_ang_vel));
    Sp jac(6, jac_ang_vel.nc);
    sp_set_rows(jac, AX, sp_add(jac1, jac2));
    return jac;
  }
  Sp JacWrtForce(const Sp& jac_force, int ee) const {  // :167-179
    V3 r = vsub(com_pos, ee_pos[ee]);
    Sp jac_tau = sp_mul(sp_neg(Cross(r)), jac_force);
    Sp jac(6, jac_force.nc);
    sp_set_rows(jac, AX, sp_neg(jac_tau));
    sp_set_rows(jac, LX, sp_neg(jac_force));
    return jac;
  }
  Sp JacWrtEEPos(const Sp& jac_ee_pos, int ee) const {  // :181-192
    Sp jac_tau = sp_mul(Cross(ee_force[ee]), sp_neg(jac_ee_pos));
    Sp jac(6, jac_tau.nc);
    sp_set_rows(jac, AX, sp_neg(jac_tau));
    return jac;
  }
};

// ---------------------------------------------------------------------------
// Constraint sets
// ---------------------------------------------------------------------------
struct Ctx {  // what SplineHolder + formulation give every constraint
  Spline* base_lin = nullptr; Spline* base_ang = nullptr;
  std::vector<Spline*> ee_motion, ee_force;
  std::vector<NodeVars*> ee_motion_nodes, ee_force_nodes;
  std::vector<Durations*> durations;
  Srbd* model = nullptr; Terrain* terrain = nullptr; Robot robot;
  bool optimize_timings = false;
};

struct CSet {
  std::string name; int rows = 0;
  virtual ~CSet() = default;
  virtual void Values(double* g) const = 0;
  virtual void Bounds(Bound* b) const = 0;
  virtual void FillBlock(const std::string& var_set, Sp& jac) const = 0;
};

std::vector<double> MakeDts(double T, double dt) {  // time_discretization_constraint.cc:37-51
  twb_f64 t = 0.0; const twb_f64 T64 = (twb_f64)T, dt64 = (twb_f64)dt; std::vector<double> dts = {t};
  for (int i = 0; i < std::floor(T64 / dt64); ++i) { t += dt64; dts.push_back(t); }
  dts.push_back(T);
  return dts;
}

struct DynamicC : CSet {  // dynamic_constraint.cc
  const Ctx* c; std::vector<double> dts; Euler eu;
  DynamicC(const Ctx* ctx, double T, double dt) : c(ctx), dts(MakeDts(T, dt)) {
    name = "dynamic"; eu.s = c->base_ang; rows = (int)dts.size() * 6;
  }
  void UpdateModel(double t) const {  // :119-137
    State3 com = c->base_lin->GetPoint(t);
    Srbd& md = *c->model;
    md.com_pos = com.p; md.com_acc = com.a;
    md.R = sp_to_dense3(eu.RotBaseToWorld(t));
    md.omega = eu.AngVel(t); md.omega_dot = eu.AngAcc(t);
    for (size_t ee = 0; ee < c->ee_motion.size(); ++ee) {
      md.ee_force[ee] = c->ee_force[ee]->GetPoint(t).p;
      md.ee_pos[ee] = c->ee_motion[ee]->GetPoint(t).p;
    }
  }
  void Values(double* g) const override {  // :59-64
    int k = 0;
    for (double t : dts) { UpdateModel(t); auto v = c->model->DynamicViolation(); for (int d = 0; d < 6; ++d) g[6 * k + d] = v[d]; ++k; }
  }
  void Bounds(Bound* b) const override { for (int i = 0; i < rows; ++i) b[i] = kBoundZero; }  // :66-71
  void FillBlock(const std::string& vs, Sp& jac) const override {  // :73-117 via time_discretization_constraint.cc:89-96
    int k = 0;
    for (double t : dts) {
      UpdateModel(t);
      int n = jac.nc;
      Sp jm(6, n);
      if (vs == "base-lin") jm = c->model->JacWrtBaseLin(c->base_lin->JacWrtNodes(t, kPos), c->base_lin->JacWrtNodes(t, kAcc));
      if (vs == "base-ang") jm = c->model->JacWrtBaseAng(eu, t);
      for (int ee = 0; ee < (int)c->ee_motion.size(); ++ee) {
        if (vs == "ee-force_" + std::to_string(ee)) jm = c->model->JacWrtForce(c->ee_force[ee]->JacWrtNodes(t, kPos), ee);
        if (vs == "ee-motion_" + std::to_string(ee)) jm = c->model->JacWrtEEPos(c->ee_motion[ee]->JacWrtNodes(t, kPos), ee);
        if (vs == "ee-schedule" + std::to_string(ee)) {
          jm = sp_add(jm, c->model->JacWrtForce(c->ee_force[ee]->JacOfPosWrtDurations(t), ee));
          jm = sp_add(jm, c->model->JacWrtEEPos(c->ee_motion[ee]->JacOfPosWrtDurations(t), ee));
        }
      }
      sp_set_rows(jac, 6 * k, jm);
      ++k;
    }
  }
};

struct RomC : CSet {  // range_of_motion_constraint.cc
  const Ctx* c; std::vector<double> dts; Euler eu; int ee; V3 max_dev, nominal;
  RomC(const Ctx* ctx, double T, double dt, int ee_) : c(ctx), dts(MakeDts(T, dt)), ee(ee_) {
    name = "rangeofmotion-" + std::to_string(ee); eu.s = c->base_ang;
    max_dev = c->robot.max_dev; nominal = c->robot.nominal[ee]; rows = (int)dts.size() * 3;
  }
  void Values(double* g) const override {  // :58-69
    int k = 0;
    for (double t : dts) {
      V3 base_W = c->base_lin->GetPoint(t).p;
      V3 pos_ee_W = c->ee_motion[ee]->GetPoint(t).p;
      Sp b_R_w = sp_transpose(eu.RotBaseToWorld(t));
      V3 r_W = vsub(pos_ee_W, base_W);
      V3 r_B = sp_mulvec3(b_R_w, r_W);
      for (int d = 0; d < 3; ++d) g[3 * k + d] = r_B[d];
      ++k;
    }
  }
  void Bounds(Bound* b) const override {  // :71-81
    for (size_t k = 0; k < dts.size(); ++k) for (int dim = 0; dim < 3; ++dim) {
      Bound bd{0.0, 0.0}; bd.lo += nominal[dim]; bd.up += nominal[dim];
      bd.up += max_dev[dim]; bd.lo -= max_dev[dim];
      b[3 * k + dim] = bd;
    }
  }
  void FillBlock(const std::string& vs, Sp& jac) const override {  // :83-109
    int k = 0;
    for (double t : dts) {
      Sp b_R_w = sp_transpose(eu.RotBaseToWorld(t));
      int row_start = 3 * k;
      if (vs == "base-lin") sp_set_rows(jac, row_start, sp_mul(sp_scale(-1, b_R_w), c->base_lin->JacWrtNodes(t, kPos)));
      if (vs == "base-ang") {
        V3 base_W = c->base_lin->GetPoint(t).p; V3 ee_pos_W = c->ee_motion[ee]->GetPoint(t).p;
        V3 r_W = vsub(ee_pos_W, base_W);
        sp_set_rows(jac, row_start, eu.DerivOfRotVecMult(t, r_W, true));
      }
      if (vs == "ee-motion_" + std::to_string(ee)) sp_set_rows(jac, row_start, sp_mul(b_R_w, c->ee_motion[ee]->JacWrtNodes(t, kPos)));
      if (vs == "ee-schedule" + std::to_string(ee)) sp_set_rows(jac, row_start, sp_mul(b_R_w, c->ee_motion[ee]->JacOfPosWrtDurations(t)));
      ++k;
    }
  }
};

struct BaseMotionC : CSet {  // base_motion_constraint.cc:38-91
  const Ctx* c; std::vector<double> dts; Bound nb[6];
  BaseMotionC(const Ctx* ctx, double T, double dt) : c(ctx), dts(MakeDts(T, dt)) {
    name = "baseMotion";
    double dev_rad = 0.05;
    nb[AX] = Bound{-dev_rad, dev_rad}; nb[AY] = Bound{-dev_rad, dev_rad}; nb[AZ] = kNoBound;
    double z_init = c->base_lin->GetPoint(0.0).p[Z];
    nb[LX] = kNoBound; nb[LY] = kNoBound; nb[LZ] = Bound{z_init - 0.02, z_init + 0.1};
    rows = (int)dts.size() * 6;
  }
  void Values(double* g) const override {
    int k = 0;
    for (double t : dts) {
      V3 l = c->base_lin->GetPoint(t).p, a = c->base_ang->GetPoint(t).p;
      for (int d = 0; d < 3; ++d) { g[6 * k + LX + d] = l[d]; g[6 * k + AX + d] = a[d]; }
      ++k;
    }
  }
  void Bounds(Bound* b) const override { for (size_t k = 0; k < dts.size(); ++k) for (int d = 0; d < 6; ++d) b[6 * k + d] = nb[d]; }
  void FillBlock(const std::string& vs, Sp& jac) const override {
    int k = 0;
    for (double t : dts) {
      if (vs == "base-ang") sp_set_rows(jac, 6 * k + AX, c->base_ang->JacWrtNodes(t, kPos));
      if (vs == "base-lin") sp_set_rows(jac, 6 * k + LX, c->base_lin->JacWrtNodes(t, kPos));
      ++k;
    }
  }
};

struct TerrainC : CSet {  // terrain_constraint.cc
  const Ctx* c; NodeVars* mo; std::vector<int> node_ids;
  TerrainC(const Ctx* ctx, int ee) : c(ctx), mo(ctx->ee_motion_nodes[ee]) {
    name = "terrain-ee-motion_" + std::to_string(ee);
    for (int id = 1; id < (int)mo->nodes.size(); ++id) node_ids.push_back(id);  // :47-57
    rows = (int)node_ids.size();
  }
  void Values(double* g) const override {  // :59-73
    int row = 0;
    for (int id : node_ids) { const V3& p = mo->nodes[id][kPos]; g[row++] = p[Z] - c->terrain->Height(p[X], p[Y]); }
  }
  void Bounds(Bound* b) const override {  // :75-91
    int row = 0;
    for (int id : node_ids) { b[row] = mo->IsConstantNode(id) ? kBoundZero : Bound{0.0, 1e20}; row++; }
  }
  void FillBlock(const std::string& vs, Sp& jac) const override {  // :93-108
    if (vs != mo->name) return;
    int row = 0;
    for (int id : node_ids) {
      jac.ref(row, mo->GetOptIndex(id, kPos, Z)) = 1.0;
      const V3& p = mo->nodes[id][kPos];
      for (int dim : {X, Y}) jac.ref(row, mo->GetOptIndex(id, kPos, dim)) = -c->terrain->DerivOfHeightWrt(dim, p[X], p[Y]);
      row++;
    }
  }
};

struct ForceC : CSet {  // force_constraint.cc
  const Ctx* c; NodeVars *fo, *mo; std::vector<int> ids; double fn_max, mu;
  ForceC(const Ctx* ctx, int ee, double force_limit) : c(ctx), fo(ctx->ee_force_nodes[ee]), mo(ctx->ee_motion_nodes[ee]) {
    name = "force-ee-force_" + std::to_string(ee);
    fn_max = force_limit; mu = Terrain::mu;
    ids = fo->NonConstantNodes();  // :52-62
    rows = (int)ids.size() * 5;
  }
  void Values(double* g) const override {  // :64-92
    int row = 0;
    for (int f_id : ids) {
      int phase = fo->GetPhase(f_id);
      V3 p = mo->nodes[mo->NodeIDAtStartOfPhase(phase)][kPos];
      V3 n = c->terrain->NormalizedBasis(0, p[X], p[Y]);
      V3 f = fo->nodes[f_id][kPos];
      g[row++] = vdot(f, n);
      V3 t1 = c->terrain->NormalizedBasis(1, p[X], p[Y]);
      g[row++] = vdot(f, vsub(t1, vscale(mu, n)));
      g[row++] = vdot(f, vadd(t1, vscale(mu, n)));
      V3 t2 = c->terrain->NormalizedBasis(2, p[X], p[Y]);
      g[row++] = vdot(f, vsub(t2, vscale(mu, n)));
      g[row++] = vdot(f, vadd(t2, vscale(mu, n)));
    }
  }
  void Bounds(Bound* b) const override {  // :94-108
    int row = 0;
    for (size_t i = 0; i < ids.size(); ++i) {
      b[row++] = Bound{0.0, fn_max}; b[row++] = kBoundSmallerZero; b[row++] = kBoundGreaterZero;
      b[row++] = kBoundSmallerZero; b[row++] = kBoundGreaterZero;
    }
  }
  void FillBlock(const std::string& vs, Sp& jac) const override {  // :110-171
    if (vs == fo->name) {
      int row = 0;
      for (int f_id : ids) {
        int phase = fo->GetPhase(f_id);
        V3 p = mo->nodes[mo->NodeIDAtStartOfPhase(phase)][kPos];
        V3 n = c->terrain->NormalizedBasis(0, p[X], p[Y]);
        V3 t1 = c->terrain->NormalizedBasis(1, p[X], p[Y]);
        V3 t2 = c->terrain->NormalizedBasis(2, p[X], p[Y]);
        for (int dim : {X, Y, Z}) {
          int idx = fo->GetOptIndex(f_id, kPos, dim);
          int rr = row;
          jac.ref(rr++, idx) = n[dim];
          jac.ref(rr++, idx) = t1[dim] - mu * n[dim];
          jac.ref(rr++, idx) = t1[dim] + mu * n[dim];
          jac.ref(rr++, idx) = t2[dim] - mu * n[dim];
          jac.ref(rr++, idx) = t2[dim] + mu * n[dim];
        }
        row += 5;
      }
    }
    if (vs == mo->name) {
      int row = 0;
      for (int f_id : ids) {
        int phase = fo->GetPhase(f_id);
        int ee_node_id = mo->NodeIDAtStartOfPhase(phase);
        V3 p = mo->nodes[ee_node_id][kPos];
        V3 f = fo->nodes[f_id][kPos];
        for (int dim : {X, Y}) {
          V3 dn = c->terrain->DerivOfNormalizedBasisWrt(0, dim, p[X], p[Y]);
          V3 dt1 = c->terrain->DerivOfNormalizedBasisWrt(1, dim, p[X], p[Y]);
          V3 dt2 = c->terrain->DerivOfNormalizedBasisWrt(2, dim, p[X], p[Y]);
          int idx = mo->GetOptIndex(ee_node_id, kPos, dim);
          int rr = row;
          jac.ref(rr++, idx) = vdot(f, dn);
          jac.ref(rr++, idx) = vdot(f, vsub(dt1, vscale(mu, dn)));
          jac.ref(rr++, idx) = vdot(f, vadd(dt1, vscale(mu, dn)));
          jac.ref(rr++, idx) = vdot(f, vsub(dt2, vscale(mu, dn)));
          jac.ref(rr++, idx) = vdot(f, vadd(dt2, vscale(mu, dn)));
        }
        row += 5;
      }
    }
  }
};

struct SwingC : CSet {  // swing_constraint.cc
  NodeVars* mo; std::vector<int> ids; double t_swing_avg = 0.3;  // swing_constraint.h:68
  SwingC(const Ctx* ctx, int ee) : mo(ctx->ee_motion_nodes[ee]) {
    name = "swing-ee-motion_" + std::to_string(ee);
    ids = mo->NonConstantNodes(); rows = (int)ids.size() * 2 * 2;  // :44-55
  }
  void Values(double* g) const override {  // :57-83
    int row = 0;
    for (int id : ids) {
      const auto& curr = mo->nodes[id];
      double prev[2] = {mo->nodes[id - 1][kPos][X], mo->nodes[id - 1][kPos][Y]};
      double next[2] = {mo->nodes[id + 1][kPos][X], mo->nodes[id + 1][kPos][Y]};
      double dist[2] = {next[0] - prev[0], next[1] - prev[1]};
      double center[2] = {prev[0] + 0.5 * dist[0], prev[1] + 0.5 * dist[1]};
      double des_vel[2] = {dist[0] / t_swing_avg, dist[1] / t_swing_avg};
      for (int dim : {X, Y}) { g[row++] = curr[kPos][dim] - center[dim]; g[row++] = curr[kVel][dim] - des_vel[dim]; }
    }
  }
  void Bounds(Bound* b) const override { for (int i = 0; i < rows; ++i) b[i] = kBoundZero; }
  void FillBlock(const std::string& vs, Sp& jac) const override {  // :91-108
    if (vs != mo->name) return;
    int row = 0;
    for (int id : ids) for (int dim : {X, Y}) {
      jac.ref(row, mo->GetOptIndex(id, kPos, dim)) = 1.0;
      jac.ref(row, mo->GetOptIndex(id + 1, kPos, dim)) = -0.5;
      jac.ref(row, mo->GetOptIndex(id - 1, kPos, dim)) = -0.5;
      row++;
      jac.ref(row, mo->GetOptIndex(id, kVel, dim)) = 1.0;
      jac.ref(row, mo->GetOptIndex(id + 1, kPos, dim)) = -1.0 / t_swing_avg;
      jac.ref(row, mo->GetOptIndex(id - 1, kPos, dim)) = +1.0 / t_swing_avg;
      row++;
    }
  }
};

struct TotalDurationC : CSet {  // total_duration_constraint.cc
  Durations* pd; double T_total;
  TotalDurationC(const Ctx* ctx, double T, int ee) : pd(ctx->durations[ee]), T_total(T) { name = "totalduration-" + std::to_string(ee); rows = 1; }
  void Values(double* g) const override {  // :48-54
    double s = 0.0; for (int i = 0; i < pd->Rows(); ++i) s += pd->durations[i]; g[0] = s;
  }
  void Bounds(Bound* b) const override { b[0] = Bound{0.1, T_total - 0.2}; }  // :56-64
  void FillBlock(const std::string& vs, Sp& jac) const override {  // :66-72
    if (vs == pd->name) for (int col = 0; col < pd->Rows(); ++col) jac.ref(0, col) = 1.0;
  }
};

struct SplineAccC : CSet {  // spline_acc_constraint.cc
  Spline* s; std::string var; int n_junctions; std::vector<double> T;
  SplineAccC(Spline* sp, const std::string& v) : s(sp), var(v) {
    name = "splineacc-" + v; n_junctions = (int)s->T.size() - 1; T = s->T; rows = 3 * n_junctions;
  }
  void Values(double* g) const override {  // :49-65
    for (int j = 0; j < n_junctions; ++j) {
      V3 ap = s->GetPoint(j, T[j]).a, an = s->GetPoint(j + 1, 0.0).a;
      for (int d = 0; d < 3; ++d) g[j * 3 + d] = ap[d] - an[d];
    }
  }
  void Bounds(Bound* b) const override { for (int i = 0; i < rows; ++i) b[i] = kBoundZero; }
  void FillBlock(const std::string& vs, Sp& jac) const override {  // :67-80
    if (vs != var) return;
    for (int j = 0; j < n_junctions; ++j)
      sp_set_rows(jac, j * 3, sp_sub(s->JacWrtNodes(j, T[j], kAcc), s->JacWrtNodes(j + 1, 0.0, kAcc)));
  }
};

struct NodeCostT {  // node_cost.cc
  NodeVars* nodes; int deriv, dim; double weight;
  double Cost() const {  // :53-63 — reference leaves `cost` uninitialised (UB); defined as 0.0 here
    double cost = 0.0;
    for (auto& n : nodes->nodes) { double val = n[deriv][dim]; cost += weight * std::pow(val, 2); }
    return cost;
  }
  void FillBlock(const std::string& vs, Sp& jac) const {  // :65-76
    if (vs != nodes->name) return;
    for (int i = 0; i < nodes->Rows(); ++i) for (auto& nvi : nodes->infos[i])
      if (nvi.deriv == deriv && nvi.dim == dim) { double val = nodes->nodes[nvi.id][deriv][dim]; jac.ref(0, i) += weight * 2.0 * val; }
  }
};

// ---------------------------------------------------------------------------
// Problem = NlpFormulation (nlp_formulation.cc) + ifopt::Problem assembly
// ---------------------------------------------------------------------------
struct Problem {
  twb_spec spec;
  Robot robot; Terrain terrain; Srbd model; Ctx ctx;
  std::vector<std::unique_ptr<VarSet>> store;  // owns every variable set incl. unused durations
  std::vector<VarSet*> vars;                   // variable sets in the NLP, in order
  std::vector<std::unique_ptr<Spline>> splines;
  std::vector<std::unique_ptr<CSet>> csets;
  std::vector<NodeCostT> costs;
  int n = 0, m = 0;
  std::vector<int> row_ptr, col_idx;           // structure from the first Jacobian

  static std::vector<double> BasePolyDurations(double T, double dt) {  // parameters.cc:82-98
    std::vector<double> v; twb_f64 t_left = (twb_f64)T; const twb_f64 eps = 1e-10, dt64 = (twb_f64)dt;
    while (t_left > eps) { twb_f64 d = t_left > dt64 ? dt64 : t_left; v.push_back(d); t_left -= dt64; }
    return v;
  }
  double TotalTime() const {  // parameters.cc:112-126
    twb_f64 T = 0.0;
    if (spec.n_ee > 0) { T = 0.0; for (int i = 0; i < spec.n_phases[0]; ++i) T += spec.phase_durations[0][i]; }
    return T;
  }
  bool OptimizeTimings() const {  // parameters.cc:128-135
    for (int i = 0; i < spec.n_constraints; ++i) if (spec.constraints[i] == TWB_C_TOTAL_TIME) return true;
    return false;
  }

  explicit Problem(const twb_spec& s) : spec(s) {
    robot = MakeRobot(spec.robot);
    terrain.id = spec.terrain;
    model.Init(robot);
    const int n_ee = spec.n_ee;
    const double T = TotalTime();
    auto v3 = [](const auto* p) { return V3{p[0], p[1], p[2]}; };
    auto dims = [](const int* f) { std::vector<int> d; for (int i = 0; i < 3; ++i) if (f[i]) d.push_back(i); return d; };
    std::vector<double> base_T = BasePolyDurations(T, spec.duration_base_polynomial);

    // ---- MakeBaseVariables, nlp_formulation.cc:95-125
    int n_nodes = (int)base_T.size() + 1;
    auto lin = NodeVars::MakeAll(n_nodes, "base-lin");
    {
      double x = spec.final_base_lin_pos[0], y = spec.final_base_lin_pos[1];
      double z = terrain.Height(x, y) - robot.nominal[0][Z];
      lin->SetByLinearInterpolation(v3(spec.initial_base_lin_pos), V3{x, y, z}, T);
      lin->AddStartBound(kPos, {X, Y, Z}, v3(spec.initial_base_lin_pos));
      lin->AddStartBound(kVel, {X, Y, Z}, v3(spec.initial_base_lin_vel));
      lin->AddFinalBound(kPos, dims(spec.bounds_final_lin_pos), v3(spec.final_base_lin_pos));
      lin->AddFinalBound(kVel, dims(spec.bounds_final_lin_vel), v3(spec.final_base_lin_vel));
    }
    auto ang = NodeVars::MakeAll(n_nodes, "base-ang");
    {
      ang->SetByLinearInterpolation(v3(spec.initial_base_ang_pos), v3(spec.final_base_ang_pos), T);
      ang->AddStartBound(kPos, {X, Y, Z}, v3(spec.initial_base_ang_pos));
      ang->AddStartBound(kVel, {X, Y, Z}, v3(spec.initial_base_ang_vel));
      ang->AddFinalBound(kPos, dims(spec.bounds_final_ang_pos), v3(spec.final_base_ang_pos));
      ang->AddFinalBound(kVel, dims(spec.bounds_final_ang_vel), v3(spec.final_base_ang_vel));
    }
    NodeVars* lin_p = lin.get(); NodeVars* ang_p = ang.get();
    vars.push_back(lin_p); vars.push_back(ang_p);
    store.push_back(std::move(lin)); store.push_back(std::move(ang));

    // ---- MakeEndeffectorVariables, :127-156
    for (int ee = 0; ee < n_ee; ++ee) {
      auto nodes = NodeVars::MakeEEMotion(spec.n_phases[ee], spec.in_contact_at_start[ee] != 0,
                                          "ee-motion_" + std::to_string(ee), spec.ee_polynomials_per_swing_phase);
      double yaw = spec.final_base_ang_pos[Z];
      M3 w_R_b = sp_to_dense3(Euler::RotBaseToWorld(V3{0.0, 0.0, yaw}));
      V3 final_ee = vadd(v3(spec.final_base_lin_pos), m3_mulvec(w_R_b, robot.nominal[ee]));
      double x = final_ee[X], y = final_ee[Y], z = terrain.Height(x, y);
      nodes->SetByLinearInterpolation(v3(spec.initial_ee_W[ee]), V3{x, y, z}, T);
      nodes->AddStartBound(kPos, {X, Y, Z}, v3(spec.initial_ee_W[ee]));
      ctx.ee_motion_nodes.push_back(nodes.get()); vars.push_back(nodes.get()); store.push_back(std::move(nodes));
    }
    // ---- MakeForceVariables, :158-181
    for (int ee = 0; ee < n_ee; ++ee) {
      auto nodes = NodeVars::MakeEEForce(spec.n_phases[ee], spec.in_contact_at_start[ee] != 0,
                                         "ee-force_" + std::to_string(ee), spec.force_polynomials_per_stance_phase);
      V3 f_stance{0.0, 0.0, model.m * model.g / n_ee};
      nodes->SetByLinearInterpolation(f_stance, f_stance, T);
      ctx.ee_force_nodes.push_back(nodes.get()); vars.push_back(nodes.get()); store.push_back(std::move(nodes));
    }
    // ---- MakeContactScheduleVariables, :183-198 (+ :76-82)
    ctx.optimize_timings = OptimizeTimings();
    for (int ee = 0; ee < n_ee; ++ee) {
      std::vector<double> tim(spec.phase_durations[ee], spec.phase_durations[ee] + spec.n_phases[ee]);
      auto d = std::make_unique<Durations>(ee, tim, spec.in_contact_at_start[ee] != 0,
                                           spec.bound_phase_duration_min, spec.bound_phase_duration_max);
      ctx.durations.push_back(d.get());
      if (ctx.optimize_timings) vars.push_back(d.get());
      store.push_back(std::move(d));
    }
    // ---- SplineHolder, spline_holder.cc:35-61
    splines.push_back(std::make_unique<Spline>(lin_p, base_T)); ctx.base_lin = splines.back().get();
    splines.push_back(std::make_unique<Spline>(ang_p, base_T)); ctx.base_ang = splines.back().get();
    for (int ee = 0; ee < n_ee; ++ee) {
      NodeVars* mo = ctx.ee_motion_nodes[ee]; NodeVars* fo = ctx.ee_force_nodes[ee]; Durations* pd = ctx.durations[ee];
      auto mT = mo->ConvertPhaseToPolyDurations(pd->durations), fT = fo->ConvertPhaseToPolyDurations(pd->durations);
      splines.push_back(std::make_unique<Spline>(mo, mT, ctx.optimize_timings ? pd : nullptr)); ctx.ee_motion.push_back(splines.back().get());
      splines.push_back(std::make_unique<Spline>(fo, fT, ctx.optimize_timings ? pd : nullptr)); ctx.ee_force.push_back(splines.back().get());
    }
    ctx.model = &model; ctx.terrain = &terrain; ctx.robot = robot;

    // ---- GetConstraints, :200-331
    for (int i = 0; i < spec.n_constraints; ++i) {
      switch (spec.constraints[i]) {
        case TWB_C_DYNAMIC: csets.push_back(std::make_unique<DynamicC>(&ctx, T, spec.dt_constraint_dynamic)); break;
        case TWB_C_EE_ROM: for (int ee = 0; ee < n_ee; ++ee) csets.push_back(std::make_unique<RomC>(&ctx, T, spec.dt_constraint_range_of_motion, ee)); break;
        case TWB_C_BASE_ROM: csets.push_back(std::make_unique<BaseMotionC>(&ctx, T, spec.dt_constraint_base_motion)); break;
        case TWB_C_TOTAL_TIME: for (int ee = 0; ee < n_ee; ++ee) csets.push_back(std::make_unique<TotalDurationC>(&ctx, T, ee)); break;
        case TWB_C_TERRAIN: for (int ee = 0; ee < n_ee; ++ee) csets.push_back(std::make_unique<TerrainC>(&ctx, ee)); break;
        case TWB_C_FORCE: for (int ee = 0; ee < n_ee; ++ee) csets.push_back(std::make_unique<ForceC>(&ctx, ee, spec.force_limit_in_normal_direction)); break;
        case TWB_C_SWING: for (int ee = 0; ee < n_ee; ++ee) csets.push_back(std::make_unique<SwingC>(&ctx, ee)); break;
        case TWB_C_BASE_ACC:
          csets.push_back(std::make_unique<SplineAccC>(ctx.base_lin, "base-lin"));
          csets.push_back(std::make_unique<SplineAccC>(ctx.base_ang, "base-ang")); break;
        default: break;
      }
    }
    // ---- GetCosts, :333-376
    for (int i = 0; i < spec.n_costs; ++i) {
      double w = spec.cost_weights[i];
      if (spec.cost_ids[i] == TWB_COST_FORCES) for (int ee = 0; ee < n_ee; ++ee) costs.push_back({ctx.ee_force_nodes[ee], kPos, Z, w});
      if (spec.cost_ids[i] == TWB_COST_EE_MOTION) for (int ee = 0; ee < n_ee; ++ee) {
        costs.push_back({ctx.ee_motion_nodes[ee], kVel, X, w}); costs.push_back({ctx.ee_motion_nodes[ee], kVel, Y, w}); }
    }
    for (auto* v : vars) n += v->Rows();
    for (auto& c : csets) m += c->rows;
    // structure: IpoptAdapter takes it from the first GetJacobianOfConstraints() (at x0)
    Sp J = Jacobian();
    row_ptr.assign(m + 1, 0);
    for (int i = 0; i < m; ++i) { row_ptr[i + 1] = row_ptr[i] + (int)J.r[i].size(); for (auto& e : J.r[i]) col_idx.push_back(e.first); }
  }

  void X0(double* x) const { int o = 0; for (auto* v : vars) { v->GetValues(x + o); o += v->Rows(); } }
  void SetVariables(const double* x) { int o = 0; for (auto* v : vars) { v->SetVariables(x + o); o += v->Rows(); } }
  void XBounds(double* lo, double* up) const {
    std::vector<Bound> b(n); int o = 0; for (auto* v : vars) { v->GetBounds(b.data() + o); o += v->Rows(); }
    for (int i = 0; i < n; ++i) { lo[i] = b[i].lo; up[i] = b[i].up; }
  }
  void GBounds(double* lo, double* up) const {
    std::vector<Bound> b(m); int o = 0; for (auto& c : csets) { c->Bounds(b.data() + o); o += c->rows; }
    for (int i = 0; i < m; ++i) { lo[i] = b[i].lo; up[i] = b[i].up; }
  }
  void Values(double* g) const { int o = 0; for (auto& c : csets) { c->Values(g + o); o += c->rows; } }
  // ConstraintSet::GetJacobian + Composite::GetJacobian [ifopt]
  Sp Jacobian() const {
    Sp J(m, n);
    int row0 = 0;
    for (auto& c : csets) {
      int col0 = 0;
      // setFromTriplets: duplicates summed; here blocks of distinct variable sets never overlap
      for (auto* v : vars) {
        Sp blk(c->rows, v->Rows());
        c->FillBlock(v->name, blk);
        for (int i = 0; i < blk.nr; ++i) for (auto& e : blk.r[i]) J.r[row0 + i].emplace_back(col0 + e.first, e.second);
        col0 += v->Rows();
      }
      row0 += c->rows;
    }
    return J;
  }
  double Cost() const { double s = 0.0; for (auto& c : costs) s += c.Cost(); return s; }
  void Grad(double* grad) const {  // Composite(cost)::GetJacobian: all terms summed into row 0
    for (int i = 0; i < n; ++i) grad[i] = 0.0;
    for (auto& c : costs) {
      int col0 = 0;
      for (auto* v : vars) {
        Sp blk(1, v->Rows()); c.FillBlock(v->name, blk);
        for (auto& e : blk.r[0]) grad[col0 + e.first] += e.second;
        col0 += v->Rows();
      }
    }
  }
  // returns 0 ok, 1 if the pattern at x differs from the structure at x0
  int Eval(const double* x, double* g, double* vals, double* cost, double* grad) {
    SetVariables(x);
    if (g) Values(g);
    int rc = 0;
    if (vals) {
      Sp J = Jacobian();
      int k = 0;
      for (int i = 0; i < m; ++i) {
        if ((int)J.r[i].size() != row_ptr[i + 1] - row_ptr[i]) { rc = 1; }
        for (auto& e : J.r[i]) { if (k < (int)col_idx.size()) { if (col_idx[k] != e.first) rc = 1; vals[k] = e.second; } ++k; }
      }
      if (k != (int)col_idx.size()) rc = 1;
    }
    if (cost) *cost = costs.empty() ? 0.0 : Cost();
    if (grad) Grad(grad);
    return rc;
  }
};

}  // namespace

// ---------------------------------------------------------------------------
// C entry points for ctypes (tests / bench cpu_baseline only)
// ---------------------------------------------------------------------------
extern "C" {

void* oracle_create(const twb_spec* spec) { return new Problem(*spec); }
void oracle_destroy(void* h) { delete static_cast<Problem*>(h); }
void oracle_dims(void* h, int* n, int* m, int* nnz) {
  auto* p = static_cast<Problem*>(h); *n = p->n; *m = p->m; *nnz = (int)p->col_idx.size();
}
void oracle_structure(void* h, int* row_ptr, int* col_idx) {
  auto* p = static_cast<Problem*>(h);
  std::copy(p->row_ptr.begin(), p->row_ptr.end(), row_ptr);
  std::copy(p->col_idx.begin(), p->col_idx.end(), col_idx);
}
void oracle_bounds(void* h, double* xl, double* xu, double* gl, double* gu) {
  auto* p = static_cast<Problem*>(h); p->XBounds(xl, xu); p->GBounds(gl, gu);
}
void oracle_x0(void* h, double* x) { static_cast<Problem*>(h)->X0(x); }
int oracle_num_varsets(void* h) { return (int)static_cast<Problem*>(h)->vars.size(); }
int oracle_varset(void* h, int i, char* name, int cap) {
  auto* p = static_cast<Problem*>(h); std::snprintf(name, cap, "%s", p->vars[i]->name.c_str()); return p->vars[i]->Rows();
}
int oracle_num_csets(void* h) { return (int)static_cast<Problem*>(h)->csets.size(); }
int oracle_cset(void* h, int i, char* name, int cap) {
  auto* p = static_cast<Problem*>(h); std::snprintf(name, cap, "%s", p->csets[i]->name.c_str()); return p->csets[i]->rows;
}
void oracle_set_terrain(void* h, int terrain) { static_cast<Problem*>(h)->terrain.id = terrain; }
int oracle_eval(void* h, const double* x, double* g, double* vals, double* cost, double* grad) {
  return static_cast<Problem*>(h)->Eval(x, g, vals, cost, grad);
}
double oracle_terrain_height(int terrain, double x, double y) { Terrain t; t.id = terrain; return t.Height(x, y); }
// grid of the TWB_GRID_CSV terrain (process-global; set before evaluating)
void oracle_set_grid(const double* heights, int rows, int cols) {
  g_grid.h.assign(heights, heights + (size_t)rows * cols); g_grid.rows = rows; g_grid.cols = cols;
}
void oracle_set_grid_map(const float* heights, int size_x, int size_y, double resolution, double pos_x, double pos_y) {
  g_gridmap.h.assign(heights, heights + (size_t)size_x * size_y); g_gridmap.sx = size_x; g_gridmap.sy = size_y;
  g_gridmap.res = resolution; g_gridmap.px = pos_x; g_gridmap.py = pos_y;
}
// fpowr::NearestPlaneLookup::GetNearestPlaneIndex (fpowr/include/fpowr/nearest_plane_lookup.h:62-84): boost::geometry::distance
// (point, polygon), default cartesian strategies, restated (boost is un-vendored): covered_by -> 0 (winding strategy over
// the ring's segments as given; a `closed` ring has no implicit closing segment), else min over the ring's segments of the
// projected-point distance (comparable = squared, one sqrt at the end).  min_distance starts at DBL_MAX, strict `<`.
static double PolygonDistanceRef(const double* v, int n, double px, double py) {
  if (n <= 0) return std::numeric_limits<double>::max();
  if (n == 1) return std::sqrt((px - v[0]) * (px - v[0]) + (py - v[1]) * (py - v[1]));
  int winding = 0; bool touches = false; double best = std::numeric_limits<double>::max();
  for (int i = 0; i + 1 < n; ++i) {
    const double ax = v[2 * i], ay = v[2 * i + 1], bx = v[2 * i + 2], by = v[2 * i + 3];
    const double side = (bx - ax) * (py - ay) - (px - ax) * (by - ay);
    if (ay <= py) { if (by > py && side > 0) ++winding; } else if (by <= py && side < 0) --winding;
    const double vx = bx - ax, vy = by - ay, wx = px - ax, wy = py - ay;
    const double c1 = wx * vx + wy * vy;
    double d2;
    if (c1 <= 0) d2 = wx * wx + wy * wy;
    else {
      const double c2 = vx * vx + vy * vy;
      if (c2 <= c1) d2 = (px - bx) * (px - bx) + (py - by) * (py - by);
      else { const double b = c1 / c2, qx = ax + b * vx, qy = ay + b * vy; d2 = (px - qx) * (px - qx) + (py - qy) * (py - qy); }
    }
    if (d2 == 0.0) touches = true;
    if (d2 < best) best = d2;
  }
  if (touches || winding != 0) return 0.0;
  return std::sqrt(best);
}
int oracle_nearest_plane(const int* poly_offsets, int n_polys, const double* verts, double x, double y) {
  double min_distance = std::numeric_limits<double>::max(); int nearest = -1;
  for (int i = 0; i < n_polys; ++i) {
    const double d = PolygonDistanceRef(verts + 2 * (size_t)poly_offsets[i], poly_offsets[i + 1] - poly_offsets[i], x, y);
    if (d < min_distance) { min_distance = d; nearest = i; }
  }
  return nearest;
}
// towr::LinearEqualityConstraint::GetValues (linear_constraint.cc:46-51): M * x (Eigen dense row-times-vector: ascending column sum)
void oracle_linear_equality(const double* M, int rows, int cols, const double* x_set, double* g) {
  for (int r = 0; r < rows; ++r) { double acc = 0.0; for (int c = 0; c < cols; ++c) acc += M[(size_t)r * cols + c] * x_set[c]; g[r] = acc; }
}
// towr::SoftConstraint::GetValues / GetJacobian (soft_constraint.cc:53-72) of rows row0 .. row0+n_rows-1 of an evaluated problem:
// cost = 0.5 (g-b)^T W (g-b), b = (upper+lower)/2 (:41-46); grad = J^T W (g-b)
void oracle_soft_constraint(void* h, const double* g, const double* vals, int row0, int n_rows, const double* w, double* cost, double* grad) {
  auto* p = static_cast<Problem*>(h);
  std::vector<double> g_lo(p->m), g_up(p->m);
  p->GBounds(g_lo.data(), g_up.data());
  double c = 0.0;
  for (int i = 0; i < p->n; ++i) grad[i] = 0.0;
  for (int r = 0; r < n_rows; ++r) {
    const double b = (g_up[row0 + r] + g_lo[row0 + r]) / 2.;
    const double d = g[row0 + r] - b, wr = w ? w[r] : 1.0;
    c += d * wr * d;
    for (int k = p->row_ptr[row0 + r]; k < p->row_ptr[row0 + r + 1]; ++k) grad[p->col_idx[k]] += vals[k] * (wr * d);
  }
  *cost = 0.5 * c;
}
void oracle_terrain_point(int terrain, double x, double y, double* out3) {
  Terrain t; t.id = terrain; out3[0] = t.Height(x, y); out3[1] = t.DerivOfHeightWrt(0, x, y); out3[2] = t.DerivOfHeightWrt(1, x, y);
}

// Batched evaluation over independent instances (CPU baseline): one Problem
// clone per thread (the reference's objects are mutable and not re-entrant).
// terrain_ids may be NULL.  Returns the OR of the per-instance return codes.
int oracle_batch_eval(const twb_spec* spec, int B, const int* terrain_ids, const double* x,
                      double* g, double* vals, double* cost, double* grad, int n_threads) {
  int rc = 0;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel reduction(| : rc)
  {
    Problem p(*spec);
    const int n = p.n, m = p.m; const size_t nnz = p.col_idx.size();
#pragma omp for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
      if (terrain_ids) p.terrain.id = terrain_ids[b];
      rc |= p.Eval(x + (size_t)b * n, g ? g + (size_t)b * m : nullptr, vals ? vals + (size_t)b * nnz : nullptr,
                   cost ? cost + b : nullptr, grad ? grad + (size_t)b * n : nullptr);
    }
  }
  return rc;
}
// fpowr::GetTrajectory (fpowr/include/fpowr/footstep_plan_extractor.h:19-53): the solution sampled every dt.
// Per sample: base lin p, v, a (9) | base orientation quaternion w, x, y, z (4) | angular velocity (3) | angular
// acceleration (3) | per foot: contact flag (1), ee-motion p, v, a (9), ee-force (3).
// The quaternion is Eigen::Quaterniond(Matrix3d) (Eigen 3.3 Quaternion.h, quaternionbase_assign_impl<Other,3,3>) restated.
static void QuaternionFromMatrix(const M3& m, double q[4]) {   // q = w, x, y, z
  double t = m.a[0][0] + m.a[1][1] + m.a[2][2];
  if (t > 0.0) {
    t = std::sqrt(t + 1.0); q[0] = 0.5 * t; t = 0.5 / t;
    q[1] = (m.a[2][1] - m.a[1][2]) * t; q[2] = (m.a[0][2] - m.a[2][0]) * t; q[3] = (m.a[1][0] - m.a[0][1]) * t;
  } else {
    int i = 0; if (m.a[1][1] > m.a[0][0]) i = 1; if (m.a[2][2] > m.a[i][i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m.a[i][i] - m.a[j][j] - m.a[k][k] + 1.0);
    q[1 + i] = 0.5 * t; t = 0.5 / t;
    q[0] = (m.a[k][j] - m.a[j][k]) * t; q[1 + j] = (m.a[j][i] + m.a[i][j]) * t; q[1 + k] = (m.a[k][i] + m.a[i][k]) * t;
  }
}
int oracle_trajectory_dims(void* h, double dt, int* n_samples, int* n_values) {
  Problem* p = static_cast<Problem*>(h);
  double T = 0.0; for (double d : p->ctx.base_lin->T) T += d;   // Spline::GetTotalTime
  int n = 0; for (double t = 0.0; t <= T + 1e-5; t += dt) ++n;
  *n_samples = n; *n_values = 19 + 13 * p->robot.n_ee;
  return 0;
}
int oracle_trajectory(void* h, const double* x, double dt, double* out) {
  Problem* p = static_cast<Problem*>(h);
  p->SetVariables(x);
  const int n_ee = p->robot.n_ee, K = 19 + 13 * n_ee;
  double T = 0.0; for (double d : p->ctx.base_lin->T) T += d;
  Euler eu; eu.s = p->ctx.base_ang;
  int k = 0;
  for (double t = 0.0; t <= T + 1e-5; t += dt, ++k) {
    double* o = out + (size_t)k * K;
    State3 lin = p->ctx.base_lin->GetPoint(t);
    for (int d = 0; d < 3; ++d) { o[d] = lin.p[d]; o[3 + d] = lin.v[d]; o[6 + d] = lin.a[d]; }
    QuaternionFromMatrix(sp_to_dense3(eu.RotBaseToWorld(t)), o + 9);
    V3 w = eu.AngVel(t), wd = eu.AngAcc(t);
    for (int d = 0; d < 3; ++d) { o[13 + d] = w[d]; o[16 + d] = wd[d]; }
    for (int ee = 0; ee < n_ee; ++ee) {
      double* e = o + 19 + 13 * ee;
      const Durations* pd = p->ctx.durations[ee];
      int phase = GetSegmentID(t, pd->durations);                       // PhaseDurations::IsContactPhase, phase_durations.cc:120-124
      e[0] = (phase % 2 == 0 ? pd->initial_contact : !pd->initial_contact) ? 1.0 : 0.0;
      State3 mo = p->ctx.ee_motion[ee]->GetPoint(t);
      for (int d = 0; d < 3; ++d) { e[1 + d] = mo.p[d]; e[4 + d] = mo.v[d]; e[7 + d] = mo.a[d]; }
      V3 f = p->ctx.ee_force[ee]->GetPoint(t).p;
      for (int d = 0; d < 3; ++d) e[10 + d] = f[d];
    }
  }
  return 0;
}
// fpowr::ExtractInitialGuess (fpowr/include/fpowr/initial_guess_extractor.h:17-34) at each of `n_times` sample times
// (ExtractInitialGuesses, :36-48): out[k] = time | state(12) | controls(36).
int oracle_initial_guess(void* h, const double* x, const double* times, int n_times, double* out) {
  Problem* p = static_cast<Problem*>(h);
  if (p->robot.n_ee > 4) return 1;
  p->SetVariables(x);
  for (int k = 0; k < n_times; ++k) {
    const double t = times[k];
    double* o = out + (size_t)k * 49;
    for (int i = 0; i < 49; ++i) o[i] = 0.0;
    o[0] = t;
    double* state = o + 1; double* controls = o + 13;
    State3 lin = p->ctx.base_lin->GetPoint(t), ang = p->ctx.base_ang->GetPoint(t);
    for (int d = 0; d < 3; ++d) { state[d] = lin.p[d]; state[3 + d] = ang.p[d]; state[6 + d] = lin.v[d]; state[9 + d] = ang.v[d]; }
    for (int i = 0; i < p->robot.n_ee; ++i) {
      State3 mo = p->ctx.ee_motion[i]->GetPoint(t);
      V3 f = p->ctx.ee_force[i]->GetPoint(t).p;
      for (int d = 0; d < 3; ++d) { controls[i * 3 + d] = mo.a[d]; controls[12 + i * 3 + d] = 0.0; controls[24 + i * 3 + d] = f[d]; }
    }
  }
  return 0;
}
// fpowr::ExtractFootstepPlan (fpowr/include/fpowr/footstep_plan_extractor.h:68-133) without the nearest-plane lookup
// (boost::geometry + ROS message, outside the tree): GetTrajectory(solution, 0.01), the first state and every state whose
// contact set differs from the previous one (HasEndEffectorContactChanged, :55-66) are footstep states; duration = time
// to the next footstep state, the last one up to time_horizon.  out[i] = t_global | duration | per foot: contact, ee p.
int oracle_footstep_plan(void* h, const double* x, double time_horizon, int max_states, int* n_states, double* out) {
  Problem* p = static_cast<Problem*>(h);
  const int n_ee = p->robot.n_ee, K = 19 + 13 * n_ee, V = 2 + 4 * n_ee;
  int n_samples = 0, n_values = 0;
  oracle_trajectory_dims(h, 0.01, &n_samples, &n_values);
  std::vector<double> traj((size_t)n_samples * K);
  oracle_trajectory(h, x, 0.01, traj.data());
  std::vector<double> t_global; std::vector<int> index;
  double t = 0.0;
  for (int k = 0; k < n_samples; ++k, t += 0.01) {
    bool changed = (k == 0);
    for (int e = 0; e < n_ee && !changed; ++e) changed = traj[(size_t)k * K + 19 + 13 * e] != traj[(size_t)(k - 1) * K + 19 + 13 * e];
    if (changed) { t_global.push_back(t); index.push_back(k); }
  }
  *n_states = (int)index.size();
  for (int i = 0; i < (int)index.size() && i < max_states; ++i) {
    double* o = out + (size_t)i * V; const double* st = traj.data() + (size_t)index[i] * K;
    o[0] = t_global[i];
    o[1] = (i + 1 < (int)index.size()) ? (t_global[i + 1] - t_global[i]) : (time_horizon - t_global[i]);
    for (int e = 0; e < n_ee; ++e) { o[2 + 4 * e] = st[19 + 13 * e]; for (int d = 0; d < 3; ++d) o[3 + 4 * e + d] = st[20 + 13 * e + d]; }
  }
  return 0;
}
int oracle_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
