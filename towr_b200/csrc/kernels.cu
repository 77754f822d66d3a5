// kernels.cu — sm_100a kernels of the batched NLP evaluation (fp64).
//
// Everything is a data-parallel map with LANE = PROBLEM INSTANCE.  The iterates are first brought into
// an instance-tiled matrix XT[tile][n+1][32] (tile = 32 consecutive instances, lane = instance), so that
// every read of a node value by a warp is one 256-byte row segment:
//
//   TransposeIn   x[B][n]   -> XT[tile][n+1][32]   (row n stays 0: "not optimised" node values)
//   DynOut        XT        -> g rows + CSR values of the dynamic constraint          (warp = sample x tile)
//   RomOut        XT        -> g rows + CSR values of the range-of-motion constraints (warp = sample x tile, all feet)
//   NodeOut       XT        -> g rows + CSR values of the node-wise sets: terrain, force, swing, spline-acc,
//                              base-motion (warp = group of consecutive nodes x tile)
//   CostKernel    XT        -> cost + gradient (only when the formulation has cost terms)
//
// In the *Out kernels a warp owns one unit of 32 instances: each lane evaluates the splines its unit
// needs (reference operation order), computes its instance's unit state into a padded shared-memory
// block (row = state slot, column = lane), the warp synchronises, and then the lanes switch roles —
// lane = 16-byte pair of output elements — and stream
//     out[instance][off + h] = state[d_h][instance] * c_h
// for the 32 instances, so that every store instruction covers 512 contiguous bytes of one instance's
// CSR value array (a unit's rows are consecutive CSR rows).  The state never leaves the SM; HBM sees x
// once and g / jac exactly once.  The Out kernels are independent and run on separate streams.
//
// Reference math restated per device function (file:line cited there).  This translation unit is
// compiled with -fmad=false: plain * and + round like the reference's scalar C++; fused
// multiply-adds are written explicitly (fma) where the algebra is re-associated anyway.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>

#include "device_tables.h"
#include "launch.h"

namespace twb {
void (*g_after_launch)(const char* label, cudaStream_t stream) = nullptr;   // profiling hook (capi.cc, TWB_PROFILE=1)
#define TWB_MARK(label, stream) do { if (g_after_launch) g_after_launch(label, stream); } while (0)
namespace {

#ifndef TWB_DYN_WARPS
#define TWB_DYN_WARPS 4
#endif
#ifndef TWB_DYN_CTAS
#define TWB_DYN_CTAS 2
#endif
#ifndef TWB_ROM_WARPS
#define TWB_ROM_WARPS 4
#endif
#ifndef TWB_ROM_CTAS
#define TWB_ROM_CTAS 3
#endif
#ifndef TWB_NODE_WARPS
#define TWB_NODE_WARPS 4
#endif
constexpr int kDynWarps = TWB_DYN_WARPS;     // consecutive samples per CTA of DynOut (one instance tile)
constexpr int kRomWarps = TWB_ROM_WARPS;     // consecutive samples per CTA of RomOut
constexpr int kNodeWarps = TWB_NODE_WARPS;   // consecutive node groups per CTA of NodeOut
constexpr int kLD = 34;                      // leading dimension of a state block: 32 instances, padded; even keeps rows 16-byte aligned

// output stores: streaming (evict-first) — the values are consumed by the host / a solver, not by these kernels
__device__ __forceinline__ void StoreOut(double* p, double v) { __stcs(p, v); }
__device__ __forceinline__ void StoreOut2(double* p, double a, double b) {
  asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}

// column `b` of a row-major [rows][ld] matrix: element r lives at p[r * ld]
struct Col {
  double* p;
  size_t ld;
  __device__ __forceinline__ double& operator[](int r) const { return p[(size_t)r * ld]; }
};
struct ConstCol {
  const double* __restrict__ p;
  __device__ __forceinline__ double operator[](int r) const { return __ldg(p + (size_t)r * 32); }
};

// ---- cubic Hermite evaluation ------------------------------------------------
// a / b, correctly rounded, from y = RN(1/b) (Markstein): two residual corrections with exact
// FMA residuals.  Replaces the ~35-instruction IEEE division sequence; the divisors (T^2, T^3)
// are structure-class constants whose reciprocals come with the sample table.
__device__ __forceinline__ double DivExact(double a, double b, double y) {
  double q = a * y;
  double r = fma(-b, q, a);
  q = fma(r, y, q);
  r = fma(-b, q, a);
  return fma(r, y, q);
}
// CubicHermitePolynomial::UpdateCoeff (polynomial.cc:97-104) followed by
// Polynomial::GetPoint (polynomial.cc:47-61): sum_c d^k/dt^k(t^c) * coeff_c, c = A..D, in the
// reference's operation order (this translation unit is compiled with -fmad=false, so the
// products and sums below round exactly like the reference's scalar code).
struct SampleRegs { double T, T2, T3, rT2, rT3, t, t2, t3; int xi[12]; };
__device__ __forceinline__ SampleRegs LoadSample(const SplineSample* __restrict__ p) {
  const double2* d = reinterpret_cast<const double2*>(p);
  const double2 a = __ldg(d), b = __ldg(d + 1), c = __ldg(d + 2), e = __ldg(d + 3);
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p) + 4);
  const uint2 w = __ldg(reinterpret_cast<const uint2*>(p) + 10);
  SampleRegs r;
  r.T = a.x; r.T2 = a.y; r.T3 = b.x; r.rT2 = b.y; r.rT3 = c.x; r.t = c.y; r.t2 = e.x; r.t3 = e.y;
  const uint32_t v[6] = {u.x, u.y, u.z, u.w, w.x, w.y};
#pragma unroll
  for (int i = 0; i < 6; ++i) { r.xi[2 * i] = (int)(v[i] & 0xFFFFu); r.xi[2 * i + 1] = (int)(v[i] >> 16); }
  return r;
}
// kWant: 0 position; 1 position + acceleration; 2 position + velocity + acceleration
template <int kWant>
__device__ __forceinline__ void EvalSpline(const SplineSample* __restrict__ sp, const ConstCol xs, double pos[3], double vel[3], double acc[3]) {
  const SampleRegs s = LoadSample(sp);
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double p0 = xs[s.xi[d]], v0 = xs[s.xi[3 + d]], p1 = xs[s.xi[6 + d]], v1 = xs[s.xi[9 + d]];
    const double C = DivExact(-(3 * (p0 - p1) + s.T * (2 * v0 + v1)), s.T2, s.rT2);
    const double D = DivExact(2 * (p0 - p1) + s.T * (v0 + v1), s.T3, s.rT3);
    pos[d] = ((p0 + s.t * v0) + s.t2 * C) + s.t3 * D;
    if (kWant == 2) vel[d] = (v0 + (2 * s.t) * C) + (3 * s.t2) * D;
    if (kWant >= 1) acc[d] = 2 * C + (6 * s.t) * D;
  }
}

// ---- Euler angles (roll x, pitch y, yaw z; applied Z-Y'-X'') -------------------
struct Trig { double sx, cx, sy, cy, sz, cz; };
__device__ __forceinline__ Trig MakeTrig(const double th[3]) {
  Trig t; sincos(th[0], &t.sx, &t.cx); sincos(th[1], &t.sy, &t.cy); sincos(th[2], &t.sz, &t.cz); return t;
}
// EulerConverter::GetRotationMatrixBaseToWorld, euler_converter.cc:207-221
__device__ __forceinline__ void RotationMatrix(const Trig& t, double R[3][3]) {
  R[0][0] = t.cy * t.cz; R[0][1] = t.cz * t.sx * t.sy - t.cx * t.sz; R[0][2] = t.sx * t.sz + t.cx * t.cz * t.sy;
  R[1][0] = t.cy * t.sz; R[1][1] = t.cx * t.cz + t.sx * t.sy * t.sz; R[1][2] = t.cx * t.sy * t.sz - t.cz * t.sx;
  R[2][0] = -t.sy;       R[2][1] = t.cy * t.sx;                      R[2][2] = t.cx * t.cy;
}
// d(R_ij)/d(theta_d): coefficients of jac_x / jac_y / jac_z in
// EulerConverter::GetDerivativeOfRotationMatrixWrtNodes, euler_converter.cc:241-268
__device__ __forceinline__ void RotationDerivative(const Trig& t, double dR[3][3][3]) {
  const double sx = t.sx, cx = t.cx, sy = t.sy, cy = t.cy, sz = t.sz, cz = t.cz;
  dR[0][0][0] = 0.0;                         dR[0][0][1] = -cz * sy;      dR[0][0][2] = -cy * sz;
  dR[0][1][0] = sx * sz + cx * cz * sy;      dR[0][1][1] = cy * cz * sx;  dR[0][1][2] = -cx * cz - sx * sy * sz;
  dR[0][2][0] = cx * sz - cz * sx * sy;      dR[0][2][1] = cx * cy * cz;  dR[0][2][2] = cz * sx - cx * sy * sz;
  dR[1][0][0] = 0.0;                         dR[1][0][1] = -sy * sz;      dR[1][0][2] = cy * cz;
  dR[1][1][0] = cx * sy * sz - cz * sx;      dR[1][1][1] = cy * sx * sz;  dR[1][1][2] = -cx * sz + cz * sx * sy;
  dR[1][2][0] = -cx * cz - sx * sy * sz;     dR[1][2][1] = cx * cy * sz;  dR[1][2][2] = sx * sz + cx * cz * sy;
  dR[2][0][0] = 0.0;                         dR[2][0][1] = -cy;           dR[2][0][2] = 0.0;
  dR[2][1][0] = cx * cy;                     dR[2][1][1] = -sx * sy;      dR[2][1][2] = 0.0;
  dR[2][2][0] = -cy * sx;                    dR[2][2][1] = -cx * sy;      dR[2][2][2] = 0.0;
}
// EulerConverter::DerivOfRotVecMult (euler_converter.cc:223-239) as a dense 3x3:
// D[row][d] = sum_col v[col] * d(R or R^T)[row][col] / d(theta_d)
template <bool kInverse>
__device__ __forceinline__ void RotVecDerivative(const double dR[3][3][3], const double v[3], double D[3][3]) {
#pragma unroll
  for (int row = 0; row < 3; ++row)
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      double s = 0.0;
#pragma unroll
      for (int col = 0; col < 3; ++col) s = fma(v[col], kInverse ? dR[col][row][d] : dR[row][col][d], s);
      D[row][d] = s;
    }
}
__device__ __forceinline__ void Mul33(const double A[3][3], const double B[3][3], double C[3][3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C[i][j] = fma(A[i][2], B[2][j], fma(A[i][1], B[1][j], A[i][0] * B[0][j]));
}
__device__ __forceinline__ void MulVec(const double A[3][3], const double v[3], double o[3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i) o[i] = A[i][0] * v[0] + A[i][1] * v[1] + A[i][2] * v[2];
}
// C = [w]x * A  (Cross(), single_rigid_body_dynamics.cc:46-57)
__device__ __forceinline__ void CrossMul(const double w[3], const double A[3][3], double C[3][3]) {
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    C[0][j] = fma(w[1], A[2][j], -w[2] * A[1][j]);
    C[1][j] = fma(w[2], A[0][j], -w[0] * A[2][j]);
    C[2][j] = fma(w[0], A[1][j], -w[1] * A[0][j]);
  }
}

// ---- DynamicConstraint sample ------------------------------------------------
// Values: DynamicConstraint::UpdateModel (dynamic_constraint.cc:119-137) +
//   SingleRigidBodyDynamics::GetDynamicViolation (single_rigid_body_dynamics.cc:76-101).
// Jacobian state: the 3x3 maps A, B, C with
//   d(angular rows)/d(base-ang nodes) = A*dtheta + B*dtheta_dot + C*dtheta_ddot
//   restating SingleRigidBodyDynamics::GetJacobianWrtBaseAng (:123-165) with
//   EulerConverter::GetDerivOfAng{Vel,Acc}WrtEulerNodes (euler_converter.cc:85-131),
//   GetDerivMwrtNodes (:168-198), GetDerivMdotwrtNodes (:270-304);
//   sum of forces, per-foot force and lever arm for the other blocks (:103-121, :167-192).
// Sk: local state rows 1.. (Sk[0..2] sum f, Sk[3..38] base-ang block, Sk[39 + 6e ..] f_e, c - p_e);
// gk: the 6 constraint values
template <int kNEE>
__device__ __forceinline__ void DynamicUnit(const Plan& P, int k, const SplineSample* __restrict__ sp, const ConstCol xs,
                                            const Col Sk, const Col gk) {
  double c[3], cdd[3], th[3], thd[3], thdd[3], unused[3];
  EvalSpline<1>(sp + 0, xs, c, unused, cdd);
  EvalSpline<2>(sp + 1, xs, th, thd, thdd);

  // feet
  double fsum[3] = {0, 0, 0}, tau[3] = {0, 0, 0};
#pragma unroll
  for (int e = 0; e < kNEE; ++e) {
    double pe[3], f[3];
    EvalSpline<0>(sp + 2 + e, xs, pe, unused, unused);
    EvalSpline<0>(sp + 2 + kNEE + e, xs, f, unused, unused);
    const double r[3] = {c[0] - pe[0], c[1] - pe[1], c[2] - pe[2]};
    tau[0] += f[1] * r[2] - f[2] * r[1];
    tau[1] += f[2] * r[0] - f[0] * r[2];
    tau[2] += f[0] * r[1] - f[1] * r[0];
#pragma unroll
    for (int d = 0; d < 3; ++d) { fsum[d] += f[d]; Sk[39 + e * 6 + d] = f[d]; Sk[39 + e * 6 + 3 + d] = r[d]; }
  }
  Sk[0] = fsum[0]; Sk[1] = fsum[1]; Sk[2] = fsum[2];

  const Trig tr = MakeTrig(th);
  const double sy = tr.sy, cy = tr.cy, sz = tr.sz, cz = tr.cz;
  double R[3][3]; RotationMatrix(tr, R);
  const double yd = thd[1], zd = thd[2];
  // EulerConverter::GetM (:133-148) and GetMdot (:150-166)
  double M[3][3] = {{cy * cz, -sz, 0.0}, {cy * sz, cz, 0.0}, {-sy, 0.0, 1.0}};
  double Md[3][3] = {{-cz * sy * yd - cy * sz * zd, -cz * zd, 0.0}, {cy * cz * zd - sy * sz * yd, -sz * zd, 0.0}, {-cy * yd, 0.0, 0.0}};
  double om[3], omd[3];
  om[0] = M[0][0] * thd[0] + M[0][1] * thd[1];
  om[1] = M[1][0] * thd[0] + M[1][1] * thd[1];
  om[2] = M[2][0] * thd[0] + thd[2];
  omd[0] = (Md[0][0] * thd[0] + Md[0][1] * thd[1]) + (M[0][0] * thdd[0] + M[0][1] * thdd[1]);
  omd[1] = (Md[1][0] * thd[0] + Md[1][1] * thd[1]) + (M[1][0] * thdd[0] + M[1][1] * thdd[1]);
  omd[2] = (Md[2][0] * thd[0]) + (M[2][0] * thdd[0] + thdd[2]);

  // I_w = R I_b R^T
  double Ib[3][3], RIb[3][3], Rt[3][3], Iw[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) { Ib[i][j] = P.I_b[i * 3 + j]; Rt[i][j] = R[j][i]; }
  Mul33(R, Ib, RIb); Mul33(RIb, Rt, Iw);

  double Iw_om[3], Iw_omd[3];
  MulVec(Iw, om, Iw_om); MulVec(Iw, omd, Iw_omd);
  {
    const double wx[3] = {om[1] * Iw_om[2] - om[2] * Iw_om[1], om[2] * Iw_om[0] - om[0] * Iw_om[2], om[0] * Iw_om[1] - om[1] * Iw_om[0]};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      gk[d] = Iw_omd[d] + wx[d] - tau[d];
      const double grav = (d == 2) ? -P.mass * P.gravity : 0.0;
      gk[3 + d] = P.mass * cdd[d] - fsum[d] - grav;
    }
  }

  // ---- base-angular Jacobian maps
  double dR[3][3][3]; RotationDerivative(tr, dR);
  // d(omega)/d(theta) (GetDerivMwrtNodes contracted with theta_dot) ; d(omega)/d(theta_dot) = M
  double Jw[3][3] = {{0.0, thd[0] * (-cz * sy), thd[0] * (-cy * sz) + thd[1] * (-cz)},
                     {0.0, thd[0] * (-sy * sz), thd[0] * (cy * cz) + thd[1] * (-sz)},
                     {0.0, thd[0] * (-cy), 0.0}};
  // d(omega_dot)/d(theta): theta_dot * dMdot/dtheta + theta_ddot * dM/dtheta
  double Jwd_th[3][3] = {
      {0.0, thd[0] * (-cy * cz * yd + sy * sz * zd) + thdd[0] * (-cz * sy),
            (thd[0] * (sy * sz * yd - cy * cz * zd) + thd[1] * (sz * zd)) + (thdd[0] * (-cy * sz) + thdd[1] * (-cz))},
      {0.0, thd[0] * (-cy * sz * yd - cz * sy * zd) + thdd[0] * (-sy * sz),
            (thd[0] * (-cz * sy * yd - cy * sz * zd) + thd[1] * (-cz * zd)) + (thdd[0] * (cy * cz) + thdd[1] * (-sz))},
      {0.0, thd[0] * (sy * yd) + thdd[0] * (-cy), 0.0}};
  // d(omega_dot)/d(theta_dot): theta_dot * dMdot/dtheta_dot + Mdot
  double Jwd_thd[3][3] = {{Md[0][0], thd[0] * (-cz * sy) + Md[0][1], thd[0] * (-cy * sz) + thd[1] * (-cz)},
                          {Md[1][0], thd[0] * (-sy * sz) + Md[1][1], thd[0] * (cy * cz) + thd[1] * (-sz)},
                          {Md[2][0], thd[0] * (-cy), 0.0}};

  double v11[3], v21[3], tmp[3];
  MulVec(Rt, omd, tmp); MulVec(Ib, tmp, v11);   // I_b R^T omega_dot
  MulVec(Rt, om, tmp);  MulVec(Ib, tmp, v21);   // I_b R^T omega
  double D11[3][3], D12[3][3], D21[3][3], D22[3][3], T1[3][3], T2[3][3];
  RotVecDerivative<false>(dR, v11, D11); RotVecDerivative<true>(dR, omd, D12);
  RotVecDerivative<false>(dR, v21, D21); RotVecDerivative<true>(dR, om, D22);

  double A[3][3], Bm[3][3], C[3][3];
  // jac1 = D11 + R I_b D12 + I_w d(omega_dot)
  Mul33(RIb, D12, T1); Mul33(Iw, Jwd_th, T2);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int d = 0; d < 3; ++d) A[i][d] = (D11[i][d] + T1[i][d]) + T2[i][d];
  Mul33(Iw, Jwd_thd, Bm);
  Mul33(Iw, M, C);
  // jac2 = [omega]x (D21 + R I_b D22 + I_w d(omega)) - [I_w omega]x d(omega)
  double Gth[3][3], X1[3][3], X2[3][3];
  Mul33(RIb, D22, T1); Mul33(Iw, Jw, T2);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int d = 0; d < 3; ++d) Gth[i][d] = (D21[i][d] + T1[i][d]) + T2[i][d];
  CrossMul(om, Gth, X1); CrossMul(Iw_om, Jw, X2);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int d = 0; d < 3; ++d) A[i][d] += X1[i][d] - X2[i][d];
  CrossMul(om, C, X1); CrossMul(Iw_om, M, X2);   // I_w d(omega)/d(theta_dot) = I_w M = C
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int d = 0; d < 3; ++d) Bm[i][d] += X1[i][d] - X2[i][d];
  // contract with the Hermite basis of the active base-ang polynomial: 3 rows x 12 node values
  const double* bb = P.dyn_ang_basis + 12 * k;
  double bpv[4], bvv[4], bav[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) { bpv[q] = __ldg(bb + q); bvv[q] = __ldg(bb + 4 + q); bav[q] = __ldg(bb + 8 + q); }
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      const int d = j % 3, q = j / 3;
      Sk[3 + i * 12 + j] = fma(C[i][d], bav[q], fma(Bm[i][d], bvv[q], A[i][d] * bpv[q]));
    }
}

// ---- RangeOfMotionConstraint sample (all feet) --------------------------------
// range_of_motion_constraint.cc:58-109: g_e = R^T (p_e - c); Jacobian state R^T and
// D_e = d(R^T r_e)/d(theta) (DerivOfRotVecMult(t, r_W, true)).  The rotation and its derivative are
// computed once per sample and shared by the feet.
// Sk: local state rows 1.. (Sk[0..8] R^T, Sk[9 + 9e ..] D_e); gk: 3 constraint values per foot
template <int kNEE>
__device__ __forceinline__ void RomUnitEval(const SplineSample* __restrict__ sp, const ConstCol xs, const Col Sk, const Col gk) {
  double c[3], th[3], unused[3];
  EvalSpline<0>(sp + 0, xs, c, unused, unused);
  EvalSpline<0>(sp + 1, xs, th, unused, unused);
  const Trig tr = MakeTrig(th);
  double R[3][3]; RotationMatrix(tr, R);
  double dR[3][3][3]; RotationDerivative(tr, dR);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int d = 0; d < 3; ++d) Sk[i * 3 + d] = R[d][i];
#pragma unroll
  for (int e = 0; e < kNEE; ++e) {
    double pe[3];
    EvalSpline<0>(sp + 2 + e, xs, pe, unused, unused);
    const double r[3] = {pe[0] - c[0], pe[1] - c[1], pe[2] - c[2]};
#pragma unroll
    for (int i = 0; i < 3; ++i) gk[3 * e + i] = R[0][i] * r[0] + R[1][i] * r[1] + R[2][i] * r[2];
    double D[3][3]; RotVecDerivative<true>(dR, r, D);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int d = 0; d < 3; ++d) Sk[9 + 9 * e + i * 3 + d] = D[i][d];
  }
}

// ---- analytic terrains: height_map_examples.cc:35-211 ------------------------
struct TerrainPoint { double h, hx, hy, hxx; };
__device__ __forceinline__ TerrainPoint EvalTerrain(int id, double x, double y) {
  TerrainPoint o{0.0, 0.0, 0.0, 0.0};
  switch (id) {
    case 1: {  // Block
      const double start = 0.7, eps = 0.03, len = 3.5, height = 0.5; const double slope = height / eps;
      if (start <= x && x <= start + eps) { o.h = slope * (x - start); o.hx = slope; }
      if (start + eps <= x && x <= start + len) o.h = height;
      break; }
    case 2:  // Stairs
      if (x >= 1.0) o.h = 0.2;
      if (x >= 1.0 + 0.4) o.h = 0.4;
      if (x >= 1.0 + 0.4 + 1.0) o.h = 0.0;
      break;
    case 3: {  // Gap
      const double gs = 1.0, w = 0.5, hh = 1.5; const double xc = gs + w / 2.0;
      const double a = (4 * hh) / (w * w), b = -(8 * hh * xc) / (w * w), c = -(hh * (w - 2 * xc) * (w + 2 * xc)) / (w * w);
      if (gs <= x && x <= gs + w) { o.h = a * x * x + b * x + c; o.hx = 2 * a * x + b; o.hxx = 2 * a; }
      break; }
    case 4: {  // Slope
      const double s0 = 1.0, up = 1.0, down = 1.0, hc = 0.7; const double slope = hc / up;
      if (x >= s0) { o.h = slope * (x - s0); o.hx = slope; }
      if (x >= s0 + up) { o.h = hc - slope * (x - (s0 + up)); o.hx = -slope; }
      if (x >= (s0 + up) + down) { o.h = 0.0; o.hx = 0.0; }
      break; }
    case 5:  // Chimney
      if (1.0 <= x && x <= 1.0 + 1.5) { o.h = 3.0 * (y - 0.5); o.hy = 3.0; }
      break;
    case 6:  // ChimneyLR
      if (0.5 <= x && x <= 0.5 + 1.0) { o.h = 2.0 * (y - 0.5); o.hy = 2.0; }
      if (0.5 + 1.0 <= x && x <= 0.5 + 2 * 1.0) { o.h = -2.0 * (y + 0.5); o.hy = -2.0; }
      break;
    default: break;  // FlatGround(0.0)
  }
  return o;
}

// TerrainConstraint, terrain_constraint.cc:59-108.  Sk: {-dh/dx, -dh/dy}; gk: the constraint value
__device__ __forceinline__ void TerrainUnitEval(const TerrainUnit& u, int terrain, const ConstCol xs, const Col Sk, const Col gk) {
  const double px = xs[u.xi[0]], py = xs[u.xi[1]], pz = xs[u.xi[2]];
  const TerrainPoint tp = EvalTerrain(terrain, px, py);
  gk[0] = pz - tp.h;
  Sk[0] = -tp.hx; Sk[1] = -tp.hy;
}

// normalised vector and HeightMap::GetDerivativeOfNormalizedBasisWrt (height_map.cc:62-91,140-146):
// out[i] = (1/|v|^2 * (|v| * delta(i,dim) - v[dim] * v_hat[i])) * dv[i]   (element-wise, as the reference)
__device__ __forceinline__ void Normalize(const double v[3], double vh[3], double* sn, double* nrm) {
  *sn = v[0] * v[0] + v[1] * v[1] + v[2] * v[2]; *nrm = sqrt(*sn);
  vh[0] = v[0] / *nrm; vh[1] = v[1] / *nrm; vh[2] = v[2] / *nrm;
}
__device__ __forceinline__ void NormalizedDeriv(const double v[3], const double vh[3], double sn, double nrm, int dim,
                                                const double dv[3], double out[3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double unit = (i == dim) ? 1.0 : 0.0;
    out[i] = (1 / sn * (nrm * unit - v[dim] * vh[i])) * dv[i];
  }
}
__device__ __forceinline__ double Dot3(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// ForceConstraint, force_constraint.cc:64-171.  Su: 25 Jacobian values; gr: the 5 constraint values
__device__ __forceinline__ void ForceUnitEval(const Plan& P, const ForceUnit& u, int terrain, const ConstCol xs, const Col Su, const Col gr) {
  const double mu = P.mu;
  const double px = xs[u.xp[0]], py = xs[u.xp[1]];
  const double f[3] = {xs[u.xf[0]], xs[u.xf[1]], xs[u.xf[2]]};
  const TerrainPoint tp = EvalTerrain(terrain, px, py);
  // HeightMap::GetNormal / GetTangent1 / GetTangent2, height_map.cc:93-138
  const double vn[3] = {-tp.hx, -tp.hy, 1.0}, vt1[3] = {1.0, 0.0, tp.hx}, vt2[3] = {0.0, 1.0, tp.hy};
  double n[3], t1[3], t2[3], sn_n, nr_n, sn_1, nr_1, sn_2, nr_2;
  Normalize(vn, n, &sn_n, &nr_n); Normalize(vt1, t1, &sn_1, &nr_1); Normalize(vt2, t2, &sn_2, &nr_2);
  double a1[3], b1[3], a2[3], b2[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    a1[d] = t1[d] - mu * n[d]; b1[d] = t1[d] + mu * n[d];
    a2[d] = t2[d] - mu * n[d]; b2[d] = t2[d] + mu * n[d];
  }
  gr[0] = Dot3(f, n); gr[1] = Dot3(f, a1); gr[2] = Dot3(f, b1); gr[3] = Dot3(f, a2); gr[4] = Dot3(f, b2);
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    Su[0 * 5 + 2 + d] = n[d]; Su[1 * 5 + 2 + d] = a1[d]; Su[2 * 5 + 2 + d] = b1[d];
    Su[3 * 5 + 2 + d] = a2[d]; Su[4 * 5 + 2 + d] = b2[d];
  }
  // second derivatives of the height: only d2h/dx2 exists in the analytic terrains
#pragma unroll
  for (int dim = 0; dim < 2; ++dim) {
    const double hxd = (dim == 0) ? tp.hxx : 0.0;   // d2h/(dx ddim)
    const double hyd = 0.0;                          // d2h/(dy ddim)
    const double dvn[3] = {-hxd, -hyd, 0.0}, dvt1[3] = {0.0, 0.0, hxd}, dvt2[3] = {0.0, 0.0, hyd};
    double dn[3], dt1[3], dt2[3];
    NormalizedDeriv(vn, n, sn_n, nr_n, dim, dvn, dn);
    NormalizedDeriv(vt1, t1, sn_1, nr_1, dim, dvt1, dt1);
    NormalizedDeriv(vt2, t2, sn_2, nr_2, dim, dvt2, dt2);
    double m1[3], p1[3], m2[3], p2[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      m1[d] = dt1[d] - mu * dn[d]; p1[d] = dt1[d] + mu * dn[d];
      m2[d] = dt2[d] - mu * dn[d]; p2[d] = dt2[d] + mu * dn[d];
    }
    Su[0 * 5 + dim] = Dot3(f, dn); Su[1 * 5 + dim] = Dot3(f, m1); Su[2 * 5 + dim] = Dot3(f, p1);
    Su[3 * 5 + dim] = Dot3(f, m2); Su[4 * 5 + dim] = Dot3(f, p2);
  }
}

// SwingConstraint::GetValues, swing_constraint.cc:57-83 (Jacobian is constant)
__device__ __forceinline__ void SwingUnitEval(const SwingUnit& u, const ConstCol xs, const Col gk) {
  const double t_swing_avg = 0.3;
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    const double prev = xs[u.xprev[d]], next = xs[u.xnext[d]];
    const double dist = next - prev;
    const double center = prev + 0.5 * dist;
    const double des_vel = dist / t_swing_avg;
    gk[2 * d] = xs[u.xc_p[d]] - center;
    gk[2 * d + 1] = xs[u.xc_v[d]] - des_vel;
  }
}

// SplineAccConstraint::GetValues, spline_acc_constraint.cc:49-65 (Jacobian constant for fixed durations)
__device__ __forceinline__ void AccUnitEval(const AccUnit& u, const ConstCol xs, const Col gk) {
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double p0 = xs[u.x0 + d], v0 = xs[u.x0 + 3 + d], p1 = xs[u.x0 + 6 + d], v1 = xs[u.x0 + 9 + d];
    const double p2 = xs[u.x0 + 12 + d], v2 = xs[u.x0 + 15 + d];
    const double Cp = DivExact(-(3 * (p0 - p1) + u.Tp * (2 * v0 + v1)), u.Tp2, u.rTp2);
    const double Dp = DivExact(2 * (p0 - p1) + u.Tp * (v0 + v1), u.Tp3, u.rTp3);
    const double Cn = DivExact(-(3 * (p1 - p2) + u.Tn * (2 * v1 + v2)), u.Tn2, u.rTn2);
    const double a_prev = 2 * Cp + (6 * u.Tp) * Dp;
    const double a_next = 2 * Cn;
    gk[d] = a_prev - a_next;
  }
}

// BaseMotionConstraint::UpdateConstraintAtInstance, base_motion_constraint.cc:56-66: rows AX.. = base-ang, LX.. = base-lin position
__device__ __forceinline__ void BaseMotionUnitEval(const Plan& P, const BaseMotionUnit& u, const ConstCol xs, const Col gk) {
  double lin[3], ang[3], unused[3];
  EvalSpline<0>(P.samples + u.sample_lin, xs, lin, unused, unused);
  EvalSpline<0>(P.samples + u.sample_ang, xs, ang, unused, unused);
#pragma unroll
  for (int d = 0; d < 3; ++d) { gk[d] = ang[d]; gk[3 + d] = lin[d]; }
}

// ---- kernels ---------------------------------------------------------------------
// instance b of the tiled iterate matrix with `rows` rows: element r lives at base[((b/32)*rows + r)*32 + b%32]
__device__ __forceinline__ ConstCol TiledCol(const double* base, int b, int rows) {
  return ConstCol{base + ((size_t)(b >> 5) * rows) * 32 + (b & 31)};
}

// x[b][i] -> XT[b/32][i][b%32]: 32x32 tiles through shared memory, coalesced on both sides; clears status
__global__ void __launch_bounds__(256) TransposeIn(const double* __restrict__ x, double* __restrict__ XT,
                                                   int* __restrict__ status, int n, int nb) {
  __shared__ double tile[32][33];
  const int i0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
  if (status && blockIdx.x == 0 && threadIdx.y == 0 && b0 + threadIdx.x < nb) status[b0 + threadIdx.x] = 0;
#pragma unroll
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int b = b0 + r, i = i0 + threadIdx.x;
    if (b < nb && i < n) tile[r][threadIdx.x] = __ldcs(x + (size_t)b * n + i);
  }
  __syncthreads();
  double* dst = XT + ((size_t)blockIdx.y * (n + 1)) * 32;
#pragma unroll
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int i = i0 + r, b = b0 + threadIdx.x;
    if (b < nb && i < n) dst[(size_t)i * 32 + threadIdx.x] = tile[threadIdx.x][r];
  }
}

// ---- warp-collective output of one unit: lanes switch from "instance" to "pair of output elements" ----
// out[j][off + h] = t[d_h][j] * c_h for the instances j = j0, j0 + jstep, .. < n_inst of the tile;
// `out` points at element 0 of the tile's first instance, `stride` is the row length (nnz or m).
__device__ __forceinline__ void StorePairs(const double* t, const OutPair* __restrict__ pairs, const OutCoef* __restrict__ coefs,
                                           int n_pairs, double* __restrict__ out, size_t stride, int j0, int jstep, int n_inst, int lane) {
  if (n_pairs <= 0) return;
  int i = lane;
  OutPair pr{0, kNoRow, kNoRow}; OutCoef cf{0.0, 0.0};
  if (i < n_pairs) { const uint2 raw = __ldg(reinterpret_cast<const uint2*>(pairs) + i); pr.off = (int)raw.x; pr.d0 = raw.y & 0xFFFFu; pr.d1 = raw.y >> 16;
                     const double2 c = __ldg(reinterpret_cast<const double2*>(coefs) + i); cf.c0 = c.x; cf.c1 = c.y; }
  for (; i < n_pairs; i += 32) {
    // prefetch the next entry of this lane under the stores of the current one
    OutPair npr{0, kNoRow, kNoRow}; OutCoef ncf{0.0, 0.0};
    if (i + 32 < n_pairs) { const uint2 raw = __ldg(reinterpret_cast<const uint2*>(pairs) + i + 32); npr.off = (int)raw.x; npr.d0 = raw.y & 0xFFFFu; npr.d1 = raw.y >> 16;
                            const double2 c = __ldg(reinterpret_cast<const double2*>(coefs) + i + 32); ncf.c0 = c.x; ncf.c1 = c.y; }
    double* o = out + pr.off;
    if (pr.d0 != kNoRow && pr.d1 != kNoRow) {
      const double* r0 = t + pr.d0 * kLD; const double* r1 = t + pr.d1 * kLD;
      if (n_inst == 32 && jstep == 1) {
#pragma unroll 16
        for (int j = 0; j < 32; ++j) StoreOut2(o + j * stride, r0[j] * cf.c0, r1[j] * cf.c1);
      } else {
#pragma unroll 4
        for (int j = j0; j < n_inst; j += jstep) StoreOut2(o + j * stride, r0[j] * cf.c0, r1[j] * cf.c1);
      }
    } else if (pr.d0 != kNoRow) {
      const double* r0 = t + pr.d0 * kLD;
      for (int j = j0; j < n_inst; j += jstep) StoreOut(o + j * stride, r0[j] * cf.c0);
    } else if (pr.d1 != kNoRow) {
      const double* r1 = t + pr.d1 * kLD;
      for (int j = j0; j < n_inst; j += jstep) StoreOut(o + 1 + j * stride, r1[j] * cf.c1);
    }
    pr = npr; cf = ncf;
  }
}
// all outputs of one unit: Jacobian values (flags & 2) and constraint values (flags & 1)
__device__ __forceinline__ void StoreUnit(const Plan& P, const double* t, const OutList& L, double* __restrict__ g_tile,
                                          double* __restrict__ jac_tile, unsigned flags, int n_inst, int lane) {
  if (flags & 2u) {
    if (P.nnz & 1) {
      StorePairs(t, P.pairs + L.jac[0], P.coefs + L.jac[0], L.n_jac[0], jac_tile, (size_t)P.nnz, 0, 2, n_inst, lane);
      StorePairs(t, P.pairs + L.jac[1], P.coefs + L.jac[1], L.n_jac[1], jac_tile, (size_t)P.nnz, 1, 2, n_inst, lane);
    } else {
      StorePairs(t, P.pairs + L.jac[0], P.coefs + L.jac[0], L.n_jac[0], jac_tile, (size_t)P.nnz, 0, 1, n_inst, lane);
    }
  }
  if (flags & 1u) {
    if (P.m & 1) {
      StorePairs(t, P.pairs + L.g[0], P.coefs + L.g[0], L.n_g[0], g_tile, (size_t)P.m, 0, 2, n_inst, lane);
      StorePairs(t, P.pairs + L.g[1], P.coefs + L.g[1], L.n_g[1], g_tile, (size_t)P.m, 1, 2, n_inst, lane);
    } else {
      StorePairs(t, P.pairs + L.g[0], P.coefs + L.g[0], L.n_g[0], g_tile, (size_t)P.m, 0, 1, n_inst, lane);
    }
  }
}
// non-finite check of this lane's own column (rows 1 .. n_rows-1); flags instance b
__device__ __forceinline__ void FlagNonFinite(const double* t, int n_rows, int lane, int* __restrict__ status, int b, int nb) {
  if (!status) return;
  double chk = 0.0;
  for (int r = 1; r < n_rows; ++r) chk = fma(t[r * kLD + lane], 0.0, chk);
  if (chk != chk && b < nb) atomicOr(status + b, 1);
}

// DynamicConstraint: blockIdx.y = instance tile, warp = one of kDynWarps CONSECUTIVE samples, so a CTA writes
// several KB of contiguous CSR values per instance (the samples' rows are adjacent) — DRAM page locality.
template <int kNEE>
__global__ void __launch_bounds__(kDynWarps * 32, TWB_DYN_CTAS) DynOut(const Plan P, const double* __restrict__ XT, double* __restrict__ g,
                                                           double* __restrict__ jac, int* __restrict__ status, int nb, unsigned flags) {
  extern __shared__ __align__(16) double out_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k = blockIdx.x * kDynWarps + warp, b0 = blockIdx.y * 32;
  if (k >= P.n_dyn) return;
  constexpr int G0 = 40 + 6 * kNEE, n_rows = G0 + 6;   // local rows: 1 | 3 | 36 | 6 per foot | g (6)
  double* t = out_smem + (size_t)warp * n_rows * kLD;
  const DynUnit* u = P.dyn + k;
  t[lane] = 1.0;
  DynamicUnit<kNEE>(P, k, P.samples + __ldg(&u->sample0), TiledCol(XT, b0 + lane, P.n + 1), Col{t + kLD + lane, kLD}, Col{t + G0 * kLD + lane, kLD});
  FlagNonFinite(t, n_rows, lane, status, b0 + lane, nb);
  __syncwarp();
  StoreUnit(P, t, u->out, g + (size_t)b0 * P.m, jac + (size_t)b0 * P.nnz, flags, min(32, nb - b0), lane);
}

// RangeOfMotionConstraint: blockIdx.y = instance tile, warp = one of kRomWarps consecutive samples, all feet
template <int kNEE>
__global__ void __launch_bounds__(kRomWarps * 32, TWB_ROM_CTAS) RomOut(const Plan P, const double* __restrict__ XT, double* __restrict__ g,
                                                           double* __restrict__ jac, int* __restrict__ status, int nb, unsigned flags) {
  extern __shared__ __align__(16) double out_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k = blockIdx.x * kRomWarps + warp, b0 = blockIdx.y * 32;
  if (k >= P.n_rom) return;
  constexpr int G0 = 10 + 9 * kNEE, n_rows = G0 + 3 * kNEE;   // local rows: 1 | R^T 9 | D_e 9 per foot | g 3 per foot
  double* t = out_smem + (size_t)warp * n_rows * kLD;
  const RomUnit* u = P.rom + k;
  t[lane] = 1.0;
  RomUnitEval<kNEE>(P.samples + __ldg(&u->sample0), TiledCol(XT, b0 + lane, P.n + 1), Col{t + kLD + lane, kLD}, Col{t + G0 * kLD + lane, kLD});
  FlagNonFinite(t, n_rows, lane, status, b0 + lane, nb);
  __syncwarp();
  StoreUnit(P, t, u->out, g + (size_t)b0 * P.m, jac + (size_t)b0 * P.nnz, flags, min(32, nb - b0), lane);
}

// node groups: blockIdx.y = instance tile, warp = one of kNodeWarps consecutive groups
__global__ void __launch_bounds__(kNodeWarps * 32) NodeOut(const Plan P, const double* __restrict__ XT, double* __restrict__ g,
                                                           double* __restrict__ jac, int* __restrict__ status,
                                                           const int* __restrict__ terrain_ids, int default_terrain, int nb,
                                                           unsigned flags) {
  extern __shared__ __align__(16) double node_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gi = blockIdx.x * kNodeWarps + warp, b0 = blockIdx.y * 32, b = b0 + lane;
  if (gi >= P.n_groups) return;
  double* t = node_smem + (size_t)warp * kNodeStateRows * kLD;
  const ConstCol xs = TiledCol(XT, b, P.n + 1);
  const NodeGroup* grp = P.groups + gi;
  const int kind = __ldg(&grp->kind), first = __ldg(&grp->first), count = __ldg(&grp->count);
  t[lane] = 1.0;
  int n_rows = 1;
  if (kind == kGroupForce) {
    const int terrain = (terrain_ids && b < nb) ? __ldg(terrain_ids + b) : default_terrain;
    for (int q = 0; q < count; ++q)
      ForceUnitEval(P, P.force[first + q], terrain, xs, Col{t + (1 + 25 * q) * kLD + lane, kLD}, Col{t + (1 + 25 * count + 5 * q) * kLD + lane, kLD});
    n_rows = 1 + 30 * count;
  } else if (kind == kGroupTerrain) {
    const int terrain = (terrain_ids && b < nb) ? __ldg(terrain_ids + b) : default_terrain;
    for (int q = 0; q < count; ++q)
      TerrainUnitEval(P.terr[first + q], terrain, xs, Col{t + (1 + 2 * q) * kLD + lane, kLD}, Col{t + (1 + 2 * count + q) * kLD + lane, kLD});
    n_rows = 1 + 3 * count;
  } else if (kind == kGroupSwing) {
    if (flags & 1u) for (int q = 0; q < count; ++q) SwingUnitEval(P.swing[first + q], xs, Col{t + (1 + 4 * q) * kLD + lane, kLD});
    n_rows = (flags & 1u) ? 1 + 4 * count : 1;
  } else if (kind == kGroupAcc) {
    if (flags & 1u) for (int q = 0; q < count; ++q) AccUnitEval(P.acc[first + q], xs, Col{t + (1 + 3 * q) * kLD + lane, kLD});
    n_rows = (flags & 1u) ? 1 + 3 * count : 1;
  } else if (kind == kGroupBaseMotion) {
    if (flags & 1u) for (int q = 0; q < count; ++q) BaseMotionUnitEval(P, P.base_motion[first + q], xs, Col{t + (1 + 6 * q) * kLD + lane, kLD});
    n_rows = (flags & 1u) ? 1 + 6 * count : 1;
  }
  FlagNonFinite(t, n_rows, lane, status, b, nb);
  __syncwarp();
  StoreUnit(P, t, grp->out, g + (size_t)b0 * P.m, jac + (size_t)b0 * P.nnz, flags, min(32, nb - b0), lane);
}

// NodeCost::GetCost summed over terms (node_cost.cc:53-63; Composite::GetValues for costs) and the
// dense gradient row (node_cost.cc:65-76), both in the reference's order.  One thread per instance.
__global__ void __launch_bounds__(128) CostKernel(const Plan P, const double* __restrict__ XT, double* __restrict__ cost,
                                                  double* __restrict__ grad, int nb) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  const ConstCol xs = TiledCol(XT, b, P.n + 1);
  double* gr = grad ? grad + (size_t)b * P.n : nullptr;
  if (gr) for (int i = 0; i < P.n; ++i) gr[i] = 0.0;
  double total_cost = 0.0, term = 0.0;
  for (int i = 0; i < P.n_cost; ++i) {
    const CostEntry ce = P.cost[i];
    if (ce.pad && i > 0) { total_cost += term; term = 0.0; }
    const double val = xs[ce.xi];
    term += ce.weight * (val * val);
    if (gr && ce.grad_col >= 0) gr[ce.grad_col] += ce.weight * 2.0 * val;
  }
  total_cost += term;
  if (cost) cost[b] = total_cost;
}

template <int kNEE>
cudaError_t LaunchDynRom(const Plan& P, const double* XT, double* g, double* jac, int* status, int nb, unsigned flags, int tiles,
                         cudaStream_t s_dyn, cudaStream_t s_rom, int* count) {
  cudaError_t e = cudaSuccess;
  if (P.n_dyn > 0) {
    const size_t smem = (size_t)kDynWarps * (46 + 6 * kNEE) * kLD * sizeof(double);
    e = cudaFuncSetAttribute(DynOut<kNEE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    DynOut<kNEE><<<dim3((P.n_dyn + kDynWarps - 1) / kDynWarps, tiles), kDynWarps * 32, smem, s_dyn>>>(P, XT, g, jac, status, nb, flags);
    ++*count; TWB_MARK("DynOut", s_dyn);
  }
  if (P.n_rom > 0) {
    const size_t smem = (size_t)kRomWarps * (10 + 12 * kNEE) * kLD * sizeof(double);
    e = cudaFuncSetAttribute(RomOut<kNEE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    RomOut<kNEE><<<dim3((P.n_rom + kRomWarps - 1) / kRomWarps, tiles), kRomWarps * 32, smem, s_rom>>>(P, XT, g, jac, status, nb, flags);
    ++*count; TWB_MARK("RomOut", s_rom);
  }
  return cudaSuccess;
}

}  // namespace

// ---- host launchers ------------------------------------------------------------------

// XT is the tiled iterate matrix of the whole batch (first tile = first instance of x / g / jac).
// Streams: `s` carries TransposeIn -> RomOut; DynOut / NodeOut (+ CostKernel) run on aux[0] / aux[1]
// after the transposition (ev[0]) and are joined back into `s` (ev[1], ev[2]).
int LaunchEval(const Plan& P, const double* x, double* XT, double* g, double* jac, double* cost, double* grad,
               int* status, const int* terrain_ids, int default_terrain, int nb, unsigned flags, cudaStream_t s,
               cudaStream_t aux0, cudaStream_t aux1, cudaEvent_t* ev, int* launches) {
  if (nb <= 0) return 0;
  int count = 0;
  const bool serial = (g_after_launch != nullptr);
  if (serial) aux0 = aux1 = s;
  const int tiles = (nb + 31) / 32;
  const unsigned out_flags = flags & 3u;
  TWB_MARK("begin", s);
  TransposeIn<<<dim3((P.n + 31) / 32, tiles), dim3(32, 8), 0, s>>>(x, XT, status, P.n, nb); ++count; TWB_MARK("TransposeIn", s);
  if (!serial) { cudaEventRecord(ev[0], s); cudaStreamWaitEvent(aux0, ev[0], 0); cudaStreamWaitEvent(aux1, ev[0], 0); }
  cudaError_t e = cudaSuccess;
  if (out_flags) {
    switch (P.n_ee) {
      case 1: e = LaunchDynRom<1>(P, XT, g, jac, status, nb, out_flags, tiles, aux0, s, &count); break;
      case 2: e = LaunchDynRom<2>(P, XT, g, jac, status, nb, out_flags, tiles, aux0, s, &count); break;
      case 4: e = LaunchDynRom<4>(P, XT, g, jac, status, nb, out_flags, tiles, aux0, s, &count); break;
      default: return (int)cudaErrorInvalidValue;
    }
    if (e != cudaSuccess) return (int)e;
    if (P.n_groups > 0) {
      const size_t smem = (size_t)kNodeWarps * kNodeStateRows * kLD * sizeof(double);
      e = cudaFuncSetAttribute(NodeOut, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
      NodeOut<<<dim3((P.n_groups + kNodeWarps - 1) / kNodeWarps, tiles), kNodeWarps * 32, smem, aux1>>>(P, XT, g, jac, status, terrain_ids, default_terrain, nb, out_flags);
      ++count; TWB_MARK("NodeOut", aux1);
    }
  }
  if ((flags & 4u) && P.n_cost > 0) {
    CostKernel<<<(nb + 127) / 128, 128, 0, aux1>>>(P, XT, cost, grad, nb); ++count; TWB_MARK("CostKernel", aux1);
  }
  if (!serial) {
    cudaEventRecord(ev[1], aux0); cudaEventRecord(ev[2], aux1);
    cudaStreamWaitEvent(s, ev[1], 0); cudaStreamWaitEvent(s, ev[2], 0);
  }
  if (launches) *launches += count;
  return (int)cudaGetLastError();
}

}  // namespace twb
