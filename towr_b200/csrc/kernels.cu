// kernels.cu — sm_100a kernels of the batched NLP evaluation (fp64).
//
// Everything is a data-parallel map with LANE = PROBLEM INSTANCE.  All per-iterate
// intermediates live in one state matrix, stored instance-tiled: ST[tile][S_size][32] (tile = 32
// consecutive instances, row = state slot, lane = instance), so every global access of every kernel
// is a 256-byte row segment and all rows of a tile sit in one compact region (DRAM page / L2 locality):
//
//   TransposeIn   x[B][n]            -> XT[tile][n+1][32]     (row n stays 0: "not optimised" node values)
//   SplineKernel  XT                 -> ST rows (spline values at every constraint sample)
//   DynKernel     ST rows            -> ST rows (g of the dynamic constraint + its Jacobian state)
//   RomKernel     ST rows            -> ST rows (g of the range-of-motion constraints + Jacobian state)
//   NodeKernel    XT                 -> ST rows (terrain / force / swing / spline-acc rows, cost)
//   FillJac       ST, desc, coef     -> jac[B][nnz]           jac[b][s] = ST[desc[s]][b] * coef[s]
//   TransposeOut  ST rows 1..m       -> g[B][m]
//
// The first five are ALU/latency-type kernels (fp64 FMA pipe, sincos); FillJac carries ~95% of the
// HBM traffic and is a pure stream.  capi.cc pipelines sub-batches over two streams so that the
// state kernels of sub-batch i+1 run under the HBM-bound fill of sub-batch i.
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

#include <cstdint>

namespace twb {
namespace {

constexpr int kThreads = 256;

// output stores: TWB_STORE_MODE 0 = default write-back, 1 = streaming (evict-first), 2 = write-through
#ifndef TWB_STORE_MODE
#define TWB_STORE_MODE 1
#endif
template <class T>
__device__ __forceinline__ void StoreOut(T* p, T v) {
#if TWB_STORE_MODE == 1
  __stcs(p, v);
#elif TWB_STORE_MODE == 2
  __stwt(p, v);
#else
  *p = v;
#endif
}

// column `b` of a row-major [rows][ld] matrix: element r lives at p[r * ld]
struct Col {
  double* p;
  size_t ld;
  __device__ __forceinline__ double& operator[](int r) const { return p[(size_t)r * ld]; }
  __device__ __forceinline__ Col at(int r) const { return Col{p + (size_t)r * ld, ld}; }
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
// kind 0: position; 1: position + acceleration; 2: position + velocity + acceleration.
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
__device__ __forceinline__ void EvalSplineToState(const SampleRegs& s, int kind, const Col xs, const Col out) {
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double p0 = xs[s.xi[d]], v0 = xs[s.xi[3 + d]], p1 = xs[s.xi[6 + d]], v1 = xs[s.xi[9 + d]];
    const double C = DivExact(-(3 * (p0 - p1) + s.T * (2 * v0 + v1)), s.T2, s.rT2);
    const double D = DivExact(2 * (p0 - p1) + s.T * (v0 + v1), s.T3, s.rT3);
    out[d] = ((p0 + s.t * v0) + s.t2 * C) + s.t3 * D;
    if (kind == 1) out[3 + d] = 2 * C + (6 * s.t) * D;
    if (kind == 2) {
      out[3 + d] = (v0 + (2 * s.t) * C) + (3 * s.t2) * D;
      out[6 + d] = 2 * C + (6 * s.t) * D;
    }
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
__device__ __forceinline__ void DynamicUnit(const Plan& P, int k, const Col S) {
  const int n_ee = P.n_ee;
  const Col Sk = S.at(P.S_dyn0 + k * P.S_dyn_stride);
  // phase-0 scratch: c, c_ddot, theta, theta_dot, theta_ddot, p_e.., f_e..
  double c[3], cdd[3], th[3], thd[3], thdd[3], pe[kMaxEE][3], fe[kMaxEE][3];
#pragma unroll
  for (int d = 0; d < 3; ++d) { c[d] = Sk[d]; cdd[d] = Sk[3 + d]; th[d] = Sk[6 + d]; thd[d] = Sk[9 + d]; thdd[d] = Sk[12 + d]; }
#pragma unroll
  for (int e = 0; e < kMaxEE; ++e)
    if (e < n_ee) {
#pragma unroll
      for (int d = 0; d < 3; ++d) { pe[e][d] = Sk[15 + 3 * e + d]; fe[e][d] = Sk[15 + 3 * n_ee + 3 * e + d]; }
    }

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

  // feet
  double fsum[3] = {0, 0, 0}, tau[3] = {0, 0, 0};
#pragma unroll
  for (int e = 0; e < kMaxEE; ++e) {
    if (e >= n_ee) break;
    const double* f = fe[e];
    const double r[3] = {c[0] - pe[e][0], c[1] - pe[e][1], c[2] - pe[e][2]};
    tau[0] += f[1] * r[2] - f[2] * r[1];
    tau[1] += f[2] * r[0] - f[0] * r[2];
    tau[2] += f[0] * r[1] - f[1] * r[0];
#pragma unroll
    for (int d = 0; d < 3; ++d) { fsum[d] += f[d]; Sk[39 + e * 6 + d] = f[d]; Sk[39 + e * 6 + 3 + d] = r[d]; }
  }
  Sk[0] = fsum[0]; Sk[1] = fsum[1]; Sk[2] = fsum[2];

  double Iw_om[3], Iw_omd[3];
  MulVec(Iw, om, Iw_om); MulVec(Iw, omd, Iw_omd);
  {
    const double wx[3] = {om[1] * Iw_om[2] - om[2] * Iw_om[1], om[2] * Iw_om[0] - om[0] * Iw_om[2], om[0] * Iw_om[1] - om[1] * Iw_om[0]};
    const Col gk = S.at(P.S_g0 + P.dyn_row0 + 6 * k);
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
// range_of_motion_constraint.cc:58-109: g = R^T (p_ee - c); Jacobian state R^T and
// D_e = d(R^T r_e)/d(theta) (DerivOfRotVecMult(t, r_W, true)).
__device__ __forceinline__ void RomUnit(const Plan& P, int k, const Col S) {
  const int n_ee = P.n_ee;
  const Col Sk = S.at(P.S_rom0 + k * P.S_rom_stride);
  double c[3], th[3], pes[kMaxEE][3];
#pragma unroll
  for (int d = 0; d < 3; ++d) { c[d] = Sk[d]; th[d] = Sk[3 + d]; }
#pragma unroll
  for (int e = 0; e < kMaxEE; ++e)
    if (e < n_ee) {
#pragma unroll
      for (int d = 0; d < 3; ++d) pes[e][d] = Sk[6 + 3 * e + d];
    }
  const Trig tr = MakeTrig(th);
  double R[3][3]; RotationMatrix(tr, R);
  double dR[3][3][3]; RotationDerivative(tr, dR);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int d = 0; d < 3; ++d) Sk[i * 3 + d] = R[d][i];
#pragma unroll
  for (int e = 0; e < kMaxEE; ++e) {
    if (e >= n_ee) break;
    const double* pe = pes[e];
    const double r[3] = {pe[0] - c[0], pe[1] - c[1], pe[2] - c[2]};
    {
      const Col ge = S.at(P.S_g0 + P.rom_row0[e] + 3 * k);
#pragma unroll
      for (int i = 0; i < 3; ++i) ge[i] = R[0][i] * r[0] + R[1][i] * r[1] + R[2][i] * r[2];
    }
    double D[3][3]; RotVecDerivative<true>(dR, r, D);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int d = 0; d < 3; ++d) Sk[9 + e * 9 + i * 3 + d] = D[i][d];
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

// TerrainConstraint, terrain_constraint.cc:59-108
__device__ __forceinline__ void TerrainUnitEval(const Plan& P, const TerrainUnit& u, int terrain, const Col xs, const Col S) {
  const double px = xs[u.xi[0]], py = xs[u.xi[1]], pz = xs[u.xi[2]];
  const TerrainPoint tp = EvalTerrain(terrain, px, py);
  S[P.S_g0 + u.g_row] = pz - tp.h;
  S[u.s_idx + 0] = -tp.hx; S[u.s_idx + 1] = -tp.hy;
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

// ForceConstraint, force_constraint.cc:64-171
__device__ __forceinline__ void ForceUnitEval(const Plan& P, const ForceUnit& u, int terrain, const Col xs, const Col S) {
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
  {
    const Col gr = S.at(P.S_g0 + u.g_row);
    gr[0] = Dot3(f, n); gr[1] = Dot3(f, a1); gr[2] = Dot3(f, b1); gr[3] = Dot3(f, a2); gr[4] = Dot3(f, b2);
  }
  const Col Su = S.at(u.s_idx);
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
__device__ __forceinline__ void SwingUnitEval(const Plan& P, const SwingUnit& u, const Col xs, const Col S) {
  const Col g = S.at(P.S_g0);
  const double t_swing_avg = 0.3;
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    const double prev = xs[u.xprev[d]], next = xs[u.xnext[d]];
    const double dist = next - prev;
    const double center = prev + 0.5 * dist;
    const double des_vel = dist / t_swing_avg;
    g[u.g_row + 2 * d] = xs[u.xc_p[d]] - center;
    g[u.g_row + 2 * d + 1] = xs[u.xc_v[d]] - des_vel;
  }
}

// SplineAccConstraint::GetValues, spline_acc_constraint.cc:49-65 (Jacobian constant for fixed durations)
__device__ __forceinline__ void AccUnitEval(const Plan& P, const AccUnit& u, const Col xs, const Col S) {
  const Col g = S.at(P.S_g0);
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double p0 = xs[u.x0 + d], v0 = xs[u.x0 + 3 + d], p1 = xs[u.x0 + 6 + d], v1 = xs[u.x0 + 9 + d];
    const double p2 = xs[u.x0 + 12 + d], v2 = xs[u.x0 + 15 + d];
    const double Cp = DivExact(-(3 * (p0 - p1) + u.Tp * (2 * v0 + v1)), u.Tp2, u.rTp2);
    const double Dp = DivExact(2 * (p0 - p1) + u.Tp * (v0 + v1), u.Tp3, u.rTp3);
    const double Cn = DivExact(-(3 * (p1 - p2) + u.Tn * (2 * v1 + v2)), u.Tn2, u.rTn2);
    const double a_prev = 2 * Cp + (6 * u.Tp) * Dp;
    const double a_next = 2 * Cn;
    g[u.g_row + d] = a_prev - a_next;
  }
}

// ---- kernels ---------------------------------------------------------------------
// instance b of a tiled matrix with `rows` rows: element r lives at base[((b/32)*rows + r)*32 + b%32]
__device__ __forceinline__ Col TiledCol(double* base, int b, int rows) {
  return Col{base + ((size_t)(b >> 5) * rows) * 32 + (b & 31), 32};
}

// x[b][i] -> XT[b/32][i][b%32]: 32x32 tiles through shared memory, coalesced on both sides
__global__ void __launch_bounds__(256) TransposeIn(const double* __restrict__ x, double* __restrict__ XT, int n, int nb) {
  __shared__ double tile[32][33];
  const int i0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int b = b0 + r, i = i0 + threadIdx.x;
    if (b < nb && i < n) tile[r][threadIdx.x] = x[(size_t)b * n + i];
  }
  __syncthreads();
  double* dst = XT + ((size_t)blockIdx.y * (n + 1)) * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int i = i0 + r, b = b0 + threadIdx.x;
    if (b < nb && i < n) dst[(size_t)i * 32 + threadIdx.x] = tile[threadIdx.x][r];
  }
}

// ST[b/32][row0 + r][b%32] -> out[b][r], r < rows  (constraint values g, cost gradient)
__global__ void __launch_bounds__(256) TransposeOut(const double* __restrict__ ST, double* __restrict__ out, int S_size,
                                                    int row0, int rows, int nb) {
  __shared__ double tile[32][33];
  const int r0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
  const double* src = ST + ((size_t)blockIdx.y * S_size + row0) * 32;
  for (int q = threadIdx.y; q < 32; q += 8) {
    const int r = r0 + q, b = b0 + threadIdx.x;
    if (r < rows && b < nb) tile[q][threadIdx.x] = src[(size_t)r * 32 + threadIdx.x];
  }
  __syncthreads();
  for (int q = threadIdx.y; q < 32; q += 8) {
    const int b = b0 + q, r = r0 + threadIdx.x;
    if (r < rows && b < nb) StoreOut(out + (size_t)b * rows + r, tile[threadIdx.x][q]);
  }
}

// one thread per (eval item = blockIdx.y, instance): all lanes of a warp share the sample table entry
__global__ void __launch_bounds__(128) SplineKernel(const Plan P, double* __restrict__ XT, double* __restrict__ ST,
                                                    int nb) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  const uint2 raw = __ldg(reinterpret_cast<const uint2*>(P.eval_items) + blockIdx.y);
  const int sample = (int)raw.x, row = (int)(raw.y & 0xFFFFu), kind = (int)(raw.y >> 16);
  const SampleRegs sr = LoadSample(P.samples + sample);
  EvalSplineToState(sr, kind, TiledCol(XT, b, P.n + 1), TiledCol(ST, b, P.S_size).at(row));
}

// one thread per (dynamic sample = blockIdx.y, instance)
__global__ void __launch_bounds__(128, 2) DynKernel(const Plan P, double* __restrict__ ST, int nb) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  DynamicUnit(P, blockIdx.y, TiledCol(ST, b, P.S_size));
}

// one thread per (range-of-motion sample = blockIdx.y, instance)
__global__ void __launch_bounds__(128, 2) RomKernel(const Plan P, double* __restrict__ ST, int nb) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  RomUnit(P, blockIdx.y, TiledCol(ST, b, P.S_size));
}

// one thread per (node unit = blockIdx.y, instance): force | terrain | swing | spline-acc | cost
__global__ void __launch_bounds__(128) NodeKernel(const Plan P, double* __restrict__ XT, double* __restrict__ ST,
                                                  const int* __restrict__ terrain_ids, int default_terrain,
                                                  double* __restrict__ cost, int nb, int want_cost) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  const Col xs = TiledCol(XT, b, P.n + 1), S = TiledCol(ST, b, P.S_size);
  int u = blockIdx.y;
  if (u < P.n_force) {
    ForceUnitEval(P, P.force[u], terrain_ids ? terrain_ids[b] : default_terrain, xs, S);
    return;
  }
  u -= P.n_force;
  if (u < P.n_terr) {
    TerrainUnitEval(P, P.terr[u], terrain_ids ? terrain_ids[b] : default_terrain, xs, S);
    return;
  }
  u -= P.n_terr;
  if (u < P.n_swing) { SwingUnitEval(P, P.swing[u], xs, S); return; }
  u -= P.n_swing;
  if (u < P.n_acc) { AccUnitEval(P, P.acc[u], xs, S); return; }
  if (want_cost) {
    // NodeCost::GetCost summed over terms (node_cost.cc:53-63; Composite::GetValues for costs) and the
    // dense gradient row (node_cost.cc:65-76), both in the reference's order; gradient rows live in ST.
    const Col gr = S.at(P.S_grad0);
    for (int i = 0; i < P.n; ++i) gr[i] = 0.0;
    double total_cost = 0.0, term = 0.0;
    for (int i = 0; i < P.n_cost; ++i) {
      const CostEntry ce = P.cost[i];
      if (ce.pad && i > 0) { total_cost += term; term = 0.0; }
      const double val = xs[ce.xi];
      term += ce.weight * (val * val);
      if (ce.grad_col >= 0) gr[ce.grad_col] += ce.weight * 2.0 * val;
    }
    total_cost += term;
    if (cost) cost[b] = total_cost;
  }
}

// jac[b][s] = ST[tile][desc[s]][b] * coef[s]   — gather state rows, transpose, scale, stream out.
// A CTA owns 32 instances (one tile of ST) and one chunk of kFillChunkSlots CSR slots.
//  stage: the chunk's DISTINCT state rows (host-built list, ~4x fewer than slots) are copied with
//         cp.async, 256 contiguous bytes per row (lane = instance), into a padded shared-memory
//         array — dozens of row copies in flight per thread, no registers held;
//  store: lane = slot pair: it reads its two rows for one instance (conflict-free: row stride 33),
//         scales by its two constants and writes 16 bytes, so each store instruction covers 512
//         contiguous bytes of one instance's CSR value array.
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}
template <bool kVec>
__global__ void __launch_bounds__(kThreads) FillJac(const Plan P, const double* __restrict__ ST, double* __restrict__ jac,
                                                    int* __restrict__ status, int nb) {
  extern __shared__ __align__(16) double fill_rows_smem[];
  double (*t)[33] = reinterpret_cast<double (*)[33]>(fill_rows_smem);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nnz = P.nnz;
  const int chunk = blockIdx.x;                                     // chunk fastest: the CTAs of one instance
  const int b0 = blockIdx.y * 32;                                   // tile run together (L2 locality of its rows)
  const double* Sb = ST + ((size_t)blockIdx.y * P.S_size) * 32 + lane;
  const int r_begin = __ldg(P.fill_row_off + chunk), n_rows = __ldg(P.fill_row_off + chunk + 1) - r_begin;
  for (int r = warp; r < n_rows; r += kThreads / 32)
    cp_async8(&t[r][lane], Sb + (size_t)__ldg(P.fill_rows + r_begin + r) * 32);
  const int s_begin = chunk * kFillChunkSlots, s_end = min(nnz, s_begin + kFillChunkSlots);
  cp_async_wait_all();
  __syncthreads();
  if (status) {   // every output is (staged value) x (finite constant): check the staged rows once
    double chk = 0.0;
    for (int r = warp; r < n_rows; r += kThreads / 32) chk = fma(t[r][lane], 0.0, chk);
    if (chk != chk && b0 + lane < nb) atomicOr(status + b0 + lane, 1);   // NaN or Inf in instance b0 + lane
  }
  const int n_inst = min(32, nb - b0);
  if (kVec) {
    for (int s = s_begin + 2 * threadIdx.x; s < s_end; s += 2 * kThreads) {
      const uint32_t loc = __ldg(reinterpret_cast<const uint32_t*>(P.fill_local + s));   // two 16-bit row positions
      const double2 cf = __ldg(reinterpret_cast<const double2*>(P.coef + s));
      const double* r0 = t[loc & 0xFFFFu];
      const double* r1 = t[loc >> 16];
      double2* out = reinterpret_cast<double2*>(jac + (size_t)b0 * nnz + s);
#pragma unroll 8
      for (int j = 0; j < n_inst; ++j) StoreOut(out + (size_t)j * (nnz >> 1), make_double2(r0[j] * cf.x, r1[j] * cf.y));
    }
  } else {
    for (int s = s_begin + threadIdx.x; s < s_end; s += kThreads) {
      const double cf = __ldg(P.coef + s);
      const double* r0 = t[__ldg(P.fill_local + s)];
      double* out = jac + (size_t)b0 * nnz + s;
      for (int j = 0; j < n_inst; ++j) StoreOut(out + (size_t)j * nnz, r0[j] * cf);
    }
  }
}

__global__ void ClearStatus(int* __restrict__ status, int nb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nb) status[i] = 0;
}

__global__ void InitOnes(double* __restrict__ ST, int S_size, int n_tiles) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_tiles * 32) ST[((size_t)(i >> 5) * S_size) * 32 + (i & 31)] = 1.0;   // state row 0 == 1
}

}  // namespace

// ---- host launchers ------------------------------------------------------------------
void (*g_after_launch)(const char* label, cudaStream_t stream) = nullptr;   // profiling hook (capi.cc, TWB_PROFILE=1)
#define TWB_MARK(label) do { if (g_after_launch) g_after_launch(label, stream); } while (0)
// XT / ST point at the first TILE of the sub-batch (sub-batches start at multiples of 32 instances).
int LaunchInitState(double* ST, int S_size, int n_tiles, cudaStream_t stream) {
  InitOnes<<<(n_tiles * 32 + 255) / 256, 256, 0, stream>>>(ST, S_size, n_tiles);
  return (int)cudaGetLastError();
}

int LaunchStateKernels(const Plan& P, const double* x, double* XT, double* ST, const int* terrain_ids,
                       int default_terrain, double* cost, int* status, int nb, bool want_cost, cudaStream_t stream,
                       int* launches) {
  if (nb <= 0) return 0;
  int count = 0;
  const int bx = (nb + 127) / 128;
  TWB_MARK("begin");
  if (status) { ClearStatus<<<(nb + 255) / 256, 256, 0, stream>>>(status, nb); ++count; TWB_MARK("ClearStatus"); }
  TransposeIn<<<dim3((P.n + 31) / 32, (nb + 31) / 32), dim3(32, 8), 0, stream>>>(x, XT, P.n, nb); ++count; TWB_MARK("TransposeIn");
  if (P.n_eval_items > 0) { SplineKernel<<<dim3(bx, P.n_eval_items), 128, 0, stream>>>(P, XT, ST, nb); ++count; TWB_MARK("SplineKernel"); }
  if (P.n_dyn > 0) { DynKernel<<<dim3(bx, P.n_dyn), 128, 0, stream>>>(P, ST, nb); ++count; TWB_MARK("DynKernel"); }
  if (P.n_rom > 0) { RomKernel<<<dim3(bx, P.n_rom), 128, 0, stream>>>(P, ST, nb); ++count; TWB_MARK("RomKernel"); }
  const int n_units = P.n_force + P.n_terr + P.n_swing + P.n_acc + (want_cost ? 1 : 0);
  if (n_units > 0) {
    NodeKernel<<<dim3(bx, n_units), 128, 0, stream>>>(P, XT, ST, terrain_ids, default_terrain, cost, nb, want_cost ? 1 : 0);
    ++count; TWB_MARK("NodeKernel");
  }
  if (launches) *launches += count;
  return (int)cudaGetLastError();
}

int LaunchFillJac(const Plan& P, const double* ST, double* jac, int* status, int nb, int n_sms, cudaStream_t stream,
                  int* launches) {
  if (nb <= 0) return 0;
  (void)n_sms;
  const bool vec = ((P.nnz & 1) == 0) && ((reinterpret_cast<uintptr_t>(jac) & 15) == 0);
  const size_t smem = (size_t)P.fill_max_rows * 33 * sizeof(double);
  const dim3 grid(P.fill_chunks, (nb + 31) / 32);
  cudaError_t e;
  TWB_MARK("begin");
  if (vec) {
    e = cudaFuncSetAttribute(FillJac<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    FillJac<true><<<grid, kThreads, smem, stream>>>(P, ST, jac, status, nb);
  } else {
    e = cudaFuncSetAttribute(FillJac<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    FillJac<false><<<grid, kThreads, smem, stream>>>(P, ST, jac, status, nb);
  }
  TWB_MARK("FillJac");
  if (launches) *launches += 1;
  return (int)cudaGetLastError();
}

int LaunchTransposeOut(const Plan& P, const double* ST, int row0, int rows, double* out, int nb, cudaStream_t stream,
                       int* launches) {
  if (nb <= 0 || rows <= 0) return 0;
  TWB_MARK("begin");
  TransposeOut<<<dim3((rows + 31) / 32, (nb + 31) / 32), dim3(32, 8), 0, stream>>>(ST, out, P.S_size, row0, rows, nb);
  TWB_MARK("TransposeOut");
  if (launches) *launches += 1;
  return (int)cudaGetLastError();
}

}  // namespace twb
