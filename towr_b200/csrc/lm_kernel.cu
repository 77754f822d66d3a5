// lm_kernel.cu — one Levenberg-Marquardt feasibility step for every instance of a batch, entirely on the device.
//
// The stand-in for the IPOPT solves BASELINE configs[0] / [4] name (hopper_example.cc:77-93, footstep_plan_server.cc:222-236;
// neither IPOPT nor ifopt exists in this image): what towr_b200/solver.py documents and tests/test_solver_loop.py walks a
// second time on the CPU —
//     r  = violation of g(x) against [g_lower, g_upper]
//     Js = diag(s) J,  rs = s r,   s_i = 1 / max(1, max_k |J_ik|)
//     (Js^T Js + mu I) dx = -Js^T rs          cg_iters conjugate-gradient iterations, matrix-free
//     x <- clip(x + dx min(1, cap / max|dx|), x_lower, x_upper)
// CTA = instance.  When they fit (config 2: 120 KB of 227 KB), the instance's row-scaled CSR values Js are kept in shared
// memory for the whole step — one pass over global memory, then 2 x cg_iters sparse products on chip (512 threads, one CTA
// per SM); larger patterns (config 4: 436 KB) are re-read from global memory / the L2 in every product (256 threads, several
// CTAs per SM).  All solver vectors live in shared memory.  Both products walk the ONE sparsity pattern of the structure class:
// Js p row by row (thread = row), Js^T u column by column through the pattern's transpose (thread = column) — fixed
// summation orders, no atomics, bit-reproducible.
#include <cuda_runtime.h>

#include <cstdint>

#include "launch.h"

namespace twb {
namespace {

// sum / max over the CTA, result broadcast to every thread (fixed tree: deterministic)
__device__ __forceinline__ double BlockSum(double v, double* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += scratch[w];
  return t;
}
__device__ __forceinline__ double BlockMax(double v, double* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = scratch[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) t = fmax(t, scratch[w]);
  return t;
}

// kSmemJ: the scaled values Js live in shared memory (after the vectors); else they are formed on the fly from global memory.
// Idx: the pattern's index arrays as 16-bit (n, m, nnz < 65 536: 90 KB instead of 180 KB, so that they stay in the L1 beside
// one CTA's shared memory — with 32-bit indices every index load was an L2 round trip: 14.6 instead of 10.5 ms per step) or 32-bit.
// thread = row (Js p) / thread = column (Js^T u): measured faster than 16 lanes per row with shuffle trees (25.4 vs 14.6 ms).
template <bool kSmemJ, class Idx>
__global__ void __launch_bounds__(512) LmStepKernel(LmPattern pat, const Idx* __restrict__ col_idx, const Idx* __restrict__ slot_t,
                                                    const Idx* __restrict__ row_t, int n, int m, int nnz, double* __restrict__ X,
                                                    const double* __restrict__ G, const double* __restrict__ JAC, const double* __restrict__ x_lower,
                                                    const double* __restrict__ x_upper, size_t bound_stride, double mu, double cap, int cg_iters,
                                                    double* __restrict__ violation) {
  extern __shared__ __align__(16) double sm[];
  double* s = sm;             // [m] row scales
  double* u = s + m;          // [m] row vector (rs, then Js p)
  double* dx = u + m;         // [n]
  double* res = dx + n;       // [n]
  double* pd = res + n;       // [n]
  double* ap = pd + n;        // [n]
  double* scratch = ap + n;   // [32]
  double* Js = scratch + 32;  // [nnz] (kSmemJ)
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const double* __restrict__ J = JAC + (size_t)b * nnz;
  const double* __restrict__ g = G + (size_t)b * m;
  // row scales, the scaled values and the scaled violation
  double vmax = 0.0;
  for (int q = tid; q < m; q += nt) {
    const int i = __ldg(pat.row_order + q);   // rows in descending length: the lanes of a warp walk rows of similar length
    const int k0 = __ldg(pat.row_ptr + i), k1 = __ldg(pat.row_ptr + i + 1);
    double big = 0.0;
    for (int k = k0; k < k1; ++k) big = fmax(big, fabs(J[k]));
    const double si = 1.0 / fmax(big, 1.0);
    if (kSmemJ) for (int k = k0; k < k1; ++k) Js[k] = J[k] * si;
    const double gi = g[i], lo = __ldg(pat.g_lower + i), up = __ldg(pat.g_upper + i);
    const double r = gi < lo ? gi - lo : gi > up ? gi - up : 0.0;
    s[i] = si;
    const double rs = si * r;
    u[i] = rs;
    vmax = fmax(vmax, fabs(rs));
  }
  vmax = BlockMax(vmax, scratch);        // (also orders the writes of s, u, Js)
  auto row_dot = [&](int i) {            // (Js p)_i
    double acc = 0.0;
    const int k0 = __ldg(pat.row_ptr + i), k1 = __ldg(pat.row_ptr + i + 1);
    if (kSmemJ) { for (int k = k0; k < k1; ++k) acc += Js[k] * pd[__ldg(col_idx + k)]; }
    else { const double si = s[i]; for (int k = k0; k < k1; ++k) acc += (J[k] * si) * pd[__ldg(col_idx + k)]; }
    return acc;
  };
  auto col_dot = [&](int j) {            // (Js^T u)_j
    double acc = 0.0;
    const int k0 = __ldg(pat.col_ptr + j), k1 = __ldg(pat.col_ptr + j + 1);
    for (int k = k0; k < k1; ++k) {
      const int r = __ldg(row_t + k), slot = __ldg(slot_t + k);
      acc += (kSmemJ ? Js[slot] : J[slot] * s[r]) * u[r];
    }
    return acc;
  };
  // b = -Js^T rs
  double rr = 0.0;
  for (int q = tid; q < n; q += nt) {
    const int j = __ldg(pat.col_order + q);
    const double bj = -col_dot(j);
    dx[j] = 0.0; res[j] = bj; pd[j] = bj;
    rr += bj * bj;
  }
  rr = BlockSum(rr, scratch);
  for (int it = 0; it < cg_iters; ++it) {
    for (int q = tid; q < m; q += nt) { const int i = __ldg(pat.row_order + q); u[i] = row_dot(i); }
    __syncthreads();
    double pap = 0.0;
    for (int q = tid; q < n; q += nt) {   // ap = Js^T (Js p) + mu p
      const int j = __ldg(pat.col_order + q);
      const double a = col_dot(j) + mu * pd[j];
      ap[j] = a;
      pap += pd[j] * a;
    }
    pap = BlockSum(pap, scratch);
    const double alpha = pap > 0.0 ? rr / pap : 0.0;
    double rr_new = 0.0;
    for (int j = tid; j < n; j += nt) {
      dx[j] += alpha * pd[j];
      const double rj = res[j] - alpha * ap[j];
      res[j] = rj;
      rr_new += rj * rj;
    }
    rr_new = BlockSum(rr_new, scratch);
    const double beta = rr > 0.0 ? rr_new / rr : 0.0;
    for (int j = tid; j < n; j += nt) pd[j] = res[j] + beta * pd[j];
    rr = rr_new;
    __syncthreads();
  }
  double big = 0.0;
  for (int j = tid; j < n; j += nt) big = fmax(big, fabs(dx[j]));
  big = BlockMax(big, scratch);
  const double scale = big > cap ? cap / big : 1.0;
  double* x = X + (size_t)b * n;
  const double* lo = x_lower + (size_t)b * bound_stride;
  const double* up = x_upper + (size_t)b * bound_stride;
  for (int j = tid; j < n; j += nt) x[j] = fmin(fmax(x[j] + dx[j] * scale, lo[j]), up[j]);
  if (tid == 0 && violation) violation[b] = vmax;
}

template <bool kSmemJ, class Idx>
int LaunchLm(const LmPattern& pat, const Idx* col_idx, const Idx* slot_t, const Idx* row_t, int threads, size_t smem, int n, int m, int nnz, double* x,
             const double* g, const double* jac, const double* xl, const double* xu, size_t bound_stride, double mu, double cap, int cg_iters,
             double* violation, int nb, cudaStream_t s) {
  cudaError_t e = cudaFuncSetAttribute(LmStepKernel<kSmemJ, Idx>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  LmStepKernel<kSmemJ, Idx><<<nb, threads, smem, s>>>(pat, col_idx, slot_t, row_t, n, m, nnz, x, g, jac, xl, xu, bound_stride, mu, cap, cg_iters, violation);
  return (int)cudaGetLastError();
}

}  // namespace

size_t LmSharedBytes(int n, int m) { return sizeof(double) * ((size_t)2 * m + 4 * (size_t)n + 32); }

int LaunchLmStep(const LmPattern& pat, int n, int m, int nnz, double* x, const double* g, const double* jac, const double* x_lower,
                 const double* x_upper, size_t bound_stride, double mu, double cap, int cg_iters, double* violation, int nb, cudaStream_t s) {
  if (nb <= 0) return 0;
  const size_t vec = LmSharedBytes(n, m), with_j = vec + sizeof(double) * (size_t)nnz;
  const bool smem_j = with_j <= 220 * 1024;   // the scaled values fit beside the vectors: one CTA of 512 threads per SM
  const int threads = smem_j ? 512 : 256;
  const size_t smem = smem_j ? with_j : vec;
  if (pat.col_idx16) {
    return smem_j ? LaunchLm<true>(pat, pat.col_idx16, pat.slot_t16, pat.row_t16, threads, smem, n, m, nnz, x, g, jac, x_lower, x_upper, bound_stride, mu, cap, cg_iters, violation, nb, s)
                  : LaunchLm<false>(pat, pat.col_idx16, pat.slot_t16, pat.row_t16, threads, smem, n, m, nnz, x, g, jac, x_lower, x_upper, bound_stride, mu, cap, cg_iters, violation, nb, s);
  }
  return smem_j ? LaunchLm<true>(pat, pat.col_idx, pat.slot_t, pat.row_t, threads, smem, n, m, nnz, x, g, jac, x_lower, x_upper, bound_stride, mu, cap, cg_iters, violation, nb, s)
                : LaunchLm<false>(pat, pat.col_idx, pat.slot_t, pat.row_t, threads, smem, n, m, nnz, x, g, jac, x_lower, x_upper, bound_stride, mu, cap, cg_iters, violation, nb, s);
}

}  // namespace twb
