// wbench.cu — write-bandwidth microbenchmark for the output layout of the evaluation kernels:
// out[B][row_len] (instance-major), every warp writes, for the 32 instances of a tile, one contiguous
// segment of `seg` doubles of each instance's row.  Explores loop order, store width and cache hints.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/wbench.cu -o /tmp/wbench && /tmp/wbench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int kHint>
__device__ __forceinline__ void St2(double* p, double a, double b) {
  if (kHint == 0) asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
  if (kHint == 1) asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
  if (kHint == 2) asm volatile("st.global.wt.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
  if (kHint == 3) asm volatile("st.global.cg.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}

// contiguous fill: the ceiling
__global__ void FillContig(double* out, size_t n2) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t step = (size_t)gridDim.x * blockDim.x;
  for (; i < n2; i += step) St2<0>(out + 2 * i, 1.0, 2.0);
}

// segments: grid.x = segment index (row_len / seg segments per row), grid.y = tile; warps_per_cta consecutive segments per CTA
// order 0: chunk-outer (for chunk: for instance), order 1: instance-outer (for instance: for chunk)
template <int kHint, int kOrder>
__global__ void FillSeg(double* out, int row_len, int seg, int n_seg, int warps) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s = blockIdx.x * warps + warp;
  if (s >= n_seg) return;
  double* base = out + (size_t)blockIdx.y * 32 * row_len + (size_t)s * seg;
  const int pairs = seg / 2;
  if (kOrder == 0) {
    for (int i = lane; i < pairs; i += 32) {
      double* o = base + 2 * i;
#pragma unroll 8
      for (int j = 0; j < 32; ++j) { St2<kHint>(o, 1.0 + j, 2.0 + i); o += row_len; }
    }
  } else {
    for (int j = 0; j < 32; ++j) {
      double* o = base + (size_t)j * row_len;
      for (int i = lane; i < pairs; i += 32) St2<kHint>(o + 2 * i, 1.0 + j, 2.0 + i);
    }
  }
}

// lane = instance: every lane walks its own row segment with 16-byte stores (no transposition needed)
template <int kHint>
__global__ void FillLaneRow(double* out, int row_len, int seg, int n_seg, int warps) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s = blockIdx.x * warps + warp;
  if (s >= n_seg) return;
  double* o = out + ((size_t)blockIdx.y * 32 + lane) * row_len + (size_t)s * seg;
#pragma unroll 8
  for (int i = 0; i < seg; i += 2) St2<kHint>(o + i, 1.0 + i, 2.0);
}

// tile-row order: CTA = (tile, instance group); its warps write whole rows back to back: warp w writes instance (w) fully (contiguous)
__global__ void FillRows(double* out, int row_len, int B) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  for (int b = blockIdx.x * warps + warp; b < B; b += gridDim.x * warps) {
    double* o = out + (size_t)b * row_len;
    for (int i = lane; i < row_len / 2; i += 32) St2<0>(o + 2 * i, 1.0, 2.0);
  }
}

// the real store phase in miniature: values come from a shared-memory state block (rows picked by a table),
// are multiplied by a coefficient and stored; occupancy is capped by the dynamic shared memory size
struct PairDesc { int off; unsigned short d0, d1; double c0, c1; };
template <int kUnroll>
__global__ void FillFromSmem(double* out, const PairDesc* __restrict__ descs, int row_len, int seg, int n_seg, int warps, int rows) {
  extern __shared__ __align__(16) double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s = blockIdx.x * warps + warp;
  if (s >= n_seg) return;
  double* t = sm + (size_t)warp * rows * 34;
  for (int r = 0; r < rows; ++r) t[r * 34 + lane] = r + lane;
  __syncwarp();
  double* base = out + (size_t)blockIdx.y * 32 * row_len;
  const int pairs = seg / 2;
  const PairDesc* dl = descs + (size_t)s * pairs;
  for (int i = lane; i < pairs; i += 32) {
    const PairDesc pd = dl[i];
    double* o = base + pd.off;
    const double* r0 = t + pd.d0 * 34; const double* r1 = t + pd.d1 * 34;
#pragma unroll kUnroll
    for (int j = 0; j < 32; j += 2) {
      const double2 a = *reinterpret_cast<const double2*>(r0 + j), b = *reinterpret_cast<const double2*>(r1 + j);
      St2<1>(o, a.x * pd.c0, b.x * pd.c1); o += row_len;
      St2<1>(o, a.y * pd.c0, b.y * pd.c1); o += row_len;
    }
  }
}

// variant for isolating what slows the real store phase: `shift` doubles of misalignment (the first/last
// `shift` elements of a segment are written as 8-byte singles, lane = instance), optional g-like side array
// (3 doubles per segment and instance), and a sub-range of segments [seg_lo, seg_hi) of every row.
__device__ int g_tile_mul = 1;   // tile order experiment: tile = (blockIdx.y * g_tile_mul) % gridDim.y
__global__ void FillExp(double* out, double* gout, const PairDesc* __restrict__ descs, int row_len, int seg, int seg_lo, int seg_hi,
                        int warps, int rows, int shift, int g_len) {
  extern __shared__ __align__(16) double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s = seg_lo + blockIdx.x * warps + warp;
  if (s >= seg_hi) return;
  double* t = sm + (size_t)warp * rows * 34;
  for (int r = 0; r < rows; ++r) t[r * 34 + lane] = r + lane;
  __syncwarp();
  const int tile = (int)(((long long)blockIdx.y * g_tile_mul) % gridDim.y);
  double* base = out + (size_t)tile * 32 * row_len;
  const int pairs = seg / 2 - (shift ? 1 : 0);
  const PairDesc* dl = descs + (size_t)s * (seg / 2);
  for (int i = lane; i < pairs; i += 32) {
    const PairDesc pd = dl[i];
    double* o = base + pd.off + 2 * shift;
    const double* r0 = t + pd.d0 * 34; const double* r1 = t + pd.d1 * 34;
#pragma unroll 8
    for (int j = 0; j < 32; j += 2) {
      const double2 a = *reinterpret_cast<const double2*>(r0 + j), b = *reinterpret_cast<const double2*>(r1 + j);
      St2<1>(o, a.x * pd.c0, b.x * pd.c1); o += row_len;
      St2<1>(o, a.y * pd.c0, b.y * pd.c1); o += row_len;
    }
  }
  if (shift) {   // boundary elements as singles: `shift` at the front, 2 - shift at the back
    double* o = base + (size_t)lane * row_len + (size_t)s * seg;
    if (s * seg + seg < row_len) { o[1] = t[34 + lane]; o[seg] = t[5 * 34 + lane]; }
  }
  if (gout) {
    double* go = gout + ((size_t)blockIdx.y * 32 + lane) * g_len + (size_t)s * 3;
    go[0] = t[3 * 34 + lane]; go[1] = t[4 * 34 + lane]; go[2] = t[5 * 34 + lane];
  }
}

// a warp writes `per_warp` segments one after the other (like the feet of a range-of-motion sample), each with its
// own dependent header -> descriptor loads; optional CTA barrier before every segment
__global__ void FillMulti(double* out, const PairDesc* __restrict__ descs, const int* __restrict__ hdr, int row_len, int seg, int n_seg,
                          int warps, int rows, int per_warp, int use_barrier, int hdr_depth) {
  extern __shared__ __align__(16) double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s0 = (blockIdx.x * warps + warp);
  const int groups = n_seg / per_warp;          // segment q of warp-item s0 is s0 + q * groups (far apart, like the feet blocks)
  const bool valid = s0 < groups;
  double* t = sm + (size_t)warp * rows * 34;
  for (int r = 0; r < rows; ++r) t[r * 34 + lane] = r + lane;
  __syncwarp();
  double* base = out + (size_t)blockIdx.y * 32 * row_len;
  const int pairs = seg / 2;
  for (int q = 0; q < per_warp; ++q) {
    if (use_barrier) __syncthreads();
    if (!valid) continue;
    int s = s0 + q * groups;
    for (int h = 0; h < hdr_depth; ++h) s = __ldg(hdr + s);     // dependent header loads (identity table)
    const PairDesc* dl = descs + (size_t)s * pairs;
    for (int i = lane; i < pairs; i += 32) {
      const PairDesc pd = dl[i];
      double* o = base + pd.off;
      const double* r0 = t + pd.d0 * 34; const double* r1 = t + pd.d1 * 34;
#pragma unroll 8
      for (int j = 0; j < 32; j += 2) {
        const double2 a = *reinterpret_cast<const double2*>(r0 + j), b = *reinterpret_cast<const double2*>(r1 + j);
        St2<1>(o, a.x * pd.c0, b.x * pd.c1); o += row_len;
        St2<1>(o, a.y * pd.c0, b.y * pd.c1); o += row_len;
      }
    }
  }
}

template <class F>
float TimeIt(F f, int reps = 10) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); f(); cudaDeviceSynchronize();
  float best = 1e9;
  for (int r = 0; r < reps; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
  return best;
}

int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 4096, row_len = 15096;   // config 2: nnz
  const size_t n = (size_t)B * row_len;
  double* out; CK(cudaMalloc(&out, n * 8));
  const double gb = n * 8 / 1e9;
  float ms = TimeIt([&] { FillContig<<<148 * 8, 256>>>(out, n / 2); });
  printf("contiguous fill (st.v2): %.1f us  %.0f GB/s\n", ms * 1e3, gb / ms * 1e3);
  ms = TimeIt([&] { cudaMemsetAsync(out, 0, n * 8); });
  printf("cudaMemset: %.1f us  %.0f GB/s\n", ms * 1e3, gb / ms * 1e3);
  ms = TimeIt([&] { FillRows<<<148 * 4, 256>>>(out, row_len, B); });
  printf("row-major rows (warp = instance row): %.1f us  %.0f GB/s\n", ms * 1e3, gb / ms * 1e3);
  const int tiles = B / 32;
  for (int seg : {88, 176, 348, 1368, 2516}) {   // doubles per (instance, warp): 88 ~ one RoM foot-sample, 176 ~ one dynamic sample
    const int n_seg = row_len / seg;
    const double frac = (double)n_seg * seg / row_len;
    for (int warps : {4, 8}) {
      dim3 grid((n_seg + warps - 1) / warps, tiles);
      printf("seg %4d doubles, %d warps/CTA:", seg, warps);
      ms = TimeIt([&] { FillSeg<0, 0><<<grid, warps * 32>>>(out, row_len, seg, n_seg, warps); }); printf("  chunk-outer %.0f", gb * frac / ms * 1e3);
      ms = TimeIt([&] { FillSeg<1, 0><<<grid, warps * 32>>>(out, row_len, seg, n_seg, warps); }); printf("  .cs %.0f", gb * frac / ms * 1e3);
      ms = TimeIt([&] { FillSeg<2, 0><<<grid, warps * 32>>>(out, row_len, seg, n_seg, warps); }); printf("  .wt %.0f", gb * frac / ms * 1e3);
      ms = TimeIt([&] { FillSeg<3, 0><<<grid, warps * 32>>>(out, row_len, seg, n_seg, warps); }); printf("  .cg %.0f", gb * frac / ms * 1e3);
      ms = TimeIt([&] { FillSeg<0, 1><<<grid, warps * 32>>>(out, row_len, seg, n_seg, warps); }); printf("  inst-outer %.0f", gb * frac / ms * 1e3);
      ms = TimeIt([&] { FillSeg<1, 1><<<grid, warps * 32>>>(out, row_len, seg, n_seg, warps); }); printf("  inst-outer.cs %.0f", gb * frac / ms * 1e3);
      ms = TimeIt([&] { FillLaneRow<0><<<grid, warps * 32>>>(out, row_len, seg, n_seg, warps); }); printf("  lane=row %.0f", gb * frac / ms * 1e3);
      printf(" GB/s\n");
    }
  }
  // occupancy sweep of the table-driven shared-memory version (seg = 88, 4 warps per CTA, 22 state rows per warp)
  {
    const int seg = 88, n_seg = row_len / seg, warps = 4, rows = 22, pairs = seg / 2;
    std::vector<PairDesc> h((size_t)n_seg * pairs);
    for (int sgm = 0; sgm < n_seg; ++sgm) for (int i = 0; i < pairs; ++i) h[(size_t)sgm * pairs + i] = PairDesc{sgm * seg + 2 * i, (unsigned short)((2 * i) % rows), (unsigned short)((2 * i + 1) % rows), 1.5, 2.5};
    PairDesc* d; CK(cudaMalloc(&d, h.size() * sizeof(PairDesc))); CK(cudaMemcpy(d, h.data(), h.size() * sizeof(PairDesc), cudaMemcpyHostToDevice));
    const double frac = (double)n_seg * seg / row_len;
    dim3 grid((n_seg + warps - 1) / warps, tiles);
    CK(cudaFuncSetAttribute(FillFromSmem<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(FillFromSmem<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(FillFromSmem<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int ctas : {1, 2, 3, 4, 6, 8, 12, 16}) {
      size_t smem = (size_t)(226 * 1024) / ctas - 1024;
      if (smem < (size_t)warps * rows * 34 * 8) smem = (size_t)warps * rows * 34 * 8;
      if (smem > 200 * 1024) smem = 200 * 1024;
      printf("smem-driven, %2d CTAs/SM (%2d warps):", ctas, ctas * warps);
      ms = TimeIt([&] { FillFromSmem<2><<<grid, warps * 32, smem>>>(out, d, row_len, seg, n_seg, warps, rows); }); printf("  unroll2 %.0f", gb * frac / ms * 1e3);
      ms = TimeIt([&] { FillFromSmem<8><<<grid, warps * 32, smem>>>(out, d, row_len, seg, n_seg, warps, rows); }); printf("  unroll8 %.0f", gb * frac / ms * 1e3);
      ms = TimeIt([&] { FillFromSmem<16><<<grid, warps * 32, smem>>>(out, d, row_len, seg, n_seg, warps, rows); }); printf("  unroll16 %.0f GB/s\n", gb * frac / ms * 1e3);
    }
  }
  {
    const int seg = 88, n_seg = row_len / seg, warps = 4, rows = 22, pairs = seg / 2;
    std::vector<PairDesc> h((size_t)n_seg * pairs);
    for (int sgm = 0; sgm < n_seg; ++sgm) for (int i = 0; i < pairs; ++i) h[(size_t)sgm * pairs + i] = PairDesc{sgm * seg + 2 * i, (unsigned short)((2 * i) % rows), (unsigned short)((2 * i + 1) % rows), 1.5, 2.5};
    PairDesc* d; CK(cudaMalloc(&d, h.size() * sizeof(PairDesc))); CK(cudaMemcpy(d, h.data(), h.size() * sizeof(PairDesc), cudaMemcpyHostToDevice));
    const int g_len = n_seg * 3;
    double* gout; CK(cudaMalloc(&gout, (size_t)B * g_len * 8));
    CK(cudaFuncSetAttribute(FillExp, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const size_t smem = (size_t)(226 * 1024) / 4 - 1024;   // 4 CTAs = 16 warps per SM
    auto run = [&](const char* label, int lo, int hi, int shift, bool with_g) {
      dim3 grid((hi - lo + warps - 1) / warps, tiles);
      const double bytes = (double)B * ((hi - lo) * seg + (with_g ? (hi - lo) * 3 : 0)) * 8 / 1e9;
      float t = TimeIt([&] { FillExp<<<grid, warps * 32, smem>>>(out, with_g ? gout : nullptr, d, row_len, seg, lo, hi, warps, rows, shift, g_len); });
      printf("%-58s %.1f us  %.0f GB/s\n", label, t * 1e3, bytes / t * 1e3);
    };
    {
      std::vector<int> hh(n_seg); for (int i = 0; i < n_seg; ++i) hh[i] = i;
      int* dh; CK(cudaMalloc(&dh, n_seg * 4)); CK(cudaMemcpy(dh, hh.data(), n_seg * 4, cudaMemcpyHostToDevice));
      CK(cudaFuncSetAttribute(FillMulti, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      for (int per_warp : {1, 4}) for (int depth : {0, 2, 6}) for (int bar : {0, 1}) for (int w : {5, 7}) {
        const int groups = n_seg / per_warp;
        dim3 grid((groups + w - 1) / w, tiles);
        const size_t smem2 = (size_t)(226 * 1024) / 2 - 1024;   // 2 CTAs per SM
        float t = TimeIt([&] { FillMulti<<<grid, w * 32, smem2>>>(out, d, dh, row_len, seg, n_seg, w, rows, per_warp, bar, depth); });
        printf("multi: %d segments/warp, %d dependent header loads, barrier %d, %d warps x 2 CTAs: %.1f us  %.0f GB/s\n", per_warp, depth, bar, w, t * 1e3,
               (double)B * groups * per_warp * seg * 8 / 1e9 / t * 1e3);
      }
    }
    run("all segments, aligned", 0, n_seg, 0, false);
    run("all segments, aligned + g side array", 0, n_seg, 0, true);
    run("all segments, shift 1 (8-byte singles at both ends)", 0, n_seg, 1, false);
    run("all segments, shift 1 + g", 0, n_seg, 1, true);
    run("middle 60% of every row, aligned", n_seg / 5, n_seg / 5 + (n_seg * 3) / 5, 0, false);
    run("middle 60% of every row, shift 1 + g", n_seg / 5, n_seg / 5 + (n_seg * 3) / 5, 1, true);
    run("first 25% of every row, aligned", 0, n_seg / 4, 0, false);
    for (int mul : {37, 63, 127}) {
      CK(cudaMemcpyToSymbol(g_tile_mul, &mul, sizeof(int)));
      char label[96]; snprintf(label, sizeof label, "first 25%%, tile order x%d mod tiles", mul);
      run(label, 0, n_seg / 4, 0, false);
      snprintf(label, sizeof label, "all segments, tile order x%d mod tiles", mul);
      run(label, 0, n_seg, 0, false);
    }
    { int one = 1; CK(cudaMemcpyToSymbol(g_tile_mul, &one, sizeof(int))); }
    // grid transposed: segment-major (all tiles of one segment group first) instead of tile-major

    {   // three kernels on three streams, each covering its own column range (the current pipeline's shape)
      cudaStream_t st[3]; for (auto& x : st) cudaStreamCreate(&x);
      cudaEvent_t e0, e1[3]; cudaEventCreate(&e0); for (auto& x : e1) cudaEventCreate(&x);
      const int cut1 = n_seg / 4, cut2 = n_seg / 4 + (n_seg * 3) / 5;
      const int lo[3] = {0, cut1, cut2}, hi[3] = {cut1, cut2, n_seg};
      const size_t smem3 = (size_t)(226 * 1024) / 6 - 1024;
      auto go = [&] {
        cudaEventRecord(e0, 0);
        for (int q = 0; q < 3; ++q) {
          cudaStreamWaitEvent(st[q], e0, 0);
          dim3 grid((hi[q] - lo[q] + warps - 1) / warps, tiles);
          FillExp<<<grid, warps * 32, smem3, st[q]>>>(out, nullptr, d, row_len, seg, lo[q], hi[q], warps, rows, 0, g_len);
          cudaEventRecord(e1[q], st[q]); cudaStreamWaitEvent(0, e1[q], 0);
        }
      };
      float t = TimeIt(go);
      printf("%-58s %.1f us  %.0f GB/s\n", "three column ranges on three streams (25% | 60% | 15%)", t * 1e3, (double)B * n_seg * seg * 8 / 1e9 / t * 1e3);
    }
  }
  return 0;
}
