// wbench_tma.cu — write-bandwidth microbenchmark, TMA edition: the same output layout as wbench.cu
// (out[B][row_len], a CTA owns one contiguous segment of `seg` doubles of each of the 32 instances of a tile), but the
// values are assembled in shared memory in OUTPUT ORDER and leave the SM as one `cp.async.bulk.global.shared::cta` per
// (instance, segment).  Every warp owns 32 / W instances of the CTA's tile and walks the CTA's whole pair list for them.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/wbench_tma.cu -o /tmp/wbench_tma && /tmp/wbench_tma
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

struct PairDesc { int off; unsigned short d0, d1; double c0, c1; };

__device__ __forceinline__ unsigned SmemAddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void BulkStore(void* gdst, const void* ssrc, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(SmemAddr(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void BulkStoreHint(void* gdst, const void* ssrc, unsigned bytes, unsigned long long pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst), "r"(SmemAddr(ssrc)), "r"(bytes), "l"(pol) : "memory");
}
__device__ __forceinline__ void BulkCommit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void BulkWaitRead() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void FenceAsync() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// kMode 0: staging filled with constants (pure TMA store pattern); 1: table-driven expansion from the state block
// kG: instances per fill step (1: 8-byte state reads, 2: 16-byte); kBuf: ring depth of staging buffers per warp
// kHint: 1 = evict-first L2 policy on the bulk stores
template <int kMode, int kG, int kBuf, int kHint>
__global__ void FillTma(double* out, const PairDesc* __restrict__ descs, int row_len, int seg, int seg_lo, int seg_hi, int rows) {
  extern __shared__ __align__(128) double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
  const int s = seg_lo + blockIdx.x;
  if (s >= seg_hi) return;
  double* t = sm;                                            // CTA state block: rows x 34
  const int seg_pad = (seg + 1) & ~1;
  double* stage = sm + (((size_t)rows * 34 + 15) & ~(size_t)15) + (size_t)warp * kBuf * kG * seg_pad;
  for (int r = warp; r < rows; r += W) t[r * 34 + lane] = r + lane;
  __syncthreads();
  unsigned long long pol = 0;
  if (kHint) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  const int per = 32 / W, j0 = warp * per;
  double* base = out + (size_t)blockIdx.y * 32 * row_len + (size_t)s * seg;
  const int pairs = seg / 2;
  const PairDesc* dl = descs + (size_t)s * pairs;
  int it = 0;
  for (int j = j0; j < j0 + per; j += kG, ++it) {
    double* buf = stage + (size_t)(it % kBuf) * kG * seg_pad;
    if (it >= kBuf) { if (lane == 0) BulkWaitRead<kBuf - 1>(); __syncwarp(); }
    for (int i = lane; i < pairs; i += 32) {
      if (kMode == 0) {
#pragma unroll
        for (int g = 0; g < kG; ++g) *reinterpret_cast<double2*>(buf + g * seg_pad + 2 * i) = make_double2(1.0 + j, 2.0 + i);
      } else {
        const PairDesc pd = dl[i];
        const int o = pd.off - s * seg;
        if (kG == 4) {
          const double2 a = *reinterpret_cast<const double2*>(t + pd.d0 * 34 + j), b = *reinterpret_cast<const double2*>(t + pd.d1 * 34 + j);
          const double2 a2 = *reinterpret_cast<const double2*>(t + pd.d0 * 34 + j + 2), b2 = *reinterpret_cast<const double2*>(t + pd.d1 * 34 + j + 2);
          *reinterpret_cast<double2*>(buf + o) = make_double2(a.x * pd.c0, b.x * pd.c1);
          *reinterpret_cast<double2*>(buf + seg_pad + o) = make_double2(a.y * pd.c0, b.y * pd.c1);
          *reinterpret_cast<double2*>(buf + 2 * seg_pad + o) = make_double2(a2.x * pd.c0, b2.x * pd.c1);
          *reinterpret_cast<double2*>(buf + 3 * seg_pad + o) = make_double2(a2.y * pd.c0, b2.y * pd.c1);
        } else if (kG == 2) {
          const double2 a = *reinterpret_cast<const double2*>(t + pd.d0 * 34 + j), b = *reinterpret_cast<const double2*>(t + pd.d1 * 34 + j);
          *reinterpret_cast<double2*>(buf + o) = make_double2(a.x * pd.c0, b.x * pd.c1);
          *reinterpret_cast<double2*>(buf + seg_pad + o) = make_double2(a.y * pd.c0, b.y * pd.c1);
        } else {
#pragma unroll
          for (int g = 0; g < kG; ++g)
            *reinterpret_cast<double2*>(buf + g * seg_pad + o) = make_double2(t[pd.d0 * 34 + j + g] * pd.c0, t[pd.d1 * 34 + j + g] * pd.c1);
        }
      }
    }
    FenceAsync();
    __syncwarp();
    if (lane == 0) {
#pragma unroll
      for (int g = 0; g < kG; ++g) {
        if (kHint) BulkStoreHint(base + (size_t)(j + g) * row_len, buf + g * seg_pad, (unsigned)seg * 8u, pol);
        else BulkStore(base + (size_t)(j + g) * row_len, buf + g * seg_pad, (unsigned)seg * 8u);
      }
      BulkCommit();
    }
  }
  if (lane == 0) BulkWaitRead<0>();
  __syncwarp();
}

// CTA-level staging: all 32 instances of the segment are assembled by the whole CTA (thread = pair, loop over instances,
// the descriptor stays in registers), then one warp issues the 32 bulk stores.
template <int kMode>
__global__ void FillTmaCta(double* out, const PairDesc* __restrict__ descs, int row_len, int seg, int seg_lo, int seg_hi, int rows, int per_cta) {
  extern __shared__ __align__(128) double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* t = sm;
  const int seg_pad = (seg + 1) & ~1;
  double* stage = sm + (((size_t)rows * 34 + 15) & ~(size_t)15);
  for (int r = warp; r < rows; r += (blockDim.x >> 5)) t[r * 34 + lane] = r + lane;
  for (int q = 0; q < per_cta; ++q) {           // per_cta consecutive segments, one after the other (like the feet)
    const int s = seg_lo + blockIdx.x * per_cta + q;
    if (s >= seg_hi) break;
    if (q > 0 && warp == 0) { if (lane == 0) BulkWaitRead<0>(); }
    __syncthreads();
    const int pairs = seg / 2;
    const PairDesc* dl = descs + (size_t)s * pairs;
    for (int i = threadIdx.x; i < pairs; i += blockDim.x) {
      if (kMode == 0) {
        for (int j = 0; j < 32; ++j) *reinterpret_cast<double2*>(stage + j * seg_pad + 2 * i) = make_double2(1.0 + j, 2.0 + i);
      } else {
        const PairDesc pd = dl[i];
        const int o = pd.off - s * seg;
        const double* r0 = t + pd.d0 * 34; const double* r1 = t + pd.d1 * 34;
#pragma unroll 8
        for (int j = 0; j < 32; j += 2) {
          const double2 a = *reinterpret_cast<const double2*>(r0 + j), b = *reinterpret_cast<const double2*>(r1 + j);
          *reinterpret_cast<double2*>(stage + j * seg_pad + o) = make_double2(a.x * pd.c0, b.x * pd.c1);
          *reinterpret_cast<double2*>(stage + (j + 1) * seg_pad + o) = make_double2(a.y * pd.c0, b.y * pd.c1);
        }
      }
    }
    FenceAsync();
    __syncthreads();
    if (warp == 0) {
      double* base = out + (size_t)blockIdx.y * 32 * row_len + (size_t)s * seg;
      BulkStore(base + (size_t)lane * row_len, stage + lane * seg_pad, (unsigned)seg * 8u);
      BulkCommit();
    }
  }
  if (warp == 0) BulkWaitRead<0>();
}

// reference: the round-1 store loop (thread = pair over the CTA, 16-byte st.global.cs)
__global__ void FillStg(double* out, const PairDesc* __restrict__ descs, int row_len, int seg, int seg_lo, int seg_hi, int rows) {
  extern __shared__ __align__(128) double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s = seg_lo + blockIdx.x;
  if (s >= seg_hi) return;
  double* t = sm;
  for (int r = warp; r < rows; r += (blockDim.x >> 5)) t[r * 34 + lane] = r + lane;
  __syncthreads();
  double* base = out + (size_t)blockIdx.y * 32 * row_len;
  const int pairs = seg / 2;
  const PairDesc* dl = descs + (size_t)s * pairs;
  for (int i = threadIdx.x; i < pairs; i += blockDim.x) {
    const PairDesc pd = dl[i];
    double* o = base + pd.off;
    const double* r0 = t + pd.d0 * 34; const double* r1 = t + pd.d1 * 34;
#pragma unroll 8
    for (int j = 0; j < 32; j += 2) {
      const double2 a = *reinterpret_cast<const double2*>(r0 + j), b = *reinterpret_cast<const double2*>(r1 + j);
      asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(o), "d"(a.x * pd.c0), "d"(b.x * pd.c1) : "memory"); o += row_len;
      asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(o), "d"(a.y * pd.c0), "d"(b.y * pd.c1) : "memory"); o += row_len;
    }
  }
}


// CTA-level groups: the CTA assembles kGI instances at a time (thread = (pair, sub-group of 2 instances) item, descriptors
// re-read per item from L1), double-buffered; thread 0 issues the group's bulk stores.
template <int kGI>
__global__ void FillTmaGroup(double* out, const PairDesc* __restrict__ descs, int row_len, int seg, int seg_lo, int seg_hi, int rows) {
  extern __shared__ __align__(128) double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s = seg_lo + blockIdx.x;
  if (s >= seg_hi) return;
  double* t = sm;
  const int seg_pad = (seg + 1) & ~1;
  double* stage = sm + (((size_t)rows * 34 + 15) & ~(size_t)15);
  for (int r = warp; r < rows; r += (blockDim.x >> 5)) t[r * 34 + lane] = r + lane;
  __syncthreads();
  const int pairs = seg / 2;
  const PairDesc* dl = descs + (size_t)s * pairs;
  double* base = out + (size_t)blockIdx.y * 32 * row_len + (size_t)s * seg;
  constexpr int kSub = kGI / 2;                 // 2-instance steps per group
  const int n_items = pairs * kSub;
  for (int grp = 0; grp < 32 / kGI; ++grp) {
    double* buf = stage + (size_t)(grp & 1) * kGI * seg_pad;
    const int jg = grp * kGI;
    for (int it = threadIdx.x; it < n_items; it += blockDim.x) {
      const int sub = it / pairs, i = it - sub * pairs;      // consecutive threads = consecutive pairs of one instance pair
      const PairDesc pd = dl[i];
      const int o = pd.off - s * seg, j = jg + 2 * sub;
      const double2 a = *reinterpret_cast<const double2*>(t + pd.d0 * 34 + j), b = *reinterpret_cast<const double2*>(t + pd.d1 * 34 + j);
      *reinterpret_cast<double2*>(buf + (size_t)(2 * sub) * seg_pad + o) = make_double2(a.x * pd.c0, b.x * pd.c1);
      *reinterpret_cast<double2*>(buf + (size_t)(2 * sub + 1) * seg_pad + o) = make_double2(a.y * pd.c0, b.y * pd.c1);
    }
    FenceAsync();
    if (threadIdx.x == 0) BulkWaitRead<0>();   // the previous group (other buffer) has been read: it may be refilled after this barrier
    __syncthreads();
    if (warp == 0 && lane < kGI) { BulkStore(base + (size_t)(jg + lane) * row_len, buf + (size_t)lane * seg_pad, (unsigned)seg * 8u); }
    if (threadIdx.x == 0) BulkCommit();
  }
  if (threadIdx.x == 0) BulkWaitRead<0>();
}

// 256-bit stores (sm_100: st.global.v4.f64, SASS STG.E.256): thread = whole 32-byte sector (two consecutive pairs of the
// list), a warp instruction covers 1 KB of one instance's row
__global__ void FillContigV4(double* out, size_t n4) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t step = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += step) asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(out + 4 * i), "d"(1.0), "d"(2.0), "d"(3.0), "d"(4.0) : "memory");
}
__global__ void FillContigV2(double* out, size_t n2) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t step = (size_t)gridDim.x * blockDim.x;
  for (; i < n2; i += step) asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(out + 2 * i), "d"(1.0), "d"(2.0) : "memory");
}
template <int kCs>
__global__ void FillStgV4(double* out, const PairDesc* __restrict__ descs, int row_len, int seg, int seg_lo, int seg_hi, int rows) {
  extern __shared__ __align__(128) double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s = seg_lo + blockIdx.x;
  if (s >= seg_hi) return;
  double* t = sm;
  for (int r = warp; r < rows; r += (blockDim.x >> 5)) t[r * 34 + lane] = r + lane;
  __syncthreads();
  double* base = out + (size_t)blockIdx.y * 32 * row_len;
  const int sectors = seg / 4;                     // (the benchmark's segments start on sectors when seg % 4 == 0; else the tail pair is skipped)
  const PairDesc* dl = descs + (size_t)s * (seg / 2);
  const int n_items = sectors * 4;                 // (sector, quarter of the tile's instances)
  for (int it = threadIdx.x; it < n_items; it += blockDim.x) {
    const int q = it / sectors, k = it - q * sectors;
    const PairDesc p0 = dl[2 * k], p1 = dl[2 * k + 1];
    double* o = base + p0.off + (size_t)(8 * q) * row_len;
    const double* r0 = t + p0.d0 * 34 + 8 * q; const double* r1 = t + p0.d1 * 34 + 8 * q;
    const double* r2 = t + p1.d0 * 34 + 8 * q; const double* r3 = t + p1.d1 * 34 + 8 * q;
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      const double2 a = *reinterpret_cast<const double2*>(r0 + j), b = *reinterpret_cast<const double2*>(r1 + j);
      const double2 c = *reinterpret_cast<const double2*>(r2 + j), d = *reinterpret_cast<const double2*>(r3 + j);
      if (kCs) asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(o), "d"(a.x * p0.c0), "d"(b.x * p0.c1), "d"(c.x * p1.c0), "d"(d.x * p1.c1) : "memory");
      else asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(o), "d"(a.x * p0.c0), "d"(b.x * p0.c1), "d"(c.x * p1.c0), "d"(d.x * p1.c1) : "memory");
      o += row_len;
      if (kCs) asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(o), "d"(a.y * p0.c0), "d"(b.y * p0.c1), "d"(c.y * p1.c0), "d"(d.y * p1.c1) : "memory");
      else asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(o), "d"(a.y * p0.c0), "d"(b.y * p0.c1), "d"(c.y * p1.c0), "d"(d.y * p1.c1) : "memory");
      o += row_len;
    }
  }
}

template <class F>
float TimeIt(F f, int reps = 10) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); f(); cudaDeviceSynchronize();
  float best = 1e9;
  for (int r = 0; r < reps; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
  CK(cudaGetLastError());
  return best;
}

int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 4096, row_len = 15096;
  const size_t n = (size_t)B * row_len;
  double* out; CK(cudaMalloc(&out, n * 8));
  const int tiles = B / 32, rows = 34 * 4, W = 4;
  const bool realistic = argc > 2 ? atoi(argv[2]) != 0 : true;
  ({ float ms = TimeIt([&] { cudaMemsetAsync(out, 0, n * 8); }); printf("cudaMemset: %.1f us  %.0f GB/s\n", ms * 1e3, n * 8 / 1e9 / ms * 1e3); });
  ({ float ms = TimeIt([&] { FillContigV2<<<148 * 8, 256>>>(out, n / 2); }); printf("contiguous fill st.cs.v2.f64: %.1f us  %.0f GB/s\n", ms * 1e3, n * 8 / 1e9 / ms * 1e3); });
  ({ float ms = TimeIt([&] { FillContigV4<<<148 * 8, 256>>>(out, n / 4); }); printf("contiguous fill st.v4.f64 (256-bit): %.1f us  %.0f GB/s\n", ms * 1e3, n * 8 / 1e9 / ms * 1e3); });
  for (int seg : {88, 176, 336, 728, 1368}) {
    const int n_seg = row_len / seg, pairs = seg / 2;
    std::vector<PairDesc> h((size_t)n_seg * pairs);
    for (int sgm = 0; sgm < n_seg; ++sgm) for (int i = 0; i < pairs; ++i)
      {
        // state rows like a range-of-motion row: blocks of 12 entries = 3 state rows x 4 basis values, rows advance block by block
        auto drow = [&](int e) { return (unsigned short)(1 + 3 * ((e / 12) % 11) + (e % 3)); };
        h[(size_t)sgm * pairs + i] = realistic ? PairDesc{sgm * seg + 2 * i, drow(2 * i), drow(2 * i + 1), 1.5, 2.5}
                                               : PairDesc{sgm * seg + 2 * i, (unsigned short)((2 * i) % rows), (unsigned short)((2 * i + 1) % rows), 1.5, 2.5};
      }
    PairDesc* d; CK(cudaMalloc(&d, h.size() * sizeof(PairDesc))); CK(cudaMemcpy(d, h.data(), h.size() * sizeof(PairDesc), cudaMemcpyHostToDevice));
    const size_t state_bytes = (((size_t)rows * 34 + 15) & ~(size_t)15) * 8;
    struct Range { const char* name; int lo, hi; };
    const Range ranges[] = {{"all", 0, n_seg}, {"mid60", n_seg / 5, n_seg / 5 + (n_seg * 3) / 5}, {"first25", 0, n_seg / 4}};
    for (const Range& rg : ranges) {
      const double gb = (double)B * (rg.hi - rg.lo) * seg * 8 / 1e9;
      dim3 grid(rg.hi - rg.lo, tiles);
      printf("seg %4d (%5d B) %-7s:", seg, seg * 8, rg.name);
      {
        CK(cudaFuncSetAttribute(FillStg, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        float ms = TimeIt([&] { FillStg<<<grid, W * 32, state_bytes>>>(out, d, row_len, seg, rg.lo, rg.hi, rows); });
        printf("  stg %.0f", gb / ms * 1e3);
        CK(cudaFuncSetAttribute(FillStgV4<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CK(cudaFuncSetAttribute(FillStgV4<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        ms = TimeIt([&] { FillStgV4<0><<<grid, W * 32, state_bytes>>>(out, d, row_len, seg, rg.lo, rg.hi, rows); });
        printf("  stg256 %.0f", gb / ms * 1e3);
        ms = TimeIt([&] { FillStgV4<1><<<grid, W * 32, state_bytes>>>(out, d, row_len, seg, rg.lo, rg.hi, rows); });
        printf("  stg256.cs %.0f", gb / ms * 1e3);
      }
#define RUN(MODE, G, NB, HINT, label)                                                                                         \
  {                                                                                                                          \
    const size_t smem = state_bytes + (size_t)W * NB * G * ((seg + 1) & ~1) * 8;                                             \
    if (smem <= 220 * 1024) {                                                                                                \
      CK(cudaFuncSetAttribute(FillTma<MODE, G, NB, HINT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));           \
      float ms = TimeIt([&] { FillTma<MODE, G, NB, HINT><<<grid, W * 32, smem>>>(out, d, row_len, seg, rg.lo, rg.hi, rows); }); \
      printf("  " label " %.0f", gb / ms * 1e3);                                                                             \
    }                                                                                                                        \
  }
      RUN(0, 1, 2, 0, "pure.g1b2") RUN(0, 2, 2, 0, "pure.g2b2")
      RUN(1, 1, 3, 0, "exp.g1b3") RUN(1, 2, 1, 0, "exp.g2b1") RUN(1, 2, 2, 0, "exp.g2b2") RUN(1, 4, 1, 0, "exp.g4b1") RUN(1, 4, 2, 0, "exp.g4b2")
#define RUNG(GI, WW, label)                                                                                                    \
  {                                                                                                                          \
    const size_t smem = state_bytes + (size_t)2 * GI * ((seg + 1) & ~1) * 8;                                                  \
    if (smem <= 220 * 1024) {                                                                                                \
      CK(cudaFuncSetAttribute(FillTmaGroup<GI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                     \
      float ms = TimeIt([&] { FillTmaGroup<GI><<<grid, WW * 32, smem>>>(out, d, row_len, seg, rg.lo, rg.hi, rows); });        \
      printf("  " label " %.0f", gb / ms * 1e3);                                                                             \
    }                                                                                                                        \
  }
      RUNG(4, 4, "grp4") RUNG(8, 4, "grp8") RUNG(16, 4, "grp16") RUNG(8, 8, "grp8.8w")
      printf(" GB/s\n");
    }
    cudaFree(d);
  }
  return 0;
}
