// pcie_probe.cu — what the host link of this box can do: pinned cudaMemcpyAsync bandwidth, device -> host and
// host -> device, alone and both directions at once, on 1 .. N GPUs concurrently (one host thread per GPU).  The
// end-to-end number of bench.py (`e2e`) is quoted against the D2H figure measured here (`e2e.ceiling`).
//   nvcc -O3 scripts/pcie_probe.cu -o build/pcie_probe && build/pcie_probe [max_gpus] [MiB]
#include <cuda_runtime.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

struct Result { double d2h = 0, h2d = 0, both_d2h = 0, both_h2d = 0; };

static void Worker(int dev, size_t bytes, int reps, std::atomic<int>* gate, int n_workers, Result* out) {
  CK(cudaSetDevice(dev));
  void *h0, *h1, *d0, *d1;
  CK(cudaMallocHost(&h0, bytes)); CK(cudaMallocHost(&h1, bytes)); CK(cudaMalloc(&d0, bytes)); CK(cudaMalloc(&d1, bytes));
  cudaStream_t s0, s1; CK(cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
  cudaEvent_t a, b, c, d; cudaEventCreate(&a); cudaEventCreate(&b); cudaEventCreate(&c); cudaEventCreate(&d);
  auto sync_all = [&] { gate->fetch_add(1); while (gate->load() % n_workers != 0) std::this_thread::yield(); };
  CK(cudaMemcpyAsync(h0, d0, bytes, cudaMemcpyDeviceToHost, s0)); CK(cudaMemcpyAsync(d1, h1, bytes, cudaMemcpyHostToDevice, s1));
  CK(cudaDeviceSynchronize());
  float ms;
  sync_all();
  cudaEventRecord(a, s0); for (int r = 0; r < reps; ++r) CK(cudaMemcpyAsync(h0, d0, bytes, cudaMemcpyDeviceToHost, s0)); cudaEventRecord(b, s0);
  CK(cudaStreamSynchronize(s0)); cudaEventElapsedTime(&ms, a, b); out->d2h = bytes * (double)reps / ms / 1e6;
  sync_all();
  cudaEventRecord(a, s1); for (int r = 0; r < reps; ++r) CK(cudaMemcpyAsync(d1, h1, bytes, cudaMemcpyHostToDevice, s1)); cudaEventRecord(b, s1);
  CK(cudaStreamSynchronize(s1)); cudaEventElapsedTime(&ms, a, b); out->h2d = bytes * (double)reps / ms / 1e6;
  sync_all();
  cudaEventRecord(a, s0); cudaEventRecord(c, s1);
  for (int r = 0; r < reps; ++r) { CK(cudaMemcpyAsync(h0, d0, bytes, cudaMemcpyDeviceToHost, s0)); CK(cudaMemcpyAsync(d1, h1, bytes, cudaMemcpyHostToDevice, s1)); }
  cudaEventRecord(b, s0); cudaEventRecord(d, s1);
  CK(cudaDeviceSynchronize());
  cudaEventElapsedTime(&ms, a, b); out->both_d2h = bytes * (double)reps / ms / 1e6;
  cudaEventElapsedTime(&ms, c, d); out->both_h2d = bytes * (double)reps / ms / 1e6;
  sync_all();
  cudaFreeHost(h0); cudaFreeHost(h1); cudaFree(d0); cudaFree(d1);
}

int main(int argc, char** argv) {
  int n_dev = 0; CK(cudaGetDeviceCount(&n_dev));
  const int max_gpus = argc > 1 ? std::min(atoi(argv[1]), n_dev) : n_dev;
  const size_t bytes = (size_t)(argc > 2 ? atoi(argv[2]) : 512) << 20;
  printf("{\"probe\": \"pcie\", \"bytes\": %zu, \"results\": [", bytes);
  bool first = true;
  for (int n = 1; n <= max_gpus; n *= 2) {
    std::vector<Result> res(n); std::vector<std::thread> th; std::atomic<int> gate{0};
    for (int g = 0; g < n; ++g) th.emplace_back(Worker, g, bytes, 4, &gate, n, &res[g]);
    for (auto& t : th) t.join();
    Result sum, mn{1e30, 1e30, 1e30, 1e30};
    for (auto& r : res) {
      sum.d2h += r.d2h; sum.h2d += r.h2d; sum.both_d2h += r.both_d2h; sum.both_h2d += r.both_h2d;
      mn.d2h = std::min(mn.d2h, r.d2h); mn.h2d = std::min(mn.h2d, r.h2d);
    }
    printf("%s{\"gpus\": %d, \"d2h_gbs_total\": %.1f, \"h2d_gbs_total\": %.1f, \"d2h_gbs_min_rank\": %.1f, \"h2d_gbs_min_rank\": %.1f, "
           "\"bidir_d2h_gbs_total\": %.1f, \"bidir_h2d_gbs_total\": %.1f}", first ? "" : ", ", n, sum.d2h, sum.h2d, mn.d2h, mn.h2d, sum.both_d2h, sum.both_h2d);
    first = false;
  }
  printf("]}\n");
  return 0;
}
