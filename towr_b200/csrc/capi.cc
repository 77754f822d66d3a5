// capi.cc — extern "C" boundary (include/towr_b200.h).  Owns device memory for
// the structure-class tables and the per-batch staging buffers.  There is no
// CPU evaluation path: without a CUDA device twb_batch_create fails.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/towr_b200.h"
#include "formulation.h"
#include "launch.h"

namespace {
thread_local std::string g_err;
int Fail(int code, const std::string& why) { g_err = why; return code; }
int CudaFail(cudaError_t e, const char* what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return TWB_ERR_CUDA;
}
template <class T>
cudaError_t Upload(const std::vector<T>& h, const T** d, std::vector<void*>* owned) {
  *d = nullptr;
  if (h.empty()) return cudaSuccess;
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, h.size() * sizeof(T));
  if (e != cudaSuccess) return e;
  owned->push_back(p);
  e = cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
  *d = static_cast<const T*>(p);
  return e;
}
}  // namespace

struct twb_problem {
  twb::Formulation f;
};

struct twb_batch {
  const twb_problem* prob = nullptr;
  int B = 0, device = 0, G = 2;
  twb::Plan plan{};
  std::vector<void*> owned;       // device allocations of the tables
  int* d_terrain = nullptr;       // per-instance terrain ids (optional)
  // staging for the host-pointer variant
  double *d_x = nullptr, *d_g = nullptr, *d_jac = nullptr, *d_cost = nullptr, *d_grad = nullptr;
  int* d_status = nullptr;
  cudaStream_t stream = nullptr;
};

extern "C" {

const char* twb_last_error(void) { return g_err.c_str(); }
const char* twb_version(void) { return "towr_b200 0.1 (sm_100a, fp64)"; }

int twb_problem_create(const twb_spec* spec, twb_problem** out) {
  if (!spec || !out) return Fail(TWB_ERR_INVALID, "null argument");
  *out = nullptr;
  twb_problem* p = new (std::nothrow) twb_problem();
  if (!p) return Fail(TWB_ERR_INVALID, "out of memory");
  std::string err;
  int rc = TWB_OK;
  try { rc = p->f.Build(*spec, &err); } catch (const std::exception& e) { rc = TWB_ERR_INVALID; err = e.what(); }
  if (rc != TWB_OK) { delete p; return Fail(rc, err); }
  *out = p;
  return TWB_OK;
}
void twb_problem_destroy(twb_problem* p) { delete p; }

int twb_problem_dims(const twb_problem* p, int* n, int* m, int* nnz) {
  if (!p) return Fail(TWB_ERR_INVALID, "null problem");
  if (n) *n = p->f.n;
  if (m) *m = p->f.m;
  if (nnz) *nnz = p->f.nnz;
  return TWB_OK;
}
int twb_problem_structure(const twb_problem* p, int* iRow, int* jCol) {
  if (!p || !iRow || !jCol) return Fail(TWB_ERR_INVALID, "null argument");
  for (int r = 0; r < p->f.m; ++r)
    for (int k = p->f.row_ptr[r]; k < p->f.row_ptr[r + 1]; ++k) { iRow[k] = r; jCol[k] = p->f.col_idx[k]; }
  return TWB_OK;
}
int twb_problem_row_ptr(const twb_problem* p, int* row_ptr) {
  if (!p || !row_ptr) return Fail(TWB_ERR_INVALID, "null argument");
  std::memcpy(row_ptr, p->f.row_ptr.data(), sizeof(int) * (p->f.m + 1));
  return TWB_OK;
}
int twb_problem_bounds(const twb_problem* p, double* xl, double* xu, double* gl, double* gu) {
  if (!p) return Fail(TWB_ERR_INVALID, "null problem");
  if (xl) std::memcpy(xl, p->f.x_lower.data(), sizeof(double) * p->f.n);
  if (xu) std::memcpy(xu, p->f.x_upper.data(), sizeof(double) * p->f.n);
  if (gl) std::memcpy(gl, p->f.g_lower.data(), sizeof(double) * p->f.m);
  if (gu) std::memcpy(gu, p->f.g_upper.data(), sizeof(double) * p->f.m);
  return TWB_OK;
}
int twb_problem_x0(const twb_problem* p, double* x0) {
  if (!p || !x0) return Fail(TWB_ERR_INVALID, "null argument");
  std::memcpy(x0, p->f.x0.data(), sizeof(double) * p->f.n);
  return TWB_OK;
}
int twb_layout_num_variable_sets(const twb_problem* p) { return p ? (int)p->f.var_sets.size() : 0; }
int twb_layout_num_constraint_sets(const twb_problem* p) { return p ? (int)p->f.con_sets.size() : 0; }
static int CopyComponent(const std::vector<twb::Component>& v, int i, char* name, int cap, int* start, int* count) {
  if (i < 0 || i >= (int)v.size()) return Fail(TWB_ERR_INVALID, "component index out of range");
  if (name && cap > 0) std::snprintf(name, cap, "%s", v[i].name.c_str());
  if (start) *start = v[i].start;
  if (count) *count = v[i].count;
  return TWB_OK;
}
int twb_layout_variable_set(const twb_problem* p, int i, char* name, int cap, int* col_start, int* n_cols) {
  if (!p) return Fail(TWB_ERR_INVALID, "null problem");
  return CopyComponent(p->f.var_sets, i, name, cap, col_start, n_cols);
}
int twb_layout_constraint_set(const twb_problem* p, int i, char* name, int cap, int* row_start, int* n_rows) {
  if (!p) return Fail(TWB_ERR_INVALID, "null problem");
  return CopyComponent(p->f.con_sets, i, name, cap, row_start, n_rows);
}

int twb_batch_create(const twb_problem* p, int batch_size, int device, twb_batch** out) {
  if (!p || !out || batch_size <= 0) return Fail(TWB_ERR_INVALID, "bad argument");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return Fail(TWB_ERR_NO_DEVICE, "no CUDA device: towr_b200 has no CPU evaluation path");
  if (device < 0 || device >= count) return Fail(TWB_ERR_INVALID, "device index out of range");
  if ((e = cudaSetDevice(device)) != cudaSuccess) return CudaFail(e, "cudaSetDevice");
  twb_batch* b = new (std::nothrow) twb_batch();
  if (!b) return Fail(TWB_ERR_INVALID, "out of memory");
  b->prob = p; b->B = batch_size; b->device = device;
  const twb::HostTables& t = p->f.tables;
  b->plan = t.plan;
#define TWB_UP(field)                                                                  \
  if ((e = Upload(t.field, &b->plan.field, &b->owned)) != cudaSuccess) {               \
    twb_batch_destroy(b);                                                              \
    return CudaFail(e, "table upload");                                                \
  }
  TWB_UP(samples) TWB_UP(eval_items) TWB_UP(terr) TWB_UP(force) TWB_UP(swing) TWB_UP(acc) TWB_UP(cost)
  TWB_UP(desc) TWB_UP(coef) TWB_UP(dyn_ang_basis)
#undef TWB_UP
  // instances per CTA: largest G in {4,2,1} that lets two CTAs share an SM's shared memory
  int max_smem = 0;
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  b->G = 1;
  for (int G : {4, 3, 2, 1}) {
    if (twb::EvalSmemBytes(b->plan, G) * 2 + 4096 <= (size_t)max_smem) { b->G = G; break; }
  }
  if (const char* env = std::getenv("TWB_INSTANCES_PER_CTA")) {
    int G = std::atoi(env);
    if (G == 1 || G == 2 || G == 3 || G == 4 || G == 6 || G == 8) b->G = G;
  }
  if (twb::EvalSmemBytes(b->plan, b->G) > (size_t)max_smem) {
    twb_batch_destroy(b);
    return Fail(TWB_ERR_UNSUPPORTED, "problem state does not fit in shared memory");
  }
  if ((e = cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking)) != cudaSuccess) {
    twb_batch_destroy(b);
    return CudaFail(e, "cudaStreamCreate");
  }
  *out = b;
  return TWB_OK;
}

void twb_batch_destroy(twb_batch* b) {
  if (!b) return;
  cudaSetDevice(b->device);
  for (void* p : b->owned) cudaFree(p);
  cudaFree(b->d_terrain);
  cudaFree(b->d_x); cudaFree(b->d_g); cudaFree(b->d_jac); cudaFree(b->d_cost); cudaFree(b->d_grad); cudaFree(b->d_status);
  if (b->stream) cudaStreamDestroy(b->stream);
  delete b;
}

int twb_batch_set_terrains(twb_batch* b, const int* terrain_ids) {
  if (!b) return Fail(TWB_ERR_INVALID, "null batch");
  cudaSetDevice(b->device);
  if (!terrain_ids) { cudaFree(b->d_terrain); b->d_terrain = nullptr; return TWB_OK; }
  for (int i = 0; i < b->B; ++i)
    if (terrain_ids[i] < 0 || terrain_ids[i] >= TWB_TERRAIN_COUNT) return Fail(TWB_ERR_INVALID, "unknown terrain id");
  cudaError_t e;
  if (!b->d_terrain && (e = cudaMalloc(&b->d_terrain, sizeof(int) * b->B)) != cudaSuccess) return CudaFail(e, "cudaMalloc");
  if ((e = cudaMemcpy(b->d_terrain, terrain_ids, sizeof(int) * b->B, cudaMemcpyHostToDevice)) != cudaSuccess)
    return CudaFail(e, "cudaMemcpy");
  return TWB_OK;
}

int twb_batch_launches_per_eval(const twb_batch* b, unsigned flags) { (void)flags; return b ? 1 : 0; }

int twb_batch_eval_device(twb_batch* b, const double* x, double* g, double* jac, double* cost, double* grad,
                          int* status, unsigned flags, void* stream) {
  if (!b || !x) return Fail(TWB_ERR_INVALID, "null argument");
  if ((flags & TWB_EVAL_G) && !g) return Fail(TWB_ERR_INVALID, "g requested but NULL");
  if ((flags & TWB_EVAL_JAC) && !jac) return Fail(TWB_ERR_INVALID, "jac requested but NULL");
  cudaError_t e = cudaSetDevice(b->device);
  if (e != cudaSuccess) return CudaFail(e, "cudaSetDevice");
  int rc = twb::LaunchEval(b->plan, b->G, x, g, jac, cost, grad, status, b->d_terrain, b->prob->f.spec.terrain,
                           b->B, flags, static_cast<cudaStream_t>(stream), nullptr);
  if (rc != 0) return CudaFail(static_cast<cudaError_t>(rc), "kernel launch");
  return TWB_OK;
}

int twb_batch_eval_host(twb_batch* b, const double* x, double* g, double* jac, double* cost, double* grad,
                        int* status, unsigned flags) {
  if (!b || !x) return Fail(TWB_ERR_INVALID, "null argument");
  cudaError_t e = cudaSetDevice(b->device);
  if (e != cudaSuccess) return CudaFail(e, "cudaSetDevice");
  const twb::Formulation& f = b->prob->f;
  const size_t B = b->B;
  const bool has_cost = f.has_cost && (flags & TWB_EVAL_COST);
#define TWB_ENSURE(ptr, count, T)                                                                             \
  if (!(ptr) && (e = cudaMalloc(reinterpret_cast<void**>(&(ptr)), sizeof(T) * (count))) != cudaSuccess)       \
    return CudaFail(e, "cudaMalloc");
  TWB_ENSURE(b->d_x, B * f.n, double)
  TWB_ENSURE(b->d_status, B, int)
  if (flags & TWB_EVAL_G) TWB_ENSURE(b->d_g, B * f.m, double)
  if (flags & TWB_EVAL_JAC) TWB_ENSURE(b->d_jac, B * f.nnz, double)
  if (has_cost) { TWB_ENSURE(b->d_cost, B, double) TWB_ENSURE(b->d_grad, B * f.n, double) }
#undef TWB_ENSURE
  cudaStream_t s = b->stream;
  if ((e = cudaMemcpyAsync(b->d_x, x, sizeof(double) * B * f.n, cudaMemcpyHostToDevice, s)) != cudaSuccess)
    return CudaFail(e, "H2D copy");
  int rc = twb_batch_eval_device(b, b->d_x, b->d_g, b->d_jac, has_cost ? b->d_cost : nullptr,
                                 has_cost ? b->d_grad : nullptr, b->d_status, flags, s);
  if (rc != TWB_OK) return rc;
  if ((flags & TWB_EVAL_G) && g) cudaMemcpyAsync(g, b->d_g, sizeof(double) * B * f.m, cudaMemcpyDeviceToHost, s);
  if ((flags & TWB_EVAL_JAC) && jac) cudaMemcpyAsync(jac, b->d_jac, sizeof(double) * B * f.nnz, cudaMemcpyDeviceToHost, s);
  if (has_cost && cost) cudaMemcpyAsync(cost, b->d_cost, sizeof(double) * B, cudaMemcpyDeviceToHost, s);
  if (has_cost && grad) cudaMemcpyAsync(grad, b->d_grad, sizeof(double) * B * f.n, cudaMemcpyDeviceToHost, s);
  if (status) cudaMemcpyAsync(status, b->d_status, sizeof(int) * B, cudaMemcpyDeviceToHost, s);
  if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return CudaFail(e, "evaluation");
  if (!has_cost && (flags & TWB_EVAL_COST)) {  // Problem::EvaluateCostFunction without cost terms: 0
    if (cost) for (size_t i = 0; i < B; ++i) cost[i] = 0.0;
    if (grad) std::memset(grad, 0, sizeof(double) * B * f.n);
  }
  return TWB_OK;
}

}  // extern "C"
