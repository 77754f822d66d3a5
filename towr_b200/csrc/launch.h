// launch.h — host entry points of kernels.cu
#ifndef TOWR_B200_LAUNCH_H_
#define TOWR_B200_LAUNCH_H_
#include <cuda_runtime.h>

#include <cstddef>

#include "device_tables.h"

namespace twb {

// dynamic shared memory one CTA of G instances needs
size_t EvalSmemBytes(const Plan& P, int G);

// Enqueues one batched evaluation on `stream`. Returns a cudaError_t as int.
int LaunchEval(const Plan& P, int G, const double* x, double* g, double* jac, double* cost, double* grad,
               int* status, const int* terrain_ids, int default_terrain, int B, unsigned flags, cudaStream_t stream,
               int* n_launches);

}  // namespace twb
#endif
