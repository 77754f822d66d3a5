// launch.h — host entry points of kernels.cu
#ifndef TOWR_B200_LAUNCH_H_
#define TOWR_B200_LAUNCH_H_
#include <cuda_runtime.h>

#include <cstddef>

#include "device_tables.h"

namespace twb {

// profiling hook: when set, called after every kernel launch (label "begin" marks the start of a group)
extern void (*g_after_launch)(const char* label, cudaStream_t stream);

// XT is [tiles][n+1][32], ST is [tiles][S_size][32] (tile = 32 consecutive instances).  XT/ST arguments
// point at the first tile of the sub-batch; x/cost/status/terrain_ids at its first instance.

// state row 0 := 1.0 in every tile (once, at batch creation)
int LaunchInitState(double* ST, int S_size, int n_tiles, cudaStream_t stream);

// transposes x, evaluates all splines and all constraint units of `nb` instances into the state matrix
int LaunchStateKernels(const Plan& P, const double* x, double* XT, double* ST, const int* terrain_ids,
                       int default_terrain, double* cost, int* status, int nb, bool want_cost, cudaStream_t stream,
                       int* launches);

// jac[b][s] = ST[..][desc[s]][b] * coef[s] for every CSR slot; ORs bit 0 into status[b] on NaN/Inf
int LaunchFillJac(const Plan& P, const double* ST, double* jac, int* status, int nb, int n_sms, cudaStream_t stream,
                  int* launches);

// out[b][r] = ST[..][row0 + r][b], r < rows (constraint values, cost gradient)
int LaunchTransposeOut(const Plan& P, const double* ST, int row0, int rows, double* out, int nb, cudaStream_t stream,
                       int* launches);

}  // namespace twb
#endif
