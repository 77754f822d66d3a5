/* The C ABI from plain C99 (no C++, no CUDA headers): builds the hopper_example.cc problem (towr/test/
 * hopper_example.cc:47-68) and checks dimensions, structure order and bounds; without a GPU twb_batch_create must
 * refuse with TWB_ERR_NO_DEVICE (there is no CPU evaluation path). */
#include <stdio.h>
#include <stdlib.h>
#include "../../include/towr_b200.h"

#define REQUIRE(c) do { if (!(c)) { printf("FAILED %s:%d: %s (%s)\n", __FILE__, __LINE__, #c, twb_last_error()); return 1; } } while (0)

int main(int argc, char** argv) {
  const int expect_gpu = argc > 1;
  twb_spec spec;
  twb_problem* prob = NULL;
  twb_batch* batch = NULL;
  int n, m, nnz, k, rc;
  const double phases[7] = {0.4, 0.2, 0.4, 0.2, 0.4, 0.2, 0.2};
  REQUIRE(twb_spec_default(&spec, TWB_MONOPED) == TWB_OK);
  spec.terrain = TWB_FLAT; spec.n_ee = 1; spec.n_phases[0] = 7; spec.in_contact_at_start[0] = 1;
  for (k = 0; k < 7; ++k) spec.phase_durations[0][k] = phases[k];
  spec.initial_base_lin_pos[2] = 0.5; spec.final_base_lin_pos[0] = 1.0; spec.final_base_lin_pos[2] = 0.5;
  REQUIRE(twb_problem_create(&spec, &prob) == TWB_OK);
  REQUIRE(twb_problem_dims(prob, &n, &m, &nnz) == TWB_OK && n == 339 && m == 399 && nnz == 5392);
  {
    int* iRow = (int*)malloc(sizeof(int) * nnz); int* jCol = (int*)malloc(sizeof(int) * nnz);
    double* gl = (double*)malloc(sizeof(double) * m); double* gu = (double*)malloc(sizeof(double) * m);
    REQUIRE(twb_problem_structure(prob, iRow, jCol) == TWB_OK);
    for (k = 1; k < nnz; ++k) REQUIRE(iRow[k] > iRow[k - 1] || (iRow[k] == iRow[k - 1] && jCol[k] > jCol[k - 1]));   /* row-major, ascending columns */
    REQUIRE(twb_problem_bounds(prob, NULL, NULL, gl, gu) == TWB_OK);
    REQUIRE(gl[10] == 0.0 && gu[10] == 0.0);   /* first dynamic row: BoundZero */
    free(iRow); free(jCol); free(gl); free(gu);
  }
  spec.n_ee = 2;                               /* a monoped with two feet: refused, with a message */
  { twb_problem* bad = NULL; REQUIRE(twb_problem_create(&spec, &bad) == TWB_ERR_INVALID && bad == NULL && twb_last_error()[0] != 0); }
  rc = twb_batch_create(prob, 4, 0, &batch);
  if (expect_gpu) REQUIRE(rc == TWB_OK && batch != NULL);
  else REQUIRE(rc == TWB_OK || (rc == TWB_ERR_NO_DEVICE && batch == NULL));
  twb_batch_destroy(batch);
  twb_problem_destroy(prob);
  printf("ok %s\n", twb_version());
  return 0;
}
