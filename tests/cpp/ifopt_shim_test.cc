// Exercises include/towr_b200_ifopt.hpp the way towr/test/hopper_example.cc drives the reference:
// fill an NlpFormulation, get variable / constraint / cost sets, ask them for values, bounds and Jacobian blocks.
//   ./ifopt_shim_test          structure and the slicing logic (no GPU needed)
//   ./ifopt_shim_test --gpu    additionally one batched evaluation; prints sums the Python test compares
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/towr_b200_ifopt.hpp"

#define REQUIRE(cond) do { if (!(cond)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } } while (0)

int main(int argc, char** argv) {
  const bool gpu = argc > 1 && std::strcmp(argv[1], "--gpu") == 0;
  using namespace towr_b200;
  // towr/test/hopper_example.cc:47-68
  NlpFormulation formulation(TWB_MONOPED, TWB_FLAT);
  formulation.initial_base_.lin.p[2] = 0.5;
  formulation.initial_ee_W_.push_back({0.0, 0.0, 0.0});
  formulation.final_base_.lin.p[0] = 1.0; formulation.final_base_.lin.p[2] = 0.5;
  const double phases[7] = {0.4, 0.2, 0.4, 0.2, 0.4, 0.2, 0.2};
  formulation.params_.n_ee = 1; formulation.params_.n_phases[0] = 7;
  for (int i = 0; i < 7; ++i) formulation.params_.phase_durations[0][i] = phases[i];
  formulation.params_.in_contact_at_start[0] = 1;

  const int B = 3;
  BatchedProblem nlp(formulation.ToSpec(), B, 0, gpu);
  REQUIRE(nlp.GetNumberOfOptimizationVariables() == 339);   // 126 + 126 + 27 + 60
  REQUIRE(nlp.GetNumberOfConstraints() == 399);
  REQUIRE(nlp.nnz() == 5392);

  const int b = 1;
  auto vars = NlpFormulation::GetVariableSets(nlp, b);
  auto cons = NlpFormulation::GetConstraints(nlp, b);
  auto costs = NlpFormulation::GetCosts(nlp, b);
  REQUIRE(vars.size() == 4 && vars[0]->GetName() == "base-lin" && vars[2]->GetName() == "ee-motion_0" && vars[3]->GetRows() == 60);
  REQUIRE(cons.size() == 7 && cons[1]->GetName() == "dynamic" && cons[1]->GetRows() == 132 && cons[4]->GetName() == "rangeofmotion-0");
  int rows = 0; for (auto& c : cons) rows += c->GetRows();
  REQUIRE(rows == 399);
  for (auto& bd : cons[1]->GetBounds()) REQUIRE(bd.lower_ == 0.0 && bd.upper_ == 0.0);      // BoundZero, dynamic_constraint.cc:66-71
  REQUIRE(vars[0]->GetBounds()[2].lower_ == 0.5 && vars[0]->GetBounds()[2].upper_ == 0.5);   // initial base z is fixed
  REQUIRE(std::fabs(vars[3]->GetValues()[4] - 20.0 * 9.80665) < 1e-12);                      // f_z = m g / n_ee, nlp_formulation.cc:171-176

  if (gpu) nlp.Evaluate(TWB_EVAL_ALL);
  else for (size_t k = 0; k < nlp.JAC.size(); ++k) nlp.JAC[k] = 1.0 + (double)(k % 5392);      // stand-in values: slot number + 1

  // ConstraintSet::GetJacobian as ifopt does it: FillJacobianBlock once per variable set; together the blocks are the CSR row
  int filled = 0; double sum = 0.0, csr_sum = 0.0;
  for (auto& c : cons)
    for (auto& v : vars) {
      twb_ifopt::Jacobian block(c->GetRows(), v->GetRows());
      c->FillJacobianBlock(v->GetName(), block);
      filled += block.nonZeros();
      for (auto& kv : block.entries()) {
        REQUIRE(kv.first.first >= 0 && kv.first.first < c->GetRows() && kv.first.second >= 0 && kv.first.second < v->GetRows());
        sum += kv.second;
      }
    }
  for (int k = 0; k < nlp.nnz(); ++k) csr_sum += nlp.JAC[(size_t)b * nlp.nnz() + k];
  REQUIRE(filled == nlp.nnz());
  REQUIRE(std::fabs(sum - csr_sum) <= 1e-9 * std::fabs(csr_sum));
  {
    twb_ifopt::Jacobian block(cons[1]->GetRows(), vars[1]->GetRows());   // dynamic w.r.t. base-ang: 3 angular rows x 12 columns per sample
    cons[1]->FillJacobianBlock("base-ang", block);
    REQUIRE(block.nonZeros() == 22 * 3 * 12);
  }
  if (gpu) {
    double g_sum = 0.0; for (double v : cons[1]->GetValues()) g_sum += v;
    std::printf("dynamic_g_sum %.17g\njac_sum %.17g\ncost %.17g\nstatus %d\n", g_sum, csr_sum, costs[0]->GetValues()[0], nlp.STATUS[b]);
    // post-processing of the current X (fpowr): footstep plan of the hopper gait, initial guesses at two times
    std::vector<int> n_states; int max_states = 0, nv = 0;
    std::vector<double> plans = nlp.ExtractFootstepPlans(2.0, &n_states, &max_states, &nv);
    REQUIRE(nv == 6 && n_states[b] == 7 && max_states >= 7);
    double dur = 0.0; for (int i = 0; i < n_states[b]; ++i) dur += plans[((size_t)b * max_states + i) * nv + 1];
    REQUIRE(std::fabs(dur - 2.0) < 1e-12);
    std::vector<double> ig = nlp.ExtractInitialGuesses({0.0, 1.0});
    REQUIRE(ig.size() == (size_t)B * 2 * 49 && ig[((size_t)b * 2 + 1) * 49] == 1.0);
    int ns = 0, nvt = 0;
    std::vector<double> traj = nlp.SampleTrajectory(0.5, &ns, &nvt);
    REQUIRE(ns == 5 && nvt == 32 && std::fabs(traj[((size_t)b * ns + 2) * nvt + 0] - ig[((size_t)b * 2 + 1) * 49 + 1]) < 1e-13);   // base x at t = 1
    std::printf("footstep_states %d\n", n_states[b]);
    // ifopt::Problem view of instance b as IpoptAdapter drives it: a new iterate re-evaluates the batch lazily, once
    GpuProblem ip(nlp, b);
    std::vector<double> x(nlp.X.begin() + (size_t)b * 339, nlp.X.begin() + (size_t)(b + 1) * 339);
    for (int i = 0; i < 339; ++i) x[i] += 0.01 * std::sin(1.0 + i);
    twb_ifopt::VectorXd g1 = ip.EvaluateConstraints(x.data());
    std::vector<double> vals(nlp.nnz());
    ip.EvalNonzerosOfJacobian(x.data(), vals.data());
    nlp.SetVariables(b, x.data()); nlp.Evaluate(TWB_EVAL_ALL);     // the explicit route
    for (int i = 0; i < 399; ++i) REQUIRE(g1(i) == nlp.G[(size_t)b * 399 + i]);
    for (int k = 0; k < nlp.nnz(); ++k) REQUIRE(vals[k] == nlp.JAC[(size_t)b * nlp.nnz() + k]);
    REQUIRE(std::fabs(g1(10 + 3)) > 1e-6);                          // the perturbed iterate violates the dynamics
    REQUIRE(ip.GetConstraints().GetComponent("dynamic")->GetJacobian().nonZeros() == nlp.row_ptr()[10 + 132] - nlp.row_ptr()[10]);
  }
  std::printf("ok\n");
  return 0;
}
