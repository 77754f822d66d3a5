// towr_b200_ifopt.hpp — header-only C++ host layer above the C ABI (towr_b200.h) that mirrors, for the NLP
// evaluation path, the interfaces the reference hands to IPOPT:
//
//   * the ifopt Component API — Component / VariableSet / ConstraintSet / CostTerm with GetValues, GetBounds,
//     GetRows, GetName, FillJacobianBlock(var_set, Jacobian&) — as used by every towr constraint class
//     (towr/include/towr/constraints/*.h, e.g. dynamic_constraint.h:62-103, and nodes_variables.h:33), and
//   * towr::NlpFormulation's public fields (towr/include/towr/nlp_formulation.h:100-105) with
//     GetVariableSets / GetConstraints / GetCosts (nlp_formulation.cc:63-376),
//
// served for batch index `b` of a twb_batch from the result of the last batched evaluation.  ifopt and Eigen are
// not available in this image, so the ifopt types are restated here in namespace twb_ifopt with the same member
// names (Eigen::VectorXd -> std::vector<double>, Eigen::SparseMatrix<double, RowMajor> -> a small row-major
// triplet block with coeffRef); a build that has ifopt only needs `namespace twb_ifopt = ifopt;` plus Eigen::Map
// in the three places marked "Eigen:".  Nothing here computes: all arithmetic happens in libtowr_b200.so.
#ifndef TOWR_B200_IFOPT_HPP_
#define TOWR_B200_IFOPT_HPP_

#include <algorithm>
#include <array>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "towr_b200.h"

namespace twb_ifopt {

using VectorXd = std::vector<double>;                       // Eigen: Eigen::VectorXd
struct Bounds { double lower_, upper_; };                   // ifopt::Bounds
using VecBound = std::vector<Bounds>;
static const double inf = 1e20;                             // ifopt::inf

// Eigen: Eigen::SparseMatrix<double, Eigen::RowMajor> restricted to what FillJacobianBlock implementations use
class Jacobian {
 public:
  Jacobian(int rows = 0, int cols = 0) : rows_(rows), cols_(cols) {}
  double& coeffRef(int r, int c) { return v_[{r, c}]; }
  int rows() const { return rows_; }
  int cols() const { return cols_; }
  int nonZeros() const { return (int)v_.size(); }
  const std::map<std::pair<int, int>, double>& entries() const { return v_; }   // row-major, ascending column
 private:
  int rows_, cols_;
  std::map<std::pair<int, int>, double> v_;
};

class Component {                                            // ifopt::Component
 public:
  using Ptr = std::shared_ptr<Component>;
  Component(int num_rows, std::string name) : num_rows_(num_rows), name_(std::move(name)) {}
  virtual ~Component() = default;
  virtual VectorXd GetValues() const = 0;
  virtual VecBound GetBounds() const = 0;
  int GetRows() const { return num_rows_; }
  std::string GetName() const { return name_; }
 private:
  int num_rows_;
  std::string name_;
};
class VariableSet : public Component { public: using Component::Component; };
class ConstraintSet : public Component {                     // ifopt::ConstraintSet
 public:
  using Component::Component;
  virtual void FillJacobianBlock(std::string var_set, Jacobian& jac_block) const = 0;
};
class CostTerm : public Component {                          // ifopt::CostTerm
 public:
  using Component::Component;
  virtual double GetCost() const = 0;
};

}  // namespace twb_ifopt

namespace towr_b200 {

inline void Check(int rc, const char* what) {
  if (rc != TWB_OK) throw std::runtime_error(std::string(what) + ": " + twb_last_error());
}

// B instances of one structure class; stands where B ifopt::Problem objects stand.  Host buffers are plain
// vectors here; a production driver would allocate them pinned (cudaHostAlloc) so the copies overlap.
class BatchedProblem {
 public:
  BatchedProblem(const twb_spec& spec, int batch_size, int device = 0, bool create_batch = true) : B_(batch_size) {
    Check(twb_problem_create(&spec, &prob_), "twb_problem_create");
    Check(twb_problem_dims(prob_, &n_, &m_, &nnz_), "twb_problem_dims");
    row_ptr_.resize(m_ + 1); iRow_.resize(nnz_); jCol_.resize(nnz_);
    Check(twb_problem_row_ptr(prob_, row_ptr_.data()), "twb_problem_row_ptr");
    Check(twb_problem_structure(prob_, iRow_.data(), jCol_.data()), "twb_problem_structure");
    xl_.resize(n_); xu_.resize(n_); gl_.resize(m_); gu_.resize(m_);
    Check(twb_problem_bounds(prob_, xl_.data(), xu_.data(), gl_.data(), gu_.data()), "twb_problem_bounds");
    X.assign((size_t)B_ * n_, 0.0); G.assign((size_t)B_ * m_, 0.0); JAC.assign((size_t)B_ * nnz_, 0.0);
    COST.assign(B_, 0.0); GRAD.assign((size_t)B_ * n_, 0.0); STATUS.assign(B_, 0);
    std::vector<double> x0(n_);
    Check(twb_problem_x0(prob_, x0.data()), "twb_problem_x0");
    for (int b = 0; b < B_; ++b) std::copy(x0.begin(), x0.end(), X.begin() + (size_t)b * n_);
    if (create_batch) Check(twb_batch_create(prob_, B_, device, &batch_), "twb_batch_create");   // TWB_ERR_NO_DEVICE without a GPU
  }
  ~BatchedProblem() { twb_batch_destroy(batch_); twb_problem_destroy(prob_); }
  BatchedProblem(const BatchedProblem&) = delete;
  BatchedProblem& operator=(const BatchedProblem&) = delete;

  // ifopt::Problem::SetVariables(const double*) for instance b
  void SetVariables(int b, const double* x) { std::copy(x, x + n_, X.begin() + (size_t)b * n_); }
  // one batched evaluation: Problem::EvaluateConstraints + EvalNonzerosOfJacobian + EvaluateCostFunction[Gradient]
  void Evaluate(unsigned flags = TWB_EVAL_ALL) {
    if (!batch_) throw std::runtime_error("BatchedProblem: no device batch");
    Check(twb_batch_eval_host(batch_, X.data(), G.data(), JAC.data(), COST.data(), GRAD.data(), STATUS.data(), flags), "twb_batch_eval_host");
  }
  // ---- solution post-processing of the current X (fpowr) ----
  // fpowr::GetTrajectory(solution, dt) for every instance: [B][n_samples][19 + 13 n_ee]
  std::vector<double> SampleTrajectory(double dt, int* n_samples = nullptr, int* n_values = nullptr) {
    if (!batch_) throw std::runtime_error("BatchedProblem: no device batch");
    int ns = 0, nv = 0;
    Check(twb_problem_trajectory_dims(prob_, dt, &ns, &nv), "twb_problem_trajectory_dims");
    std::vector<double> out((size_t)B_ * ns * nv);
    Check(twb_batch_sample_trajectory_host(batch_, X.data(), dt, out.data()), "twb_batch_sample_trajectory_host");
    if (n_samples) *n_samples = ns;
    if (n_values) *n_values = nv;
    return out;
  }
  // fpowr::ExtractInitialGuesses at the given sample times: [B][times.size()][49] = time | state[12] | controls[36]
  std::vector<double> ExtractInitialGuesses(const std::vector<double>& times) {
    if (!batch_) throw std::runtime_error("BatchedProblem: no device batch");
    std::vector<double> out((size_t)B_ * times.size() * 49);
    Check(twb_batch_initial_guess_host(batch_, X.data(), times.data(), (int)times.size(), out.data()), "twb_batch_initial_guess_host");
    return out;
  }
  // fpowr::ExtractFootstepPlan without the plane lookup: [B][max_states][2 + 4 n_ee], n_states[b] footstep states each
  std::vector<double> ExtractFootstepPlans(double time_horizon, std::vector<int>* n_states, int* max_states = nullptr, int* n_values = nullptr) {
    if (!batch_) throw std::runtime_error("BatchedProblem: no device batch");
    int ms = 0, nv = 0;
    Check(twb_problem_footstep_plan_dims(prob_, &ms, &nv), "twb_problem_footstep_plan_dims");
    std::vector<double> out((size_t)B_ * ms * nv);
    n_states->assign(B_, 0);
    Check(twb_batch_footstep_plan_host(batch_, X.data(), time_horizon, n_states->data(), out.data()), "twb_batch_footstep_plan_host");
    if (max_states) *max_states = ms;
    if (n_values) *n_values = nv;
    return out;
  }
  int GetNumberOfOptimizationVariables() const { return n_; }
  int GetNumberOfConstraints() const { return m_; }
  int nnz() const { return nnz_; }
  int batch_size() const { return B_; }
  const std::vector<int>& row_ptr() const { return row_ptr_; }
  const std::vector<int>& jCol() const { return jCol_; }
  const std::vector<int>& iRow() const { return iRow_; }
  const std::vector<double>& x_lower() const { return xl_; }
  const std::vector<double>& x_upper() const { return xu_; }
  const std::vector<double>& g_lower() const { return gl_; }
  const std::vector<double>& g_upper() const { return gu_; }
  const twb_problem* handle() const { return prob_; }

  std::vector<double> X, G, JAC, COST, GRAD;   // [B][n], [B][m], [B][nnz], [B], [B][n]
  std::vector<int> STATUS;

 private:
  twb_problem* prob_ = nullptr;
  twb_batch* batch_ = nullptr;
  int B_, n_ = 0, m_ = 0, nnz_ = 0;
  std::vector<int> row_ptr_, iRow_, jCol_;
  std::vector<double> xl_, xu_, gl_, gu_;
};

// "base-lin", "ee-motion_0", "ee-schedule1", ... (NodesVariables / PhaseDurations) of instance b
class GpuVariableSet : public twb_ifopt::VariableSet {
 public:
  GpuVariableSet(const BatchedProblem* p, int b, const std::string& name, int col0, int n) : VariableSet(n, name), p_(p), b_(b), col0_(col0) {}
  twb_ifopt::VectorXd GetValues() const override {
    const double* x = p_->X.data() + (size_t)b_ * p_->GetNumberOfOptimizationVariables() + col0_;
    return twb_ifopt::VectorXd(x, x + GetRows());
  }
  twb_ifopt::VecBound GetBounds() const override {
    twb_ifopt::VecBound v(GetRows());
    for (int i = 0; i < GetRows(); ++i) v[i] = {p_->x_lower()[col0_ + i], p_->x_upper()[col0_ + i]};
    return v;
  }
  int col0() const { return col0_; }
 private:
  const BatchedProblem* p_; int b_, col0_;
};

// "dynamic", "rangeofmotion-0", "terrain-ee-motion_0", ... (towr::DynamicConstraint etc.) of instance b
class GpuConstraintSet : public twb_ifopt::ConstraintSet {
 public:
  GpuConstraintSet(const BatchedProblem* p, int b, const std::string& name, int row0, int rows,
                   std::map<std::string, std::pair<int, int>> var_cols)
      : ConstraintSet(rows, name), p_(p), b_(b), row0_(row0), var_cols_(std::move(var_cols)) {}
  twb_ifopt::VectorXd GetValues() const override {
    const double* g = p_->G.data() + (size_t)b_ * p_->GetNumberOfConstraints() + row0_;
    return twb_ifopt::VectorXd(g, g + GetRows());
  }
  twb_ifopt::VecBound GetBounds() const override {
    twb_ifopt::VecBound v(GetRows());
    for (int i = 0; i < GetRows(); ++i) v[i] = {p_->g_lower()[row0_ + i], p_->g_upper()[row0_ + i]};
    return v;
  }
  // the (rows x n_var_set) block of this set's Jacobian w.r.t. one variable set: a slice of the CSR value array
  void FillJacobianBlock(std::string var_set, twb_ifopt::Jacobian& jac) const override {
    auto it = var_cols_.find(var_set);
    if (it == var_cols_.end()) return;
    const int c0 = it->second.first, nc = it->second.second;
    const double* v = p_->JAC.data() + (size_t)b_ * p_->nnz();
    for (int r = 0; r < GetRows(); ++r)
      for (int k = p_->row_ptr()[row0_ + r]; k < p_->row_ptr()[row0_ + r + 1]; ++k) {
        const int c = p_->jCol()[k];
        if (c0 <= c && c < c0 + nc) jac.coeffRef(r, c - c0) = v[k];
      }
  }
 private:
  const BatchedProblem* p_; int b_, row0_;
  std::map<std::string, std::pair<int, int>> var_cols_;
};

// the summed NodeCost terms of instance b (ifopt sums all cost terms into one row)
class GpuCostTerm : public twb_ifopt::CostTerm {
 public:
  GpuCostTerm(const BatchedProblem* p, int b) : CostTerm(1, "cost"), p_(p), b_(b) {}
  double GetCost() const override { return p_->COST[b_]; }
  twb_ifopt::VectorXd GetValues() const override { return {GetCost()}; }
  twb_ifopt::VecBound GetBounds() const override { return {{-twb_ifopt::inf, +twb_ifopt::inf}}; }
  const double* GetGradient() const { return p_->GRAD.data() + (size_t)b_ * p_->GetNumberOfOptimizationVariables(); }
 private:
  const BatchedProblem* p_; int b_;
};

// towr::NlpFormulation: same public fields; GetVariableSets / GetConstraints / GetCosts return the GPU-backed views
struct State3 { double p[3] = {0, 0, 0}, v[3] = {0, 0, 0}; };
struct BaseState { State3 lin, ang; };
class NlpFormulation {
 public:
  using VariablePtrVec = std::vector<std::shared_ptr<twb_ifopt::VariableSet>>;
  using ConstraintPtrVec = std::vector<std::shared_ptr<twb_ifopt::ConstraintSet>>;
  using CostPtrVec = std::vector<std::shared_ptr<twb_ifopt::CostTerm>>;

  explicit NlpFormulation(int robot = TWB_MONOPED, int terrain = TWB_FLAT) { Check(twb_spec_default(&params_, robot), "twb_spec_default"); params_.terrain = terrain; }

  BaseState initial_base_, final_base_;             // nlp_formulation.h:100-101
  std::vector<std::array<double, 3>> initial_ee_W_; // :102
  twb_spec params_;                                 // towr::Parameters + model_ + terrain_ (:103-105)

  // GaitGenerator::SetCombo + GetPhaseDurations / IsInContactAtStart for every foot
  void SetGait(int n_ee, int combo, double t_total) { Check(twb_spec_set_gait(&params_, n_ee, combo, t_total), "twb_spec_set_gait"); }
  void OptimizePhaseDurations() { Check(twb_spec_optimize_phase_durations(&params_), "twb_spec_optimize_phase_durations"); }

  twb_spec ToSpec() const {
    twb_spec s = params_;
    for (int k = 0; k < 3; ++k) {
      s.initial_base_lin_pos[k] = initial_base_.lin.p[k]; s.initial_base_lin_vel[k] = initial_base_.lin.v[k];
      s.initial_base_ang_pos[k] = initial_base_.ang.p[k]; s.initial_base_ang_vel[k] = initial_base_.ang.v[k];
      s.final_base_lin_pos[k] = final_base_.lin.p[k]; s.final_base_lin_vel[k] = final_base_.lin.v[k];
      s.final_base_ang_pos[k] = final_base_.ang.p[k]; s.final_base_ang_vel[k] = final_base_.ang.v[k];
    }
    if ((int)initial_ee_W_.size() != s.n_ee) throw std::runtime_error("initial_ee_W_ / ee_in_contact_at_start_ size mismatch");
    for (int e = 0; e < s.n_ee; ++e) for (int k = 0; k < 3; ++k) s.initial_ee_W[e][k] = initial_ee_W_[e][k];
    return s;
  }
  static VariablePtrVec GetVariableSets(const BatchedProblem& p, int b) {
    VariablePtrVec v; char name[64]; int c0, nc;
    for (int i = 0; i < twb_layout_num_variable_sets(p.handle()); ++i) {
      Check(twb_layout_variable_set(p.handle(), i, name, 64, &c0, &nc), "twb_layout_variable_set");
      v.push_back(std::make_shared<GpuVariableSet>(&p, b, name, c0, nc));
    }
    return v;
  }
  static ConstraintPtrVec GetConstraints(const BatchedProblem& p, int b) {
    std::map<std::string, std::pair<int, int>> cols; char name[64]; int a0, na;
    for (int i = 0; i < twb_layout_num_variable_sets(p.handle()); ++i) {
      Check(twb_layout_variable_set(p.handle(), i, name, 64, &a0, &na), "twb_layout_variable_set");
      cols[name] = {a0, na};
    }
    ConstraintPtrVec v;
    for (int i = 0; i < twb_layout_num_constraint_sets(p.handle()); ++i) {
      Check(twb_layout_constraint_set(p.handle(), i, name, 64, &a0, &na), "twb_layout_constraint_set");
      v.push_back(std::make_shared<GpuConstraintSet>(&p, b, name, a0, na, cols));
    }
    return v;
  }
  static CostPtrVec GetCosts(const BatchedProblem& p, int b) { return {std::make_shared<GpuCostTerm>(&p, b)}; }
};

}  // namespace towr_b200
#endif  // TOWR_B200_IFOPT_HPP_
