// Conformance of include/towr_b200_ifopt.hpp with ifopt 2.0's abstract interfaces.
//
// ifopt is not available in this image, so its class DECLARATIONS are restated below verbatim (ifopt 2.0.x:
// ifopt_core/include/ifopt/{bounds.h, composite.h, variable_set.h, constraint_set.h, cost_term.h}; the interface towr
// compiles against: towr/include/towr/variables/nodes_variables.h:33, constraints/time_discretization_constraint.h:36,
// costs/node_cost.h:36) over a two-class stand-in for Eigen, with the few out-of-line definitions the link needs.  The
// product header is then compiled in its "external ifopt" mode (TWB_IFOPT_EXTERNAL), i.e. exactly as a build that
// has the real ifopt would compile it: GpuVariableSet / GpuConstraintSet / GpuCostTerm derive from THESE classes.
// static_asserts prove they are concrete and that every pure virtual of the reference interface is overridden with
// the reference's signature; the run drives them through ifopt's own entry points (Composite::SetVariables,
// ConstraintSet::LinkWithVariables / GetJacobian, CostTerm::GetValues) without a GPU (structure + slicing).
#include <cassert>
#include <cmath>
#include <cstdio>
#include <iostream>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

// ---- stand-in for the two Eigen types the interface mentions ---------------------------------------------------
namespace Eigen {
enum { RowMajor = 1 };
class VectorXd {
 public:
  VectorXd() = default;
  explicit VectorXd(int n) : v_(n, 0.0) {}
  static VectorXd Zero(int n) { return VectorXd(n); }
  int size() const { return (int)v_.size(); }
  int rows() const { return (int)v_.size(); }
  double& operator()(int i) { return v_[i]; }
  double operator()(int i) const { return v_[i]; }
  double& operator[](int i) { return v_[i]; }
  double operator[](int i) const { return v_[i]; }
  VectorXd middleRows(int start, int n) const { VectorXd o(n); for (int i = 0; i < n; ++i) o(i) = v_[start + i]; return o; }
  double* data() { return v_.data(); }
  const double* data() const { return v_.data(); }
 private:
  std::vector<double> v_;
};
template <typename Scalar, int Options>
class SparseMatrix {
 public:
  SparseMatrix(int rows = 0, int cols = 0) : rows_(rows), cols_(cols) {}
  Scalar& coeffRef(int r, int c) { return v_[{r, c}]; }
  int rows() const { return rows_; }
  int cols() const { return cols_; }
  int nonZeros() const { return (int)v_.size(); }
  const std::map<std::pair<int, int>, Scalar>& entries() const { return v_; }
  void middleRowsAssign(int row0, const SparseMatrix& block) { for (auto& kv : block.v_) v_[{row0 + kv.first.first, kv.first.second}] = kv.second; }
 private:
  int rows_, cols_;
  std::map<std::pair<int, int>, Scalar> v_;
};
}  // namespace Eigen

// ---- ifopt 2.0 declarations, verbatim -------------------------------------------------------------------------------
namespace ifopt {

struct Bounds {   // ifopt/bounds.h
  Bounds(double lower = 0.0, double upper = 0.0) { lower_ = lower; upper_ = upper; }
  double lower_;
  double upper_;
  void operator+=(double scalar) { lower_ += scalar; upper_ += scalar; }
  void operator-=(double scalar) { lower_ -= scalar; upper_ -= scalar; }
};
static const double inf = 1.0e20;
static const Bounds NoBound = Bounds(-inf, +inf);
static const Bounds BoundZero = Bounds(0.0, 0.0);
static const Bounds BoundGreaterZero = Bounds(0.0, +inf);
static const Bounds BoundSmallerZero = Bounds(-inf, 0.0);

class Component {   // ifopt/composite.h
 public:
  using Ptr = std::shared_ptr<Component>;
  using Jacobian = Eigen::SparseMatrix<double, Eigen::RowMajor>;
  using VectorXd = Eigen::VectorXd;
  using VecBound = std::vector<Bounds>;

  Component(int num_rows, const std::string& name);
  virtual ~Component() = default;

  virtual VectorXd GetValues() const = 0;
  virtual VecBound GetBounds() const = 0;
  virtual void SetVariables(const VectorXd& x) = 0;
  virtual Jacobian GetJacobian() const = 0;
  int GetRows() const;
  std::string GetName() const;
  virtual void Print(double tolerance, int& index_start) const;
  void SetRows(int num_rows);
  static const int kSpecifyLater = -1;

 private:
  int num_rows_ = kSpecifyLater;
  std::string name_;
};

class Composite : public Component {   // ifopt/composite.h
 public:
  using Ptr = std::shared_ptr<Composite>;
  using ComponentVec = std::vector<Component::Ptr>;

  Composite(const std::string& name, bool is_cost);
  virtual ~Composite() = default;

  VectorXd GetValues() const override;
  Jacobian GetJacobian() const override;
  VecBound GetBounds() const override;
  void SetVariables(const VectorXd& x) override;
  void PrintAll() const;

  const Component::Ptr GetComponent(std::string name) const;
  template <typename T>
  std::shared_ptr<T> GetComponent(const std::string& name) const;
  void AddComponent(const Component::Ptr&);
  void ClearComponents();
  const ComponentVec GetComponents() const;

 private:
  ComponentVec components_;
  bool is_cost_;
};

class VariableSet : public Component {   // ifopt/variable_set.h
 public:
  VariableSet(int n_var, const std::string& name);
  virtual ~VariableSet() = default;

  // doesn't exist for variables, generated run-time error when used
  Jacobian GetJacobian() const final { throw std::runtime_error("not implemented for variables"); };
};

class ConstraintSet : public Component {   // ifopt/constraint_set.h
 public:
  using Ptr = std::shared_ptr<ConstraintSet>;
  using VariablesPtr = Composite::Ptr;

  ConstraintSet(int n_constraints, const std::string& name);
  virtual ~ConstraintSet() = default;

  void LinkWithVariables(const VariablesPtr& x);
  Jacobian GetJacobian() const final;
  virtual void FillJacobianBlock(std::string var_set, Jacobian& jac_block) const = 0;

 protected:
  const VariablesPtr GetVariables() const { return variables_; };

 private:
  VariablesPtr variables_;
  virtual void InitVariableDependedQuantities(const VariablesPtr& x_init){};

  // doesn't exist for constraints, generated run-time error when used
  void SetVariables(const VectorXd& x) final { assert(false); };
};

class CostTerm : public ConstraintSet {   // ifopt/cost_term.h
 public:
  CostTerm(const std::string& name);
  virtual ~CostTerm() = default;

 private:
  virtual double GetCost() const = 0;

 public:
  VectorXd GetValues() const final;
  VecBound GetBounds() const final;
  void Print(double tol, int& index) const final;
};

// ---- definitions (ifopt_core/src/{composite.cc, leaves.cc}), reduced to what the link needs
Component::Component(int num_rows, const std::string& name) { num_rows_ = num_rows; name_ = name; }
int Component::GetRows() const { return num_rows_; }
void Component::SetRows(int num_rows) { num_rows_ = num_rows; }
std::string Component::GetName() const { return name_; }
void Component::Print(double, int& index) const { std::cout << name_ << "  " << num_rows_ << " rows from " << index << "\n"; index += num_rows_; }
Composite::Composite(const std::string& name, bool is_cost) : Component(0, name) { is_cost_ = is_cost; }
void Composite::AddComponent(const Component::Ptr& c) {
  components_.push_back(c);
  if (is_cost_) SetRows(1); else SetRows(GetRows() + c->GetRows());
}
void Composite::ClearComponents() { components_.clear(); SetRows(0); }
const Component::Ptr Composite::GetComponent(std::string name) const {
  for (const auto& c : components_) if (c->GetName() == name) return c;
  assert(false);
  return Component::Ptr();
}
template <typename T> std::shared_ptr<T> Composite::GetComponent(const std::string& name) const {
  Component::Ptr c = GetComponent(name);
  return std::dynamic_pointer_cast<T>(c);
}
Composite::VectorXd Composite::GetValues() const {
  VectorXd g_all = VectorXd::Zero(GetRows());
  int row = 0;
  for (const auto& c : components_) {
    int n_rows = c->GetRows();
    VectorXd g = c->GetValues();
    for (int i = 0; i < n_rows; ++i) g_all(row + i) += g(i);
    if (!is_cost_) row += n_rows;
  }
  return g_all;
}
void Composite::SetVariables(const VectorXd& x) {
  int row = 0;
  for (auto& c : components_) {
    int n_rows = c->GetRows();
    c->SetVariables(x.middleRows(row, n_rows));
    row += n_rows;
  }
}
Composite::Jacobian Composite::GetJacobian() const {
  int n_var = components_.empty() ? 0 : components_.front()->GetJacobian().cols();
  Jacobian jacobian(GetRows(), n_var);
  int row = 0;
  for (const auto& c : components_) {
    jacobian.middleRowsAssign(is_cost_ ? 0 : row, c->GetJacobian());
    if (!is_cost_) row += c->GetRows();
  }
  return jacobian;
}
Composite::VecBound Composite::GetBounds() const {
  VecBound bounds_;
  for (const auto& c : components_) { VecBound b = c->GetBounds(); bounds_.insert(bounds_.end(), b.begin(), b.end()); }
  return bounds_;
}
const Composite::ComponentVec Composite::GetComponents() const { return components_; }
void Composite::PrintAll() const { int index = 0; for (auto c : components_) c->Print(0.001, index); }
VariableSet::VariableSet(int n_var, const std::string& name) : Component(n_var, name) {}
ConstraintSet::ConstraintSet(int row_count, const std::string& name) : Component(row_count, name) {}
ConstraintSet::Jacobian ConstraintSet::GetJacobian() const {
  Jacobian jacobian(GetRows(), variables_->GetRows());
  int col = 0;
  for (const auto& vars : variables_->GetComponents()) {
    int n = vars->GetRows();
    Jacobian jac = Jacobian(GetRows(), n);
    FillJacobianBlock(vars->GetName(), jac);
    for (auto& kv : jac.entries()) jacobian.coeffRef(kv.first.first, col + kv.first.second) = kv.second;   // triplets with column offset
    col += n;
  }
  return jacobian;
}
void ConstraintSet::LinkWithVariables(const VariablesPtr& x) { variables_ = x; InitVariableDependedQuantities(x); }
CostTerm::CostTerm(const std::string& name) : ConstraintSet(1, name) {}
CostTerm::VectorXd CostTerm::GetValues() const { VectorXd cost(1); cost(0) = GetCost(); return cost; }
CostTerm::VecBound CostTerm::GetBounds() const { return VecBound(GetRows(), NoBound); }
void CostTerm::Print(double, int&) const { std::cout << GetName() << " cost " << GetValues()(0) << "\n"; }

}  // namespace ifopt

// ---- the product header, compiled against the declarations above ----------------------------------------------------
#define TWB_IFOPT_EXTERNAL ifopt
#include "../../include/towr_b200_ifopt.hpp"

using towr_b200::GpuConstraintSet;
using towr_b200::GpuCostTerm;
using towr_b200::GpuVariableSet;

static_assert(std::is_base_of<ifopt::VariableSet, GpuVariableSet>::value, "variable views are ifopt::VariableSet");
static_assert(std::is_base_of<ifopt::ConstraintSet, GpuConstraintSet>::value, "constraint views are ifopt::ConstraintSet");
static_assert(std::is_base_of<ifopt::CostTerm, GpuCostTerm>::value, "the cost view is an ifopt::CostTerm");
static_assert(!std::is_abstract<GpuVariableSet>::value, "GpuVariableSet overrides every pure virtual (GetValues, GetBounds, SetVariables)");
static_assert(!std::is_abstract<GpuConstraintSet>::value, "GpuConstraintSet overrides GetValues, GetBounds, FillJacobianBlock");
static_assert(!std::is_abstract<GpuCostTerm>::value, "GpuCostTerm overrides GetCost, FillJacobianBlock");
// the overriders have the reference's signatures (a mismatch would have hidden, not overridden, the virtual)
static_assert(std::is_same<decltype(&GpuVariableSet::GetValues), ifopt::Component::VectorXd (GpuVariableSet::*)() const>::value, "GetValues");
static_assert(std::is_same<decltype(&GpuVariableSet::GetBounds), ifopt::Component::VecBound (GpuVariableSet::*)() const>::value, "GetBounds");
static_assert(std::is_same<decltype(&GpuVariableSet::SetVariables), void (GpuVariableSet::*)(const ifopt::Component::VectorXd&)>::value, "SetVariables");
static_assert(std::is_same<decltype(&GpuConstraintSet::FillJacobianBlock), void (GpuConstraintSet::*)(std::string, ifopt::Component::Jacobian&) const>::value,
              "FillJacobianBlock");

#define REQUIRE(cond) do { if (!(cond)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } } while (0)

int main() {
  using namespace towr_b200;
  NlpFormulation formulation(TWB_MONOPED, TWB_FLAT);   // towr/test/hopper_example.cc:47-68
  formulation.initial_base_.lin.p[2] = 0.5;
  formulation.initial_ee_W_.push_back({0.0, 0.0, 0.0});
  formulation.final_base_.lin.p[0] = 1.0; formulation.final_base_.lin.p[2] = 0.5;
  const double phases[7] = {0.4, 0.2, 0.4, 0.2, 0.4, 0.2, 0.2};
  formulation.params_.n_ee = 1; formulation.params_.n_phases[0] = 7;
  for (int i = 0; i < 7; ++i) formulation.params_.phase_durations[0][i] = phases[i];
  formulation.params_.in_contact_at_start[0] = 1;

  BatchedProblem batch(formulation.ToSpec(), 2, 0, /*create_batch=*/false);
  for (size_t k = 0; k < batch.JAC.size(); ++k) batch.JAC[k] = 1.0 + (double)(k % batch.nnz());   // stand-in values: slot number + 1
  GpuProblem nlp(batch, 1);   // ifopt::Problem as hopper_example.cc:70-75 assembles it
  REQUIRE(nlp.GetNumberOfOptimizationVariables() == 339 && nlp.GetNumberOfConstraints() == 399 && !nlp.HasCostTerms());
  REQUIRE((int)nlp.GetBoundsOnConstraints().size() == 399 && (int)nlp.GetBoundsOnOptimizationVariables().size() == 339);

  // Composite::SetVariables -> NodesVariables::SetVariables of every set -> row 1 of the batch's X
  ifopt::Component::VectorXd x = nlp.GetVariableValues();
  for (int i = 0; i < x.size(); ++i) x(i) = 0.001 * i;
  nlp.GetOptVariables()->SetVariables(x);
  for (int i = 0; i < 339; ++i) REQUIRE(batch.X[339 + i] == 0.001 * i);
  REQUIRE(batch.X[0] != 0.001 * 0 || batch.X[5] != 0.005);   // instance 0 untouched (still x0)
  ifopt::Component::VectorXd back = nlp.GetVariableValues();
  for (int i = 0; i < 339; ++i) REQUIRE(back(i) == x(i));

  // ConstraintSet::GetJacobian (ifopt's own final method) over the linked variable composite: per set, the row block of the CSR
  int total = 0;
  for (const auto& c : nlp.GetConstraints().GetComponents()) {
    ifopt::Component::Jacobian jac = c->GetJacobian();
    REQUIRE(jac.rows() == c->GetRows() && jac.cols() == 339);
    total += jac.nonZeros();
  }
  REQUIRE(total == batch.nnz());
  ifopt::Component::Jacobian all = nlp.GetJacobianOfConstraints();
  REQUIRE(all.nonZeros() == batch.nnz());
  int k = 0;
  for (auto& kv : all.entries()) { REQUIRE(kv.first.first == batch.iRow()[k] && kv.first.second == batch.jCol()[k] && kv.second == 1.0 + k); ++k; }   // row-major, ascending column

  // CostTerm::GetValues (final in ifopt) -> GpuCostTerm::GetCost; the variable sets reject GetJacobian like the reference's
  batch.COST[1] = 42.0;
  REQUIRE(nlp.GetCosts().GetValues()(0) == 42.0);
  bool threw = false;
  try { nlp.GetOptVariables()->GetComponent("base-lin")->GetJacobian(); } catch (const std::runtime_error&) { threw = true; }
  REQUIRE(threw);
  REQUIRE(nlp.GetOptVariables()->GetComponent<GpuVariableSet>("ee-force_0")->GetRows() == 60);
  nlp.PrintCurrent();
  std::printf("ok\n");
  return 0;
}
