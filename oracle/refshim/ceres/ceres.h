// Stand-in for <ceres/ceres.h>: exactly the interfaces the reference's factor classes derive from -- declarations only,
// no solver.  TEST INFRASTRUCTURE (oracle/_ref build); the product's own ceres-shaped shim lives in viw-fusion_b200/host/ceres/.
#pragma once
#include <vector>
#include <cstdint>
#include <cmath>
#include <limits>
#include <algorithm>
#include <map>
#include <string>
namespace ceres {
typedef int int32;
class CostFunction {
  public:
    CostFunction() : num_residuals_(0) {}
    virtual ~CostFunction() {}
    virtual bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const = 0;
    const std::vector<int32> &parameter_block_sizes() const { return parameter_block_sizes_; }
    int num_residuals() const { return num_residuals_; }
  protected:
    std::vector<int32> *mutable_parameter_block_sizes() { return &parameter_block_sizes_; }
    void set_num_residuals(int n) { num_residuals_ = n; }
  private:
    std::vector<int32> parameter_block_sizes_;
    int num_residuals_;
};
template <int kNumResiduals, int... Ns> class SizedCostFunction : public CostFunction {
  public:
    SizedCostFunction() { set_num_residuals(kNumResiduals); *mutable_parameter_block_sizes() = std::vector<int32>{Ns...}; }
    virtual ~SizedCostFunction() {}
};
class LocalParameterization {
  public:
    virtual ~LocalParameterization() {}
    virtual bool Plus(const double *x, const double *delta, double *x_plus_delta) const = 0;
    virtual bool ComputeJacobian(const double *x, double *jacobian) const = 0;
    virtual int GlobalSize() const = 0;
    virtual int LocalSize() const = 0;
};
class LossFunction {
  public:
    virtual ~LossFunction() {}
    virtual void Evaluate(double sq_norm, double out[3]) const = 0;
};
// ceres::HuberLoss (third party, loss_function.cc): rho(s) = s for s <= delta^2, 2 delta sqrt(s) - delta^2 above
class HuberLoss : public LossFunction {
  public:
    explicit HuberLoss(double a) : a_(a), b_(a * a) {}
    virtual void Evaluate(double s, double rho[3]) const {
        if (s > b_) {
            const double r = std::sqrt(s);
            rho[0] = 2.0 * a_ * r - b_;
            rho[1] = std::max(std::numeric_limits<double>::min(), a_ / r);
            rho[2] = -rho[1] / (2.0 * s);
        } else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
    }
  private:
    const double a_, b_;
};
// ---- the solver-side API as estimator.cpp / initial_sfm.h spell it.  Problem RECORDS what it is given (oracle/refshim/ref_glue.cpp reads
// the record to pin the problem assembly); Solve() calls a hook the glue installs (it plays back a solution computed elsewhere).
template <typename Functor, int kNumResiduals, int... Ns> class AutoDiffCostFunction : public SizedCostFunction<kNumResiduals, Ns...> {
  public:
    explicit AutoDiffCostFunction(Functor *f) : functor_(f) {}
    virtual ~AutoDiffCostFunction() { delete functor_; }
    virtual bool Evaluate(double const *const *, double *, double **) const { return false; }     // GlobalSFM's bundle adjustment is not exercised
  private:
    Functor *functor_;
};
enum LinearSolverType { DENSE_NORMAL_CHOLESKY, DENSE_QR, SPARSE_NORMAL_CHOLESKY, DENSE_SCHUR, SPARSE_SCHUR, ITERATIVE_SCHUR, CGNR };
enum TrustRegionStrategyType { LEVENBERG_MARQUARDT, DOGLEG };
enum TerminationType { CONVERGENCE, NO_CONVERGENCE, FAILURE, USER_SUCCESS, USER_FAILURE };
struct IterationSummary { int iteration; double cost; };
class Problem {
  public:
    struct Options {};
    struct Block { int size; LocalParameterization *lp; bool constant; int order; };
    struct Residual { CostFunction *cost; LossFunction *loss; std::vector<double *> blocks; };
    Problem() {}
    explicit Problem(const Options &) {}
    ~Problem();                                                   // defined by the glue: Ceres' Problem owns its cost / loss / parameterization objects
    void AddParameterBlock(double *values, int size, LocalParameterization *lp = nullptr) {
        for (auto &b : order_) if (b == values) { Block &k = blocks_[values]; k.size = size; if (lp) k.lp = lp; return; }
        Block k; k.size = size; k.lp = lp; k.constant = false; k.order = (int)order_.size(); blocks_[values] = k; order_.push_back(values);
    }
    void SetParameterBlockConstant(double *values) { blocks_[values].constant = true; }
    void SetParameterization(double *values, LocalParameterization *lp) { blocks_[values].lp = lp; }
    template <typename... Ts> void *AddResidualBlock(CostFunction *cost, LossFunction *loss, double *x0, Ts *... xs) { return AddResidualBlock(cost, loss, std::vector<double *>{x0, xs...}); }
    void *AddResidualBlock(CostFunction *cost, LossFunction *loss, const std::vector<double *> &blocks) {
        const std::vector<int32> &sizes = cost->parameter_block_sizes();
        for (size_t i = 0; i < blocks.size(); i++) { bool known = false; for (auto &b : order_) if (b == blocks[i]) known = true; if (!known) AddParameterBlock(blocks[i], i < sizes.size() ? sizes[i] : 0); }
        residuals_.push_back(Residual{cost, loss, blocks});
        return nullptr;
    }
    std::map<double *, Block> blocks_;
    std::vector<double *> order_;
    std::vector<Residual> residuals_;
};
class Solver {
  public:
    struct Options {
        LinearSolverType linear_solver_type = SPARSE_NORMAL_CHOLESKY;
        TrustRegionStrategyType trust_region_strategy_type = LEVENBERG_MARQUARDT;
        int max_num_iterations = 50, num_threads = 1;
        double max_solver_time_in_seconds = 1e9;
        bool minimizer_progress_to_stdout = false, use_nonmonotonic_steps = false;
    };
    struct Summary {
        std::vector<IterationSummary> iterations;
        double initial_cost = 0, final_cost = 0;
        TerminationType termination_type = NO_CONVERGENCE;
        std::string BriefReport() const { return "refshim: solution played back"; }
        std::string FullReport() const { return BriefReport(); }
    };
};
void Solve(const Solver::Options &options, Problem *problem, Solver::Summary *summary);     // defined by the glue
}  // namespace ceres
