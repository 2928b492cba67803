// ceres/ceres.h -- header-only shim of the ceres:: subset that Estimator::optimization() uses
// (estimator.cpp:1388-1660), lowering a ceres::Problem to the viwb C ABI (include/viwb.h) so that the window solve
// runs on the GPU.  Put viw-fusion_b200/host first on the include path and link libviwb.so; see INTEGRATION.md.
//
// Reproduced API: CostFunction {Evaluate, num_residuals, parameter_block_sizes,
// mutable_parameter_block_sizes, set_num_residuals}, SizedCostFunction<kRes, N...>, LocalParameterization {Plus,
// ComputeJacobian, GlobalSize, LocalSize}, LossFunction::Evaluate, HuberLoss, Problem {AddParameterBlock (2 overloads),
// SetParameterBlockConstant, AddResidualBlock (pointer packs and std::vector<double*>)}, Solver::Options /
// Solver::Summary, Solve().  Ownership as in Ceres: the Problem deletes its cost functions, loss functions and local
// parameterizations; parameter memory stays with the caller and is updated in place.
//
// Only the reference's seven hot-path factor classes (viw-fusion_b200/host/factor/*.h) can be lowered; any other
// CostFunction makes Solve() return FAILURE (termination_type) instead of silently solving on the CPU.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <vector>
#include "../../../include/viwb.h"

namespace ceres {

class CostFunction {
  public:
    virtual ~CostFunction() {}
    virtual bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const = 0;
    int num_residuals() const { return num_residuals_; }
    const std::vector<int32_t> &parameter_block_sizes() const { return parameter_block_sizes_; }
    // viwb extension: which device factor this is (-1: unknown, -2: marginalization prior) and its constant record
    virtual int viwb_factor_type() const { return -1; }
    virtual const double *viwb_record() const { return nullptr; }
    virtual const viwb_prior *viwb_prior_data() const { return nullptr; }
  protected:
    std::vector<int32_t> *mutable_parameter_block_sizes() { return &parameter_block_sizes_; }
    void set_num_residuals(int n) { num_residuals_ = n; }
  private:
    int num_residuals_ = 0;
    std::vector<int32_t> parameter_block_sizes_;
};

template <int kNumResiduals, int... Ns>
class SizedCostFunction : public CostFunction {
  public:
    SizedCostFunction() { set_num_residuals(kNumResiduals); *mutable_parameter_block_sizes() = std::vector<int32_t>{Ns...}; }
};

// estimator.h reaches initial/initial_sfm.h:58, which names this template inside ReprojectionError3D::Create.  The initial structure-from-motion
// is not on the window path (it stays with the real Ceres in its own target); here the name only has to compile, and a cost
// function built from it cannot be lowered, so Solve() fails loudly on it like on any other unknown class.
template <typename Functor, int kNumResiduals, int... Ns>
class AutoDiffCostFunction : public SizedCostFunction<kNumResiduals, Ns...> {
  public:
    explicit AutoDiffCostFunction(Functor *f) : functor_(f) {}
    ~AutoDiffCostFunction() override { delete functor_; }
    bool Evaluate(double const *const *, double *, double **) const override { return false; }
  private:
    Functor *functor_;
};

class LocalParameterization {
  public:
    virtual ~LocalParameterization() {}
    virtual bool Plus(const double *x, const double *delta, double *x_plus_delta) const = 0;
    virtual bool ComputeJacobian(const double *x, double *jacobian) const = 0;
    virtual int GlobalSize() const = 0;
    virtual int LocalSize() const = 0;
    // viwb extension: bit i set -> delta[i] is zeroed in Plus (PoseSubsetParameterization / OrientationSubsetParameterization)
    virtual unsigned viwb_subset_mask() const { return 0u; }
};

class LossFunction {
  public:
    virtual ~LossFunction() {}
    virtual void Evaluate(double sq_norm, double out[3]) const = 0;
    virtual double viwb_huber_delta() const { return -1.0; }
};
class HuberLoss : public LossFunction {
  public:
    explicit HuberLoss(double a) : a_(a), b_(a * a) {}
    void Evaluate(double s, double rho[3]) const override;      // defined in viwb_shim_impl.h (device call)
    double viwb_huber_delta() const override { return a_; }
  private:
    double a_, b_;
};

enum LinearSolverType { DENSE_NORMAL_CHOLESKY, DENSE_QR, SPARSE_NORMAL_CHOLESKY, DENSE_SCHUR, SPARSE_SCHUR, ITERATIVE_SCHUR, CGNR };
enum TrustRegionStrategyType { LEVENBERG_MARQUARDT, DOGLEG };
enum TerminationType { CONVERGENCE, NO_CONVERGENCE, FAILURE, USER_SUCCESS, USER_FAILURE };

struct IterationSummary { int iteration = 0; double cost = 0; };

class Problem;
class Solver {
  public:
    struct Options {
        LinearSolverType linear_solver_type = DENSE_SCHUR;
        TrustRegionStrategyType trust_region_strategy_type = DOGLEG;
        int max_num_iterations = 50;
        double max_solver_time_in_seconds = 1e9;
        int num_threads = 1;
        bool minimizer_progress_to_stdout = false;
        bool use_nonmonotonic_steps = false;
        bool use_explicit_schur_complement = false;
        double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
        double initial_trust_region_radius = 1e4, max_trust_region_radius = 1e16, min_trust_region_radius = 1e-32;
        double min_relative_decrease = 1e-3, min_lm_diagonal = 1e-6, max_lm_diagonal = 1e32;
        int max_num_consecutive_invalid_steps = 5;
        bool jacobi_scaling = true;
    };
    struct Summary {
        std::vector<IterationSummary> iterations;
        TerminationType termination_type = FAILURE;
        double initial_cost = 0, final_cost = 0;
        int num_successful_steps = 0;
        std::string message;
        std::string BriefReport() const;
        std::string FullReport() const { return BriefReport(); }
    };
};

class Problem {
  public:
    struct Options {};
    Problem() {}
    explicit Problem(const Options &) {}
    ~Problem();
    void AddParameterBlock(double *values, int size) { AddParameterBlock(values, size, nullptr); }
    void AddParameterBlock(double *values, int size, LocalParameterization *lp);
    void SetParameterBlockConstant(double *values) { blocks_[values].constant = true; }
    void SetParameterBlockVariable(double *values) { blocks_[values].constant = false; }
    template <typename... Ts>
    void *AddResidualBlock(CostFunction *cost, LossFunction *loss, double *x0, Ts *...xs) { return AddResidualBlock(cost, loss, std::vector<double *>{x0, xs...}); }
    void *AddResidualBlock(CostFunction *cost, LossFunction *loss, const std::vector<double *> &blocks);
    int NumParameterBlocks() const { return (int)blocks_.size(); }
    int NumResidualBlocks() const { return (int)residuals_.size(); }

    struct Block { int size = 0; LocalParameterization *lp = nullptr; bool constant = false; int order = 0; };
    struct Residual { CostFunction *cost; LossFunction *loss; std::vector<double *> blocks; };
    std::map<double *, Block> blocks_;
    std::vector<Residual> residuals_;
  private:
    Problem(const Problem &) = delete;
};

void Solve(const Solver::Options &options, Problem *problem, Solver::Summary *summary);

}  // namespace ceres

// Adapters for cost functions / manifolds / losses that are NOT the shim's own classes -- above all the reference's own factor classes when
// estimator.cpp is compiled unmodified (viw-fusion_b200/host/viwb_reference_adapter.h installs them).  A class the shim does not know and no
// adapter claims still makes Solve() fail loudly.
namespace viwb_shim {
struct Lowered { int type = -1; const double *record = nullptr; const viwb_prior *prior = nullptr; bool prior_ids_exact = true; };   // type as CostFunction::viwb_factor_type(); prior_ids_exact: prior->block_id are real block ids (an adapter may only know the block sizes)
typedef bool (*CostAdapter)(const ceres::CostFunction *, Lowered *);
typedef int (*ManifoldAdapter)(const ceres::LocalParameterization *);      // subset mask, or -1: not mine
typedef double (*LossAdapter)(const ceres::LossFunction *);                // Huber delta, or -1: not mine
typedef void (*SolveBegin)();                                              // called at the start of every Solve (adapters drop their scratch records)
inline CostAdapter &cost_adapter() { static CostAdapter f = nullptr; return f; }
inline ManifoldAdapter &manifold_adapter() { static ManifoldAdapter f = nullptr; return f; }
inline LossAdapter &loss_adapter() { static LossAdapter f = nullptr; return f; }
inline SolveBegin &solve_begin() { static SolveBegin f = nullptr; return f; }
inline Lowered lower(const ceres::CostFunction *c) {
    Lowered l; l.type = c->viwb_factor_type(); l.record = c->viwb_record(); l.prior = c->viwb_prior_data();
    if (l.type == -1 && cost_adapter()) { Lowered a; if (cost_adapter()(c, &a)) l = a; }
    return l;
}
inline unsigned subset_mask(const ceres::LocalParameterization *p) {
    if (!p) return 0u;
    if (manifold_adapter()) { const int m = manifold_adapter()(p); if (m >= 0) return (unsigned)m; }
    return p->viwb_subset_mask();
}
inline double huber_delta(const ceres::LossFunction *l) {
    const double d = l->viwb_huber_delta();
    if (d < 0 && loss_adapter()) return loss_adapter()(l);
    return d;
}
}  // namespace viwb_shim

#include "../viwb_shim_impl.h"
