// Stand-in for <ceres/ceres.h>: exactly the interfaces the reference's factor classes derive from -- declarations only,
// no solver.  TEST INFRASTRUCTURE (oracle/_ref build); the product's own ceres-shaped shim lives in viw-fusion_b200/host/ceres/.
#pragma once
#include <vector>
#include <cstdint>
#include <cmath>
#include <limits>
#include <algorithm>
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
}  // namespace ceres
