// viwb_factors.h -- the reference's hot-path factor and manifold classes (same names, constructor signatures and
// Evaluate/Plus contracts) re-implemented over the viwb C ABI.  Every Evaluate / marginalize runs on the GPU.
// Constructors are templates over "anything indexable with (i)" so that they accept Eigen vectors when Eigen is
// present (the reference's call sites, estimator.cpp:1528-1638) and plain test vectors when it is not.
//
//   ProjectionTwoFrameOneCamFactor / TwoFrameTwoCam / OneFrameTwoCam   factor/projection*Factor.h:21-36
//   IMUFactor(IntegrationBase*)                                       factor/imu_factor.h:23-29
//   WheelFactor(WheelIntegrationBase*)                                factor/wheel_factor.h:20-26
//   PlaneFactor()                                                     factor/plane_factor.h:21-24
//   PoseLocalParameterization, PoseSubsetParameterization, Orientation*Parameterization   factor/*parameterization.h
//   ResidualBlockInfo, MarginalizationInfo, MarginalizationFactor     factor/marginalization_factor.h:24-93
#pragma once
#include <cmath>
#include <type_traits>
#include <unordered_map>
#include "../ceres/ceres.h"

namespace viwb_shim {
inline bool gpu_evaluate(int type, const double *rec, double const *const *parameters, double *residuals, double **jacobians) {
    viwb_context *ctx = context();
    return ctx && viwb_factor_evaluate(ctx, type, &globals(), rec, parameters, residuals, jacobians) == VIWB_OK;
}
inline void quat_mul(const double *a, const double *b, double *o) {      // [x,y,z,w]
    o[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1]; o[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
    o[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0]; o[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
}
inline void plus_quat(const double *q, const double *dth, unsigned mask, int off, double *out) {   // q * deltaQ(dtheta), normalised
    double d[3]; for (int i = 0; i < 3; i++) d[i] = ((mask >> (off + i)) & 1u) ? 0.0 : dth[i];
    double dq[4] = {d[0] / 2, d[1] / 2, d[2] / 2, 1.0}, n = std::sqrt(dq[0] * dq[0] + dq[1] * dq[1] + dq[2] * dq[2] + 1.0);
    for (int i = 0; i < 4; i++) dq[i] /= n;
    quat_mul(q, dq, out);
    n = std::sqrt(out[0] * out[0] + out[1] * out[1] + out[2] * out[2] + out[3] * out[3]);
    for (int i = 0; i < 4; i++) out[i] /= n;
}
}  // namespace viwb_shim

// ---------------------------------------------------------------------------------------------- projection factors
template <int TYPE, int... Ns>
class ViwbProjectionFactor : public ceres::SizedCostFunction<2, Ns...> {
  public:
    template <class V3, class V2>
    ViwbProjectionFactor(const V3 &pts_i, const V3 &pts_j, const V2 &vel_i, const V2 &vel_j, double td_i, double td_j) {
        for (int k = 0; k < 3; k++) { rec_[k] = pts_i(k); rec_[3 + k] = pts_j(k); }
        rec_[6] = vel_i(0); rec_[7] = vel_i(1); rec_[8] = vel_j(0); rec_[9] = vel_j(1); rec_[10] = td_i; rec_[11] = td_j;
    }
    bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const override { return viwb_shim::gpu_evaluate(TYPE, rec_, parameters, residuals, jacobians); }
    int viwb_factor_type() const override { return TYPE; }
    const double *viwb_record() const override { return rec_; }
  private:
    double rec_[VIWB_VIS_OBS_DOUBLES];
};
typedef ViwbProjectionFactor<VIWB_F_PROJ_2F1C, 7, 7, 7, 1, 1> ProjectionTwoFrameOneCamFactor;
typedef ViwbProjectionFactor<VIWB_F_PROJ_2F2C, 7, 7, 7, 7, 1, 1> ProjectionTwoFrameTwoCamFactor;
typedef ViwbProjectionFactor<VIWB_F_PROJ_1F2C, 7, 7, 1, 1> ProjectionOneFrameTwoCamFactor;

// ---------------------------------------------------------------------------------------------- IMU / wheel / plane
class IMUFactor : public ceres::SizedCostFunction<15, 7, 9, 7, 9> {
  public:
    // PreInt = IntegrationBase (integration_base.h:197-214): sum_dt, delta_p, delta_q, delta_v, linearized_ba/bg, jacobian, covariance
    template <class PreInt, class = typename std::enable_if<!std::is_arithmetic<PreInt>::value>::type> explicit IMUFactor(PreInt *p) {
        rec_[0] = p->sum_dt;
        for (int k = 0; k < 3; k++) { rec_[1 + k] = p->delta_p(k); rec_[8 + k] = p->delta_v(k); rec_[11 + k] = p->linearized_ba(k); rec_[14 + k] = p->linearized_bg(k); }
        rec_[4] = p->delta_q.x(); rec_[5] = p->delta_q.y(); rec_[6] = p->delta_q.z(); rec_[7] = p->delta_q.w();
        const int blk[5][2] = {{0, 9}, {0, 12}, {3, 12}, {6, 9}, {6, 12}};
        for (int b = 0; b < 5; b++) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) rec_[17 + 9 * b + 3 * i + j] = p->jacobian(blk[b][0] + i, blk[b][1] + j);
        for (int i = 0; i < 15; i++) for (int j = 0; j < 15; j++) rec_[62 + 15 * i + j] = p->covariance(i, j);
    }
    explicit IMUFactor(const double *record) { std::memcpy(rec_, record, sizeof rec_); }
    bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const override { return viwb_shim::gpu_evaluate(VIWB_F_IMU, rec_, parameters, residuals, jacobians); }
    int viwb_factor_type() const override { return VIWB_F_IMU; }
    const double *viwb_record() const override { return rec_; }
  private:
    double rec_[VIWB_IMU_DOUBLES];
};
class WheelFactor : public ceres::SizedCostFunction<6, 7, 7, 7, 1, 1, 1, 1> {
  public:
    // PreInt = WheelIntegrationBase (wheel_integration_base.h:220-243)
    template <class PreInt, class = typename std::enable_if<!std::is_arithmetic<PreInt>::value>::type> explicit WheelFactor(PreInt *p) {
        for (int k = 0; k < 3; k++) { rec_[k] = p->delta_p(k); rec_[65 + k] = p->linearized_vel(k); rec_[68 + k] = p->linearized_gyr(k); rec_[71 + k] = p->vel_1(k); rec_[74 + k] = p->gyr_1(k); }
        rec_[3] = p->delta_q.x(); rec_[4] = p->delta_q.y(); rec_[5] = p->delta_q.z(); rec_[6] = p->delta_q.w();
        for (int i = 0; i < 6; i++) for (int j = 0; j < 3; j++) rec_[7 + 3 * i + j] = p->jacobian(i, j);
        for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) rec_[25 + 6 * i + j] = p->covariance(i, j);
        rec_[61] = p->linearized_sx; rec_[62] = p->linearized_sy; rec_[63] = p->linearized_sw; rec_[64] = p->linearized_td; rec_[77] = p->sum_dt;
    }
    explicit WheelFactor(const double *record) { std::memcpy(rec_, record, sizeof rec_); }
    bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const override { return viwb_shim::gpu_evaluate(VIWB_F_WHEEL, rec_, parameters, residuals, jacobians); }
    int viwb_factor_type() const override { return VIWB_F_WHEEL; }
    const double *viwb_record() const override { return rec_; }
  private:
    double rec_[VIWB_WHEEL_DOUBLES];
};
class PlaneFactor : public ceres::SizedCostFunction<3, 7, 7, 4, 1> {
  public:
    PlaneFactor() {}
    bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const override { return viwb_shim::gpu_evaluate(VIWB_F_PLANE, nullptr, parameters, residuals, jacobians); }
    int viwb_factor_type() const override { return VIWB_F_PLANE; }
};

// ---------------------------------------------------------------------------------------------- manifolds
// Plus is host-side glue for API fidelity (pose_local_parameterization.cpp:12-27); the solver applies it on the GPU.
class PoseSubsetParameterization : public ceres::LocalParameterization {
  public:
    explicit PoseSubsetParameterization(const std::vector<int> &constant_parameters) : mask_(0) { for (int c : constant_parameters) if (c >= 0 && c < 6) mask_ |= 1u << c; }
    bool Plus(const double *x, const double *delta, double *out) const override {
        for (int i = 0; i < 3; i++) out[i] = x[i] + (((mask_ >> i) & 1u) ? 0.0 : delta[i]);
        viwb_shim::plus_quat(x + 3, delta + 3, mask_, 3, out + 3);
        return true;
    }
    bool ComputeJacobian(const double *, double *j) const override { for (int i = 0; i < 42; i++) j[i] = 0.0; for (int i = 0; i < 6; i++) j[i * 6 + i] = 1.0; return true; }
    int GlobalSize() const override { return 7; }
    int LocalSize() const override { return 6; }
    unsigned viwb_subset_mask() const override { return mask_; }
  private:
    unsigned mask_;
};
class PoseLocalParameterization : public PoseSubsetParameterization { public: PoseLocalParameterization() : PoseSubsetParameterization(std::vector<int>()) {} };
class OrientationSubsetParameterization : public ceres::LocalParameterization {
  public:
    explicit OrientationSubsetParameterization(const std::vector<int> &constant_parameters) : mask_(0) { for (int c : constant_parameters) if (c >= 0 && c < 3) mask_ |= 1u << c; }
    bool Plus(const double *x, const double *delta, double *out) const override { viwb_shim::plus_quat(x, delta, mask_, 0, out); return true; }
    bool ComputeJacobian(const double *, double *j) const override { for (int i = 0; i < 12; i++) j[i] = 0.0; for (int i = 0; i < 3; i++) j[i * 3 + i] = 1.0; return true; }
    int GlobalSize() const override { return 4; }
    int LocalSize() const override { return 3; }
    unsigned viwb_subset_mask() const override { return mask_; }
  private:
    unsigned mask_;
};
class OrientationLocalParameterization : public OrientationSubsetParameterization { public: OrientationLocalParameterization() : OrientationSubsetParameterization(std::vector<int>()) {} };

// ---------------------------------------------------------------------------------------------- marginalization
struct ResidualBlockInfo {
    ResidualBlockInfo(ceres::CostFunction *c, ceres::LossFunction *l, std::vector<double *> blocks, std::vector<int> drop)
        : cost_function(c), loss_function(l), parameter_blocks(blocks), drop_set(drop) {}
    ceres::CostFunction *cost_function; ceres::LossFunction *loss_function;
    std::vector<double *> parameter_blocks; std::vector<int> drop_set;
};

class MarginalizationInfo {
  public:
    MarginalizationInfo() : valid(true), m(0), n(0), sum_block_size(0) { std::memset(&prior_, 0, sizeof prior_); }
    ~MarginalizationInfo() { for (auto *f : factors) { delete f->cost_function; delete f; } }
    void addResidualBlockInfo(ResidualBlockInfo *info) { factors.push_back(info); }
    void preMarginalize() {}                         // evaluation happens inside marginalize() on the GPU
    void marginalize();                              // marginalization_factor.cpp:183-312 -> viwb_marginalize
    std::vector<double *> getParameterBlocks(std::unordered_map<long, double *> &addr_shift);   // :314-334
    int localSize(int size) const { return size == 7 ? 6 : size; }
    std::vector<ResidualBlockInfo *> factors;
    bool valid; int m, n, sum_block_size;
    std::vector<int> keep_block_size, keep_block_idx; std::vector<double *> keep_block_data;
    const viwb_prior *prior() const { return &prior_; }
  private:
    viwb_prior prior_; std::vector<double> x0_, J_, r_;
    std::vector<double *> kept_addr_;       // address (before addr_shift) of each kept block, aligned with prior_.block_id
    friend class MarginalizationFactor;
};

class MarginalizationFactor : public ceres::CostFunction {
  public:
    explicit MarginalizationFactor(MarginalizationInfo *info) : marginalization_info(info) {
        for (int s : info->keep_block_size) mutable_parameter_block_sizes()->push_back(s);
        set_num_residuals(info->n);
    }
    bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const override;
    int viwb_factor_type() const override { return -2; }
    const viwb_prior *viwb_prior_data() const override { return marginalization_info->prior(); }
    MarginalizationInfo *marginalization_info;
};

#include "viwb_marginalization_impl.h"
