// ref_glue.cpp -- TEST INFRASTRUCTURE ONLY.  C entry points over the reference's OWN factor classes, compiled unmodified from
// /root/reference/vins_estimator/src/factor/ (oracle/Makefile `ref` target; third-party headers replaced by oracle/refshim/).
// The entry points have the contract of vo_factor_evaluate / viwb_factor_evaluate (include/viwb.h), so that tests can run the same
// inputs through the reference code, the restated oracle and the CUDA library.  Nothing here is product code.
#include "factor/projectionTwoFrameOneCamFactor.h"
#include "factor/projectionTwoFrameTwoCamFactor.h"
#include "factor/projectionOneFrameTwoCamFactor.h"
#include "factor/imu_factor.h"
#include "factor/pose_local_parameterization.h"
#include "factor/pose_subset_parameterization.h"
#include "factor/orientation_subset_parameterization.h"
#include "factor/wheel_factor.h"
#include "factor/plane_factor.h"
#include "../../include/viwb.h"

// ---- the globals of estimator/parameters.cpp that the factor code reads (that file itself needs ROS + OpenCV and is not compiled)
double ACC_N, ACC_W, GYR_N, GYR_W;
double VEL_N_wheel, GYR_N_wheel, SX, SY, SW;
double ROLL_N, PITCH_N, ZPW_N, ROLL_N_INV, PITCH_N_INV, ZPW_N_INV;
Eigen::Vector3d G;
std::vector<Eigen::Matrix3d> RIC;
std::vector<Eigen::Vector3d> TIC;
Eigen::Matrix3d RIO;
Eigen::Vector3d TIO;
double TD, TD_WHEEL;
int ESTIMATE_EXTRINSIC, ESTIMATE_EXTRINSIC_WHEEL, ESTIMATE_INTRINSIC_WHEEL, ESTIMATE_TD, ESTIMATE_TD_WHEEL, USE_IMU, USE_WHEEL, USE_PLANE, STEREO;

static void fill_pre_integration(IntegrationBase &pre, const double *c) {
    // record layout of include/viwb.h (VIWB_IMU_DOUBLES)
    pre.sum_dt = c[0];
    pre.delta_p = Eigen::Vector3d(c[1], c[2], c[3]);
    pre.delta_q = Eigen::Quaterniond(c[7], c[4], c[5], c[6]);
    pre.delta_v = Eigen::Vector3d(c[8], c[9], c[10]);
    pre.linearized_ba = Eigen::Vector3d(c[11], c[12], c[13]);
    pre.linearized_bg = Eigen::Vector3d(c[14], c[15], c[16]);
    pre.jacobian.setIdentity();
    const int br[5] = {O_P, O_P, O_R, O_V, O_V}, bc[5] = {O_BA, O_BG, O_BG, O_BA, O_BG};
    for (int k = 0; k < 5; k++) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) pre.jacobian(br[k] + i, bc[k] + j) = c[17 + 9 * k + 3 * i + j];
    for (int i = 0; i < 15; i++) for (int j = 0; j < 15; j++) pre.covariance(i, j) = c[62 + 15 * i + j];
}

extern "C" int ref_factor_evaluate(int type, const viwb_globals *g, const double *c, const double *const *parameters, double *residuals, double **jacobians) {
    G = Eigen::Vector3d(g->G[0], g->G[1], g->G[2]);
    Eigen::Matrix2d si; si << g->vis_sqrt_info[0], g->vis_sqrt_info[1], g->vis_sqrt_info[2], g->vis_sqrt_info[3];
    if (type >= VIWB_F_PROJ_2F1C && type <= VIWB_F_PROJ_1F2C) {
        const Eigen::Vector3d pi(c[0], c[1], c[2]), pj(c[3], c[4], c[5]);
        const Eigen::Vector2d vi(c[6], c[7]), vj(c[8], c[9]);
        if (type == VIWB_F_PROJ_2F1C) { ProjectionTwoFrameOneCamFactor::sqrt_info = si; ProjectionTwoFrameOneCamFactor f(pi, pj, vi, vj, c[10], c[11]); return f.Evaluate(parameters, residuals, jacobians) ? 0 : 1; }
        if (type == VIWB_F_PROJ_2F2C) { ProjectionTwoFrameTwoCamFactor::sqrt_info = si; ProjectionTwoFrameTwoCamFactor f(pi, pj, vi, vj, c[10], c[11]); return f.Evaluate(parameters, residuals, jacobians) ? 0 : 1; }
        ProjectionOneFrameTwoCamFactor::sqrt_info = si; ProjectionOneFrameTwoCamFactor f(pi, pj, vi, vj, c[10], c[11]); return f.Evaluate(parameters, residuals, jacobians) ? 0 : 1;
    }
    if (type == VIWB_F_IMU) {
        IntegrationBase pre(Eigen::Vector3d::Zero(), Eigen::Vector3d::Zero(), Eigen::Vector3d::Zero(), Eigen::Vector3d::Zero());
        fill_pre_integration(pre, c);
        IMUFactor f(&pre);
        return f.Evaluate(parameters, residuals, jacobians) ? 0 : 1;
    }
    if (type == VIWB_F_WHEEL) {
        // record layout of include/viwb.h (VIWB_WHEEL_DOUBLES); linearized_vel / linearized_gyr are const members set by the constructor
        WheelIntegrationBase pre(Eigen::Vector3d(c[65], c[66], c[67]), Eigen::Vector3d(c[68], c[69], c[70]), c[61], c[62], c[63], c[64]);
        pre.delta_p = Eigen::Vector3d(c[0], c[1], c[2]);
        pre.delta_q = Eigen::Quaterniond(c[6], c[3], c[4], c[5]);
        for (int i = 0; i < 6; i++) for (int j = 0; j < 3; j++) pre.jacobian(i, j) = c[7 + 3 * i + j];
        for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) pre.covariance(i, j) = c[25 + 6 * i + j];
        pre.vel_1 = Eigen::Vector3d(c[71], c[72], c[73]);
        pre.gyr_1 = Eigen::Vector3d(c[74], c[75], c[76]);
        pre.sum_dt = c[77];
        WheelFactor f(&pre);
        return f.Evaluate(parameters, residuals, jacobians) ? 0 : 1;
    }
    if (type == VIWB_F_PLANE) {
        PITCH_N_INV = g->plane_sqrt_info[0]; ROLL_N_INV = g->plane_sqrt_info[1]; ZPW_N_INV = g->plane_sqrt_info[2];
        PlaneFactor f;
        return f.Evaluate(parameters, residuals, jacobians) ? 0 : 1;
    }
    return 2;
}

// WheelIntegrationBase::push_back over a sample buffer -> the 78-double record (same contract as vo_wheel_preintegrate)
extern "C" void ref_wheel_preintegrate(int n, const double *dt, const double *vel, const double *gyr, const double *s, double td, const double *noise, double *rec) {
    VEL_N_wheel = noise[0]; GYR_N_wheel = noise[1];
    WheelIntegrationBase pre(Eigen::Vector3d(vel[0], vel[1], vel[2]), Eigen::Vector3d(gyr[0], gyr[1], gyr[2]), s[0], s[1], s[2], td);
    for (int k = 0; k < n; k++) pre.push_back(dt[k], Eigen::Vector3d(vel[3 * (k + 1)], vel[3 * (k + 1) + 1], vel[3 * (k + 1) + 2]), Eigen::Vector3d(gyr[3 * (k + 1)], gyr[3 * (k + 1) + 1], gyr[3 * (k + 1) + 2]));
    for (int i = 0; i < 3; i++) { rec[i] = pre.delta_p(i); rec[65 + i] = pre.linearized_vel(i); rec[68 + i] = pre.linearized_gyr(i); rec[71 + i] = pre.vel_1(i); rec[74 + i] = pre.gyr_1(i); }
    rec[3] = pre.delta_q.x(); rec[4] = pre.delta_q.y(); rec[5] = pre.delta_q.z(); rec[6] = pre.delta_q.w();
    for (int i = 0; i < 6; i++) for (int j = 0; j < 3; j++) rec[7 + 3 * i + j] = pre.jacobian(i, j);
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) rec[25 + 6 * i + j] = pre.covariance(i, j);
    rec[61] = pre.linearized_sx; rec[62] = pre.linearized_sy; rec[63] = pre.linearized_sw; rec[64] = pre.linearized_td; rec[77] = pre.sum_dt;
}

// IntegrationBase::push_back over a sample buffer -> the 287-double record (same contract as vo_imu_preintegrate)
extern "C" void ref_imu_preintegrate(int n, const double *dt, const double *acc, const double *gyr, const double *ba, const double *bg, const double *noise, double *rec) {
    ACC_N = noise[0]; GYR_N = noise[1]; ACC_W = noise[2]; GYR_W = noise[3];
    IntegrationBase pre(Eigen::Vector3d(acc[0], acc[1], acc[2]), Eigen::Vector3d(gyr[0], gyr[1], gyr[2]), Eigen::Vector3d(ba[0], ba[1], ba[2]), Eigen::Vector3d(bg[0], bg[1], bg[2]));
    for (int k = 0; k < n; k++) pre.push_back(dt[k], Eigen::Vector3d(acc[3 * (k + 1)], acc[3 * (k + 1) + 1], acc[3 * (k + 1) + 2]), Eigen::Vector3d(gyr[3 * (k + 1)], gyr[3 * (k + 1) + 1], gyr[3 * (k + 1) + 2]));
    rec[0] = pre.sum_dt;
    for (int i = 0; i < 3; i++) { rec[1 + i] = pre.delta_p(i); rec[8 + i] = pre.delta_v(i); rec[11 + i] = pre.linearized_ba(i); rec[14 + i] = pre.linearized_bg(i); }
    rec[4] = pre.delta_q.x(); rec[5] = pre.delta_q.y(); rec[6] = pre.delta_q.z(); rec[7] = pre.delta_q.w();
    const int br[5] = {O_P, O_P, O_R, O_V, O_V}, bc[5] = {O_BA, O_BG, O_BG, O_BA, O_BG};
    for (int k = 0; k < 5; k++) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) rec[17 + 9 * k + 3 * i + j] = pre.jacobian(br[k] + i, bc[k] + j);
    for (int i = 0; i < 15; i++) for (int j = 0; j < 15; j++) rec[62 + 15 * i + j] = pre.covariance(i, j);
}

// LocalParameterization::Plus / ComputeJacobian of the three manifolds; kind 0 = PoseLocal, 1 = PoseSubset, 2 = OrientationSubset;
// mask bit i = tangent component i held constant.  jacobian: global x local, row-major (may be NULL)
extern "C" int ref_manifold(int kind, unsigned mask, const double *x, const double *delta, double *x_plus_delta, double *jacobian) {
    std::vector<int> constant;
    for (int i = 0; i < 6; i++) if (mask & (1u << i)) constant.push_back(i);
    ceres::LocalParameterization *p = nullptr;
    if (kind == 0) p = new PoseLocalParameterization();
    else if (kind == 1) p = new PoseSubsetParameterization(constant);
    else if (kind == 2) p = new OrientationSubsetParameterization(constant);
    else return 2;
    bool ok = p->Plus(x, delta, x_plus_delta);
    if (jacobian) ok = p->ComputeJacobian(x, jacobian) && ok;
    delete p;
    return ok ? 0 : 1;
}
