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
#include "factor/marginalization_factor.h"
#include "estimator/feature_manager.h"
#include "camodocal/camera_models/PinholeCamera.h"
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
double INIT_DEPTH = 5.0, MIN_PARALLAX = 10.0 / 460.0;
int NUM_OF_CAM = 2, ROW, COL, MULTIPLE_THREAD, ONLY_INITIAL_WITH_WHEEL;
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

// ---------------------------------------------------------------------------------------------------- marginalization
// The arithmetic is the reference's (ResidualBlockInfo::Evaluate with the loss corrector, MarginalizationInfo::preMarginalize /
// marginalize with its four threads, MarginalizationFactor::Evaluate: factor/marginalization_factor.cpp, compiled unmodified).  What
// this glue restates is only the bookkeeping of Estimator::optimization() that hands the factors over (estimator.cpp:1669-1790 for
// MARGIN_OLD, :1803-1842 for MARGIN_SECOND_NEW), driven by the same viwb_problem tables the oracle and the CUDA library consume.
static MarginalizationInfo *info_from_prior(const viwb_prior *pr, std::vector<double *> &x0_blocks) {
    MarginalizationInfo *info = new MarginalizationInfo();
    info->n = pr->n; info->m = 0; info->valid = pr->valid != 0;
    for (int k = 0; k < pr->num_blocks; k++) {
        const int b = pr->block_id[k], size = viwb_block_size(b);
        double *d = new double[size];
        memcpy(d, pr->x0 + viwb_block_offset(b), sizeof(double) * size);
        x0_blocks.push_back(d);
        info->keep_block_size.push_back(size);
        info->keep_block_idx.push_back(pr->block_idx[k]);
        info->keep_block_data.push_back(d);
    }
    info->linearized_jacobians.resize(pr->n, pr->n);
    info->linearized_residuals.resize(pr->n);
    for (int i = 0; i < pr->n; i++) { info->linearized_residuals(i) = pr->r[i]; for (int j = 0; j < pr->n; j++) info->linearized_jacobians(i, j) = pr->J[(size_t)i * pr->n + j]; }
    return info;
}

// MarginalizationFactor::Evaluate: same contract as vo_prior_evaluate (jacobian n x VIWB_STATE_FIXED row-major, or NULL)
extern "C" int ref_prior_evaluate(const viwb_prior *pr, const double *state, double *residuals, double *jacobian) {
    std::vector<double *> x0;
    MarginalizationInfo *info = info_from_prior(pr, x0);
    MarginalizationFactor f(info);
    std::vector<const double *> params; std::vector<std::vector<double>> jb; std::vector<double *> jp;
    for (int k = 0; k < pr->num_blocks; k++) { const int b = pr->block_id[k]; params.push_back(state + viwb_block_offset(b)); jb.emplace_back((size_t)pr->n * viwb_block_size(b)); }
    for (auto &v : jb) jp.push_back(v.data());
    const bool ok = f.Evaluate(params.data(), residuals, jacobian ? jp.data() : nullptr);
    if (jacobian) {
        memset(jacobian, 0, sizeof(double) * pr->n * VIWB_STATE_FIXED);
        for (int k = 0; k < pr->num_blocks; k++) { const int b = pr->block_id[k], size = viwb_block_size(b), off = viwb_block_offset(b);
            for (int i = 0; i < pr->n; i++) for (int j = 0; j < size; j++) jacobian[(size_t)i * VIWB_STATE_FIXED + off + j] = jb[k][(size_t)i * size + j]; }
    }
    for (double *d : x0) delete[] d;
    delete info;
    return ok ? 0 : 1;
}

// Estimator::optimization()'s marginalization step.  Outputs: mn = {m, n, number of kept blocks}; kept block ids / column offsets (idx - m);
// J (n x n row-major) = linearized_jacobians, r (n) = linearized_residuals.  Returns 0, or 3 when nothing is left to keep.
extern "C" int ref_marginalize(const viwb_problem *p, const double *state_in, int margin_flag, int32_t *mn, int32_t *block_id, int32_t *block_idx, double *J, double *r) {
    const viwb_globals *g = &p->globals;
    G = Eigen::Vector3d(g->G[0], g->G[1], g->G[2]);
    Eigen::Matrix2d si; si << g->vis_sqrt_info[0], g->vis_sqrt_info[1], g->vis_sqrt_info[2], g->vis_sqrt_info[3];
    ProjectionTwoFrameOneCamFactor::sqrt_info = si; ProjectionTwoFrameTwoCamFactor::sqrt_info = si; ProjectionOneFrameTwoCamFactor::sqrt_info = si;
    PITCH_N_INV = g->plane_sqrt_info[0]; ROLL_N_INV = g->plane_sqrt_info[1]; ZPW_N_INV = g->plane_sqrt_info[2];
    std::vector<double> st(state_in, state_in + VIWB_STATE_FIXED + p->num_landmarks);
    auto blk = [&](int b) { return st.data() + viwb_block_offset(b); };
    auto lm = [&](int k) { return st.data() + VIWB_STATE_FIXED + k; };
    ceres::LossFunction *loss = new ceres::HuberLoss(g->huber_delta);
    MarginalizationInfo *info = new MarginalizationInfo();
    MarginalizationInfo *last = nullptr; std::vector<double *> x0;
    std::vector<IntegrationBase *> imu_pre; std::vector<WheelIntegrationBase *> wheel_pre;
    const int WS = 10;
    if (p->prior && p->prior->valid) {
        last = info_from_prior(p->prior, x0);
        std::vector<double *> blocks; std::vector<int> drop;
        for (int k = 0; k < p->prior->num_blocks; k++) {
            const int b = p->prior->block_id[k];
            blocks.push_back(blk(b));
            if (margin_flag == VIWB_MARGIN_OLD ? (b == 0 || b == 11) : (b == WS - 1)) drop.push_back(k);
        }
        info->addResidualBlockInfo(new ResidualBlockInfo(new MarginalizationFactor(last), NULL, blocks, drop));
    }
    if (margin_flag == VIWB_MARGIN_OLD) {
        for (int f = 0; f < p->num_imu; f++) if (p->imu_frame_i[f] == 0 && p->imu_frame_j[f] == 1) {
            IntegrationBase *pre = new IntegrationBase(Eigen::Vector3d::Zero(), Eigen::Vector3d::Zero(), Eigen::Vector3d::Zero(), Eigen::Vector3d::Zero());
            fill_pre_integration(*pre, p->imu_data + (size_t)f * VIWB_IMU_DOUBLES); imu_pre.push_back(pre);
            if (pre->sum_dt < 10.0) info->addResidualBlockInfo(new ResidualBlockInfo(new IMUFactor(pre), NULL, std::vector<double *>{blk(0), blk(11), blk(1), blk(12)}, std::vector<int>{0, 1}));
        }
        for (int f = 0; f < p->num_wheel; f++) if (p->wheel_frame_i[f] == 0 && p->wheel_frame_j[f] == 1) {
            const double *c = p->wheel_data + (size_t)f * VIWB_WHEEL_DOUBLES;
            WheelIntegrationBase *pre = new WheelIntegrationBase(Eigen::Vector3d(c[65], c[66], c[67]), Eigen::Vector3d(c[68], c[69], c[70]), c[61], c[62], c[63], c[64]);
            pre->delta_p = Eigen::Vector3d(c[0], c[1], c[2]); pre->delta_q = Eigen::Quaterniond(c[6], c[3], c[4], c[5]);
            for (int i = 0; i < 6; i++) for (int j = 0; j < 3; j++) pre->jacobian(i, j) = c[7 + 3 * i + j];
            for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) pre->covariance(i, j) = c[25 + 6 * i + j];
            pre->vel_1 = Eigen::Vector3d(c[71], c[72], c[73]); pre->gyr_1 = Eigen::Vector3d(c[74], c[75], c[76]); pre->sum_dt = c[77];
            wheel_pre.push_back(pre);
            if (pre->sum_dt < 10.0) info->addResidualBlockInfo(new ResidualBlockInfo(new WheelFactor(pre), NULL, std::vector<double *>{blk(0), blk(1), blk(24), blk(27), blk(28), blk(29), blk(31)}, std::vector<int>{0}));
        }
        for (int f = 0; f < p->num_plane; f++) if (p->plane_frame[f] == 0)
            info->addResidualBlockInfo(new ResidualBlockInfo(new PlaneFactor(), NULL, std::vector<double *>{blk(0), blk(24), blk(25), blk(26)}, std::vector<int>{0}));
        for (int f = 0; f < p->num_vis; f++) {
            if (p->vis_frame_i[f] != 0) continue;
            const double *c = p->vis_obs + (size_t)f * VIWB_VIS_OBS_DOUBLES;
            const Eigen::Vector3d pi(c[0], c[1], c[2]), pj(c[3], c[4], c[5]); const Eigen::Vector2d vi(c[6], c[7]), vj(c[8], c[9]);
            const int j = p->vis_frame_j[f]; double *l = lm(p->vis_landmark[f]);
            if (p->vis_type[f] == VIWB_F_PROJ_2F1C)
                info->addResidualBlockInfo(new ResidualBlockInfo(new ProjectionTwoFrameOneCamFactor(pi, pj, vi, vj, c[10], c[11]), loss, std::vector<double *>{blk(0), blk(j), blk(22), l, blk(30)}, std::vector<int>{0, 3}));
            else if (p->vis_type[f] == VIWB_F_PROJ_2F2C)
                info->addResidualBlockInfo(new ResidualBlockInfo(new ProjectionTwoFrameTwoCamFactor(pi, pj, vi, vj, c[10], c[11]), loss, std::vector<double *>{blk(0), blk(j), blk(22), blk(23), l, blk(30)}, std::vector<int>{0, 4}));
            else
                info->addResidualBlockInfo(new ResidualBlockInfo(new ProjectionOneFrameTwoCamFactor(pi, pj, vi, vj, c[10], c[11]), loss, std::vector<double *>{blk(22), blk(23), l, blk(30)}, std::vector<int>{2}));
        }
    }
    int rc = 0;
    if (info->factors.empty()) rc = 3;
    else {
        info->preMarginalize();
        info->marginalize();
        std::unordered_map<long, double *> addr_shift;
        for (auto &it : info->parameter_block_idx) addr_shift[it.first] = reinterpret_cast<double *>(it.first);
        std::vector<double *> kept = info->getParameterBlocks(addr_shift);
        mn[0] = info->m; mn[1] = info->n; mn[2] = (int)kept.size();
        for (size_t k = 0; k < kept.size(); k++) {
            const long off = kept[k] - st.data();
            int b = -1;
            for (int q = 0; q < VIWB_NUM_FIXED_BLOCKS; q++) if (viwb_block_offset(q) == off) b = q;
            block_id[k] = b; block_idx[k] = info->keep_block_idx[k] - info->m;
        }
        for (int i = 0; i < info->n; i++) { r[i] = info->linearized_residuals(i); for (int j = 0; j < info->n; j++) J[(size_t)i * info->n + j] = info->linearized_jacobians(i, j); }
    }
    delete info;                       // deletes the factors and the cost functions it was given
    if (last) { for (double *d : x0) delete[] d; /* `last` itself went with its MarginalizationFactor's owner; the info object is ours */ delete last; }
    for (auto q : imu_pre) delete q;
    for (auto q : wheel_pre) delete q;
    delete loss;
    return rc;
}

// ---------------------------------------------------------------------------------------------------- self-test hooks of the stand-ins
// tests/test_refshim.py checks oracle/refshim/mini_eigen.h and mini_sophus.h against numpy / scipy on their own, so that agreement
// between the compiled reference code and the oracle cannot be explained by a shared mistake in the matrix header.
extern "C" void ref_selftest_linalg15(const double *A_rm, double *inv_rm, double *llt_rm, double *prod_rm) {
    Eigen::Matrix<double, 15, 15> A;
    for (int i = 0; i < 15; i++) for (int j = 0; j < 15; j++) A(i, j) = A_rm[15 * i + j];
    const Eigen::Matrix<double, 15, 15> inv = A.inverse();
    const Eigen::Matrix<double, 15, 15> L = Eigen::LLT<Eigen::Matrix<double, 15, 15>>(A).matrixL();
    Eigen::MatrixXd F = Eigen::MatrixXd::Zero(15, 15);
    F.block<3, 3>(3, 6) = A.block<3, 3>(0, 0); F.block<3, 3>(0, 0) = Eigen::Matrix3d::Identity() * 2.0;
    const Eigen::Matrix<double, 15, 15> P = F * A * F.transpose() + A.transpose();
    for (int i = 0; i < 15; i++) for (int j = 0; j < 15; j++) { inv_rm[15 * i + j] = inv(i, j); llt_rm[15 * i + j] = L(i, j); prod_rm[15 * i + j] = P(i, j); }
}
extern "C" void ref_selftest_eig(int n, const double *A_rm, double *w, double *V_rm) {
    Eigen::MatrixXd A(n, n);
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) A(i, j) = A_rm[(size_t)n * i + j];
    Eigen::SelfAdjointEigenSolver<Eigen::MatrixXd> s(A);
    for (int i = 0; i < n; i++) { w[i] = s.eigenvalues()(i); for (int j = 0; j < n; j++) V_rm[(size_t)n * i + j] = s.eigenvectors()(i, j); }
}
// q = [x, y, z, w]; out: R(q) row-major (9), q^-1 (4), q * v (3), quaternion of R (4), (q * p) product (4), SO3::exp(v) quaternion (4), SO3(q).log() (3)
extern "C" void ref_selftest_rotations(const double *q_, const double *p_, const double *v_, double *out) {
    const Eigen::Quaterniond q(q_[3], q_[0], q_[1], q_[2]), p(p_[3], p_[0], p_[1], p_[2]);
    const Eigen::Vector3d v(v_[0], v_[1], v_[2]);
    const Eigen::Matrix3d R = q.toRotationMatrix();
    int k = 0;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) out[k++] = R(i, j);
    const Eigen::Quaterniond qi = q.inverse(); out[k++] = qi.x(); out[k++] = qi.y(); out[k++] = qi.z(); out[k++] = qi.w();
    const Eigen::Vector3d qv = q * v; for (int i = 0; i < 3; i++) out[k++] = qv(i);
    const Eigen::Quaterniond qr(R); out[k++] = qr.x(); out[k++] = qr.y(); out[k++] = qr.z(); out[k++] = qr.w();
    const Eigen::Quaterniond qp = q * p; out[k++] = qp.x(); out[k++] = qp.y(); out[k++] = qp.z(); out[k++] = qp.w();
    const Eigen::Quaterniond e = Sophus::SO3d::exp(v).unit_quaternion(); out[k++] = e.x(); out[k++] = e.y(); out[k++] = e.z(); out[k++] = e.w();
    const Eigen::Vector3d l = Sophus::SO3d(q).log(); for (int i = 0; i < 3; i++) out[k++] = l(i);
}

// ---------------------------------------------------------------------------------------------------- FeatureManager (estimator/feature_manager.cpp)
static void state_poses(const double *st, Eigen::Vector3d *Ps, Eigen::Matrix3d *Rs, Eigen::Vector3d *tic, Eigen::Matrix3d *ric) {
    for (int i = 0; i <= 10; i++) { Ps[i] = Eigen::Vector3d(st[7 * i], st[7 * i + 1], st[7 * i + 2]); Rs[i] = Eigen::Quaterniond(st[7 * i + 6], st[7 * i + 3], st[7 * i + 4], st[7 * i + 5]).toRotationMatrix(); }
    for (int c = 0; c < 2; c++) { const double *e = st + 176 + 7 * c; tic[c] = Eigen::Vector3d(e[0], e[1], e[2]); ric[c] = Eigen::Quaterniond(e[6], e[3], e[4], e[5]).toRotationMatrix(); }
}
// FeatureManager::triangulate on features built from (first observation, second observation): same contract as vo_triangulate
extern "C" int ref_triangulate(const double *state, int n, const int32_t *stereo, const int32_t *frame, const double *pt0, const double *pt1, double init_depth, double *depth) {
    Eigen::Vector3d Ps[11], tic[2]; Eigen::Matrix3d Rs[11], ric[2];
    state_poses(state, Ps, Rs, tic, ric);
    INIT_DEPTH = init_depth; STEREO = 1;
    FeatureManager fm(Rs);
    fm.setRic(ric);
    for (int k = 0; k < n; k++) {
        Eigen::Matrix<double, 7, 1> a, b;
        a << pt0[2 * k], pt0[2 * k + 1], 1.0, 0.0, 0.0, 0.0, 0.0;
        b << pt1[2 * k], pt1[2 * k + 1], 1.0, 0.0, 0.0, 0.0, 0.0;
        FeaturePerId f(k, frame[k]);
        f.feature_per_frame.push_back(FeaturePerFrame(a, 0.0));
        if (stereo[k]) f.feature_per_frame[0].rightObservation(b);
        else f.feature_per_frame.push_back(FeaturePerFrame(b, 0.0));
        fm.feature.push_back(f);
    }
    fm.triangulate(10, Ps, Rs, tic, ric);
    int k = 0;
    for (auto &f : fm.feature) depth[k++] = f.estimated_depth;
    return 0;
}
// FeatureManager::removeBackShiftDepth on features hosted in frame 0 with three observations each: same contract as vo_shift_depth
extern "C" int ref_shift_depth(int n, const double *uv, const double *depth_in, const double *marg_R, const double *marg_P, const double *new_R, const double *new_P,
                               double init_depth, double *depth_out) {
    Eigen::Matrix3d Rs[11], mR, nR; Eigen::Vector3d mP(marg_P[0], marg_P[1], marg_P[2]), nP(new_P[0], new_P[1], new_P[2]);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { mR(i, j) = marg_R[3 * i + j]; nR(i, j) = new_R[3 * i + j]; }
    INIT_DEPTH = init_depth;
    FeatureManager fm(Rs);
    for (int k = 0; k < n; k++) {
        Eigen::Matrix<double, 7, 1> a; a << uv[3 * k], uv[3 * k + 1], uv[3 * k + 2], 0.0, 0.0, 0.0, 0.0;
        FeaturePerId f(k, 0);
        for (int o = 0; o < 3; o++) f.feature_per_frame.push_back(FeaturePerFrame(a, 0.0));
        f.estimated_depth = depth_in[k];
        fm.feature.push_back(f);
    }
    fm.removeBackShiftDepth(mR, mP, nR, nP);
    if ((int)fm.feature.size() != n) return 1;
    int k = 0;
    for (auto &f : fm.feature) depth_out[k++] = f.estimated_depth;
    return 0;
}

// ---------------------------------------------------------------------------------------------------- camera model (camera_models/src/camera_models/PinholeCamera.cc)
// FeatureTracker::undistortedPts (featureTracker/feature_tracker.cpp:606-617) is three lines around PinholeCamera::liftProjective; the
// lift and the distortion model are the reference's.  cam = {fx, fy, cx, cy, k1, k2, p1, p2}; pts / un: n x 2 floats (cv::Point2f).
extern "C" void ref_undistorted_pts(const double *cam, int width, int height, int n, const float *pts, float *un) {
    camodocal::PinholeCamera c("cam", width, height, cam[4], cam[5], cam[6], cam[7], cam[0], cam[1], cam[2], cam[3]);
    for (int i = 0; i < n; i++) {
        Eigen::Vector2d a(pts[2 * i], pts[2 * i + 1]);
        Eigen::Vector3d b;
        c.liftProjective(a, b);
        un[2 * i] = (float)(b.x() / b.z()); un[2 * i + 1] = (float)(b.y() / b.z());
    }
}
// PinholeCamera::spaceToPlane (projection with distortion): P n x 3 doubles -> pixel n x 2 doubles (FeatureTracker::setPrediction, :729)
extern "C" void ref_space_to_plane(const double *cam, int width, int height, int n, const double *P, double *uv) {
    camodocal::PinholeCamera c("cam", width, height, cam[4], cam[5], cam[6], cam[7], cam[0], cam[1], cam[2], cam[3]);
    for (int i = 0; i < n; i++) { Eigen::Vector2d p; c.spaceToPlane(Eigen::Vector3d(P[3 * i], P[3 * i + 1], P[3 * i + 2]), p); uv[2 * i] = p.x(); uv[2 * i + 1] = p.y(); }
}
