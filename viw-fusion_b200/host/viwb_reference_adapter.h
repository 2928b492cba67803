// viwb_reference_adapter.h -- lets the shim lower the REFERENCE'S OWN factor classes, so that estimator.cpp compiles and runs unmodified:
// only the include path changes (viw-fusion_b200/host before the real Ceres), no `#ifdef` in the estimator sources.
//
// Include this header in ONE translation unit of vins_estimator AFTER the reference's factor headers (it names their classes):
//     #include "factor/imu_factor.h" ... "factor/marginalization_factor.h", "factor/pose_subset_parameterization.h", ...
//     #include "viwb_reference_adapter.h"
// and call viwb_shim::install_reference_adapter() once (e.g. from Estimator::setParameter()).  ceres::Solve then recognises
//   ProjectionTwoFrameOneCamFactor / ...TwoFrameTwoCamFactor / ...OneFrameTwoCamFactor   (projection*Factor.h: pts_i, pts_j, velocity_i/j, td_i/j)
//   IMUFactor, WheelFactor (pre_integration members, factor/integration_base.h:197-214, wheel_integration_base.h:220-243), PlaneFactor,
//   MarginalizationFactor (marginalization_info: linearized_jacobians / residuals, keep_block_*; factor/marginalization_factor.h:62-80),
//   PoseLocalParameterization, PoseSubsetParameterization, OrientationSubsetParameterization (their constancy masks),
// by dynamic_cast, and builds the constant records of include/viwb.h from their public members.  The sqrt-information statics
// (ProjectionTwoFrameOneCamFactor::sqrt_info, estimator.cpp:157-159), G and the plane weights are read into viwb_shim::globals().
// Needs: the subset parameterizations' `constancy_mask_` is private in the reference -- compile this one translation unit with
// -fno-access-control, or add a one-line accessor there.
#pragma once
#include <deque>
#include <vector>

namespace viwb_shim {

// exact block ids of the priors produced by factor/marginalization_factor_device.cpp (when that translation unit replaces the reference's
// marginalization_factor.cpp); weak: absent when the stock CPU marginalization is linked
std::map<const void *, std::vector<int>> &device_prior_ids() __attribute__((weak));

struct ReferenceScratch {
    std::deque<std::vector<double>> records;          // constant records built for the current Solve
    std::deque<viwb_prior> priors;
    std::deque<std::vector<double>> prior_storage;
};
inline ReferenceScratch &reference_scratch() { static thread_local ReferenceScratch s; return s; }

template <class F> inline const double *reference_visual_record(const F *f) {
    std::vector<double> r(VIWB_VIS_OBS_DOUBLES);
    for (int k = 0; k < 3; k++) { r[k] = f->pts_i(k); r[3 + k] = f->pts_j(k); }
    r[6] = f->velocity_i(0); r[7] = f->velocity_i(1); r[8] = f->velocity_j(0); r[9] = f->velocity_j(1); r[10] = f->td_i; r[11] = f->td_j;
    reference_scratch().records.push_back(r);
    return reference_scratch().records.back().data();
}

inline bool reference_cost_adapter(const ceres::CostFunction *c, Lowered *out) {
    ReferenceScratch &S = reference_scratch();
    if (auto *f = dynamic_cast<const ProjectionTwoFrameOneCamFactor *>(c)) { out->type = VIWB_F_PROJ_2F1C; out->record = reference_visual_record(f); return true; }
    if (auto *f = dynamic_cast<const ProjectionTwoFrameTwoCamFactor *>(c)) { out->type = VIWB_F_PROJ_2F2C; out->record = reference_visual_record(f); return true; }
    if (auto *f = dynamic_cast<const ProjectionOneFrameTwoCamFactor *>(c)) { out->type = VIWB_F_PROJ_1F2C; out->record = reference_visual_record(f); return true; }
    if (auto *f = dynamic_cast<const IMUFactor *>(c)) {
        const IntegrationBase *p = f->pre_integration;
        std::vector<double> r(VIWB_IMU_DOUBLES);
        r[0] = p->sum_dt;
        for (int k = 0; k < 3; k++) { r[1 + k] = p->delta_p(k); r[8 + k] = p->delta_v(k); r[11 + k] = p->linearized_ba(k); r[14 + k] = p->linearized_bg(k); }
        r[4] = p->delta_q.x(); r[5] = p->delta_q.y(); r[6] = p->delta_q.z(); r[7] = p->delta_q.w();
        const int br[5] = {O_P, O_P, O_R, O_V, O_V}, bc[5] = {O_BA, O_BG, O_BG, O_BA, O_BG};
        for (int k = 0; k < 5; k++) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r[17 + 9 * k + 3 * i + j] = p->jacobian(br[k] + i, bc[k] + j);
        for (int i = 0; i < 15; i++) for (int j = 0; j < 15; j++) r[62 + 15 * i + j] = p->covariance(i, j);
        S.records.push_back(r); out->type = VIWB_F_IMU; out->record = S.records.back().data(); return true;
    }
    if (auto *f = dynamic_cast<const WheelFactor *>(c)) {
        const WheelIntegrationBase *p = f->pre_integration;
        std::vector<double> r(VIWB_WHEEL_DOUBLES);
        for (int k = 0; k < 3; k++) { r[k] = p->delta_p(k); r[65 + k] = p->linearized_vel(k); r[68 + k] = p->linearized_gyr(k); r[71 + k] = p->vel_1(k); r[74 + k] = p->gyr_1(k); }
        r[3] = p->delta_q.x(); r[4] = p->delta_q.y(); r[5] = p->delta_q.z(); r[6] = p->delta_q.w();
        for (int i = 0; i < 6; i++) for (int j = 0; j < 3; j++) r[7 + 3 * i + j] = p->jacobian(i, j);
        for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) r[25 + 6 * i + j] = p->covariance(i, j);
        r[61] = p->linearized_sx; r[62] = p->linearized_sy; r[63] = p->linearized_sw; r[64] = p->linearized_td; r[77] = p->sum_dt;
        S.records.push_back(r); out->type = VIWB_F_WHEEL; out->record = S.records.back().data(); return true;
    }
    if (dynamic_cast<const PlaneFactor *>(c)) { out->type = VIWB_F_PLANE; out->record = nullptr; return true; }
    if (auto *f = dynamic_cast<const MarginalizationFactor *>(c)) {
        const MarginalizationInfo *info = f->marginalization_info;
        const int n = info->n, nb = (int)info->keep_block_size.size();
        viwb_prior pr; std::memset(&pr, 0, sizeof pr);
        pr.valid = info->valid ? 1 : 0; pr.n = n; pr.num_blocks = nb;
        S.prior_storage.emplace_back(VIWB_STATE_FIXED, 0.0); std::vector<double> &x0 = S.prior_storage.back();
        S.prior_storage.emplace_back((size_t)n * n); std::vector<double> &J = S.prior_storage.back();
        S.prior_storage.emplace_back((size_t)n); std::vector<double> &r = S.prior_storage.back();
        // block ids: exact when the prior came from the device marginalization (in column order, like keep_block_* after getParameterBlocks), else
        // provisional -- any fixed block of the right size (Solve() re-maps the kept blocks by the addresses of the residual's parameter blocks)
        const std::vector<int> *exact = nullptr;
        if (&device_prior_ids != nullptr) { auto it = device_prior_ids().find(info); if (it != device_prior_ids().end() && (int)it->second.size() == nb) exact = &it->second; }
        bool used[VIWB_NUM_FIXED_BLOCKS] = {false};
        for (int i = 0; i < nb; i++) {
            const int size = info->keep_block_size[i];
            int slot = -1;
            if (exact) { slot = (*exact)[i]; if (slot < 0 || slot >= VIWB_NUM_FIXED_BLOCKS || viwb_block_size(slot) != size) return false; }
            for (int b = 0; b < VIWB_NUM_FIXED_BLOCKS && slot < 0; b++) if (!used[b] && viwb_block_size(b) == size) slot = b;
            if (slot < 0) return false;
            used[slot] = true;
            pr.block_id[i] = slot; pr.block_idx[i] = info->keep_block_idx[i] - info->m;
            std::memcpy(x0.data() + viwb_block_offset(slot), info->keep_block_data[i], sizeof(double) * size);
        }
        for (int i = 0; i < n; i++) { r[i] = info->linearized_residuals(i); for (int j = 0; j < n; j++) J[(size_t)i * n + j] = info->linearized_jacobians(i, j); }
        pr.x0 = x0.data(); pr.J = J.data(); pr.r = r.data();
        S.priors.push_back(pr);
        out->type = -2; out->prior = &S.priors.back(); out->prior_ids_exact = exact != nullptr; return true;
    }
    return false;
}

inline int reference_manifold_adapter(const ceres::LocalParameterization *p) {
    if (auto *s = dynamic_cast<const PoseSubsetParameterization *>(p)) { int m = 0; for (size_t i = 0; i < s->constancy_mask_.size() && i < 6; i++) if (s->constancy_mask_[i]) m |= 1 << i; return m; }
    if (auto *s = dynamic_cast<const OrientationSubsetParameterization *>(p)) { int m = 0; for (size_t i = 0; i < s->constancy_mask_.size() && i < 3; i++) if (s->constancy_mask_[i]) m |= 1 << i; return m; }
    if (dynamic_cast<const PoseLocalParameterization *>(p)) return 0;
    return -1;
}

inline void reference_solve_begin() {
    ReferenceScratch &S = reference_scratch();
    S.records.clear(); S.priors.clear(); S.prior_storage.clear();
    viwb_globals &g = globals();                                   // the reference's static / global weights, as they are right now
    for (int i = 0; i < 3; i++) g.G[i] = G(i);
    const Eigen::Matrix2d &si = ProjectionTwoFrameOneCamFactor::sqrt_info;
    g.vis_sqrt_info[0] = si(0, 0); g.vis_sqrt_info[1] = si(0, 1); g.vis_sqrt_info[2] = si(1, 0); g.vis_sqrt_info[3] = si(1, 1);
    g.plane_sqrt_info[0] = PITCH_N_INV; g.plane_sqrt_info[1] = ROLL_N_INV; g.plane_sqrt_info[2] = ZPW_N_INV;
}

inline void install_reference_adapter() {
    cost_adapter() = reference_cost_adapter;
    manifold_adapter() = reference_manifold_adapter;
    solve_begin() = reference_solve_begin;
}

}  // namespace viwb_shim
