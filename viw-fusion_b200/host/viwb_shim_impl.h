// viwb_shim_impl.h -- lowering of a shim ceres::Problem to include/viwb.h tables and the call into libviwb.so.
// Included at the end of ceres/ceres.h.  One process-wide GPU context per calling thread (a context is single-threaded:
// optimization() runs on processThread, trackImage on sync_thread; rosNodeTestWheel.cpp:84-145).
#pragma once
#include <algorithm>
#include <cstdio>

namespace viwb_shim {

inline viwb_context *context() {
    static thread_local viwb_context *ctx = nullptr;
    if (!ctx && viwb_create(0, &ctx) != VIWB_OK) { ctx = nullptr; std::fprintf(stderr, "viwb: no CUDA device -- there is no CPU fallback\n"); }
    return ctx;
}
// globals the reference keeps in static members / global variables (set once from readParameters):
//   ProjectionTwoFrameOneCamFactor::sqrt_info etc. (estimator.cpp:157-159), G (parameters.cpp:32,149), plane weights (plane_factor.h:52)
inline viwb_globals &globals() { static viwb_globals g = [] { viwb_globals t; viwb_default_globals(&t); return t; }(); return g; }

}  // namespace viwb_shim

namespace ceres {

inline void HuberLoss::Evaluate(double s, double rho[3]) const {
    // ceres::HuberLoss (API fidelity for ResidualBlockInfo::Evaluate, marginalization_factor.cpp:53; the GPU path applies it in lin_vis)
    if (s > b_) { const double r = std::sqrt(s); rho[0] = 2.0 * a_ * r - b_; rho[1] = a_ / r; rho[2] = -rho[1] / (2.0 * s); }
    else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
}

inline std::string Solver::Summary::BriefReport() const {
    char buf[256];
    std::snprintf(buf, sizeof buf, "viwb(B200) iterations: %d, initial cost: %e, final cost: %e, termination: %s", (int)iterations.size(), initial_cost,
                  final_cost, termination_type == CONVERGENCE ? "CONVERGENCE" : termination_type == NO_CONVERGENCE ? "NO_CONVERGENCE" : "FAILURE");
    return buf;
}

inline Problem::~Problem() {
    std::set<CostFunction *> cs; std::set<LossFunction *> ls; std::set<LocalParameterization *> ps;
    for (auto &r : residuals_) { cs.insert(r.cost); if (r.loss) ls.insert(r.loss); }
    for (auto &b : blocks_) if (b.second.lp) ps.insert(b.second.lp);
    for (auto *c : cs) delete c;
    for (auto *l : ls) delete l;
    for (auto *p : ps) delete p;
}
inline void Problem::AddParameterBlock(double *values, int size, LocalParameterization *lp) {
    Block &b = blocks_[values];
    if (b.size == 0) b.order = (int)blocks_.size() - 1;
    b.size = size;
    if (lp) b.lp = lp;
}
inline void *Problem::AddResidualBlock(CostFunction *cost, LossFunction *loss, const std::vector<double *> &blocks) {
    const std::vector<int32_t> &sizes = cost->parameter_block_sizes();
    for (size_t i = 0; i < blocks.size(); i++) { Block &b = blocks_[blocks[i]]; if (b.size == 0) { b.order = (int)blocks_.size() - 1; b.size = i < sizes.size() ? sizes[i] : 0; } }
    residuals_.push_back(Residual{cost, loss, blocks});
    return &residuals_.back();
}

// The lowering: classify every parameter block by the factor slots it appears in, order window poses / speed-biases by
// address (para_Pose[i], para_SpeedBias[i] are rows of one array, estimator.h:191-192), landmarks by first appearance
// (= feature_index, estimator.cpp:1587-1593), then emit the tables of include/viwb.h.
inline void Solve(const Solver::Options &options, Problem *problem, Solver::Summary *summary) {
    summary->iterations.clear(); summary->termination_type = FAILURE;
    if (options.linear_solver_type != DENSE_SCHUR || options.trust_region_strategy_type != DOGLEG) { summary->message = "viwb supports DENSE_SCHUR + DOGLEG (the reference configuration) only"; return; }
    viwb_context *ctx = viwb_shim::context();
    if (!ctx) { summary->message = "no CUDA device"; return; }
    std::set<double *> poses, sbs; std::vector<double *> landmarks; std::map<double *, int> lm_index;
    double *ex0 = nullptr, *ex1 = nullptr, *exw = nullptr, *sx = nullptr, *sy = nullptr, *sw = nullptr, *tdw = nullptr, *td = nullptr, *plane_r = nullptr, *plane_z = nullptr;
    const viwb_prior *prior = nullptr; const Problem::Residual *prior_res = nullptr;
    double huber = -1.0;
    if (viwb_shim::solve_begin()) viwb_shim::solve_begin()();
    std::vector<viwb_shim::Lowered> low(problem->residuals_.size());
    for (size_t k = 0; k < problem->residuals_.size(); k++) low[k] = viwb_shim::lower(problem->residuals_[k].cost);
    size_t rk = 0;
    for (auto &r : problem->residuals_) {
        const int t = low[rk].type;
        const std::vector<double *> &b = r.blocks;
        if (r.loss) huber = viwb_shim::huber_delta(r.loss);
        switch (t) {
        case VIWB_F_PROJ_2F1C: poses.insert(b[0]); poses.insert(b[1]); ex0 = b[2]; if (!lm_index.count(b[3])) { lm_index[b[3]] = (int)landmarks.size(); landmarks.push_back(b[3]); } td = b[4]; break;
        case VIWB_F_PROJ_2F2C: poses.insert(b[0]); poses.insert(b[1]); ex0 = b[2]; ex1 = b[3]; if (!lm_index.count(b[4])) { lm_index[b[4]] = (int)landmarks.size(); landmarks.push_back(b[4]); } td = b[5]; break;
        case VIWB_F_PROJ_1F2C: ex0 = b[0]; ex1 = b[1]; if (!lm_index.count(b[2])) { lm_index[b[2]] = (int)landmarks.size(); landmarks.push_back(b[2]); } td = b[3]; break;
        case VIWB_F_IMU: poses.insert(b[0]); sbs.insert(b[1]); poses.insert(b[2]); sbs.insert(b[3]); break;
        case VIWB_F_WHEEL: poses.insert(b[0]); poses.insert(b[1]); exw = b[2]; sx = b[3]; sy = b[4]; sw = b[5]; tdw = b[6]; break;
        case VIWB_F_PLANE: poses.insert(b[0]); exw = b[1]; plane_r = b[2]; plane_z = b[3]; break;
        case -2: prior = low[rk].prior; prior_res = &r; break;
        default: summary->message = "unknown CostFunction: only the hot-path factor classes can be lowered to the GPU"; return;
        }
        rk++;
    }
    // Window poses: para_Pose[i] are the rows of ONE array (estimator.h:191), so they sit on a lattice of 7 doubles around any pose a factor names
    // (or around the first size-7 block registered, estimator.cpp:1394-1397, when no factor names one yet).  Every other size-7 block is an
    // extrinsic: those no factor names are told apart by registration order (camera extrinsics first, estimator.cpp:1405-1437, then the wheel
    // extrinsic, :1439-1470) and by para_Ex_Pose[1] following para_Ex_Pose[0] in memory (estimator.h:193).
    {
        std::vector<std::pair<int, double *>> sevens;        // (registration order, address)
        for (auto &kv : problem->blocks_) if (kv.second.size == 7) sevens.push_back({kv.second.order, kv.first});
        std::sort(sevens.begin(), sevens.end());
        double *anchor = !poses.empty() ? *poses.begin() : nullptr;
        for (auto &sv : sevens) if (!anchor && sv.second != ex0 && sv.second != ex1 && sv.second != exw) anchor = sv.second;
        for (auto &sv : sevens) {
            double *q = sv.second;
            if (q == ex0 || q == ex1 || q == exw || poses.count(q)) continue;
            const std::ptrdiff_t d = q - anchor;
            if (anchor && d % 7 == 0 && d / 7 >= -(std::ptrdiff_t)VIWB_WINDOW_SIZE && d / 7 <= (std::ptrdiff_t)VIWB_WINDOW_SIZE) { poses.insert(q); continue; }
            if (!ex0) ex0 = q; else if (!ex1 && q == ex0 + 7) ex1 = q; else if (!exw) exw = q;
            else { summary->message = "a size-7 parameter block is neither a window pose nor an extrinsic"; return; }
        }
    }
    for (auto &kv : problem->blocks_) if (kv.second.size == 9) sbs.insert(kv.first);
    std::vector<double *> pose_v(poses.begin(), poses.end()), sb_v(sbs.begin(), sbs.end());      // std::set<double*> iterates in address order
    if (pose_v.empty() || pose_v.size() > VIWB_NUM_FRAMES || sb_v.size() > VIWB_NUM_FRAMES || landmarks.size() > VIWB_MAX_LANDMARKS) { summary->message = "window too large (or empty)"; return; }
    for (size_t i = 0; i < pose_v.size(); i++) if (pose_v[i] - pose_v[0] != 7 * (std::ptrdiff_t)i) { summary->message = "window poses are not consecutive rows of one array"; return; }
    for (size_t i = 0; i < sb_v.size(); i++) if (sb_v[i] - sb_v[0] != 9 * (std::ptrdiff_t)i) { summary->message = "speed-bias blocks are not consecutive rows of one array"; return; }
    std::map<double *, int> id;
    for (size_t i = 0; i < pose_v.size(); i++) id[pose_v[i]] = VIWB_BLK_POSE0 + (int)i;
    for (size_t i = 0; i < sb_v.size(); i++) id[sb_v[i]] = VIWB_BLK_SPEEDBIAS0 + (int)i;
    auto put = [&](double *p, int b) { if (p) id[p] = b; };
    put(ex0, VIWB_BLK_EX_POSE0); put(ex1, VIWB_BLK_EX_POSE1); put(exw, VIWB_BLK_EX_WHEEL); put(plane_r, VIWB_BLK_PLANE_R); put(plane_z, VIWB_BLK_PLANE_Z);
    put(sx, VIWB_BLK_SX); put(sy, VIWB_BLK_SY); put(sw, VIWB_BLK_SW); put(td, VIWB_BLK_TD); put(tdw, VIWB_BLK_TD_WHEEL);
    // blocks that are registered but appear in no factor (td_wheel without wheel factors, ...) are classified by size/order
    for (auto &kv : problem->blocks_) if (!id.count(kv.first) && !lm_index.count(kv.first)) {
        if (kv.second.size == 4) put(kv.first, VIWB_BLK_PLANE_R);
    }
    viwb_problem pb; std::memset(&pb, 0, sizeof pb);
    pb.frame_count = (int)pose_v.size() - 1; pb.num_landmarks = (int)landmarks.size();
    pb.globals = viwb_shim::globals(); if (huber > 0) pb.globals.huber_delta = huber;
    std::vector<double> state(VIWB_STATE_FIXED + landmarks.size(), 0.0);
    for (auto &kv : id) {
        const Problem::Block &blk = problem->blocks_[kv.first];
        pb.block_flags[kv.second] = VIWB_BLOCK_PRESENT | (blk.constant ? VIWB_BLOCK_CONSTANT : 0u);
        pb.subset_mask[kv.second] = (uint8_t)viwb_shim::subset_mask(blk.lp);
        std::memcpy(state.data() + viwb_block_offset(kv.second), kv.first, sizeof(double) * viwb_block_size(kv.second));
    }
    for (size_t k = 0; k < landmarks.size(); k++) state[VIWB_STATE_FIXED + k] = landmarks[k][0];
    std::vector<int32_t> vt, vl, vi, vj, ii, ij, wi, wj, pf; std::vector<double> vobs, idata, wdata;
    viwb_prior pr_local; std::vector<double> pr_x0;
    rk = 0;
    for (auto &r : problem->residuals_) {
        const int t = low[rk].type; const std::vector<double *> &b = r.blocks; const double *rec = low[rk].record; rk++;
        if (t >= 0 && t <= VIWB_F_PROJ_1F2C) {
            const int li = t == VIWB_F_PROJ_2F1C ? 3 : t == VIWB_F_PROJ_2F2C ? 4 : 2;
            vt.push_back(t); vl.push_back(lm_index[b[li]]);
            if (t == VIWB_F_PROJ_1F2C) { vi.push_back(0); vj.push_back(0); }      // frame indices are irrelevant for the same-frame stereo factor
            else { vi.push_back(id[b[0]]); vj.push_back(id[b[1]]); }
            vobs.insert(vobs.end(), rec, rec + VIWB_VIS_OBS_DOUBLES);
        } else if (t == VIWB_F_IMU) { ii.push_back(id[b[0]]); ij.push_back(id[b[2]]); idata.insert(idata.end(), rec, rec + VIWB_IMU_DOUBLES); }
        else if (t == VIWB_F_WHEEL) { wi.push_back(id[b[0]]); wj.push_back(id[b[1]]); wdata.insert(wdata.end(), rec, rec + VIWB_WHEEL_DOUBLES); }
        else if (t == VIWB_F_PLANE) pf.push_back(id[b[0]]);
    }
    // the 1F2C factor belongs to its landmark's host frame: patch its frame indices from the landmark's other factors
    { std::vector<int> host(landmarks.size(), -1);
      for (size_t f = 0; f < vt.size(); f++) if (vt[f] != VIWB_F_PROJ_1F2C) host[vl[f]] = vi[f];
      for (size_t f = 0; f < vt.size(); f++) if (vt[f] == VIWB_F_PROJ_1F2C && host[vl[f]] >= 0) { vi[f] = host[vl[f]]; vj[f] = host[vl[f]]; } }
    if (prior && prior->valid) {      // re-express the prior's kept blocks (identified by address in the reference) as block ids
        pr_local = *prior; pr_x0.assign(VIWB_STATE_FIXED, 0.0);
        for (int i = 0; i < prior->num_blocks && prior_res; i++) {
            const auto hit = id.find(prior_res->blocks[i]);
            if (hit == id.end()) { summary->message = "the prior keeps a parameter block that is not part of the problem"; return; }
            const int bid = hit->second;
            pr_local.block_id[i] = bid;
            std::memcpy(pr_x0.data() + viwb_block_offset(bid), prior->x0 + viwb_block_offset(prior->block_id[i]), sizeof(double) * viwb_block_size(bid));
        }
        pr_local.x0 = pr_x0.data();
        pb.prior = &pr_local;
    }
    pb.num_vis = (int)vt.size(); pb.vis_type = vt.data(); pb.vis_landmark = vl.data(); pb.vis_frame_i = vi.data(); pb.vis_frame_j = vj.data(); pb.vis_obs = vobs.data();
    pb.num_imu = (int)ii.size(); pb.imu_frame_i = ii.data(); pb.imu_frame_j = ij.data(); pb.imu_data = idata.data();
    pb.num_wheel = (int)wi.size(); pb.wheel_frame_i = wi.data(); pb.wheel_frame_j = wj.data(); pb.wheel_data = wdata.data();
    pb.num_plane = (int)pf.size(); pb.plane_frame = pf.data();
    viwb_options opt; viwb_default_options(&opt);
    opt.max_num_iterations = options.max_num_iterations;
    opt.max_solver_time_in_seconds = options.max_solver_time_in_seconds >= 1e8 ? 0.0 : options.max_solver_time_in_seconds;
    opt.function_tolerance = options.function_tolerance; opt.gradient_tolerance = options.gradient_tolerance; opt.parameter_tolerance = options.parameter_tolerance;
    opt.initial_trust_region_radius = options.initial_trust_region_radius; opt.max_trust_region_radius = options.max_trust_region_radius;
    opt.min_trust_region_radius = options.min_trust_region_radius; opt.min_relative_decrease = options.min_relative_decrease;
    opt.min_lm_diagonal = options.min_lm_diagonal; opt.max_lm_diagonal = options.max_lm_diagonal;
    opt.max_num_consecutive_invalid_steps = options.max_num_consecutive_invalid_steps; opt.jacobi_scaling = options.jacobi_scaling ? 1 : 0;
    viwb_summary sm;
    const int rc = viwb_window_solve(ctx, &pb, state.data(), &opt, &sm);
    if (rc != VIWB_OK) { summary->message = viwb_last_error(ctx); return; }
    // parameter memory is caller-owned and updated in place (non-constant blocks only, like Ceres)
    for (auto &kv : id) if (!problem->blocks_[kv.first].constant) std::memcpy(kv.first, state.data() + viwb_block_offset(kv.second), sizeof(double) * viwb_block_size(kv.second));
    for (size_t k = 0; k < landmarks.size(); k++) landmarks[k][0] = state[VIWB_STATE_FIXED + k];
    summary->iterations.resize(sm.num_iterations);
    for (int i = 0; i < sm.num_iterations; i++) summary->iterations[i].iteration = i;
    summary->termination_type = sm.termination_type == VIWB_CONVERGENCE ? CONVERGENCE : sm.termination_type == VIWB_NO_CONVERGENCE ? NO_CONVERGENCE : FAILURE;
    summary->initial_cost = sm.initial_cost; summary->final_cost = sm.final_cost; summary->num_successful_steps = sm.num_successful_steps;
}

}  // namespace ceres
