// viwb_marginalization_impl.h -- MarginalizationInfo / MarginalizationFactor over viwb_marginalize / viwb_prior_evaluate.
#pragma once

inline void MarginalizationInfo::marginalize() {
    // Lower the collected factors (they ARE the marginalization set of estimator.cpp:1670-1776) to a viwb_problem that
    // contains only them, with the dropped pose as frame 0 (MARGIN_OLD) or the prior alone (MARGIN_SECOND_NEW).
    ceres::Problem tmp;
    bool has_non_prior = false;
    std::set<double *> dropped;
    for (auto *f : factors) {
        tmp.AddResidualBlock(f->cost_function, f->loss_function, f->parameter_blocks);
        if (f->cost_function->viwb_factor_type() != -2) has_non_prior = true;
        for (int d : f->drop_set) dropped.insert(f->parameter_blocks[d]);
    }
    // reuse Solve()'s classification by running the same lowering with zero iterations is not possible (it would not
    // marginalise), so lower here explicitly
    std::set<double *> poses, sbs; std::vector<double *> landmarks; std::map<double *, int> lm_index;
    double *ex0 = nullptr, *ex1 = nullptr, *exw = nullptr, *sx = nullptr, *sy = nullptr, *sw = nullptr, *tdw = nullptr, *td = nullptr, *pr_ = nullptr, *pz = nullptr;
    const viwb_prior *prior = nullptr; const ResidualBlockInfo *prior_info = nullptr;
    for (auto *f : factors) {
        const int t = f->cost_function->viwb_factor_type(); const std::vector<double *> &b = f->parameter_blocks;
        auto lm = [&](double *p) { if (!lm_index.count(p)) { lm_index[p] = (int)landmarks.size(); landmarks.push_back(p); } };
        switch (t) {
        case VIWB_F_PROJ_2F1C: poses.insert(b[0]); poses.insert(b[1]); ex0 = b[2]; lm(b[3]); td = b[4]; break;
        case VIWB_F_PROJ_2F2C: poses.insert(b[0]); poses.insert(b[1]); ex0 = b[2]; ex1 = b[3]; lm(b[4]); td = b[5]; break;
        case VIWB_F_PROJ_1F2C: ex0 = b[0]; ex1 = b[1]; lm(b[2]); td = b[3]; break;
        case VIWB_F_IMU: poses.insert(b[0]); sbs.insert(b[1]); poses.insert(b[2]); sbs.insert(b[3]); break;
        case VIWB_F_WHEEL: poses.insert(b[0]); poses.insert(b[1]); exw = b[2]; sx = b[3]; sy = b[4]; sw = b[5]; tdw = b[6]; break;
        case VIWB_F_PLANE: poses.insert(b[0]); exw = b[1]; pr_ = b[2]; pz = b[3]; break;
        case -2: prior = f->cost_function->viwb_prior_data(); prior_info = f; break;
        default: valid = false; tmp.residuals_.clear(); return;
        }
    }
    // the prior's kept blocks: classify by size (7 = window pose unless it is a known extrinsic, 9 = speed-bias, ...)
    std::vector<int> prior_ids;
    if (prior_info) for (size_t i = 0; i < prior_info->parameter_blocks.size(); i++) {
        double *p = prior_info->parameter_blocks[i]; const int bid = prior->block_id[i];
        if (bid <= 10) poses.insert(p); else if (bid <= 21) sbs.insert(p);
        else if (bid == VIWB_BLK_EX_POSE0) ex0 = p; else if (bid == VIWB_BLK_EX_POSE1) ex1 = p; else if (bid == VIWB_BLK_EX_WHEEL) exw = p;
        else if (bid == VIWB_BLK_PLANE_R) pr_ = p; else if (bid == VIWB_BLK_PLANE_Z) pz = p; else if (bid == VIWB_BLK_SX) sx = p; else if (bid == VIWB_BLK_SY) sy = p;
        else if (bid == VIWB_BLK_SW) sw = p; else if (bid == VIWB_BLK_TD) td = p; else if (bid == VIWB_BLK_TD_WHEEL) tdw = p;
    }
    tmp.residuals_.clear();                                   // tmp must not delete the cost functions it borrowed
    if (dropped.empty()) { valid = false; return; }          // marginalization_factor.cpp:205-210
    std::map<double *, int> id;
    // frame indices: poses (and speed-biases) keep the window index they have by address; the estimator's arrays are
    // contiguous, so the index is recovered from the address distance to the lowest pose seen
    auto assign = [&](const std::set<double *> &s, int base, int stride) {
        if (s.empty()) return;
        double *lo = *s.begin();
        // MARGIN_SECOND_NEW keeps poses 0..8,10: the window index must be the real one, taken from the prior's ids
        for (double *p : s) id[p] = base + (int)((p - lo) / stride);
    };
    assign(poses, VIWB_BLK_POSE0, 7); assign(sbs, VIWB_BLK_SPEEDBIAS0, 9);
    if (prior_info) {        // trust the prior's own ids for its blocks (they are exact); shift the rest consistently
        int shift_p = 0, shift_s = 0; bool hp = false, hs = false;
        for (size_t i = 0; i < prior_info->parameter_blocks.size(); i++) {
            double *p = prior_info->parameter_blocks[i]; const int bid = prior->block_id[i];
            if (bid <= 10 && !hp) { shift_p = bid - id[p]; hp = true; }
            if (bid > 10 && bid <= 21 && !hs) { shift_s = bid - id[p]; hs = true; }
        }
        for (auto &kv : id) { if (kv.second <= 10) kv.second += shift_p; else kv.second += shift_s; }
    }
    auto put = [&](double *p, int b) { if (p) id[p] = b; };
    put(ex0, VIWB_BLK_EX_POSE0); put(ex1, VIWB_BLK_EX_POSE1); put(exw, VIWB_BLK_EX_WHEEL); put(pr_, VIWB_BLK_PLANE_R); put(pz, VIWB_BLK_PLANE_Z);
    put(sx, VIWB_BLK_SX); put(sy, VIWB_BLK_SY); put(sw, VIWB_BLK_SW); put(td, VIWB_BLK_TD); put(tdw, VIWB_BLK_TD_WHEEL);
    viwb_problem pb; std::memset(&pb, 0, sizeof pb);
    pb.frame_count = VIWB_WINDOW_SIZE; pb.num_landmarks = (int)landmarks.size(); pb.globals = viwb_shim::globals();
    std::vector<double> state(VIWB_STATE_FIXED + landmarks.size(), 0.0);
    for (int i = 0; i < VIWB_NUM_FRAMES; i++) state[7 * i + 6] = 1.0;
    state[176 + 6] = state[183 + 6] = state[190 + 6] = 1.0; state[200] = 1.0;
    for (auto &kv : id) { pb.block_flags[kv.second] = VIWB_BLOCK_PRESENT; std::memcpy(state.data() + viwb_block_offset(kv.second), kv.first, sizeof(double) * viwb_block_size(kv.second)); }
    for (size_t k = 0; k < landmarks.size(); k++) state[VIWB_STATE_FIXED + k] = landmarks[k][0];
    std::vector<int32_t> vt, vl, vi, vj, ii, ij, wi, wj, pf; std::vector<double> vobs, idata, wdata;
    double huber = -1.0;
    for (auto *f : factors) {
        const int t = f->cost_function->viwb_factor_type(); const std::vector<double *> &b = f->parameter_blocks; const double *rec = f->cost_function->viwb_record();
        if (f->loss_function) huber = f->loss_function->viwb_huber_delta();
        if (t >= 0 && t <= VIWB_F_PROJ_1F2C) {
            const int li = t == VIWB_F_PROJ_2F1C ? 3 : t == VIWB_F_PROJ_2F2C ? 4 : 2;
            vt.push_back(t); vl.push_back(lm_index[b[li]]);
            if (t == VIWB_F_PROJ_1F2C) { vi.push_back(0); vj.push_back(0); } else { vi.push_back(id[b[0]]); vj.push_back(id[b[1]]); }
            vobs.insert(vobs.end(), rec, rec + VIWB_VIS_OBS_DOUBLES);
        } else if (t == VIWB_F_IMU) { ii.push_back(id[b[0]]); ij.push_back(id[b[2]]); idata.insert(idata.end(), rec, rec + VIWB_IMU_DOUBLES); }
        else if (t == VIWB_F_WHEEL) { wi.push_back(id[b[0]]); wj.push_back(id[b[1]]); wdata.insert(wdata.end(), rec, rec + VIWB_WHEEL_DOUBLES); }
        else if (t == VIWB_F_PLANE) pf.push_back(id[b[0]]);
    }
    if (huber > 0) pb.globals.huber_delta = huber;
    viwb_prior pin; std::vector<double> pin_x0;
    if (prior && prior->valid) { pin = *prior; pb.prior = &pin; }
    pb.num_vis = (int)vt.size(); pb.vis_type = vt.data(); pb.vis_landmark = vl.data(); pb.vis_frame_i = vi.data(); pb.vis_frame_j = vj.data(); pb.vis_obs = vobs.data();
    pb.num_imu = (int)ii.size(); pb.imu_frame_i = ii.data(); pb.imu_frame_j = ij.data(); pb.imu_data = idata.data();
    pb.num_wheel = (int)wi.size(); pb.wheel_frame_i = wi.data(); pb.wheel_frame_j = wj.data(); pb.wheel_data = wdata.data();
    pb.num_plane = (int)pf.size(); pb.plane_frame = pf.data();
    const int flag = has_non_prior || (dropped.size() && id.count(*dropped.begin()) && id[*dropped.begin()] == 0) ? VIWB_MARGIN_OLD : VIWB_MARGIN_SECOND_NEW;
    x0_.assign(VIWB_STATE_FIXED, 0.0); J_.assign((size_t)VIWB_MAX_PRIOR_DIM * VIWB_MAX_PRIOR_DIM, 0.0); r_.assign(VIWB_MAX_PRIOR_DIM, 0.0);
    prior_.x0 = x0_.data(); prior_.J = J_.data(); prior_.r = r_.data();
    viwb_context *ctx = viwb_shim::context();
    if (!ctx || viwb_marginalize(ctx, &pb, state.data(), flag, &prior_) != VIWB_OK || !prior_.valid) { valid = false; return; }
    n = prior_.n;
    m = 0; for (double *p : dropped) m += (id.count(p) ? viwb_block_marg_size(id[p]) : 1);
    // remember the (pre-shift) address of every kept block: id before the slide -> address
    std::map<int, double *> addr_of; for (auto &kv : id) addr_of[kv.second] = kv.first;
    kept_addr_.clear();
    for (int i = 0; i < prior_.num_blocks; i++) {
        int old_id = prior_.block_id[i];
        if (flag == VIWB_MARGIN_OLD) { if (old_id <= 9 || (old_id >= 11 && old_id <= 20)) old_id += 1; }      // ids were shifted by the slide
        else { if (old_id == 9 || old_id == 20) old_id += 1; }
        kept_addr_.push_back(addr_of.count(old_id) ? addr_of[old_id] : nullptr);
    }
}

inline std::vector<double *> MarginalizationInfo::getParameterBlocks(std::unordered_map<long, double *> &addr_shift) {
    std::vector<double *> keep_block_addr;
    keep_block_size.clear(); keep_block_idx.clear(); keep_block_data.clear();
    for (int i = 0; i < prior_.num_blocks; i++) {
        const int bid = prior_.block_id[i];
        keep_block_size.push_back(viwb_block_size(bid));
        keep_block_idx.push_back(prior_.block_idx[i] + m);
        keep_block_data.push_back(x0_.data() + viwb_block_offset(bid));
        keep_block_addr.push_back(kept_addr_[i] ? addr_shift[reinterpret_cast<long>(kept_addr_[i])] : nullptr);
    }
    sum_block_size = 0; for (int s : keep_block_size) sum_block_size += s;
    return keep_block_addr;
}

inline bool MarginalizationFactor::Evaluate(double const *const *parameters, double *residuals, double **jacobians) const {
    const viwb_prior *p = marginalization_info->prior();
    std::vector<double> state(VIWB_STATE_FIXED, 0.0), jac;
    for (int i = 0; i < p->num_blocks; i++) std::memcpy(state.data() + viwb_block_offset(p->block_id[i]), parameters[i], sizeof(double) * viwb_block_size(p->block_id[i]));
    if (jacobians) jac.resize((size_t)p->n * VIWB_STATE_FIXED);
    viwb_context *ctx = viwb_shim::context();
    if (!ctx || viwb_prior_evaluate(ctx, p, state.data(), residuals, jacobians ? jac.data() : nullptr) != VIWB_OK) return false;
    if (jacobians) for (int i = 0; i < p->num_blocks; i++) if (jacobians[i]) {
        const int gs = viwb_block_size(p->block_id[i]), off = viwb_block_offset(p->block_id[i]);
        for (int r = 0; r < p->n; r++) for (int c = 0; c < gs; c++) jacobians[i][r * gs + c] = jac[(size_t)r * VIWB_STATE_FIXED + off + c];
    }
    return true;
}
