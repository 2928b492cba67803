// viwb_marginalization_lower.h -- the marginalization set (estimator.cpp:1670-1842) lowered to a viwb_problem and marginalised on the GPU
// (viwb_marginalize).  Shared by the shim's own MarginalizationInfo (viwb_marginalization_impl.h) and by
// factor/marginalization_factor_device.cpp, the replacement translation unit for the REFERENCE'S OWN MarginalizationInfo class.
#pragma once
#include <map>
#include <set>
#include <vector>
#include <cstring>

// ---- shared lowering: the marginalization set (estimator.cpp:1670-1842) -> viwb_problem -> viwb_marginalize on the GPU.
// Used by the shim's own MarginalizationInfo below and by factor/marginalization_factor_device.cpp, the replacement translation unit for the
// REFERENCE'S OWN MarginalizationInfo class (adapter route).  Factors are classified through viwb_shim::lower(): the shim's classes by their
// virtuals, the reference's classes by the installed adapter.
namespace viwb_shim {
struct MargFactor { const ceres::CostFunction *cost; const ceres::LossFunction *loss; const std::vector<double *> *blocks; const std::vector<int> *drop; };
struct MargResult {
    bool valid = false; int m = 0;
    viwb_prior prior; std::vector<double> x0, J, r;
    std::vector<double *> kept_addr;       // address (before addr_shift) of each kept block, aligned with prior.block_id
};
inline void marginalize_factors(const std::vector<MargFactor> &factors, MargResult &res) {
    res.valid = false; std::memset(&res.prior, 0, sizeof res.prior);
    bool has_non_prior = false;
    std::set<double *> dropped;
    std::vector<Lowered> low(factors.size());
    for (size_t k = 0; k < factors.size(); k++) {
        low[k] = lower(factors[k].cost);
        if (low[k].type != -2) has_non_prior = true;
        for (int d : *factors[k].drop) dropped.insert((*factors[k].blocks)[d]);
    }
    std::set<double *> poses, sbs; std::vector<double *> landmarks; std::map<double *, int> lm_index;
    double *ex0 = nullptr, *ex1 = nullptr, *exw = nullptr, *sx = nullptr, *sy = nullptr, *sw = nullptr, *tdw = nullptr, *td = nullptr, *pr_ = nullptr, *pz = nullptr;
    const viwb_prior *prior = nullptr; const std::vector<double *> *prior_blocks = nullptr; bool prior_exact = true;
    for (size_t k = 0; k < factors.size(); k++) {
        const int t = low[k].type; const std::vector<double *> &b = *factors[k].blocks;
        auto lm = [&](double *p) { if (!lm_index.count(p)) { lm_index[p] = (int)landmarks.size(); landmarks.push_back(p); } };
        switch (t) {
        case VIWB_F_PROJ_2F1C: poses.insert(b[0]); poses.insert(b[1]); ex0 = b[2]; lm(b[3]); td = b[4]; break;
        case VIWB_F_PROJ_2F2C: poses.insert(b[0]); poses.insert(b[1]); ex0 = b[2]; ex1 = b[3]; lm(b[4]); td = b[5]; break;
        case VIWB_F_PROJ_1F2C: ex0 = b[0]; ex1 = b[1]; lm(b[2]); td = b[3]; break;
        case VIWB_F_IMU: poses.insert(b[0]); sbs.insert(b[1]); poses.insert(b[2]); sbs.insert(b[3]); break;
        case VIWB_F_WHEEL: poses.insert(b[0]); poses.insert(b[1]); exw = b[2]; sx = b[3]; sy = b[4]; sw = b[5]; tdw = b[6]; break;
        case VIWB_F_PLANE: poses.insert(b[0]); exw = b[1]; pr_ = b[2]; pz = b[3]; break;
        case -2: prior = low[k].prior; prior_blocks = factors[k].blocks; prior_exact = low[k].prior_ids_exact; break;
        default: return;
        }
    }
    // the prior's kept blocks carry their block ids -- unless the prior was built by code that only knows the block SIZES (the reference's own CPU
    // marginalization behind the adapter): then window poses are told from extrinsics by the pose array's lattice of 7 doubles (estimator.h:191), and
    // every scalar block must be named by one of the other factors of the set
    if (prior_blocks && !prior_exact) {
        double *anchor = poses.empty() ? nullptr : *poses.begin();
        for (size_t i = 0; i < prior_blocks->size(); i++) {
            double *p = (*prior_blocks)[i]; const int size = viwb_block_size(prior->block_id[i]);
            if (size == 9) sbs.insert(p);
            else if (size == 7) {
                if (p == ex0 || p == ex1 || p == exw || poses.count(p)) continue;
                const std::ptrdiff_t d = anchor ? p - anchor : 1;
                if (anchor && d % 7 == 0 && d / 7 >= -(std::ptrdiff_t)VIWB_WINDOW_SIZE && d / 7 <= (std::ptrdiff_t)VIWB_WINDOW_SIZE) poses.insert(p);
                else if (!ex0) ex0 = p; else if (!ex1 && p == ex0 + 7) ex1 = p; else if (!exw) exw = p; else return;
            } else if (size == 4) pr_ = p;
            else if (p != td && p != tdw && p != sx && p != sy && p != sw && p != pz) return;      // an unnamed scalar block: cannot be told apart by size
        }
    }
    if (prior_blocks && prior_exact) for (size_t i = 0; i < prior_blocks->size(); i++) {
        double *p = (*prior_blocks)[i]; const int bid = prior->block_id[i];
        if (bid <= 10) poses.insert(p); else if (bid <= 21) sbs.insert(p);
        else if (bid == VIWB_BLK_EX_POSE0) ex0 = p; else if (bid == VIWB_BLK_EX_POSE1) ex1 = p; else if (bid == VIWB_BLK_EX_WHEEL) exw = p;
        else if (bid == VIWB_BLK_PLANE_R) pr_ = p; else if (bid == VIWB_BLK_PLANE_Z) pz = p; else if (bid == VIWB_BLK_SX) sx = p; else if (bid == VIWB_BLK_SY) sy = p;
        else if (bid == VIWB_BLK_SW) sw = p; else if (bid == VIWB_BLK_TD) td = p; else if (bid == VIWB_BLK_TD_WHEEL) tdw = p;
    }
    if (dropped.empty()) return;          // marginalization_factor.cpp:205-210
    std::map<double *, int> id;
    // frame indices: poses (and speed-biases) keep the window index they have by address; the estimator's arrays are
    // contiguous, so the index is recovered from the address distance to the lowest pose seen
    auto assign = [&](const std::set<double *> &st, int base, int stride) {
        if (st.empty()) return;
        double *lo = *st.begin();
        for (double *p : st) id[p] = base + (int)((p - lo) / stride);
    };
    assign(poses, VIWB_BLK_POSE0, 7); assign(sbs, VIWB_BLK_SPEEDBIAS0, 9);
    if (prior_blocks && prior_exact) {        // trust the prior's own ids for its blocks (they are exact); shift the rest consistently
        int shift_p = 0, shift_s = 0; bool hp = false, hs = false;
        for (size_t i = 0; i < prior_blocks->size(); i++) {
            double *p = (*prior_blocks)[i]; const int bid = prior->block_id[i];
            if (bid <= 10 && !hp) { shift_p = bid - id[p]; hp = true; }
            if (bid > 10 && bid <= 21 && !hs) { shift_s = bid - id[p]; hs = true; }
        }
        for (auto &kv : id) { if (kv.second <= 10) kv.second += shift_p; else kv.second += shift_s; }
    }
    auto put = [&](double *p, int b) { if (p) id[p] = b; };
    put(ex0, VIWB_BLK_EX_POSE0); put(ex1, VIWB_BLK_EX_POSE1); put(exw, VIWB_BLK_EX_WHEEL); put(pr_, VIWB_BLK_PLANE_R); put(pz, VIWB_BLK_PLANE_Z);
    put(sx, VIWB_BLK_SX); put(sy, VIWB_BLK_SY); put(sw, VIWB_BLK_SW); put(td, VIWB_BLK_TD); put(tdw, VIWB_BLK_TD_WHEEL);
    viwb_problem pb; std::memset(&pb, 0, sizeof pb);
    pb.frame_count = VIWB_WINDOW_SIZE; pb.num_landmarks = (int)landmarks.size(); pb.globals = globals();
    std::vector<double> state(VIWB_STATE_FIXED + landmarks.size(), 0.0);
    for (int i = 0; i < VIWB_NUM_FRAMES; i++) state[7 * i + 6] = 1.0;
    state[176 + 6] = state[183 + 6] = state[190 + 6] = 1.0; state[200] = 1.0;
    for (auto &kv : id) { pb.block_flags[kv.second] = VIWB_BLOCK_PRESENT; std::memcpy(state.data() + viwb_block_offset(kv.second), kv.first, sizeof(double) * viwb_block_size(kv.second)); }
    for (size_t k = 0; k < landmarks.size(); k++) state[VIWB_STATE_FIXED + k] = landmarks[k][0];
    std::vector<int32_t> vt, vl, vi, vj, ii, ij, wi, wj, pf; std::vector<double> vobs, idata, wdata;
    double huber = -1.0;
    for (size_t k = 0; k < factors.size(); k++) {
        const int t = low[k].type; const std::vector<double *> &b = *factors[k].blocks; const double *rec = low[k].record;
        if (factors[k].loss) huber = huber_delta(factors[k].loss);
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
    if (prior && prior->valid) {      // the prior's kept blocks under the ids of this lowering (the adapter hands out provisional ids by size)
        pin = *prior; pin_x0.assign(VIWB_STATE_FIXED, 0.0);
        for (int i = 0; i < prior->num_blocks; i++) {
            const auto hit = id.find((*prior_blocks)[i]);
            if (hit == id.end()) return;
            pin.block_id[i] = hit->second;
            std::memcpy(pin_x0.data() + viwb_block_offset(hit->second), prior->x0 + viwb_block_offset(prior->block_id[i]), sizeof(double) * viwb_block_size(hit->second));
        }
        pin.x0 = pin_x0.data(); pb.prior = &pin;
    }
    pb.num_vis = (int)vt.size(); pb.vis_type = vt.data(); pb.vis_landmark = vl.data(); pb.vis_frame_i = vi.data(); pb.vis_frame_j = vj.data(); pb.vis_obs = vobs.data();
    pb.num_imu = (int)ii.size(); pb.imu_frame_i = ii.data(); pb.imu_frame_j = ij.data(); pb.imu_data = idata.data();
    pb.num_wheel = (int)wi.size(); pb.wheel_frame_i = wi.data(); pb.wheel_frame_j = wj.data(); pb.wheel_data = wdata.data();
    pb.num_plane = (int)pf.size(); pb.plane_frame = pf.data();
    const int flag = has_non_prior || (dropped.size() && id.count(*dropped.begin()) && id[*dropped.begin()] == 0) ? VIWB_MARGIN_OLD : VIWB_MARGIN_SECOND_NEW;
    res.x0.assign(VIWB_STATE_FIXED, 0.0); res.J.assign((size_t)VIWB_MAX_PRIOR_DIM * VIWB_MAX_PRIOR_DIM, 0.0); res.r.assign(VIWB_MAX_PRIOR_DIM, 0.0);
    res.prior.x0 = res.x0.data(); res.prior.J = res.J.data(); res.prior.r = res.r.data();
    viwb_context *ctx = context();
    if (!ctx || viwb_marginalize(ctx, &pb, state.data(), flag, &res.prior) != VIWB_OK || !res.prior.valid) return;
    res.m = 0; for (double *p : dropped) res.m += (id.count(p) ? viwb_block_marg_size(id[p]) : 1);
    // remember the (pre-shift) address of every kept block: id before the slide -> address
    std::map<int, double *> addr_of; for (auto &kv : id) addr_of[kv.second] = kv.first;
    res.kept_addr.clear();
    for (int i = 0; i < res.prior.num_blocks; i++) {
        int old_id = res.prior.block_id[i];
        if (flag == VIWB_MARGIN_OLD) { if (old_id <= 9 || (old_id >= 11 && old_id <= 20)) old_id += 1; }      // ids were shifted by the slide
        else { if (old_id == 9 || old_id == 20) old_id += 1; }
        res.kept_addr.push_back(addr_of.count(old_id) ? addr_of[old_id] : nullptr);
    }
    res.valid = true;
}
}  // namespace viwb_shim

