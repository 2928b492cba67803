// A translation unit shaped like Estimator::optimization() (estimator.cpp:1383-1896) compiled against the viwb
// ceres shim: same ceres:: calls, same factor classes, same MarginalizationInfo flow -- only the state lives in
// plain arrays (Eigen is not in this image).  Reads a window dumped by tests/test_shim.py, writes the solved
// parameter blocks and the new prior.  Linked against libviwb.so (GPU) or the test-only emulation build.
#include <cstdio>
#include <cstdlib>
#include <unordered_map>
#include <vector>
#include "factor/viwb_factors.h"

struct V { std::vector<double> d; double operator()(int i) const { return d[i]; } };
static const int WINDOW_SIZE = 10, NUM_OF_F = 1000;
static double para_Pose[WINDOW_SIZE + 1][7], para_SpeedBias[WINDOW_SIZE + 1][9], para_Feature[NUM_OF_F][1], para_Ex_Pose[2][7],
    para_Ex_Pose_wheel[1][7], para_plane_R[1][4], para_plane_Z[1][1], para_Ix_sx_wheel[1][1], para_Ix_sy_wheel[1][1], para_Ix_sw_wheel[1][1],
    para_Td[1][1], para_Td_wheel[1][1];

template <class T> static std::vector<T> rd(FILE *f) { int n; if (fread(&n, 4, 1, f) != 1) exit(2); std::vector<T> v(n); if (n && fread(v.data(), sizeof(T), n, f) != (size_t)n) exit(2); return v; }

int main(int argc, char **argv) {
    if (argc < 3) return 1;
    FILE *f = fopen(argv[1], "rb"); if (!f) return 1;
    std::vector<int> hdr = rd<int>(f);       // frame_count, nlm, num_cam, use_wheel, use_plane, est_ex, ex_mask, est_td, has_prior, margin_flag
    std::vector<double> state = rd<double>(f), vobs = rd<double>(f), imu = rd<double>(f), wheel = rd<double>(f), gl = rd<double>(f);
    std::vector<int> vt = rd<int>(f), vl = rd<int>(f), vi = rd<int>(f), vj = rd<int>(f), pflags = rd<int>(f);
    std::vector<int> pr_hdr = rd<int>(f), pr_id = rd<int>(f), pr_idx = rd<int>(f); std::vector<double> pr_x0 = rd<double>(f), pr_J = rd<double>(f), pr_r = rd<double>(f);
    fclose(f);
    const int frame_count = hdr[0], nlm = hdr[1], NUM_OF_CAM = hdr[2], USE_WHEEL = hdr[3], USE_PLANE = hdr[4], ESTIMATE_EXTRINSIC = hdr[5], ex_mask = hdr[6], ESTIMATE_TD = hdr[7];
    const int has_prior = hdr[8], margin_flag = hdr[9];
    // vector2double()
    for (int i = 0; i <= WINDOW_SIZE; i++) { for (int k = 0; k < 7; k++) para_Pose[i][k] = state[7 * i + k]; for (int k = 0; k < 9; k++) para_SpeedBias[i][k] = state[77 + 9 * i + k]; }
    for (int c = 0; c < 2; c++) for (int k = 0; k < 7; k++) para_Ex_Pose[c][k] = state[176 + 7 * c + k];
    for (int k = 0; k < 7; k++) para_Ex_Pose_wheel[0][k] = state[190 + k];
    for (int k = 0; k < 4; k++) para_plane_R[0][k] = state[197 + k];
    para_plane_Z[0][0] = state[201]; para_Ix_sx_wheel[0][0] = state[202]; para_Ix_sy_wheel[0][0] = state[203]; para_Ix_sw_wheel[0][0] = state[204];
    para_Td[0][0] = state[205]; para_Td_wheel[0][0] = state[206];
    for (int k = 0; k < nlm; k++) para_Feature[k][0] = state[207 + k];
    viwb_globals &g = viwb_shim::globals();
    for (int k = 0; k < 3; k++) g.G[k] = gl[k];
    for (int k = 0; k < 4; k++) g.vis_sqrt_info[k] = gl[3 + k];
    for (int k = 0; k < 3; k++) g.plane_sqrt_info[k] = gl[7 + k];
    // the marginalization info of the previous optimization() (last_marginalization_info / ..._parameter_blocks)
    MarginalizationInfo *last_marginalization_info = nullptr; std::vector<double *> last_marginalization_parameter_blocks;
    auto addr_of_block = [&](int b) -> double * {
        return b <= 10 ? para_Pose[b] : b <= 21 ? para_SpeedBias[b - 11] : b <= 23 ? para_Ex_Pose[b - 22] : b == 24 ? para_Ex_Pose_wheel[0] :
               b == 25 ? para_plane_R[0] : b == 26 ? para_plane_Z[0] : b == 27 ? para_Ix_sx_wheel[0] : b == 28 ? para_Ix_sy_wheel[0] :
               b == 29 ? para_Ix_sw_wheel[0] : b == 30 ? para_Td[0] : para_Td_wheel[0]; };
    struct LoadedInfo : MarginalizationInfo { };      // a prior loaded from disk stands in for last_marginalization_info
    viwb_prior loaded; std::vector<double> lx0 = pr_x0, lJ = pr_J, lr = pr_r;
    struct PriorFactor : ceres::CostFunction {        // MarginalizationFactor over a loaded prior
        const viwb_prior *p; explicit PriorFactor(const viwb_prior *pp) : p(pp) { for (int i = 0; i < pp->num_blocks; i++) mutable_parameter_block_sizes()->push_back(viwb_block_size(pp->block_id[i])); set_num_residuals(pp->n); }
        bool Evaluate(double const *const *, double *, double **) const override { return false; }
        int viwb_factor_type() const override { return -2; } const viwb_prior *viwb_prior_data() const override { return p; } };
    if (has_prior) {
        loaded.valid = 1; loaded.n = pr_hdr[0]; loaded.num_blocks = (int)pr_id.size();
        for (size_t i = 0; i < pr_id.size(); i++) { loaded.block_id[i] = pr_id[i]; loaded.block_idx[i] = pr_idx[i]; last_marginalization_parameter_blocks.push_back(addr_of_block(pr_id[i])); }
        loaded.x0 = lx0.data(); loaded.J = lJ.data(); loaded.r = lr.data();
    }
    // ---------------- optimization(): problem assembly exactly in the reference's order (estimator.cpp:1388-1638)
    ceres::Problem problem;
    ceres::LossFunction *loss_function = new ceres::HuberLoss(1.0);
    for (int i = 0; i < frame_count + 1; i++) {
        ceres::LocalParameterization *local_parameterization = new PoseLocalParameterization();
        problem.AddParameterBlock(para_Pose[i], 7, local_parameterization);
        problem.AddParameterBlock(para_SpeedBias[i], 9);
    }
    for (int i = 0; i < NUM_OF_CAM; i++) {
        ceres::LocalParameterization *lp;
        if (ESTIMATE_EXTRINSIC) { std::vector<int> c; for (int k = 0; k < 7; k++) if ((ex_mask >> k) & 1) c.push_back(k); lp = new PoseSubsetParameterization(c); }
        else lp = new PoseLocalParameterization();
        problem.AddParameterBlock(para_Ex_Pose[i], 7, lp);
        if (!ESTIMATE_EXTRINSIC) problem.SetParameterBlockConstant(para_Ex_Pose[i]);
    }
    if (USE_WHEEL) {
        problem.AddParameterBlock(para_Ex_Pose_wheel[0], 7, new PoseLocalParameterization());
        problem.SetParameterBlockConstant(para_Ex_Pose_wheel[0]);
        problem.AddParameterBlock(para_Ix_sx_wheel[0], 1); problem.AddParameterBlock(para_Ix_sy_wheel[0], 1); problem.AddParameterBlock(para_Ix_sw_wheel[0], 1);
        problem.SetParameterBlockConstant(para_Ix_sx_wheel[0]); problem.SetParameterBlockConstant(para_Ix_sy_wheel[0]); problem.SetParameterBlockConstant(para_Ix_sw_wheel[0]);
    }
    if (USE_PLANE) {
        problem.AddParameterBlock(para_plane_R[0], 4, new OrientationSubsetParameterization(std::vector<int>{2}));
        problem.AddParameterBlock(para_plane_Z[0], 1);
    }
    problem.AddParameterBlock(para_Td[0], 1);
    problem.AddParameterBlock(para_Td_wheel[0], 1);
    if (!ESTIMATE_TD) problem.SetParameterBlockConstant(para_Td[0]);
    problem.SetParameterBlockConstant(para_Td_wheel[0]);
    if (has_prior) problem.AddResidualBlock(new PriorFactor(&loaded), NULL, last_marginalization_parameter_blocks);
    for (int i = 0; i < frame_count; i++) { int j = i + 1; problem.AddResidualBlock(new IMUFactor(imu.data() + (size_t)i * 287), NULL, para_Pose[i], para_SpeedBias[i], para_Pose[j], para_SpeedBias[j]); }
    if (USE_WHEEL) for (int i = 0; i < frame_count; i++) { int j = i + 1;
        problem.AddResidualBlock(new WheelFactor(wheel.data() + (size_t)i * 78), NULL, para_Pose[i], para_Pose[j], para_Ex_Pose_wheel[0], para_Ix_sx_wheel[0], para_Ix_sy_wheel[0], para_Ix_sw_wheel[0], para_Td_wheel[0]); }
    if (USE_PLANE) for (int i = 0; i < frame_count; i++) problem.AddResidualBlock(new PlaneFactor(), NULL, para_Pose[i], para_Ex_Pose_wheel[0], para_plane_R[0], para_plane_Z[0]);
    auto make_vis = [&](size_t k, ceres::Problem *pb, MarginalizationInfo *mi) {
        const double *o = vobs.data() + k * 12;
        V pi{{o[0], o[1], o[2]}}, pj{{o[3], o[4], o[5]}}, vi_{{o[6], o[7]}}, vj_{{o[8], o[9]}};
        const int imu_i = vi[k], imu_j = vj[k], fidx = vl[k];
        if (vt[k] == 0) { auto *fac = new ProjectionTwoFrameOneCamFactor(pi, pj, vi_, vj_, o[10], o[11]);
            if (pb) pb->AddResidualBlock(fac, loss_function, para_Pose[imu_i], para_Pose[imu_j], para_Ex_Pose[0], para_Feature[fidx], para_Td[0]);
            else mi->addResidualBlockInfo(new ResidualBlockInfo(fac, loss_function, std::vector<double *>{para_Pose[imu_i], para_Pose[imu_j], para_Ex_Pose[0], para_Feature[fidx], para_Td[0]}, std::vector<int>{0, 3})); }
        else if (vt[k] == 1) { auto *fac = new ProjectionTwoFrameTwoCamFactor(pi, pj, vi_, vj_, o[10], o[11]);
            if (pb) pb->AddResidualBlock(fac, loss_function, para_Pose[imu_i], para_Pose[imu_j], para_Ex_Pose[0], para_Ex_Pose[1], para_Feature[fidx], para_Td[0]);
            else mi->addResidualBlockInfo(new ResidualBlockInfo(fac, loss_function, std::vector<double *>{para_Pose[imu_i], para_Pose[imu_j], para_Ex_Pose[0], para_Ex_Pose[1], para_Feature[fidx], para_Td[0]}, std::vector<int>{0, 4})); }
        else { auto *fac = new ProjectionOneFrameTwoCamFactor(pi, pj, vi_, vj_, o[10], o[11]);
            if (pb) pb->AddResidualBlock(fac, loss_function, para_Ex_Pose[0], para_Ex_Pose[1], para_Feature[fidx], para_Td[0]);
            else mi->addResidualBlockInfo(new ResidualBlockInfo(fac, loss_function, std::vector<double *>{para_Ex_Pose[0], para_Ex_Pose[1], para_Feature[fidx], para_Td[0]}, std::vector<int>{2})); }
    };
    for (size_t k = 0; k < vt.size(); k++) make_vis(k, &problem, nullptr);
    ceres::Solver::Options options;
    options.linear_solver_type = ceres::DENSE_SCHUR;
    options.trust_region_strategy_type = ceres::DOGLEG;
    options.max_num_iterations = 8;
    ceres::Solver::Summary summary;
    ceres::Solve(options, &problem, &summary);
    std::printf("%s\n", summary.BriefReport().c_str());
    // (double2vector / vector2double happen in estimator.cpp itself; the test applies them through the C ABI afterwards)
    // ---------------- MARGIN_OLD block (estimator.cpp:1670-1818)
    MarginalizationInfo *marginalization_info = new MarginalizationInfo();
    if (margin_flag == 0) {
        if (has_prior) {
            std::vector<int> drop_set;
            for (int i = 0; i < (int)last_marginalization_parameter_blocks.size(); i++)
                if (last_marginalization_parameter_blocks[i] == para_Pose[0] || last_marginalization_parameter_blocks[i] == para_SpeedBias[0]) drop_set.push_back(i);
            marginalization_info->addResidualBlockInfo(new ResidualBlockInfo(new PriorFactor(&loaded), NULL, last_marginalization_parameter_blocks, drop_set));
        }
        marginalization_info->addResidualBlockInfo(new ResidualBlockInfo(new IMUFactor(imu.data() + 0), NULL, std::vector<double *>{para_Pose[0], para_SpeedBias[0], para_Pose[1], para_SpeedBias[1]}, std::vector<int>{0, 1}));
        if (USE_WHEEL) marginalization_info->addResidualBlockInfo(new ResidualBlockInfo(new WheelFactor(wheel.data() + 0), NULL,
            std::vector<double *>{para_Pose[0], para_Pose[1], para_Ex_Pose_wheel[0], para_Ix_sx_wheel[0], para_Ix_sy_wheel[0], para_Ix_sw_wheel[0], para_Td_wheel[0]}, std::vector<int>{0}));
        if (USE_PLANE) marginalization_info->addResidualBlockInfo(new ResidualBlockInfo(new PlaneFactor(), NULL, std::vector<double *>{para_Pose[0], para_Ex_Pose_wheel[0], para_plane_R[0], para_plane_Z[0]}, std::vector<int>{0}));
        for (size_t k = 0; k < vt.size(); k++) if (vi[k] == 0) make_vis(k, nullptr, marginalization_info);
        marginalization_info->preMarginalize();
        marginalization_info->marginalize();
        std::unordered_map<long, double *> addr_shift;
        for (int i = 1; i <= WINDOW_SIZE; i++) { addr_shift[reinterpret_cast<long>(para_Pose[i])] = para_Pose[i - 1]; addr_shift[reinterpret_cast<long>(para_SpeedBias[i])] = para_SpeedBias[i - 1]; }
        for (int i = 0; i < 2; i++) addr_shift[reinterpret_cast<long>(para_Ex_Pose[i])] = para_Ex_Pose[i];
        addr_shift[reinterpret_cast<long>(para_Ex_Pose_wheel[0])] = para_Ex_Pose_wheel[0];
        addr_shift[reinterpret_cast<long>(para_Ix_sx_wheel[0])] = para_Ix_sx_wheel[0]; addr_shift[reinterpret_cast<long>(para_Ix_sy_wheel[0])] = para_Ix_sy_wheel[0];
        addr_shift[reinterpret_cast<long>(para_Ix_sw_wheel[0])] = para_Ix_sw_wheel[0]; addr_shift[reinterpret_cast<long>(para_plane_R[0])] = para_plane_R[0];
        addr_shift[reinterpret_cast<long>(para_plane_Z[0])] = para_plane_Z[0]; addr_shift[reinterpret_cast<long>(para_Td[0])] = para_Td[0]; addr_shift[reinterpret_cast<long>(para_Td_wheel[0])] = para_Td_wheel[0];
        std::vector<double *> parameter_blocks = marginalization_info->getParameterBlocks(addr_shift);
        last_marginalization_info = marginalization_info; last_marginalization_parameter_blocks = parameter_blocks;
    }
    // ---------------- results
    FILE *o = fopen(argv[2], "wb");
    auto wr = [&](const double *p, int n) { fwrite(&n, 4, 1, o); fwrite(p, 8, n, o); };
    std::vector<double> out(207 + nlm);
    for (int i = 0; i <= WINDOW_SIZE; i++) { for (int k = 0; k < 7; k++) out[7 * i + k] = para_Pose[i][k]; for (int k = 0; k < 9; k++) out[77 + 9 * i + k] = para_SpeedBias[i][k]; }
    for (int c = 0; c < 2; c++) for (int k = 0; k < 7; k++) out[176 + 7 * c + k] = para_Ex_Pose[c][k];
    for (int k = 0; k < 7; k++) out[190 + k] = para_Ex_Pose_wheel[0][k];
    for (int k = 0; k < 4; k++) out[197 + k] = para_plane_R[0][k];
    out[201] = para_plane_Z[0][0]; out[202] = para_Ix_sx_wheel[0][0]; out[203] = para_Ix_sy_wheel[0][0]; out[204] = para_Ix_sw_wheel[0][0]; out[205] = para_Td[0][0]; out[206] = para_Td_wheel[0][0];
    for (int k = 0; k < nlm; k++) out[207 + k] = para_Feature[k][0];
    wr(out.data(), (int)out.size());
    double meta[4] = {(double)summary.iterations.size(), summary.final_cost, last_marginalization_info ? (double)last_marginalization_info->n : 0.0, last_marginalization_info ? (double)last_marginalization_info->valid : 0.0};
    wr(meta, 4);
    if (last_marginalization_info && last_marginalization_info->valid) {
        const viwb_prior *p = last_marginalization_info->prior();
        std::vector<double> ids; for (int i = 0; i < p->num_blocks; i++) { ids.push_back(p->block_id[i]); ids.push_back(p->block_idx[i]);
            // the block list the estimator would keep must address the shifted blocks
            ids.push_back(last_marginalization_parameter_blocks[i] == addr_of_block(p->block_id[i]) ? 1.0 : 0.0); }
        wr(ids.data(), (int)ids.size()); wr(p->J, p->n * p->n); wr(p->r, p->n);
    }
    fclose(o);
    delete marginalization_info;
    return 0;
}
