// ref_estimator_glue.cpp -- TEST INFRASTRUCTURE ONLY.  Drives the reference's OWN Estimator::optimization() (estimator/estimator.cpp,
// compiled unmodified from /root/reference; ROS / OpenCV / camodocal / ceres replaced by the name stand-ins of oracle/refshim/) on a
// window given as include/viwb.h tables:
//   * the Estimator object is filled from the tables (states, feature manager, pre-integrations, last marginalization prior),
//   * optimization() runs as written: vector2double(), the whole ceres::Problem assembly, ceres::Solve, double2vector(), the
//     marginalization with its factor hand-over and address shift,
//   * ceres::Problem here RECORDS what it is given and ceres::Solve is a hook: it checks the record against the tables (the parity of the
//     problem assembly) and PLAYS BACK a solution computed elsewhere (the restated solver's; Ceres itself is not in the
//     image, so a-3 stays unpinned) into the parameter arrays,
//   * the outputs are the re-anchored window (double2vector then vector2double, a-1) and the new prior (a-13 incl. its orchestration).
// Compiled with -fno-access-control (the glue reads private members); nothing here is product code.
#include "estimator/estimator.h"
#include "factor/pose_subset_parameterization.h"
#include "factor/orientation_subset_parameterization.h"
#include "../../include/viwb.h"
#include <cstring>
#include <new>
#ifdef VIWB_PRODUCT_SHIM
// Third build (oracle/Makefile `ref_product`): <ceres/ceres.h> is the PRODUCT's shim (viw-fusion_b200/host), not the recording stand-in, and the
// reference's own factor / manifold classes are lowered by the product's adapter.  ceres::Solve below is then the product's, and what runs is
// the reference's unmodified estimator.cpp on the library the process has loaded (emulation or libviwb.so).
#include "viwb_reference_adapter.h"
#endif

CameraExtrinsicAdjustType CAM_EXT_ADJ_TYPE;
WheelExtrinsicAdjustType WHEEL_EXT_ADJ_TYPE;
#ifdef VIWB_PRODUCT_SHIM
double SOLVER_TIME = 1e9;              // the shim reads >= 1e8 as "no wall-clock limit" (parity runs must not depend on the clock)
#else
double SOLVER_TIME = 0.04;
#endif
int NUM_ITERATIONS = 8;
int SHOW_TRACK = 0;
double BIAS_ACC_THRESHOLD = 0.1, BIAS_GYR_THRESHOLD = 0.1, F_THRESHOLD = 1.0, OFFSET_SIM = 0.0;
int MAX_CNT = 150, MIN_DIST = 30, FLOW_BACK = 1, ROLLING_SHUTTER = 0;
std::string EX_CALIB_RESULT_PATH, IN_CALIB_RESULT_PATH, INTRINSIC_ITERATE_PATH, EXTRINSIC_WHEEL_ITERATE_PATH, EXTRINSIC_CAM_ITERATE_PATH, PROCESS_TIME_PATH, TD_WHEEL_PATH, TD_PATH,
    VINS_RESULT_PATH, GROUNDTRUTH_PATH, OUTPUT_FOLDER, IMU_TOPIC, WHEEL_TOPIC, IMAGE0_TOPIC, IMAGE1_TOPIC, FEATURE0_TOPIC, FEATURE1_TOPIC, GROUNDTRUTH_TOPIC, FISHEYE_MASK;
std::vector<std::string> CAM_NAMES;
map<int, Eigen::Vector3d> pts_gt;

extern double ACC_N, ACC_W, GYR_N, GYR_W, VEL_N_wheel, GYR_N_wheel, SX, SY, SW, ROLL_N_INV, PITCH_N_INV, ZPW_N_INV, TD, TD_WHEEL, INIT_DEPTH;
extern int NUM_OF_CAM, MULTIPLE_THREAD, ONLY_INITIAL_WITH_WHEEL, ESTIMATE_EXTRINSIC, ESTIMATE_EXTRINSIC_WHEEL, ESTIMATE_INTRINSIC_WHEEL, ESTIMATE_TD, ESTIMATE_TD_WHEEL, USE_IMU, USE_WHEEL, USE_PLANE, STEREO;

namespace {
struct Playback {
    const viwb_problem *p = nullptr;
    const double *solved = nullptr;
    Estimator *e = nullptr;
    const double *input = nullptr;
    int32_t *record = nullptr;        // [0..6] residual blocks: prior, imu, wheel, plane, 2F1C, 2F2C, 1F2C; [7] parameter blocks; [8] structure mismatches;
                                      // [9] vector2double mismatches; [10] visual rows whose constants / blocks differ from the tables
    bool active = false;
    int (*solve_cb)(const viwb_problem *, double *) = nullptr;      // when set, ceres::Solve hands the window to it instead of playing `solved` back
    std::vector<double> cb_state;
} g_pb;

double *block_ptr(Estimator *e, int b) {
    if (b >= 0 && b <= 10) return e->para_Pose[b];
    if (b >= 11 && b <= 21) return e->para_SpeedBias[b - 11];
    switch (b) {
        case 22: return e->para_Ex_Pose[0]; case 23: return e->para_Ex_Pose[1]; case 24: return e->para_Ex_Pose_wheel[0]; case 25: return e->para_plane_R[0];
        case 26: return e->para_plane_Z[0]; case 27: return e->para_Ix_sx_wheel[0]; case 28: return e->para_Ix_sy_wheel[0]; case 29: return e->para_Ix_sw_wheel[0];
        case 30: return e->para_Td[0]; case 31: return e->para_Td_wheel[0];
    }
    return nullptr;
}
int block_of(Estimator *e, const double *ptr) { for (int b = 0; b < VIWB_NUM_FIXED_BLOCKS; b++) if (block_ptr(e, b) == ptr) return b; return -1; }
int landmark_of(Estimator *e, const double *ptr) { const long k = (ptr - &e->para_Feature[0][0]); return (k >= 0 && k < NUM_OF_F) ? (int)k : -1; }
bool near(double a, double b) { return a == b || std::fabs(a - b) <= 1e-15 * std::max(1.0, std::fabs(b)); }
}  // namespace

#ifndef VIWB_PRODUCT_SHIM
ceres::Problem::~Problem() {}         // the objects a played-back problem holds are few and leak on purpose (test process)

void ceres::Solve(const Solver::Options &options, Problem *problem, Solver::Summary *summary) {
    if (!g_pb.active) { summary->termination_type = FAILURE; return; }
    Estimator *e = g_pb.e; const viwb_problem *p = g_pb.p; int32_t *rec = g_pb.record;
    // ---- vector2double (estimator.cpp:1155-1222): the arrays the solver sees must hold the window that was handed in
    int v2d = 0;
    for (int b = 0; b < VIWB_NUM_FIXED_BLOCKS; b++) {
        if (!(p->block_flags[b] & VIWB_BLOCK_PRESENT)) continue;
        const double *q = block_ptr(e, b), *x = g_pb.input + viwb_block_offset(b);
        for (int k = 0; k < viwb_block_size(b); k++) if (!near(q[k], x[k])) { v2d++; if (getenv("VIW_REF_DEBUG")) fprintf(stderr, "vector2double: block %d[%d] = %.17g, handed in %.17g\n", b, k, q[k], x[k]); }
    }
    for (int k = 0; k < p->num_landmarks; k++) if (!near(e->para_Feature[k][0], 1.0 / (1.0 / g_pb.input[VIWB_STATE_FIXED + k]))) v2d++;
    rec[9] = v2d;
    // ---- the record against the tables: parameter blocks (presence, constancy, manifold) ...
    int bad = 0;
    rec[7] = (int)problem->order_.size();
    for (int b = 0; b < VIWB_NUM_FIXED_BLOCKS; b++) {
        auto it = problem->blocks_.find(block_ptr(e, b));
        const bool present = it != problem->blocks_.end(), want = (p->block_flags[b] & VIWB_BLOCK_PRESENT) != 0;
        if (present != want) { bad++; continue; }
        if (!present) continue;
        const Problem::Block &k = it->second;
        if (k.size != viwb_block_size(b)) bad++;
        if (k.constant != ((p->block_flags[b] & VIWB_BLOCK_CONSTANT) != 0)) bad++;
        unsigned mask = 0;
        if (auto *ps = dynamic_cast<PoseSubsetParameterization *>(k.lp)) { for (size_t i = 0; i < ps->constancy_mask_.size() && i < 6; i++) if (ps->constancy_mask_[i]) mask |= 1u << i; }
        else if (auto *os = dynamic_cast<OrientationSubsetParameterization *>(k.lp)) { for (size_t i = 0; i < os->constancy_mask_.size() && i < 3; i++) if (os->constancy_mask_[i]) mask |= 1u << i; }
        else if (viwb_block_size(b) == 7 && !dynamic_cast<PoseLocalParameterization *>(k.lp)) bad++;
        if (!k.constant && mask != p->subset_mask[b]) bad++;
    }
    // ... and the residual blocks, in the order optimization() adds them (the tables list the factors in the same order)
    int cnt[7] = {0, 0, 0, 0, 0, 0, 0}, vis_bad = 0, iv = 0, ii = 0, iw = 0, ipl = 0;
    for (const Problem::Residual &r : problem->residuals_) {
        if (dynamic_cast<MarginalizationFactor *>(r.cost)) { cnt[0]++; continue; }
        if (dynamic_cast<IMUFactor *>(r.cost)) {
            if (ii >= p->num_imu || block_of(e, r.blocks[0]) != p->imu_frame_i[ii] || block_of(e, r.blocks[2]) != p->imu_frame_j[ii]) bad++;
            ii++; cnt[1]++; continue;
        }
        if (dynamic_cast<WheelFactor *>(r.cost)) {
            if (iw >= p->num_wheel || block_of(e, r.blocks[0]) != p->wheel_frame_i[iw] || block_of(e, r.blocks[1]) != p->wheel_frame_j[iw]) bad++;
            iw++; cnt[2]++; continue;
        }
        if (dynamic_cast<PlaneFactor *>(r.cost)) { if (ipl >= p->num_plane || block_of(e, r.blocks[0]) != p->plane_frame[ipl]) bad++; ipl++; cnt[3]++; continue; }
        int type = -1; const Eigen::Vector3d *pi = nullptr, *pj = nullptr, *vi = nullptr, *vj = nullptr; double tdi = 0, tdj = 0;
        if (auto *f = dynamic_cast<ProjectionTwoFrameOneCamFactor *>(r.cost)) { type = 0; pi = &f->pts_i; pj = &f->pts_j; vi = &f->velocity_i; vj = &f->velocity_j; tdi = f->td_i; tdj = f->td_j; }
        else if (auto *f = dynamic_cast<ProjectionTwoFrameTwoCamFactor *>(r.cost)) { type = 1; pi = &f->pts_i; pj = &f->pts_j; vi = &f->velocity_i; vj = &f->velocity_j; tdi = f->td_i; tdj = f->td_j; }
        else if (auto *f = dynamic_cast<ProjectionOneFrameTwoCamFactor *>(r.cost)) { type = 2; pi = &f->pts_i; pj = &f->pts_j; vi = &f->velocity_i; vj = &f->velocity_j; tdi = f->td_i; tdj = f->td_j; }
        if (type < 0) { bad++; continue; }
        cnt[4 + type]++;
        if (iv >= p->num_vis) { vis_bad++; continue; }
        const double *c = p->vis_obs + (size_t)iv * VIWB_VIS_OBS_DOUBLES;
        bool ok = p->vis_type[iv] == type && r.loss != nullptr;
        const int nb = (int)r.blocks.size();
        const int lmk = landmark_of(e, r.blocks[nb - 2]);
        ok = ok && lmk == p->vis_landmark[iv] && block_of(e, r.blocks[nb - 1]) == 30;
        if (type != 2) ok = ok && block_of(e, r.blocks[0]) == p->vis_frame_i[iv] && block_of(e, r.blocks[1]) == p->vis_frame_j[iv] && block_of(e, r.blocks[2]) == 22;
        else ok = ok && block_of(e, r.blocks[0]) == 22 && block_of(e, r.blocks[1]) == 23;
        if (type == 1) ok = ok && block_of(e, r.blocks[3]) == 23;
        for (int k = 0; k < 3; k++) ok = ok && near((*pi)(k), c[k]) && near((*pj)(k), c[3 + k]);
        for (int k = 0; k < 2; k++) ok = ok && near((*vi)(k), c[6 + k]) && near((*vj)(k), c[8 + k]);
        ok = ok && near(tdi, c[10]) && near(tdj, c[11]);
        if (!ok) vis_bad++;
        iv++;
    }
    for (int k = 0; k < 7; k++) rec[k] = cnt[k];
    if (ii != p->num_imu || iw != p->num_wheel || ipl != p->num_plane || iv != p->num_vis) bad++;
    if ((cnt[0] == 1) != (p->prior && p->prior->valid)) bad++;
    if (options.linear_solver_type != DENSE_SCHUR || options.trust_region_strategy_type != DOGLEG || options.max_num_iterations != NUM_ITERATIONS) bad++;
    rec[8] = bad; rec[10] = vis_bad;
    // ---- the solve itself: either a solution computed beforehand, or a callback that receives the tables and the window as vector2double()
    //      left it (tests hang the library under test there: the unmodified estimator.cpp then runs on the CUDA backend)
    if (g_pb.solve_cb) {
        g_pb.cb_state.assign(g_pb.input, g_pb.input + VIWB_STATE_FIXED + p->num_landmarks);
        for (int b = 0; b < VIWB_NUM_FIXED_BLOCKS; b++) if (problem->blocks_.count(block_ptr(e, b))) memcpy(g_pb.cb_state.data() + viwb_block_offset(b), block_ptr(e, b), sizeof(double) * viwb_block_size(b));
        for (int k = 0; k < p->num_landmarks; k++) g_pb.cb_state[VIWB_STATE_FIXED + k] = e->para_Feature[k][0];
        if (g_pb.solve_cb(p, g_pb.cb_state.data()) != 0) { summary->termination_type = FAILURE; return; }
        g_pb.solved = g_pb.cb_state.data();
    }
    // ---- play the solution back into the arrays, as ceres::Solve leaves them
    for (int b = 0; b < VIWB_NUM_FIXED_BLOCKS; b++)
        if (problem->blocks_.count(block_ptr(e, b))) memcpy(block_ptr(e, b), g_pb.solved + viwb_block_offset(b), sizeof(double) * viwb_block_size(b));
    for (int k = 0; k < p->num_landmarks; k++) e->para_Feature[k][0] = g_pb.solved[VIWB_STATE_FIXED + k];
    summary->iterations.resize(1);
    summary->termination_type = CONVERGENCE;
}
#endif  // !VIWB_PRODUCT_SHIM

static Eigen::Quaterniond quat_at(const double *q) { return Eigen::Quaterniond(q[3], q[0], q[1], q[2]); }
static MarginalizationInfo *estimator_prior(const viwb_prior *pr) {
    MarginalizationInfo *info = new MarginalizationInfo();
    info->n = pr->n; info->m = 0; info->valid = pr->valid != 0;
    for (int k = 0; k < pr->num_blocks; k++) {
        const int b = pr->block_id[k], size = viwb_block_size(b);
        double *d = new double[size];
        memcpy(d, pr->x0 + viwb_block_offset(b), sizeof(double) * size);
        info->keep_block_size.push_back(size); info->keep_block_idx.push_back(pr->block_idx[k]); info->keep_block_data.push_back(d);
    }
    info->linearized_jacobians.resize(pr->n, pr->n); info->linearized_residuals.resize(pr->n);
    for (int i = 0; i < pr->n; i++) { info->linearized_residuals(i) = pr->r[i]; for (int j = 0; j < pr->n; j++) info->linearized_jacobians(i, j) = pr->J[(size_t)i * pr->n + j]; }
    return info;
}
static int subset_type(unsigned mask) {                              // the constant tangent components of PoseSubsetParameterization -> the config enum
    switch (mask) { case 1u << 2: return 3; case 0x7: return 1; case 0x38: return 0; case 0x3c: return 4; default: return 2; }
}

// Builds an Estimator from the tables.  Returns nullptr on a table the reference bookkeeping cannot represent.
static Estimator *make_estimator(const viwb_problem *p, const double *st, int margin_flag) {
    const uint8_t *fl = p->block_flags;
    auto present = [&](int b) { return (fl[b] & VIWB_BLOCK_PRESENT) != 0; };
    auto free_ = [&](int b) { return present(b) && !(fl[b] & VIWB_BLOCK_CONSTANT); };
    USE_IMU = present(11); NUM_OF_CAM = present(23) ? 2 : 1; STEREO = NUM_OF_CAM == 2; USE_WHEEL = present(24); USE_PLANE = present(25);
    ONLY_INITIAL_WITH_WHEEL = 0; MULTIPLE_THREAD = 0;
    ESTIMATE_EXTRINSIC = free_(22); ESTIMATE_EXTRINSIC_WHEEL = free_(24); ESTIMATE_INTRINSIC_WHEEL = free_(27); ESTIMATE_TD = free_(30); ESTIMATE_TD_WHEEL = free_(31);
    CAM_EXT_ADJ_TYPE = (CameraExtrinsicAdjustType)subset_type(p->subset_mask[22]);
    WHEEL_EXT_ADJ_TYPE = (WheelExtrinsicAdjustType)subset_type(p->subset_mask[24]);
    const viwb_globals *g = &p->globals;
    G = Eigen::Vector3d(g->G[0], g->G[1], g->G[2]);
    Eigen::Matrix2d si; si << g->vis_sqrt_info[0], g->vis_sqrt_info[1], g->vis_sqrt_info[2], g->vis_sqrt_info[3];
    ProjectionTwoFrameOneCamFactor::sqrt_info = si; ProjectionTwoFrameTwoCamFactor::sqrt_info = si; ProjectionOneFrameTwoCamFactor::sqrt_info = si;
    PITCH_N_INV = g->plane_sqrt_info[0]; ROLL_N_INV = g->plane_sqrt_info[1]; ZPW_N_INV = g->plane_sqrt_info[2];
    void *mem = calloc(1, sizeof(Estimator));
    Estimator *e = new (mem) Estimator();                               // never destroyed: FeatureTracker & friends are name stubs
    e->frame_count = p->frame_count; e->solver_flag = Estimator::NON_LINEAR;
    e->marginalization_flag = margin_flag == VIWB_MARGIN_OLD ? Estimator::MARGIN_OLD : Estimator::MARGIN_SECOND_NEW;
    for (int i = 0; i <= 10; i++) {
        e->Ps[i] = Eigen::Vector3d(st[7 * i], st[7 * i + 1], st[7 * i + 2]); e->Rs[i] = quat_at(st + 7 * i + 3).toRotationMatrix();
        const double *sb = st + 77 + 9 * i;
        e->Vs[i] = Eigen::Vector3d(sb[0], sb[1], sb[2]); e->Bas[i] = Eigen::Vector3d(sb[3], sb[4], sb[5]); e->Bgs[i] = Eigen::Vector3d(sb[6], sb[7], sb[8]);
        e->Headers[i] = 0.05 * i;
    }
    for (int c = 0; c < 2; c++) { const double *x = st + 176 + 7 * c; e->tic[c] = Eigen::Vector3d(x[0], x[1], x[2]); e->ric[c] = quat_at(x + 3).toRotationMatrix(); }
    e->f_manager.setRic(e->ric);
    e->tio = Eigen::Vector3d(st[190], st[191], st[192]); e->rio = quat_at(st + 193).toRotationMatrix();
    e->rpw = quat_at(st + 197).toRotationMatrix(); e->zpw = st[201];
    e->sx = st[202]; e->sy = st[203]; e->sw = st[204]; e->td = st[205]; e->td_wheel = st[206];
    e->openExEstimation = free_(22); e->openExWheelEstimation = free_(24); e->openIxEstimation = free_(27); e->openPlaneEstimation = free_(25);
    // pre-integrations of the intervals (j-1, j)
    for (int f = 0; f < p->num_imu; f++) {
        const int j = p->imu_frame_j[f]; const double *c = p->imu_data + (size_t)f * VIWB_IMU_DOUBLES;
        if (p->imu_frame_i[f] != j - 1) return nullptr;
        IntegrationBase *pre = new IntegrationBase(Eigen::Vector3d::Zero(), Eigen::Vector3d::Zero(), Eigen::Vector3d(c[11], c[12], c[13]), Eigen::Vector3d(c[14], c[15], c[16]));
        pre->sum_dt = c[0]; pre->delta_p = Eigen::Vector3d(c[1], c[2], c[3]); pre->delta_q = Eigen::Quaterniond(c[7], c[4], c[5], c[6]); pre->delta_v = Eigen::Vector3d(c[8], c[9], c[10]);
        pre->jacobian.setIdentity();
        const int br[5] = {O_P, O_P, O_R, O_V, O_V}, bc[5] = {O_BA, O_BG, O_BG, O_BA, O_BG};
        for (int k = 0; k < 5; k++) for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) pre->jacobian(br[k] + a, bc[k] + b) = c[17 + 9 * k + 3 * a + b];
        for (int a = 0; a < 15; a++) for (int b = 0; b < 15; b++) pre->covariance(a, b) = c[62 + 15 * a + b];
        e->pre_integrations[j] = pre;
    }
    for (int f = 0; f < p->num_wheel; f++) {
        const int j = p->wheel_frame_j[f]; const double *c = p->wheel_data + (size_t)f * VIWB_WHEEL_DOUBLES;
        if (p->wheel_frame_i[f] != j - 1) return nullptr;
        WheelIntegrationBase *pre = new WheelIntegrationBase(Eigen::Vector3d(c[65], c[66], c[67]), Eigen::Vector3d(c[68], c[69], c[70]), c[61], c[62], c[63], c[64]);
        pre->delta_p = Eigen::Vector3d(c[0], c[1], c[2]); pre->delta_q = Eigen::Quaterniond(c[6], c[3], c[4], c[5]);
        for (int a = 0; a < 6; a++) for (int b = 0; b < 3; b++) pre->jacobian(a, b) = c[7 + 3 * a + b];
        for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) pre->covariance(a, b) = c[25 + 6 * a + b];
        pre->vel_1 = Eigen::Vector3d(c[71], c[72], c[73]); pre->gyr_1 = Eigen::Vector3d(c[74], c[75], c[76]); pre->sum_dt = c[77];
        e->pre_integrations_wheel[j] = pre;
    }
    // feature manager: one FeaturePerId per landmark, observations rebuilt from the factor rows (host observation = pts_i)
    std::vector<FeaturePerId> feats;
    for (int k = 0; k < p->num_landmarks; k++) feats.emplace_back(k, -1);
    for (int f = 0; f < p->num_vis; f++) {
        const double *c = p->vis_obs + (size_t)f * VIWB_VIS_OBS_DOUBLES;
        FeaturePerId &F = feats[p->vis_landmark[f]];
        const int host = p->vis_frame_i[f], j = p->vis_type[f] == VIWB_F_PROJ_1F2C ? host : p->vis_frame_j[f];
        if (F.start_frame < 0) {
            F.start_frame = host;
            Eigen::Matrix<double, 7, 1> a; a << c[0], c[1], c[2], 0.0, 0.0, c[6], c[7];
            F.feature_per_frame.push_back(FeaturePerFrame(a, c[10]));
        }
        Eigen::Matrix<double, 7, 1> zero; zero.setZero();
        while ((int)F.feature_per_frame.size() <= j - F.start_frame) F.feature_per_frame.push_back(FeaturePerFrame(zero, 0.0));
        FeaturePerFrame &slot = F.feature_per_frame[j - F.start_frame];
        Eigen::Matrix<double, 7, 1> o; o << c[3], c[4], c[5], 0.0, 0.0, c[8], c[9];
        if (p->vis_type[f] == VIWB_F_PROJ_2F1C) { slot.point = Eigen::Vector3d(c[3], c[4], c[5]); slot.velocity = Eigen::Vector2d(c[8], c[9]); slot.cur_td = c[11]; }
        else { slot.rightObservation(o); if (p->vis_type[f] == VIWB_F_PROJ_2F2C) slot.cur_td = c[11]; }
    }
    for (int k = 0; k < p->num_landmarks; k++) {
        if (feats[k].start_frame < 0) return nullptr;
        feats[k].estimated_depth = 1.0 / st[VIWB_STATE_FIXED + k];
        e->f_manager.feature.push_back(feats[k]);
    }
    if (p->prior && p->prior->valid) {
        e->last_marginalization_info = estimator_prior(p->prior);
        for (int k = 0; k < p->prior->num_blocks; k++) e->last_marginalization_parameter_blocks.push_back(block_ptr(e, p->prior->block_id[k]));
    }
    return e;
}

// Estimator::optimization() of the reference on the window `state_in`; `state_solved` = what ceres::Solve is to leave in the arrays.
// state_out: the window after double2vector() and another vector2double() (the convention of vo_optimization / viwb_optimization);
// mn = {m, n, kept blocks}; block ids after the address shift; J / r = the new prior; record[11] as documented at Playback.
static int estimator_optimization(const viwb_problem *p, const double *state_in, const double *state_solved, int (*cb)(const viwb_problem *, double *), int margin_flag,
                                  double *state_out, int32_t *mn, int32_t *block_id, int32_t *block_idx, double *J, double *r, int32_t *record, double *prior_eval = nullptr);
extern "C" int ref_estimator_optimization(const viwb_problem *p, const double *state_in, const double *state_solved, int margin_flag, double *state_out,
                                          int32_t *mn, int32_t *block_id, int32_t *block_idx, double *J, double *r, int32_t *record) {
    return estimator_optimization(p, state_in, state_solved, nullptr, margin_flag, state_out, mn, block_id, block_idx, J, r, record);
}
// the same with the solve delegated: cb(problem tables, window in / solved window out) -> 0 on success
extern "C" int ref_estimator_optimization_with(const viwb_problem *p, const double *state_in, int (*cb)(const viwb_problem *, double *), int margin_flag, double *state_out,
                                               int32_t *mn, int32_t *block_id, int32_t *block_idx, double *J, double *r, int32_t *record) {
    return estimator_optimization(p, state_in, nullptr, cb, margin_flag, state_out, mn, block_id, block_idx, J, r, record);
}
// prior_eval (optional, 3 doubles): MarginalizationFactor::Evaluate (marginalization_factor.cpp:349-397 -- or whatever translation unit defines it in this
// build) of the NEW prior at a perturbed copy of its blocks: {|res|^2, sum over blocks |J_b^T res|^2, number of residuals}.  Both numbers are invariant
// under the orthogonal freedom of the square-root factorisation, so builds with different eigen-decompositions can be compared.
extern "C" int ref_estimator_optimization_prior_eval(const viwb_problem *p, const double *state_in, int margin_flag, double *state_out,
                                                     int32_t *mn, int32_t *block_id, int32_t *block_idx, double *J, double *r, int32_t *record, double *prior_eval) {
    return estimator_optimization(p, state_in, nullptr, nullptr, margin_flag, state_out, mn, block_id, block_idx, J, r, record, prior_eval);
}
static int estimator_optimization(const viwb_problem *p, const double *state_in, const double *state_solved, int (*cb)(const viwb_problem *, double *), int margin_flag,
                                  double *state_out, int32_t *mn, int32_t *block_id, int32_t *block_idx, double *J, double *r, int32_t *record, double *prior_eval) {
    if (p->frame_count != 10) return 2;
    Estimator *e = make_estimator(p, state_in, margin_flag);
    if (!e) return 3;
#ifdef VIWB_PRODUCT_SHIM
    viwb_shim::install_reference_adapter();
#endif
    g_pb.solve_cb = cb;
    g_pb.p = p; g_pb.solved = state_solved; g_pb.e = e; g_pb.input = state_in; g_pb.record = record; g_pb.active = true;
    for (int k = 0; k < 11; k++) record[k] = -1;
    e->optimization();
    g_pb.active = false; g_pb.solve_cb = nullptr;
    e->vector2double();
    memcpy(state_out, state_in, sizeof(double) * (VIWB_STATE_FIXED + p->num_landmarks));
    for (int b = 0; b < VIWB_NUM_FIXED_BLOCKS; b++) if (p->block_flags[b] & VIWB_BLOCK_PRESENT) memcpy(state_out + viwb_block_offset(b), block_ptr(e, b), sizeof(double) * viwb_block_size(b));
    for (int k = 0; k < p->num_landmarks; k++) state_out[VIWB_STATE_FIXED + k] = e->para_Feature[k][0];
    mn[0] = mn[1] = mn[2] = 0;
    MarginalizationInfo *info = e->last_marginalization_info;
    if (info && info->valid && info->n > 0 && info->linearized_jacobians.rows() == info->n) {
        mn[0] = info->m; mn[1] = info->n; mn[2] = (int)e->last_marginalization_parameter_blocks.size();
        for (int k = 0; k < mn[2]; k++) { block_id[k] = block_of(e, e->last_marginalization_parameter_blocks[k]); block_idx[k] = info->keep_block_idx[k] - info->m; }
        for (int i = 0; i < info->n; i++) { r[i] = info->linearized_residuals(i); for (int j = 0; j < info->n; j++) J[(size_t)i * info->n + j] = info->linearized_jacobians(i, j); }
        if (prior_eval) {
            MarginalizationFactor factor(info);
            const int nb = (int)e->last_marginalization_parameter_blocks.size();
            std::vector<std::vector<double>> xs(nb), jac(nb);
            std::vector<const double *> params(nb); std::vector<double *> jp(nb);
            std::vector<double> res(info->n, 0.0);
            for (int k = 0; k < nb; k++) {
                const int size = info->keep_block_size[k];
                xs[k].assign(e->last_marginalization_parameter_blocks[k], e->last_marginalization_parameter_blocks[k] + size);
                const int bid = block_of(e, e->last_marginalization_parameter_blocks[k]);                 // the block's identity, not its place in this build's list
                for (int c = 0; c < size && c < 3; c++) xs[k][c] += 1e-3 * (1 + ((bid + c) % 3));        // a fixed, build-independent perturbation (positions / scalars only: unit quaternions stay unit)
                jac[k].assign((size_t)info->n * size, 0.0); params[k] = xs[k].data(); jp[k] = jac[k].data();
            }
            prior_eval[0] = prior_eval[1] = 0.0; prior_eval[2] = info->n;
            if (!factor.Evaluate(params.data(), res.data(), jp.data())) return 7;
            for (int i = 0; i < info->n; i++) prior_eval[0] += res[i] * res[i];
            for (int k = 0; k < nb; k++) { const int size = info->keep_block_size[k]; for (int c = 0; c < size; c++) { double g = 0.0; for (int i = 0; i < info->n; i++) g += jac[k][(size_t)i * size + c] * res[i]; prior_eval[1] += g * g; } }
        }
    }
    return 0;
}

// Estimator::outliersRejection (estimator.cpp:2127-2185) on the window `state`: out[num_landmarks] = 1 where the feature id is reported
extern "C" int ref_estimator_outliers(const viwb_problem *p, const double *state, uint8_t *out) {
    Estimator *e = make_estimator(p, state, VIWB_MARGIN_OLD);
    if (!e) return 3;
    std::set<int> removeIndex;
    e->outliersRejection(removeIndex);
    for (int k = 0; k < p->num_landmarks; k++) out[k] = removeIndex.count(k) ? 1 : 0;
    return 0;
}

// ---------------------------------------------------------------------------------------------------- VisualIMUAlignment (initial/initial_aligment.cpp:336-344)
// The reference's own solveGyroscopeBias (incl. its repropagate of every pre-integration) + LinearAlignment[WithWheel] + RefineGravity[WithWheel],
// compiled unmodified.  Frames in time order; interval i (frame i -> i+1) has counts[i] IMU steps, sample rows as in ref_imu_preintegrate.
// Outputs: delta_bg (what Bgs[] gained), the repropagated 287-double records, g, x, the return value.
extern std::vector<Eigen::Vector3d> TIC;
extern Eigen::Matrix3d RIO; extern Eigen::Vector3d TIO;
extern "C" int ref_visual_imu_alignment(int F, const double *R, const double *T, const int32_t *counts, const double *dt, const double *acc, const double *gyr, const double *noise,
                                        const double *bg0, const double *wheel_rec, const double *tic, const double *rio, const double *tio, const double *gvec,
                                        double *delta_bg, double *imu_rec, double *g_out, double *x_out, int32_t *x_size) {
    ACC_N = noise[0]; GYR_N = noise[1]; ACC_W = noise[2]; GYR_W = noise[3];
    G = Eigen::Vector3d(gvec[0], gvec[1], gvec[2]);
    TIC.assign(1, Eigen::Vector3d(tic[0], tic[1], tic[2]));
    USE_WHEEL = wheel_rec != nullptr;
    if (wheel_rec) { for (int i = 0; i < 3; i++) { TIO(i) = tio[i]; for (int j = 0; j < 3; j++) RIO(i, j) = rio[3 * i + j]; } }
    map<double, ImageFrame> frames;
    Eigen::Vector3d Bgs[WINDOW_SIZE + 1];
    for (int i = 0; i <= WINDOW_SIZE; i++) Bgs[i] = Eigen::Vector3d(bg0[0], bg0[1], bg0[2]);
    int s0 = 0;
    for (int f = 0; f < F; f++) {
        ImageFrame fr; fr.t = 0.05 * f; fr.is_key_frame = true; fr.pre_integration = nullptr; fr.pre_integration_wheel = nullptr;
        for (int i = 0; i < 3; i++) { fr.T(i) = T[3 * f + i]; for (int j = 0; j < 3; j++) fr.R(i, j) = R[9 * f + 3 * i + j]; }
        if (f > 0) {
            const int it = f - 1, r0 = s0 + it;
            IntegrationBase *pre = new IntegrationBase(Eigen::Vector3d(acc[3 * r0], acc[3 * r0 + 1], acc[3 * r0 + 2]), Eigen::Vector3d(gyr[3 * r0], gyr[3 * r0 + 1], gyr[3 * r0 + 2]),
                                                       Eigen::Vector3d::Zero(), Bgs[0]);
            for (int k = 0; k < counts[it]; k++) { const int r = r0 + k + 1; pre->push_back(dt[s0 + k], Eigen::Vector3d(acc[3 * r], acc[3 * r + 1], acc[3 * r + 2]), Eigen::Vector3d(gyr[3 * r], gyr[3 * r + 1], gyr[3 * r + 2])); }
            s0 += counts[it];
            fr.pre_integration = pre;
            if (wheel_rec) {
                const double *c = wheel_rec + (size_t)it * VIWB_WHEEL_DOUBLES;
                WheelIntegrationBase *w = new WheelIntegrationBase(Eigen::Vector3d(c[65], c[66], c[67]), Eigen::Vector3d(c[68], c[69], c[70]), c[61], c[62], c[63], c[64]);
                w->delta_p = Eigen::Vector3d(c[0], c[1], c[2]);
                fr.pre_integration_wheel = w;
            }
        }
        frames[fr.t] = fr;
    }
    Eigen::Vector3d g; Eigen::VectorXd x;
    const bool ok = VisualIMUAlignment(frames, Bgs, g, x);
    for (int i = 0; i < 3; i++) { delta_bg[i] = Bgs[0](i) - bg0[i]; g_out[i] = g(i); }
    *x_size = (int)x.size();
    for (int i = 0; i < (int)x.size(); i++) x_out[i] = x(i);
    int it = 0;
    for (auto &kv : frames) {
        IntegrationBase *p = kv.second.pre_integration;
        if (!p) continue;
        double *rec = imu_rec + (size_t)it * VIWB_IMU_DOUBLES; it++;
        rec[0] = p->sum_dt;
        for (int i = 0; i < 3; i++) { rec[1 + i] = p->delta_p(i); rec[8 + i] = p->delta_v(i); rec[11 + i] = p->linearized_ba(i); rec[14 + i] = p->linearized_bg(i); }
        rec[4] = p->delta_q.x(); rec[5] = p->delta_q.y(); rec[6] = p->delta_q.z(); rec[7] = p->delta_q.w();
        const int br[5] = {O_P, O_P, O_R, O_V, O_V}, bc[5] = {O_BA, O_BG, O_BG, O_BA, O_BG};
        for (int k = 0; k < 5; k++) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) rec[17 + 9 * k + 3 * i + j] = p->jacobian(br[k] + i, bc[k] + j);
        for (int i = 0; i < 15; i++) for (int j = 0; j < 15; j++) rec[62 + 15 * i + j] = p->covariance(i, j);
    }
    return ok ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------------- FeatureTracker (featureTracker/feature_tracker.cpp)
// The reference's own trackImage(), compiled unmodified.  The OpenCV routines it calls are forwarded to callbacks (refshim/opencv2/opencv.hpp)
// that the test points at the real cv2, so the tracker code runs on the library the reference links.
extern "C" void ref_set_cv_callbacks(cv::LkCallback lk, cv::GfttCallback gftt, cv::CircleCallback circle) { cv::g_lk_cb = lk; cv::g_gftt_cb = gftt; cv::g_circle_cb = circle; }
extern int ROW, COL;
extern "C" void *ref_tracker_create(const double *cam0, const double *cam1, int width, int height, int max_cnt, int min_dist, int flow_back) {
    ROW = height; COL = width; MAX_CNT = max_cnt; MIN_DIST = min_dist; FLOW_BACK = flow_back; SHOW_TRACK = 0;
    FeatureTracker *t = new FeatureTracker();
    auto mk = [&](const double *c) { return camodocal::CameraPtr(new camodocal::PinholeCamera("cam", width, height, c[4], c[5], c[6], c[7], c[0], c[1], c[2], c[3])); };
    t->m_camera.push_back(mk(cam0));
    if (cam1) { t->m_camera.push_back(mk(cam1)); t->stereo_cam = 1; }
    return t;
}
extern "C" void ref_tracker_destroy(void *h) { delete (FeatureTracker *)h; }
// one trackImage(); outputs in the tracker's own vector order: ids / track_cnt / {x, y, u, v, vx, vy} rows of both cameras
// hasPrediction / predict_pts as FeatureTracker::setPrediction() leaves them (feature_tracker.cpp:715-736): n points aligned with prev_pts
extern "C" int ref_tracker_set_prediction(void *h, int n, const float *pts) {
    FeatureTracker *t = (FeatureTracker *)h;
    if (n != (int)t->prev_pts.size()) return 1;
    t->hasPrediction = true; t->predict_pts.clear();
    for (int i = 0; i < n; i++) t->predict_pts.push_back(cv::Point2f(pts[2 * i], pts[2 * i + 1]));
    return 0;
}
extern "C" int ref_tracker_track(void *h, double time, unsigned char *left, unsigned char *right, int cap, int32_t *n_left, int32_t *ids, int32_t *cnt, float *feat,
                                 int32_t *n_right, int32_t *ids_r, float *feat_r) {
    FeatureTracker *t = (FeatureTracker *)h;
    cv::Mat l(ROW, COL, CV_8UC1, (void *)left), r = right ? cv::Mat(ROW, COL, CV_8UC1, (void *)right) : cv::Mat();
    t->trackImage(time, l, r);
    const int n = (int)t->ids.size(), m = (int)t->ids_right.size();
    if (n > cap || m > cap) return 1;
    *n_left = n; *n_right = m;
    for (int i = 0; i < n; i++) {
        ids[i] = t->ids[i]; cnt[i] = t->track_cnt[i];
        float *o = feat + 6 * i; o[0] = t->cur_un_pts[i].x; o[1] = t->cur_un_pts[i].y; o[2] = t->cur_pts[i].x; o[3] = t->cur_pts[i].y; o[4] = t->pts_velocity[i].x; o[5] = t->pts_velocity[i].y;
    }
    for (int i = 0; i < m; i++) {
        ids_r[i] = t->ids_right[i];
        float *o = feat_r + 6 * i; o[0] = t->cur_un_right_pts[i].x; o[1] = t->cur_un_right_pts[i].y; o[2] = t->cur_right_pts[i].x; o[3] = t->cur_right_pts[i].y;
        o[4] = t->right_pts_velocity[i].x; o[5] = t->right_pts_velocity[i].y;
    }
    return 0;
}

// The permutation std::sort (this libstdc++) gives FeatureTracker::setMask's vector for these track counts: setMask sorts
// (track_cnt, (point, id)) with `a.first > b.first` (feature_tracker.cpp:70-73), which leaves the order of equal counts to the library.
extern "C" void ref_std_sort_order(int n, const int32_t *cnt, int32_t *order) {
    std::vector<std::pair<int, int>> v;
    for (int i = 0; i < n; i++) v.push_back(std::make_pair(cnt[i], i));
    std::sort(v.begin(), v.end(), [](const std::pair<int, int> &a, const std::pair<int, int> &b) { return a.first > b.first; });
    for (int i = 0; i < n; i++) order[i] = v[i].second;
}
