/*
 * viw_oracle.h -- CPU FP64 restatement of VIW-Fusion's window-solve hot path.
 *
 * THIS IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product (libviwb.so) never links or calls it.
 *
 * PINNING: the factor, manifold, pre-integration, prior, marginalization, triangulation and re-anchoring functions are checked against THE
 * REFERENCE'S OWN SOURCES compiled unmodified into oracle/_ref/libviw_ref.so (oracle/Makefile `ref`, header stand-ins in oracle/refshim/;
 * tests/test_reference_factors.py), Estimator::optimization() of estimator.cpp included.  PARITY UNPINNED for the arithmetic inside ceres::Solve: the reference ships no golden vectors / known-answer tests for this path
 *  and ceres-solver is not in the image; the trust-region / dense-Schur arithmetic restates
 * ceres-solver 1.14.0 (README.md:41; docker/Dockerfile:3 pins 1.12.0 -- same refactored TrustRegionMinimizer) from its
 * published algorithm (trust_region_minimizer.cc, dogleg_strategy.cc, corrector.cc, schur_complement_solver.cc) and is anchored
 * by the reference's own sanctioned check (analytic vs forward-difference Jacobians, projectionTwoFrameTwoCamFactor.cpp:237-303),
 * by scipy.optimize.least_squares at convergence and by the marginalization identities (marginalization_factor.cpp:310-311)
 * -- see tests/test_oracle_*.py.
 *
 * All functions consume the POD tables of include/viwb.h.
 */
#ifndef VIW_ORACLE_H
#define VIW_ORACLE_H
#include "../include/viwb.h"

#ifdef __cplusplus
extern "C" {
#endif

#define VO_MAX_TRACE 64
typedef struct vo_trace_entry {
    int32_t iteration, step_is_valid, step_is_successful, reused;
    double cost, cost_change, model_cost_change, relative_decrease, step_norm, x_norm;
    double radius, mu, gradient_max_norm, dogleg_step_norm, alpha;
} vo_trace_entry;
typedef struct vo_trace { int32_t count; vo_trace_entry e[VO_MAX_TRACE]; } vo_trace;

/* ceres::CostFunction::Evaluate of the six analytic factors (same contract as viwb_factor_evaluate). */
int vo_factor_evaluate(int factor_type, const viwb_globals *globals, const double *consts,
                       const double *const *parameters, double *residuals, double **jacobians);
/* MarginalizationFactor::Evaluate (marginalization_factor.cpp:349-397); jacobian n x 207 or NULL. */
int vo_prior_evaluate(const viwb_prior *prior, const double *state, double *residuals, double *jacobian);
/* ceres::HuberLoss::Evaluate (third party) : rho[3] */
void vo_huber(double delta, double s, double rho[3]);

/* Normal equations at `state` over the fixed tangent layout (same contract as viwb_debug_normal_equations). */
int vo_normal_equations(const viwb_problem *problem, const double *state, double *H, double *g, double *lm, double *cost);
/* Robustified cost 0.5*sum rho(|r|^2) and the stacked corrected residual vector (for scipy cross-checks).
 * residuals may be NULL; returns the number of residual rows. */
int vo_cost(const viwb_problem *problem, const double *state, double *cost, double *residuals);

/* ceres::Solve with DENSE_SCHUR + DOGLEG (estimator.cpp:1643-1658). trace may be NULL. */
int vo_window_solve(const viwb_problem *problem, double *state, const viwb_options *options, viwb_summary *summary,
                    vo_trace *trace);
/* x (+) delta on the whole state: PoseLocalParameterization::Plus etc. delta is in the fixed tangent
 * layout [192] followed by one entry per landmark. */
void vo_state_plus(const viwb_problem *problem, const double *state, const double *delta, double *out);
/* double2vector()+vector2double() gauge re-anchoring (estimator.cpp:1224-1276). */
int vo_gauge_reanchor(const viwb_problem *problem, const double *state_before, double *state);
/* MarginalizationInfo (marginalization_factor.cpp:98-334, estimator.cpp:1669-1893).
 * A_out/b_out (optional, (m+n)^2 and (m+n)) return the pre-Schur system, mn_out = {m, n}. */
int vo_marginalize(const viwb_problem *problem, const double *state, int margin_flag, viwb_prior *prior_out,
                   double *A_out, double *b_out, int32_t *mn_out);
/* solve + reanchor + marginalize = Estimator::optimization(). num_threads is used for marginalization's
 * A,b construction (NUM_THREADS = 4, marginalization_factor.h:22). */
int vo_optimization(const viwb_problem *problem, double *state, const viwb_options *options, int margin_flag,
                    viwb_summary *summary, viwb_prior *prior_out);

/* Estimator::outliersRejection (estimator.cpp:2115-2185): out[num_landmarks] */
int vo_outlier_rejection(const viwb_problem *problem, const double *state, double focal_length, double threshold_px, uint8_t *out);

/* FeatureManager::triangulate (two-view branches, feature_manager.cpp:309-385) and removeBackShiftDepth (:457-493) */
int vo_triangulate(const double *state, int n, const int32_t *stereo, const int32_t *frame, const double *pt0, const double *pt1, double init_depth, double *depth);
int vo_shift_depth(int n, const double *uv, const double *depth_in, const double *marg_R, const double *marg_P, const double *new_R, const double *new_P,
                   double init_depth, double *depth_out);

/* n independent windows on `threads` pthreads, `repeat` passes (bench.py CPU arm); returns the number of optimisations run */
long vo_optimization_throughput(int n, const viwb_problem *problems, const double *const *states, const int32_t *flags,
                                const viwb_options *opt, int threads, int repeat);
/* the same with results: solved states, iteration counts and per window {n, |J_lin|_F^2, |J_lin^T r_lin|^2} of the new prior (any pointer may be NULL) */
long vo_optimization_many(int n, const viwb_problem *problems, const double *const *states, const int32_t *flags,
                          const viwb_options *opt, int threads, int repeat, double *const *out_states, int32_t *out_iters, double *out_prior);

/* IntegrationBase::propagate (integration_base.h:63-167) on a buffer of samples -> the 287-double record.
 * acc/gyr have (n+1) rows (sample 0 = acc_0/gyr_0), dt has n entries; noise = {ACC_N, GYR_N, ACC_W, GYR_W}. */
void vo_imu_preintegrate(int n, const double *dt, const double *acc, const double *gyr, const double *ba,
                         const double *bg, const double *noise, double *record);
/* WheelIntegrationBase::propagate (wheel_integration_base.h:67-177) -> the 78-double record.
 * noise = {VEL_N_wheel, GYR_N_wheel}; s = {sx, sy, sw}; td = linearized_td. */
void vo_wheel_preintegrate(int n, const double *dt, const double *vel, const double *gyr, const double *s,
                           double td, const double *noise, double *record);

/* Symmetric eigen decomposition used by marginalization (stands in for Eigen::SelfAdjointEigenSolver):
 * A (n x n row-major, symmetric) -> eigenvalues ascending w[n], eigenvectors in columns of V (row-major). */
int vo_sym_eig(int n, const double *A, double *w, double *V);

void vo_default_options(viwb_options *opt);
void vo_default_globals(viwb_globals *g);

#ifdef __cplusplus
}
#endif
#endif
