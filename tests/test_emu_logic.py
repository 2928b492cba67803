"""`not gpu`: the device source of libviwb.so, compiled for the host (tests/emu), against the CPU oracle.
This checks the kernels' algebra / indexing and the host-side lowering on the CPU-only box; the GPU tests
(test_gpu_parity.py) run the same checks on the real library through the C ABI."""
import pytest

import parity_checks as pc
from viwb import lib as viwb_lib


@pytest.fixture(scope="module")
def emu():
    from emu import build_emu
    ctx = viwb_lib.Context(0, build_emu.build())
    yield ctx
    ctx.close()


def test_factor_evaluate(emu, oracle):
    pc.check_factor_evaluate(emu, oracle, 4, max_each=2)


@pytest.mark.parametrize("cid", [1, 4])
def test_normal_equations(emu, oracle, cid):
    pc.check_normal_equations(emu, oracle, cid)


@pytest.mark.parametrize("cid", [1, 2, 3, 4, 6])     # 6 = stereo only, USE_IMU = 0 (estimator.cpp:1398-1403)
def test_solve(emu, oracle, cid):
    pc.check_solve(emu, oracle, cid)


def test_solve_with_prior(emu, oracle):
    pc.check_solve(emu, oracle, 4, prior_chain=True)


@pytest.mark.parametrize("cid", [4, 6])         # 6: double2vector's !USE_IMU branch (estimator.cpp:1288-1297)
def test_reanchor(emu, oracle, cid):
    pc.check_reanchor(emu, oracle, cid)


@pytest.mark.parametrize("cid", [1, 4, 6])
def test_marginalize(emu, oracle, cid):
    pc.check_marginalize(emu, oracle, cid)


@pytest.mark.parametrize("cid", [4, 6])         # both marginalisation flags; 6: drop sets without speed-bias blocks (estimator.cpp:1692,1795)
def test_sequence(emu, oracle, cid):
    pc.check_sequence(emu, oracle, cid)


def test_batch(emu, oracle):
    pc.check_batch_matches_single(emu, oracle)


def test_edge_cases(emu, oracle):
    pc.check_edge_cases(emu, oracle)


def test_lk(emu):
    pc.check_lk(emu)


def test_track_checked(emu):
    pc.check_track_checked(emu)


def test_unsorted_table_and_parallel_lowering(emu, oracle):
    pc.check_unsorted_table_and_threads(emu, oracle)


def test_solver_time_limit(emu, oracle):
    pc.check_solver_time_limit(emu, oracle)


def test_lk_batch(emu):
    pc.check_lk_batch(emu, streams=2, w=200, h=160, min_both=8)


def test_preintegration(emu, oracle):
    pc.check_preintegration(emu, oracle)


def test_outlier_rejection(emu, oracle):
    pc.check_outlier_rejection(emu, oracle)


def test_small_edges(emu, oracle):
    pc.check_small_edges(emu, oracle)


def test_triangulation(emu, oracle):
    pc.check_triangulation(emu, oracle)


def test_batch_properties_small(emu, oracle):
    pc.check_full_batch_properties(emu, oracle, distinct=2, copies=2)


def test_undistort_velocity(emu):
    pc.check_undistort_velocity(emu)


def test_set_mask(emu):
    pc.check_set_mask(emu)


def test_good_features_to_track(emu):
    pc.check_good_features(emu, full=False)           # the small images: the emulation runs every pixel on one host thread


def test_detector_batch(emu):
    pc.check_detector_batch(emu, streams=2)


def test_detector_resident(emu):
    pc.check_detector_resident(emu)


def test_tracker_session(emu):
    pc.check_tracker_session(emu, streams=2, w=240, h=180, ticks=4, max_cnt=40, min_dist=18)


def test_tracker_session_mono_no_flow_back(emu):
    pc.check_tracker_session(emu, streams=1, w=200, h=160, ticks=3, max_cnt=30, min_dist=15, stereo=False, flow_back=False, predict=False)


def test_tracker_edges(emu):
    pc.check_tracker_edges(emu)


@pytest.mark.parametrize("cid", [2, 4])
def test_factor_evaluate_vs_reference_code(emu, reference_code, cid):
    """every factor class of the library against the reference's own Evaluate (oracle/_ref: the reference sources, compiled)"""
    assert pc.check_factor_evaluate(emu, reference_code, cid, max_each=4) < 1e-9


@pytest.mark.parametrize("cid", [2, 4, 6])
def test_marginalize_vs_reference_code(emu, oracle, reference_code, cid):
    pc.check_marginalize_vs_reference_code(emu, oracle, reference_code, cid)


@pytest.mark.parametrize("cid", [2, 4, 6])
def test_reference_estimator_runs_on_this_backend(emu, oracle, reference_code, cid):
    """estimator.cpp of the reference, unmodified, with ceres::Solve answered by the library: north_star's drop-in, literally"""
    pc.check_reference_estimator_on_this_backend(emu, oracle, cid)


@pytest.mark.parametrize("cid", [1, 2, 3, 4])
def test_visual_imu_alignment(emu, oracle, cid):
    pc.check_visual_imu_alignment(emu, oracle, cid)


@pytest.mark.parametrize("cid", [2, 4])
def test_visual_imu_alignment_vs_reference_code(emu, reference_code, cid):
    pc.check_visual_imu_alignment_vs_reference_code(emu, cid)


@pytest.mark.parametrize("cid", [2, 4, 6])
def test_reference_estimator_on_product_shim(emu, reference_code, cid):
    """estimator.cpp of the reference, unmodified, compiled against the product's ceres shim + reference adapter"""
    from emu import build_emu
    pc.check_reference_estimator_on_product_shim(emu, build_emu.build(), cid)


def test_reference_tracker_runs_on_this_backend(emu, reference_code):
    """feature_tracker.cpp of the reference, unmodified, with its cv:: flow and corner calls answered by the library"""
    pc.check_reference_tracker_on_this_backend(emu)


@pytest.mark.parametrize("cid", [2, 4, 6])
def test_reference_estimator_on_product_shim_with_device_marginalization(emu, reference_code, cid):
    """adapter route with factor/marginalization_factor.cpp replaced by the product's marginalization_factor_device.cpp: the reference's own
    MarginalizationInfo class, marginalize() on the library under test"""
    from emu import build_emu
    pc.check_reference_estimator_on_product_shim(emu, build_emu.build(), cid, dev=True)


@pytest.mark.parametrize("cid", [2, 4, 6])
def test_device_marginalization_prior_factor_evaluate(emu, reference_code, cid):
    """MarginalizationFactor::Evaluate of the device translation unit (viwb_prior_evaluate underneath) against the reference's own Evaluate
    (marginalization_factor.cpp:349-397) on the prior each route has just produced, at the same perturbed blocks: |res|^2 and sum_b |J_b^T res|^2
    (both invariant under the orthogonal freedom of the square-root factorisation) agree"""
    import viw_ref
    from emu import build_emu
    from viwb import abi, synth
    path = build_emu.build()
    prob, st, gt = synth.make_window(cid, 0)
    a = viw_ref.prior_factor_digest_on_product_shim(path, prob, st, abi.MARGIN_OLD, False)
    b = viw_ref.prior_factor_digest_on_product_shim(path, prob, st, abi.MARGIN_OLD, True)
    assert a[2] == b[2] and a[2] > 0
    assert abs(a[0] - b[0]) <= 1e-6 * a[0] and abs(a[1] - b[1]) <= 1e-6 * a[1], (a, b)
