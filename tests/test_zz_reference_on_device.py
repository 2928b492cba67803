"""The reference's own, unmodified estimator.cpp / feature_tracker.cpp (oracle/_ref) running on the library under test, on the GPU.
The emulation twins of these tests live in tests/test_emu_logic.py and are green; these were added after the round's GPU minutes were spent,
so they sit in the file pytest collects last: with `-x` a surprise here cannot hide any test that has already been seen green on a B200."""
import pytest

import parity_checks as pc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cid", [2, 4, 6])
def test_reference_estimator_runs_on_this_backend(gpu_ctx, oracle, reference_code, cid):
    """estimator.cpp of the reference, unmodified, with ceres::Solve answered by the library: north_star's drop-in, literally"""
    pc.check_reference_estimator_on_this_backend(gpu_ctx, oracle, cid)


def test_reference_tracker_runs_on_this_backend(gpu_ctx, reference_code):
    """feature_tracker.cpp of the reference, unmodified, with its cv:: flow and corner calls answered by the library"""
    pc.check_reference_tracker_on_this_backend(gpu_ctx)


@pytest.mark.parametrize("cid", [1, 2, 3, 4, 6])
def test_reference_estimator_on_product_shim(gpu_ctx, reference_code, cid):
    """estimator.cpp of the reference, unmodified, compiled against the product's ceres shim + reference adapter, solving on the GPU"""
    from viwb import lib as viwb_lib
    pc.check_reference_estimator_on_product_shim(gpu_ctx, viwb_lib.DEFAULT_LIB, cid)


@pytest.mark.parametrize("cid", [1, 2, 3, 4])
def test_visual_imu_alignment(gpu_ctx, oracle, cid):
    """SURVEY 8 f-4 ii: gyroscope bias, repropagation and the linear alignment with gravity refinement, on the device"""
    pc.check_visual_imu_alignment(gpu_ctx, oracle, cid)


@pytest.mark.parametrize("cid", [3, 4])
def test_visual_imu_alignment_vs_reference_code(gpu_ctx, reference_code, cid):
    pc.check_visual_imu_alignment_vs_reference_code(gpu_ctx, cid)


@pytest.mark.parametrize("cid", [2, 4, 6])
def test_reference_estimator_on_product_shim_with_device_marginalization(gpu_ctx, reference_code, cid):
    """the same with factor/marginalization_factor.cpp replaced by the product's marginalization_factor_device.cpp: the reference's own
    MarginalizationInfo class, its marginalize() on the GPU"""
    from viwb import lib as viwb_lib
    pc.check_reference_estimator_on_product_shim(gpu_ctx, viwb_lib.DEFAULT_LIB, cid, dev=True)

