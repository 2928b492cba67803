import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "viw-fusion_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import viw_oracle
    viw_oracle.lib()
    return viw_oracle


@pytest.fixture(scope="session")
def gpu_ctx():
    """A libviwb context on cuda:0; fails loudly if the CUDA extension is missing."""
    from viwb import lib as viwb_lib
    ctx = viwb_lib.Context(0)
    yield ctx
    ctx.close()


class _ReferenceCode:
    """The reference's own compiled factor / marginalization code (oracle/_ref) behind the oracle module's call signatures, so that the
    parity checks written against the oracle can run against the reference itself."""

    def __init__(self, vr):
        self.vr = vr

    def factor_evaluate(self, ftype, globals_, consts, params, want_jac=True, null_jac=()):
        from viwb import abi
        return self.vr.factor_evaluate(ftype, globals_, consts, params, abi.FACTOR_BLOCK_SIZES[ftype], abi.FACTOR_RESIDUALS[ftype], want_jac, null_jac)

    def prior_evaluate(self, prior, state, want_jac=True):
        return self.vr.prior_evaluate(prior, state, want_jac)

    def marginalize(self, problem, state, flag):
        return self.vr.marginalize(problem, state, flag)


@pytest.fixture(scope="session")
def reference_code():
    import viw_ref
    if not viw_ref.available():
        pytest.skip("neither /root/reference nor a prebuilt oracle/_ref/libviw_ref.so")
    viw_ref.lib()
    return _ReferenceCode(viw_ref)
