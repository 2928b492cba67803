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
