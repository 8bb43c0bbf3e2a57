import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session", params=["cube", "rect"])
def golden(request):
    return dict(np.load(os.path.join(GOLDEN, f"ref_{request.param}.npz")))


@pytest.fixture(scope="session")
def orc():
    from oracle import orc as o
    o.build()
    return o


def has_gpu():
    try:
        import ctypes
        from ctopprm_learn_b200 import capi
        n = ctypes.c_int(0)
        return capi.lib().ctp_device_count(ctypes.byref(n)) == 0 and n.value > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def c1_map():
    from ctopprm_learn_b200 import synth
    vox, c, e = synth.make_config_map("C1")
    return vox, c, e
