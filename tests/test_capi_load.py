"""The C-ABI library loads and exports every symbol include/ctopprm_b200.h declares
(no compute without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, has_gpu
from ctopprm_learn_b200 import capi


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "ctopprm_b200.h")).read()
    return sorted(set(re.findall(r"^(?:int|const char \*)\s*(ctp_[a-z0-9_]+)\s*\(", src, flags=re.M)))


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    names = declared_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(L, name), f"{name} declared in include/ctopprm_b200.h but not exported"
    assert set(names) == set(capi.exported_symbols())
    assert L.ctp_version() >= 100


def test_no_cpu_fallback_without_device():
    if has_gpu():
        pytest.skip("GPU present")
    vox = np.zeros((4, 4, 4), dtype=np.float32)
    with pytest.raises(capi.CtpError) as ei:
        capi.DeviceMap(vox, np.zeros(3), np.ones(3))
    assert ei.value.code in (-2, -1)  # CUDA error / no such device; never a silent CPU path


def test_argument_validation_precedes_device_use():
    L = capi.lib()
    h = ctypes.c_void_p()
    assert L.ctp_map_create(None, 8, None, None, 0, 1, ctypes.byref(h)) == -1
    assert b"null" in L.ctp_last_error()
    assert L.ctp_clearance(None, None, 0, None) == -1
