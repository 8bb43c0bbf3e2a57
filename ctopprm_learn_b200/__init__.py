"""ctopprm_learn_b200 -- B200-native drop-in for CTopPRM's data-parallel hot path.

The product is ``lib/libctopprm_b200.so`` (hand-written sm_100a CUDA behind the C ABI declared
in ``include/ctopprm_b200.h``) and the C++ class surface in ``include/ctopprm/``.  This Python
package is the thin ctypes binding that tests and ``bench.py`` drive it through; it never
computes anything itself and fails loudly when the library or a GPU is missing.
"""
from .capi import CtpError, DeviceMap, lib, lib_path  # noqa: F401
from . import synth  # noqa: F401

__all__ = ["CtpError", "DeviceMap", "lib", "lib_path", "synth"]
