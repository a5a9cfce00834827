"""The C-ABI boundary without a GPU: the shared library loads, exports every function include/snarkos_b200.h declares
(the reference-side `-sys` crate binds exactly these), the Python symbol table covers them, and -- on a box without a
CUDA device -- every compute entry fails loudly instead of falling back to a CPU path."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "snarkos_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from snarkos_b200 import _lib
    lib = _lib.lib()
    names = declared_functions()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/snarkos_b200.h but not exported: {missing}"
    unbound = [n for n in names if n not in _lib.SYMBOLS]
    assert not unbound, f"declared but without a ctypes signature in snarkos_b200/_lib.py: {unbound}"
    assert lib.b200_abi_version() == 2


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from snarkos_b200 import _lib
    lib = _lib.lib()
    out = np.zeros(144, dtype=np.uint8)
    pts = np.zeros(104, dtype=np.uint8)
    sc = np.zeros(32, dtype=np.uint8)
    e = lib.b200_msm_g1_bls12_377(out.ctypes.data_as(ctypes.c_void_p), pts.ctypes.data_as(ctypes.c_void_p), 1,
                                  sc.ctypes.data_as(ctypes.c_void_p), 104)
    assert e.code != 0 and b"no CPU fallback" in e.msg
    data = np.zeros(4, dtype=np.uint64)
    e = lib.b200_ntt_fr_bls12_377(data.ctypes.data_as(ctypes.c_void_p), 0, 1, 1, 0, 0)
    assert e.code != 0
    with pytest.raises(_lib.B200Error):
        _lib.check(lib.b200_init(-1))


def test_options_and_counters_need_no_device():
    """tuning knobs are plain process state (environment read once, b200_set_option afterwards)"""
    from snarkos_b200 import _lib
    import snarkos_b200 as S
    S.set_option("msm_window_bits", 17)
    S.set_option("msm_window_bits", 0)
    S.set_option("ntt_plan", "8,8,8")
    S.set_option("ntt_plan", "")
    with pytest.raises(_lib.B200Error):
        S.set_option("no_such_knob", 1)
    assert S.counter("kernel_launches") >= 0
    assert S.counter("msm_xyzz_fallbacks") >= 0
    with pytest.raises(_lib.B200Error):
        S.counter("no_such_counter")
