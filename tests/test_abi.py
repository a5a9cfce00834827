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


def test_rust_binding_covers_the_header():
    """rust/snarkvm-algorithms-b200/src/lib.rs cannot be compiled here (no toolchain): at least every function the
    header declares is bound there with the same number of arguments"""
    hdr = open(os.path.join(ROOT, "include", "snarkos_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    rs = open(os.path.join(ROOT, "rust", "snarkvm-algorithms-b200", "src", "lib.rs")).read()
    rs = re.sub(r"//[^\n]*", "", rs)

    def arity(args: str) -> int:
        args = args.strip()
        return 0 if args in ("", "void") else args.count(",") + 1

    c_decl = {m.group(1): arity(m.group(2)) for m in re.finditer(r"\b(b200_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", hdr)}
    rs_decl = {m.group(1): arity(m.group(2)) for m in re.finditer(r"pub fn (b200_[a-z0-9_]+)\s*\(([^)]*)\)", rs)}
    assert set(declared_functions()) == set(c_decl)
    missing = sorted(set(c_decl) - set(rs_decl))
    assert not missing, f"not bound in the -sys crate: {missing}"
    extra = sorted(set(rs_decl) - set(c_decl))
    assert not extra, f"bound in the -sys crate but not declared in the header: {extra}"
    wrong = {n: (c_decl[n], rs_decl[n]) for n in c_decl if c_decl[n] != rs_decl[n]}
    assert not wrong, f"argument count differs (header, rust): {wrong}"
