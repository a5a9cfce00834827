"""ctypes loader for snarkos_b200/libsnarkos_b200.so (the C ABI declared in include/snarkos_b200.h).

There is deliberately no fallback of any kind: if the shared library is missing, or no CUDA device is
present when a compute entry point is called, an exception is raised."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsnarkos_b200.so")


class B200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"snarkos_b200 error {code}: {msg}")
        self.code = code


class b200_error_t(ctypes.Structure):
    _fields_ = [("code", ctypes.c_int32), ("msg", ctypes.c_char_p)]


# every symbol include/snarkos_b200.h declares: name -> (restype, argtypes)
_vp, _sz, _u32, _u64, _i = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int
SYMBOLS = {
    "b200_init": (b200_error_t, [_i]),
    "b200_shutdown": (None, []),
    "b200_abi_version": (_u32, []),
    "b200_release_scratch": (b200_error_t, []),
    "b200_set_option": (b200_error_t, [ctypes.c_char_p, ctypes.c_char_p]),
    "b200_get_counter": (b200_error_t, [ctypes.c_char_p, ctypes.POINTER(_u64)]),
    "b200_msm_submit": (b200_error_t, [_vp, _sz, _vp, _sz, ctypes.POINTER(_u64)]),
    "b200_msm_wait": (b200_error_t, [_u64, _vp]),
    "b200_msm_g1_bls12_377": (b200_error_t, [_vp, _vp, _sz, _vp, _sz]),
    "b200_msm_g1_bls12_377_device": (b200_error_t, [_vp, _vp, _sz, _vp, _sz, _vp]),
    "b200_msm_batch_g1_bls12_377": (b200_error_t, [_vp, _vp, _vp, _vp, _sz, _sz]),
    "b200_msm_batch_g1_bls12_377_device": (b200_error_t, [_vp, _vp, _vp, _vp, _sz, _sz, _sz, _vp]),
    "b200_msm_register_bases": (b200_error_t, [_vp, _sz, _sz, ctypes.POINTER(_u64)]),
    "b200_msm_register_bases_device": (b200_error_t, [_vp, _sz, _sz, _vp, ctypes.POINTER(_u64)]),
    "b200_msm_register_bases_tabulated": (b200_error_t, [_vp, _sz, _sz, _u32, ctypes.POINTER(_u64)]),
    "b200_msm_register_bases_tabulated_device": (b200_error_t, [_vp, _sz, _sz, _u32, _vp, ctypes.POINTER(_u64)]),
    "b200_msm_registered": (b200_error_t, [_vp, _u64, _vp, _sz]),
    "b200_msm_registered_device": (b200_error_t, [_vp, _u64, _vp, _sz, _vp]),
    "b200_msm_release_bases": (b200_error_t, [_u64]),
    "b200_kzg_commit": (b200_error_t, [_vp, _u64, _vp, _sz]),
    "b200_kzg_commit_device": (b200_error_t, [_vp, _u64, _vp, _sz, _vp]),
    "b200_kzg_commit_batch": (b200_error_t, [_vp, _u64, _vp, ctypes.POINTER(_u64), _sz]),
    "b200_kzg_commit_batch_device": (b200_error_t, [_vp, _u64, _vp, ctypes.POINTER(_u64), _sz, _vp]),
    "b200_kzg_open": (b200_error_t, [_vp, _u64, _vp, _sz, _vp, _vp]),
    "b200_kzg_open_device": (b200_error_t, [_vp, _u64, _vp, _sz, _vp, _vp, _vp]),
    "b200_g1_batch_normalize": (b200_error_t, [_vp, _vp, _sz, _sz]),
    "b200_g1_batch_normalize_device": (b200_error_t, [_vp, _vp, _sz, _sz, _vp]),
    "b200_g1_compress": (b200_error_t, [_vp, _vp, _sz]),
    "b200_g1_compress_device": (b200_error_t, [_vp, _vp, _sz, _vp]),
    "b200_fr_linear_combination_device": (b200_error_t, [_vp, _vp, ctypes.POINTER(_u64), _sz, _vp, _sz, _vp]),
    "b200_fr_divide_by_linear_device": (b200_error_t, [_vp, _vp, _sz, _vp, _vp, _vp]),
    "b200_msm_window_bits": (_u32, [_sz]),
    "b200_msm_affine_rounds": (_u32, [_sz]),
    "b200_msm_describe": (None, [_sz, ctypes.POINTER(_u32)]),
    "b200_g1_sum_jacobian_device": (b200_error_t, [_vp, _vp, _sz, _vp]),
    "b200_ntt_fr_bls12_377": (b200_error_t, [_vp, _u32, _sz, _sz, _i, _i]),
    "b200_ntt_fr_bls12_377_device": (b200_error_t, [_vp, _u32, _sz, _sz, _i, _i, _vp]),
    "b200_fr_vec_op_device": (b200_error_t, [_i, _vp, _vp, _vp, _vp, _sz, _i, _vp]),
    "b200_fr_batch_inverse_device": (b200_error_t, [_vp, _sz, _vp]),
    "b200_fr_divide_by_vanishing_on_coset_device": (b200_error_t, [_vp, _u32, _u32, _vp]),
    "b200_fr_mul_powers_device": (b200_error_t, [_vp, _u32, _i, _i, _u64, _u64, _u64, _u64, _vp]),
    "b200_fr_exchange_transpose_device": (b200_error_t, [_vp, ctypes.POINTER(_vp), _u32, _u32, _u64, _u64, _u32, _i, _i, _u64, _vp]),
    "b200_fr_exchange_transpose_part_device": (b200_error_t, [_vp, ctypes.POINTER(_vp), _u32, _u32, _u64, _u64, _u64, _u64, _u64, _u64, _u32, _i, _i, _u64, _u32, _vp]),
    "b200_ntt_rows_exchange_device": (b200_error_t, [_vp, ctypes.POINTER(_vp), _u32, _u32, _u64, _u32, _u32, _i, _i, _u64, _vp]),
    "b200_peer_buffer_alloc": (b200_error_t, [_sz, ctypes.POINTER(_vp), _vp]),
    "b200_peer_buffer_open": (b200_error_t, [_vp, ctypes.POINTER(_vp)]),
    "b200_peer_buffer_close": (b200_error_t, [_vp]),
    "b200_peer_buffer_free": (b200_error_t, [_vp]),
    "b200_g1_synthetic_bases_device": (b200_error_t, [_vp, _sz, _sz, _u64, _vp]),
    "b200_debug_field_op": (b200_error_t, [_i, _vp, _vp, _vp, _sz]),
    "b200_debug_g1_op": (b200_error_t, [_i, _vp, _vp, _vp, _sz, _sz]),
    "b200_debug_microbench": (b200_error_t, [_i, _u32, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double)]),
    "b200_profile_begin": (None, []),
    "b200_profile_end": (b200_error_t, [ctypes.c_char_p, _sz]),
    "b200_kernel_launch_count": (_u64, []),
}

_LIB = None


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C snarkos_b200/csrc). There is no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)           # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def check(err: b200_error_t) -> None:
    if err.code != 0:
        raise B200Error(err.code, (err.msg or b"?").decode())


def init(device: int = -1) -> None:
    check(lib().b200_init(device))


class profile:
    """with profile() as p: ...calls...; p.stages -> [(stage, ms), ...] measured with CUDA events on the stream."""

    def __enter__(self):
        self.stages = []
        lib().b200_profile_begin()
        return self

    def __exit__(self, *exc):
        buf = ctypes.create_string_buffer(1 << 16)
        check(lib().b200_profile_end(buf, len(buf)))
        for item in buf.value.decode().split(";"):
            if item:
                k, v = item.split("=")
                self.stages.append((k, float(v)))
        return False

    def totals(self):
        out = {}
        for k, v in self.stages:
            out[k] = out.get(k, 0.0) + v
        return out


def set_option(key: str, value) -> None:
    """b200_set_option: change a tuning knob at run time (the B200_* environment is only read on first use)."""
    check(lib().b200_set_option(key.encode(), str(value).encode()))


def counter(name: str) -> int:
    v = _u64(0)
    check(lib().b200_get_counter(name.encode(), ctypes.byref(v)))
    return int(v.value)


def release_scratch() -> None:
    """b200_release_scratch: hand all cached scratch memory back to the driver (and trim its pool)"""
    check(lib().b200_release_scratch())


def kernel_launch_count() -> int:
    return int(lib().b200_kernel_launch_count())
