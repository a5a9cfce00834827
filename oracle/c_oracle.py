"""ctypes binding of oracle/liboracle.so (TEST INFRASTRUCTURE ONLY -- see oracle/oracle.c header).

numpy arrays of little-endian u64 limbs in, numpy arrays out.  Used by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs; never by the
product package.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        vp, sz, i32 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
        for name in ("oracle_fr_mul", "oracle_fq_mul", "oracle_fr_add", "oracle_fr_sub",
                     "oracle_fq_add", "oracle_fq_sub"):
            getattr(L, name).argtypes = [vp, vp, vp, sz]
            getattr(L, name).restype = None
        for name in ("oracle_fq_inv", "oracle_fr_inv", "oracle_fr_to_mont", "oracle_fr_from_mont"):
            getattr(L, name).argtypes = [vp, vp, sz]
            getattr(L, name).restype = None
        L.oracle_g1_to_affine.argtypes = [vp, vp, sz, sz]
        L.oracle_g1_to_affine.restype = None
        L.oracle_g1_compress.argtypes = [vp, vp, sz]
        L.oracle_g1_compress.restype = None
        L.oracle_g1_mul_u64.argtypes = [vp, vp, vp, sz, sz]
        L.oracle_g1_mul_u64.restype = None
        L.oracle_g1_is_on_curve.argtypes = [vp]
        L.oracle_g1_is_on_curve.restype = i32
        L.oracle_msm.argtypes = [vp, vp, sz, sz, vp, i32]
        L.oracle_msm.restype = None
        L.oracle_msm_batched.argtypes = [vp, vp, sz, sz, vp, i32]
        L.oracle_msm_batched.restype = None
        L.oracle_msm_many.argtypes = [vp, vp, vp, vp, sz, sz, i32]
        L.oracle_msm_many.restype = None
        L.oracle_g1_sequence.argtypes = [vp, ctypes.c_uint64, sz, sz, i32]
        L.oracle_g1_sequence.restype = None
        L.oracle_msm_window_bits.argtypes = [sz]
        L.oracle_msm_window_bits.restype = i32
        L.oracle_ntt.argtypes = [vp, i32, sz, sz, i32, i32, i32]
        L.oracle_ntt.restype = None
        L.oracle_num_threads.restype = i32
        _LIB = L
    return _LIB


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def _binop(name, a, b, limbs):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, limbs)
    b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, limbs)
    out = np.empty_like(a)
    getattr(lib(), name)(_p(out), _p(a), _p(b), a.shape[0])
    return out


def fr_mul(a, b): return _binop("oracle_fr_mul", a, b, 4)
def fr_add(a, b): return _binop("oracle_fr_add", a, b, 4)
def fr_sub(a, b): return _binop("oracle_fr_sub", a, b, 4)
def fq_mul(a, b): return _binop("oracle_fq_mul", a, b, 6)
def fq_add(a, b): return _binop("oracle_fq_add", a, b, 6)
def fq_sub(a, b): return _binop("oracle_fq_sub", a, b, 6)


def _unop(name, a, limbs):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, limbs)
    out = np.empty_like(a)
    getattr(lib(), name)(_p(out), _p(a), a.shape[0])
    return out


def fq_inv(a): return _unop("oracle_fq_inv", a, 6)
def fr_inv(a): return _unop("oracle_fr_inv", a, 4)
def fr_to_mont(a): return _unop("oracle_fr_to_mont", a, 4)
def fr_from_mont(a): return _unop("oracle_fr_from_mont", a, 4)


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def msm(bases: np.ndarray, scalars: np.ndarray, stride: int = 104, nthreads: int = 0) -> np.ndarray:
    """bases: uint8 [n*stride] snarkVM G1Affine images; scalars: uint64 [n,4] canonical.
    Returns the 144-byte Jacobian result as uint8[144]."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8).reshape(-1)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    n = scalars.shape[0]
    assert bases.size >= n * stride
    out = np.zeros(144, dtype=np.uint8)
    lib().oracle_msm(_p(out), _p(bases), n, stride, _p(scalars), nthreads)
    return out


def msm_batched(bases: np.ndarray, scalars: np.ndarray, stride: int = 104, nthreads: int = 0) -> np.ndarray:
    """snarkVM batched::msm restated (affine pair additions sharing one inversion per batch) -- the variant
    VariableBase::msm runs for BLS12-377 G1; same interface as msm()."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8).reshape(-1)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    n = scalars.shape[0]
    assert bases.size >= n * stride
    out = np.zeros(144, dtype=np.uint8)
    lib().oracle_msm_batched(_p(out), _p(bases), n, stride, _p(scalars), nthreads)
    return out


def msm_many(bases: np.ndarray, scalars: np.ndarray, offsets, stride: int = 104, nthreads: int = 0) -> np.ndarray:
    """independent MSMs over [offsets[m], offsets[m+1]), one task per MSM (the reference verifies a block's transactions
    rayon-parallel); returns uint8 [nmsm, 144]"""
    bases = np.ascontiguousarray(bases, dtype=np.uint8).reshape(-1)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    off = np.ascontiguousarray(offsets, dtype=np.uint64)
    out = np.zeros((off.size - 1, 144), dtype=np.uint8)
    lib().oracle_msm_many(_p(out), _p(bases), _p(scalars), _p(off), off.size - 1, stride, nthreads)
    return out


def g1_sequence(k0: int, n: int, stride: int = 104, nthreads: int = 0) -> np.ndarray:
    """(k0 + i) * G for i < n as affine images [n, stride]: n distinct points at ~1 us each"""
    out = np.zeros((n, stride), dtype=np.uint8)
    lib().oracle_g1_sequence(_p(out), k0, n, stride, nthreads)
    return out


def g1_to_affine(jac: np.ndarray, stride: int = 104) -> np.ndarray:
    jac = np.ascontiguousarray(jac, dtype=np.uint8).reshape(-1, 144)
    out = np.zeros((jac.shape[0], stride), dtype=np.uint8)
    lib().oracle_g1_to_affine(_p(out), _p(jac), jac.shape[0], stride)
    return out


def g1_compress(jac: np.ndarray) -> np.ndarray:
    """Jacobian images [n, 144] -> compressed G1Affine encodings uint8 [n, 48] (Varuna proof wire format of a commitment)"""
    jac = np.ascontiguousarray(jac, dtype=np.uint8).reshape(-1, 144)
    out = np.zeros((jac.shape[0], 48), dtype=np.uint8)
    lib().oracle_g1_compress(_p(out), _p(jac), jac.shape[0])
    return out


def g1_mul_u64(base_affine: np.ndarray, k: np.ndarray, stride: int = 104) -> np.ndarray:
    k = np.ascontiguousarray(k, dtype=np.uint64).reshape(-1)
    base_affine = np.ascontiguousarray(base_affine, dtype=np.uint8)
    out = np.zeros((k.size, stride), dtype=np.uint8)
    lib().oracle_g1_mul_u64(_p(out), _p(base_affine), _p(k), k.size, stride)
    return out


def g1_is_on_curve(affine: np.ndarray) -> bool:
    affine = np.ascontiguousarray(affine, dtype=np.uint8)
    return bool(lib().oracle_g1_is_on_curve(_p(affine)))


def ntt(data: np.ndarray, log_n: int, batch: int = 1, direction: int = 0, coset: int = 0,
        nthreads: int = 0, batch_stride: int | None = None) -> np.ndarray:
    """data: uint64 [batch * stride, 4] Montgomery Fr; returns a transformed copy."""
    out = np.array(data, dtype=np.uint64, copy=True).reshape(-1, 4)
    n = 1 << log_n
    stride = n if batch_stride is None else batch_stride
    assert out.shape[0] >= (batch - 1) * stride + n
    lib().oracle_ntt(_p(out), log_n, batch, stride, direction, coset, nthreads)
    return out
