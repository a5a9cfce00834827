"""Host-side mirror of snarkVM's `snarkvm_algorithms::msm::VariableBase` for BLS12-377 G1
[UPSTREAM algorithms/src/msm/variable_base/mod.rs; SURVEY.md 8a row a1], calling the CUDA library through
the C ABI.  Buffers use snarkVM's in-memory layout unchanged: bases are G1Affine images (104-byte stride,
Montgomery x @0, y @48, infinity flag @96), scalars are canonical BigInteger256, the result is the 144-byte
Jacobian G1Projective (X, Y, Z Montgomery).
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np

from . import _lib

AFFINE_STRIDE = 104
SCALAR_BYTES = 32
PROJECTIVE_BYTES = 144

try:  # torch is plumbing only (device memory / streams); numpy inputs work without it
    import torch
except Exception:  # pragma: no cover
    torch = None


def _is_cuda_tensor(x) -> bool:
    return torch is not None and isinstance(x, torch.Tensor) and x.is_cuda


def _np_ptr(a: np.ndarray) -> ctypes.c_void_p:
    return a.ctypes.data_as(ctypes.c_void_p)


def _stream_ptr() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev_out(device) -> "torch.Tensor":
    return torch.empty(PROJECTIVE_BYTES + 16, dtype=torch.uint8, device=device)[:PROJECTIVE_BYTES]


def _host_bytes(x) -> np.ndarray:
    if torch is not None and isinstance(x, torch.Tensor):
        x = x.numpy()                      # CPU tensor (e.g. pinned): zero-copy view
    return np.ascontiguousarray(x).view(np.uint8).reshape(-1)


class VariableBase:
    """`VariableBase::msm(bases, scalars) -> G1Projective`."""

    @staticmethod
    def msm(bases, scalars, stride: int = AFFINE_STRIDE):
        """sum_i scalars[i] * bases[i].

        Host inputs (numpy arrays or CPU/pinned torch tensors): uint8 bases [n * stride], scalars [n, 4] uint64
        or [n * 32] uint8 -> returns a numpy uint8[144]; the library does the H2D / D2H copies.
        CUDA torch tensors (device resident) -> returns a CUDA uint8[144] tensor, enqueued on the current
        stream.  Like snarkVM, uses min(len(bases), len(scalars)) terms."""
        L = _lib.lib()
        if _is_cuda_tensor(bases) or _is_cuda_tensor(scalars):
            if not (_is_cuda_tensor(bases) and _is_cuda_tensor(scalars)):
                raise TypeError("bases and scalars must both be CUDA tensors or both be host arrays")
            b = bases.contiguous().view(torch.uint8).reshape(-1)
            s = scalars.contiguous().view(torch.uint8).reshape(-1)
            n = min(b.numel() // stride, s.numel() // SCALAR_BYTES)
            out = _dev_out(b.device)
            _lib.check(L.b200_msm_g1_bls12_377_device(ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(b.data_ptr()), n,
                                                      ctypes.c_void_p(s.data_ptr()), stride, _stream_ptr()))
            return out
        b = _host_bytes(bases)
        s = _host_bytes(scalars)
        n = min(b.size // stride, s.size // SCALAR_BYTES)
        out = np.zeros(PROJECTIVE_BYTES, dtype=np.uint8)
        _lib.check(L.b200_msm_g1_bls12_377(_np_ptr(out), _np_ptr(b), n, _np_ptr(s), stride))
        return out


def msm_batch(bases, scalars, offsets, stride: int = AFFINE_STRIDE) -> np.ndarray:
    """Many independent small MSMs in one call: MSM m sums points [offsets[m], offsets[m + 1]).  Host buffers in,
    numpy uint8 [nmsm, 144] out.  This is the aggregation a batch verifier needs: the per-transaction linear
    combinations of a block's transactions (`KZG10::batch_check`) as ONE launch set."""
    b = _host_bytes(bases)
    s = _host_bytes(scalars)
    off = np.ascontiguousarray(offsets, dtype=np.uint64)
    nmsm = off.size - 1
    if nmsm < 0 or (nmsm >= 0 and int(off[-1]) * stride > b.size) or int(off[-1]) * SCALAR_BYTES > s.size:
        raise ValueError("offsets do not fit the point / scalar buffers")
    out = np.zeros((max(nmsm, 0), PROJECTIVE_BYTES), dtype=np.uint8)
    _lib.check(_lib.lib().b200_msm_batch_g1_bls12_377(_np_ptr(out), _np_ptr(b), _np_ptr(s), _np_ptr(off), nmsm, stride))
    return out


class ResidentBases:
    """Device-resident base set (e.g. the SRS powers_of_beta_g that KZG10::commit multiplies against on every
    call [UPSTREAM algorithms/src/polycommit/kzg10/mod.rs; SURVEY.md 8f rank 1])."""

    def __init__(self, bases, stride: int = AFFINE_STRIDE, tabulate: bool = False, window_bits: int = 0):
        """tabulate=True additionally stores 2^(c*w) * P_i for every window (nwin x the memory, e.g. 17.7 GB for 2^24
        points): later MSMs share one bucket set across windows and skip the fold."""
        L = _lib.lib()
        h = ctypes.c_uint64(0)
        if _is_cuda_tensor(bases):
            b = bases.contiguous().view(torch.uint8).reshape(-1)
            self.n = b.numel() // stride
            if tabulate:
                _lib.check(L.b200_msm_register_bases_tabulated_device(ctypes.c_void_p(b.data_ptr()), self.n, stride, window_bits,
                                                                      _stream_ptr(), ctypes.byref(h)))
            else:
                _lib.check(L.b200_msm_register_bases_device(ctypes.c_void_p(b.data_ptr()), self.n, stride, _stream_ptr(),
                                                            ctypes.byref(h)))
        else:
            b = _host_bytes(bases)
            self.n = b.size // stride
            if tabulate:
                _lib.check(L.b200_msm_register_bases_tabulated(_np_ptr(b), self.n, stride, window_bits, ctypes.byref(h)))
            else:
                _lib.check(L.b200_msm_register_bases(_np_ptr(b), self.n, stride, ctypes.byref(h)))
        self.handle: Optional[int] = h.value

    def msm(self, scalars):
        L = _lib.lib()
        if self.handle is None:
            raise ValueError("bases were released")
        if _is_cuda_tensor(scalars):
            s = scalars.contiguous().view(torch.uint8).reshape(-1)
            n = s.numel() // SCALAR_BYTES
            out = _dev_out(s.device)
            _lib.check(L.b200_msm_registered_device(ctypes.c_void_p(out.data_ptr()), self.handle,
                                                    ctypes.c_void_p(s.data_ptr()), n, _stream_ptr()))
            return out
        s = _host_bytes(scalars)
        n = s.size // SCALAR_BYTES
        out = np.zeros(PROJECTIVE_BYTES, dtype=np.uint8)
        _lib.check(L.b200_msm_registered(_np_ptr(out), self.handle, _np_ptr(s), n))
        return out

    def release(self) -> None:
        if self.handle is not None:
            _lib.check(_lib.lib().b200_msm_release_bases(self.handle))
            self.handle = None


def sum_projective(points) -> "torch.Tensor":
    """Sum of k Jacobian points given as a CUDA uint8 tensor [k, 144] (multi-GPU partial-sum combine)."""
    p = points.contiguous().view(torch.uint8).reshape(-1)
    k = p.numel() // PROJECTIVE_BYTES
    out = _dev_out(p.device)
    _lib.check(_lib.lib().b200_g1_sum_jacobian_device(ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(p.data_ptr()), k,
                                                       _stream_ptr()))
    return out


def synthetic_bases(n: int, seed: int = 1234567890, stride: int = AFFINE_STRIDE, device=None) -> "torch.Tensor":
    """n G1Affine images k_i * G (k_i = splitmix64(seed, i)), generated on the device."""
    device = device or torch.device("cuda", torch.cuda.current_device())
    out = torch.empty(n * stride, dtype=torch.uint8, device=device)
    _lib.check(_lib.lib().b200_g1_synthetic_bases_device(ctypes.c_void_p(out.data_ptr()), n, stride, seed, _stream_ptr()))
    return out


def g1_batch_normalize(points, stride: int = AFFINE_STRIDE):
    """`Projective::batch_normalization` + `to_affine`: k Jacobian images ([k, 144] uint8) -> G1Affine images
    [k, stride] (x, y Montgomery, infinity flag at byte 96).  numpy in -> numpy out; CUDA tensor in -> CUDA tensor."""
    L = _lib.lib()
    if _is_cuda_tensor(points):
        p = points.contiguous().view(torch.uint8).reshape(-1)
        k = p.numel() // PROJECTIVE_BYTES
        out = torch.empty((k, stride), dtype=torch.uint8, device=p.device)
        _lib.check(L.b200_g1_batch_normalize_device(ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(p.data_ptr()), k, stride, _stream_ptr()))
        return out
    p = _host_bytes(points)
    k = p.size // PROJECTIVE_BYTES
    out = np.zeros((k, stride), dtype=np.uint8)
    _lib.check(L.b200_g1_batch_normalize(_np_ptr(out), _np_ptr(p), k, stride))
    return out


def g1_compress(points):
    """Compressed G1Affine encodings [k, 48] of k Jacobian images ([k, 144] uint8): the bytes a commitment has inside a
    serialised Varuna proof (`ToBytes::write_le` of a G1Affine)."""
    L = _lib.lib()
    if _is_cuda_tensor(points):
        p = points.contiguous().view(torch.uint8).reshape(-1)
        k = p.numel() // PROJECTIVE_BYTES
        out = torch.empty((k, 48), dtype=torch.uint8, device=p.device)
        _lib.check(L.b200_g1_compress_device(ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(p.data_ptr()), k, _stream_ptr()))
        return out
    p = _host_bytes(points)
    k = p.size // PROJECTIVE_BYTES
    out = np.zeros((k, 48), dtype=np.uint8)
    _lib.check(L.b200_g1_compress(_np_ptr(out), _np_ptr(p), k))
    return out
