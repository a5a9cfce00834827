"""Host-side mirror of the MSM-bearing part of snarkVM's `polycommit::kzg10::KZG10` [UPSTREAM
algorithms/src/polycommit/kzg10/mod.rs; SURVEY.md 8a row a9, 8f rank 1]: `commit` / `commit_lagrange` against powers
that stay resident in HBM.  The SRS is fixed for the lifetime of the process, so it is uploaded (and re-packed) once;
each commit moves only the polynomial."""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np

from . import _lib
from .msm import PROJECTIVE_BYTES, ResidentBases, _dev_out, _host_bytes, _is_cuda_tensor, _np_ptr, _stream_ptr, sum_projective

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


class Powers:
    """`Powers { powers_of_beta_g, powers_of_beta_times_gamma_g }` (or the Lagrange-basis equivalents) on the device."""

    def __init__(self, powers_of_beta_g, powers_of_beta_times_gamma_g=None, stride: int = 104, tabulate: bool = False):
        """tabulate=True also stores the window multiples 2^(c*w) * beta^i G of the (fixed) SRS in HBM, which makes every
        commit cheaper (one bucket set, no fold) at nwin x the memory."""
        self.powers_of_beta_g = ResidentBases(powers_of_beta_g, stride, tabulate=tabulate)
        self.powers_of_beta_times_gamma_g = (ResidentBases(powers_of_beta_times_gamma_g, stride)
                                             if powers_of_beta_times_gamma_g is not None else None)

    def release(self) -> None:
        self.powers_of_beta_g.release()
        if self.powers_of_beta_times_gamma_g is not None:
            self.powers_of_beta_times_gamma_g.release()


def _commit_one(rb: ResidentBases, coeffs):
    L = _lib.lib()
    if _is_cuda_tensor(coeffs):
        c = coeffs.contiguous().view(torch.uint8).reshape(-1)
        n = c.numel() // 32
        out = _dev_out(c.device)
        _lib.check(L.b200_kzg_commit_device(ctypes.c_void_p(out.data_ptr()), rb.handle, ctypes.c_void_p(c.data_ptr()), n, _stream_ptr()))
        return out
    c = _host_bytes(coeffs)
    out = np.zeros(PROJECTIVE_BYTES, dtype=np.uint8)
    _lib.check(L.b200_kzg_commit(_np_ptr(out), rb.handle, _np_ptr(c), c.size // 32))
    return out


class KZG10:
    @staticmethod
    def commit(powers: Powers, polynomial_coeffs, blinding_coeffs: Optional[object] = None):
        """`KZG10::commit(powers, polynomial, hiding_bound, rng)`'s group arithmetic: sum_i coeffs[i] * beta^i G, plus
        sum_i blinding[i] * gamma beta^i G when a blinding polynomial is given (the caller draws it, as snarkVM's
        `KZGRandomness::rand` does on the CPU).  Coefficients are Montgomery Fr limbs ([n, 4] uint64 / CUDA tensors).
        Returns the G1Projective image (numpy uint8[144] for host inputs, CUDA tensor for device inputs)."""
        if blinding_coeffs is not None and powers.powers_of_beta_times_gamma_g is None:
            raise ValueError("hiding commitment needs powers_of_beta_times_gamma_g")
        c = _commit_one(powers.powers_of_beta_g, polynomial_coeffs)
        if blinding_coeffs is None:
            return c
        r = _commit_one(powers.powers_of_beta_times_gamma_g, blinding_coeffs)
        if _is_cuda_tensor(c):
            return sum_projective(torch.stack([c, r]))
        both = torch.from_numpy(np.stack([c, r])).cuda()
        out = sum_projective(both)
        torch.cuda.synchronize()
        return out.cpu().numpy()

    @staticmethod
    def commit_batch(powers: Powers, polynomials):
        """k commitments against the same powers in ONE launch set (`SonicKZG10::commit` over the labeled polynomials
        of a Varuna round [UPSTREAM algorithms/src/polycommit/sonic_pc/mod.rs]): `polynomials` is a list of Montgomery
        coefficient arrays ([n_m, 4] uint64, or CUDA tensors -- all of one kind).  Returns [k, 144] uint8 (numpy for
        host inputs, a CUDA tensor for device inputs)."""
        L = _lib.lib()
        rb = powers.powers_of_beta_g
        k = len(polynomials)
        if k == 0:
            return np.zeros((0, PROJECTIVE_BYTES), dtype=np.uint8)
        if _is_cuda_tensor(polynomials[0]):
            dev = polynomials[0].device
            flat = [p.contiguous().reshape(-1).view(torch.uint8) if p.numel() else torch.empty(0, dtype=torch.uint8, device=dev)
                    for p in polynomials]
            lens = [f.numel() // 32 for f in flat]
            c = torch.cat(flat) if k > 1 else flat[0]
            off = (ctypes.c_uint64 * (k + 1))(*np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64).tolist())
            out = torch.empty((k, PROJECTIVE_BYTES), dtype=torch.uint8, device=c.device)
            _lib.check(L.b200_kzg_commit_batch_device(ctypes.c_void_p(out.data_ptr()), rb.handle,
                                                      ctypes.c_void_p(c.data_ptr()), off, k, _stream_ptr()))
            return out
        flat = [_host_bytes(p) for p in polynomials]
        lens = [f.size // 32 for f in flat]
        c = np.ascontiguousarray(np.concatenate(flat)) if k > 1 else flat[0]
        off = (ctypes.c_uint64 * (k + 1))(*np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64).tolist())
        out = np.zeros((k, PROJECTIVE_BYTES), dtype=np.uint8)
        _lib.check(L.b200_kzg_commit_batch(_np_ptr(out), rb.handle, _np_ptr(c), off, k))
        return out

    @staticmethod
    def open(powers: Powers, polynomial_coeffs, point, blinding_coeffs: Optional[object] = None):
        """`KZG10::open(powers, polynomial, point, rand)`'s arithmetic: the commitment to the witness polynomial
        (p(X) - p(z)) / (X - z) -- plus, for a hiding commitment, the witness of the blinding polynomial against the
        gamma powers -- and the evaluations.  Host inputs: Montgomery [n, 4] uint64 coefficients, [4] uint64 point.
        Returns (w: uint8[144], p(z): uint64[4], random_v: uint64[4] or None)."""
        L = _lib.lib()
        if _is_cuda_tensor(polynomial_coeffs):
            # device-resident form: coefficients and point are CUDA tensors; returns (w [144] uint8, p(z) [1, 4] int64) on
            # the device, nothing synchronises (the hiding part is a second call by the caller)
            c = polynomial_coeffs.contiguous()
            n = c.numel() * c.element_size() // 32
            out = _dev_out(c.device)
            ev = torch.empty((1, 4), dtype=torch.int64, device=c.device)
            _lib.check(L.b200_kzg_open_device(ctypes.c_void_p(out.data_ptr()), powers.powers_of_beta_g.handle,
                                              ctypes.c_void_p(c.data_ptr()), n, ctypes.c_void_p(point.data_ptr()),
                                              ctypes.c_void_p(ev.data_ptr()), _stream_ptr()))
            return out, ev, None
        z = np.ascontiguousarray(point, dtype=np.uint64).reshape(4)

        def one(rb, coeffs):
            c = _host_bytes(coeffs)
            out = np.zeros(PROJECTIVE_BYTES, dtype=np.uint8)
            ev = np.zeros(4, dtype=np.uint64)
            _lib.check(L.b200_kzg_open(_np_ptr(out), rb.handle, _np_ptr(c), c.size // 32, _np_ptr(z), _np_ptr(ev)))
            return out, ev

        w, v = one(powers.powers_of_beta_g, polynomial_coeffs)
        if blinding_coeffs is None:
            return w, v, None
        if powers.powers_of_beta_times_gamma_g is None:
            raise ValueError("hiding opening needs powers_of_beta_times_gamma_g")
        wr, rv = one(powers.powers_of_beta_times_gamma_g, blinding_coeffs)
        both = torch.from_numpy(np.stack([w, wr])).cuda()
        out = sum_projective(both)
        torch.cuda.synchronize()
        return out.cpu().numpy(), v, rv

    # commit_lagrange is the same sum against the Lagrange-basis powers with evaluations as "coefficients"
    commit_lagrange = commit
