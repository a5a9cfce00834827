"""Host-side mirror of snarkVM's `snarkvm_algorithms::fft::EvaluationDomain<Fr>` for BLS12-377
[UPSTREAM algorithms/src/fft/domain.rs; SURVEY.md 8a rows a3-a6], calling the CUDA library through the C ABI.
Data are Fr elements in snarkVM's in-memory form: 32 bytes, 4 x u64 little-endian Montgomery limbs.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib

FR_TWO_ADICITY = 47
FR_BYTES = 32
# BLS12-377 scalar field: r, the multiplicative generator snarkVM uses, and its 2^47-th root of unity
FR_MODULUS = 0x12AB655E9A2CA55660B44D1E5C37B00159AA76FED00000010A11800000000001
FR_GENERATOR = 22
FR_TWO_ADIC_ROOT_OF_UNITY = pow(FR_GENERATOR, (FR_MODULUS - 1) >> FR_TWO_ADICITY, FR_MODULUS)


def fr_to_mont_limbs(x: int) -> np.ndarray:
    """canonical integer -> snarkVM's in-memory Fr: 4 x u64 little-endian limbs of x * 2^256 mod r"""
    v = (x % FR_MODULUS) * (1 << 256) % FR_MODULUS
    return np.array([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)
MAX_LOG_SIZE = 28            # library limit (2^28 elements = 8 GiB); snarkVM's own limit is the 2-adicity, 47

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


def _is_cuda_tensor(x) -> bool:
    return torch is not None and isinstance(x, torch.Tensor) and x.is_cuda


class EvaluationDomain:
    """`EvaluationDomain::new(num_coeffs)`: the smallest power-of-two domain holding num_coeffs."""

    def __init__(self, num_coeffs: int):
        size, log = 1, 0
        while size < max(int(num_coeffs), 1):
            size <<= 1
            log += 1
        if log > FR_TWO_ADICITY:
            raise ValueError("EvaluationDomain::new: domain larger than 2^47 (snarkVM returns None)")
        self.size = size
        self.log_size_of_group = log
        # the remaining fields of snarkVM's struct, as canonical integers (fr_to_mont_limbs gives the in-memory form);
        # the CUDA library derives the same values for its twiddle tables (ntt.cu: get_tables)
        self.size_as_field_element = size % FR_MODULUS
        self.size_inv = pow(size, -1, FR_MODULUS)
        self.group_gen = pow(FR_TWO_ADIC_ROOT_OF_UNITY, 1 << (FR_TWO_ADICITY - log), FR_MODULUS)
        self.group_gen_inv = pow(self.group_gen, -1, FR_MODULUS)
        self.generator_inv = pow(FR_GENERATOR, -1, FR_MODULUS)

    def elements(self):
        """`EvaluationDomain::elements()`: 1, g, g^2, ... (canonical integers)"""
        x = 1
        for _ in range(self.size):
            yield x
            x = x * self.group_gen % FR_MODULUS

    def evaluate_vanishing_polynomial(self, tau: int) -> int:
        """Z_H(tau) = tau^size - 1"""
        return (pow(tau, self.size, FR_MODULUS) - 1) % FR_MODULUS

    # out-of-place forms (`fft`, `ifft`, `coset_fft`, `coset_ifft`): the input is left alone
    def fft(self, coeffs):
        return self._run(self._copy(coeffs), 0, 0)

    def ifft(self, evals):
        return self._run(self._copy(evals), 1, 0)

    def coset_fft(self, coeffs):
        return self._run(self._copy(coeffs), 0, 1)

    def coset_ifft(self, evals):
        return self._run(self._copy(evals), 1, 1)

    @staticmethod
    def _copy(x):
        if torch is not None and isinstance(x, torch.Tensor):
            return x.clone()
        return np.array(x, copy=True)

    @classmethod
    def new(cls, num_coeffs: int) -> "EvaluationDomain":
        return cls(num_coeffs)

    # ---- the four in-place entry points -------------------------------------------------------
    def fft_in_place(self, coeffs):
        return self._run(coeffs, 0, 0)

    def ifft_in_place(self, evals):
        return self._run(evals, 1, 0)

    def coset_fft_in_place(self, coeffs):
        return self._run(coeffs, 0, 1)

    def coset_ifft_in_place(self, evals):
        return self._run(evals, 1, 1)

    # ---- implementation -----------------------------------------------------------------------
    def _run(self, data, direction: int, coset: int):
        """Host data, [m, 4] uint64 with m <= size: zero-padded to the domain like `coeffs.resize(size, 0)`; the
        transformed array is returned (numpy cannot grow in place).  Host [batch, size, 4] transforms every row
        (CPU / pinned torch tensors of that shape are transformed truly in place).
        CUDA tensors holding a whole number of domain-sized polynomials are transformed in place on the
        current stream and returned."""
        if self.log_size_of_group > MAX_LOG_SIZE:
            raise _lib.B200Error(-3, "domain larger than the library limit 2^28")
        L = _lib.lib()
        n = self.size
        if _is_cuda_tensor(data):
            t = data
            if not t.is_contiguous():
                raise ValueError("device data must be contiguous")
            nbytes = t.numel() * t.element_size()
            if nbytes == 0 or nbytes % (n * FR_BYTES):
                raise ValueError("device data must hold a whole number of domain-sized polynomials")
            batch = nbytes // (n * FR_BYTES)
            _lib.check(L.b200_ntt_fr_bls12_377_device(ctypes.c_void_p(t.data_ptr()), self.log_size_of_group, batch, n,
                                                      direction, coset,
                                                      ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
            return t
        if torch is not None and isinstance(data, torch.Tensor):
            # CPU (ideally pinned) tensor: in place, no intermediate copy
            nbytes = data.numel() * data.element_size()
            if not data.is_contiguous() or nbytes == 0 or nbytes % (n * FR_BYTES):
                raise ValueError("host tensor must be contiguous and hold whole domain-sized polynomials")
            batch = nbytes // (n * FR_BYTES)
            _lib.check(L.b200_ntt_fr_bls12_377(ctypes.c_void_p(data.data_ptr()), self.log_size_of_group, batch, n,
                                               direction, coset))
            return data
        a = np.asarray(data)
        if a.dtype != np.uint64:
            a = a.view(np.uint64)
        if a.ndim == 2:
            if a.shape[0] > n:
                raise ValueError("input longer than the domain")       # snarkVM: caller bug
            buf = np.zeros((n, 4), dtype=np.uint64)
            buf[: a.shape[0]] = a
            batch = 1
        elif a.ndim == 3:
            if a.shape[1] != n:
                raise ValueError("batched input must already have the domain size")
            buf = np.ascontiguousarray(a).copy()
            batch = a.shape[0]
        else:
            raise ValueError("expected [m, 4] or [batch, size, 4] uint64 limbs")
        _lib.check(L.b200_ntt_fr_bls12_377(buf.ctypes.data_as(ctypes.c_void_p), self.log_size_of_group, batch, n,
                                           direction, coset))
        return buf
