"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the BLS12-377 MSM / Fr-NTT hot path.

This module restates, with Python big integers, the arithmetic that snarkOS reaches
through snarkVM (git dep `snarkvm` rev dea322b, crates snarkvm-{fields,curves,algorithms}
1.0.0 -- /root/reference/Cargo.toml:44-49, Cargo.lock:3578-3606).  The snarkVM sources are
NOT vendored in /root/reference and there is no Rust toolchain in the build image, so the
oracle follows the *published* algorithms and snarkVM's data conventions as recorded in
SURVEY.md section 8 / appendix A:

  * Fr / Fq Montgomery form                       [UPSTREAM fields/src/fp_256.rs, fp_384.rs]
  * y^2 = x^3 + 1 short-Weierstrass G1             [UPSTREAM curves/src/bls12_377/g1.rs]
  * VariableBase::msm  = sum_i s_i * P_i           [UPSTREAM algorithms/src/msm/variable_base/mod.rs]
  * EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}_in_place, natural order in/out
                                                   [UPSTREAM algorithms/src/fft/domain.rs]

PARITY STATUS: "parity unpinned" at the snarkOS boundary (the reference holds no
known-answer vectors for this path); the oracle is pinned against the
offline-derived known-answer vectors of SURVEY.md appendix A (tests/golden/kat.json) and
against first-principles identities (r*G = O, O(n^2) DFT, naive double-and-add).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product path (snarkos_b200/) never does.
"""
from __future__ import annotations

import hashlib
from typing import Iterable, List, Optional, Sequence, Tuple

# --------------------------------------------------------------------------------------
# Parameters (derived from the BLS12-377 seed; SURVEY.md section 8c "verified constants")
# --------------------------------------------------------------------------------------
X_SEED = 0x8508C00000000001
R_MOD = X_SEED**4 - X_SEED**2 + 1                      # scalar field modulus r (253 bits)
P_MOD = ((X_SEED - 1) ** 2 * R_MOD) // 3 + X_SEED      # base field modulus p (377 bits)

assert R_MOD == 0x12AB655E9A2CA55660B44D1E5C37B00159AA76FED00000010A11800000000001
assert P_MOD == int(
    "1ae3a4617c510eac63b05c06ca1493b1a22d9f300f5138f1ef3622fba094800170b5d44300000008508c00000000001", 16)

FR_BITS, FQ_BITS = 253, 377
FR_R = (1 << 256) % R_MOD          # Montgomery radix for Fr (4 x u64 limbs)
FQ_R = (1 << 384) % P_MOD          # Montgomery radix for Fq (6 x u64 limbs)
FR_R_INV = pow(FR_R, -1, R_MOD)
FQ_R_INV = pow(FQ_R, -1, P_MOD)

FR_TWO_ADICITY = 47
FR_GENERATOR = 22                  # multiplicative generator of Fr* used by snarkVM
FR_TWO_ADIC_ROOT = pow(FR_GENERATOR, (R_MOD - 1) >> FR_TWO_ADICITY, R_MOD)

G1_B = 1                           # y^2 = x^3 + 1
G1_GEN = (
    81937999373150964239938255573465948239988671502647976594219695644855304257327692006745978603320413799295628339695,
    241266749859715473739788878240585681733927191168601896383759122102112907357779751001206799952863815012735208165030,
)

AFFINE_STRIDE = 104                # snarkVM G1Affine {x@0, y@48, infinity@96}


# --------------------------------------------------------------------------------------
# Montgomery form + byte layouts (what crosses the C ABI)
# --------------------------------------------------------------------------------------
def fr_to_mont(a: int) -> int:
    return (a * FR_R) % R_MOD


def fr_from_mont(a: int) -> int:
    return (a * FR_R_INV) % R_MOD


def fq_to_mont(a: int) -> int:
    return (a * FQ_R) % P_MOD


def fq_from_mont(a: int) -> int:
    return (a * FQ_R_INV) % P_MOD


def mont_mul(a: int, b: int, mod: int, rinv: int) -> int:
    """Montgomery product a*b*R^-1 mod m (what Fp256/Fp384::mul_assign computes on Montgomery limbs)."""
    return (a * b * rinv) % mod


def fr_bytes_mont(a: int) -> bytes:
    """canonical Fr value -> 32 B little-endian Montgomery limbs (snarkVM in-memory Fp256)."""
    return fr_to_mont(a).to_bytes(32, "little")


def fr_from_bytes_mont(b: bytes) -> int:
    v = int.from_bytes(b, "little")
    assert v < R_MOD, "non-reduced Fr limb pattern"
    return fr_from_mont(v)


def scalar_bytes(a: int) -> bytes:
    """canonical scalar -> BigInteger256 (32 B little-endian, NOT Montgomery)."""
    assert 0 <= a < R_MOD
    return a.to_bytes(32, "little")


def affine_bytes(pt: Optional[Tuple[int, int]], stride: int = AFFINE_STRIDE) -> bytes:
    """affine point (canonical ints) or None (=infinity) -> snarkVM G1Affine memory image."""
    out = bytearray(stride)
    if pt is None:
        out[96] = 1
    else:
        out[0:48] = fq_to_mont(pt[0]).to_bytes(48, "little")
        out[48:96] = fq_to_mont(pt[1]).to_bytes(48, "little")
    return bytes(out)


def affine_from_bytes(b: bytes) -> Optional[Tuple[int, int]]:
    if b[96]:
        return None
    x = int.from_bytes(b[0:48], "little")
    y = int.from_bytes(b[48:96], "little")
    assert x < P_MOD and y < P_MOD
    return fq_from_mont(x), fq_from_mont(y)


def jacobian_from_bytes(b: bytes) -> Optional[Tuple[int, int]]:
    """144 B Jacobian (X,Y,Z Montgomery) -> affine canonical ints, None for Z = 0.

    'bit-exact MSM' is defined on this normalised form."""
    X = int.from_bytes(b[0:48], "little")
    Y = int.from_bytes(b[48:96], "little")
    Z = int.from_bytes(b[96:144], "little")
    assert X < P_MOD and Y < P_MOD and Z < P_MOD, "non-reduced Fq limbs crossed the ABI"
    if Z == 0:
        return None
    X, Y, Z = fq_from_mont(X), fq_from_mont(Y), fq_from_mont(Z)
    zi = pow(Z, -1, P_MOD)
    zi2 = zi * zi % P_MOD
    return X * zi2 % P_MOD, Y * zi2 * zi % P_MOD


def g1_compress(pt: Optional[Tuple[int, int]]) -> bytes:
    """Compressed G1Affine (48 B), as `CanonicalSerialize::serialize_compressed` / `ToBytes::write_le` emit it inside a
    Varuna proof [UPSTREAM curves/src/templates/short_weierstrass_jacobian/affine.rs (serialize_with_mode, Compress::Yes),
    utilities/src/serialize/flags.rs (SWFlags)]: the canonical x coordinate little-endian in ceil((377 + 2) / 8) = 48
    bytes with the flags OR-ed into the LAST byte -- bit 7 (PositiveY) set iff y > -y as canonical integers, bit 6
    (Infinity) set for the point at infinity, whose x is serialised as 0.
    RECALLED from upstream, not read from source (the dependency is not on disk): pinned only once the Rust harness
    (rust/snarkvm-algorithms-b200/tests/parity.rs::compressed_encoding_matches_to_bytes_le) has run."""
    if pt is None:
        out = bytearray(48)
        out[47] |= 1 << 6
        return bytes(out)
    x, y = pt
    out = bytearray(x.to_bytes(48, "little"))
    if y > P_MOD - y:
        out[47] |= 1 << 7
    return bytes(out)


def g1_decompress(b: bytes) -> Optional[Tuple[int, int]]:
    """inverse of g1_compress (from_x_coordinate + flag); raises on a non-canonical or off-curve encoding"""
    assert len(b) == 48
    flags = b[47] & 0xC0
    x = int.from_bytes(bytes(b[:47]) + bytes([b[47] & 0x3F]), "little")
    if flags & 0x40:
        if flags & 0x80 or x != 0:
            raise ValueError("invalid infinity encoding")
        return None
    if x >= P_MOD:
        raise ValueError("x not reduced")
    y2 = (x * x * x + G1_B) % P_MOD
    y = fq_sqrt(y2)
    if y is None:
        raise ValueError("x is not on the curve")
    greater = y > P_MOD - y
    if bool(flags & 0x80) != greater:
        y = P_MOD - y
    return x, y


def fq_sqrt(a: int) -> Optional[int]:
    """square root in Fq by Tonelli-Shanks (p - 1 = 2^46 * t); None for a non-residue"""
    a %= P_MOD
    if a == 0:
        return 0
    if pow(a, (P_MOD - 1) // 2, P_MOD) != 1:
        return None
    s, t = 0, P_MOD - 1
    while t % 2 == 0:
        s, t = s + 1, t // 2
    z = 2
    while pow(z, (P_MOD - 1) // 2, P_MOD) != P_MOD - 1:
        z += 1
    m, c, u, r = s, pow(z, t, P_MOD), pow(a, t, P_MOD), pow(a, (t + 1) // 2, P_MOD)
    while u != 1:
        i, v = 0, u
        while v != 1:
            v, i = v * v % P_MOD, i + 1
        b = pow(c, 1 << (m - i - 1), P_MOD)
        m, c = i, b * b % P_MOD
        u, r = u * c % P_MOD, r * b % P_MOD
    return r


# --------------------------------------------------------------------------------------
# G1 arithmetic, affine chord-and-tangent on canonical ints (None = infinity)
# --------------------------------------------------------------------------------------
Point = Optional[Tuple[int, int]]


def is_on_curve(pt: Point) -> bool:
    if pt is None:
        return True
    x, y = pt
    return (y * y - x * x * x - G1_B) % P_MOD == 0


def g1_neg(pt: Point) -> Point:
    if pt is None:
        return None
    return pt[0], (-pt[1]) % P_MOD


def g1_add(a: Point, b: Point) -> Point:
    if a is None:
        return b
    if b is None:
        return a
    x1, y1 = a
    x2, y2 = b
    if x1 == x2:
        if (y1 + y2) % P_MOD == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, P_MOD) % P_MOD
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, P_MOD) % P_MOD
    x3 = (lam * lam - x1 - x2) % P_MOD
    y3 = (lam * (x1 - x3) - y1) % P_MOD
    return x3, y3


# Jacobian (X, Y, Z) over canonical ints: fast enough for the naive oracle MSM.
def _jac_double(P):
    X, Y, Z = P
    if Z == 0:
        return P
    A = X * X % P_MOD
    B = Y * Y % P_MOD
    C = B * B % P_MOD
    D = 2 * ((X + B) * (X + B) - A - C) % P_MOD
    E = 3 * A % P_MOD
    F = E * E % P_MOD
    X3 = (F - 2 * D) % P_MOD
    Y3 = (E * (D - X3) - 8 * C) % P_MOD
    Z3 = 2 * Y * Z % P_MOD
    return X3, Y3, Z3


def _jac_add_affine(P, q: Point):
    if q is None:
        return P
    X1, Y1, Z1 = P
    x2, y2 = q
    if Z1 == 0:
        return x2, y2, 1
    Z1Z1 = Z1 * Z1 % P_MOD
    U2 = x2 * Z1Z1 % P_MOD
    S2 = y2 * Z1 * Z1Z1 % P_MOD
    H = (U2 - X1) % P_MOD
    Rr = (S2 - Y1) % P_MOD
    if H == 0:
        if Rr == 0:
            return _jac_double(P)
        return 1, 1, 0
    HH = H * H % P_MOD
    HHH = H * HH % P_MOD
    V = X1 * HH % P_MOD
    X3 = (Rr * Rr - HHH - 2 * V) % P_MOD
    Y3 = (Rr * (V - X3) - Y1 * HHH) % P_MOD
    Z3 = Z1 * H % P_MOD
    return X3, Y3, Z3


def _jac_to_affine(P) -> Point:
    X, Y, Z = P
    if Z == 0:
        return None
    zi = pow(Z, -1, P_MOD)
    zi2 = zi * zi % P_MOD
    return X * zi2 % P_MOD, Y * zi2 * zi % P_MOD


def g1_mul(pt: Point, k: int) -> Point:
    """double-and-add k*pt (MSB first), the n<15 path of snarkVM's batched::msm restated."""
    if pt is None or k == 0:
        return None
    if k < 0:
        return g1_mul(g1_neg(pt), -k)
    acc = (1, 1, 0)
    for bit in bin(k)[2:]:
        acc = _jac_double(acc)
        if bit == "1":
            acc = _jac_add_affine(acc, pt)
    return _jac_to_affine(acc)


def msm_naive(bases: Sequence[Point], scalars: Sequence[int]) -> Point:
    """sum_i s_i * P_i by independent double-and-add (the 'naive' arm of snarkVM's test_msm)."""
    acc: Point = None
    for p, s in zip(bases, scalars):
        acc = g1_add(acc, g1_mul(p, s % R_MOD))
    return acc


def msm_pippenger(bases: Sequence[Point], scalars: Sequence[int], c: Optional[int] = None) -> Point:
    """Bucket-method MSM with snarkVM's window rule c = ln(n) + 2 (standard::msm restated).

    Unsigned windows, buckets accumulated in Jacobian, running-sum bucket reduction,
    windows folded high -> low with c doublings."""
    import math
    n = min(len(bases), len(scalars))
    if n == 0:
        return None
    if c is None:
        c = 1 if n < 32 else int(math.log(n)) + 2
    nwin = (FR_BITS + c - 1) // c
    total = (1, 1, 0)
    for w in reversed(range(nwin)):
        for _ in range(c):
            total = _jac_double(total)
        buckets = [(1, 1, 0)] * ((1 << c) - 1)
        for p, s in zip(bases[:n], scalars[:n]):
            d = (s >> (w * c)) & ((1 << c) - 1)
            if d and p is not None:
                buckets[d - 1] = _jac_add_affine(buckets[d - 1], p)
        running: Point = None
        acc: Point = None
        # running-sum: sum_b b * B_b
        for b in reversed(buckets):
            running = g1_add(running, _jac_to_affine(b))
            acc = g1_add(acc, running)
        total = _jac_add_affine(total, acc)
    return _jac_to_affine(total)


# --------------------------------------------------------------------------------------
# EvaluationDomain over Fr (canonical ints in, canonical ints out)
# --------------------------------------------------------------------------------------
class EvaluationDomain:
    """Restates snarkVM's EvaluationDomain::<Fr>::new(n) and its (i)FFT entry points
    [UPSTREAM algorithms/src/fft/domain.rs; SURVEY.md section 8a rows a3-a6, appendix A.1]."""

    def __init__(self, num_coeffs: int):
        size = 1
        log = 0
        while size < max(num_coeffs, 1):
            size <<= 1
            log += 1
        if log > FR_TWO_ADICITY:
            raise ValueError("domain too large")
        self.size = size
        self.log_size_of_group = log
        self.group_gen = pow(FR_TWO_ADIC_ROOT, 1 << (FR_TWO_ADICITY - log), R_MOD)
        self.group_gen_inv = pow(self.group_gen, -1, R_MOD)
        self.size_inv = pow(size, -1, R_MOD)
        self.generator = FR_GENERATOR
        self.generator_inv = pow(FR_GENERATOR, -1, R_MOD)

    # O(n log n) radix-2, natural order in and out
    @staticmethod
    def _ntt(a: List[int], omega: int) -> List[int]:
        n = len(a)
        if n == 1:
            return list(a)
        logn = n.bit_length() - 1
        # bit-reverse then DIT butterflies
        a = list(a)
        for i in range(n):
            j = int(bin(i)[2:].zfill(logn)[::-1], 2) if logn else 0
            if i < j:
                a[i], a[j] = a[j], a[i]
        m = 1
        while m < n:
            wm = pow(omega, n // (2 * m), R_MOD)
            for k in range(0, n, 2 * m):
                w = 1
                for j in range(m):
                    t = a[k + j + m] * w % R_MOD
                    u = a[k + j]
                    a[k + j] = (u + t) % R_MOD
                    a[k + j + m] = (u - t) % R_MOD
                    w = w * wm % R_MOD
            m *= 2
        return a

    def _pad(self, coeffs: Sequence[int]) -> List[int]:
        assert len(coeffs) <= self.size, "input longer than the domain is a caller bug"
        return [c % R_MOD for c in coeffs] + [0] * (self.size - len(coeffs))

    def fft(self, coeffs: Sequence[int]) -> List[int]:
        return self._ntt(self._pad(coeffs), self.group_gen)

    def ifft(self, evals: Sequence[int]) -> List[int]:
        out = self._ntt(self._pad(evals), self.group_gen_inv)
        return [v * self.size_inv % R_MOD for v in out]

    def coset_fft(self, coeffs: Sequence[int]) -> List[int]:
        a = self._pad(coeffs)
        g = 1
        for i in range(len(a)):            # distribute_powers(coeffs, g = 22)
            a[i] = a[i] * g % R_MOD
            g = g * self.generator % R_MOD
        return self._ntt(a, self.group_gen)

    def coset_ifft(self, evals: Sequence[int]) -> List[int]:
        a = self.ifft(evals)
        g = 1
        for i in range(len(a)):            # distribute_powers(coeffs, g^-1)
            a[i] = a[i] * g % R_MOD
            g = g * self.generator_inv % R_MOD
        return a

    def dft_naive(self, coeffs: Sequence[int]) -> List[int]:
        """O(n^2) definition out[k] = sum_j in[j] * w^(jk): the oracle's own oracle."""
        a = self._pad(coeffs)
        n = self.size
        pw = [pow(self.group_gen, e, R_MOD) for e in range(n)]
        return [sum(a[j] * pw[(j * k) % n] for j in range(n)) % R_MOD for k in range(n)]


# --------------------------------------------------------------------------------------
# Deterministic synthetic inputs: splitmix64 stream
# --------------------------------------------------------------------------------------
class SplitMix64:
    def __init__(self, seed: int):
        self.s = seed & 0xFFFFFFFFFFFFFFFF

    def next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)

    def below(self, bound: int, bits: int) -> int:
        """uniform in [0, bound) by rejection on `bits`-bit draws."""
        mask = (1 << bits) - 1
        while True:
            v = 0
            for i in range((bits + 63) // 64):
                v |= self.next() << (64 * i)
            v &= mask
            if v < bound:
                return v


def random_fr(rng: SplitMix64, n: int) -> List[int]:
    return [rng.below(R_MOD, FR_BITS) for _ in range(n)]


def random_points(rng: SplitMix64, n: int) -> List[Point]:
    """n pseudo-random G1 points: k_i * G built incrementally (P_{i+1} = P_i + d*G style walk
    would be correlated, so use independent small multiples of a few seeded generators)."""
    pts: List[Point] = []
    base = g1_mul(G1_GEN, rng.below(R_MOD, FR_BITS) | 1)
    step = g1_mul(G1_GEN, rng.below(R_MOD, FR_BITS) | 1)
    cur = base
    for _ in range(n):
        pts.append(cur)
        cur = g1_add(cur, step)
    return pts


def digest(data: bytes) -> str:
    return hashlib.sha256(data).hexdigest()
