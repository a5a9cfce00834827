"""Shared helpers for the parity tests: byte-level conversions between Python ints (oracle side)
and the numpy buffers that cross the C ABI."""
import json
import os

import numpy as np

from oracle import bls12_377 as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_kat():
    raw = json.load(open(os.path.join(GOLDEN, "kat.json")))
    out = {}
    for k, v in raw.items():
        if isinstance(v, list):
            out[k] = [int(x, 0) if isinstance(x, str) and x.startswith("0x") else int(x) for x in v]
        elif isinstance(v, str) and v.startswith("0x"):
            out[k] = int(v, 16)
        else:
            out[k] = int(v)
    return out


def ints_to_limbs(vals, nlimbs):
    """list of ints -> uint64 [n, nlimbs] little-endian"""
    buf = b"".join(int(v).to_bytes(8 * nlimbs, "little") for v in vals)
    return np.frombuffer(buf, dtype=np.uint64).reshape(-1, nlimbs).copy()


def limbs_to_ints(arr):
    arr = np.ascontiguousarray(arr, dtype=np.uint64)
    nl = arr.shape[-1]
    raw = arr.reshape(-1, nl).tobytes()
    return [int.from_bytes(raw[i * 8 * nl:(i + 1) * 8 * nl], "little") for i in range(arr.reshape(-1, nl).shape[0])]


def fr_mont_array(vals):
    """canonical Fr ints -> Montgomery limbs uint64 [n,4] (what EvaluationDomain holds in memory)"""
    return ints_to_limbs([O.fr_to_mont(v) for v in vals], 4)


def fr_from_mont_array(arr):
    return [O.fr_from_mont(v) for v in limbs_to_ints(arr)]


def scalars_array(vals):
    return ints_to_limbs(vals, 4)


def bases_array(points, stride=O.AFFINE_STRIDE):
    return np.frombuffer(b"".join(O.affine_bytes(p, stride) for p in points), dtype=np.uint8).copy()


def jac_bytes_to_affine(jac):
    return O.jacobian_from_bytes(bytes(np.ascontiguousarray(jac, dtype=np.uint8).reshape(-1)[:144]))


# ---------------------------------------------------------------------------------------------
# synthetic-input mirrors (numpy) of what the library generates on the device
# ---------------------------------------------------------------------------------------------
def splitmix64_at(seed: int, idx: np.ndarray) -> np.ndarray:
    """k_i of b200_g1_synthetic_bases_device: splitmix64 output number i+1 of the stream seeded with `seed`."""
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + (idx.astype(np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def dot_mod_r(scalars_u64x4: np.ndarray, k_u64: np.ndarray) -> int:
    """sum_i s_i * k_i mod r, exactly, with vectorised 32-bit partial products (size-independent MSM check:
    if P_i = k_i * G then sum_i s_i P_i = (sum_i s_i k_i mod r) * G)."""
    s32 = np.ascontiguousarray(scalars_u64x4, dtype=np.uint64).view(np.uint32).reshape(-1, 8).astype(np.uint64)
    k32 = np.ascontiguousarray(k_u64, dtype=np.uint64).view(np.uint32).reshape(-1, 2).astype(np.uint64)
    mask = np.uint64(0xFFFFFFFF)
    total = 0
    for a in range(8):
        for b in range(2):
            p = s32[:, a] * k32[:, b]
            part = (int((p >> np.uint64(32)).sum(dtype=np.uint64)) << 32) + int((p & mask).sum(dtype=np.uint64))
            total += part << (32 * (a + b))
    return total % O.R_MOD


def random_scalars_np(rng: np.random.Generator, n: int) -> np.ndarray:
    """n canonical scalars < 2^252 < r as uint64 [n, 4] (uniform on [0, 2^252): fine for synthetic load)."""
    s = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, 4), dtype=np.uint64)
    s[:, 3] &= np.uint64((1 << 60) - 1)
    return s


def random_fr_mont_np(rng: np.random.Generator, shape) -> np.ndarray:
    """Montgomery Fr limbs: any value < r is a valid representative; draw < 2^252."""
    s = rng.integers(0, 1 << 63, size=tuple(shape) + (4,), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=tuple(shape) + (4,), dtype=np.uint64)
    s[..., 3] &= np.uint64((1 << 60) - 1)
    return s
