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
