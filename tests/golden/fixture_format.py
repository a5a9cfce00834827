"""Reader / writer of the fixture files that pin parity to snarkVM itself (format: tests/golden/FIXTURES.md; the Rust
writer is rust/snarkvm-algorithms-b200/src/bin/make_fixtures.rs).  Field elements and points are raw memory images of
the snarkVM types, i.e. exactly the bytes that cross the C ABI."""
import glob
import os
import struct

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MSM_MAGIC, NTT_MAGIC = b"B2MSM001", b"B2NTT001"


def write_msm(path, bases: bytes, scalars: bytes, n: int, result_jac: bytes, result_affine: bytes, compressed: bytes,
              stride: int = 104, offsets=(0, 48, 96)):
    assert len(bases) == n * stride and len(scalars) == n * 32
    assert len(result_jac) == 144 and len(result_affine) == stride and len(compressed) == 48
    with open(path, "wb") as f:
        f.write(MSM_MAGIC + struct.pack("<6Q", n, stride, *offsets, 144))
        f.write(bases + scalars + result_jac + result_affine + compressed)


def read_msm(path):
    raw = open(path, "rb").read()
    assert raw[:8] == MSM_MAGIC, f"{path}: not an MSM fixture"
    n, stride, ox, oy, oinf, pb = struct.unpack_from("<6Q", raw, 8)
    at = 56
    out = {"n": n, "stride": stride, "offsets": (ox, oy, oinf), "proj_bytes": pb, "path": path}
    for key, size in (("bases", n * stride), ("scalars", n * 32), ("result_jac", pb), ("result_affine", stride), ("compressed", 48)):
        out[key] = raw[at:at + size]
        at += size
    assert at == len(raw), f"{path}: trailing or missing bytes"
    return out


def write_ntt(path, log_n: int, inp: bytes, outs):
    n_in = len(inp) // 32
    assert len(outs) == 4 and all(len(o) == 32 << log_n for o in outs)
    with open(path, "wb") as f:
        f.write(NTT_MAGIC + struct.pack("<2Q", log_n, n_in) + inp + b"".join(outs))


def read_ntt(path):
    raw = open(path, "rb").read()
    assert raw[:8] == NTT_MAGIC, f"{path}: not an NTT fixture"
    log_n, n_in = struct.unpack_from("<2Q", raw, 8)
    at = 24
    out = {"log_n": log_n, "n_in": n_in, "path": path, "input": np.frombuffer(raw, dtype=np.uint64, count=4 * n_in, offset=at).reshape(n_in, 4)}
    at += 32 * n_in
    for key in ("fft", "ifft", "coset_fft", "coset_ifft"):
        out[key] = np.frombuffer(raw, dtype=np.uint64, count=4 << log_n, offset=at).reshape(1 << log_n, 4)
        at += 32 << log_n
    assert at == len(raw), f"{path}: trailing or missing bytes"
    return out


def find(prefix: str):
    """fixtures whose file name starts with `prefix`_ ('snarkvm': written by snarkVM; 'oraclefmt': format samples written
    by make_format_samples.py from the Python big-int oracle)"""
    return sorted(glob.glob(os.path.join(HERE, f"{prefix}_msm_*.bin"))), sorted(glob.glob(os.path.join(HERE, f"{prefix}_ntt_*.bin")))
