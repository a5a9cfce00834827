"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in snarkos_b200/dist.py: point-range sharding +
partial-sum gather for the MSM, and the exchange / index arithmetic of the four-step NTT.  The local kernels are
replaced by the oracle (injectable `ops`), so what is tested here is exactly the code that runs between the kernels
on a multi-GPU box."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import bls12_377 as O
from oracle import c_oracle as C
from tests import helpers as H


class OracleOps:
    """CPU stand-ins for the CUDA kernels (test only)."""

    def msm(self, bases, scalars, stride=104):
        out = C.msm(bases.numpy(), scalars.numpy().view(np.uint64), stride=stride)
        return torch.from_numpy(out.copy())

    def sum_projective(self, points):
        acc = None
        for row in points.numpy():
            acc = O.g1_add(acc, H.jac_bytes_to_affine(row))
        buf = bytearray(144)
        if acc is None:
            buf[0:48] = O.fq_to_mont(1).to_bytes(48, "little")
            buf[48:96] = O.fq_to_mont(1).to_bytes(48, "little")
        else:
            buf[0:48] = O.fq_to_mont(acc[0]).to_bytes(48, "little")
            buf[48:96] = O.fq_to_mont(acc[1]).to_bytes(48, "little")
            buf[96:144] = O.fq_to_mont(1).to_bytes(48, "little")
        return torch.from_numpy(np.frombuffer(bytes(buf), dtype=np.uint8).copy())

    def ntt_rows(self, mat, log_len, direction):
        a = mat.numpy().view(np.uint64)
        rows = a.shape[0]
        a[:] = C.ntt(a.reshape(-1, 4), log_len, batch=rows, direction=direction).reshape(a.shape)

    def mul_powers(self, mat, log_n, direction, kind, rows, cols, row_base, col_base):
        a = mat.numpy().view(np.uint64).reshape(-1, 4)
        d = O.EvaluationDomain(1 << log_n)
        if kind == 0:
            w = d.group_gen_inv if direction else d.group_gen
            ex = [((row_base + r) * (col_base + c)) % (1 << log_n) for r in range(rows) for c in range(cols)]
        else:
            w = d.generator_inv if direction else d.generator
            ex = [row_base + i for i in range(rows * cols)]
        f = H.fr_mont_array([pow(w, e, O.R_MOD) for e in ex])
        a[:] = C.fr_mul(a, f)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from snarkos_b200 import dist as D
        ops = OracleOps()
        # ---- MSM: shard by point range, gather + add --------------------------------------------------------
        n = 75
        rng = O.SplitMix64(11)
        pts = O.random_points(rng, n)
        sc = O.random_fr(rng, n)
        lo, hi = D.shard_range(n, rank, world)
        bases = torch.from_numpy(H.bases_array(pts[lo:hi]))
        scal = torch.from_numpy(H.scalars_array(sc[lo:hi]).view(np.int64))
        got = D.msm_sharded(bases, scal, ops=ops)
        assert H.jac_bytes_to_affine(got.numpy()) == O.msm_naive(pts, sc), "sharded MSM mismatch"
        # ---- four-step NTT: every (direction, coset) kind, natural out; and the k1-slab layout ---------------
        for log_n, log_n1 in ((6, 3), (7, 3), (8, 5)):
            N = 1 << log_n
            x = O.random_fr(O.SplitMix64(100 + log_n), N)
            per = N // world
            d = O.EvaluationDomain(N)
            for direction, coset, f in ((0, 0, d.fft), (1, 0, d.ifft), (0, 1, d.coset_fft), (1, 1, d.coset_ifft)):
                blk = torch.from_numpy(H.fr_mont_array(x[rank * per:(rank + 1) * per]).view(np.int64))
                keep = blk.clone()
                out = D.ntt_distributed(blk, log_n, direction, coset, natural_out=True, ops=ops, log_n1=log_n1)
                want = f(x)[rank * per:(rank + 1) * per]
                assert H.fr_from_mont_array(out.numpy().view(np.uint64)) == want, (log_n, direction, coset)
                assert torch.equal(blk, keep), "the input block is an input: coset scaling must not modify it"
            blk = torch.from_numpy(H.fr_mont_array(x[rank * per:(rank + 1) * per]).view(np.int64))
            slab = D.ntt_distributed(blk, log_n, 0, 0, natural_out=False, ops=ops, log_n1=log_n1)
            n1 = 1 << log_n1
            full = d.fft(x)
            got = H.fr_from_mont_array(slab.numpy().view(np.uint64).reshape(-1, 4))
            rows = n1 // world
            n2 = N // n1
            want = [full[(rank * rows + r) + n1 * k2] for r in range(rows) for k2 in range(n2)]
            assert got == want, "k1-slab layout mismatch"
            try:                                               # refused before any exchange is enqueued (both ranks raise)
                D.ntt_distributed(blk, log_n, 1, 1, natural_out=False, ops=ops, log_n1=log_n1)
                raise AssertionError("inverse coset with slab output must be refused")
            except ValueError:
                pass
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: " + repr(e) + "\n" + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_shard_range_tiles():
    from snarkos_b200 import dist as D
    for n in (0, 1, 7, 64, 1000):
        for world in (1, 2, 3, 8):
            r = [D.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1


def test_two_rank_gloo_msm_and_four_step_ntt():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, status in results:
        assert status == "ok", f"rank {rank}: {status}"
