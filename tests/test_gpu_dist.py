"""GPU checks of snarkos_b200/dist.py with the real CUDA ops.  With one visible GPU the exchanges degenerate to local
re-tilings (world = 1) but the whole four-step schedule, the twiddle / coset block-scaling kernel and the partial-sum
combine run on the device; tools/multi_gpu_check.py runs the same checks under torchrun on several GPUs."""
import numpy as np
import pytest

from oracle import c_oracle as C
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("log_n,log_n1", [(10, 5), (15, 7), (20, 10)])
@pytest.mark.parametrize("direction,coset", [(0, 0), (1, 0), (0, 1), (1, 1)])
def test_four_step_schedule_single_rank(log_n, log_n1, direction, coset):
    import torch
    from snarkos_b200 import dist as D
    n = 1 << log_n
    x = H.random_fr_mont_np(np.random.default_rng(log_n + direction), (n,))
    blk = torch.from_numpy(x.view(np.int64)).cuda()
    out = D.ntt_distributed(blk, log_n, direction, coset, natural_out=True, log_n1=log_n1)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().view(np.uint64), C.ntt(x, log_n, direction=direction, coset=coset))


def test_mul_powers_kernel_vs_python():
    import ctypes
    import torch
    import snarkos_b200 as S
    from oracle import bls12_377 as O
    log_n, rows, cols, rb, cb = 12, 5, 7, 300, 11
    vals = O.random_fr(O.SplitMix64(2), rows * cols)
    for direction in (0, 1):
        d = O.EvaluationDomain(1 << log_n)
        w = d.group_gen_inv if direction else d.group_gen
        g = d.generator_inv if direction else d.generator
        for kind, want in ((0, [v * pow(w, ((rb + r) * (cb + c)) % (1 << log_n), O.R_MOD) % O.R_MOD for r in range(rows) for c in range(cols) for v in [vals[r * cols + c]]]),
                           (1, [v * pow(g, rb + i, O.R_MOD) % O.R_MOD for i, v in enumerate(vals)])):
            t = torch.from_numpy(H.fr_mont_array(vals).view(np.int64)).cuda()
            S._lib.check(S.lib().b200_fr_mul_powers_device(ctypes.c_void_p(t.data_ptr()), log_n, direction, kind, rows, cols, rb, cb,
                                                           ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
            torch.cuda.synchronize()
            assert H.fr_from_mont_array(t.cpu().numpy().view(np.uint64)) == want


def test_msm_sharded_single_rank():
    import torch
    import snarkos_b200 as S
    from snarkos_b200 import dist as D
    n = 1 << 11
    bases = S.synthetic_bases(n, seed=5)
    sc = H.random_scalars_np(np.random.default_rng(1), n)
    out = D.msm_sharded(bases, torch.from_numpy(sc.view(np.int64)).cuda())
    torch.cuda.synchronize()
    assert H.jac_bytes_to_affine(out.cpu().numpy()) == H.jac_bytes_to_affine(C.msm(bases.cpu().numpy(), sc))


@pytest.mark.parametrize("log_n,world,parts", [(10, 2, 1), (12, 4, 1), (13, 2, 1), (16, 8, 1), (12, 4, 2), (16, 8, 4), (16, 2, 3)])
@pytest.mark.parametrize("direction", [0, 1])
def test_fused_exchange_virtual_world(log_n, world, parts, direction):
    """b200_fr_exchange_transpose_device (transpose + all-to-all + twiddle in one kernel) with `world` virtual ranks on
    ONE GPU: every rank's slabs are separate buffers of this device, the destination pointer table is the same one a
    real run fills with CUDA-IPC peer pointers.  The full fused four-step schedule must equal the oracle's NTT."""
    import ctypes
    import torch
    from snarkos_b200 import dist as D
    ops = D.CudaOps()
    n = 1 << log_n
    n1_log = log_n // 2
    n2_log = log_n - n1_log
    n1, n2, per = 1 << n1_log, 1 << n2_log, n // world
    x = H.random_fr_mont_np(np.random.default_rng(7 * log_n + direction), (n,))
    blocks = [torch.from_numpy(x[r * per:(r + 1) * per].view(np.int64)).cuda() for r in range(world)]
    B = [torch.empty((per, 4), dtype=torch.int64, device="cuda") for _ in range(world)]
    Cb = [torch.empty((per, 4), dtype=torch.int64, device="cuda") for _ in range(world)]
    dst_b = (ctypes.c_void_p * world)(*[t.data_ptr() for t in B])
    dst_c = (ctypes.c_void_p * world)(*[t.data_ptr() for t in Cb])
    rb, rc = n2 // world, n1 // world

    def cuts(total):                                       # `parts` ragged pieces of [0, total)
        edges = [total * i // parts for i in range(parts + 1)]
        return [(a, b - a) for a, b in zip(edges, edges[1:]) if b > a]

    for r in range(world):
        if parts == 1:
            D.exchange_transpose(blocks[r], dst_b, world, r, rc, n2, log_n, direction, False)
        else:                                              # column sub-ranges of every destination (chunks of B's rows)
            for lo, cnt in cuts(rb):
                D.exchange_transpose_part(blocks[r], dst_b, world, r, rc, 0, rc, n2, lo, cnt, log_n, direction, False)
    for r in range(world):
        ops.ntt_rows(B[r].view(rb, n1, 4), n1_log, direction)
    for r in range(world):
        if parts == 1:
            D.exchange_transpose(B[r], dst_c, world, r, rb, n1, log_n, direction, True, r * rb)
        else:                                              # row ranges of the source slab
            for lo, cnt in cuts(rb):
                D.exchange_transpose_part(B[r], dst_c, world, r, rb, lo, cnt, n1, 0, rc, log_n, direction, True, r * rb)
    for r in range(world):
        ops.ntt_rows(Cb[r].view(rc, n2, 4), n2_log, direction)
    for r in range(world):
        for lo, cnt in cuts(rc):
            D.exchange_transpose_part(Cb[r], dst_b, world, r, rc, lo, cnt, n2, 0, rb, log_n, direction, False)
    torch.cuda.synchronize()
    got = torch.cat(B).cpu().numpy().view(np.uint64)
    assert np.array_equal(got, C.ntt(x, log_n, direction=direction))


def test_peer_exchange_single_rank():
    """PeerExchange / ntt_distributed_fused with world = 1 (no process group): IPC export of the slabs, zero-copy
    tensor views, all four transform kinds against the oracle"""
    import torch
    from snarkos_b200 import dist as D
    log_n = 12
    n = 1 << log_n
    x = H.random_fr_mont_np(np.random.default_rng(3), (n,))
    fab = D.PeerExchange(n)
    for direction, coset in ((0, 0), (1, 0), (0, 1), (1, 1)):
        blk = torch.from_numpy(x.view(np.int64)).cuda()
        out = D.ntt_distributed_fused(blk, log_n, fab, direction, coset)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy().view(np.uint64), C.ntt(x, log_n, direction=direction, coset=coset))
        blk = torch.from_numpy(x.view(np.int64)).cuda()
        out = D.ntt_distributed_overlapped(blk, log_n, fab, direction, coset, chunks=4)      # two streams, chunked slabs
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy().view(np.uint64), C.ntt(x, log_n, direction=direction, coset=coset))
    fab.close()


@pytest.mark.parametrize("log_n,world", [(10, 2), (12, 4), (13, 2), (16, 8), (20, 4), (22, 8)])
@pytest.mark.parametrize("direction", [0, 1])
def test_transforms_fused_with_exchange_virtual_world(log_n, world, direction):
    """b200_ntt_rows_exchange_device: the row transforms' last pass IS the exchange (butterflies, twiddle and the stores
    into the destination ranks' slabs in one launch set).  `world` virtual ranks on one GPU, full four-step schedule
    with both fused steps == the oracle's transform (and == the schedule with separate exchange kernels)."""
    import ctypes
    import torch
    from snarkos_b200 import dist as D
    n = 1 << log_n
    n1_log = log_n // 2
    n2_log = log_n - n1_log
    n1, n2, per = 1 << n1_log, 1 << n2_log, n // world
    x = H.random_fr_mont_np(np.random.default_rng(11 * log_n + direction), (n,))
    blocks = [torch.from_numpy(x[r * per:(r + 1) * per].view(np.int64)).cuda() for r in range(world)]
    B = [torch.zeros((per, 4), dtype=torch.int64, device="cuda") for _ in range(world)]
    Cb = [torch.zeros((per, 4), dtype=torch.int64, device="cuda") for _ in range(world)]
    dst_b = (ctypes.c_void_p * world)(*[t.data_ptr() for t in B])
    dst_c = (ctypes.c_void_p * world)(*[t.data_ptr() for t in Cb])
    rb, rc = n2 // world, n1 // world
    for r in range(world):
        D.exchange_transpose(blocks[r], dst_b, world, r, rc, n2, log_n, direction, False)
    torch.cuda.synchronize()
    keep_b = [t.clone() for t in B]
    for r in range(world):                                  # transforms of B's rows + twiddle + exchange into every C
        D.ntt_rows_exchange(B[r], dst_c, world, r, rb, n1_log, log_n, direction, True, r * rb)
    torch.cuda.synchronize()
    for r in range(world):
        assert torch.equal(B[r], keep_b[r])                 # the source slab is left unchanged
    for r in range(world):                                  # transforms of C's rows + exchange into every B (natural order)
        D.ntt_rows_exchange(Cb[r], dst_b, world, r, rc, n2_log, log_n, direction, False)
    torch.cuda.synchronize()
    got = torch.cat(B).cpu().numpy().view(np.uint64)
    if log_n <= 16:
        assert np.array_equal(got, C.ntt(x, log_n, direction=direction))
    else:
        import snarkos_b200 as S
        t = torch.from_numpy(x.view(np.int64)).cuda()
        d = S.EvaluationDomain(n)
        want = d.ifft_in_place(t) if direction else d.fft_in_place(t)
        torch.cuda.synchronize()
        assert np.array_equal(got, want.cpu().numpy().view(np.uint64))
