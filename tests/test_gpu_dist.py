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
