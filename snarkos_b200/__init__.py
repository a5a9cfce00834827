"""snarkos_b200 -- B200-native (sm_100a) BLS12-377 G1 MSM and Fr NTT: the data-parallel hot path snarkOS reaches
through snarkVM's Varuna prover / verifier, behind snarkVM's own interface names.

    VariableBase.msm(bases, scalars)                       <- snarkvm_algorithms::msm::VariableBase::msm
    EvaluationDomain(n).{fft,ifft,coset_fft,coset_ifft}_in_place  <- snarkvm_algorithms::fft::EvaluationDomain

The compute lives in snarkos_b200/libsnarkos_b200.so (CUDA, C ABI in include/snarkos_b200.h).  No CPU fallback.
"""
from ._lib import B200Error, counter, init, kernel_launch_count, lib, profile, release_scratch, set_option  # noqa: F401
from .fft import EvaluationDomain  # noqa: F401
from .msm import ResidentBases, VariableBase, g1_batch_normalize, g1_compress, msm_batch, sum_projective, synthetic_bases  # noqa: F401
from .kzg10 import KZG10, Powers  # noqa: F401
from . import dist, poly, varuna  # noqa: F401,E402
