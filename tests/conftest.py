import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


# run-time tuning knobs of the library (b200_set_option); the B200_* environment is only read on first use, so tests
# change knobs through this fixture, which restores the defaults afterwards
_OPTION_DEFAULTS = {
    "msm_window_bits": 0, "msm_glv": 1, "msm_affine_rounds": -1, "msm_slices": 1, "msm_chunk": 0, "msm_seg_len": 0, "msm_reduce_quad_max": 8192, "msm_host_pipeline": 1,
    "msm_host_first_log": 20, "msm_host_chunk_log": 23, "msm_stream_two": 1, "msm_auto_table": 1, "msm_list_budget_bytes": 0, "msm_queue_threshold": 0, "msm_queue_linger_us": 100,
    "ntt_plan": "", "ntt_tile_log": 0, "ntt_radix4": 1, "ntt_boundary_tables": 1, "ntt_host_pipeline": 1, "ntt_variant": 0,
    "staged_copies": 1,
}


@pytest.fixture
def b200_opt():
    import snarkos_b200 as S
    changed = []

    def set_(key, value):
        assert key in _OPTION_DEFAULTS, key
        S.set_option(key, value)
        changed.append(key)

    yield set_
    for key in changed:
        S.set_option(key, _OPTION_DEFAULTS[key])
