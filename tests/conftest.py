import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real sm_100 (B200) GPU; run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_models():
    """Session cache of calibrated oracle models (fp32, CPU): {scale: (unfused_model, state_dict)}."""
    from oracle import yolo11_ref as R
    cache = {}

    def get(scale: str):
        if scale not in cache:
            m = R.build(scale, init="calibrated", seed=0)
            sd = {k: v.clone() for k, v in m.state_dict().items()}
            cache[scale] = (m, sd)
        return cache[scale]

    return get
