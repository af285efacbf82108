import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """Without a CUDA device the gpu-marked tests are skipped (plain `pytest` stays green on a CPU box); on a GPU box nothing is
    skipped, and the product path itself still fails loudly if the CUDA library is missing."""
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (B200)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _fresh_bc_registry():
    """BC ids are process-global and only grow (reference: boundary_condition_registry.py); restart them per test."""
    from xlb_b200.operator.boundary_condition.boundary_condition_registry import boundary_condition_registry as reg

    reg.next_id = 1
    reg.id_to_bc.clear()
    reg.bc_to_id.clear()
    yield
