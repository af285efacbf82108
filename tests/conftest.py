import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(autouse=True)
def _fresh_bc_registry():
    """BC ids are process-global and only grow (reference: boundary_condition_registry.py); restart them per test."""
    from xlb_b200.operator.boundary_condition.boundary_condition_registry import boundary_condition_registry as reg

    reg.next_id = 1
    reg.id_to_bc.clear()
    reg.bc_to_id.clear()
    yield
