import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 via gpurun)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as g
    return g.load_package()


@pytest.fixture(autouse=True)
def _cpu_tests_start_from_a_fixed_rng_state(request):
    """Every CPU test starts from the same torch / numpy / random state, so a test's draws do not depend on which tests
    ran before it (an unseeded draw once put a pre-activation within rounding of zero in one suite order only).  The GPU
    tests keep the generator state they were measured with."""
    if "gpu" not in request.keywords:
        import random

        import numpy as np
        import torch
        torch.manual_seed(20261019)
        np.random.seed(20261019)
        random.seed(20261019)
    yield


@pytest.fixture(autouse=True, scope="module")
def _gpu_test_files_start_from_a_fresh_process_rng_state():
    """On a GPU box every test FILE starts from the generator state of a fresh process (torch's default seed, CPU and
    CUDA generators): the files were measured one process each (scripts/gpu_tests.sh), and `pytest tests -m gpu` in ONE
    process should feed the tests that draw without a seed the same inputs."""
    import torch
    if torch.cuda.is_available():
        torch.manual_seed(67280421310721)       # c10::detail default_rng_seed_val
    yield
