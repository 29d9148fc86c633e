import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cuda_dev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The C-ABI library must exist for both the CPU and the GPU suites."""
    import shutil
    lib = os.path.join(ROOT, "pfpn_b200", "libpfpn_b200.so")
    if shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"):
        if not os.path.exists(lib):
            pytest.skip("no nvcc and no prebuilt libpfpn_b200.so: nothing to test against", allow_module_level=False)
        return  # prebuilt library travels with the tree (gpurun box); nothing to rebuild
    import __graft_entry__
    __graft_entry__.build()
