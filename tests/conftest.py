import os
import sys

import numpy as np
import pytest

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="session")
def built_library():
    """The library is a build artefact (git-ignored): build it before anything dlopens it.  Tests that need it are
    skipped, not failed, on a machine that has neither the built library nor nvcc."""
    import shutil
    from macarons_b200 import build
    if not build.up_to_date() and not (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")):
        if os.path.exists(build.lib_path()):
            return build.lib_path()          # stale but present and no compiler: use what travelled with the snapshot
        pytest.skip("libmacarons_b200.so is not built and nvcc is not available")
    return build.build()


# Only tests that load libmacarons_b200.so depend on the build: every GPU test, and the CPU tests of these modules
# (ABI / symbol checks, host-side logic behind the library, the bench contract).  Pure oracle / golden tests do not.
_LIBRARY_MODULES = ("test_abi_cpu", "test_dropin_cpu", "test_macarons_host_cpu", "test_bench_contract_cpu",
                    "test_parallel_gloo", "test_autograd_cpu", "test_scene_cpu")


def pytest_collection_modifyitems(config, items):
    for item in items:
        mod = getattr(item.module, "__name__", "").rsplit(".", 1)[-1]
        if item.get_closest_marker("gpu") is not None or mod in _LIBRARY_MODULES:
            item.fixturenames.append("built_library")
