import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "sam-decoding_b200"), os.path.join(REPO, "oracle"), os.path.join(REPO, "tests"), REPO):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # a fresh checkout has no built library yet (it is git-ignored): build it the way __graft_entry__.build() does,
    # so that the suite does not depend on the order the driver runs things in.  nvcc cross-compiles without a GPU.
    lib = os.path.join(REPO, "sam-decoding_b200", "samd_b200", "libsamd_b200.so")
    if not os.path.exists(lib):
        import shutil
        import subprocess
        if shutil.which("nvcc") and shutil.which("make"):
            subprocess.run(["make", "-C", os.path.join(REPO, "sam-decoding_b200", "csrc")], check=False,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(REPO, "tests", "golden")
