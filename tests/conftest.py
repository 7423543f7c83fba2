import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "emu_only: runs only against the CPU emulation build (tools/cuemu)")
    config.addinivalue_line("markers", "needs_torch_cuda: GPU test that also needs torch.cuda (device tensors, NCCL)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        skip_emu = pytest.mark.skip(reason="written for the CPU emulation of the kernels (host pointers as device pointers)")
        for it in items:
            if "emu_only" in it.keywords:
                it.add_marker(skip_emu)
        return
    if os.environ.get("OAK_B200_TEST_EMU") == "1":
        # tests/test_emulated_kernels.py re-runs GPU parity tests against the CPU emulation build of the kernel
        # sources (tools/cuemu) in a subprocess with this variable set: only tests needing torch.cuda stay skipped
        skip_cuda = pytest.mark.skip(reason="needs torch.cuda (not available in the CPU emulation of the kernels)")
        for it in items:
            if "needs_torch_cuda" in it.keywords:
                it.add_marker(skip_cuda)
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
