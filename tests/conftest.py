import os
import shutil
import subprocess
import sys
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def _cuda_available():
    """torch.cuda.is_available(), retried for a few seconds when the machine HAS an NVIDIA GPU: on the GPU box the check was once
    seen returning False for a process started right after another one released the device (the whole `-m gpu` suite would then
    be skipped silently)."""
    import torch
    if torch.cuda.is_available():
        return True
    smi = shutil.which("nvidia-smi")
    if smi is None:
        return False
    try:
        has_gpu = "GPU " in subprocess.run([smi, "-L"], capture_output=True, text=True, timeout=20).stdout
    except Exception:
        return False
    for _ in range(10 if has_gpu else 0):
        time.sleep(3.0)
        if torch.cuda.is_available():
            return True
    return False


def pytest_collection_modifyitems(config, items):
    if not any("gpu" in it.keywords for it in items):
        return
    if _cuda_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
