import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _gpu_box() -> bool:
    """True when an NVIDIA GPU is visible to the driver tools (cheap; does not initialise CUDA in this process)."""
    import shutil
    import subprocess
    if not shutil.which("nvidia-smi"):
        return False
    try:
        return subprocess.run(["nvidia-smi", "-L"], capture_output=True, timeout=20).returncode == 0
    except Exception:  # noqa: BLE001
        return False


def pytest_collection_modifyitems(config, items):
    if _gpu_box():
        # a failed cuInit is sticky inside a process: wait (in subprocesses) until the driver answers before this
        # process touches CUDA, so that a transient refusal cannot silently skip the GPU suite
        from cstp_b200.parallel import wait_for_cuda
        wait_for_cuda()
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture
def emulated_engine(monkeypatch):
    """cstp_b200.engine with its kernel layer replaced by the CPU emulator (host-orchestration tests only)."""
    from cstp_b200 import engine
    from tests import emulate_ops
    monkeypatch.setattr(engine, "ops", emulate_ops)
    return engine


GOLDEN = os.path.join(ROOT, "tests", "golden")
