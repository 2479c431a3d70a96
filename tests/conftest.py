"""pytest configuration: markers + repo root on sys.path."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


import pytest  # noqa: E402


@pytest.fixture()
def emu(monkeypatch):
    """Route the package's ctypes calls to the thread-emulated build of the kernels (CPU-only test aid)."""
    from tests.emu_support import load_emu
    from py_neuromodulation_b200 import _lib

    lib = load_emu()
    monkeypatch.setattr(_lib, "_LIB", lib)
    yield lib
