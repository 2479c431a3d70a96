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


def _gpu_available() -> bool:
    try:
        from py_neuromodulation_b200 import _lib

        return _lib.device_count() > 0
    except Exception:
        return False


@pytest.fixture(params=[pytest.param("emu"), pytest.param("gpu", marks=pytest.mark.gpu)])
def backend(request, monkeypatch):
    """Parity tests run twice: on the thread-emulated kernels (CPU suite) and on the real CUDA library (-m gpu)."""
    from py_neuromodulation_b200 import _lib

    if request.param == "emu":
        from tests.emu_support import load_emu

        monkeypatch.setattr(_lib, "_LIB", load_emu())
    else:
        monkeypatch.setattr(_lib, "_LIB", None)
        lib = _lib.load()  # raises if the CUDA library is missing: GPU tests must never pass on a fallback
        assert _lib.device_count() > 0, "no CUDA device visible"
        assert lib is not None
    return request.param
