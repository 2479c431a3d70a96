"""TEST INFRASTRUCTURE: build / load the thread-emulated library (tests/emu) and point the package at it.

The emulated build compiles the *same* CUDA sources with g++ and runs every CTA with host threads
(tests/emu/nm_emu.h).  It exists so that kernel logic is exercised by the CPU test-suite of a container
without a GPU; the package itself never loads it (``py_neuromodulation_b200/_lib.py`` only knows the CUDA
library) -- tests swap it in explicitly through the ``emu`` fixture.
"""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
EMU_LIB = ROOT / "tests" / "emu" / "_build" / "libnmb200_emu.so"


def build_emu() -> Path:
    subprocess.run(["make", "-C", str(ROOT / "py_neuromodulation_b200" / "csrc"), "emu"], check=True, capture_output=True)
    return EMU_LIB


def load_emu():
    from py_neuromodulation_b200 import _lib

    build_emu()
    return _lib.declare(ctypes.CDLL(str(EMU_LIB)))
