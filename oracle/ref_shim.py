"""TEST INFRASTRUCTURE (oracle) -- import the UNMODIFIED reference hot-path modules.

Needs the reference files: ``/root/reference`` (authoring container only) or the unmodified copy that
``oracle/build_ref.py`` places under the git-ignored ``baseline/_ref`` so that it travels to the GPU box.
It is used by ``tests/golden/make_golden.py`` to generate the committed fixtures, by CPU tests (skipped
when the reference is absent) that pin ``oracle/np_oracle.py`` against the live reference, and by the
CPU legs of ``bench.py`` (``cpu_baseline`` / ``--impl reference``); never by ``-m gpu`` tests or ``smoke()``.

Why a shim: ``import py_neuromodulation`` fails here (``__init__.py:12`` needs
package metadata; ``stream/__init__.py:2`` imports ``mne``; ``__init__.py:77,88``
pull matplotlib / the GUI).  The hot-path files themselves import fine once

* bare package objects are pre-created for ``py_neuromodulation``,
  ``py_neuromodulation.stream`` and ``py_neuromodulation.analysis`` so their
  ``__init__`` does not run, and
* ``mne.filter`` resolves to ``oracle/mne_filter_restated.py``.

No reference source is copied; modules are executed from where they lie.
"""

from __future__ import annotations

import importlib
import logging
import sys
import types
from pathlib import Path, PurePath

_PKG = "py_neuromodulation"
# the authoring container has the reference checkout; elsewhere (GPU box) the same unmodified files may have been placed under the
# git-ignored baseline/_ref by oracle/build_ref.py -- only bench.py's CPU legs use that copy
_CANDIDATES = (Path("/root/reference"), Path(__file__).resolve().parents[1] / "baseline" / "_ref")


def _find_root() -> Path:
    for r in _CANDIDATES:
        if (r / _PKG / "stream" / "data_processor.py").is_file():
            return r
    return _CANDIDATES[0]


REFERENCE_ROOT = _find_root()


def reference_available() -> bool:
    return (REFERENCE_ROOT / _PKG / "stream" / "data_processor.py").is_file()


def _bare_package(name: str, path: Path) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__path__ = [str(path)]  # type: ignore[attr-defined]
    mod.__package__ = name
    sys.modules[name] = mod
    return mod


def _install_mne_stub() -> None:
    if "mne" in sys.modules and not getattr(sys.modules["mne"], "__nm_oracle_stub__", False):
        return  # a real MNE is present: use it
    from oracle import mne_filter_restated as restated

    mne = types.ModuleType("mne")
    mne.__nm_oracle_stub__ = True  # type: ignore[attr-defined]
    mne.__path__ = []  # type: ignore[attr-defined]
    filt = types.ModuleType("mne.filter")
    filt.create_filter = restated.create_filter  # type: ignore[attr-defined]
    filt._overlap_add_filter = restated._overlap_add_filter  # type: ignore[attr-defined]
    filt.resample = restated.resample  # type: ignore[attr-defined]
    mne.filter = filt  # type: ignore[attr-defined]
    sys.modules["mne"] = mne
    sys.modules["mne.filter"] = filt


def load_reference() -> types.ModuleType:
    """Return the shimmed ``py_neuromodulation`` package object (idempotent)."""
    if not reference_available():
        raise RuntimeError("reference tree not present (neither /root/reference nor baseline/_ref)")
    if _PKG in sys.modules and getattr(sys.modules[_PKG], "__nm_oracle_shim__", False):
        return sys.modules[_PKG]
    try:
        import mne  # noqa: F401
    except ImportError:
        _install_mne_stub()

    root = REFERENCE_ROOT / _PKG
    nm = _bare_package(_PKG, root)
    nm.__nm_oracle_shim__ = True  # type: ignore[attr-defined]
    nm.PYNM_DIR = PurePath(root)  # type: ignore[attr-defined]
    nm.user_features = {}  # type: ignore[attr-defined]

    log_mod = importlib.import_module(f"{_PKG}.utils.logging")
    nm.logger = log_mod.NMLogger(_PKG)  # type: ignore[attr-defined]
    nm.logger.setLevel(logging.WARNING)

    _bare_package(f"{_PKG}.stream", root / "stream")
    _bare_package(f"{_PKG}.analysis", root / "analysis")

    settings_mod = importlib.import_module(f"{_PKG}.stream.settings")
    nm.NMSettings = settings_mod.NMSettings  # type: ignore[attr-defined]
    nm.features = importlib.import_module(f"{_PKG}.features")  # type: ignore[attr-defined]
    nm.processing = importlib.import_module(f"{_PKG}.processing")  # type: ignore[attr-defined]
    nm.filter = importlib.import_module(f"{_PKG}.filter")  # type: ignore[attr-defined]
    nm.utils = importlib.import_module(f"{_PKG}.utils")  # type: ignore[attr-defined]
    nm.utils.channels = importlib.import_module(f"{_PKG}.utils.channels")  # type: ignore[attr-defined]
    nm.io = importlib.import_module(f"{_PKG}.utils.io")  # type: ignore[attr-defined]
    nm.utils.io = nm.io  # type: ignore[attr-defined]
    dp = importlib.import_module(f"{_PKG}.stream.data_processor")
    nm.DataProcessor = dp.DataProcessor  # type: ignore[attr-defined]
    gen = importlib.import_module(f"{_PKG}.stream.generator")
    nm.RawDataGenerator = gen.RawDataGenerator  # type: ignore[attr-defined]
    fp = importlib.import_module(f"{_PKG}.features.feature_processor")
    nm.add_custom_feature = fp.add_custom_feature  # type: ignore[attr-defined]
    nm.remove_custom_feature = fp.remove_custom_feature  # type: ignore[attr-defined]
    return nm


def load_reference_stream():
    """Also import the unmodified ``stream/stream.py`` (needs sklearn via analysis.decode)."""
    nm = load_reference()
    if not hasattr(nm, "Stream"):
        importlib.import_module(f"{_PKG}.analysis.decode")
        st = importlib.import_module(f"{_PKG}.stream.stream")
        nm.Stream = st.Stream  # type: ignore[attr-defined]
    return nm
