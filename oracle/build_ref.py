"""TEST INFRASTRUCTURE (oracle) -- make the UNMODIFIED reference available to the CPU arm of ``bench.py`` on the GPU box.

``/root/reference`` exists only in the authoring container.  The offline install the contract names,

    python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref /root/reference

fails here (the reference's build backend ``hatchling`` is not in the image and there is no network), so this script places
what that install would have placed: the package's own files, byte for byte, under the git-ignored ``baseline/_ref/`` (listed in
.gitignore, not in .gpurunignore, so it travels with ``gpurun`` / the driver's snapshot but never enters the history).  Nothing
in the product imports it; ``oracle/ref_shim.py`` loads it exactly like ``/root/reference`` (bare package objects + the
restated ``mne.filter`` entry points -- MNE itself is not installed).  Only the closure of the hot path is placed.
"""
from __future__ import annotations

import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
SRC = Path("/root/reference")
DST = ROOT / "baseline" / "_ref"
PKG = "py_neuromodulation"
# sub-packages the shimmed hot path imports (oracle/ref_shim.py); GUI, analysis plots, LSL and the examples stay behind
PARTS = ["features", "filter", "processing", "utils", "stream/settings.py", "stream/data_processor.py", "stream/generator.py",
         "stream/stream.py", "stream/backend_interface.py", "analysis/decode.py", "default_settings.yaml"]


def try_pip() -> str | None:
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links", "/opt/wheelhouse",
           "--target", str(DST), str(SRC)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode == 0 and (DST / PKG / "stream" / "data_processor.py").is_file():
        return None
    tail = (r.stderr or r.stdout).strip().splitlines()
    return tail[-1] if tail else f"pip exited with {r.returncode}"


def place() -> None:
    if not (SRC / PKG).is_dir():
        print("[build_ref] /root/reference not present: keeping whatever baseline/_ref holds")
        return
    if DST.exists():
        shutil.rmtree(DST)
    DST.mkdir(parents=True)
    why = try_pip()
    if why is None:
        (DST / "PROVENANCE.txt").write_text("pip install --target of /root/reference\n")
        return
    for part in PARTS:
        s, d = SRC / PKG / part, DST / PKG / part
        d.parent.mkdir(parents=True, exist_ok=True)
        if s.is_dir():
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        else:
            shutil.copy2(s, d)
    (DST / "PROVENANCE.txt").write_text(
        "Unmodified files of /root/reference/py_neuromodulation (hot-path closure), placed by oracle/build_ref.py because\n"
        f"the offline pip install failed: {why}\n")
    print(f"[build_ref] pip install unavailable ({why}); placed the hot-path closure under {DST}")


if __name__ == "__main__":
    place()
