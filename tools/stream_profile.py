"""Where does nm.Stream.run spend its wall-clock time?  cProfile of the call a reference user makes (C3 workload).

    python tools/stream_profile.py [c3|default] [n_channels] [seconds]
"""
import cProfile
import pstats
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
import py_neuromodulation_b200 as nm  # noqa: E402


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
    n_ch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    secs = int(sys.argv[3]) if len(sys.argv) > 3 else 300
    settings = bench.c3_settings() if cfg == "c3" else nm.NMSettings.get_default()
    x = bench.synth(n_ch, secs * 1000, seed=0)
    stream = nm.Stream(sfreq=1000, data=x, settings=settings, line_noise=50, verbose=False)
    with tempfile.TemporaryDirectory() as td:
        stream.run(out_dir=td, experiment_name="p", save_csv=False)
        t0 = time.perf_counter()
        stream.run(out_dir=td, experiment_name="p", save_csv=False)
        print(f"Stream.run wall: {1e3 * (time.perf_counter() - t0):.1f} ms")
        pr = cProfile.Profile()
        pr.enable()
        stream.run(out_dir=td, experiment_name="p", save_csv=False)
        pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(45)


if __name__ == "__main__":
    main()
