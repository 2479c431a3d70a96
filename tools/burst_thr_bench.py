"""Burst-threshold kernel: incremental vs direct selection (time per launch, rebuild statistics).

    python tools/burst_thr_bench.py [n_ch] [duration_s]
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import py_neuromodulation_b200 as nm  # noqa: E402
from py_neuromodulation_b200.stream.generator import window_grid  # noqa: E402
from py_neuromodulation_b200.utils.channels import get_default_channels_from_data  # noqa: E402


def main():
    n_ch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    dur = int(sys.argv[2]) if len(sys.argv) > 2 else 120
    x = np.random.default_rng(0).random((n_ch, dur * 1000), dtype=np.float32)
    s = nm.NMSettings.get_default().reset()
    s.features.bursts = True
    dp = nm.DataProcessor(sfreq=1000, settings=s, channels=get_default_channels_from_data(x), line_noise=50, verbose=False)
    starts, lengths, _ = window_grid(x.shape[1], 1000, 10, 1000)
    pipe = dp.plan(1000).pipe
    pipe.upload(x)
    res = {}
    for mode in (True, False):
        pipe.set_burst_threshold_mode(mode)
        for rep in range(2):
            pipe.reset_state()
            pipe.set_profiling(rep == 1)
            out = pipe.run(starts, download=True)
        prof = pipe.profile()
        res[mode] = out.copy()
        print(f"incremental={mode}: {len(starts)} windows, chunk {pipe.chunk_windows}, threshold kernel {prof['burst_threshold'][0]:.3f} ms in "
              f"{prof['burst_threshold'][1]} launches, stats (rebuilds, direct) = {pipe.burst_threshold_stats()}")
    print("identical:", np.array_equal(res[True], res[False]))


if __name__ == "__main__":
    main()
