"""Per-window latency of the streaming entry (DataProcessor.process -> nm_process_window), default settings.

    python tools/stream_latency.py [n_channels ...]
"""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import py_neuromodulation_b200 as nm  # noqa: E402
from py_neuromodulation_b200.utils.channels import get_default_channels_from_data  # noqa: E402


def main():
    chans = [int(a) for a in sys.argv[1:]] or [8, 64, 256]
    for n_ch in chans:
        x = np.random.default_rng(0).random((n_ch, 1000 + 100 * 400))
        s = nm.NMSettings.get_default()
        dp = nm.DataProcessor(sfreq=1000, settings=s, channels=get_default_channels_from_data(x), line_noise=50, verbose=False)
        plan = dp.plan(1000)
        times_dict, times_raw = [], []
        for k in range(400):
            w = x[:, 100 * k : 100 * k + 1000]
            t0 = time.perf_counter()
            dp.process(w)
            times_dict.append(time.perf_counter() - t0)
        dp2 = nm.DataProcessor(sfreq=1000, settings=s, channels=get_default_channels_from_data(x), line_noise=50, verbose=False)
        pipe = dp2.plan(1000).pipe
        for k in range(400):
            w = np.ascontiguousarray(x[:, 100 * k : 100 * k + 1000])
            t0 = time.perf_counter()
            pipe.process_window(w)
            times_raw.append(time.perf_counter() - t0)
        td, tr = np.array(times_dict[50:]) * 1e3, np.array(times_raw[50:]) * 1e3
        print(f"{n_ch:4d} ch, F = {plan.pipe.F}: DataProcessor.process median {np.median(td):.3f} ms (p95 {np.percentile(td, 95):.3f}); "
              f"nm_process_window median {np.median(tr):.3f} ms (p95 {np.percentile(tr, 95):.3f}); launches/window "
              f"{pipe.kernel_launches / 400:.1f}")


if __name__ == "__main__":
    main()
