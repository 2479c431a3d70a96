"""Per-family kernel time (CUDA events, one untimed profiling pass) for a named configuration.

    python tools/profile_families.py default 256 60      # all default features, 256 ch, 60 s @ 1 kHz
    python tools/profile_families.py c3 256 300
"""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import py_neuromodulation_b200 as nm  # noqa: E402
from py_neuromodulation_b200.stream.generator import window_grid  # noqa: E402
from py_neuromodulation_b200.utils.channels import get_default_channels_from_data  # noqa: E402


def main():
    cfg, n_ch, dur = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    sfreq = float(sys.argv[4]) if len(sys.argv) > 4 else 1000.0
    s = nm.NMSettings.get_default()
    if cfg == "c3":
        s.reset()
        for f in ("fft", "bandpass_filter", "raw_hjorth", "linelength"):
            s.features[f] = True
    elif cfg == "c4":
        s.reset()
        for f in ("fft", "welch", "stft", "bursts", "sharpwave_analysis"):
            s.features[f] = True
    s.raw_resampling_settings.resample_freq_hz = sfreq
    x = np.random.default_rng(0).random((n_ch, int(dur * sfreq)), dtype=np.float32)
    dp = nm.DataProcessor(sfreq=sfreq, settings=s, channels=get_default_channels_from_data(x), line_noise=50, verbose=False)
    starts, lengths, _ = window_grid(x.shape[1], sfreq, s.sampling_rate_features_hz, s.segment_length_features_ms)
    plan = dp.plan(int(lengths[0]))
    pipe = plan.pipe
    pipe.upload(x)
    for _ in range(2):
        pipe.reset_state()
        pipe.run(starts, download=False)
    pipe.synchronize()
    pipe.reset_state()
    t0 = time.perf_counter()
    pipe.timer_start()
    pipe.run(starts, download=False)
    ms = pipe.timer_stop()
    wall = (time.perf_counter() - t0) * 1e3
    stats = pipe.burst_threshold_stats() if s.features.bursts else None
    pipe.reset_state()
    pipe.set_profiling(True)
    pipe.run(starts, download=False)
    pipe.synchronize()
    prof = pipe.profile()
    print(f"{cfg}: {n_ch} ch x {dur} s @ {sfreq:g} Hz, {starts.size} windows, F = {pipe.F}, chunk = {pipe.chunk_windows} windows")
    print(f"  device time {ms:.2f} ms (wall {wall:.2f} ms) -> {starts.size / ms * 1e3:.0f} windows/s")
    if stats:
        print(f"  burst thresholds: {stats[0]} bracket rebuilds, {stats[1]} windows by direct selection of {starts.size * n_ch * len(s.bursts_settings.frequency_bands)} row-windows")
    for k, (t, n) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        print(f"  {k:16s} {t:9.3f} ms  {n:5d} launches")


if __name__ == "__main__":
    main()
