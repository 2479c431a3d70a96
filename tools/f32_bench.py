"""C3 workload in the float32 mode of the FIR families (packed float32 pairs): resident step time + per-family profile.

    python tools/f32_bench.py [steps]
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
import py_neuromodulation_b200 as nm  # noqa: E402
from py_neuromodulation_b200.stream.generator import window_grid  # noqa: E402
from py_neuromodulation_b200.utils.channels import get_default_channels_from_data  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
x = bench.synth(256, 300_000, seed=0)
s = bench.c3_settings()
ch = get_default_channels_from_data(x)
starts, lengths, _ = window_grid(x.shape[1], 1000.0, s.sampling_rate_features_hz, s.segment_length_features_ms)
res = {}
for prec in ("f64", "f32", "f32x2"):
    dp = nm.DataProcessor(sfreq=1000, settings=s, channels=ch, line_noise=50, verbose=False, precision=prec)
    pipe = dp.plan(1000).pipe
    pipe.upload(x)
    for _ in range(3):
        pipe.run(starts, download=False)
    pipe.synchronize()
    pipe.timer_start()
    for _ in range(steps):
        pipe.prepare_resident()
        pipe.run(starts, download=False)
    ms = pipe.timer_stop() / steps
    pipe.set_profiling(True)
    pipe.prepare_resident()
    pipe.run(starts, download=False)
    pipe.synchronize()
    prof = {k: round(v[0], 3) for k, v in pipe.profile().items()}
    pipe.set_profiling(False)
    res[prec] = pipe.run(starts[:64])
    cols = dp.plan(1000).columns
    print(f"{prec}: {ms:.2f} ms per step -> {starts.size / ms * 1e3:.0f} windows/s; profile {prof}")
for prec in ("f32", "f32x2"):
  a, b = res[prec], res["f64"]
  print(prec)
  err = np.abs(a - b) / np.maximum(np.abs(b), 1.0)
  wi, ci = np.unravel_index(np.argmax(err), err.shape)
  print(f"float32 vs float64 on 64 windows: max err {err.max():.3e} (gate 1e-5) at window {wi} {cols[ci]}: {a[wi, ci]!r} vs {b[wi, ci]!r}")
  fam = {}
  for j, k in enumerate(cols):
    key = "bandpass" if "_bandpass_" in k else ("fft" if "_fft_" in k else k.split("_")[-1])
    fam[key] = max(fam.get(key, 0.0), float(err[:, j].max()))
  print("  worst per family:", {k: f"{v:.2e}" for k, v in fam.items()})
