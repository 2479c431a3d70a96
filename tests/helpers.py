"""Shared helpers for the test-suite (golden loading, tolerance rule, synthetic inputs)."""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"


def load_golden(name: str) -> dict:
    z = np.load(GOLDEN / f"{name}.npz", allow_pickle=False)
    out = {}
    for k in z.files:
        v = z[k]
        if v.dtype.kind in "US" and v.shape == ():
            s = str(v)
            try:
                out[k] = json.loads(s)
            except json.JSONDecodeError:
                out[k] = s
        elif v.shape == ():
            out[k] = v.item()
        else:
            out[k] = v
    return out


def f32(x):
    return np.asarray(x, dtype=np.float64).astype(np.float32).astype(np.float64)


def uniform(seed, c, t):
    return f32(np.random.default_rng(seed).random((c, t)))


def neural_like(seed, c, t, sfreq=1000.0):
    """pink-ish noise + 6/20 Hz bursts + 50 Hz line (same as make_golden.py)."""
    rng = np.random.default_rng(seed)
    tt = np.arange(t) / sfreq
    x = np.cumsum(rng.standard_normal((c, t)), axis=1) * 0.01 + rng.standard_normal((c, t)) * 0.05
    gate = ((tt % 1.0) < 0.3).astype(float)
    for ci in range(c):
        ph = rng.random() * 2 * np.pi
        x[ci] += gate * np.sin(2 * np.pi * 20 * tt + ph) * (0.5 + 0.5 * rng.random())
        x[ci] += np.roll(gate, int(0.5 * sfreq)) * np.sin(2 * np.pi * 6 * tt + ph)
        x[ci] += 0.5 * np.sin(2 * np.pi * 50 * tt + ci)
    return f32(x)


def assert_parity(got, ref, rtol=1e-5, what=""):
    """SURVEY.md section 8d parity rule: |got-ref| <= rtol*max(|ref|, 1) for log-type features is too
    lax for small linear features, so: pure relative where |ref| >= 1e-3, absolute rtol*1e-3... no --
    the rule used everywhere in this suite is |got-ref| <= rtol * max(|ref|, floor) with
    floor = 1 for keys the caller declares log-like and floor = scale otherwise; NaN/inf must match."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, f"{what}: shape {got.shape} vs {ref.shape}"
    nan_g, nan_r = np.isnan(got), np.isnan(ref)
    assert np.array_equal(nan_g, nan_r), f"{what}: NaN pattern differs"
    inf_r = np.isinf(ref)
    assert np.array_equal(got[inf_r], ref[inf_r]), f"{what}: inf pattern differs"
    fin = ~(nan_r | inf_r)
    err = np.abs(got[fin] - ref[fin])
    tol = rtol * np.maximum(np.abs(ref[fin]), 1.0)
    bad = err > tol
    if bad.any():
        i = np.argmax(err / tol)
        raise AssertionError(f"{what}: {bad.sum()} of {fin.sum()} values out of tolerance; worst got={got[fin][i]!r} ref={ref[fin][i]!r}")


def rel_err(got, ref):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    fin = np.isfinite(ref)
    return float(np.max(np.abs(got[fin] - ref[fin]) / np.maximum(np.abs(ref[fin]), 1e-300))) if fin.any() else 0.0


# ---------------------------------------------------------------------------------------------------------------- tolerance rule
# err = |got - ref| / max(|ref|, floor(column)):
#   * log10 outputs (FFT / Welch / STFT band features, band-pass `activity` with the log transform) and rolling z-scores are
#     differences of O(1) quantities: the ABSOLUTE error is the meaningful figure there -> floor 1;
#   * every other feature (Hjorth, line length, raw, burst amplitudes / durations, sharp-wave features, linear band power) is
#     judged PURELY RELATIVE, with a floor of 1e-12 that only matters for exact zeros.
LOG_LIKE = ("_fft_", "_welch_", "_stft_", "_bandpass_activity_")
# differences of samples / times whose operands are O(signal): cancellation makes the absolute error (relative to the operand
# scale) the figure the arithmetic can guarantee
DIFF_LIKE = ("_slope_ratio_", "_sharpness_", "_prominence_", "_raw")
REL_FLOOR = 1e-12


def column_floors(cols, normalized: bool = False, log_like=LOG_LIKE) -> np.ndarray:
    if normalized:
        return np.ones(len(cols))
    return np.array([1.0 if any(t in k for t in log_like) else (1e-3 if (k.endswith("_raw") or any(t in k for t in DIFF_LIKE[:3])) else REL_FLOOR)
                     for k in cols])


def parity_err(cols, got, ref, normalized: bool = False) -> np.ndarray:
    """Per-entry error of a (n, F) or (F,) result against the reference under the rule above (NaN / inf entries -> 0)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    floors = column_floors(list(cols), normalized)
    fin = np.isfinite(ref)
    err = np.abs(np.where(fin, got - ref, 0.0)) / np.maximum(np.abs(np.where(fin, ref, 1.0)), floors)
    return np.where(fin, err, 0.0)
