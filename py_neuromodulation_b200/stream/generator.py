"""Window grid of the offline stream (reference: ``stream/generator.py``).

``window_grid`` returns the whole grid at once for the batched GPU path; ``RawDataGenerator`` keeps the
reference's iterator interface on top of it.  Index arithmetic is done in float64 exactly like the
reference (``stride * k`` and ``start + segment_length`` are floats, truncated with ``int``), so
non-integer strides (e.g. 3 Hz at 1 kHz -> 0, 333, 666, 1000 ...) come out identical.
"""

from __future__ import annotations

import numpy as np


def window_grid(n_samples: int, sfreq: float, sampling_rate_features_hz: float, segment_length_features_ms: float):
    """-> (starts int64[n], lengths int64[n], time_ms float64[n]) with time = ceil(t_last * 1000 + 1)."""
    seg = segment_length_features_ms / 1000 * sfreq
    stride = sfreq / sampling_rate_features_hz
    # the reference's loop (k = 0, 1, ... until the window would end behind the recording), evaluated for all k at once: the same
    # float64 products / sums / truncations element by element (2 991 windows: 5 ms as a Python loop, 0.1 ms here)
    n_max = max(0, int((n_samples - seg) / stride) + 2) if stride > 0 else 0
    k = np.arange(n_max, dtype=np.float64)
    start = stride * k
    end = start + seg
    i0 = start.astype(np.int64)
    i1 = end.astype(np.int64)
    over = np.flatnonzero(i1 > n_samples)
    n = int(over[0]) if over.size else n_max
    start, end, i0, i1 = start[:n], end[:n], i0[:n], i1[:n]
    # last element of np.arange(start, end): start + (ceil(end - start) - 1)
    n_ts = np.ceil(end - start)
    t_last = (start + (n_ts - 1)) / sfreq
    return i0, i1 - i0, np.ceil(t_last * 1000 + 1)


class RawDataGenerator:
    """Iterator over ``(timestamps, data[:, i0:i1])`` mimicking online acquisition."""

    def __init__(self, data: np.ndarray, sfreq: float, sampling_rate_features_hz: float, segment_length_features_ms: float) -> None:
        self.batch_counter = 0
        self.data = data
        self.sfreq = sfreq
        self.segment_length = segment_length_features_ms / 1000 * sfreq
        self.stride = sfreq / sampling_rate_features_hz

    def __iter__(self):
        return self

    def __next__(self):
        start = self.stride * self.batch_counter
        end = start + self.segment_length
        self.batch_counter += 1
        i0, i1 = int(start), int(end)
        if i1 > self.data.shape[1]:
            raise StopIteration
        return np.arange(start, end) / self.sfreq, self.data[:, i0:i1]
