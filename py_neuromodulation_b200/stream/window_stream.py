"""Streaming driver for live sources (reference: the per-batch loop of ``stream/stream.py:280-330`` fed by
``stream/mnelsl_stream.py``).

``WindowStream`` wraps the C ABI's slot ring (``nm_stream_*``): a producer writes every incoming window into a page-locked slot
(``next_input()``), ``submit()`` enqueues it without blocking -- the transfer and kernels of window g + 1 overlap whatever the host
does with window g -- and ``collect()`` returns the oldest outstanding feature row.  Rows are appended to one pre-allocated
float64 matrix (no per-window dict, no msgpack round trip); ``to_frame()`` hands out the same DataFrame ``Stream.run`` returns.

    ws = WindowStream(data_processor, window_samples=1000)
    for window in source:                 # (n_raw_rows, 1000) arrays
        ws.next_input()[...] = window     # or let the source write into the slot directly
        ws.submit(time_ms)
        if ws.in_flight == ws.slots:
            row = ws.collect()            # float64 view, one entry per ws.columns
    ws.drain(); df = ws.to_frame()
"""

from __future__ import annotations

from collections import deque

import numpy as np


class WindowStream:
    def __init__(self, data_processor, window_samples: int, slots: int = 2, f32: bool = False, graph: bool | None = None,
                 capacity: int = 4096) -> None:
        if data_processor.user_feature_names:
            raise NotImplementedError("user-defined Python features need the per-window dict path (DataProcessor.process)")
        self.dp = data_processor
        self.plan = data_processor.plan(int(window_samples))
        if self.plan.pipe is None:
            raise ValueError("no feature is enabled")
        self.pipe = self.plan.pipe
        self.columns = list(self.plan.columns)
        self.slots = int(slots)
        self.pipe.stream_open(self.slots, f32=f32, graph=graph)
        self._next = 0
        self._pending: deque[tuple[int, float]] = deque()
        self._rows = np.empty((int(capacity), len(self.columns) + 1), dtype=np.float64)  # features + time
        self.n_rows = 0

    @property
    def in_flight(self) -> int:
        return len(self._pending)

    def next_input(self) -> np.ndarray:
        """Page-locked ``(n_raw_rows, window_samples)`` block of the slot the next ``submit`` will send."""
        if len(self._pending) == self.slots:
            raise RuntimeError("every slot holds an un-collected window: call collect() first")
        return self.pipe.stream_input(self._next)

    def submit(self, time_ms: float = float("nan")) -> None:
        self.pipe.stream_submit(self._next)
        self._pending.append((self._next, float(time_ms)))
        self._next = (self._next + 1) % self.slots

    def collect(self) -> np.ndarray:
        """Feature row of the oldest outstanding window (blocks until it is there); also appended to the table."""
        slot, t = self._pending.popleft()
        row = self.pipe.stream_wait(slot)
        if self.n_rows == self._rows.shape[0]:
            self._rows = np.concatenate([self._rows, np.empty_like(self._rows)])
        dst = self._rows[self.n_rows]
        dst[:-1] = row
        dst[-1] = t
        self.n_rows += 1
        return dst[:-1]

    def drain(self) -> None:
        while self._pending:
            self.collect()

    def process(self, window: np.ndarray, time_ms: float = float("nan")) -> np.ndarray:
        """Synchronous convenience: copy ``window`` into the next slot, submit, wait."""
        self.next_input()[...] = window
        self.submit(time_ms)
        return self.collect()

    def to_frame(self):
        import pandas as pd

        return pd.DataFrame(self._rows[: self.n_rows], columns=self.columns + ["time"], copy=False)

    def close(self) -> None:
        self.drain()
        self.pipe.stream_close()
