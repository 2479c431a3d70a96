"""Shared machinery of the GPU-backed feature plugins.

A plugin instance owns (lazily, one per window length it is called with) a single-family
:class:`~py_neuromodulation_b200._pipeline.Pipeline` without preprocessing; ``calc_feature``
uploads the (channels, time) window, runs the family's kernels and returns the reference's
``{name: value}`` dict in the reference's insertion order.
"""

from __future__ import annotations

import numpy as np


class GpuPlugin:
    def __init__(self) -> None:
        self._pipes: dict[int, object] = {}

    def _specs(self, window_samples: int) -> list:  # pragma: no cover - interface
        raise NotImplementedError

    def _keys(self, specs) -> list[str]:
        out: list[str] = []
        for s in specs:
            out += s.keys()
        return out

    def _pipeline(self, n_ch: int, window_samples: int):
        from .._pipeline import Pipeline

        pipe = self._pipes.get(window_samples)
        if pipe is None:
            if n_ch != len(self.ch_names):
                raise ValueError(f"data has {n_ch} channels but the plugin was built for {len(self.ch_names)}")
            specs = self._specs(window_samples)
            keys = self._keys(specs)
            pipe = Pipeline(n_ch, n_ch, window_samples, keys)
            for s in specs:
                s.attach(pipe)
            pipe.finalize()
            pipe._specs = specs
            self._pipes[window_samples] = pipe
        return pipe

    def _post(self, key: str, value: float):
        return value

    def calc_feature(self, data: np.ndarray) -> dict:
        data = np.asarray(data)
        if data.ndim != 2:
            raise ValueError("data must be (channels, time)")
        pipe = self._pipeline(data.shape[0], data.shape[1])
        values = pipe.process_window(data.astype(np.float64, copy=False))
        return {k: self._post(k, v) for k, v in zip(pipe.columns, values)}
