"""Hjorth parameters and last raw sample (reference: ``features/hjorth_raw.py``)."""

from __future__ import annotations

from collections.abc import Sequence

from ..utils.types import NMFeature
from ._gpu_plugin import GpuPlugin


class Hjorth(GpuPlugin, NMFeature):
    def __init__(self, settings, ch_names: Sequence[str], sfreq: float) -> None:
        self.ch_names = list(ch_names)
        GpuPlugin.__init__(self)

    def _specs(self, window_samples: int):
        from .._pipeline import ScanSpec

        return [ScanSpec(self.ch_names, hjorth=True)]

    def _keys(self, specs):
        return specs[0].keys_hjorth()


class Raw(GpuPlugin, NMFeature):
    def __init__(self, settings, ch_names: Sequence[str], sfreq: float) -> None:
        self.ch_names = list(ch_names)
        GpuPlugin.__init__(self)

    def _specs(self, window_samples: int):
        from .._pipeline import ScanSpec

        return [ScanSpec(self.ch_names, raw=True)]

    def _keys(self, specs):
        return specs[0].keys_raw()
