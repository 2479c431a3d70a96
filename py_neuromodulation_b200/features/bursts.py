"""Burst features (reference: ``features/bursts.py``)."""

from __future__ import annotations

from collections.abc import Sequence
from typing import TYPE_CHECKING

from pydantic import field_validator

from ..utils.pydantic_extensions import NMField, create_validation_error
from ..utils.types import BoolSelector, NMBaseModel, NMFeature
from ._gpu_plugin import GpuPlugin

if TYPE_CHECKING:
    from ..stream.settings import NMSettings


class BurstFeatures(BoolSelector):
    duration: bool = True
    amplitude: bool = True
    burst_rate_per_s: bool = True
    in_burst: bool = True


class BurstsSettings(NMBaseModel):
    threshold: float = NMField(default=75, ge=0)
    time_duration_s: float = NMField(default=30, ge=0, custom_metadata={"unit": "s"})
    frequency_bands: list[str] = ["low_beta", "high_beta", "low_gamma"]
    burst_features: BurstFeatures = BurstFeatures()

    @field_validator("frequency_bands")
    @classmethod
    def _underscores(cls, bands):
        return [b.replace(" ", "_") for b in bands]


def check_burst_bands(settings: "NMSettings") -> None:
    for band in settings.bursts_settings.frequency_bands:
        if band not in list(settings.frequency_ranges_hz.keys()):
            raise create_validation_error(
                f"bursting {band} needs to be defined in settings['frequency_ranges_hz']",
                location=["burst_settings", "frequency_bands"],
            )


class Bursts(GpuPlugin, NMFeature):
    """Stateful: the envelope history lives in the GPU pipeline and survives between calls, like the
    reference's ring buffer.  The history is a true ring of the last ``time_duration_s``."""

    def __init__(self, settings: "NMSettings", ch_names: Sequence[str], sfreq: float) -> None:
        settings.validate()
        check_burst_bands(settings)
        self.settings = settings.bursts_settings
        self.sfreq = sfreq
        self.ch_names = list(ch_names)
        self._nm_settings = settings
        GpuPlugin.__init__(self)
        self._spec = None
        self._spec = self._specs(0)[0]
        self.fband_names = self._spec.bands
        self.samples_overlap = self._spec.samples_overlap
        self.num_max_samples_ring_buffer = self._spec.ring

    def _specs(self, window_samples: int):
        from .._pipeline import BurstsSpec

        if self._spec is not None:
            return [self._spec]
        return [BurstsSpec(self._nm_settings, self.ch_names, self.sfreq)]

    def _post(self, key: str, value: float):
        return bool(value) if key.endswith("_in_burst") else value
