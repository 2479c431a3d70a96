"""FFT / Welch / STFT band features (reference: ``features/oscillatory.py``)."""

from __future__ import annotations

from collections.abc import Sequence
from typing import TYPE_CHECKING

import numpy as np

from ..utils.pydantic_extensions import NMField
from ..utils.types import BoolSelector, NMBaseModel, NMFeature
from ._gpu_plugin import GpuPlugin

if TYPE_CHECKING:
    from ..stream.settings import NMSettings


class OscillatoryFeatures(BoolSelector):
    mean: bool = True
    median: bool = False
    std: bool = False
    max: bool = False


class OscillatorySettings(NMBaseModel):
    windowlength_ms: int = NMField(1000, gt=0, custom_metadata={"unit": "ms"})
    log_transform: bool = True
    features: OscillatoryFeatures = OscillatoryFeatures(mean=True, median=False, std=False, max=False)
    return_spectrum: bool = True


class OscillatoryFeature(GpuPlugin, NMFeature):
    osc_feature_name: str = ""
    settings_attr: str = ""

    def __init__(self, settings: "NMSettings", ch_names: Sequence[str], sfreq: int) -> None:
        settings.validate()
        self.settings: OscillatorySettings = getattr(settings, self.settings_attr)
        self.sfreq = int(sfreq)
        self.ch_names = list(ch_names)
        self.frequency_ranges = settings.frequency_ranges_hz
        assert self.settings.windowlength_ms <= settings.segment_length_features_ms, (
            f"oscillatory feature windowlength_ms = ({self.settings.windowlength_ms})needs to be smaller than"
            f"settings['segment_length_features_ms'] = {settings.segment_length_features_ms}",
        )
        self._nm_settings = settings
        GpuPlugin.__init__(self)

    def _specs(self, window_samples: int):
        from .._pipeline import SpectralSpec, band_items

        return [SpectralSpec(self.osc_feature_name, self.settings, band_items(self._nm_settings), self.ch_names, self.sfreq,
                             window_samples)]


class FFT(OscillatoryFeature):
    osc_feature_name = "fft"
    settings_attr = "fft_settings"


class Welch(OscillatoryFeature):
    osc_feature_name = "welch"
    settings_attr = "welch_settings"


class STFT(OscillatoryFeature):
    osc_feature_name = "stft"
    settings_attr = "stft_settings"
