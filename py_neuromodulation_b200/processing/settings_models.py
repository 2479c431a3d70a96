"""Settings models of preprocessors / postprocessors that are not (yet) on the GPU path.

``ResamplerSettings``, ``FilterSettings`` (8f-3) and ``ProjectionSettings`` (out of
scope) exist so that settings files round-trip; see ``processing/data_preprocessor.py`` for what is
executed.
"""

from __future__ import annotations

from typing import Literal

from pydantic import Field

from ..utils.pydantic_extensions import NMField
from ..utils.types import BoolSelector, FrequencyRange, NMBaseModel


class ResamplerSettings(NMBaseModel):
    resample_freq_hz: float = NMField(default=1000, gt=0, custom_metadata={"unit": "Hz"})


FILTER_KEYS = ("bandstop_filter", "bandpass_filter", "lowpass_filter", "highpass_filter")


class FilterSettings(BoolSelector):
    bandstop_filter: bool = True
    bandpass_filter: bool = True
    lowpass_filter: bool = True
    highpass_filter: bool = True

    bandstop_filter_settings: FrequencyRange = FrequencyRange(100, 160)
    bandpass_filter_settings: FrequencyRange = FrequencyRange(3, 200)
    lowpass_filter_cutoff_hz: float = Field(default=200)
    highpass_filter_cutoff_hz: float = Field(default=3)

    def get_filter_tuple(self, filter_name: str):
        if filter_name == "bandstop_filter":
            # like the reference (filter_preprocessing.py:27-29): the range is returned as (low, high)
            return (self.bandstop_filter_settings.frequency_low_hz, self.bandstop_filter_settings.frequency_high_hz)
        if filter_name == "bandpass_filter":
            return (self.bandpass_filter_settings.frequency_low_hz, self.bandpass_filter_settings.frequency_high_hz)
        if filter_name == "lowpass_filter":
            return (None, self.lowpass_filter_cutoff_hz)
        if filter_name == "highpass_filter":
            return (self.highpass_filter_cutoff_hz, None)
        raise ValueError(filter_name)


class ProjectionSettings(NMBaseModel):
    max_dist_mm: float = Field(default=20.0, gt=0.0)
