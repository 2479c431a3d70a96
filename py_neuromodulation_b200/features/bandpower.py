"""Band-pass power features (reference: ``features/bandpower.py``)."""

from __future__ import annotations

from collections.abc import Sequence
from typing import TYPE_CHECKING

from pydantic import field_validator

from ..utils.pydantic_extensions import NMErrorList, NMField, create_validation_error
from ..utils.types import BoolSelector, NMBaseModel, NMFeature
from ._gpu_plugin import GpuPlugin

if TYPE_CHECKING:
    from ..stream.settings import NMSettings


class BandpowerFeatures(BoolSelector):
    activity: bool = True
    mobility: bool = False
    complexity: bool = False


class BandPowerSettings(NMBaseModel):
    segment_lengths_ms: dict[str, int] = NMField(
        default={"theta": 1000, "alpha": 500, "low beta": 333, "high beta": 333, "low gamma": 100, "high gamma": 100, "HFA": 100},
        custom_metadata={"field_type": "FrequencySegmentLength"},
    )
    bandpower_features: BandpowerFeatures = BandpowerFeatures()
    log_transform: bool = True
    kalman_filter: bool = False

    @field_validator("bandpower_features")
    @classmethod
    def _at_least_one(cls, value: BandpowerFeatures):
        if not value.get_enabled():
            raise create_validation_error(
                error_message="Set at least one bandpower_feature to True.",
                location=["bandpass_filter_settings", "bandpower_features"],
            )
        return value

    def validate_fbands(self, settings: "NMSettings") -> NMErrorList:
        from .. import logger

        errors = NMErrorList()
        for band, seg in self.segment_lengths_ms.items():
            if band not in settings.frequency_ranges_hz:
                logger.info(
                    f"Frequency band {band} in bandpass_filter_settings.segment_lengths_ms is not defined in "
                    "settings.frequency_ranges_hz"
                )
            if not seg <= settings.segment_length_features_ms:
                errors.add_error(
                    f"segment length {seg} needs to be smaller than  settings['segment_length_features_ms'] = "
                    f"{settings.segment_length_features_ms}",
                    location=["bandpass_filter_settings", "segment_lengths_ms", band],
                )
        for band in settings.frequency_ranges_hz.keys():
            if band not in self.segment_lengths_ms:
                errors.add_error(
                    f"frequency range {band} needs to be defined in settings.bandpass_filter_settings.segment_lengths_ms",
                    location=["bandpass_filter_settings", "segment_lengths_ms", band],
                )
        return errors


class BandPower(GpuPlugin, NMFeature):
    def __init__(self, settings: "NMSettings", ch_names: Sequence[str], sfreq: float, use_kf: bool | None = None) -> None:
        settings.validate()
        self.bp_settings: BandPowerSettings = settings.bandpass_filter_settings
        self.sfreq = sfreq
        self.ch_names = list(ch_names)
        if use_kf or (use_kf is None and self.bp_settings.kalman_filter):
            raise NotImplementedError("Kalman-filtered band power is out of scope of the B200 hot path")
        self._nm_settings = settings
        GpuPlugin.__init__(self)
        self._spec = self._specs(0)[0]  # designs the FIR bank now, like the reference constructor
        self.feature_params = self._spec.keys()

    def _specs(self, window_samples: int):
        from .._pipeline import BandpowerSpec, band_items

        if getattr(self, "_spec", None) is not None:
            return [self._spec]
        return [BandpowerSpec(self.bp_settings, band_items(self._nm_settings), self.ch_names, self.sfreq)]
