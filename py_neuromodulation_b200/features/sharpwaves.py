"""Sharp-wave features (reference: ``features/sharpwaves.py``)."""

from __future__ import annotations

from collections.abc import Sequence
from typing import TYPE_CHECKING

from pydantic import model_validator

from ..utils.types import BoolSelector, FrequencyRange, NMBaseModel, NMFeature
from ._gpu_plugin import GpuPlugin

if TYPE_CHECKING:
    from ..stream.settings import NMSettings


class PeakDetectionSettings(NMBaseModel):
    estimate: bool = True
    distance_troughs_ms: float = 10
    distance_peaks_ms: float = 5


class SharpwaveFeatures(BoolSelector):
    peak_left: bool = False
    peak_right: bool = False
    num_peaks: bool = False
    trough: bool = False
    width: bool = False
    prominence: bool = True
    interval: bool = True
    decay_time: bool = False
    rise_time: bool = False
    sharpness: bool = True
    rise_steepness: bool = False
    decay_steepness: bool = False
    slope_ratio: bool = False


class SharpwaveEstimators(NMBaseModel):
    mean: list[str] = ["interval"]
    median: list[str] = []
    max: list[str] = ["prominence", "sharpness"]
    min: list[str] = []
    var: list[str] = []

    def keys(self):
        return ["mean", "median", "max", "min", "var"]

    def values(self):
        return [self.mean, self.median, self.max, self.min, self.var]


class SharpwaveSettings(NMBaseModel):
    sharpwave_features: SharpwaveFeatures = SharpwaveFeatures()
    filter_ranges_hz: list[FrequencyRange] = [FrequencyRange(5, 80), FrequencyRange(5, 30)]
    detect_troughs: PeakDetectionSettings = PeakDetectionSettings()
    detect_peaks: PeakDetectionSettings = PeakDetectionSettings()
    estimator: SharpwaveEstimators = SharpwaveEstimators()
    apply_estimator_between_peaks_and_troughs: bool = True

    def disable_all_features(self) -> None:
        self.sharpwave_features.disable_all()
        for est in self.estimator.keys():
            self.estimator[est] = []

    @model_validator(mode="after")
    def _every_feature_has_an_estimator(self):
        listed = [ft for group in self.estimator.values() for ft in group]
        for ft in self.sharpwave_features.get_enabled():
            assert ft in listed, f"Add estimator key for {ft}"
        return self


class SharpwaveAnalyzer(GpuPlugin, NMFeature):
    def __init__(self, settings: "NMSettings", ch_names: Sequence[str], sfreq: float) -> None:
        self.sw_settings = settings.sharpwave_analysis_settings
        self.sfreq = sfreq
        self.ch_names = list(ch_names)
        settings.validate()
        self._nm_settings = settings
        GpuPlugin.__init__(self)
        self._spec = None
        self._spec = self._specs(0)[0]
        self.filter_names = [name for name, _ in self._spec.filters]
        self.used_features = self._spec.used

    def _specs(self, window_samples: int):
        from .._pipeline import SharpwaveSpec

        if self._spec is not None:
            return [self._spec]
        return [SharpwaveSpec(self._nm_settings, self.ch_names, self.sfreq)]
