"""Settings model of the (out-of-scope) Kalman smoother, kept so that settings files round-trip."""

from __future__ import annotations

from typing import TYPE_CHECKING

from ..utils.pydantic_extensions import NMErrorList
from ..utils.types import NMBaseModel

if TYPE_CHECKING:
    from ..stream.settings import NMSettings


class KalmanSettings(NMBaseModel):
    Tp: float = 0.1
    sigma_w: float = 0.7
    sigma_v: float = 1.0
    frequency_bands: list[str] = ["theta", "alpha", "low_beta", "high_beta", "low_gamma", "high_gamma", "HFA"]

    def validate_fbands(self, settings: "NMSettings") -> NMErrorList:
        errors = NMErrorList()
        if not all(item in settings.frequency_ranges_hz for item in self.frequency_bands):
            errors.add_error(
                "Frequency bands for Kalman filter must also be specified in bandpass_filter_settings.",
                location=["kalman_filter_settings", "frequency_bands"],
            )
        return errors
