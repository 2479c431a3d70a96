"""Core protocol types of the plugin API (reference: ``utils/types.py``)."""

from __future__ import annotations

from collections.abc import Sequence
from math import isnan
from os import PathLike
from typing import TYPE_CHECKING, Literal

from pydantic import Field, model_validator

from .pydantic_extensions import NMBaseModel, NMField  # noqa: F401  (re-exported)

if TYPE_CHECKING:
    import numpy as np

_PathLike = str | PathLike

FEATURE_NAME = Literal[
    "raw_hjorth", "return_raw", "bandpass_filter", "stft", "fft", "welch", "sharpwave_analysis",
    "fooof", "nolds", "coherence", "bursts", "linelength", "mne_connectivity", "bispectrum",
]
PREPROCESSOR_NAME = Literal["preprocessing_filter", "notch_filter", "raw_resampling", "re_referencing", "raw_normalization"]
NORM_METHOD = Literal["mean", "median", "zscore", "zscore-median", "quantile", "power", "robust", "minmax"]


class NMFeature:
    """Feature plugin protocol (reference ``utils/types.py:59-77``): duck-typed ctor + ``calc_feature``."""

    def __init__(self, settings, ch_names: Sequence[str], sfreq: int | float) -> None: ...

    def calc_feature(self, data: "np.ndarray") -> dict:
        """(channels, time) window -> ``{feature_name: value}``"""
        ...


class NMPreprocessor:
    """Preprocessor protocol (reference ``utils/types.py:80-81``)."""

    def process(self, data: "np.ndarray") -> "np.ndarray": ...


class FrequencyRange(NMBaseModel):
    frequency_low_hz: float = Field(gt=0)
    frequency_high_hz: float = Field(gt=0)

    def __getitem__(self, item):  # type: ignore[override]
        if item == 0:
            return self.frequency_low_hz
        if item == 1:
            return self.frequency_high_hz
        if isinstance(item, str):
            return getattr(self, item)
        raise IndexError(f"Index {item} out of range")

    def as_tuple(self) -> tuple[float, float]:
        return (self.frequency_low_hz, self.frequency_high_hz)

    def __iter__(self):  # type: ignore[override]
        return iter(self.as_tuple())

    @model_validator(mode="before")
    @classmethod
    def _from_pair(cls, value):
        if isinstance(value, dict):
            if "frequency_low_hz" in value and "frequency_high_hz" in value:
                return value
        elif isinstance(value, Sequence) and not isinstance(value, str) and len(value) == 2:
            return {"frequency_low_hz": value[0], "frequency_high_hz": value[1]}
        raise ValueError(
            f"Value for FrequencyRange must be a dictionary, or a sequence of 2 numeric values, but got {value} instead."
        )

    @model_validator(mode="after")
    def _ordered(self):
        if not (isnan(self.frequency_high_hz) or isnan(self.frequency_low_hz)):
            assert self.frequency_high_hz > self.frequency_low_hz, "Frequency high must be greater than frequency low"
        return self


class BoolSelector(NMBaseModel):
    """A model whose boolean fields switch things on and off, in declaration order."""

    def get_enabled(self) -> list[str]:
        out = [name for name in type(self).model_fields if isinstance(self[name], bool) and self[name]]
        extra = self.__pydantic_extra__ or {}
        out += [name for name, v in extra.items() if isinstance(v, bool) and v]
        return out

    def enable_all(self) -> None:
        for name in type(self).model_fields:
            if isinstance(self[name], bool):
                self[name] = True

    def disable_all(self) -> None:
        for name in type(self).model_fields:
            if isinstance(self[name], bool):
                self[name] = False

    def __iter__(self):  # type: ignore[override]
        return iter(self.model_dump().keys())

    @classmethod
    def list_all(cls) -> list[str]:
        return list(cls.model_fields.keys())

    @classmethod
    def print_all(cls) -> None:
        for name in cls.list_all():
            print(name)
