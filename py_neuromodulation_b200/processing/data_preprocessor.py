"""Preprocessor chain (reference: ``processing/data_preprocessor.py``).

The chain is executed in the fixed order of ``PREPROCESSOR_DICT`` -- NOT in the order of the
``settings.preprocessing`` list -- exactly like the reference (data_preprocessor.py:44-52).
In scope on the GPU: ``preprocessing_filter``, ``notch_filter``, ``raw_resampling`` (any ratio; everything downstream keeps
the ORIGINAL sampling rate like the reference, stream/data_processor.py:55,77-81), ``re_referencing`` and ``raw_normalization``
(mean / median / zscore / zscore-median).  The scikit-learn raw normalisers raise NotImplementedError.
"""

from __future__ import annotations

from typing import TYPE_CHECKING

from ..utils.types import NMPreprocessor

if TYPE_CHECKING:
    import numpy as np

    from ..stream.settings import NMSettings

PREPROCESSOR_DICT: dict[str, str] = {
    "preprocessing_filter": "PreprocessingFilter",
    "notch_filter": "NotchFilter",
    "raw_resampling": "Resampler",
    "re_referencing": "ReReferencer",
    "raw_normalization": "RawNormalizer",
}


def preprocessing_plan(settings: "NMSettings", sfreq: float) -> list[str]:
    """Names of the preprocessors that actually do work, in execution order."""
    for name in settings.preprocessing:
        if name not in PREPROCESSOR_DICT:
            raise ValueError(f"Invalid preprocessing method '{name}'. Must be one of {PREPROCESSOR_DICT.keys()}")
    plan = []
    for name in PREPROCESSOR_DICT:
        if name not in settings.preprocessing:
            continue
        if name == "raw_resampling":
            # ratio 1 is the identity, like the reference (processing/resample.py:36-38)
            if float(settings.raw_resampling_settings.resample_freq_hz / sfreq) == 1.0:
                continue
        if name == "raw_normalization":
            method = settings.raw_normalization_settings.normalization_method
            if method not in ("mean", "median", "zscore", "zscore-median", "minmax", "robust"):
                # (quantile: the reference's QuantileTransformer subsamples the history at random; power: per-window likelihood search)
                raise NotImplementedError(f"raw_normalization with the scikit-learn method '{method}' is out of scope")
        plan.append(name)
    return plan


class DataPreprocessor:
    """Holds ``NMPreprocessor`` instances with the reference's per-window ``process_data`` interface."""

    def __init__(self, settings: "NMSettings", channels, sfreq: float, line_noise: float | None = None) -> None:
        from ..filter.notch_filter import NotchFilter
        from .filter_preprocessing import PreprocessingFilter
        from .normalization import RawNormalizer
        from .rereference import ReReferencer
        from .resample import Resampler

        self.preprocessors: list[NMPreprocessor] = []
        for name in preprocessing_plan(settings, sfreq):
            if name == "preprocessing_filter":
                self.preprocessors.append(PreprocessingFilter(settings=settings, sfreq=sfreq))
            elif name == "notch_filter":
                self.preprocessors.append(NotchFilter(sfreq=sfreq, line_noise=line_noise))
            elif name == "raw_resampling":
                self.preprocessors.append(Resampler(sfreq=sfreq, resample_freq_hz=settings.raw_resampling_settings.resample_freq_hz))
            elif name == "re_referencing":
                self.preprocessors.append(ReReferencer(sfreq=sfreq, channels=channels))
            elif name == "raw_normalization":
                self.preprocessors.append(RawNormalizer(sfreq=sfreq, settings=settings))

    def process_data(self, data: "np.ndarray") -> "np.ndarray":
        for pre in self.preprocessors:
            data = pre.process(data)
        return data
