"""Rolling normalisation settings + the feature normaliser (reference: ``processing/normalization.py``)."""

from __future__ import annotations

from typing import TYPE_CHECKING, get_args

import numpy as np

from ..utils.pydantic_extensions import NMField
from ..utils.types import NORM_METHOD, NMBaseModel

if TYPE_CHECKING:
    from ..stream.settings import NMSettings

GPU_NORM_METHODS = ("mean", "median", "zscore", "zscore-median")  # raw normaliser
# feature normaliser: + the scikit-learn transformers of the reference (processing/normalization.py:58-70), restated in
# csrc/nm_norm.cuh: MinMaxScaler, RobustScaler, QuantileTransformer(n_quantiles=300); PowerTransformer stays out of scope
GPU_FEATURE_NORM_METHODS = GPU_NORM_METHODS + ("minmax", "robust", "quantile")
# raw normaliser: MinMaxScaler / RobustScaler too (csrc/nm_rawnorm.cuh); `quantile` is not reproducible in the reference itself there
# (QuantileTransformer subsamples 10 000 of the 30 000 history samples at random)
GPU_RAW_NORM_METHODS = GPU_NORM_METHODS + ("minmax", "robust")


def check_feature_norm_method(method: str, n_keep: int) -> None:
    if method not in GPU_FEATURE_NORM_METHODS:
        raise NotImplementedError(f"normalisation method '{method}' (scikit-learn PowerTransformer) is out of scope")
    if method == "quantile" and n_keep > 300:
        raise NotImplementedError("the quantile normaliser covers histories of at most 300 windows (n_quantiles = 300): "
                                  "normalization_time_s * sampling_rate_features_hz <= 300")


class NormalizationSettings(NMBaseModel):
    normalization_time_s: float = NMField(30, gt=0, custom_metadata={"unit": "s"})
    normalization_method: NORM_METHOD = NMField(default="zscore")
    clip: float = NMField(default=3, ge=0, custom_metadata={"unit": "a.u."})

    @staticmethod
    def list_normalization_methods() -> list[str]:
        return list(get_args(NORM_METHOD))


class FeatureNormalizationSettings(NormalizationSettings):
    normalize_psd: bool = False


class FeatureNormalizer:
    """Stand-alone feature normaliser with the reference's ``process(vector) -> vector`` interface.

    History and arithmetic live on the GPU (``csrc/nm_norm.cuh``); inside ``DataProcessor`` the same kernel is
    part of the fused pipeline instead.
    """

    def __init__(self, settings: "NMSettings") -> None:
        self.settings = settings.feature_normalization_settings.validate()
        self.method = self.settings.normalization_method
        self.num_samples_normalize = int(self.settings.normalization_time_s * settings.sampling_rate_features_hz)
        check_feature_norm_method(self.method, self.num_samples_normalize)
        self._pipe = None

    def process(self, data: np.ndarray) -> np.ndarray:
        from .._pipeline import IdentityNormPipeline

        v = np.asarray(data, dtype=np.float64).ravel()
        if self._pipe is None:
            self._pipe = IdentityNormPipeline(v.size, GPU_FEATURE_NORM_METHODS.index(self.method), float(self.settings.clip or 0.0),
                                              self.num_samples_normalize)
        return self._pipe.step(v)


class RawNormalizer:
    """Rolling normalisation of the preprocessed samples (reference class of the same name, ``process(data) -> data``).

    Stateful like the reference: window 0 passes through and seeds the history, later windows append their last
    ``int(sfreq / sampling_rate_features_hz)`` samples and are normalised against the whole history
    (``csrc/nm_rawnorm.cuh``; the medians come from the sliding order-statistic kernel of the burst thresholds).
    Of the scikit-learn transformers ``minmax`` and ``robust`` are restated on the GPU; ``quantile`` / ``power`` raise ``NotImplementedError``.
    """

    GPU_METHODS = GPU_RAW_NORM_METHODS

    def __init__(self, sfreq: float, settings: "NMSettings", **kwargs) -> None:
        self.settings = settings.raw_normalization_settings.validate()
        self.method = self.settings.normalization_method
        if self.method not in self.GPU_METHODS:
            raise NotImplementedError(
                f"raw normalisation method '{self.method}' is not on the B200 path (supported: {', '.join(self.GPU_METHODS)})")
        self.add_samples = int(sfreq / settings.sampling_rate_features_hz)
        self.num_samples_normalize = int(self.settings.normalization_time_s * sfreq)
        self._pipes: dict = {}

    def process(self, data: np.ndarray) -> np.ndarray:
        from .._pipeline import Pipeline, ScanSpec

        data = np.asarray(data, dtype=np.float64)
        n_ch, w = data.shape
        pipe = self._pipes.get((n_ch, w))
        if pipe is None:
            pipe = Pipeline(n_ch, n_ch, w, ["_unused"])
            pipe.set_raw_normalizer(self.method, self.settings.clip, self.num_samples_normalize, self.add_samples)
            ScanSpec([f"_c{i}" for i in range(n_ch)]).attach(pipe)
            pipe.finalize()
            self._pipes[(n_ch, w)] = pipe
        return pipe.preprocess_window(data)
