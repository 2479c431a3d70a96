"""Feature registry (reference: ``features/feature_processor.py``): name -> plugin class, user features."""

from __future__ import annotations

from typing import TYPE_CHECKING, Type

from ..utils.types import NMFeature

if TYPE_CHECKING:
    import numpy as np

    from ..stream.settings import NMSettings

FEATURE_DICT: dict[str, str] = {
    "raw_hjorth": "Hjorth",
    "return_raw": "Raw",
    "bandpass_filter": "BandPower",
    "stft": "STFT",
    "fft": "FFT",
    "welch": "Welch",
    "sharpwave_analysis": "SharpwaveAnalyzer",
    "fooof": "FooofAnalyzer",
    "nolds": "Nolds",
    "coherence": "Coherence",
    "bursts": "Bursts",
    "linelength": "LineLength",
    "mne_connectivity": "MNEConnectivity",
    "bispectrum": "Bispectra",
}

OUT_OF_SCOPE = ("fooof", "nolds", "coherence", "mne_connectivity", "bispectrum")


def feature_class(feature_name: str):
    from importlib import import_module

    if feature_name in OUT_OF_SCOPE:
        raise NotImplementedError(
            f"feature '{feature_name}' is outside the B200 hot path (third-party iterative fits; SURVEY.md section 2 row 23)"
        )
    return getattr(import_module("py_neuromodulation_b200.features"), FEATURE_DICT[feature_name])


class FeatureProcessors:
    """One plugin instance per enabled feature, in ``FeatureSelector`` order, then the user features.

    This is the per-plugin path (each plugin runs its own kernels); ``DataProcessor`` fuses the
    built-in plugins into one GPU pipeline instead and only falls back to this class for user features.
    """

    def __init__(self, settings: "NMSettings", ch_names: list[str], sfreq: float) -> None:
        from .. import user_features

        self.features: dict[str, NMFeature] = {}
        for name in settings.features.get_enabled():
            if name in user_features:
                continue
            self.features[name] = feature_class(name)(settings, ch_names, sfreq)
        for name, cls in user_features.items():
            self.features[name] = cls(settings, ch_names, sfreq)

    def register_new_feature(self, feature_name: str, feature: NMFeature) -> None:
        self.features[feature_name] = feature

    def estimate_features(self, data: "np.ndarray") -> dict:
        out: dict = {}
        for feature in self.features.values():
            out.update(feature.calc_feature(data))
        return out

    def get_feature(self, fname: str) -> NMFeature:
        return self.features[fname]


def add_custom_feature(feature_name: str, new_feature: Type[NMFeature]) -> None:
    """Register a user feature class; it is enabled in every live ``NMSettings`` (reference :90-110)."""
    from .. import user_features
    from ..stream.settings import NMSettings

    user_features[feature_name] = new_feature
    NMSettings._add_feature(feature_name)


def remove_custom_feature(feature_name: str) -> None:
    from .. import user_features
    from ..stream.settings import NMSettings

    user_features.pop(feature_name)
    NMSettings._remove_feature(feature_name)
