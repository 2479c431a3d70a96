"""py_neuromodulation_b200 -- B200-native (sm_100a) implementation of py_neuromodulation's per-window
feature-extraction hot path behind the reference's plugin API.

``import py_neuromodulation_b200 as nm`` mirrors the names a user of the reference needs for that path:
``nm.Stream``, ``nm.NMSettings``, ``nm.DataProcessor``, the feature plugin classes, ``nm.add_custom_feature``.
All sample arithmetic runs in ``csrc/libnmb200.so`` (CUDA); there is no CPU fallback.
"""

from __future__ import annotations

import logging as _logging
from pathlib import PurePath

__version__ = "0.1.0"

PYNM_DIR = PurePath(__file__).parent

# user-defined feature classes registered through add_custom_feature
user_features: dict = {}

from .utils.logging import NMLogger  # noqa: E402

logger = NMLogger(__name__, level=_logging.WARNING)

from .stream.settings import NMSettings, get_default_settings, get_fast_compute, reset_settings  # noqa: E402
from .stream.data_processor import DataProcessor  # noqa: E402
from .stream.stream import Stream  # noqa: E402
from .features.feature_processor import FeatureProcessors, add_custom_feature, remove_custom_feature  # noqa: E402
from .utils import types, io  # noqa: E402
from . import stream, features, filter, processing, utils  # noqa: E402

from .features import (  # noqa: E402
    BandPower, BandPowerSettings, Bursts, BurstsSettings, FFT, STFT, Welch, OscillatorySettings, Hjorth, Raw, LineLength,
    SharpwaveAnalyzer, SharpwaveSettings, BispectraSettings, CoherenceSettings, FooofSettings, MNEConnectivitySettings,
    NoldsSettings,
)

__all__ = [
    "Stream", "DataProcessor", "NMSettings", "FeatureProcessors", "add_custom_feature", "remove_custom_feature",
    "get_default_settings", "get_fast_compute", "reset_settings", "logger", "user_features", "PYNM_DIR",
    "FFT", "Welch", "STFT", "BandPower", "Hjorth", "Raw", "LineLength", "Bursts", "SharpwaveAnalyzer",
]
