from .bandpower import BandPower, BandPowerSettings  # noqa: F401
from .bursts import Bursts, BurstsSettings  # noqa: F401
from .hjorth_raw import Hjorth, Raw  # noqa: F401
from .linelength import LineLength  # noqa: F401
from .oscillatory import FFT, STFT, OscillatorySettings, Welch  # noqa: F401
from .out_of_scope import (  # noqa: F401
    BispectraSettings,
    CoherenceSettings,
    FooofSettings,
    MNEConnectivitySettings,
    NoldsSettings,
)
from .sharpwaves import SharpwaveAnalyzer, SharpwaveSettings  # noqa: F401
from .feature_processor import FEATURE_DICT, FeatureProcessors, add_custom_feature, remove_custom_feature  # noqa: F401
