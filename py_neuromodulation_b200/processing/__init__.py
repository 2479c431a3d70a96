from .data_preprocessor import PREPROCESSOR_DICT, DataPreprocessor  # noqa: F401
from .filter_preprocessing import PreprocessingFilter  # noqa: F401
from .normalization import FeatureNormalizationSettings, FeatureNormalizer, NormalizationSettings, RawNormalizer  # noqa: F401
from .rereference import ReReferencer  # noqa: F401
from .settings_models import FilterSettings, ProjectionSettings, ResamplerSettings  # noqa: F401
from ..filter.notch_filter import NotchFilter  # noqa: F401
