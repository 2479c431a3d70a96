"""``NMSettings`` -- the configuration surface of the hot path (reference: ``stream/settings.py``)."""

from __future__ import annotations

from pathlib import Path
from typing import Any, ClassVar, get_args

from pydantic import ValidationError, model_validator
from pydantic.functional_validators import ModelWrapValidatorHandler

from ..features.bandpower import BandPowerSettings
from ..features.bursts import BurstsSettings
from ..features.oscillatory import OscillatorySettings
from ..features.out_of_scope import (
    BispectraSettings,
    CoherenceSettings,
    FooofSettings,
    MNEConnectivitySettings,
    NoldsSettings,
)
from ..features.sharpwaves import SharpwaveSettings
from ..filter.kalman_settings import KalmanSettings
from ..processing.normalization import FeatureNormalizationSettings, NormalizationSettings
from ..processing.settings_models import FilterSettings, ProjectionSettings, ResamplerSettings
from ..utils.pydantic_extensions import NMErrorList, NMField
from ..utils.types import NORM_METHOD, PREPROCESSOR_NAME, BoolSelector, FrequencyRange, NMBaseModel, _PathLike

DEFAULT_SETTINGS_FILE = Path(__file__).resolve().parents[1] / "default_settings.yaml"


class FeatureSelector(BoolSelector):
    raw_hjorth: bool = True
    return_raw: bool = True
    bandpass_filter: bool = False
    stft: bool = False
    fft: bool = True
    welch: bool = True
    sharpwave_analysis: bool = True
    fooof: bool = False
    nolds: bool = False
    coherence: bool = False
    bursts: bool = True
    linelength: bool = True
    mne_connectivity: bool = False
    bispectrum: bool = False


class PostprocessingSettings(BoolSelector):
    feature_normalization: bool = True
    project_cortex: bool = False
    project_subcortex: bool = False


DEFAULT_PREPROCESSORS: list[PREPROCESSOR_NAME] = ["raw_resampling", "notch_filter", "re_referencing"]


def _strip_gui_metadata(data):
    """Settings may arrive from the GUI wrapped as {"__value__": ...} with "__unit__"-style siblings."""
    if isinstance(data, dict):
        if "__value__" in data:
            return data["__value__"]
        return {k: _strip_gui_metadata(v) for k, v in data.items() if not k.startswith("__")}
    if isinstance(data, (list, tuple, set)):
        return type(data)(_strip_gui_metadata(v) for v in data)
    return data


class NMSettings(NMBaseModel):
    _instances: ClassVar[list["NMSettings"]] = []

    sampling_rate_features_hz: float = NMField(default=10, gt=0, custom_metadata={"unit": "Hz"})
    segment_length_features_ms: float = NMField(default=1000, gt=0, custom_metadata={"unit": "ms"})
    frequency_ranges_hz: dict[str, FrequencyRange] = {
        "theta": FrequencyRange(4, 8),
        "alpha": FrequencyRange(8, 12),
        "low_beta": FrequencyRange(13, 20),
        "high_beta": FrequencyRange(20, 35),
        "low_gamma": FrequencyRange(60, 80),
        "high_gamma": FrequencyRange(90, 200),
        "HFA": FrequencyRange(200, 400),
    }

    preprocessing: list[PREPROCESSOR_NAME] = NMField(
        default=DEFAULT_PREPROCESSORS,
        custom_metadata={"field_type": "PreprocessorList", "valid_values": list(get_args(PREPROCESSOR_NAME))},
    )
    raw_resampling_settings: ResamplerSettings = ResamplerSettings()
    preprocessing_filter: FilterSettings = FilterSettings()
    raw_normalization_settings: NormalizationSettings = NormalizationSettings()

    postprocessing: PostprocessingSettings = PostprocessingSettings()
    feature_normalization_settings: FeatureNormalizationSettings = FeatureNormalizationSettings()
    project_cortex_settings: ProjectionSettings = ProjectionSettings(max_dist_mm=20)
    project_subcortex_settings: ProjectionSettings = ProjectionSettings(max_dist_mm=5)

    features: FeatureSelector = FeatureSelector()

    fft_settings: OscillatorySettings = OscillatorySettings()
    welch_settings: OscillatorySettings = OscillatorySettings()
    stft_settings: OscillatorySettings = OscillatorySettings()
    bandpass_filter_settings: BandPowerSettings = BandPowerSettings()
    kalman_filter_settings: KalmanSettings = KalmanSettings()
    bursts_settings: BurstsSettings = BurstsSettings()
    sharpwave_analysis_settings: SharpwaveSettings = SharpwaveSettings()
    mne_connectivity_settings: MNEConnectivitySettings = MNEConnectivitySettings()
    coherence_settings: CoherenceSettings = CoherenceSettings()
    fooof_settings: FooofSettings = FooofSettings()
    nolds_features: NoldsSettings = NoldsSettings()
    bispectrum_settings: BispectraSettings = BispectraSettings()

    def __init__(self, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        from .. import user_features

        for name in user_features.keys():
            setattr(self.features, name, True)
        NMSettings._instances.append(self)

    @classmethod
    def _add_feature(cls, feature: str) -> None:
        for inst in cls._instances:
            setattr(inst.features, feature, True)

    @classmethod
    def _remove_feature(cls, feature: str) -> None:
        for inst in cls._instances:
            try:
                delattr(inst.features, feature)
            except AttributeError:
                pass

    @model_validator(mode="wrap")  # type: ignore[arg-type]
    def _validate_settings(self, handler: ModelWrapValidatorHandler) -> Any:
        errors = NMErrorList()
        self = _strip_gui_metadata(self)
        try:
            self = handler(self)
        except ValidationError as exc:
            self = NMSettings.unvalidated(**self)  # keep going to collect the cross-field errors too
            errors.extend(NMErrorList(exc.errors()))

        if len(self.features.get_enabled()) == 0:
            errors.add_error("At least one feature must be selected.")

        self.frequency_ranges_hz = {k.replace(" ", "_"): v for k, v in self.frequency_ranges_hz.items()}

        if self.features.bandpass_filter:
            errors.extend(self.bandpass_filter_settings.validate_fbands(self))
            if self.bandpass_filter_settings.kalman_filter:
                errors.extend(self.kalman_filter_settings.validate_fbands(self))

        if len(errors) > 0:
            raise errors.create_error()
        return self

    # -- presets
    def reset(self) -> "NMSettings":
        self.features.disable_all()
        self.preprocessing = list(DEFAULT_PREPROCESSORS)
        self.postprocessing.disable_all()
        return self

    def set_fast_compute(self) -> "NMSettings":
        self.reset()
        self.features.fft = True
        self.preprocessing = list(DEFAULT_PREPROCESSORS)
        self.postprocessing.feature_normalization = True
        self.postprocessing.project_cortex = False
        self.postprocessing.project_subcortex = False
        return self

    def enable_all_features(self) -> "NMSettings":
        self.features.enable_all()
        return self

    def disable_all_features(self) -> "NMSettings":
        self.features.disable_all()
        return self

    @staticmethod
    def get_fast_compute() -> "NMSettings":
        return NMSettings.get_default().set_fast_compute()

    # -- loading / saving
    @classmethod
    def load(cls, settings: "NMSettings | _PathLike | None") -> "NMSettings":
        if isinstance(settings, cls):
            return settings.validate()
        if settings is None:
            return cls.get_default()
        return cls.from_file(str(settings))

    @staticmethod
    def from_file(PATH: _PathLike) -> "NMSettings":
        path = Path(PATH)
        if path.is_dir():
            for child in path.iterdir():
                if child.is_file() and child.suffix in (".json", ".yaml"):
                    path = child
                    break
        if not path.is_dir() and not path.is_file():
            for child in path.parent.iterdir():
                ext = child.suffix.lower()
                if child.is_file() and ext in (".json", ".yaml") and child.name == path.stem + "_SETTINGS" + ext:
                    path = child
                    break
        if path.suffix == ".json":
            import json

            with open(path) as f:
                model_dict = json.load(f)
        elif path.suffix == ".yaml":
            import yaml

            with open(path) as f:
                model_dict = yaml.safe_load(f)
        else:
            raise ValueError("File format not supported.")
        return NMSettings(**model_dict)

    @staticmethod
    def get_default() -> "NMSettings":
        return NMSettings.from_file(DEFAULT_SETTINGS_FILE)

    @staticmethod
    def list_normalization_methods() -> list[str]:
        return list(get_args(NORM_METHOD))

    def save(self, out_dir: _PathLike = ".", prefix: str = "", format: str = "yaml") -> None:
        filename = f"{prefix}_SETTINGS.{format}" if prefix else f"SETTINGS.{format}"
        path_out = Path(out_dir) / prefix / filename
        with open(path_out, "w") as f:
            if format == "json":
                f.write(self.model_dump_json(indent=4))
            elif format == "yaml":
                import yaml

                # same text as yaml.dump(...): the libyaml emitter when PyYAML was built with it (10x faster)
                yaml.dump(self.model_dump(), f, default_flow_style=None, Dumper=getattr(yaml, "CDumper", yaml.Dumper))
            else:
                raise ValueError("File format not supported.")


def get_default_settings() -> NMSettings:
    return NMSettings.get_default()


def reset_settings(settings: NMSettings) -> NMSettings:
    return settings.reset()


def get_fast_compute() -> NMSettings:
    return NMSettings.get_fast_compute()


def test_settings(settings: NMSettings) -> NMSettings:
    return settings.validate()
