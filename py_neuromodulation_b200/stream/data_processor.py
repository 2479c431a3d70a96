"""Window processor (reference: ``stream/data_processor.py``).

``DataProcessor.process(window)`` keeps the reference's one-window interface; ``process_windows`` is the
batched entry used by ``Stream.run`` for offline arrays.  Both drive ONE fused GPU pipeline per window
length: nan_to_num -> pick -> re-reference -> notch -> all enabled built-in features -> rolling
normalisation -> NaN re-insertion.  Feature order = ``FeatureSelector`` field order (the reference's dict
insertion order), then user-defined features.
"""

from __future__ import annotations

from time import time
from typing import TYPE_CHECKING

import numpy as np

from ..utils import io
from ..utils.types import _PathLike
from .settings import NMSettings

if TYPE_CHECKING:
    import pandas as pd

_SCAN = {"raw_hjorth": "hjorth", "return_raw": "raw", "linelength": "linelength"}


def build_specs(settings, names, sfreq, window_samples: int, user_feature_names=()):
    """Feature-family specs + output columns in the reference's dict-insertion order.  No GPU involved."""
    from .._pipeline import BandpowerSpec, BurstsSpec, ScanSpec, SharpwaveSpec, SpectralSpec, band_items
    from ..features.bursts import check_burst_bands
    from ..features.feature_processor import OUT_OF_SCOPE

    s, fs = settings, sfreq
    enabled = [f for f in s.features.get_enabled() if f not in user_feature_names]
    for f in enabled:
        if f in OUT_OF_SCOPE:
            raise NotImplementedError(f"feature '{f}' is outside the B200 hot path (SURVEY.md section 2 row 23)")
    scan = ScanSpec(names, hjorth="raw_hjorth" in enabled, raw="return_raw" in enabled, linelength="linelength" in enabled)
    specs: list = []
    columns: list[str] = []
    for f in enabled:
        if f == "raw_hjorth":
            columns += scan.keys_hjorth()
        elif f == "return_raw":
            columns += scan.keys_raw()
        elif f == "linelength":
            columns += scan.keys_linelength()
        else:
            if f in ("fft", "welch", "stft"):
                assert getattr(s, f"{f}_settings").windowlength_ms <= s.segment_length_features_ms
                spec = SpectralSpec(f, getattr(s, f"{f}_settings"), band_items(s), names, fs, window_samples)
            elif f == "bandpass_filter":
                spec = BandpowerSpec(s.bandpass_filter_settings, band_items(s), names, fs)
            elif f == "bursts":
                check_burst_bands(s)
                spec = BurstsSpec(s, names, fs)
            elif f == "sharpwave_analysis":
                spec = SharpwaveSpec(s, names, fs)
            else:  # pragma: no cover
                raise NotImplementedError(f)
            specs.append(spec)
            columns += spec.keys()
    return scan, specs, columns, enabled


class _Plan:
    """Everything that depends on the window length."""

    def __init__(self, dp: "DataProcessor", window_samples: int, with_normalizer: bool, nan_reinsert: bool = True) -> None:
        from .._pipeline import Pipeline

        s, names = dp.settings, dp.ch_names_used_features
        # behind the resampler the families see round(W * ratio) samples per window -- designed with the ORIGINAL sampling rate
        # (the reference never updates sfreq_raw: stream/data_processor.py:55,77-81)
        raw_window = int(window_samples)
        window_samples = dp.feature_window(raw_window)
        scan, specs, columns, enabled = build_specs(s, names, dp.sfreq_raw, window_samples, dp.user_feature_names)
        self.columns = columns
        self.has_bursts = "bursts" in enabled
        self.bool_columns = [i for i, k in enumerate(columns) if k.endswith("_in_burst")] if self.has_bursts else []
        self.normalized = False
        if not columns:
            self.pipe = None
            return
        pipe = Pipeline(dp.n_raw_rows, len(names), window_samples, columns, device=dp.device)
        pipe.set_precision(dp.precision)
        pipe.set_fused(-1 if dp.fused is None else (2 if dp.fused == "front" else int(bool(dp.fused))))
        pipe.set_pick(dp.feature_idx)
        if dp.reref_factored is not None:  # channel-sharded run: coefficients come from the GLOBAL channel table
            pipe.set_reref_factored(*dp.reref_factored)
        elif "re_referencing" in dp.preproc_plan:
            pipe.set_reref(dp.ref_matrix)
        dp.attach_resampler(pipe, raw_window)
        pipe.set_prefilters(dp.prefilter_taps)
        if "notch_filter" in dp.preproc_plan:
            pipe.set_notch(dp.notch_taps)
        if dp.rawnorm_cfg is not None:
            pipe.set_raw_normalizer(*dp.rawnorm_cfg)
        if scan.hjorth or scan.raw or scan.linelength:
            scan.attach(pipe)
        for spec in specs:
            spec.attach(pipe)
        if with_normalizer and dp.normalize:
            ns = s.feature_normalization_settings
            cols = columns if ns.normalize_psd else [k for k in columns if "psd" not in k]
            pipe.add_feature_normalizer(ns.normalization_method, ns.clip, dp.norm_keep, cols)
            self.normalized = True
        if nan_reinsert:
            pipe.set_nan_columns(dp.nan_names_by_raw_row)
        pipe.finalize()
        self.pipe = pipe


class DataProcessor:
    def __init__(
        self,
        sfreq: float,
        settings: NMSettings | _PathLike,
        channels: "pd.DataFrame | _PathLike",
        coord_names: list | None = None,
        coord_list: list | None = None,
        line_noise: float | None = None,
        path_grids: _PathLike | None = None,
        verbose: bool = True,
        device: int = 0,
        reref_factored: tuple | None = None,
        precision: str | None = None,
        fused: "bool | str | None" = None,
    ) -> None:
        from .. import user_features
        from ..filter.notch_filter import NotchFilter
        from ..processing.data_preprocessor import preprocessing_plan
        from ..processing.normalization import check_feature_norm_method
        from ..processing.rereference import build_reference_matrix

        # arithmetic of the linear FIR families: "f64" (default; ~1e-12 from the float64 reference) or "f32" (float32 inside the
        # FFT convolution, 1e-5 relative); env NMB200_PRECISION overrides the default for callers that cannot pass the argument
        import os

        # kernel organisation: None = library default / NMB200_FUSED, True = one persistent kernel per (window, channel pair)
        # (csrc/nm_fused.cuh), "front" = notch + scan + segment DFT in one kernel, band-pass bank separate (the default organisation),
        # False = one kernel per stage
        self.fused = fused
        self.precision = precision or os.environ.get("NMB200_PRECISION", "f64")
        if self.precision not in ("f64", "f32", "f32x2"):
            raise ValueError("precision must be 'f64', 'f32' or 'f32x2'")
        self.settings = NMSettings.load(settings)
        self.channels = io.load_channels(channels)
        self.sfreq_features: float = self.settings.sampling_rate_features_hz
        self._sfreq_raw_orig: float = sfreq
        self.sfreq_raw: float = sfreq // 1
        self.line_noise = line_noise
        self.path_grids = path_grids
        self.verbose = verbose
        self.device = device
        self.reref_factored = reref_factored
        self.projection = None
        if self.settings.postprocessing.project_cortex or self.settings.postprocessing.project_subcortex:
            raise NotImplementedError("grid-point projection is outside the B200 hot path (SURVEY.md section 2 row 21)")

        ch = self.channels
        good_used = (ch["used"] == 1) & (ch["status"] == "good")
        self.ch_names_used: list[str] = ch.loc[good_used, "new_name"].tolist()
        status_l, new_names_l = ch["status"].tolist(), ch["new_name"].tolist()  # (plain lists: no per-row pandas look-ups)
        self.feature_idx: list[int] = [
            int(i) for i in np.where(ch["used"].astype(bool) & ~ch["target"].astype(bool))[0] if status_l[i] == "good"
        ]
        self.n_raw_rows = int(ch.shape[0])
        # names of the rows that are actually processed (identical to ch_names_used unless a used channel is a target)
        self.ch_names_used_features = [new_names_l[i] for i in self.feature_idx]
        if len(self.ch_names_used) == len(self.ch_names_used_features):
            self.ch_names_used_features = list(self.ch_names_used)
        # NaN re-insertion (stream/data_processor.py:297-306): a NaN in raw row r turns every feature whose key contains that
        # row's channel name into NaN.  The reference indexes ch_names_used with a mask over ALL raw rows, which is only
        # defined when every row is used and good (it raises IndexError otherwise); here every used, good row maps to its OWN
        # channel name and unused / bad rows to nothing -- identical whenever the reference is defined.
        good_used_l = good_used.tolist()
        self.nan_names_by_raw_row = [new_names_l[r] if good_used_l[r] else None for r in range(self.n_raw_rows)]

        self.preproc_plan = preprocessing_plan(self.settings, self.sfreq_raw)
        self.notch_taps = None
        if "notch_filter" in self.preproc_plan:
            self.notch_taps = NotchFilter(self.sfreq_raw, line_noise).filter_bank
            if self.notch_taps is None:
                self.preproc_plan.remove("notch_filter")
        self.prefilter_taps = None
        if "preprocessing_filter" in self.preproc_plan:
            from ..processing.filter_preprocessing import PreprocessingFilter

            self.prefilter_taps = PreprocessingFilter(self.settings, self.sfreq_raw).stage_taps()
            if self.prefilter_taps is None:
                self.preproc_plan.remove("preprocessing_filter")
        self.rawnorm_cfg = None
        if "raw_normalization" in self.preproc_plan:
            rs = self.settings.raw_normalization_settings.validate()
            self.rawnorm_cfg = (rs.normalization_method, rs.clip, int(rs.normalization_time_s * self.sfreq_raw),
                                int(self.sfreq_raw / self.settings.sampling_rate_features_hz))
        self.resample_ratio = None
        if "raw_resampling" in self.preproc_plan:
            self.resample_ratio = float(self.settings.raw_resampling_settings.resample_freq_hz / self.sfreq_raw)
        self.ref_matrix = None
        if "re_referencing" in self.preproc_plan:
            self.ref_matrix = build_reference_matrix(ch)
            if self.ref_matrix is None:
                self.preproc_plan.remove("re_referencing")

        self.normalize = bool(self.settings.postprocessing.feature_normalization)
        if self.normalize:
            ns = self.settings.feature_normalization_settings.validate()
            self.norm_keep = int(ns.normalization_time_s * self.settings.sampling_rate_features_hz)
            check_feature_norm_method(ns.normalization_method, self.norm_keep)

        self.user_feature_names = list(user_features.keys())
        self._user_plugins = {name: cls(self.settings, self.ch_names_used_features, self.sfreq_raw) for name, cls in user_features.items()}
        self._user_norm = None
        self._plans: dict[tuple, _Plan] = {}
        self._stream_window: int | None = None
        self.cnt_samples = 0
        # validate the plug-in settings now (bands, filters, estimators) like the reference constructor does
        self._probe = _ProbeOnly(self)

    # ------------------------------------------------------------------ resampler
    def feature_window(self, raw_window: int) -> int:
        """Samples per window the feature families see (``mne.filter.resample``: ``round(ratio * n)``)."""
        if self.resample_ratio is None:
            return int(raw_window)
        from ..processing.resample import resample_geometry

        return int(resample_geometry(int(raw_window), self.resample_ratio)["final_len"])

    def attach_resampler(self, pipe, raw_window: int) -> None:
        if self.resample_ratio is not None:
            from ..processing.resample import integer_decimation, resample_operator

            pipe.set_resampler(resample_operator(int(raw_window), self.resample_ratio), integer_decimation(self.resample_ratio))

    # ------------------------------------------------------------------ plans
    def plan(self, window_samples: int, with_normalizer: bool = True, nan_reinsert: bool = True) -> _Plan:
        key = (int(window_samples), bool(with_normalizer), bool(nan_reinsert))
        if key not in self._plans:
            self._plans[key] = _Plan(self, int(window_samples), with_normalizer, nan_reinsert)
        return self._plans[key]

    @property
    def stateful(self) -> bool:
        """Stages that carry state from window to window (one history per window length would be wrong)."""
        return bool(self.normalize or self.rawnorm_cfg is not None or "bursts" in self.settings.features.get_enabled())

    def reset_state(self) -> None:
        for p in self._plans.values():
            if p.pipe is not None:
                p.pipe.reset_state()
        self._user_norm = None
        self._stream_window = None

    # ------------------------------------------------------------------ one window (reference interface)
    def process(self, data: np.ndarray) -> dict[str, float]:
        start_time = time()
        data = np.asarray(data)
        if self._stream_window is None:
            self._stream_window = int(data.shape[1])
        elif self._stream_window != int(data.shape[1]) and self.stateful:
            # the reference keeps ONE normaliser / burst history across window lengths; a pipeline here is built per length
            raise NotImplementedError(
                f"window length changed from {self._stream_window} to {data.shape[1]} samples while a stateful stage (feature / raw "
                "normalisation, bursts) is enabled: use an integer segment length / stride or the batched Stream.run path"
            )
        plan = self.plan(data.shape[1])
        feats: dict = {}
        if plan.pipe is not None:
            values = plan.pipe.process_window(data.astype(np.float64, copy=False))
            feats = dict(zip(plan.columns, values))
            if not plan.normalized:
                for i in plan.bool_columns:
                    feats[plan.columns[i]] = bool(values[i])
        if self._user_plugins:
            feats.update(self._user_features(plan, data))
        if self.verbose:
            from .. import logger

            logger.info("Last batch took: %.3f seconds to process", time() - start_time)
        return feats

    def _user_features(self, plan: _Plan, data: np.ndarray) -> dict:
        """User-defined Python plugins see the same preprocessed window the GPU families see."""
        from .._pipeline import IdentityNormPipeline
        from ..processing.normalization import GPU_FEATURE_NORM_METHODS

        # a pipeline of its own: the feature pipeline has already advanced its (stateful) raw normaliser for this window, and
        # preprocessing the window a second time on it would advance it twice
        pre_pipe = self._preprocess_only(data.shape[1])
        pre = pre_pipe.preprocess_window(data.astype(np.float64, copy=False))
        out: dict = {}
        for plugin in self._user_plugins.values():
            out.update(plugin.calc_feature(pre))
        if self.normalize and out:
            ns = self.settings.feature_normalization_settings
            keys = [k for k in out if ns.normalize_psd or "psd" not in k]
            if keys:
                if self._user_norm is None:
                    self._user_norm = IdentityNormPipeline(len(keys), GPU_FEATURE_NORM_METHODS.index(ns.normalization_method),
                                                           float(ns.clip or 0.0), self.norm_keep, device=self.device)
                normed = self._user_norm.step(np.array([float(out[k]) for k in keys]))
                out.update(zip(keys, normed))
        nan_rows = np.isnan(data).any(axis=1)
        if nan_rows.any():
            for r in np.flatnonzero(nan_rows):
                name = self.nan_names_by_raw_row[r]
                if name is not None:
                    for k in out:
                        if name in k:
                            out[k] = np.nan
        return out

    def _preprocess_only(self, window_samples: int):
        from .._pipeline import Pipeline, ScanSpec

        key = ("pre", window_samples)
        if key not in self._plans:
            names = self.ch_names_used_features
            pipe = Pipeline(self.n_raw_rows, len(names), self.feature_window(window_samples), ["_unused"], device=self.device)
            pipe.set_pick(self.feature_idx)
            if "re_referencing" in self.preproc_plan:
                pipe.set_reref(self.ref_matrix)
            self.attach_resampler(pipe, window_samples)
            pipe.set_prefilters(self.prefilter_taps)
            if "notch_filter" in self.preproc_plan:
                pipe.set_notch(self.notch_taps)
            if self.rawnorm_cfg is not None:
                pipe.set_raw_normalizer(*self.rawnorm_cfg)
            ScanSpec(names).attach(pipe)
            pipe.finalize()
            holder = _Plan.__new__(_Plan)
            holder.pipe, holder.columns, holder.bool_columns, holder.normalized, holder.has_bursts = pipe, [], [], False, False
            self._plans[key] = holder
        return self._plans[key].pipe

    # ------------------------------------------------------------------ batched offline entry
    def process_windows(self, data: np.ndarray, starts: np.ndarray, window_samples: int, with_normalizer: bool = True,
                        upload: bool = True, out: np.ndarray | None = None, nan_reinsert: bool = True):
        """All windows ``[starts[k], starts[k] + W)`` of a resident recording -> ``(columns, (n, F) float64)``."""
        plan = self.plan(window_samples, with_normalizer, nan_reinsert)
        if plan.pipe is None:
            return [], np.empty((len(starts), 0))
        if upload:
            plan.pipe.upload(data)
        return plan.columns, plan.pipe.run(starts, out=out)

    # ------------------------------------------------------------------ side files
    def save_sidecar(self, out_dir: _PathLike, prefix: str = "", additional_args: dict | None = None) -> None:
        sidecar: dict = {"original_fs": self._sfreq_raw_orig, "final_fs": self.sfreq_raw, "sfreq": self.sfreq_features}
        if additional_args is not None:
            sidecar = sidecar | additional_args
        io.save_sidecar(sidecar, out_dir, prefix)

    def save_settings(self, out_dir: _PathLike, prefix: str = "") -> None:
        self.settings.save(out_dir, prefix)

    def save_channels(self, out_dir: _PathLike, prefix: str) -> None:
        io.save_channels(self.channels, out_dir, prefix)

    def save_features(self, feature_arr: "pd.DataFrame", out_dir: _PathLike = "", prefix: str = "") -> None:
        io.save_features(feature_arr, out_dir, prefix)


class _ProbeOnly:
    """Runs the settings checks the reference plug-in constructors perform, without touching the GPU."""

    def __init__(self, dp: DataProcessor) -> None:
        from ..features.bursts import check_burst_bands
        from ..features.feature_processor import OUT_OF_SCOPE

        s = dp.settings
        enabled = s.features.get_enabled()
        for f in enabled:
            if f in OUT_OF_SCOPE:
                raise NotImplementedError(f"feature '{f}' is outside the B200 hot path (SURVEY.md section 2 row 23)")
        for f in ("fft", "welch", "stft"):
            if f in enabled:
                assert getattr(s, f"{f}_settings").windowlength_ms <= s.segment_length_features_ms, (
                    f"oscillatory feature windowlength_ms = ({getattr(s, f'{f}_settings').windowlength_ms})needs to be smaller than"
                    f"settings['segment_length_features_ms'] = {s.segment_length_features_ms}"
                )
        if "bursts" in enabled:
            check_burst_bands(s)
        if "sharpwave_analysis" in enabled:
            for fr in s.sharpwave_analysis_settings.filter_ranges_hz:
                assert fr[1] < dp.sfreq_raw, f"Filter range has to be smaller than sfreq, got sfreq {dp.sfreq_raw} and filter range {fr}"
