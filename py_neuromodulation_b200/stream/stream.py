"""Offline stream driver (reference: ``stream/stream.py``).

``Stream.run`` keeps the reference's signature and outputs (DataFrame, ``*_FEATURES.csv``,
``*_SETTINGS.yaml``, ``*_SIDECAR.json``, ``*_channels.csv``).  For an in-memory array -- the hot path of
this package -- the whole recording is uploaded once and ALL windows are processed by the fused GPU
pipeline in one call (``DataProcessor.process_windows``); the per-window Python loop of the reference
(stream.py:280-330) only remains for user-defined features, decoders and GUI back-ends.
LSL streaming is out of scope.
"""

from __future__ import annotations

import time
from pathlib import Path
from typing import TYPE_CHECKING, Any

import numpy as np

from ..utils.file_writer import FeatureTableWriter, MsgPackFileWriter
from ..utils.types import _PathLike
from .data_processor import DataProcessor
from .generator import RawDataGenerator, window_grid
from .settings import NMSettings

if TYPE_CHECKING:
    import pandas as pd

_USE_FREQ_RANGES = ["bandpass_filter", "stft", "fft", "welch", "bursts", "coherence", "nolds", "bispectrum"]


class Stream:
    def __init__(
        self,
        sfreq: float,
        channels: "pd.DataFrame | _PathLike | None" = None,
        data: "np.ndarray | pd.DataFrame | None" = None,
        settings: NMSettings | _PathLike | None = None,
        line_noise: float | None = 50,
        sampling_rate_features_hz: float | None = None,
        path_grids: _PathLike | None = None,
        coord_names: list | None = None,
        coord_list: list | None = None,
        verbose: bool = False,
    ) -> None:
        from ..utils import io
        from ..utils.channels import get_default_channels_from_data

        self.settings: NMSettings = NMSettings.load(settings)
        if channels is None and data is None:
            raise ValueError("Either `channels` or `data` must be passed to `Stream`.")
        if channels is None and data is not None:
            channels = get_default_channels_from_data(data)
        self.channels = io.load_channels(channels)
        if self.channels.query("used == 1 and target == 0").shape[0] == 0:
            raise ValueError(
                "No channels selected for analysis that have column 'used' = 1 and 'target' = 0. Please check your channels"
            )
        if any(f in _USE_FREQ_RANGES for f in self.settings.features.get_enabled()):
            assert all(fb.frequency_high_hz < sfreq / 2 for fb in self.settings.frequency_ranges_hz.values()), (
                "If a feature that uses frequency ranges is selected, the frequency band ranges need to be smaller than the "
                f"nyquist frequency.\nGot sfreq = {sfreq} and fband ranges:\n {self.settings.frequency_ranges_hz}"
            )
        if sampling_rate_features_hz is not None:
            self.settings.sampling_rate_features_hz = sampling_rate_features_hz

        self.path_grids = path_grids
        self.verbose = verbose
        self.sfreq = sfreq
        self.line_noise = line_noise
        self.coord_names = coord_names
        self.coord_list = coord_list
        self.sess_right = None
        self.projection = None
        self.model = None
        self.is_running = False
        self.data = data
        self._processor_key = None
        self.data_processor = self._new_processor()

    def _fingerprint(self):
        """Everything the processor is built from.  ``run`` rebuilds the processor from the CURRENT settings / channels like the
        reference (edits after ``__init__`` count) -- but only when they did change: filter design, the re-reference
        factorisation and the GPU plans of an unchanged configuration are kept (their state is reset)."""
        import json

        from .. import user_features

        ch = self.channels
        return (json.dumps(self.settings.model_dump(), sort_keys=True, default=str), tuple(ch.columns), repr(ch.to_numpy().tolist()),
                float(self.sfreq), self.line_noise, str(self.path_grids), repr(self.coord_names), repr(self.coord_list),
                bool(self.verbose), tuple(user_features.keys()))

    def _new_processor(self) -> DataProcessor:
        key = self._fingerprint()
        if self._processor_key == key and getattr(self, "data_processor", None) is not None:
            self.data_processor.reset_state()
            return self.data_processor
        dp = DataProcessor(
            sfreq=self.sfreq, settings=self.settings, channels=self.channels, path_grids=self.path_grids,
            coord_names=self.coord_names, coord_list=self.coord_list, line_noise=self.line_noise, verbose=self.verbose,
        )
        self._processor_key = key
        return dp

    # ------------------------------------------------------------------ helpers
    def _handle_data(self, data: "np.ndarray | pd.DataFrame") -> np.ndarray:
        names_expected = self.channels["name"].to_list()
        if isinstance(data, np.ndarray):
            if len(names_expected) != data.shape[0]:
                raise ValueError(
                    "If data is passed as an array, the first dimension must match the number of channel names in `channels`.\n"
                    f" Number of data channels (data.shape[0]): {data.shape[0]}\n"
                    f' Length of channels["name"]: {len(names_expected)}.'
                )
            return data
        names_data = data.columns.to_list()
        if not (len(names_expected) == len(names_data) and sorted(names_expected) == sorted(names_data)):
            raise ValueError(
                "If data is passed as a DataFrame, thecolumn names must match the channel names in `channels`.\n"
                f"Input dataframe column names: {names_data}\n"
                f'Expected (from channels["name"]): : {names_expected}.'
            )
        return data.to_numpy().transpose()

    def _targets(self) -> tuple[list[int], list[str]]:
        idx = [int(i) for i in self.channels[self.channels["target"] == 1].index]
        return idx, self.channels.loc[idx, "name"].to_list()

    def _add_target(self, feature_dict: dict, data: np.ndarray) -> None:
        for i, name in zip(*self._targets()):
            feature_dict[name] = data[i, -1]

    # ------------------------------------------------------------------ run
    def run(
        self,
        data: "np.ndarray | pd.DataFrame | None" = None,
        out_dir: _PathLike = "",
        experiment_name: str = "sub",
        is_stream_lsl: bool = False,
        stream_lsl_name: str | None = None,
        save_csv: bool = True,
        save_interval: int = 10,
        return_df: bool = True,
        simulate_real_time: bool = False,
        decoder: Any | None = None,
        backend_interface: Any | None = None,
        delete_ind_batch_files_after_stream: bool = True,
    ) -> "pd.DataFrame":
        self.is_stream_lsl = is_stream_lsl
        self.stream_lsl_name = stream_lsl_name
        self.save_csv = save_csv
        self.save_interval = save_interval
        self.return_df = return_df
        self.out_dir = Path.cwd() if not out_dir else Path(out_dir)
        self.experiment_name = experiment_name

        if is_stream_lsl:
            raise NotImplementedError("LSL streaming is outside the B200 hot path (SURVEY.md section 2 row 24)")
        if data is not None:
            data = self._handle_data(data)
        elif self.data is not None:
            data = self._handle_data(self.data)
        else:
            raise ValueError("No data passed to run function.")

        # like the reference, the processor is rebuilt from the CURRENT settings / channels (edits after __init__ count)
        self.data_processor = self._new_processor()
        self.batch_count = 0
        per_window = bool(self.data_processor.user_feature_names) or decoder is not None or backend_interface is not None
        if per_window:
            feature_df = self._run_per_window(data, out_dir, experiment_name, simulate_real_time, decoder, backend_interface,
                                              delete_ind_batch_files_after_stream)
        else:
            feature_df = self._run_batched(data, out_dir, experiment_name)
        self._save_after_stream()
        self.is_running = False
        return feature_df

    def _run_batched(self, data: np.ndarray, out_dir: _PathLike, experiment_name: str):
        dp = self.data_processor
        s = self.settings
        starts, lengths, times = window_grid(data.shape[1], self.sfreq, s.sampling_rate_features_hz, s.segment_length_features_ms)
        writer = FeatureTableWriter(name=experiment_name, out_dir=out_dir)
        if starts.size == 0:
            if self.save_csv:
                pass
            raise ValueError("No data to load")
        self.is_running = True
        distinct = np.unique(lengths)
        t_idx, t_names = self._targets()
        full = None
        if distinct.size == 1:
            # the GPU rows land directly in the left column block of the final table (row pitch = all columns): no second copy
            plan = dp.plan(int(distinct[0]))
            n_feat = len(plan.columns)
            full = np.empty((starts.size, n_feat + 1 + len(t_names)), dtype=np.float64)
            columns, matrix = dp.process_windows(data, starts, int(distinct[0]), out=full[:, :n_feat] if n_feat else None)
        else:
            # non-integer segment length / stride: two window lengths alternate.  Features are computed per length
            # without the normaliser, merged back in window order, and normalised in a second GPU pass.
            columns, matrix = None, None
            if dp.rawnorm_cfg is not None:
                raise NotImplementedError("raw normalisation needs a constant window length (stateful sample history)")
            for w in distinct:
                sel = np.flatnonzero(lengths == w)
                plan = dp.plan(int(w), with_normalizer=False, nan_reinsert=False)
                if plan.has_bursts:
                    raise NotImplementedError("burst features need a constant window length (stateful envelope history)")
                cols, part = dp.process_windows(data, starts[sel], int(w), with_normalizer=False, nan_reinsert=False)
                if matrix is None:
                    columns, matrix = cols, np.empty((starts.size, len(cols)))
                elif cols != columns:
                    raise ValueError("feature names differ between window lengths")
                matrix[sel] = part
            if dp.normalize and matrix.shape[1]:
                matrix = self._normalize_matrix(columns, matrix)
            # NaN re-insertion comes AFTER the normaliser (stream/data_processor.py:263-306): the history keeps the values
            # computed from the nan_to_num'ed samples
            self._reinsert_nan(data, starts, lengths, columns, matrix)
        self.batch_count = int(starts.size)

        all_cols = list(columns) + ["time"] + t_names
        if full is None:
            full = np.empty((starts.size, len(all_cols)), dtype=np.float64)
            full[:, : len(columns)] = matrix
        full[:, len(columns)] = times
        ends = starts + lengths - 1
        for j, ti in enumerate(t_idx):
            full[:, len(columns) + 1 + j] = np.asarray(data[ti, ends], dtype=np.float64)
        frame = writer.to_frame(all_cols, full)
        if self.save_csv:
            writer.save_csv(frame)
        return frame if self.return_df else {}

    def _reinsert_nan(self, data: np.ndarray, starts: np.ndarray, lengths: np.ndarray, columns: list[str], matrix: np.ndarray) -> None:
        """NaN for every feature of a channel that has a NaN inside the window (bookkeeping on the NaN mask only)."""
        names = self.data_processor.nan_names_by_raw_row
        rows = [r for r in range(data.shape[0]) if names[r] is not None and np.isnan(data[r]).any()]
        for r in rows:
            cs = np.concatenate(([0], np.cumsum(np.isnan(data[r]))))
            hit = (cs[starts + lengths] - cs[starts]) > 0
            if hit.any():
                cols = [i for i, k in enumerate(columns) if names[r] in k]
                matrix[np.ix_(np.flatnonzero(hit), cols)] = np.nan

    def _normalize_matrix(self, columns: list[str], matrix: np.ndarray) -> np.ndarray:
        """Rolling normalisation of a finished (n_windows, F) matrix: columns travel as 'channels', windows as time."""
        from .._pipeline import Pipeline, ScanSpec

        dp = self.data_processor
        ns = self.settings.feature_normalization_settings
        keep = [i for i, k in enumerate(columns) if ns.normalize_psd or "psd" not in k]
        if not keep:
            return matrix
        n = matrix.shape[0]
        names = [f"f{i}" for i in range(len(keep))]
        cols = [f"{c}_raw" for c in names]
        pipe = Pipeline(len(keep), len(keep), 3, cols, device=dp.device)
        ScanSpec(names, raw=True).attach(pipe)
        pipe.add_feature_normalizer(ns.normalization_method, ns.clip, dp.norm_keep, cols)
        pipe.finalize()
        rec = np.zeros((len(keep), n + 2))
        rec[:, 2:] = matrix[:, keep].T
        nan_mask = np.isnan(rec)
        pipe.upload(rec)
        normed = pipe.run(np.arange(n))
        out = matrix.copy()
        out[:, keep] = normed
        out[:, keep] = np.where(nan_mask[:, 2:].T, np.nan, out[:, keep])  # NaN channels stay NaN (re-inserted after normalisation)
        pipe.close()
        return out

    def _run_per_window(self, data, out_dir, experiment_name, simulate_real_time, decoder, backend_interface, delete_files):
        from .. import logger

        file_writer = MsgPackFileWriter(name=experiment_name, out_dir=out_dir)
        generator = RawDataGenerator(data, self.sfreq, self.settings.sampling_rate_features_hz, self.settings.segment_length_features_ms)
        for timestamps, data_batch in generator:
            self.is_running = True
            if backend_interface:
                if simulate_real_time:
                    time.sleep(1 / self.settings.sampling_rate_features_hz)
                if backend_interface.check_control_signals() == "stop":
                    break
            if data_batch is None:
                break
            logger.debug("Processing new data batch")
            feature_dict = self.data_processor.process(data_batch)
            if decoder is not None:
                ch_to_decode = self.channels.query("used == 1").iloc[0]["name"]
                feature_dict = decoder.predict(feature_dict, ch_to_decode, fft_bands_only=True)
            feature_dict["time"] = np.ceil(timestamps[-1] * 1000 + 1)
            self._add_target(feature_dict, data_batch)
            file_writer.insert_data(feature_dict)
            if backend_interface:
                backend_interface.send_features(feature_dict)
                backend_interface.send_raw_data(self._prepare_raw_data_dict(data_batch))
            self.batch_count += 1
            if self.batch_count % self.save_interval == 0:
                file_writer.save()
        file_writer.save()
        if self.save_csv:
            file_writer.save_as_csv(save_all_combined=True)
        feature_df = file_writer.load_all() if self.return_df else {}
        if delete_files:
            file_writer.delete_ind_files()
        return feature_df

    def _prepare_raw_data_dict(self, data_batch: np.ndarray) -> dict[str, Any]:
        new_samples = int(1000 / self.settings.sampling_rate_features_hz * self.sfreq / 1000)
        return {"raw_data": {ch: list(data_batch[i, -new_samples:]) for i, ch in enumerate(self.channels["name"])}}

    # ------------------------------------------------------------------ side files
    def _save_after_stream(self) -> None:
        self._save_sidecar()
        self._save_settings()
        self._save_channels()

    def _save_features(self, feature_arr: "pd.DataFrame") -> None:
        from ..utils import io

        io.save_features(feature_arr, self.out_dir, self.experiment_name)

    def _save_channels(self) -> None:
        self.data_processor.save_channels(self.out_dir, self.experiment_name)

    def _save_settings(self) -> None:
        self.data_processor.save_settings(self.out_dir, self.experiment_name)

    def _save_sidecar(self) -> None:
        self.data_processor.save_sidecar(self.out_dir, self.experiment_name, {"sess_right": self.sess_right})
