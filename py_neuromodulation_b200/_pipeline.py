"""Host-side construction of GPU pipelines.

A *spec* object describes one feature family for a given channel list / sampling rate / window
length: it knows the output keys in the reference's dict-insertion order and how to register
itself on a :class:`Pipeline` (which wraps one ``nm_pipeline`` handle of the C ABI).  The plugin
classes in :mod:`py_neuromodulation_b200.features` and the window processor in
:mod:`py_neuromodulation_b200.stream.data_processor` are thin layers over these specs.

Nothing here computes features: all arithmetic on samples happens in ``libnmb200.so``.
"""

from __future__ import annotations

import ctypes as C
import math
from collections.abc import Sequence

import numpy as np

from . import _lib

OSC_ESTIMATORS = ("mean", "median", "std", "max")
SW_FEATURES = ("peak_left", "peak_right", "num_peaks", "trough", "width", "prominence", "interval", "decay_time",
               "rise_time", "sharpness", "rise_steepness", "decay_steepness", "slope_ratio")
SW_ESTIMATORS = ("mean", "median", "max", "min", "var")
NORM_METHODS = ("mean", "median", "zscore", "zscore-median")  # raw normaliser (csrc/nm_rawnorm.cuh)
# feature normaliser (csrc/nm_norm.cuh): the numpy methods + the scikit-learn transformers the reference wraps, restated on the GPU
# ("power" -- Yeo-Johnson with a per-window maximum-likelihood search -- stays out of scope)
FEATURE_NORM_METHODS = NORM_METHODS + ("minmax", "robust", "quantile")
# raw normaliser: MinMaxScaler / RobustScaler on the sliding order-statistic kernel; the reference's QuantileTransformer draws a RANDOM
# subsample of 10 000 of the 30 000 history samples per window (no reproducible answer), PowerTransformer as above
RAW_NORM_METHODS = NORM_METHODS + ("minmax", "robust")


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a: np.ndarray, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


# ------------------------------------------------------------------------------------- pipeline
class Pipeline:
    """One ``nm_pipeline``: fixed raw-row count, feature-channel count, window length and column list."""

    def __init__(self, n_raw_rows: int, n_ch: int, window_samples: int, columns: Sequence[str], device: int = 0) -> None:
        self.lib = _lib.load()
        self.columns = list(columns)
        self.col_of = {k: i for i, k in enumerate(self.columns)}
        if len(self.col_of) != len(self.columns):
            raise ValueError("duplicate feature names in the column list")
        self.n_raw_rows, self.n_ch, self.W, self.F = int(n_raw_rows), int(n_ch), int(window_samples), len(self.columns)
        self.W_in = self.W  # raw samples per window (differs from W behind a resampler)
        handle = C.c_void_p()
        _lib.check(self.lib.nm_pipeline_create(int(device), self.n_raw_rows, self.n_ch, self.W, self.F, C.byref(handle)))
        self._h = handle
        self._keep: list = []  # arrays whose memory must outlive the C calls
        self.finalized = False

    # -- life cycle
    def close(self) -> None:
        if getattr(self, "_h", None):
            self.lib.nm_pipeline_destroy(self._h)
            self._h = None

    def __del__(self) -> None:  # pragma: no cover - interpreter shutdown order
        try:
            self.close()
        except Exception:
            pass

    def finalize(self) -> "Pipeline":
        _lib.check(self.lib.nm_finalize(self._h))
        self.finalized = True
        return self

    def reset_state(self) -> None:
        _lib.check(self.lib.nm_reset_state(self._h))

    # -- preprocessing
    def set_pick(self, pick: Sequence[int]) -> None:
        a = _i32(pick)
        assert a.shape == (self.n_ch,)
        _lib.check(self.lib.nm_set_pick(self._h, _ptr(a, C.c_int)))

    def set_reref(self, ref_matrix: np.ndarray | None, min_group: int = 4) -> None:
        """Factor the dense (C, C) reference matrix into group sums + sparse remainder (see csrc/nm_prep.cuh)."""
        if ref_matrix is None:
            return
        m = _f64(ref_matrix)
        if m.shape != (self.n_ch, self.n_ch):
            raise ValueError(
                f"re-reference matrix is {m.shape} but {self.n_ch} channels are processed: the channel table selects "
                "different rows for re-referencing (used == 1) and for features (used, not target, good)"
            )
        groups, group_of, gcoef, rem = factor_reference_matrix(m, min_group=min_group)
        ptr = [0]
        cols: list[int] = []
        vals: list[float] = []
        for i in range(self.n_ch):
            nz = np.flatnonzero(rem[i])
            cols.extend(int(j) for j in nz)
            vals.extend(float(rem[i, j]) for j in nz)
            ptr.append(len(cols))
        a_go, a_gc = _i32(group_of), _f64(gcoef)
        a_ptr, a_col, a_val = _i32(ptr), _i32(cols if cols else [0]), _f64(vals if vals else [0.0])
        _lib.check(self.lib.nm_set_reref(self._h, groups, _ptr(a_go, C.c_int), _ptr(a_gc, C.c_double), _ptr(a_ptr, C.c_int),
                                         _ptr(a_col, C.c_int), _ptr(a_val, C.c_double)))

    def set_reref_factored(self, n_groups: int, group_of, gcoef, sp_ptr, sp_col, sp_val) -> None:
        a_go, a_gc = _i32(group_of), _f64(gcoef)
        a_ptr, a_col, a_val = _i32(sp_ptr), _i32(sp_col), _f64(sp_val)
        _lib.check(self.lib.nm_set_reref(self._h, int(n_groups), _ptr(a_go, C.c_int), _ptr(a_gc, C.c_double),
                                         _ptr(a_ptr, C.c_int), _ptr(a_col, C.c_int), _ptr(a_val, C.c_double)))

    def set_resampler(self, operator: np.ndarray, fft_decim: int = 0) -> None:
        """``raw_resampling``: dense ``(window_samples, n_in)`` operator of ``mne.filter.resample`` (processing/resample.py).  The
        pipeline then takes windows of ``n_in`` raw samples; must precede ``set_prefilters`` / ``set_notch``.  ``fft_decim = D``
        declares the operator as MNE's default FFT down-sampler by the integer factor D (the library verifies it and then runs it
        as two transforms per channel pair instead of the dense GEMM)."""
        a = _f64(operator)
        if a.ndim != 2 or a.shape[0] != self.W:
            raise ValueError(f"resampling operator must be ({self.W}, n_in), got {a.shape}")
        _lib.check(self.lib.nm_set_resampler(self._h, int(a.shape[1]), _ptr(a, C.c_double), int(fft_decim)))
        self.W_in = int(a.shape[1])

    def set_notch(self, taps: np.ndarray | None) -> None:
        if taps is None:
            return
        a = _f64(taps)
        _lib.check(self.lib.nm_set_notch(self._h, _ptr(a, C.c_double), int(a.size)))

    def set_precision(self, precision: str) -> None:
        """'f64' (default), 'f32' (float32 arithmetic inside the FFT convolution of the linear families: notch, band power) or
        'f32x2' (the same on packed float32 pairs, two channel pairs per item)."""
        if precision not in ("f64", "f32", "f32x2"):
            raise ValueError("precision must be 'f64', 'f32' or 'f32x2'")
        _lib.check(self.lib.nm_set_precision(self._h, ("f64", "f32", "f32x2").index(precision)))

    def set_fused(self, mode: int) -> None:
        """Organisation of the window chain: 2 = front kernel (bulk-copy staged raw rows, folded re-reference, notch + scan + segment
        DFT on chip; band-pass bank separate -- the default), 1 = the whole chain in ONE persistent kernel, 0 = one kernel per
        stage, -1 = as the environment says (NMB200_FUSED, default 2)."""
        _lib.check(self.lib.nm_set_fused(self._h, int(mode)))

    def set_raw_normalizer(self, method: str, clip: float, n_keep: int, add_samples: int) -> None:
        """RawNormalizer in front of the features (mean / median / zscore / zscore-median; scikit-learn methods are out of scope)."""
        if method not in RAW_NORM_METHODS:
            raise NotImplementedError(f"raw normalisation method '{method}' (scikit-learn quantile / power transformer) is out of scope")
        _lib.check(self.lib.nm_set_raw_normalizer(self._h, RAW_NORM_METHODS.index(method), float(clip or 0.0), int(n_keep), int(add_samples)))

    def set_prefilters(self, stages: Sequence[np.ndarray] | None) -> None:
        """PreprocessingFilter stages (one tap vector each, possibly of different lengths), applied in order before the notch."""
        for taps in stages or []:
            a = _f64(np.ravel(taps))
            _lib.check(self.lib.nm_add_prefilter(self._h, _ptr(a, C.c_double), int(a.size)))

    def set_nan_columns(self, names_by_raw_row: Sequence[str | None]) -> None:
        """names_by_raw_row[r] = channel name whose features become NaN when raw row r holds a NaN (or None).

        The reference matches by SUBSTRING of the feature key (stream/data_processor.py:299-303)."""
        # all occurrences of every name in ONE joined string (str.find runs at memory speed; a Python-level `name in key` over
        # rows x columns costs 25 ms at 256 channels x 3072 columns)
        sep = "\x00"
        joined = sep.join(self.columns)
        key_start = np.cumsum([0] + [len(k) + 1 for k in self.columns[:-1]]) if self.columns else np.zeros(0, dtype=np.int64)
        ptr = [0]
        cols: list[int] = []
        for name in names_by_raw_row:
            if name is not None and name != "" and sep not in name:
                hits, at = [], joined.find(name)
                while at >= 0:
                    hits.append(at)
                    at = joined.find(name, at + 1)
                if hits:
                    cols.extend(np.unique(np.searchsorted(key_start, hits, side="right") - 1).tolist())
            elif name is not None:
                cols.extend(i for i, key in enumerate(self.columns) if name in key)
            ptr.append(len(cols))
        a_ptr, a_cols = _i32(ptr), _i32(cols if cols else [0])
        _lib.check(self.lib.nm_set_nan_columns(self._h, _ptr(a_ptr, C.c_int), _ptr(a_cols, C.c_int)))

    def colmap(self, keys: Sequence[str | None]) -> np.ndarray:
        return _i32([self.col_of.get(k, -1) if k is not None else -1 for k in keys])

    # -- data path
    def upload(self, data: np.ndarray) -> None:
        if data.ndim != 2 or data.shape[0] != self.n_raw_rows:
            raise ValueError(f"expected an array of shape ({self.n_raw_rows}, n_samples), got {data.shape}")
        if data.dtype == np.float32:
            a = np.ascontiguousarray(data)
            _lib.check(self.lib.nm_upload_f32(self._h, a.ctypes.data_as(C.c_void_p), a.shape[1], a.shape[1]))
        else:
            a = np.ascontiguousarray(data, dtype=np.float64)
            _lib.check(self.lib.nm_upload_f64(self._h, a.ctypes.data_as(C.c_void_p), a.shape[1], a.shape[1]))
        self._keep_data = a

    def run(self, starts: Sequence[int], out: np.ndarray | None = None, download: bool = True) -> np.ndarray | None:
        """``out`` may be a column block ``big[:, :F]`` of a wider C-contiguous float64 matrix: the rows are then written with the
        wider matrix's row pitch (``nm_set_output_pitch``), so a caller that appends columns needs no second copy."""
        s = np.ascontiguousarray(starts, dtype=np.int64)
        n = int(s.size)
        if download:
            if out is None:
                out = np.empty((n, self.F), dtype=np.float64)
            assert out.shape == (n, self.F) and out.dtype == np.float64
            pitch = self.F
            if not out.flags.c_contiguous:
                assert out.strides[1] == 8 and out.strides[0] % 8 == 0 and out.strides[0] >= 8 * self.F, "unsupported output layout"
                pitch = out.strides[0] // 8
            _lib.check(self.lib.nm_set_output_pitch(self._h, 0 if pitch == self.F else pitch))
            try:
                _lib.check(self.lib.nm_run_windows(self._h, _ptr(s, C.c_longlong), n, out.ctypes.data_as(C.c_void_p)))
            finally:
                if pitch != self.F:
                    self.lib.nm_set_output_pitch(self._h, 0)
            return out
        _lib.check(self.lib.nm_run_windows(self._h, _ptr(s, C.c_longlong), n, None))
        return None

    def download(self, n_windows: int, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((n_windows, self.F), dtype=np.float64)
        _lib.check(self.lib.nm_download(self._h, out.ctypes.data_as(C.c_void_p), int(n_windows)))
        return out

    def process_window(self, window: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(window, dtype=np.float64)
        if a.shape != (self.n_raw_rows, self.W_in):
            raise ValueError(f"expected a window of shape ({self.n_raw_rows}, {self.W_in}), got {a.shape}")
        out = np.empty(self.F, dtype=np.float64)
        _lib.check(self.lib.nm_process_window(self._h, a.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)))
        return out

    def preprocess_window(self, window: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(window, dtype=np.float64)
        if a.shape != (self.n_raw_rows, self.W_in):
            raise ValueError(f"expected a window of shape ({self.n_raw_rows}, {self.W_in}), got {a.shape}")
        out = np.empty((self.n_ch, self.W), dtype=np.float64)
        _lib.check(self.lib.nm_preprocess_window(self._h, a.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)))
        return out

    # -- streaming entry (csrc/nm_stream.cuh): ring of page-locked slots, asynchronous submit / wait, CUDA-graph replay
    def stream_open(self, slots: int = 2, f32: bool = False, graph: bool | None = None) -> None:
        """``graph``: True = replay a captured CUDA graph per window, False = eager launches, None = NMB200_STREAM_GRAPH (default on)."""
        _lib.check(self.lib.nm_stream_open(self._h, int(slots), int(f32), -1 if graph is None else int(bool(graph))))
        self._stream_slots, self._stream_dtype = int(slots), (np.float32 if f32 else np.float64)

    def stream_input(self, slot: int) -> np.ndarray:
        """Writable ``(n_raw_rows, W_in)`` view of the slot's page-locked input block: produce the next window straight into it."""
        ptr, n = C.c_void_p(), C.c_longlong()
        _lib.check(self.lib.nm_stream_input(self._h, int(slot), C.byref(ptr), C.byref(n)))
        buf = (C.c_char * n.value).from_address(ptr.value)
        return np.frombuffer(buf, dtype=self._stream_dtype).reshape(self.n_raw_rows, self.W_in)

    def stream_submit(self, slot: int) -> None:
        """Enqueue the slot's window (H2D, every kernel, D2H of the feature row); returns without waiting for the GPU."""
        _lib.check(self.lib.nm_stream_submit(self._h, int(slot)))

    def stream_wait(self, slot: int) -> np.ndarray:
        """Block until the slot's feature row has landed; returns a view of the slot's page-locked output (valid until the slot
        is submitted again)."""
        ptr = C.c_void_p()
        _lib.check(self.lib.nm_stream_wait(self._h, int(slot), C.byref(ptr)))
        return np.ctypeslib.as_array((C.c_double * self.F).from_address(ptr.value))

    def stream_stats(self) -> dict[str, int]:
        w, g, n, pt = C.c_longlong(), C.c_longlong(), C.c_longlong(), C.c_longlong()
        _lib.check(self.lib.nm_stream_stats(self._h, C.byref(w), C.byref(g), C.byref(n), C.byref(pt)))
        return {"windows": w.value, "graph_launches": g.value, "graph_kernel_nodes": n.value, "patched_arguments": pt.value}

    def stream_close(self) -> None:
        _lib.check(self.lib.nm_stream_close(self._h))

    def add_feature_normalizer(self, method: str, clip: float, n_keep: int, columns: Sequence[str]) -> None:
        cols = _i32([self.col_of[k] for k in columns] or [0])
        _lib.check(self.lib.nm_add_feature_normalizer(self._h, FEATURE_NORM_METHODS.index(method), float(clip or 0.0), int(n_keep),
                                                      len(columns), _ptr(cols, C.c_int)))

    # -- measurement
    def timer_start(self) -> None:
        _lib.check(self.lib.nm_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_double(0)
        _lib.check(self.lib.nm_timer_stop(self._h, C.byref(ms)))
        return ms.value

    PROFILE_FAMILIES = ("prep", "notch", "scan", "spectral", "bandpower", "sharpwave", "burst_envelope", "burst_threshold",
                        "burst_features", "normalizer", "nan", "fused", "resample")

    def prepare_resident(self) -> None:
        _lib.check(self.lib.nm_prepare_resident(self._h))

    def synchronize(self) -> None:
        _lib.check(self.lib.nm_synchronize(self._h))

    def set_profiling(self, enabled: bool) -> None:
        _lib.check(self.lib.nm_set_profiling(self._h, int(enabled)))

    def profile(self) -> dict[str, tuple[float, int]]:
        n = len(self.PROFILE_FAMILIES)
        ms = (C.c_double * n)()
        cnt = (C.c_longlong * n)()
        self.lib.nm_get_profile(self._h, ms, cnt, n)
        return {name: (ms[i], cnt[i]) for i, name in enumerate(self.PROFILE_FAMILIES) if cnt[i]}

    @property
    def chunk_windows(self) -> int:
        return int(self.lib.nm_chunk_windows(self._h))

    def set_burst_threshold_mode(self, incremental: bool) -> None:
        """Incremental sliding quantile (default) or per-window re-selection from the whole history (same results)."""
        _lib.check(self.lib.nm_set_burst_threshold_mode(self._h, int(incremental)))

    def burst_threshold_stats(self) -> tuple[int, int]:
        """(bracket rebuilds, windows served by the direct selection) summed over rows since the last reset."""
        a, b = C.c_longlong(), C.c_longlong()
        _lib.check(self.lib.nm_burst_threshold_stats(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def describe_plan(self) -> str:
        """One line per family: which kernel serves it (specialised / runtime-plan / generic), sizes, shared memory."""
        buf = C.create_string_buffer(4096)
        self.lib.nm_describe_plan(self._h, buf, len(buf))
        return buf.value.decode()

    @property
    def kernel_launches(self) -> int:
        return int(self.lib.nm_kernel_launches(self._h))

    def result_device_ptr(self) -> tuple[int, int, int]:
        p, rows, cols = C.c_void_p(), C.c_longlong(), C.c_int()
        _lib.check(self.lib.nm_result_device_ptr(self._h, C.byref(p), C.byref(rows), C.byref(cols)))
        return int(p.value or 0), rows.value, cols.value


def factor_reference_matrix(m: np.ndarray, min_group: int = 4, max_groups: int = 8):
    """Split ``m`` into ``gcoef @ indicator(groups) + remainder``.

    Columns that carry the same non-zero value in a row (the common-average pattern ``-1/(n-1)``)
    are served by one coefficient on the per-sample sum of that column group; whatever is left
    (diagonal corrections, bipolar references) stays in the sparse remainder.  Exact by construction:
    ``gcoef[i, g] * 1[j in g] + rem[i, j] == m[i, j]`` for every entry.
    """
    c = m.shape[0]
    # candidate groups: distinct supports of "large" equal-valued runs in the rows
    seen: dict[bytes, frozenset[int]] = {}
    for i in range(c):
        row = m[i]
        vals, counts = np.unique(row[row != 0], return_counts=True)
        for v, cnt in zip(vals, counts):
            if cnt >= min_group:
                # the row's own channel is excluded from its average; adding it back makes the supports of
                # all rows of one channel type identical
                mask = row == v
                mask[i] = True
                key = mask.tobytes()
                if key not in seen:
                    seen[key] = frozenset(np.flatnonzero(mask).tolist())
    # channel types produce identical supports once the row's own index is included (dict keeps first-seen order)
    uniq: list[frozenset[int]] = list(seen.values())
    # keep disjoint groups only (a channel belongs to at most one group)
    groups: list[frozenset[int]] = []
    for s in sorted(uniq, key=len, reverse=True):
        if len(groups) < max_groups and all(s.isdisjoint(g) for g in groups):
            groups.append(s)
    group_of = np.full(c, -1, dtype=np.int32)
    for g, members in enumerate(groups):
        for j in members:
            group_of[j] = g
    gcoef = np.zeros((c, len(groups)))
    rem = m.copy()
    member_idx = [np.array(sorted(members), dtype=np.int64) for members in groups]
    for i in range(c):
        for g, members in enumerate(member_idx):
            idx = members[members != i]
            if idx.size < min_group:
                continue
            vals, counts = np.unique(m[i, idx], return_counts=True)
            v = vals[np.argmax(counts)]
            if v == 0 or counts.max() < min_group:
                continue
            gcoef[i, g] = v
            rem[i, members] = m[i, members] - v
    return len(groups), group_of, gcoef, rem


# ------------------------------------------------------------------------------------- family specs
def band_items(settings) -> list[tuple[str, tuple[float, float]]]:
    return [(name, (float(fr[0]), float(fr[1]))) for name, fr in settings.frequency_ranges_hz.items()]


class ScanSpec:
    """Hjorth / Raw / LineLength; the three plugins keep their own key blocks but share one kernel."""

    def __init__(self, ch_names: Sequence[str], hjorth: bool = False, raw: bool = False, linelength: bool = False) -> None:
        self.ch_names = list(ch_names)
        self.hjorth, self.raw, self.linelength = hjorth, raw, linelength

    def keys_hjorth(self) -> list[str]:
        out = []
        for ch in self.ch_names:
            out += [f"{ch}_RawHjorth_Activity", f"{ch}_RawHjorth_Mobility", f"{ch}_RawHjorth_Complexity"]
        return out

    def keys_raw(self) -> list[str]:
        return ["_".join([ch, "raw"]) for ch in self.ch_names]

    def keys_linelength(self) -> list[str]:
        return [f"{ch}_LineLength" for ch in self.ch_names]

    def attach(self, pipe: Pipeline) -> None:
        slots: list[str | None] = []
        for ch in self.ch_names:
            slots += [f"{ch}_RawHjorth_Activity" if self.hjorth else None, f"{ch}_RawHjorth_Mobility" if self.hjorth else None,
                      f"{ch}_RawHjorth_Complexity" if self.hjorth else None, f"{ch}_raw" if self.raw else None,
                      f"{ch}_LineLength" if self.linelength else None]
        cm = pipe.colmap(slots)
        _lib.check(pipe.lib.nm_add_scan(pipe._h, int(self.hjorth), int(self.raw), int(self.linelength), _ptr(cm, C.c_int)))


def _periodic_window(name: str, n: int) -> np.ndarray:
    k = np.arange(n)
    if name == "hann":
        return 0.5 - 0.5 * np.cos(2.0 * np.pi * k / n)
    if name == "hamming":
        return 0.54 - 0.46 * np.cos(2.0 * np.pi * k / n)
    raise ValueError(name)


class SpectralSpec:
    """FFT / Welch / STFT (features/oscillatory.py) for a given window length."""

    def __init__(self, kind: str, osc_settings, bands: list[tuple[str, tuple[float, float]]], ch_names: Sequence[str],
                 sfreq: float, window_samples: int) -> None:
        self.kind = kind
        self.ch_names = list(ch_names)
        self.cfg = osc_settings
        w = int(window_samples)
        fs = int(sfreq)
        self.log = bool(osc_settings.log_transform)
        self.est = [e for e in osc_settings.features.get_enabled() if e in OSC_ESTIMATORS]
        self.want_spectrum = bool(osc_settings.return_spectrum)
        if kind == "fft":
            n_cfg = int(math.floor(osc_settings.windowlength_ms / 1000 * sfreq))
            freqs = np.fft.rfftfreq(n_cfg, 1 / math.floor(fs))
            self.nper = min(n_cfg, w)
            self.nseg, self.hop, self.start = 1, 0, w - self.nper
            self.ext_even, self.ext_len, self.detrend, self.power, self.scale, self.keep = 0, 0, 0, 0, 1.0, 0
            self.win = None
            closed = False
        elif kind == "welch":
            freqs = np.fft.rfftfreq(fs, 1 / fs)
            self.nper = min(fs, w)  # scipy clips nperseg to the input length
            noverlap = self.nper // 2
            self.hop = self.nper - noverlap
            self.nseg = (w - noverlap) // self.hop
            self.start = 0
            self.win = _periodic_window("hann", self.nper)
            self.ext_even, self.ext_len, self.detrend, self.power, self.keep = 0, 0, 1, 1, 0
            self.scale = 1.0 / (fs * float((self.win * self.win).sum()))
            closed = False
        elif kind == "stft":
            nperseg = int(osc_settings.windowlength_ms)
            freqs = np.fft.rfftfreq(nperseg, 1 / fs)
            self.nper = min(nperseg, w)
            noverlap = self.nper // 2
            self.hop = self.nper - noverlap
            half = self.nper // 2
            ext = w + 2 * half
            nadd = (-(ext - self.nper) % self.hop) % self.nper
            self.nseg = (ext + nadd - self.nper) // self.hop + 1
            self.start = -half
            self.win = _periodic_window("hamming", self.nper)
            self.ext_even, self.ext_len, self.detrend, self.power, self.keep = 1, half, 0, 0, 1
            self.scale = 1.0 / float(self.win.sum())
            closed = True
        else:
            raise ValueError(kind)
        self.freqs = freqs
        self.nbins = self.nper // 2 + 1
        self.bands = []
        for name, (lo, hi) in bands:
            sel = np.where((freqs >= lo) & ((freqs <= hi) if closed else (freqs < hi)))[0]
            k_lo, k_hi = (int(sel[0]), int(sel[-1]) + 1) if sel.size else (0, 0)
            if k_hi > self.nbins:
                raise IndexError(
                    f"{kind}: band '{name}' needs bin {k_hi - 1} but the {self.nper}-sample spectrum has {self.nbins} bins"
                )
            self.bands.append((name, k_lo, k_hi))
        if self.want_spectrum and len(freqs) > self.nbins:
            raise IndexError(f"{kind}: return_spectrum needs {len(freqs)} bins but the spectrum has {self.nbins}")

    def keys(self) -> list[str]:
        out: list[str] = []
        for band, _, _ in self.bands:
            for est in self.est:
                out += [f"{ch}_{self.kind}_{band}_{est}" for ch in self.ch_names]
        if self.want_spectrum:
            for ch in self.ch_names:
                # dict semantics: a repeated key (int(f) collides for sub-Hz bin spacing) keeps its first position
                out += list(dict.fromkeys(f"{ch}_{self.kind}_psd_{int(f)}" for f in self.freqs))
        return out

    def attach(self, pipe: Pipeline) -> None:
        per_ch = len(self.bands) * 4 + self.nbins
        slots: list[str | None] = [None] * (len(self.ch_names) * per_ch)
        for ci, ch in enumerate(self.ch_names):
            for bi, (band, _, _) in enumerate(self.bands):
                for est in self.est:
                    slots[ci * per_ch + bi * 4 + OSC_ESTIMATORS.index(est)] = f"{ch}_{self.kind}_{band}_{est}"
            if self.want_spectrum:
                last: dict[str, int] = {}
                for k, f in enumerate(self.freqs):
                    last[f"{ch}_{self.kind}_psd_{int(f)}"] = k  # later bins overwrite earlier ones ("last wins")
                for key, k in last.items():
                    slots[ci * per_ch + len(self.bands) * 4 + k] = key
        cm = pipe.colmap(slots)
        lo = _i32([b[1] for b in self.bands] or [0])
        hi = _i32([b[2] for b in self.bands] or [0])
        cfg = _lib.SpectralCfg()
        cfg.nper, cfg.nseg, cfg.hop, cfg.start = self.nper, self.nseg, self.hop, self.start
        cfg.ext_even, cfg.ext_len, cfg.detrend, cfg.power = self.ext_even, self.ext_len, self.detrend, self.power
        cfg.scale, cfg.log, cfg.keep_segments = self.scale, int(self.log), self.keep
        cfg.n_bands = len(self.bands)
        cfg.est_mask = sum(1 << OSC_ESTIMATORS.index(e) for e in self.est)
        cfg.want_spectrum = int(self.want_spectrum)
        win = None if self.win is None else _f64(self.win)
        cfg.win = _ptr(win, C.c_double) if win is not None else None
        cfg.band_lo, cfg.band_hi, cfg.colmap = _ptr(lo, C.c_int), _ptr(hi, C.c_int), _ptr(cm, C.c_int)
        _lib.check(pipe.lib.nm_add_spectral(pipe._h, C.byref(cfg)))


class BandpowerSpec:
    """features/bandpower.py: FIR bank over all frequency ranges + tail-variance features."""

    def __init__(self, bp_settings, bands, ch_names: Sequence[str], sfreq: float) -> None:
        from .filter.mne_filter import MNEFilter

        if bp_settings.kalman_filter:
            raise NotImplementedError("bandpass_filter_settings.kalman_filter is out of scope of the B200 hot path")
        self.ch_names = list(ch_names)
        self.bands = bands
        self.feats = bp_settings.bandpower_features.get_enabled()
        self.log = bool(bp_settings.log_transform)
        self.bank = MNEFilter([fr for _, fr in bands], sfreq, filter_length=sfreq - 1).filter_bank
        seg = {k.replace(" ", "_"): v for k, v in bp_settings.segment_lengths_ms.items()}
        self.seglen = [int(np.floor(sfreq / 1000 * seg[name])) for name, _ in bands]

    def keys(self) -> list[str]:
        return ["_".join([ch, "bandpass", ft, band]) for ch in self.ch_names for band, _ in self.bands for ft in self.feats]

    def attach(self, pipe: Pipeline) -> None:
        order = ("activity", "mobility", "complexity")
        slots = []
        for ch in self.ch_names:
            for band, _ in self.bands:
                slots += ["_".join([ch, "bandpass", ft, band]) if ft in self.feats else None for ft in order]
        cm, taps, seg = pipe.colmap(slots), _f64(self.bank), _i32(self.seglen)
        _lib.check(pipe.lib.nm_add_bandpower(pipe._h, len(self.bands), _ptr(taps, C.c_double), taps.shape[1], _ptr(seg, C.c_int),
                                             int("activity" in self.feats), int("mobility" in self.feats),
                                             int("complexity" in self.feats), int(self.log), _ptr(cm, C.c_int)))


class BurstsSpec:
    """features/bursts.py."""

    SLOTS = ("duration_mean", "duration_max", "amplitude_mean", "amplitude_max", "burst_rate_per_s", "in_burst")

    def __init__(self, settings, ch_names: Sequence[str], sfreq: float) -> None:
        from .filter.mne_filter import MNEFilter

        bs = settings.bursts_settings
        self.ch_names = list(ch_names)
        self.bands = list(bs.frequency_bands)
        self.sfreq = sfreq
        self.seg_s = settings.segment_length_features_ms / 1000
        self.samples_overlap = int(sfreq * self.seg_s / settings.sampling_rate_features_hz)
        self.ring = int(sfreq * bs.time_duration_s)
        self.q = bs.threshold / 100
        self.feats = bs.burst_features.get_enabled()
        ranges = [(settings.frequency_ranges_hz[b][0], settings.frequency_ranges_hz[b][1]) for b in self.bands]
        self.bank = MNEFilter(ranges, sfreq, filter_length=sfreq - 1).filter_bank

    def _names(self, ch: str, band: str) -> dict[str, list[str]]:
        p = f"{ch}_bursts_{band}_"
        return {"duration": [p + "duration_mean", p + "duration_max"], "amplitude": [p + "amplitude_mean", p + "amplitude_max"],
                "burst_rate_per_s": [p + "burst_rate_per_s"], "in_burst": [p + "in_burst"]}

    def keys(self) -> list[str]:
        out = []
        for ch in self.ch_names:
            for band in self.bands:
                names = self._names(ch, band)
                for ft in self.feats:
                    out += names[ft]
        return out

    def attach(self, pipe: Pipeline) -> None:
        enabled = set(self.keys())
        slots = []
        for ch in self.ch_names:
            for band in self.bands:
                for s in self.SLOTS:
                    k = f"{ch}_bursts_{band}_{s}"
                    slots.append(k if k in enabled else None)
        cm, taps = pipe.colmap(slots), _f64(self.bank)
        _lib.check(pipe.lib.nm_add_bursts(pipe._h, len(self.bands), _ptr(taps, C.c_double), taps.shape[1], self.samples_overlap,
                                          self.ring, float(self.q), float(self.sfreq), float(self.seg_s), _ptr(cm, C.c_int)))


class SharpwaveSpec:
    """features/sharpwaves.py."""

    def __init__(self, settings, ch_names: Sequence[str], sfreq: float) -> None:
        from .filter.fir_design import design_fir

        sw = settings.sharpwave_analysis_settings
        self.sw = sw
        self.ch_names = list(ch_names)
        self.sfreq = sfreq
        self.filters = []
        for fr in sw.filter_ranges_hz:
            assert fr[1] < sfreq, f"Filter range has to be smaller than sfreq, got sfreq {sfreq} and filter range {fr}"
            if fr[0] is None:
                raise NotImplementedError("sharp-wave 'no_filter' ranges are not supported")
            self.filters.append((f"range_{fr[0]:.0f}_{fr[1]:.0f}", design_fir(sfreq, fr[0], fr[1])))
        if len({len(t) for _, t in self.filters}) != 1:
            # The reference's branch for unequal lengths (features/sharpwaves.py:249-253) hands fftconvolve a (channels, taps) filter
            # TILE next to the (channels, samples) data, i.e. it runs a 2-D convolution that sums neighbouring CHANNELS into every
            # row (checked against the unmodified reference).  That is not a per-channel FIR and is not reproduced here.
            raise NotImplementedError("sharp-wave filters of different lengths: the reference convolves across channels in that "
                                      "branch (2-D fftconvolve); choose filter ranges with equal transition bands")
        self.used = sw.sharpwave_features.get_enabled()
        est_of = {ft: [e for e in SW_ESTIMATORS if ft in getattr(sw.estimator, e)]
                  for e in SW_ESTIMATORS for ft in getattr(sw.estimator, e)}
        self.combos = [(ft, e) for ft in self.used for e in est_of[ft]]
        self.pair = bool(sw.apply_estimator_between_peaks_and_troughs)
        # one polarity only (detect_peaks.estimate / detect_troughs.estimate): fine for the un-paired "_analyze_Peak/_Trough" keys;
        # the paired form indexes both values and raises in the reference too (features/sharpwaves.py:297-302)
        self.polarities = [pol for pol, on in (("Peak", sw.detect_peaks.estimate), ("Trough", sw.detect_troughs.estimate)) if on]
        if self.pair and len(self.polarities) != 2:
            raise IndexError("list index out of range: apply_estimator_between_peaks_and_troughs needs both detect_peaks.estimate "
                             "and detect_troughs.estimate (the reference raises the same IndexError)")
        self.num_peaks = bool(sw.sharpwave_features.num_peaks)

    def keys(self) -> list[str]:
        out: list[str] = []
        if self.pair:
            for ch in self.ch_names:
                for fname, _ in self.filters:
                    out += [f"{ch}_Sharpwave_{e.title()}_{ft}_{fname}" for ft, e in self.combos if ft != "num_peaks"]
            if self.num_peaks:
                out += [f"{ch}_Sharpwave_num_peaks_{fname}" for ch in self.ch_names for fname, _ in self.filters]
        else:
            for ch in self.ch_names:
                for fname, _ in self.filters:
                    # the reference flattens {key: {"Peak": v, "Trough": v}} in key-insertion order (features/sharpwaves.py:323-326):
                    # both polarities of a key are adjacent, num_peaks sits at its position among the enabled features
                    for ft, e in self.combos:
                        for pol in self.polarities:
                            k = (f"{ch}_Sharpwave_num_peaks_{fname}" if ft == "num_peaks"
                                 else f"{ch}_Sharpwave_{e.title()}_{ft}_{fname}") + "_analyze_" + pol
                            if k not in out:
                                out.append(k)
        return out

    def attach(self, pipe: Pipeline) -> None:
        n_combo = len(self.combos)
        slots: list[str | None] = []

        def both(base: str) -> list[str | None]:  # (the kernel analyses both polarities; a disabled one has no column)
            return [base + "_analyze_" + pol if pol in self.polarities else None for pol in ("Peak", "Trough")]

        for ch in self.ch_names:
            for fname, _ in self.filters:
                for ft, e in self.combos:
                    base = f"{ch}_Sharpwave_{e.title()}_{ft}_{fname}"
                    if ft == "num_peaks":
                        slots += [None, None]
                    elif self.pair:
                        slots += [base, None]
                    else:
                        slots += both(base)
                base = f"{ch}_Sharpwave_num_peaks_{fname}"
                listed = self.num_peaks and (self.pair or any(ft == "num_peaks" for ft, _ in self.combos))
                if not listed:
                    slots += [None, None]
                elif self.pair:
                    slots += [base, None]
                else:
                    slots += both(base)
        cm = pipe.colmap(slots)
        taps = _f64(np.vstack([t for _, t in self.filters]))
        feat = _i32([SW_FEATURES.index(ft) for ft, _ in self.combos] or [0])
        est = _i32([SW_ESTIMATORS.index(e) for _, e in self.combos] or [0])
        dt = self.sw.detect_troughs
        _lib.check(pipe.lib.nm_add_sharpwave(
            pipe._h, len(self.filters), _ptr(taps, C.c_double), taps.shape[1], int(math.ceil(dt.distance_peaks_ms)),
            int(math.ceil(dt.distance_troughs_ms)), int(5 * (1000 / self.sfreq)), float(1000 / self.sfreq), n_combo,
            _ptr(feat, C.c_int), _ptr(est, C.c_int), int(self.pair), int(self.num_peaks), _ptr(cm, C.c_int)))


# ------------------------------------------------------------------------------------- stand-alone FIR
def filter_rows(owner, data: np.ndarray, mode: str) -> np.ndarray:
    """Run ``nm_fir_apply`` for :class:`MNEFilter` ('same') and :class:`NotchFilter` ('reflect')."""
    lib = _lib.load()
    data = np.ascontiguousarray(data, dtype=np.float64)
    taps = _f64(np.atleast_2d(owner.filter_bank))
    n_ch, w = data.shape
    out = np.empty((n_ch, taps.shape[0], w), dtype=np.float64)
    _lib.check(lib.nm_fir_apply(0, _ptr(taps, C.c_double), taps.shape[0], taps.shape[1], 0 if mode == "same" else 1,
                                data.ctypes.data_as(C.c_void_p), n_ch, w, out.ctypes.data_as(C.c_void_p)))
    return out if mode == "same" else out[:, 0, :]


def reref_rows(owner, data: np.ndarray) -> np.ndarray:
    """``ReReferencer.process`` for a stand-alone instance: a preprocessing-only pipeline per window length."""
    n_ch, w = data.shape
    pipe = owner._pipes.get(w)
    if pipe is None:
        pipe = Pipeline(n_ch, n_ch, w, ["_unused"])
        pipe.set_reref(owner.ref_matrix)
        ScanSpec([f"_c{i}" for i in range(n_ch)]).attach(pipe)
        pipe.finalize()
        owner._pipes[w] = pipe
    return pipe.preprocess_window(data)


class IdentityNormPipeline:
    """Feature normaliser for a bare vector: every entry travels as a 3-sample constant 'channel' whose last
    sample is the feature (scan kernel, ``raw``), followed by the rolling-normalisation kernel."""

    def __init__(self, n: int, method_index: int, clip: float, n_keep: int, device: int = 0) -> None:
        names = [f"f{i}" for i in range(n)]
        cols = [f"{c}_raw" for c in names]
        self.pipe = Pipeline(n, n, 3, cols, device=device)
        ScanSpec(names, raw=True).attach(self.pipe)
        self.pipe.add_feature_normalizer(FEATURE_NORM_METHODS[method_index], clip, n_keep, cols)
        self.pipe.finalize()

    def step(self, v: np.ndarray) -> np.ndarray:
        return self.pipe.process_window(np.repeat(v[:, None], 3, axis=1))
