"""TEST INFRASTRUCTURE (oracle) -- CPU restatement of the reference hot path.

A NumPy/SciPy restatement of py_neuromodulation's per-window feature-extraction loop
(``Stream.run -> DataProcessor.process -> notch / re-reference -> features``) written
from the behaviour documented in SURVEY.md section 8a; every function cites the reference
``file:line`` (relative to ``/root/reference/py_neuromodulation``) it follows.  It calls
the same third-party primitives the reference calls (``scipy.fft.rfft``,
``scipy.signal.{welch,stft,fftconvolve,hilbert,find_peaks}``) and the restated
``mne.filter`` of ``oracle/mne_filter_restated.py``.

Pinning: ``tests/test_oracle_golden.py`` checks this module against fixtures generated
from the UNMODIFIED reference files (``tests/golden/make_golden.py`` via
``oracle/ref_shim.py``) and, when ``/root/reference`` is present, against the live
reference.  The ``mne.filter`` part is "parity unpinned" (no MNE install, no stored
coefficients in the reference) -- see ``oracle/mne_filter_restated.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module; the product never does.

Settings are consumed as the plain nested ``dict`` produced by ``NMSettings.model_dump()``
(same layout as ``default_settings.yaml``), so the oracle depends on neither the reference
package nor the product package.
"""

from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np

from oracle import mne_filter_restated as mne_filter

# field order of the reference's selectors (stream/settings.py:41-55 etc.)
FEATURE_ORDER = [
    "raw_hjorth", "return_raw", "bandpass_filter", "stft", "fft", "welch", "sharpwave_analysis",
    "fooof", "nolds", "coherence", "bursts", "linelength", "mne_connectivity", "bispectrum",
]
PREPROC_ORDER = ["preprocessing_filter", "notch_filter", "raw_resampling", "re_referencing", "raw_normalization"]
OSC_EST_ORDER = ["mean", "median", "std", "max"]
BP_FEAT_ORDER = ["activity", "mobility", "complexity"]
BURST_FEAT_ORDER = ["duration", "amplitude", "burst_rate_per_s", "in_burst"]
SW_FEAT_ORDER = [
    "peak_left", "peak_right", "num_peaks", "trough", "width", "prominence", "interval",
    "decay_time", "rise_time", "sharpness", "rise_steepness", "decay_steepness", "slope_ratio",
]
SW_EST_ORDER = ["mean", "median", "max", "min", "var"]


def _enabled(sel: dict, order: list[str]) -> list[str]:
    out = [k for k in order if sel.get(k) is True]
    out += [k for k, v in sel.items() if k not in order and v is True]  # user features last
    return out


def _band_tuple(v) -> tuple[float, float]:
    if isinstance(v, dict):
        return float(v["frequency_low_hz"]), float(v["frequency_high_hz"])
    return float(v[0]), float(v[1])


def bands_of(settings: dict) -> "OrderedDict[str, tuple[float, float]]":
    return OrderedDict((k.replace(" ", "_"), _band_tuple(v)) for k, v in settings["frequency_ranges_hz"].items())


# ----------------------------------------------------------------------------- a1: window grid
def window_grid(n_samples: int, sfreq: float, rate_hz: float, segment_ms: float):
    """stream/generator.py:34-53 + stream/stream.py:310 -> list of (i0, i1, time_ms)."""
    seg = segment_ms / 1000 * sfreq
    stride = sfreq / rate_hz
    out = []
    k = 0
    while True:
        start = stride * k
        end = start + seg
        k += 1
        i0, i1 = int(start), int(end)
        if i1 > n_samples:
            break
        ts_last = np.arange(start, end)[-1] / sfreq
        out.append((i0, i1, float(np.ceil(ts_last * 1000 + 1))))
    return out


# ----------------------------------------------------------------------------- channel table
def default_channels(n_ch: int) -> dict:
    """utils/channels.py:257-309 (all ecog, average reference, ``ch{i}_avgref``)."""
    names = [f"ch{i}" for i in range(n_ch)]
    return {
        "name": names,
        "rereference": ["average"] * n_ch,
        "used": [1] * n_ch,
        "target": [0] * n_ch,
        "type": ["ecog"] * n_ch,
        "status": ["good"] * n_ch,
        "new_name": [f"{n}_avgref" for n in names],
    }


def channel_info(ch: dict):
    """stream/data_processor.py:141-160."""
    n = len(ch["name"])
    names_used = [ch["new_name"][i] for i in range(n) if ch["used"][i] == 1 and ch["status"][i] == "good"]
    feature_idx = [i for i in range(n) if ch["used"][i] and not ch["target"][i] and ch["status"][i] == "good"]
    return names_used, feature_idx


def reref_matrix(ch: dict):
    """processing/rereference.py:31-86 -> (n_good_used, n_good_used) matrix or None."""
    used = [i for i in range(len(ch["name"])) if ch["used"][i] == 1]
    if len(used) in (0, 1):
        return None
    names = [ch["name"][i] for i in used]
    types = [ch["type"][i] for i in used]
    refs = [ch["rereference"][i] for i in used]
    status = [ch["status"][i] for i in used]
    n = len(used)
    m = np.zeros((n, n))
    for i in range(n):
        m[i, i] = 1
        ref = refs[i]
        if ref is None or (isinstance(ref, float) and math.isnan(ref)) or str(ref).lower() == "none" or status[i] != "good":
            continue
        if str(ref).lower() == "average":
            idx = [j for j in range(n) if types[j] == types[i] and status[j] == "good" and j != i]
        else:
            idx = []
            for rc in str(ref).split("&"):
                if rc not in names:
                    raise ValueError(f"One or more of the reference channels are not part of the recording channels: {rc}")
                if rc == names[i]:
                    raise ValueError(f"You cannot rereference to the same channel: {rc}")
                idx.append(names.index(rc))
        m[i, idx] = -1 / len(idx)
    good = [i for i in range(n) if status[i] == "good"]
    return m[np.ix_(good, good)]


# ----------------------------------------------------------------------------- a4: notch
def design_notch(sfreq: float, line_noise: float | None, freqs=None, notch_widths=3, trans_bandwidth=6.8):
    """filter/notch_filter.py:9-76 -> taps or None."""
    if line_noise is None and freqs is None:
        raise ValueError("Either line_noise or freqs must be defined if notch_filter is activated.")
    if freqs is None:
        freqs = np.arange(line_noise, sfreq / 2, line_noise, dtype=int)
    freqs = np.asarray(freqs)
    if freqs.size > 0 and freqs[-1] >= sfreq / 2:
        freqs = freqs[:-1]
    if freqs.size == 0:
        return None
    if notch_widths is None:
        widths = freqs / 200.0
    else:
        widths = np.atleast_1d(notch_widths)
        if np.any(widths < 0):
            raise ValueError("notch_widths must be >= 0")
        if len(widths) == 1:
            widths = widths[0] * np.ones_like(freqs)
        elif len(widths) != len(freqs):
            raise ValueError("notch_widths must be None, scalar, or the same length as freqs")
    tb_half = trans_bandwidth / 2.0
    lows = [f - w / 2.0 - tb_half for f, w in zip(freqs, widths)]
    highs = [f + w / 2.0 + tb_half for f, w in zip(freqs, widths)]
    return mne_filter.create_filter(
        None, sfreq, l_freq=highs, h_freq=lows, filter_length=int(sfreq - 1),
        l_trans_bandwidth=tb_half, h_trans_bandwidth=tb_half,
    )


def apply_notch(x: np.ndarray, taps) -> np.ndarray:
    """filter/notch_filter.py:78-93."""
    if taps is None:
        return x
    return mne_filter._overlap_add_filter(x, taps, phase="zero", copy=True, pad="reflect_limited")


# ----------------------------------------------------------------------------- a10: FIR bank
def design_bank(f_ranges, sfreq: float, filter_length=None, l_tb=4, h_tb=4) -> np.ndarray:
    """filter/mne_filter.py:35-80 (incl. the ValueError fallback to auto length/bandwidths)."""
    if filter_length is None:
        filter_length = "999ms"
    if isinstance(filter_length, float):
        filter_length = int(filter_length)
    bank = []
    for lo, hi in f_ranges:
        try:
            h = mne_filter.create_filter(None, sfreq, l_freq=lo, h_freq=hi, l_trans_bandwidth=l_tb,
                                         h_trans_bandwidth=h_tb, filter_length=filter_length)
        except ValueError:
            h = mne_filter.create_filter(None, sfreq, l_freq=lo, h_freq=hi)
        bank.append(h)
    return np.vstack(bank)


def apply_bank(x: np.ndarray, bank: np.ndarray) -> np.ndarray:
    """filter/mne_filter.py:82-128: (C, W) -> (C, n_filters, W) 'same' FFT convolution."""
    from scipy.signal import fftconvolve

    if x.ndim == 1:
        x = x[None, :]
    nf = bank.shape[0]
    xt = np.tile(x[:, None, :], (1, nf, 1))
    ft = np.tile(bank[None, :, :], (x.shape[0], 1, 1))
    out = fftconvolve(xt, ft, axes=2, mode="same")
    if x.shape[1] != out.shape[-1]:
        mid = out.shape[-1] // 2
        out = out[:, :, mid - x.shape[1] // 2 : mid + x.shape[1] // 2]
    return out


# ----------------------------------------------------------------------------- a7-a9: oscillatory
_EST = {"mean": np.nanmean, "median": np.nanmedian, "std": np.nanstd, "max": np.nanmax}


class OscOracle:
    """features/oscillatory.py:37-250 (FFT / Welch / STFT)."""

    def __init__(self, kind: str, settings: dict, ch_names, sfreq):
        from scipy.fft import rfftfreq

        self.kind = kind
        self.cfg = settings[f"{kind}_settings"]
        self.sfreq = int(sfreq)
        self.ch_names = list(ch_names)
        assert self.cfg["windowlength_ms"] <= settings["segment_length_features_ms"]
        bands = bands_of(settings)
        if kind == "fft":
            self.n_win = int(np.floor(self.cfg["windowlength_ms"] / 1000 * sfreq))
            self.freqs = rfftfreq(self.n_win, 1 / np.floor(self.sfreq))
            closed = False
        elif kind == "welch":
            self.freqs = rfftfreq(self.sfreq, 1 / self.sfreq)
            closed = False
        else:
            self.nperseg = self.cfg["windowlength_ms"]
            self.freqs = rfftfreq(self.nperseg, 1 / self.sfreq)
            closed = True
        self.idx = []
        for name, (lo, hi) in bands.items():
            if closed:
                sel = np.where((self.freqs >= lo) & (self.freqs <= hi))[0]
            else:
                sel = np.where((self.freqs >= lo) & (self.freqs < hi))[0]
            self.idx.append((name, sel))
        self.est = _enabled(self.cfg["features"], OSC_EST_ORDER)

    def spectrum(self, data: np.ndarray) -> np.ndarray:
        if self.kind == "fft":
            from scipy.fft import rfft

            z = np.abs(rfft(data[:, -self.n_win:]))
        elif self.kind == "welch":
            from scipy.signal import welch

            _, z = welch(data, fs=self.sfreq, window="hann", nperseg=self.sfreq, noverlap=None)
        else:
            from scipy.signal import stft

            _, _, zxx = stft(data, fs=self.sfreq, window="hamming", nperseg=self.nperseg, boundary="even")
            z = np.abs(zxx)
        if self.cfg["log_transform"]:
            with np.errstate(divide="ignore"):
                z = np.log10(z)
        return z

    def calc(self, data: np.ndarray) -> dict:
        z = self.spectrum(data)
        out: dict = {}
        axis = (1, 2) if self.kind == "stft" else 1
        for band, sel in self.idx:
            zb = z[:, sel, :] if self.kind == "stft" else z[:, sel]
            for est in self.est:
                with np.errstate(all="ignore"):
                    import warnings

                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore")
                        res = _EST[est](zb, axis=axis)
                for ci, ch in enumerate(self.ch_names):
                    out[f"{ch}_{self.kind}_{band}_{est}"] = res[ci]
        if self.cfg["return_spectrum"]:
            for ci, ch in enumerate(self.ch_names):
                row = z[ci].mean(axis=1) if self.kind == "stft" else z[ci]
                for k, f in enumerate(self.freqs):
                    out[f"{ch}_{self.kind}_psd_{int(f)}"] = row[k]
        return out


# ----------------------------------------------------------------------------- a11: band power
def _var0(v):
    with np.errstate(all="ignore"):
        return np.var(v)


class BandPowerOracle:
    """features/bandpower.py:98-207 (Kalman filter not restated: off by default, out of scope)."""

    def __init__(self, settings: dict, ch_names, sfreq):
        self.cfg = settings["bandpass_filter_settings"]
        if self.cfg.get("kalman_filter"):
            raise NotImplementedError("Kalman smoothing is out of scope (SURVEY.md section 2 row 20)")
        self.sfreq = sfreq
        self.ch_names = list(ch_names)
        bands = bands_of(settings)
        self.bank = design_bank(list(bands.values()), sfreq, filter_length=sfreq - 1)
        seg = {k.replace(" ", "_"): v for k, v in self.cfg["segment_lengths_ms"].items()}
        feats = _enabled(self.cfg["bandpower_features"], BP_FEAT_ORDER)
        self.params = []
        for ci, ch in enumerate(self.ch_names):
            for bi, band in enumerate(bands.keys()):
                seglen = int(np.floor(sfreq / 1000 * seg[band]))
                for ft in feats:
                    self.params.append((ci, bi, seglen, ft, "_".join([ch, "bandpass", ft, band])))

    def calc(self, data: np.ndarray) -> dict:
        y = apply_bank(data, self.bank)
        out = {}
        with np.errstate(all="ignore"):
            for ci, bi, seglen, ft, name in self.params:
                seg = y[ci, bi, -seglen:]
                if ft == "activity":
                    v = _var0(seg)
                    if self.cfg["log_transform"]:
                        v = np.log10(v)
                elif ft == "mobility":
                    v = np.sqrt(_var0(np.diff(seg)) / _var0(seg))
                else:
                    d1 = np.diff(seg)
                    v1 = _var0(d1)
                    mob = np.sqrt(v1 / _var0(seg))
                    v = np.sqrt(_var0(np.diff(d1)) / v1) / mob
                out[name] = np.nan_to_num(v)
        return out


# ----------------------------------------------------------------------------- a12/a13: scan features
def hjorth(data: np.ndarray, ch_names) -> dict:
    """features/hjorth_raw.py:24-42."""
    with np.errstate(all="ignore"):
        v0 = np.var(data, axis=-1)
        d1 = np.diff(data, axis=-1)
        d2 = np.diff(d1, axis=-1)
        v1 = np.var(d1, axis=-1)
        v2 = np.var(d2, axis=-1)
        act = np.nan_to_num(v0)
        mob = np.nan_to_num(np.sqrt(v1 / v0))
        comp = np.nan_to_num(np.sqrt(v2 / v1) / mob)
    out = {}
    for ci, ch in enumerate(ch_names):
        out[f"{ch}_RawHjorth_Activity"] = act[ci]
        out[f"{ch}_RawHjorth_Mobility"] = mob[ci]
        out[f"{ch}_RawHjorth_Complexity"] = comp[ci]
    return out


def raw_last(data: np.ndarray, ch_names) -> dict:
    """features/hjorth_raw.py:51-57."""
    return {f"{ch}_raw": data[ci, -1] for ci, ch in enumerate(ch_names)}


def linelength(data: np.ndarray, ch_names) -> dict:
    """features/linelength.py:11-21 (the reference divides by (W-1) twice)."""
    ll = np.mean(np.abs(np.diff(data, axis=-1)) / (data.shape[1] - 1), axis=-1)
    return {f"{ch}_LineLength": ll[ci] for ci, ch in enumerate(ch_names)}


# ----------------------------------------------------------------------------- a14: bursts
def quantile_linear(buf: np.ndarray, q: float) -> np.ndarray:
    """NumPy 'linear' quantile along the last axis without touching ``buf`` (the *intended*
    semantics of features/bursts.py:171; see SURVEY.md headline facts for the in-place defect)."""
    n = buf.shape[-1]
    virt = n * q + (1 + q * (1 - 1 - 1)) - 1  # numpy _compute_virtual_index, alpha = beta = 1
    lo = int(math.floor(virt))
    g = virt - lo
    lo = min(max(lo, 0), n - 1)
    hi = min(lo + 1, n - 1)
    part = np.partition(buf, (lo, hi), axis=-1)
    a, b = part[..., lo], part[..., hi]
    diff = b - a
    lerp = a + diff * g
    if g >= 0.5:
        lerp = b - diff * (1 - g)
    return np.where(diff == 0, a, lerp)


class BurstsOracle:
    """features/bursts.py:60-298.

    ``faithful=True`` reproduces the reference bit for bit, including the in-place partition
    of the history buffer (exact ring semantics only while the buffer is not full);
    ``faithful=False`` (default) is the "fixed oracle": same code with the buffer copied
    before the quantile, i.e. a true ring of the last ``time_duration_s``.
    """

    def __init__(self, settings: dict, ch_names, sfreq, faithful: bool = False):
        self.cfg = settings["bursts_settings"]
        self.faithful = faithful
        self.sfreq = sfreq
        self.ch_names = list(ch_names)
        bands = bands_of(settings)
        self.bands = [b.replace(" ", "_") for b in self.cfg["frequency_bands"]]
        for b in self.bands:
            if b not in bands:
                raise ValueError(f"bursting {b} needs to be defined in settings['frequency_ranges_hz']")
        self.seg_s = settings["segment_length_features_ms"] / 1000
        self.samples_overlap = int(sfreq * self.seg_s / settings["sampling_rate_features_hz"])
        self.bank = design_bank([bands[b] for b in self.bands], sfreq, filter_length=sfreq - 1)
        self.n_ring = int(sfreq * self.cfg["time_duration_s"])
        self.buf = np.empty((len(self.ch_names), len(self.bands), 0))
        self.batch = 0
        self.feats = _enabled(self.cfg["burst_features"], BURST_FEAT_ORDER)

    def envelope(self, data: np.ndarray) -> np.ndarray:
        from scipy.signal import hilbert

        return np.abs(np.array(hilbert(apply_bank(data, self.bank))))

    def calc(self, data: np.ndarray) -> dict:
        env = self.envelope(data)
        take = env.shape[-1] if self.batch == 0 else self.samples_overlap
        self.batch += 1
        self.buf = np.concatenate((self.buf, env[:, :, -take:]), axis=2)[:, :, -self.n_ring:]
        q = self.cfg["threshold"] / 100
        if self.faithful:
            from numpy.lib._function_base_impl import _quantile

            thr = _quantile(self.buf, q)
        else:
            thr = quantile_linear(self.buf, q)
        return self.features_from(env, thr)

    def features_from(self, env: np.ndarray, thr: np.ndarray) -> dict:
        nc, nb, w = env.shape
        b = env >= thr[:, :, None]
        res = {}
        for ci, ch in enumerate(self.ch_names):
            for bi, band in enumerate(self.bands):
                row, e = b[ci, bi], env[ci, bi]
                edges = np.flatnonzero(np.diff(np.concatenate(([False], row, [False])).astype(np.int8)))
                starts, ends = edges[0::2], edges[1::2]  # [start, end)
                n_trans = int(np.sum(np.diff(row, prepend=False)))
                num = n_trans // 2
                valid = ends < w  # runs that do not touch the last sample
                dur_mean = (row.sum() / num / self.sfreq) if num != 0 else 0.0
                lens = (ends - starts)[valid]
                dur_max = (lens.max() / self.sfreq) if lens.size else 0.0
                amp_max = float((e * row).max())
                means = [e[s:t].sum() / (t - s) for s, t in zip(starts[valid], ends[valid])]
                amp_mean = float(np.mean(means)) if len(means) else 0.0
                for ft in self.feats:
                    if ft == "duration":
                        res[f"{ch}_bursts_{band}_duration_mean"] = dur_mean
                        res[f"{ch}_bursts_{band}_duration_max"] = dur_max
                    elif ft == "amplitude":
                        res[f"{ch}_bursts_{band}_amplitude_mean"] = amp_mean
                        res[f"{ch}_bursts_{band}_amplitude_max"] = amp_max
                    elif ft == "burst_rate_per_s":
                        res[f"{ch}_bursts_{band}_burst_rate_per_s"] = dur_mean / self.seg_s
                    elif ft == "in_burst":
                        res[f"{ch}_bursts_{band}_in_burst"] = bool(row[-1])
        return res


# ----------------------------------------------------------------------------- a15: sharp waves
_SW_EST = {"mean": np.mean, "median": np.median, "max": np.max, "min": np.min, "var": np.var}


def design_sharpwave_filters(settings: dict, sfreq: float):
    """features/sharpwaves.py:117-142 -> list of (name, taps)."""
    out = []
    for fr in settings["sharpwave_analysis_settings"]["filter_ranges_hz"]:
        lo, hi = _band_tuple(fr)
        assert hi < sfreq
        out.append((f"range_{lo:.0f}_{hi:.0f}", mne_filter.create_filter(None, sfreq, l_freq=lo, h_freq=hi)))
    return out


def waveform_features(x: np.ndarray, sw: dict, sfreq: float) -> dict:
    """features/sharpwaves.py:330-465 for one filtered row of one polarity."""
    from scipy.signal import find_peaks

    feats = sw["sharpwave_features"]
    peaks = find_peaks(x, distance=sw["detect_troughs"]["distance_peaks_ms"])[0]
    troughs = find_peaks(-x, distance=sw["detect_troughs"]["distance_troughs_ms"])[0]
    ptr = first_valid = last_valid = 0
    left, right = [], []
    for i in range(len(troughs)):
        while ptr < peaks.size and peaks[ptr] < troughs[i]:
            ptr += 1
        if ptr - 1 < 0:
            first_valid = i + 1
            continue
        if ptr == peaks.size:
            continue
        last_valid = i
        left.append(peaks[ptr - 1])
        right.append(peaks[ptr])
    troughs = troughs[first_valid : last_valid + 1]
    left = np.array(left, dtype=int)
    right = np.array(right, dtype=int)
    res: dict = {}
    need_prom = feats["prominence"]
    if feats["peak_left"] or need_prom:
        res["peak_left"] = x[left]
    if feats["peak_right"] or need_prom:
        res["peak_right"] = x[right]
    if feats["trough"] or need_prom:
        res["trough"] = x[troughs]
    ms = 1000 / sfreq
    if feats["interval"]:
        res["interval"] = np.concatenate((np.zeros(1), np.diff(troughs))) * ms
    if feats["sharpness"]:
        off = int(5 * ms)
        tv = troughs[np.logical_and(troughs - off > 0, troughs + off < x.shape[0])]
        res["sharpness"] = x[tv] - 0.5 * (x[tv - off] + x[tv + off])
    if feats["num_peaks"]:
        res["num_peaks"] = [troughs.shape[0]]
    need_rise = feats["rise_steepness"] or feats["slope_ratio"]
    need_decay = feats["decay_steepness"] or feats["slope_ratio"]
    if need_rise or need_decay:
        steep = np.concatenate((np.zeros(1), np.diff(x)))
        rise = np.array([np.max(np.abs(steep[l : t + 1])) for l, t in zip(left, troughs)])
        decay = np.array([np.max(np.abs(steep[t : r + 1])) for t, r in zip(troughs, right)])
        if need_rise:
            res["rise_steepness"] = rise
        if need_decay:
            res["decay_steepness"] = decay
        if feats["slope_ratio"]:
            res["slope_ratio"] = rise - decay
    if need_prom:
        res["prominence"] = np.abs((res["peak_right"] + res["peak_left"]) / 2 - res["trough"])
    if feats["decay_time"]:
        res["decay_time"] = (left - troughs) * ms
    if feats["rise_time"]:
        res["rise_time"] = (right - troughs) * ms
    if feats["width"]:
        res["width"] = right - left
    return res


class SharpwaveOracle:
    """features/sharpwaves.py:100-328."""

    def __init__(self, settings: dict, ch_names, sfreq):
        self.sw = settings["sharpwave_analysis_settings"]
        self.sfreq = sfreq
        self.ch_names = list(ch_names)
        self.filters = design_sharpwave_filters(settings, sfreq)
        self.used = _enabled(self.sw["sharpwave_features"], SW_FEAT_ORDER)
        est = self.sw["estimator"]
        for ft in self.used:
            assert any(ft in est[e] for e in SW_EST_ORDER), f"Add estimator key for {ft}"
        est_of = {}
        for e in SW_EST_ORDER:
            for ft in est[e]:
                est_of.setdefault(ft, [])
        for ft in est_of:
            est_of[ft] = [e for e in SW_EST_ORDER if ft in est[e]]
        self.combos = [(ft, e) for ft in self.used for e in est_of[ft]]

    def filtered(self, data: np.ndarray) -> np.ndarray:
        from scipy.signal import fftconvolve

        rows = [np.stack([fftconvolve(data[c], h, mode="same") for _, h in self.filters]) for c in range(data.shape[0])]
        return np.stack(rows)

    def calc(self, data: np.ndarray) -> dict:
        y = self.filtered(data)
        per_key: "OrderedDict[str, dict]" = OrderedDict()
        key_est: "OrderedDict[str, str]" = OrderedDict()
        passes = []
        if self.sw["detect_peaks"]["estimate"]:
            passes.append(("Peak", 1.0))
        if self.sw["detect_troughs"]["estimate"]:
            passes.append(("Trough", -1.0))
        for ci, ch in enumerate(self.ch_names):
            for fi, (fname, _) in enumerate(self.filters):
                for ft, e in self.combos:
                    key_est[f"{ch}_Sharpwave_{e.title()}_{ft}_{fname}"] = e
                for pname, sign in passes:
                    wf = waveform_features(sign * y[ci, fi], self.sw, self.sfreq)
                    for ft, e in self.combos:
                        if ft == "num_peaks":
                            per_key.setdefault(f"{ch}_Sharpwave_num_peaks_{fname}", {})[pname] = wf[ft][0]
                            continue
                        arr = wf[ft]
                        val = _SW_EST[e](arr) if len(arr) != 0 else 0
                        per_key.setdefault(f"{ch}_Sharpwave_{e.title()}_{ft}_{fname}", {})[pname] = val
        out = {}
        if self.sw["apply_estimator_between_peaks_and_troughs"]:
            for key, e in key_est.items():
                vals = list(per_key.get(key, {}).values())
                if len(vals) == 0:
                    continue
                out[key] = _SW_EST[e]([vals[0], vals[1]])
            if self.sw["sharpwave_features"]["num_peaks"]:
                for ch in self.ch_names:
                    for fname, _ in self.filters:
                        k = f"{ch}_Sharpwave_num_peaks_{fname}"
                        out[k] = np.mean([per_key[k]["Peak"], per_key[k]["Trough"]])
        else:
            for key, sub in per_key.items():
                for pname, v in sub.items():
                    out[key + "_analyze_" + pname] = v
        return out


# ----------------------------------------------------------------------------- 8f-1: feature normaliser
def sklearn_restated(method: str, hist: np.ndarray, v: np.ndarray) -> np.ndarray:
    """processing/normalization.py:173-190 (`norm_sklearn`): `transformer.fit(nan_to_num(previous)).transform(current)` for the
    transformers of lines 58-70, restated from scikit-learn 1.9 `sklearn/preprocessing/_data.py` (scikit-learn is a third-party
    dependency of the reference; pinned here by fixtures the unmodified reference generated WITH the real scikit-learn):
    MinMaxScaler (`X * scale_ + min_`), RobustScaler (median, 25-75 percentile range), QuantileTransformer(n_quantiles=300,
    uniform output: mean of the ascending and the descending `np.interp`, bounds applied afterwards).  Scales below 10 * eps
    become 1 (`_handle_zeros_in_scale`)."""
    tiny = 10 * np.finfo(np.float64).eps
    v = np.asarray(v, dtype=np.float64)
    if method == "minmax":
        mn, mx = hist.min(axis=0), hist.max(axis=0)
        rng = mx - mn
        rng[rng < tiny] = 1.0
        scale = 1.0 / rng
        return v * scale + (0.0 - mn * scale)
    if method == "robust":
        centre = np.nanmedian(hist, axis=0)
        q = np.nanpercentile(hist, (25.0, 75.0), axis=0)
        scale = q[1] - q[0]
        scale[scale < tiny] = 1.0
        return (v - centre) / scale
    n = hist.shape[0]
    nq = max(1, min(300, n))
    refs = np.linspace(0, 1, nq, endpoint=True)
    quant = np.maximum.accumulate(np.nanpercentile(hist, refs * 100, axis=0))
    out = np.array(v, dtype=np.float64)
    for j in range(out.size):
        qj = quant[:, j]
        x = out[j]
        lower, upper = x == qj[0], x == qj[-1]
        if not np.isnan(x):
            out[j] = 0.5 * (np.interp(x, qj, refs) - np.interp(-x, -qj[::-1], -refs[::-1]))
        if upper:
            out[j] = 1.0
        if lower:
            out[j] = 0.0
    return out


class FeatureNormalizerOracle:
    """processing/normalization.py:81-111,151-190 ('feature' type; numpy methods and the restated scikit-learn transformers)."""

    def __init__(self, settings: dict):
        cfg = settings["feature_normalization_settings"]
        self.method = cfg["normalization_method"]
        self.clip = cfg["clip"]
        self.n_keep = int(cfg["normalization_time_s"] * settings["sampling_rate_features_hz"])
        self.prev = np.empty((0, 0))

    def process(self, v: np.ndarray) -> np.ndarray:
        if self.prev.size == 0:
            self.prev = v
            return v
        self.prev = np.vstack((self.prev, v))  # data[-0:] == whole vector
        has_nan = np.any(np.isnan(sum(self.prev)))
        mean = (np.nanmean if has_nan else np.mean)(self.prev, axis=0)
        with np.errstate(all="ignore"):
            if self.method == "mean":
                out = (v - mean) / mean
            elif self.method == "median":
                med = (np.nanmedian if has_nan else np.median)(self.prev, axis=0)
                out = (v - med) / med
            elif self.method in ("zscore", "zscore-median"):
                std = (np.nanstd if has_nan else np.std)(self.prev, axis=0)
                std[std == 0] = 1
                centre = mean if self.method == "zscore" else (np.nanmedian if has_nan else np.median)(self.prev, axis=0)
                out = (v - centre) / std
            elif self.method in ("minmax", "robust", "quantile"):
                out = sklearn_restated(self.method, np.nan_to_num(self.prev), v)
            else:
                raise NotImplementedError(f"sklearn normaliser '{self.method}' is out of scope")
        if self.clip:
            out = out.clip(min=-self.clip, max=self.clip)
        self.prev = self.prev[-self.n_keep + 1:]
        return np.nan_to_num(out)


# ----------------------------------------------------------------------------- 8f-3: preprocessing filter, raw normaliser
PREFILTER_ORDER = ["bandstop_filter", "bandpass_filter", "lowpass_filter", "highpass_filter"]


def design_prefilters(settings: dict, sfreq: float) -> list[np.ndarray]:
    """processing/filter_preprocessing.py:44-77: one single-filter MNEFilter per enabled stage, filter_length = sfreq - 1;
    the band stages first (in FilterSettings field order), then low-pass, then high-pass.  Note that the reference hands
    the 'bandstop' range to create_filter as (l_freq, h_freq) = (low, high), i.e. it designs a band-PASS there."""
    cfg = settings["preprocessing_filter"]
    enabled = [k for k in PREFILTER_ORDER if cfg.get(k) is True]
    ranges = []
    for name in enabled:
        if name in ("bandstop_filter", "bandpass_filter"):
            ranges.append(_band_tuple(cfg[name + "_settings"]))
    if "lowpass_filter" in enabled:
        ranges.append((None, cfg["lowpass_filter_cutoff_hz"]))
    if "highpass_filter" in enabled:
        ranges.append((cfg["highpass_filter_cutoff_hz"], None))
    return [design_bank([r], sfreq, filter_length=int(sfreq - 1)) for r in ranges]


def apply_prefilters(x: np.ndarray, banks: list[np.ndarray]) -> np.ndarray:
    """processing/filter_preprocessing.py:79-94: the stages are applied one after the other ('same' FFT convolution)."""
    for b in banks:
        x = apply_bank(x, b)[:, 0, :]
    return x


class RawNormalizerOracle:
    """processing/normalization.py:30-111 ('raw' type; numpy methods only): window 0 passes through and seeds the history;
    window k >= 1 appends its last int(sfreq / rate) samples, is normalised against the WHOLE history (its own new
    samples included), clipped, and the history is trimmed to normalization_time_s * sfreq - 1 samples."""

    def __init__(self, settings: dict, sfreq: float):
        cfg = settings["raw_normalization_settings"]
        self.method = cfg["normalization_method"]
        self.clip = cfg["clip"]
        self.add = int(sfreq / settings["sampling_rate_features_hz"])
        self.n_keep = int(cfg["normalization_time_s"] * sfreq)
        self.prev = np.empty((0, 0))

    def process(self, data: np.ndarray) -> np.ndarray:
        if self.prev.size == 0:
            self.prev = data.T
            return data
        d = data.T
        self.prev = np.vstack((self.prev, d[-self.add:]))
        has_nan = np.any(np.isnan(sum(self.prev)))
        mean = (np.nanmean if has_nan else np.mean)(self.prev, axis=0)
        with np.errstate(all="ignore"):
            if self.method == "mean":
                out = (d - mean) / mean
            elif self.method == "median":
                med = (np.nanmedian if has_nan else np.median)(self.prev, axis=0)
                out = (d - med) / med
            elif self.method in ("zscore", "zscore-median"):
                std = (np.nanstd if has_nan else np.std)(self.prev, axis=0)
                std[std == 0] = 1
                centre = mean if self.method == "zscore" else (np.nanmedian if has_nan else np.median)(self.prev, axis=0)
                out = (d - centre) / std
            elif self.method in ("minmax", "robust"):
                out = sklearn_restated(self.method, np.nan_to_num(self.prev), d)  # (samples x channels: column-wise like sklearn)
            else:
                raise NotImplementedError(f"sklearn normaliser '{self.method}' is out of scope")
        if self.clip:
            out = out.clip(min=-self.clip, max=self.clip)
        self.prev = self.prev[-self.n_keep + 1:]
        return np.nan_to_num(out).T


# ----------------------------------------------------------------------------- a2/a3/a6: window processor
class WindowOracle:
    """stream/data_processor.py:19-90,238-311 + processing/data_preprocessor.py:21-84 +
    features/feature_processor.py:31-84 for the in-scope preprocessors and features."""

    IN_SCOPE_FEATURES = {"raw_hjorth", "return_raw", "bandpass_filter", "stft", "fft", "welch",
                         "sharpwave_analysis", "bursts", "linelength"}

    def __init__(self, sfreq: float, settings: dict, channels: dict | None = None, n_channels: int | None = None,
                 line_noise: float | None = 50, faithful_bursts: bool = False):
        self.settings = settings
        self.sfreq = sfreq // 1
        self.ch = channels if channels is not None else default_channels(n_channels)
        self.names, self.feature_idx = channel_info(self.ch)
        pre = [p for p in PREPROC_ORDER if p in settings["preprocessing"]]
        self.notch = None
        self.ref = None
        self.prefilters = None
        self.rawnorm = None
        self.resample_up = None
        self.pre = []
        for p in pre:
            if p == "preprocessing_filter":
                self.prefilters = design_prefilters(settings, self.sfreq)
                self.pre.append(p)
            elif p == "raw_normalization":
                self.rawnorm = RawNormalizerOracle(settings, self.sfreq)
                self.pre.append(p)
            elif p == "notch_filter":
                self.notch = design_notch(self.sfreq, line_noise)
                self.pre.append(p)
            elif p == "re_referencing":
                self.ref = reref_matrix(self.ch)
                self.pre.append(p)
            elif p == "raw_resampling":
                # processing/resample.py:28-60: ratio 1 is the identity; everything downstream keeps the ORIGINAL sfreq
                # (stream/data_processor.py:55,77-81 never update sfreq_raw) -- reproduced, not repaired
                ratio = float(settings["raw_resampling_settings"]["resample_freq_hz"] / self.sfreq)
                if ratio != 1.0:
                    self.resample_up = ratio
                    self.pre.append(p)
            else:
                raise NotImplementedError(f"unknown preprocessor {p}")
        self.plugins = []
        for f in _enabled(settings["features"], FEATURE_ORDER):
            if f not in self.IN_SCOPE_FEATURES:
                raise NotImplementedError(f"feature {f} is out of scope (SURVEY.md section 2 row 23)")
            self.plugins.append((f, self._make(f, faithful_bursts)))
        self.norm = FeatureNormalizerOracle(settings) if settings["postprocessing"]["feature_normalization"] else None
        self.non_psd = None

    def _make(self, f: str, faithful_bursts: bool):
        s, n, fs = self.settings, self.names, self.sfreq
        if f == "raw_hjorth":
            return lambda d: hjorth(d, n)
        if f == "return_raw":
            return lambda d: raw_last(d, n)
        if f == "linelength":
            return lambda d: linelength(d, n)
        if f in ("fft", "welch", "stft"):
            return OscOracle(f, s, n, fs).calc
        if f == "bandpass_filter":
            return BandPowerOracle(s, n, fs).calc
        if f == "bursts":
            return BurstsOracle(s, n, fs, faithful=faithful_bursts).calc
        return SharpwaveOracle(s, n, fs).calc

    def preprocess(self, data: np.ndarray) -> np.ndarray:
        d = np.nan_to_num(data)[self.feature_idx, :]
        for p in self.pre:
            if p == "preprocessing_filter":
                d = apply_prefilters(d, self.prefilters)
            elif p == "notch_filter":
                d = apply_notch(d, self.notch)
            elif p == "raw_normalization":
                d = self.rawnorm.process(d)
            elif p == "raw_resampling":
                from oracle.mne_filter_restated import resample

                d = resample(d.astype(np.float64), up=self.resample_up, down=1.0)
            elif self.ref is not None:
                d = self.ref @ d
        return d

    def process(self, data: np.ndarray) -> dict:
        nan_ch = np.isnan(data).any(axis=1)
        d = self.preprocess(data)
        feats: dict = {}
        for _, fn in self.plugins:
            feats.update(fn(d))
        if self.norm is not None:
            vals = np.fromiter(feats.values(), dtype=np.float64)
            if not self.settings["feature_normalization_settings"]["normalize_psd"]:
                if self.non_psd is None:
                    self.non_psd = [i for i, k in enumerate(feats) if "psd" not in k]
                    self.psd = sorted(set(range(len(feats))) - set(self.non_psd))
                normed = np.empty(vals.shape[0])
                normed[self.non_psd] = self.norm.process(vals[self.non_psd])
                normed[self.psd] = vals[self.psd]
            else:
                normed = self.norm.process(vals)
            feats = {k: normed[i] for i, k in enumerate(feats)}
        if nan_ch.sum() > 0:
            hit = [k for ch in list(np.array(self.names)[nan_ch]) for k in feats if ch in k]
            for k in hit:
                feats[k] = np.nan
        return feats


def run_offline(data: np.ndarray, sfreq: float, settings: dict, channels: dict | None = None,
                line_noise: float | None = 50, faithful_bursts: bool = False, max_windows: int | None = None):
    """stream/stream.py:198-345 without the file writer: returns (column names, (n_windows, F+1) matrix)."""
    proc = WindowOracle(sfreq, settings, channels, n_channels=data.shape[0], line_noise=line_noise,
                        faithful_bursts=faithful_bursts)
    ch = proc.ch
    targets = [i for i in range(len(ch["name"])) if ch["target"][i] == 1]
    rows, cols = [], None
    grid = window_grid(data.shape[1], sfreq, settings["sampling_rate_features_hz"], settings["segment_length_features_ms"])
    for wi, (i0, i1, t_ms) in enumerate(grid):
        if max_windows is not None and wi >= max_windows:
            break
        win = data[:, i0:i1]
        f = proc.process(win)
        f["time"] = t_ms
        for ti in targets:
            f[ch["name"][ti]] = win[ti, -1]
        if cols is None:
            cols = list(f.keys())
        rows.append([float(f[k]) for k in cols])
    return cols, np.array(rows, dtype=np.float64)
