"""Generate the committed golden fixtures from the UNMODIFIED reference.

Run in the authoring container only (needs ``/root/reference``):

    python tests/golden/make_golden.py

Every case stores (a) the seeded input, rounded once to float32 and up-cast to float64 as
SURVEY.md section 8d prescribes, (b) the settings as ``NMSettings.model_dump()`` JSON, (c) the
ordered output keys and (d) the float64 values produced by the reference files executed
through ``oracle/ref_shim.py`` (``mne.filter`` = ``oracle/mne_filter_restated.py``; MNE is
not installed -> that part is "parity unpinned", see the oracle header).
Fixtures are small ``.npz`` files under ``tests/golden/``.
"""

from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
OUT = Path(__file__).resolve().parent

from oracle.ref_shim import load_reference, load_reference_stream  # noqa: E402


def f32(x):
    return np.asarray(x, dtype=np.float64).astype(np.float32).astype(np.float64)


def uniform(seed, c, t):
    return f32(np.random.default_rng(seed).random((c, t)))


def neural_like(seed, c, t, sfreq=1000.0):
    """pink-ish noise + 6/20 Hz bursts + 50 Hz line."""
    rng = np.random.default_rng(seed)
    tt = np.arange(t) / sfreq
    x = np.cumsum(rng.standard_normal((c, t)), axis=1) * 0.01 + rng.standard_normal((c, t)) * 0.05
    gate = ((tt % 1.0) < 0.3).astype(float)
    for ci in range(c):
        ph = rng.random() * 2 * np.pi
        x[ci] += gate * np.sin(2 * np.pi * 20 * tt + ph) * (0.5 + 0.5 * rng.random())
        x[ci] += np.roll(gate, int(0.5 * sfreq)) * np.sin(2 * np.pi * 6 * tt + ph)
        x[ci] += 0.5 * np.sin(2 * np.pi * 50 * tt + ci)
    return f32(x)


def dump_settings(s) -> str:
    return json.dumps(s.model_dump())


def save(name, **arrs):
    np.savez_compressed(OUT / f"{name}.npz", **arrs)
    print("wrote", name, {k: (v.shape if hasattr(v, "shape") else None) for k, v in arrs.items()})


def dict_to_arrays(d):
    keys = list(d.keys())
    return json.dumps(keys), np.array([float(d[k]) for k in keys], dtype=np.float64)


def plugin_cases(nm):
    F = nm.features
    ch = ["c0", "c1", "c2", "c3"]
    variants = {}

    s = nm.NMSettings.get_default()
    variants["default"] = s

    s = nm.NMSettings.get_default()
    for osc in ("fft_settings", "welch_settings", "stft_settings"):
        for est in ("mean", "median", "std", "max"):
            s[osc].features[est] = True
    s.fft_settings.return_spectrum = True
    s.stft_settings.return_spectrum = True
    s.welch_settings.return_spectrum = True
    s.bandpass_filter_settings.bandpower_features.mobility = True
    s.bandpass_filter_settings.bandpower_features.complexity = True
    sw = s.sharpwave_analysis_settings
    for ft in list(sw.sharpwave_features.model_fields.keys()):
        sw.sharpwave_features[ft] = True
    sw.estimator.mean = ["interval", "num_peaks", "width", "rise_time", "peak_left"]
    sw.estimator.median = ["decay_time", "trough", "interval"]
    sw.estimator.max = ["prominence", "sharpness", "rise_steepness", "peak_right"]
    sw.estimator.min = ["decay_steepness", "sharpness"]
    sw.estimator.var = ["slope_ratio", "prominence"]
    variants["allest"] = s

    s = nm.NMSettings.get_default()
    s.frequency_ranges_hz = {
        "theta": [4, 8], "alpha": [8, 12], "low beta": [13, 20], "high beta": [20, 35],
        "low gamma": [60, 80], "high gamma": [90, 200], "HFA": [200, 400],
    }
    s.fft_settings.log_transform = False
    s.welch_settings.log_transform = False
    s.stft_settings.log_transform = False
    s.bandpass_filter_settings.log_transform = False
    s.bursts_settings.frequency_bands = ["low_beta", "high_beta", "low_gamma"]
    s.bursts_settings.threshold = 60
    s = s.validate()
    variants["sevenbands_nolog"] = s

    classes = {
        "fft": F.FFT, "welch": F.Welch, "stft": F.STFT, "hjorth": F.Hjorth, "raw": F.Raw,
        "linelength": F.LineLength, "bandpower": F.BandPower, "bursts": F.Bursts, "sharpwave": F.SharpwaveAnalyzer,
    }
    inputs = {"uniform": uniform(0, 4, 1000), "neural": neural_like(1, 4, 1000)}
    for vname, st in variants.items():
        for iname, x in inputs.items():
            arrs = {"x": x.astype(np.float32), "settings": dump_settings(st), "ch_names": json.dumps(ch), "sfreq": 1000.0}
            for cname, cls in classes.items():
                obj = cls(st, ch, 1000)
                k, v = dict_to_arrays(obj.calc_feature(x.copy()))
                arrs[f"{cname}_keys"], arrs[f"{cname}_vals"] = k, v
            save(f"plugins_{vname}_{iname}", **arrs)

    # sfreq = 2000 (taps 1999 / 3301), window 2000
    st = nm.NMSettings.get_default()
    x = neural_like(2, 3, 2000, 2000.0)
    ch3 = ["a", "b", "c"]
    arrs = {"x": x.astype(np.float32), "settings": dump_settings(st), "ch_names": json.dumps(ch3), "sfreq": 2000.0}
    for cname, cls in classes.items():
        obj = cls(st, ch3, 2000)
        k, v = dict_to_arrays(obj.calc_feature(x.copy()))
        arrs[f"{cname}_keys"], arrs[f"{cname}_vals"] = k, v
    save("plugins_default_neural_2k", **arrs)


def preprocess_cases(nm):
    import pandas as pd

    x = neural_like(3, 6, 1000)
    ch = nm.utils.channels.get_default_channels_from_data(x)
    ch.loc[1, "rereference"] = "ch0"
    ch.loc[2, "rereference"] = "ch0&ch3"
    ch.loc[4, "rereference"] = "None"
    ch.loc[5, "status"] = "bad"
    ch.loc[5, "used"] = 0
    notch = nm.filter.NotchFilter(1000, 50)
    y = notch.process(x.copy())
    rr = nm.processing.ReReferencer(1000, ch)
    used = x[:5]
    z = rr.process(notch.process(used.copy()))
    save("preprocess_notch_reref", x=x.astype(np.float32), notch_taps=notch.filter_bank, notch_out=y,
         channels=ch.to_json(), ref_matrix=rr.ref_matrix, reref_out=z)
    for sf in (150, 200, 500, 2000):
        nf = nm.filter.NotchFilter(sf, 50)
        xs = uniform(sf, 2, int(sf))
        save(f"notch_sf{sf}", x=xs.astype(np.float32), taps=nf.filter_bank, out=nf.process(xs.copy()), sfreq=float(sf))
    # short window (W < filter length) and long window
    nf = nm.filter.NotchFilter(1000, 60)
    for w in (370, 2500):
        xs = uniform(w, 2, w)
        save(f"notch_w{w}", x=xs.astype(np.float32), taps=nf.filter_bank, out=nf.process(xs.copy()), sfreq=1000.0)


def window_processor_cases(nm):
    # default settings (incl. feature normalisation), neural-like, 5 channels, 40 windows
    for name, make, nwin in (("default", nm.NMSettings.get_default, 40), ("fast", nm.NMSettings.get_fast_compute, 40)):
        st = make()
        x = neural_like(4, 5, 1000 + 100 * (nwin - 1))
        ch = nm.utils.channels.get_default_channels_from_data(x)
        dp = nm.DataProcessor(sfreq=1000, settings=st, channels=ch, line_noise=50, verbose=False)
        gen = nm.RawDataGenerator(x, 1000, st.sampling_rate_features_hz, st.segment_length_features_ms)
        rows, keys = [], None
        for _, batch in gen:
            d = dp.process(batch)
            if keys is None:
                keys = list(d.keys())
            rows.append([float(d[k]) for k in keys])
        save(f"dataprocessor_{name}", x=x.astype(np.float32), settings=dump_settings(st), keys=json.dumps(keys),
             vals=np.array(rows), sfreq=1000.0)

    # C3-like feature set, no normalisation, with a NaN channel span and a target channel
    st = nm.NMSettings.get_default().reset()
    st.features.fft = True
    st.features.bandpass_filter = True
    st.features.raw_hjorth = True
    st.features.linelength = True
    st.features.return_raw = True
    st.postprocessing.feature_normalization = False
    x = uniform(5, 4, 3000)
    x[1, 1500:1600] = np.nan
    ch = nm.utils.channels.get_default_channels_from_data(x)
    dp = nm.DataProcessor(sfreq=1000, settings=st, channels=ch, line_noise=50, verbose=False)
    gen = nm.RawDataGenerator(x, 1000, st.sampling_rate_features_hz, st.segment_length_features_ms)
    rows, keys = [], None
    for _, batch in gen:
        d = dp.process(batch)
        if keys is None:
            keys = list(d.keys())
        rows.append([float(d[k]) for k in keys])
    save("dataprocessor_c3_nan", x=x.astype(np.float32), settings=dump_settings(st), keys=json.dumps(keys),
         vals=np.array(rows), sfreq=1000.0)


def _run_windows(nm, st, x, line_noise=50, sfreq=1000):
    ch = nm.utils.channels.get_default_channels_from_data(x)
    dp = nm.DataProcessor(sfreq=sfreq, settings=st, channels=ch, line_noise=line_noise, verbose=False)
    gen = nm.RawDataGenerator(x, sfreq, st.sampling_rate_features_hz, st.segment_length_features_ms)
    rows, keys = [], None
    for _, batch in gen:
        d = dp.process(batch)
        if keys is None:
            keys = list(d.keys())
        rows.append([float(d[k]) for k in keys])
    return keys, np.array(rows)


def next_row_cases(nm):
    """PreprocessingFilter (processing/filter_preprocessing.py) and RawNormalizer
    (processing/normalization.py) in front of a small feature set, unmodified reference through the shim."""
    def base():
        st = nm.NMSettings.get_default().reset()
        for f in ("fft", "raw_hjorth", "linelength", "return_raw"):
            st.features[f] = True
        st.postprocessing.feature_normalization = False
        return st

    x = neural_like(21, 4, 1000 + 100 * 24)
    st = base()
    st.preprocessing = ["preprocessing_filter", "notch_filter", "re_referencing"]  # default FilterSettings: all four stages
    keys, vals = _run_windows(nm, st, x)
    save("dataprocessor_prefilter_default", x=x.astype(np.float32), settings=dump_settings(st), keys=json.dumps(keys), vals=vals, sfreq=1000.0)

    st = base()
    st.preprocessing = ["preprocessing_filter", "re_referencing"]
    st.preprocessing_filter.bandstop_filter = False
    st.preprocessing_filter.bandpass_filter = False
    st.preprocessing_filter.lowpass_filter_cutoff_hz = 90
    st.preprocessing_filter.highpass_filter_cutoff_hz = 5
    keys, vals = _run_windows(nm, st, x)
    save("dataprocessor_prefilter_lphp", x=x.astype(np.float32), settings=dump_settings(st), keys=json.dumps(keys), vals=vals, sfreq=1000.0)

    x = neural_like(22, 4, 1000 + 100 * 44)
    for method in ("zscore", "mean", "median", "zscore-median"):
        st = base()
        st.preprocessing = ["notch_filter", "re_referencing", "raw_normalization"]
        st.raw_normalization_settings.normalization_time_s = 2.5   # history is trimmed inside the run
        st.raw_normalization_settings.normalization_method = method
        st.raw_normalization_settings.clip = 2.0 if method == "zscore" else 3.0
        keys, vals = _run_windows(nm, st, x)
        save(f"dataprocessor_rawnorm_{method.replace('-', '_')}", x=x.astype(np.float32), settings=dump_settings(st), keys=json.dumps(keys),
             vals=vals, sfreq=1000.0)


def sklearn_norm_cases(nm):
    """Feature normalisation through the scikit-learn transformers the reference wraps (processing/normalization.py:58-70,173-190):
    MinMaxScaler, RobustScaler, QuantileTransformer(n_quantiles=300); history of 30 windows, trimmed inside the run; one channel
    with a flat stretch (constant features -> zero scales)."""
    x = neural_like(31, 4, 1000 + 100 * 44)
    x[2, :2600] = x[2, 0]  # constant samples for the first windows of channel 2
    for method in ("minmax", "robust", "quantile"):
        st = nm.NMSettings.get_default().reset()
        for f in ("fft", "raw_hjorth", "linelength", "return_raw"):
            st.features[f] = True
        st.postprocessing.feature_normalization = True
        st.feature_normalization_settings.normalization_method = method
        st.feature_normalization_settings.normalization_time_s = 3.0
        st.feature_normalization_settings.clip = 0.8 if method == "robust" else 3.0
        keys, vals = _run_windows(nm, st, x)
        save(f"dataprocessor_featnorm_{method}", x=x.astype(np.float32), settings=dump_settings(st), keys=json.dumps(keys), vals=vals,
             sfreq=1000.0)


def sklearn_rawnorm_cases(nm):
    """RawNormalizer through the scikit-learn MinMaxScaler / RobustScaler (processing/normalization.py:58-70,173-190): history of
    2.5 s of samples, trimmed inside the run."""
    x = neural_like(32, 4, 1000 + 100 * 44)
    for method in ("minmax", "robust"):
        st = nm.NMSettings.get_default().reset()
        for f in ("fft", "raw_hjorth", "linelength", "return_raw"):
            st.features[f] = True
        st.postprocessing.feature_normalization = False
        st.preprocessing = ["notch_filter", "re_referencing", "raw_normalization"]
        st.raw_normalization_settings.normalization_time_s = 2.5
        st.raw_normalization_settings.normalization_method = method
        st.raw_normalization_settings.clip = 1.5 if method == "robust" else 3.0
        keys, vals = _run_windows(nm, st, x)
        save(f"dataprocessor_rawnorm_{method}", x=x.astype(np.float32), settings=dump_settings(st), keys=json.dumps(keys), vals=vals,
             sfreq=1000.0)


def stream_cases():
    nm = load_reference_stream()
    import tempfile

    # README demo: 5 ch x 10 s, rate 3 Hz, default settings
    x = uniform(6, 5, 10000)
    with tempfile.TemporaryDirectory() as td:
        st = nm.Stream(sfreq=1000, data=x, sampling_rate_features_hz=3)
        df = st.run(out_dir=td, experiment_name="demo")
    save("stream_readme_demo", x=x.astype(np.float32), keys=json.dumps(list(df.columns)), vals=df.to_numpy(dtype=np.float64),
         sfreq=1000.0, rate=3.0)

    # float sampling rate / short segment (tests/test_timing.py:43-73): variable window lengths
    fs = 1111.111
    x = uniform(7, 1, int(2 * fs))
    s = nm.NMSettings.get_fast_compute()
    s.segment_length_features_ms = 333
    s.features.fft = False
    s.features.raw_hjorth = True
    s.preprocessing = ["notch_filter", "re_referencing"]  # resampling (ratio != 1) is a "next" row
    with tempfile.TemporaryDirectory() as td:
        st = nm.Stream(sfreq=fs, data=x, sampling_rate_features_hz=200, settings=s)
        df = st.run(out_dir=td, experiment_name="floatfs")
    save("stream_float_fs", x=x.astype(np.float32), settings=dump_settings(st.settings), keys=json.dumps(list(df.columns)),
         vals=df.to_numpy(dtype=np.float64), sfreq=fs, rate=200.0)


def resample_cases(nm):
    """``raw_resampling`` with a ratio != 1 (processing/resample.py:28-60; the reference's DEFAULT preprocessing
    list resamples to 1 kHz).  Unmodified reference through the shim; ``mne.filter.resample`` itself is the restatement in
    ``oracle/mne_filter_restated.py`` (MNE is not installed), so the fixtures pin the reference's handling of the resampled rows --
    every plug-in keeps the ORIGINAL sampling rate (stream/data_processor.py:55,77-81) -- not MNE's arithmetic."""
    # untouched default settings at 2 kHz (C5's configuration): 2000 -> 1000 samples per window, z-scored features
    st = nm.NMSettings.get_default()
    x = neural_like(31, 4, 2000 + 200 * 34, sfreq=2000.0)
    keys, vals = _run_windows(nm, st, x, sfreq=2000)
    save("dataprocessor_resample_2k_default", x=x.astype(np.float32), settings=dump_settings(st), keys=json.dumps(keys), vals=vals, sfreq=2000.0)
    # ratio 0.8 (1250 Hz: padded length 2048 -> 1638, not a power of two), un-normalised, plus band-pass power and STFT
    st = nm.NMSettings.get_default()
    st.postprocessing.feature_normalization = False
    st.features.bandpass_filter = True
    st.features.stft = True
    x = neural_like(32, 3, 1250 + 125 * 30, sfreq=1250.0)
    keys, vals = _run_windows(nm, st, x, sfreq=1250)
    save("dataprocessor_resample_1250", x=x.astype(np.float32), settings=dump_settings(st), keys=json.dumps(keys), vals=vals, sfreq=1250.0)
    # up-sampling (500 Hz -> 1 kHz) with the raw normaliser behind the resampler
    st = nm.NMSettings.get_default().reset()
    for f in ("fft", "raw_hjorth", "linelength", "return_raw", "sharpwave_analysis"):
        st.features[f] = True
    st.postprocessing.feature_normalization = False
    st.preprocessing = ["raw_resampling", "notch_filter", "re_referencing", "raw_normalization"]
    st.raw_normalization_settings.normalization_time_s = 2.0
    x = neural_like(33, 3, 500 + 50 * 24, sfreq=500.0)
    keys, vals = _run_windows(nm, st, x, sfreq=500)
    save("dataprocessor_resample_up_rawnorm", x=x.astype(np.float32), settings=dump_settings(st), keys=json.dumps(keys), vals=vals, sfreq=500.0)


def sharpwave_option_cases(nm):
    """Sharp-wave option the default fixtures do not reach (features/sharpwaves.py:269-272): a single polarity with the un-paired
    key form.  (Filters of different lengths are not a fixture: that branch of the reference convolves across channels.)"""
    x = neural_like(41, 3, 1000 + 100 * 7)
    for pol in ("peaks", "troughs"):
        st = nm.NMSettings.get_default().reset()
        st.features.sharpwave_analysis = True
        st.postprocessing.feature_normalization = False
        st.sharpwave_analysis_settings.apply_estimator_between_peaks_and_troughs = False
        (st.sharpwave_analysis_settings.detect_troughs if pol == "peaks" else st.sharpwave_analysis_settings.detect_peaks).estimate = False
        keys, vals = _run_windows(nm, st, x)
        save(f"dataprocessor_sharpwave_{pol}_only", x=x.astype(np.float32), settings=dump_settings(st), keys=json.dumps(keys), vals=vals, sfreq=1000.0)


def burst_history_case(nm):
    """Bursts across 320 windows (history overflows after window 290): faithful reference values."""
    st = nm.NMSettings.get_default()
    ch = ["u", "v"]
    nwin = 320
    x = neural_like(8, 2, 1000 + 100 * (nwin - 1))
    b = nm.features.Bursts(st, ch, 1000)
    rows, keys = [], None
    for k in range(nwin):
        d = b.calc_feature(x[:, 100 * k : 100 * k + 1000].copy())
        if keys is None:
            keys = list(d.keys())
        rows.append([float(d[kk]) for kk in keys])
    save("bursts_history_320", x=x.astype(np.float32), settings=dump_settings(st), ch_names=json.dumps(ch), keys=json.dumps(keys),
         vals=np.array(rows), sfreq=1000.0)


def real_data_case(nm):
    """6 channels x 6 s excerpt of the reference's bundled BrainVision recording (data, not source)."""
    f = Path("/root/reference/py_neuromodulation/data/sub-testsub/ses-EphysMedOff/ieeg/"
             "sub-testsub_ses-EphysMedOff_task-gripforce_run-0_ieeg.eeg")
    if not f.is_file():
        print("real-data file not found, skipped")
        return
    raw = np.fromfile(f, "<f4").reshape(-1, 10).T * 1e-7
    x = f32(raw[3:9, 2000:8000])  # the six ECoG channels
    st = nm.NMSettings.get_default()
    st.postprocessing.feature_normalization = False
    ch = nm.utils.channels.get_default_channels_from_data(x)
    dp = nm.DataProcessor(sfreq=1000, settings=st, channels=ch, line_noise=60, verbose=False)
    gen = nm.RawDataGenerator(x, 1000, st.sampling_rate_features_hz, st.segment_length_features_ms)
    rows, keys = [], None
    for _, batch in gen:
        d = dp.process(batch)
        if keys is None:
            keys = list(d.keys())
        rows.append([float(d[k]) for k in keys])
    save("dataprocessor_realdata", x=x.astype(np.float32), settings=dump_settings(st), keys=json.dumps(keys),
         vals=np.array(rows), sfreq=1000.0, line_noise=60.0)


if __name__ == "__main__":
    nm = load_reference()
    if len(sys.argv) > 1 and sys.argv[1] == "next":  # only the SURVEY 8f "next row" fixtures
        next_row_cases(nm)
        raise SystemExit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "sklearn_norm":
        sklearn_norm_cases(nm)
        sklearn_rawnorm_cases(nm)
        raise SystemExit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "sharpwave":
        sharpwave_option_cases(nm)
        raise SystemExit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "resample":
        resample_cases(nm)
        raise SystemExit(0)
    plugin_cases(nm)
    preprocess_cases(nm)
    resample_cases(nm)
    sharpwave_option_cases(nm)
    window_processor_cases(nm)
    sklearn_norm_cases(nm)
    sklearn_rawnorm_cases(nm)
    burst_history_case(nm)
    real_data_case(nm)
    stream_cases()
