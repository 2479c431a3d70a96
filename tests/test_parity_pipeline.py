"""Parity of the fused window processor (DataProcessor.process / process_windows) and of Stream.run against
fixtures generated from the unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest

import py_neuromodulation_b200 as nm
from oracle import np_oracle as orc
from py_neuromodulation_b200.stream.generator import window_grid
from py_neuromodulation_b200.utils.channels import get_default_channels_from_data
from tests.helpers import load_golden, uniform, neural_like, parity_err

TOL = 1e-8  # tests/helpers.py::parity_err: pure relative for linear features, absolute for log10 outputs and z-scores


def check_matrix(cols, mat, keys_ref, ref, what, tol=TOL, normalized=True):
    assert list(cols) == list(keys_ref), f"{what}: column order differs"
    assert mat.shape == ref.shape, f"{what}: {mat.shape} vs {ref.shape}"
    assert np.array_equal(np.isnan(mat), np.isnan(ref)), f"{what}: NaN pattern differs"
    inf = np.isinf(ref)
    assert np.array_equal(mat[inf], ref[inf]), f"{what}: inf pattern differs"
    err = parity_err(cols, mat, ref, normalized)
    w, c = np.unravel_index(np.argmax(err), err.shape)
    assert err.max() <= tol, f"{what}: window {w} {cols[c]}: got {mat[w, c]!r} ref {ref[w, c]!r}"
    for j, k in enumerate(cols):
        if k.endswith("_in_burst") and not np.isnan(ref[:, j]).any():
            pass  # checked exactly in the non-normalised cases below


def run_dp(g, n_windows=None, with_norm=True, fused=None):
    x = g["x"].astype(np.float64)
    s = nm.NMSettings(**g["settings"])
    dp = nm.DataProcessor(sfreq=g["sfreq"], settings=s, channels=get_default_channels_from_data(x),
                          line_noise=g.get("line_noise", 50), verbose=False, fused=fused)
    starts, lengths, _ = window_grid(x.shape[1], g["sfreq"], s.sampling_rate_features_hz, s.segment_length_features_ms)
    if n_windows:
        starts = starts[:n_windows]
    cols, mat = dp.process_windows(x, starts, int(lengths[0]))
    return dp, cols, mat, starts


@pytest.mark.parametrize("name,n_emu", [("dataprocessor_c3_nan", None), ("dataprocessor_fast", None), ("dataprocessor_default", 24),
                                        ("dataprocessor_realdata", 12), ("dataprocessor_prefilter_default", None),
                                        ("dataprocessor_prefilter_lphp", None), ("dataprocessor_rawnorm_zscore", None),
                                        ("dataprocessor_rawnorm_mean", None), ("dataprocessor_rawnorm_median", None),
                                        ("dataprocessor_rawnorm_zscore_median", None),
                                        # raw_resampling with a ratio != 1 (processing/resample.py:28-60; stale sampling rate downstream)
                                        ("dataprocessor_resample_2k_default", 8), ("dataprocessor_resample_1250", 10),
                                        ("dataprocessor_resample_up_rawnorm", None),
                                        # sharp waves with one polarity only (features/sharpwaves.py:269-272)
                                        ("dataprocessor_sharpwave_peaks_only", None),
                                        ("dataprocessor_sharpwave_troughs_only", None),
                                        # MinMaxScaler / RobustScaler / QuantileTransformer feature normalisation
                                        # (processing/normalization.py:58-70,173-190), restated in csrc/nm_norm.cuh
                                        ("dataprocessor_featnorm_minmax", None), ("dataprocessor_featnorm_robust", None),
                                        ("dataprocessor_featnorm_quantile", None),
                                        # RawNormalizer through MinMaxScaler / RobustScaler: sliding min / max and 25 / 50 / 75th
                                        # percentiles from the order-statistic kernel of the burst thresholds
                                        ("dataprocessor_rawnorm_minmax", None), ("dataprocessor_rawnorm_robust", None)])
def test_window_processor_matches_reference_golden(backend, name, n_emu):
    g = load_golden(name)
    n = n_emu if backend == "emu" else None  # the thread emulator is slow: fewer windows on CPU, all on the GPU
    _, cols, mat, starts = run_dp(g, n)
    normalized = bool(g["settings"]["postprocessing"]["feature_normalization"]) or "raw_normalization" in g["settings"]["preprocessing"]
    check_matrix(cols, mat, g["keys"], g["vals"][: len(starts)], name, normalized=normalized)
    if name == "dataprocessor_realdata":  # not normalised: integer-valued burst outputs must be bit-exact
        for j, k in enumerate(cols):
            if k.endswith("_in_burst") or k.endswith("_duration_max"):
                assert np.array_equal(mat[:, j], g["vals"][: len(starts), j]), k


@pytest.mark.parametrize("mode", [True, "front"])
@pytest.mark.parametrize("name,n_emu", [("dataprocessor_c3_nan", None), ("dataprocessor_fast", None), ("dataprocessor_default", 12),
                                        ("dataprocessor_realdata", 8)])
def test_fused_window_kernel_matches_reference_golden(backend, name, n_emu, mode):
    """The same fixtures through the bulk-copy staged kernels of csrc/nm_fused.cuh (raw rows through cp.async.bulk + mbarrier,
    re-reference folded into the load): the single persistent kernel (notch -> scan -> DFT band features -> band-pass bank on
    chip) and the front kernel (the default: notch + scan + DFT, bank separate) -- both agree with the staged kernels."""
    g = load_golden(name)
    n = n_emu if backend == "emu" else None
    dp, cols, mat, starts = run_dp(g, n, fused=mode)
    # common average over < 5 channels is handed over as a sparse matrix (no group sum): not foldable into the load, so the
    # staged kernels serve that pipeline whatever was asked for
    kern = "nm_front_kernel" if mode == "front" else "nm_fused_kernel"
    assert (kern in dp.plan(1000).pipe.describe_plan()) == (g["x"].shape[0] >= 5)
    normalized = bool(g["settings"]["postprocessing"]["feature_normalization"])
    check_matrix(cols, mat, g["keys"], g["vals"][: len(starts)], name + " (fused)", normalized=normalized)
    _, _, mat0, _ = run_dp(g, n, fused=False)
    assert np.array_equal(np.isnan(mat), np.isnan(mat0))
    assert parity_err(cols, mat, mat0, normalized).max() < 1e-11  # (same arithmetic, different summation trees)


def test_fused_window_kernel_float64_recording_and_odd_channels(backend):
    """float64 uploads (what reference users pass), an odd channel count and window starts that are not multiples of 4 samples
    (the bulk copies start at the 16-byte aligned sample below the window)."""
    x = neural_like(21, 5, 2600)
    s = nm.NMSettings.get_fast_compute()
    s.features.raw_hjorth = True
    s.features.linelength = True
    s.features.return_raw = True
    s.features.bandpass_filter = True
    s.features.welch = True
    s.postprocessing.feature_normalization = False
    s.sampling_rate_features_hz = 7  # stride 142.86 samples: starts 0, 142, 285, 428, ...
    for dtype, mode in ((np.float64, True), (np.float32, True), (np.float64, "front"), (np.float32, "front")):
        xs = x.astype(dtype)
        dp = nm.DataProcessor(sfreq=1000, settings=s, channels=get_default_channels_from_data(x), line_noise=50, verbose=False, fused=mode)
        starts, lengths, _ = window_grid(x.shape[1], 1000, 7, 1000)
        cols, mat = dp.process_windows(xs, starts, 1000)
        assert ("nm_front_kernel" if mode == "front" else "nm_fused_kernel") in dp.plan(1000).pipe.describe_plan()
        ref_cols, ref = orc.run_offline(xs.astype(np.float64), 1000, s.model_dump())
        assert ref_cols[: len(cols)] == cols
        assert parity_err(cols, mat, ref[:, : len(cols)]).max() < 1e-9


def test_streaming_process_equals_batch(backend):
    """DataProcessor.process window by window (stateful bursts + normaliser) == one batched run."""
    g = load_golden("dataprocessor_default")
    n = 8
    dp, cols, mat, starts = run_dp(g, n)
    x = g["x"].astype(np.float64)
    s = nm.NMSettings(**g["settings"])
    dp2 = nm.DataProcessor(sfreq=1000, settings=s, channels=get_default_channels_from_data(x), line_noise=50, verbose=False)
    for k in range(n):
        d = dp2.process(x[:, starts[k] : starts[k] + 1000])
        assert list(d.keys()) == cols
        v = np.array([float(t) for t in d.values()])
        assert np.max(np.abs(v - mat[k]) / np.maximum(np.abs(mat[k]), 1)) < 1e-12, k
        ref = g["vals"][k]
        assert np.max(np.abs(v - ref) / np.maximum(np.abs(ref), 1)) < TOL, k


def test_stream_readme_demo(backend, tmp_path):
    """README demo of the reference: 5 ch x 10 s @ 1 kHz, 3 Hz features, default settings -> 28 x 156 DataFrame."""
    g = load_golden("stream_readme_demo")
    x = g["x"].astype(np.float64)
    stream = nm.Stream(sfreq=1000, data=x, sampling_rate_features_hz=3)
    df = stream.run(out_dir=tmp_path, experiment_name="demo")
    assert df.shape == (28, 156)
    assert list(df.columns) == g["keys"]
    assert list(df["time"].iloc[:4]) == [1000, 1334, 1667, 2000] and df["time"].iloc[-1] == 10000
    check_matrix(list(df.columns), df.to_numpy(), g["keys"], g["vals"], "readme demo")
    for suffix in ("_FEATURES.csv", "_SETTINGS.yaml", "_SIDECAR.json", "_channels.csv"):
        assert (tmp_path / "demo" / f"demo{suffix}").is_file(), suffix
    import pandas as pd

    back = pd.read_csv(tmp_path / "demo" / "demo_FEATURES.csv")
    assert list(back.columns) == g["keys"] and back.shape == df.shape


def test_stream_untouched_defaults_at_2khz(backend, tmp_path):
    """The reference's default preprocessing list resamples to 1 kHz (default_settings.yaml:46-49): ``nm.Stream(sfreq=2000, data)``
    with untouched settings runs the resampler (processing/resample.py:28-60) and designs every plug-in with the stale 2 kHz rate."""
    g = load_golden("dataprocessor_resample_2k_default")
    x = g["x"].astype(np.float64)
    stream = nm.Stream(sfreq=2000, data=x)
    df = stream.run(out_dir=tmp_path, experiment_name="rs2k", save_csv=False)
    assert list(df.columns) == g["keys"] + ["time"]
    assert df.shape[0] == g["vals"].shape[0]
    plan = stream.data_processor.plan(2000)
    # integer down-sampling runs as an FFT convolution with an ideal low-pass + decimating store, not as the dense GEMM
    assert "resampler: nm_convx_kernel 2000 -> 1000" in plan.pipe.describe_plan()
    check_matrix(g["keys"], df[g["keys"]].to_numpy(dtype=np.float64), g["keys"], g["vals"], "2 kHz defaults")


def test_standalone_resampler_matches_oracle(backend):
    """``Resampler.process`` as a stand-alone preprocessor (same constructor as the reference class) against the restated
    ``mne.filter.resample``: down-sampling by 2, ratio 0.9 and up-sampling by 2.5, odd lengths."""
    from oracle.mne_filter_restated import resample
    from py_neuromodulation_b200.processing.resample import Resampler

    for sfreq, target, n in ((2000, 1000, 600), (1111.111, 1000, 371), (400, 1000, 233)):
        x = neural_like(5, 3, n, sfreq=sfreq)
        got = Resampler(sfreq=sfreq, resample_freq_hz=target).process(x)
        ref = resample(x, up=float(target / sfreq), down=1.0)
        assert got.shape == ref.shape
        assert np.max(np.abs(got - ref)) <= 1e-12 * max(1.0, np.max(np.abs(ref)))
    same = neural_like(6, 2, 100)
    assert Resampler(sfreq=1000, resample_freq_hz=1000).process(same) is same  # ratio 1: identity, like the reference


def test_stream_float_sampling_rate_variable_window_length(backend, tmp_path):
    """reference tests/test_timing.py:43-73: sfreq 1111.111, 333 ms segments -> windows of 369 and 370 samples."""
    g = load_golden("stream_float_fs")
    x = g["x"].astype(np.float64)
    s = nm.NMSettings(**g["settings"])
    stream = nm.Stream(sfreq=g["sfreq"], data=x, sampling_rate_features_hz=g["rate"], settings=s)
    df = stream.run(out_dir=tmp_path, experiment_name="floatfs")
    assert list(df.columns) == g["keys"]
    assert np.array_equal(df["time"].to_numpy(), g["vals"][:, g["keys"].index("time")])
    check_matrix(list(df.columns), df.to_numpy(), g["keys"], g["vals"], "float fs", tol=1e-7)


def test_bursts_history_two_tier_contract(backend):
    """exact agreement with the UNMODIFIED reference while its history buffer is not full
    (windows 0..290), agreement with the fixed oracle (true ring) afterwards."""
    g = load_golden("bursts_history_320")
    x = g["x"].astype(np.float64)
    n = 320 if backend == "gpu" else 300
    s = nm.NMSettings(**g["settings"])
    b = nm.Bursts(s, g["ch_names"], 1000)
    fixed = orc.BurstsOracle(g["settings"], g["ch_names"], 1000, faithful=False)
    for k in range(n):
        win = x[:, 100 * k : 100 * k + 1000]
        out = b.calc_feature(win)
        ref_fixed = fixed.calc(win)
        assert list(out.keys()) == g["keys"]
        v = np.array([float(t) for t in out.values()])
        rf = np.array([float(t) for t in ref_fixed.values()])
        assert np.max(np.abs(v - rf) / np.maximum(np.abs(rf), 1)) < 1e-12, f"window {k} vs fixed oracle"
        if k <= 290:
            r = g["vals"][k]
            assert np.max(np.abs(v - r) / np.maximum(np.abs(r), 1)) < 1e-12, f"window {k} vs unmodified reference"
            for j, key in enumerate(g["keys"]):
                if key.endswith(("_in_burst", "_duration_max", "_duration_mean")):
                    assert v[j] == r[j], (k, key)


def test_nan_handling_like_reference_tests(backend, tmp_path):
    """reference tests/test_nan_values.py: a NaN channel -> all its features NaN, others finite; a NaN span ->
    NaN only for the windows that overlap it."""
    x = uniform(3, 2, 3000)
    x[0, 1000:2000] = np.nan
    s = nm.NMSettings.get_fast_compute()
    s.features.raw_hjorth = True
    s.features.linelength = True
    stream = nm.Stream(sfreq=1000, data=x, settings=s)
    df = stream.run(out_dir=tmp_path, experiment_name="nan")
    t = df["time"].to_numpy()
    hit = (t > 1000) & (t < 2000 + 1000)
    ch0 = df.filter(like="ch0")
    ch1 = df.filter(like="ch1")
    assert ch0[hit].isna().all().all() and ch0[~hit].notna().all().all()
    assert ch1.notna().all().all()
    # and the oracle agrees on the values
    cols, ref = orc.run_offline(x, 1000, s.model_dump())
    check_matrix(list(df.columns), df.to_numpy(), cols, ref, "nan span")


def test_target_channel_and_bad_channel(backend, tmp_path):
    """target channels are appended raw (stream.py:145-170); bad / unused channels are skipped."""
    from py_neuromodulation_b200.utils.channels import set_channels

    x = uniform(9, 5, 2500)
    ch = set_channels(["ecog_0", "ecog_1", "ecog_2", "lfp_x", "MOV_RIGHT"], ["ecog", "ecog", "ecog", "dbs", "misc"],
                      reference="default", bads=["ecog_2"], target_keywords=["mov"])
    assert list(ch["used"]) == [1, 1, 0, 1, 0] and list(ch["target"]) == [0, 0, 0, 0, 1]
    s = nm.NMSettings.get_fast_compute()
    s.postprocessing.feature_normalization = False
    stream = nm.Stream(sfreq=1000, data=x, channels=ch, settings=s)
    df = stream.run(out_dir=tmp_path, experiment_name="tgt")
    assert "MOV_RIGHT" in df.columns and df.columns[-2] == "time"
    starts, lengths, _ = window_grid(2500, 1000, 10, 1000)
    assert np.array_equal(df["MOV_RIGHT"].to_numpy(), x[4, starts + lengths - 1])
    assert not any("ecog_2" in c for c in df.columns)
    chd = {k: list(v) for k, v in ch.to_dict(orient="list").items()}
    cols, ref = orc.run_offline(x, 1000, s.model_dump(), channels=chd)
    check_matrix(list(df.columns), df.to_numpy(), cols, ref, "targets")


def test_custom_python_feature_runs_next_to_gpu_features(backend, tmp_path):
    """reference examples/plot_2_example_add_feature.py: a duck-typed user feature sees the preprocessed window."""

    class ChannelMean:
        def __init__(self, settings, ch_names, sfreq):
            self.ch_names = ch_names

        def calc_feature(self, data):
            return {f"{ch}_chmean": float(np.mean(data[i])) for i, ch in enumerate(self.ch_names)}

    nm.add_custom_feature("channel_mean", ChannelMean)
    try:
        x = uniform(4, 3, 1600)
        s = nm.NMSettings.get_fast_compute()
        s.postprocessing.feature_normalization = False
        stream = nm.Stream(sfreq=1000, data=x, settings=s)
        df = stream.run(out_dir=tmp_path, experiment_name="custom")
        assert "ch0_avgref_chmean" in df.columns
        wo = orc.WindowOracle(1000, {**s.model_dump(), "features": {**s.model_dump()["features"], "channel_mean": False}},
                              n_channels=3)
        pre = wo.preprocess(x[:, :1000])
        assert abs(df["ch0_avgref_chmean"].iloc[0] - pre[0].mean()) < 1e-12
    finally:
        nm.remove_custom_feature("channel_mean")


def _burst_stress_signal(seed, c, t):
    """Amplitude jumps (bracket rebuilds), an exactly constant stretch (ties -> direct selection), slow drift."""
    rng = np.random.default_rng(seed)
    tt = np.arange(t) / 1000.0
    x = rng.standard_normal((c, t)) * 0.1
    x += np.sin(2 * np.pi * 17 * tt) * (0.2 + 2.0 * ((tt % 7.0) > 5.0))      # strong beta episodes
    x *= 1.0 + 0.5 * np.sin(2 * np.pi * 0.05 * tt)                             # slow amplitude drift
    x[:, 12000:16500] = 0.0                                                    # flat line: envelope ties at 0
    x[:, 30000:31000] *= 50.0                                                  # artefact
    return x.astype(np.float32).astype(np.float64)


@pytest.mark.parametrize("duration_s,n_windows,n_ch", [(3.0, 420, 3), (30.0, 150, 3), (3.0, 420, 1), (30.0, 200, 1)])
def test_burst_thresholds_incremental_equals_direct(backend, duration_s, n_windows, n_ch):
    """The sliding order-statistic state (bracket + FIFO queue) must reproduce the per-window re-selection bit for
    bit: short history (constant expiry), chunk boundaries, streamed single windows, rebuilds, ties.  With few rows
    (one channel: 2 rows on the 4 emulated / 148 real SMs) the windows of a launch are split into ranges that run on
    CTAs of their own, each later range starting with a bracket rebuild (`nm_burst_thr_split`)."""
    x = _burst_stress_signal(3, n_ch, 46000)
    s = nm.NMSettings.get_default().reset()
    s.features.bursts = True
    s.bursts_settings.time_duration_s = duration_s
    s.postprocessing.feature_normalization = False
    s.preprocessing = []
    outs = []
    for incremental in (True, False):
        dp = nm.DataProcessor(sfreq=1000, settings=s, channels=get_default_channels_from_data(x), line_noise=50, verbose=False)
        starts, lengths, _ = window_grid(x.shape[1], 1000, s.sampling_rate_features_hz, s.segment_length_features_ms)
        starts = starts[:n_windows]
        plan = dp.plan(int(lengths[0]))
        plan.pipe.set_burst_threshold_mode(incremental)
        cols, mat = dp.process_windows(x, starts, int(lengths[0]))
        outs.append(mat)
        if incremental:  # the sliding path must carry most windows: 6 rows x n_windows row-windows in total
            rebuilds, direct = plan.pipe.burst_threshold_stats()
            print(f"burst thresholds: {rebuilds} rebuilds, {direct} direct windows of {2 * n_ch * n_windows}")
            assert rebuilds + direct < 0.35 * 2 * n_ch * n_windows
    assert np.array_equal(outs[0], outs[1]), np.argwhere(outs[0] != outs[1])[:5]
    # and against the oracle (true-ring variant) on the same windows
    ref_cols, ref = orc.run_offline(x, 1000, s.model_dump(), max_windows=n_windows)  # faithful_bursts=False: true ring
    ref = ref[:, : len(cols)]
    check_matrix(cols, outs[0], ref_cols[: len(cols)], ref, "burst stress")
    for j, k in enumerate(cols):
        if k.endswith("_in_burst") or k.endswith("_duration_max"):
            assert np.array_equal(outs[0][:, j], ref[:, j]), k


def test_burst_thresholds_streaming_equals_batch(backend):
    """One window per call (state saved / restored around every launch) == one batched run."""
    x = _burst_stress_signal(5, 2, 20000)
    s = nm.NMSettings.get_default().reset()
    s.features.bursts = True
    s.bursts_settings.time_duration_s = 2.0
    s.postprocessing.feature_normalization = False
    s.preprocessing = []
    ch = get_default_channels_from_data(x)
    dp = nm.DataProcessor(sfreq=1000, settings=s, channels=ch, line_noise=50, verbose=False)
    starts, lengths, _ = window_grid(x.shape[1], 1000, s.sampling_rate_features_hz, s.segment_length_features_ms)
    starts = starts[:120]
    cols, mat = dp.process_windows(x, starts, 1000)
    dp2 = nm.DataProcessor(sfreq=1000, settings=s, channels=ch, line_noise=50, verbose=False)
    for k, st in enumerate(starts):
        d = dp2.process(x[:, st : st + 1000])
        v = np.array([float(t) for t in d.values()])
        assert np.array_equal(v, mat[k]), k


def test_pipelined_upload_and_chunked_download(backend):
    """Recordings of >= 65 536 samples are uploaded in time slices that are re-referenced lazily, and finished chunks are
    copied back while later chunks compute: windows straddling slice boundaries, a NaN span and the last samples of the
    recording must still match the oracle window by window."""
    T = 70200
    x = neural_like(11, 3, T)
    x[1, 30000:30010] = np.nan
    s = nm.NMSettings.get_default().reset()
    for f in ("fft", "raw_hjorth", "linelength", "return_raw"):
        s.features[f] = True
    s.postprocessing.feature_normalization = False
    dp = nm.DataProcessor(sfreq=1000, settings=s, channels=get_default_channels_from_data(x), line_noise=50, verbose=False)
    starts = np.concatenate([np.arange(0, T - 1000, 1733), [8959 - 500, 8960, 17920 - 1, T - 1000]]).astype(np.int64)
    cols, mat = dp.process_windows(x, starts, 1000)
    ora = orc.WindowOracle(1000, s.model_dump(), n_channels=3, line_noise=50)
    for k, st in enumerate(starts):
        f = ora.process(x[:, st : st + 1000])
        ref = np.array([float(f[c]) for c in cols])
        got = mat[k]
        assert np.array_equal(np.isnan(got), np.isnan(ref)), (k, st)
        fin = np.isfinite(ref)
        assert np.max(np.abs(got[fin] - ref[fin]) / np.maximum(np.abs(ref[fin]), 1.0)) < TOL, (k, st)
    # a second run on the same pipeline (buffers reused, events re-recorded) gives the same matrix
    cols2, mat2 = dp.process_windows(x, starts, 1000)
    assert np.array_equal(mat, mat2, equal_nan=True)


def test_standalone_preprocessing_filter_matches_oracle(backend):
    """PreprocessingFilter.process outside a DataProcessor (reference class of the same name and signature)."""
    from py_neuromodulation_b200.processing import PreprocessingFilter

    x = neural_like(31, 3, 1000)
    s = nm.NMSettings.get_default()
    pf = PreprocessingFilter(s, 1000)
    assert len(pf.filters) == 4
    got = pf.process(x)
    ref = orc.apply_prefilters(x, orc.design_prefilters(s.model_dump(), 1000))
    assert got.shape == ref.shape == x.shape
    assert np.max(np.abs(got - ref)) <= 1e-11 * max(1.0, np.max(np.abs(ref)))


def test_raw_normalizer_streaming_and_standalone(backend):
    """RawNormalizer state across calls: window-by-window DataProcessor.process == batched run == stand-alone class ==
    oracle; the scikit-learn variants are refused loudly."""
    from py_neuromodulation_b200.processing import RawNormalizer

    g = load_golden("dataprocessor_rawnorm_zscore")
    x = g["x"].astype(np.float64)
    s = nm.NMSettings(**g["settings"])
    ch = get_default_channels_from_data(x)
    dp = nm.DataProcessor(sfreq=1000, settings=s, channels=ch, line_noise=50, verbose=False)
    starts, _, _ = window_grid(x.shape[1], 1000, s.sampling_rate_features_hz, s.segment_length_features_ms)
    for k in range(30):
        d = dp.process(x[:, starts[k] : starts[k] + 1000])
        v = np.array([float(d[c]) for c in g["keys"]])
        ref = g["vals"][k]
        assert np.max(np.abs(v - ref) / np.maximum(np.abs(ref), 1.0)) < TOL, k
    # stand-alone class on raw windows (no other preprocessing) against the oracle's restatement
    rn = RawNormalizer(sfreq=1000, settings=s)
    ora = orc.RawNormalizerOracle(s.model_dump(), 1000)
    for k in range(30):
        w = x[:, starts[k] : starts[k] + 1000]
        got, ref = rn.process(w), ora.process(w)
        assert np.max(np.abs(got - ref)) < 1e-10, k
    s2 = nm.NMSettings(**g["settings"])
    s2.raw_normalization_settings.normalization_method = "power"  # scikit-learn PowerTransformer: not offered (minmax / robust are)
    with pytest.raises(NotImplementedError):
        nm.DataProcessor(sfreq=1000, settings=s2, channels=ch, line_noise=50, verbose=False)


@pytest.mark.parametrize("name", ["dataprocessor_c3_nan", "dataprocessor_fast", "dataprocessor_prefilter_default"])
def test_float32_linear_mode_within_north_star_tolerance(backend, name):
    """precision="f32": float32 arithmetic inside the notch / band-pass FFT convolutions (moments and outputs float64).
    Tolerance = the task's 1e-5 relative (|ref| >= 1: relative, else absolute 1e-5), NaN pattern identical."""
    g = load_golden(name)
    x = g["x"].astype(np.float64)
    s = nm.NMSettings(**g["settings"])
    s.postprocessing.feature_normalization = False  # z-scores of a handful of windows amplify any input difference
    dp = nm.DataProcessor(sfreq=g["sfreq"], settings=s, channels=get_default_channels_from_data(x), line_noise=50, verbose=False,
                          precision="f32")
    starts, lengths, _ = window_grid(x.shape[1], g["sfreq"], s.sampling_rate_features_hz, s.segment_length_features_ms)
    cols, mat = dp.process_windows(x, starts, int(lengths[0]))
    dp64 = nm.DataProcessor(sfreq=g["sfreq"], settings=s, channels=get_default_channels_from_data(x), line_noise=50, verbose=False)
    _, ref = dp64.process_windows(x, starts, int(lengths[0]))
    assert np.array_equal(np.isnan(mat), np.isnan(ref))
    fin = np.isfinite(ref)
    err = np.abs(mat[fin] - ref[fin]) / np.maximum(np.abs(ref[fin]), 1.0)
    assert err.max() < 1e-5, float(err.max())
    assert err.max() > 0.0  # the float32 kernels really ran


def test_c_abi_error_paths_are_loud_and_recoverable(backend):
    """Errors come back as RuntimeError with the library's message (rc != 0 + nm_last_error) and leave the pipeline usable."""
    from py_neuromodulation_b200 import _lib
    from py_neuromodulation_b200._pipeline import Pipeline, ScanSpec

    names = ["a", "b"]
    spec = ScanSpec(names, hjorth=True, raw=True, linelength=True)
    cols = spec.keys_hjorth() + spec.keys_raw() + spec.keys_linelength()
    pipe = Pipeline(2, 2, 200, cols)
    with pytest.raises(RuntimeError, match="finalize"):
        pipe.upload(np.zeros((2, 1000)))
    spec.attach(pipe)
    pipe.finalize()
    with pytest.raises(RuntimeError, match="already finalized"):
        pipe.finalize()
    with pytest.raises(RuntimeError, match="upload a recording"):
        pipe.run([0])
    x = np.random.default_rng(0).random((2, 1000))
    pipe.upload(x)
    with pytest.raises(RuntimeError, match="outside the recording"):
        pipe.run([0, 900])
    with pytest.raises(RuntimeError, match="bad recording geometry"):
        pipe.upload(np.zeros((2, 100)))  # shorter than one window
    with pytest.raises(ValueError):
        pipe.upload(np.zeros((3, 1000)))  # wrong number of rows: refused by the host wrapper
    lib = _lib.load()
    assert lib.nm_set_output_pitch(pipe._h, 1) != 0 and b"pitch" in lib.nm_last_error()
    out = pipe.run([0, 400, 800])  # still works after the failures
    ref = orc.hjorth(x[:, 800:1000], names)
    assert abs(out[2, cols.index("a_RawHjorth_Activity")] - ref["a_RawHjorth_Activity"]) < 1e-12
    with pytest.raises(RuntimeError, match="out of range"):
        Pipeline(2, 2, 200, cols).set_pick([0, 5])
    with pytest.raises(RuntimeError):
        Pipeline(2, 3, 200, cols)  # more feature channels than raw rows


def test_default_configurations_are_served_by_the_specialised_kernels(backend):
    """nm_describe_plan: at the default 1 kHz / 2 kHz geometries every FIR family runs nm_convx_kernel and every segment
    DFT nm_specx_kernel (the runtime-plan / generic kernels are for unusual sizes only)."""
    for sfreq in (1000.0, 2000.0):
        x = np.zeros((4, int(sfreq) * 2))
        s = nm.NMSettings.get_default()
        s.features.bandpass_filter = True
        s.features.stft = True
        s.raw_resampling_settings.resample_freq_hz = sfreq
        dp = nm.DataProcessor(sfreq=sfreq, settings=s, channels=get_default_channels_from_data(x), line_noise=50, verbose=False)
        plan = dp.plan(int(sfreq)).pipe.describe_plan()
        lines = plan.splitlines()[1:]
        assert len(lines) == 7, plan  # notch+scan, fft, welch, stft, bandpower, sharpwave, bursts
        assert all(("nm_convx_kernel" in ln) or ("nm_specx_kernel" in ln) or ("nm_notchx_kernel" in ln) for ln in lines), plan
        # the notch gets its rows through the TMA engine (cp.async.bulk + mbarrier staged variant of nm_convx_kernel)
        assert lines[0].startswith("notch+scan: nm_notchx_kernel"), plan
