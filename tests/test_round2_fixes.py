"""Regression tests for the round-2 review findings on the host side of the window processor
(stream/data_processor.py, stream/stream.py): stateful stages must advance once per window, NaN re-insertion maps a raw row
to its OWN channel and follows the normaliser, and window-length changes are rejected where a per-length history would be wrong."""
import numpy as np
import pytest

import py_neuromodulation_b200 as nm
from oracle import np_oracle as orc
from py_neuromodulation_b200.utils.channels import get_default_channels_from_data
from tests.helpers import neural_like, parity_err, uniform


class _ChannelMean:
    def __init__(self, settings, ch_names, sfreq):
        self.ch_names = ch_names

    def calc_feature(self, data):
        return {f"{ch}_chmean": float(np.mean(data[i])) for i, ch in enumerate(self.ch_names)}


def _rawnorm_settings():
    s = nm.NMSettings.get_fast_compute()
    s.features.raw_hjorth = True
    s.postprocessing.feature_normalization = False
    s.preprocessing = ["notch_filter", "re_referencing", "raw_normalization"]
    s.raw_normalization_settings.normalization_time_s = 2
    return s


def test_user_feature_does_not_advance_the_raw_normalizer_twice(backend):
    """A custom Python feature next to the GPU features (reference features/feature_processor.py:52-53) must not change the
    built-in features: both see a RawNormalizer (processing/normalization.py:44-51) that advanced ONCE per window."""
    x = neural_like(3, 3, 2200)
    wins = [x[:, 100 * k : 100 * k + 1000] for k in range(6)]

    def run():
        dp = nm.DataProcessor(sfreq=1000, settings=_rawnorm_settings(), channels=get_default_channels_from_data(x), line_noise=50, verbose=False)
        return [dp.process(w) for w in wins]

    plain = run()
    nm.add_custom_feature("channel_mean", _ChannelMean)
    try:
        both = run()
    finally:
        nm.remove_custom_feature("channel_mean")
    wo = orc.WindowOracle(1000, _rawnorm_settings().model_dump(), n_channels=3, line_noise=50)
    for k, (a, b) in enumerate(zip(plain, both)):
        assert [key for key in b if not key.endswith("_chmean")] == list(a)
        for key, v in a.items():
            assert b[key] == v, (k, key)  # identical pipeline state -> identical bits
        pre = wo.preprocess(wins[k])  # the oracle's normaliser advances once per window as well
        for i in range(3):
            got = b[f"ch{i}_avgref_chmean"]
            assert abs(got - pre[i].mean()) < 1e-9 * max(1.0, abs(pre[i].mean())), (k, i)


def test_nan_reinsertion_maps_each_raw_row_to_its_own_channel(backend):
    """stream/data_processor.py:253,297-306 with an unused / bad channel in front of used ones: a NaN in the unused row touches
    nothing, a NaN in a used row turns exactly that channel's features into NaN."""
    x = uniform(5, 3, 1000)
    ch = get_default_channels_from_data(x)
    ch.loc[1, "status"] = "bad"
    ch.loc[1, "used"] = 0
    s = nm.NMSettings.get_fast_compute()
    s.postprocessing.feature_normalization = False
    dp = nm.DataProcessor(sfreq=1000, settings=s, channels=ch, line_noise=50, verbose=False)
    clean = dp.process(x)
    assert not any(np.isnan(v) for v in clean.values())
    names = list(ch["new_name"])
    x1 = x.copy()
    x1[1, 10:20] = np.nan  # unused row
    f1 = dp.process(x1)
    assert f1 == clean
    x2 = x.copy()
    x2[2, 500] = np.nan
    f2 = dp.process(x2)
    for k, v in f2.items():
        assert np.isnan(v) == (names[2] in k), k


def test_window_length_change_with_stateful_stage_is_rejected(backend, tmp_path):
    x = uniform(6, 2, 2400)
    s = nm.NMSettings.get_fast_compute()  # feature normalisation on
    dp = nm.DataProcessor(sfreq=1000, settings=s, channels=get_default_channels_from_data(x), line_noise=50, verbose=False)
    dp.process(x[:, :1000])
    with pytest.raises(NotImplementedError):
        dp.process(x[:, :1001])
    # batched path: alternating window lengths (float sfreq) + RawNormalizer has no single sample history
    s2 = _rawnorm_settings()
    s2.segment_length_features_ms = 333
    s2.fft_settings.windowlength_ms = 333
    stream = nm.Stream(sfreq=1111.111, data=uniform(7, 2, 3000), sampling_rate_features_hz=3, settings=s2)
    with pytest.raises(NotImplementedError):
        stream.run(out_dir=tmp_path, experiment_name="x")


def test_variable_window_length_normaliser_keeps_values_of_nan_windows(backend, tmp_path):
    """Alternating window lengths (reference tests/test_timing.py:43-73) + rolling normalisation + a NaN span: the normaliser's
    history holds the values computed from the nan_to_num'ed samples and NaN is re-inserted afterwards
    (stream/data_processor.py:255-306), so the windows FOLLOWING the span match the reference as well."""
    x = uniform(8, 3, 4000)
    x[1, 1500:1510] = np.nan
    s = nm.NMSettings.get_fast_compute()
    s.segment_length_features_ms = 333
    s.fft_settings.windowlength_ms = 333
    s.preprocessing = ["notch_filter", "re_referencing"]
    s.sampling_rate_features_hz = 3
    stream = nm.Stream(sfreq=1111.111, data=x, sampling_rate_features_hz=3, settings=s)
    df = stream.run(out_dir=tmp_path, experiment_name="floatfs_nan")
    cols, ref = orc.run_offline(x, 1111.111, s.model_dump())
    assert list(df.columns) == cols
    mat = df.to_numpy()
    assert np.array_equal(np.isnan(mat), np.isnan(ref))
    assert np.isnan(ref).any() and not np.isnan(ref[-1]).any()
    assert parity_err(cols, mat, ref, normalized=True).max() < 1e-7


def test_stream_keeps_the_processor_of_an_unchanged_configuration(backend, tmp_path):
    """``Stream.run`` rebuilds its processor from the CURRENT settings like the reference (stream/stream.py:206-222) -- edits after
    ``__init__`` count -- but an unchanged configuration keeps its filter design and GPU plans; stateful stages restart."""
    x = neural_like(9, 3, 2600)
    stream = nm.Stream(sfreq=1000, data=x)  # default settings: bursts + rolling z-score are stateful
    dp0 = stream.data_processor
    a = stream.run(out_dir=tmp_path, experiment_name="a", save_csv=False)
    assert stream.data_processor is dp0
    b = stream.run(out_dir=tmp_path, experiment_name="a", save_csv=False)
    assert stream.data_processor is dp0
    assert np.array_equal(a.to_numpy(), b.to_numpy(), equal_nan=True)  # history reset between runs
    stream.settings.features.welch = False  # an edit after construction must be honoured
    c = stream.run(out_dir=tmp_path, experiment_name="a", save_csv=False)
    assert stream.data_processor is not dp0
    assert not any("_welch_" in k for k in c.columns) and any("_welch_" in k for k in a.columns)
    common = [k for k in c.columns if "_fft_" in k]
    assert np.array_equal(c[common].to_numpy(), a[common].to_numpy(), equal_nan=True)


def test_pipeline_writes_into_a_column_block_of_a_wider_table(backend):
    """``Pipeline.run(out=big[:, :F])`` (nm_set_output_pitch): with and without the sequential normaliser."""
    x = neural_like(4, 3, 2400)
    for norm in (False, True):
        s = nm.NMSettings.get_fast_compute()
        s.postprocessing.feature_normalization = norm
        dp = nm.DataProcessor(sfreq=1000, settings=s, channels=get_default_channels_from_data(x), line_noise=50, verbose=False)
        starts = np.arange(0, 1400, 100)
        cols, dense = dp.process_windows(x, starts, 1000)
        dp.reset_state()
        big = np.full((starts.size, len(cols) + 3), -7.0)
        _, view = dp.process_windows(x, starts, 1000, out=big[:, : len(cols)])
        assert np.array_equal(big[:, : len(cols)], dense, equal_nan=True) and np.all(big[:, len(cols):] == -7.0)


@pytest.mark.parametrize("graph", [False, True])
def test_window_stream_slot_ring_equals_batched_run(backend, graph):
    """Streaming entry: windows pushed through the page-locked slot ring -- two in flight, eager launches or CUDA-graph
    replay with patched window counters -- give the rows of ONE batched run: burst history, raw-sample and feature normalisers
    advance exactly once per window (reference: stream/data_processor.py:238-311 called per batch by stream/stream.py:280-330)."""
    from py_neuromodulation_b200.stream.window_stream import WindowStream

    x = neural_like(11, 5, 1000 + 100 * 23)
    s = nm.NMSettings.get_default()  # bursts + sharp waves + FFT/Welch + Hjorth + rolling z-score
    s.preprocessing = ["notch_filter", "re_referencing", "raw_normalization"]
    s.raw_normalization_settings.normalization_time_s = 1.5
    s.feature_normalization_settings.normalization_time_s = 1.2  # history shorter than the run: the fixed-capacity ring wraps
    ch = get_default_channels_from_data(x)
    starts = np.arange(0, 100 * 24, 100)
    ref_cols, ref = nm.DataProcessor(sfreq=1000, settings=s, channels=ch, line_noise=50, verbose=False).process_windows(x, starts, 1000)

    dp = nm.DataProcessor(sfreq=1000, settings=s, channels=ch, line_noise=50, verbose=False)
    ws = WindowStream(dp, 1000, slots=2, graph=graph, capacity=8)  # (capacity 8: the table grows twice)
    assert ws.columns == list(ref_cols)
    for k, st in enumerate(starts):
        ws.next_input()[...] = x[:, st : st + 1000]
        ws.submit(time_ms=float(st + 1000))
        if ws.in_flight == ws.slots:
            ws.collect()
    ws.drain()
    stats = ws.pipe.stream_stats()
    df = ws.to_frame()
    ws.close()
    got = df[ws.columns].to_numpy()
    assert got.shape == ref.shape and list(df["time"]) == [float(t + 1000) for t in starts]
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    assert parity_err(ref_cols, got, ref, normalized=True).max() < 1e-11
    for j, k in enumerate(ref_cols):
        if k.endswith("_in_burst"):
            assert np.array_equal(got[:, j], ref[:, j]), k
    assert stats["windows"] == len(starts)
    if backend == "gpu" and graph:  # captured once per slot, replayed for the rest, counters patched into the kernel nodes
        assert stats["graph_launches"] == len(starts) - 2 and stats["graph_kernel_nodes"] > 5 and stats["patched_arguments"] > 0
    else:
        assert stats["graph_launches"] == 0 or backend == "gpu"


@pytest.mark.parametrize("precision", ["f32x2", "f32"])
@pytest.mark.parametrize("n_ch", [5, 6, 7, 8])
def test_packed_float32_kernels_all_quad_remainders(backend, n_ch, precision):
    """float32 mode of the FIR families: two channel PAIRS share one item (packed float32 pairs, csrc/nm_convx.cuh); every remainder
    of the channel count modulo 4 leaves a differently filled last quad.  Gate: the task's 1e-5 (relative for |ref| >= 1)."""
    x = neural_like(50 + n_ch, n_ch, 1000 + 100 * 5)
    s = nm.NMSettings.get_default().reset()
    for f in ("fft", "bandpass_filter", "raw_hjorth", "linelength", "return_raw"):
        s.features[f] = True
    s.postprocessing.feature_normalization = False
    ch = get_default_channels_from_data(x)
    starts = np.arange(0, 600, 100)
    dp32 = nm.DataProcessor(sfreq=1000, settings=s, channels=ch, line_noise=50, verbose=False, precision=precision)
    cols, got = dp32.process_windows(x, starts, 1000)
    ref_cols, ref = orc.run_offline(x, 1000, s.model_dump())
    assert ref_cols[: len(cols)] == cols
    ref = ref[:, : len(cols)]
    err = np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)
    assert np.isfinite(got).all() and err.max() < 1e-5, (float(err.max()), cols[int(np.argmax(err.max(axis=0)))])
    _, got64 = nm.DataProcessor(sfreq=1000, settings=s, channels=ch, line_noise=50, verbose=False).process_windows(x, starts, 1000)
    assert np.abs(got - got64).max() > 0.0  # the float32 kernels really ran


@pytest.mark.parametrize("sfreq,precision", [(1000, "f64"), (2000, "f64"), (1000, "f32"), (1000, "f32x2")])
def test_same_mode_banks_on_three_times_power_of_two_plans(backend, monkeypatch, sfreq, precision):
    """'same'-mode FIR banks (band-pass power, burst envelopes) need a circular length of W + (L-1)/2 only and run on the
    12 x R1 x 16 plans of csrc/nm_convx.cuh (1536 points at 1 kHz, 3072 at 2 kHz) instead of 2048 / 4096: same results as the
    power-of-two plans (NMB200_MIXED_RADIX=0) to rounding, and as the oracle; odd channel count (half-filled last pair)."""
    W = int(sfreq)
    x = neural_like(77, 5, W + 100 * 7 * (W // 1000))
    s = nm.NMSettings.get_default().reset()
    for f in ("bandpass_filter", "bursts", "raw_hjorth"):
        s.features[f] = True
    if precision != "f64":
        s.features.bursts = False  # (threshold decisions keep every float32 pipeline's notch in float64: test the banks alone)
    s.postprocessing.feature_normalization = False
    s.raw_resampling_settings.resample_freq_hz = sfreq
    ch = get_default_channels_from_data(x)
    starts = np.arange(0, 7) * (W // 10)
    outs = {}
    for mixed in ("1", "0"):
        monkeypatch.setenv("NMB200_MIXED_RADIX", mixed)
        dp = nm.DataProcessor(sfreq=sfreq, settings=s, channels=ch, line_noise=50, verbose=False, precision=precision)
        cols, outs[mixed] = dp.process_windows(x, starts, W)
        plan = dp.plan(W).pipe.describe_plan()
        want = (f"P={3 * 512 * (W // 1000)}" if mixed == "1" else f"P={2048 * (W // 1000)}")
        assert all(want in line for line in plan.splitlines() if line.startswith(("bandpower", "bursts"))), plan
    tol = 1e-11 if precision == "f64" else 2e-5
    d = np.abs(outs["1"] - outs["0"]) / np.maximum(np.abs(outs["0"]), 1e-3 if precision == "f64" else 1.0)
    assert d.max() < tol, (float(d.max()), cols[int(np.argmax(d.max(axis=0)))])
    assert np.abs(outs["1"] - outs["0"]).max() > 0.0  # two different transforms really ran
    ref_cols, ref = orc.run_offline(x, sfreq, s.model_dump(), max_windows=len(starts))
    assert ref_cols[: len(cols)] == cols
    err = np.abs(outs["1"] - ref[:, : len(cols)]) / np.maximum(np.abs(ref[:, : len(cols)]), 1.0)
    assert err.max() < (1e-9 if precision == "f64" else 1e-5), float(err.max())


@pytest.mark.parametrize("method", ["minmax", "robust", "quantile"])
def test_sklearn_feature_normalizers_restated_on_the_gpu(backend, method):
    """processing/normalization.py:58-70,173-190: the reference hands the history to a scikit-learn transformer.  The stand-alone
    `FeatureNormalizer` (same kernel as inside the pipeline) against the oracle's restatement on vectors with ties (quantised
    values), a constant column and a history that is trimmed many times; `power` and long quantile histories raise."""
    rng = np.random.default_rng(5)
    s = nm.NMSettings.get_default()
    s.feature_normalization_settings.normalization_method = method
    s.feature_normalization_settings.normalization_time_s = 1.2  # 12 windows
    s.feature_normalization_settings.clip = 0.9 if method == "robust" else 3
    fn = nm.FeatureNormalizer(s) if hasattr(nm, "FeatureNormalizer") else None
    if fn is None:
        from py_neuromodulation_b200.processing.normalization import FeatureNormalizer
        fn = FeatureNormalizer(s)
    ora = orc.FeatureNormalizerOracle(s.model_dump())
    worst = 0.0
    for k in range(40):
        v = rng.standard_normal(9)
        v[1] = np.round(v[1] * 2) / 2          # ties
        v[2] = 0.25                            # constant column: zero scale -> 1
        v[3] = np.round(v[3])                  # heavy ties
        got = fn.process(v.copy())
        ref = ora.process(v.copy())
        assert np.array_equal(np.isnan(got), np.isnan(ref)), k
        worst = max(worst, float(np.nanmax(np.abs(got - ref))))
    assert worst < 1e-12, worst
    from py_neuromodulation_b200.processing.normalization import FeatureNormalizer
    s.feature_normalization_settings.normalization_method = "power"
    with pytest.raises(NotImplementedError):
        FeatureNormalizer(s)
    s.feature_normalization_settings.normalization_method = "quantile"
    s.feature_normalization_settings.normalization_time_s = 31  # 310 windows > n_quantiles
    with pytest.raises(NotImplementedError):
        FeatureNormalizer(s)


@pytest.mark.parametrize("method", ["zscore", "minmax", "quantile", "median", "robust"])
def test_feature_normalizer_across_chunks_of_a_batched_run(backend, method):
    """150 windows = three chunks of a batched run: the O(n_keep) methods normalise every chunk right behind its kernels (its rows
    are then shipped while later chunks compute), the order-statistic methods in one sliding pass at the end; history of 50
    windows, a NaN span (features NaN after the normaliser, history built from the nan_to_num'ed samples)."""
    from py_neuromodulation_b200.stream.generator import window_grid
    x = neural_like(3, 3, 1000 + 100 * 149)
    x[1, 7000:7300] = np.nan
    s = nm.NMSettings.get_default().reset()
    for f in ("raw_hjorth", "return_raw", "linelength"):
        s.features[f] = True
    s.preprocessing = ["re_referencing"]
    s.postprocessing.feature_normalization = True
    s.feature_normalization_settings.normalization_method = method
    s.feature_normalization_settings.normalization_time_s = 5.0
    dp = nm.DataProcessor(sfreq=1000, settings=s, channels=get_default_channels_from_data(x), line_noise=50, verbose=False)
    starts, _, _ = window_grid(x.shape[1], 1000, 10, 1000)
    cols, got = dp.process_windows(x, starts, 1000)
    assert dp.plan(1000).pipe.chunk_windows < len(starts)
    ref_cols, ref = orc.run_offline(x, 1000, s.model_dump())
    assert ref_cols[: len(cols)] == cols
    ref = ref[:, : len(cols)]
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    assert float(np.nanmax(np.abs(got - ref))) < 1e-10


@pytest.mark.parametrize("method", ["minmax", "robust"])
def test_standalone_raw_normalizer_with_sklearn_methods(backend, method):
    """`RawNormalizer.process(window)` (reference class name and interface) with the restated MinMaxScaler / RobustScaler: sliding
    minimum / maximum / percentiles of a 1.7 s sample history against the oracle, window by window."""
    from py_neuromodulation_b200.processing.normalization import RawNormalizer
    x = neural_like(8, 3, 1000 + 100 * 39)
    s = nm.NMSettings.get_default()
    s.raw_normalization_settings.normalization_method = method
    s.raw_normalization_settings.normalization_time_s = 1.7
    rn = RawNormalizer(1000, s)
    ora = orc.RawNormalizerOracle(s.model_dump(), 1000)
    changed = 0.0
    for k in range(40):
        w = x[:, 100 * k : 100 * k + 1000]
        got, ref = rn.process(w), ora.process(w)
        assert np.max(np.abs(got - ref)) < 1e-10, k
        changed = max(changed, float(np.max(np.abs(got - w))))
    assert changed > 0.1  # the windows really were normalised
