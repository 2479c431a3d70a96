"""GPU-only: the BASELINE.json configurations at full size, checked (a) against the oracle on sampled windows and
(b) through size-independent properties (bit-reproducibility, amplitude scaling laws, channel-permutation
equivariance, batch == streaming)."""
import numpy as np
import pytest

import py_neuromodulation_b200 as nm
from oracle import np_oracle as orc
from py_neuromodulation_b200.stream.generator import window_grid
from py_neuromodulation_b200.utils.channels import get_default_channels_from_data
from tests.helpers import neural_like

pytestmark = pytest.mark.gpu


@pytest.fixture()
def gpu(monkeypatch):
    from py_neuromodulation_b200 import _lib

    monkeypatch.setattr(_lib, "_LIB", None)
    _lib.load()
    assert _lib.device_count() > 0
    return True


def c3_settings():
    s = nm.NMSettings.get_default().reset()
    s.features.fft = True
    s.features.bandpass_filter = True
    s.features.raw_hjorth = True
    s.features.linelength = True
    return s


def rel(a, b):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0))


def test_c3_full_size_vs_oracle_and_properties(gpu):
    """256 ch x 300 s @ 1 kHz, notch + CAR, FFT + band-pass + Hjorth + line length: 2 991 windows, F = 3 072."""
    rng = np.random.default_rng(0)
    x = rng.random((256, 300_000), dtype=np.float32)
    s = c3_settings()
    dp = nm.DataProcessor(sfreq=1000, settings=s, channels=get_default_channels_from_data(x), line_noise=50, verbose=False)
    starts, lengths, _ = window_grid(x.shape[1], 1000, 10, 1000)
    assert starts.size == 2991
    cols, mat = dp.process_windows(x, starts, 1000)
    assert mat.shape == (2991, 3072) and np.isfinite(mat).all()
    # (a) oracle on sampled windows (the same float32 samples, up-cast)
    wo = orc.WindowOracle(1000, s.model_dump(), n_channels=256, line_noise=50)
    for k in (0, 1, 1499, 2990):
        ref = wo.process(x[:, starts[k] : starts[k] + 1000].astype(np.float64))
        assert list(ref.keys()) == cols
        assert rel(mat[k], np.array(list(ref.values()))) < 1e-9, k
    # (b1) bit-reproducible
    _, again = dp.process_windows(x, starts, 1000)
    assert np.array_equal(mat, again)
    # (b2) amplitude scaling laws on a slice of the run (alpha = 4 is exact in binary floating point)
    sub = starts[:200]
    _, base = dp.process_windows(x, sub, 1000)
    _, scaled = dp.process_windows(x * np.float32(4.0), sub, 1000)
    c = np.array(cols)
    is_act = np.char.endswith(c, "RawHjorth_Activity")
    is_inv = np.char.endswith(c, "RawHjorth_Mobility") | np.char.endswith(c, "RawHjorth_Complexity")
    is_ll = np.char.endswith(c, "LineLength")
    is_fft = np.char.find(c, "_fft_") >= 0
    is_bp = np.char.find(c, "_bandpass_activity_") >= 0
    assert rel(scaled[:, is_act], 16.0 * base[:, is_act]) < 1e-12
    assert rel(scaled[:, is_inv], base[:, is_inv]) < 1e-12
    assert rel(scaled[:, is_ll], 4.0 * base[:, is_ll]) < 1e-12
    assert rel(scaled[:, is_fft], base[:, is_fft] + np.log10(4.0)) < 1e-12
    assert rel(scaled[:, is_bp], base[:, is_bp] + 2 * np.log10(4.0)) < 1e-12
    assert (is_act | is_inv | is_ll | is_fft | is_bp).all()
    # (b3) channel permutation equivariance (common average over all channels)
    perm = rng.permutation(256)
    _, pm = dp.process_windows(x[perm], sub, 1000)
    by_name = {k: i for i, k in enumerate(cols)}
    idx = np.array([by_name[k.replace(f"ch{int(k[2:k.index('_')])}_", f"ch{perm[int(k[2:k.index('_')])]}_", 1)] for k in cols])
    assert rel(pm, base[:, idx]) < 1e-9


def test_c2_fft_only_2khz(gpu):
    """64 ch x 60 s @ 2 kHz, FFT band power only (fast compute), resampling set to the identity."""
    x = neural_like(5, 64, 120_000, 2000.0).astype(np.float32)
    s = nm.NMSettings.get_fast_compute()
    s.raw_resampling_settings.resample_freq_hz = 2000
    s.postprocessing.feature_normalization = False
    dp = nm.DataProcessor(sfreq=2000, settings=s, channels=get_default_channels_from_data(x), line_noise=50, verbose=False)
    starts, lengths, _ = window_grid(x.shape[1], 2000, 10, 1000)
    assert starts.size == 591 and lengths[0] == 2000
    cols, mat = dp.process_windows(x, starts, 2000)
    assert mat.shape == (591, 256)
    wo = orc.WindowOracle(2000, s.model_dump(), n_channels=64, line_noise=50)
    for k in (0, 300, 590):
        ref = wo.process(x[:, starts[k] : starts[k] + 2000].astype(np.float64))
        assert list(ref.keys()) == cols and rel(mat[k], np.array(list(ref.values()))) < 1e-9


def test_c4_oscillatory_bursts_sharpwave_one_shard(gpu):
    """The per-GPU share of C4: 32 ch x 30 s, FFT + Welch + STFT + bursts + sharp waves; 291 windows = the regime in
    which the reference's burst history is exact, so every window is compared, integers bit-exactly."""
    x = neural_like(6, 32, 30_000).astype(np.float32)
    s = nm.NMSettings.get_default().reset()
    for f in ("fft", "welch", "stft", "bursts", "sharpwave_analysis"):
        s.features[f] = True
    dp = nm.DataProcessor(sfreq=1000, settings=s, channels=get_default_channels_from_data(x), line_noise=50, verbose=False)
    starts, _, _ = window_grid(x.shape[1], 1000, 10, 1000)
    assert starts.size == 291
    cols, mat = dp.process_windows(x, starts, 1000)
    wo = orc.WindowOracle(1000, s.model_dump(), n_channels=32, line_noise=50, faithful_bursts=True)  # UNMODIFIED burst code path
    xd = x.astype(np.float64)
    worst = 0.0
    for k in range(291):
        ref = wo.process(xd[:, starts[k] : starts[k] + 1000])
        r = np.array([float(v) for v in ref.values()])
        if k == 0:
            assert list(ref.keys()) == cols
        worst = max(worst, rel(mat[k], r))
        for j, key in enumerate(cols):
            if key.endswith(("_in_burst", "_duration_max")):
                assert mat[k, j] == r[j], (k, key)
    assert worst < 1e-9, worst


def test_c5_shard_default_features_2khz(gpu):
    """One GPU's share of C5 scaled in time: 128 ch x 40 s @ 2 kHz, full default feature set incl. normalisation."""
    x = neural_like(7, 128, 80_000, 2000.0).astype(np.float32)
    s = nm.NMSettings.get_default()
    s.raw_resampling_settings.resample_freq_hz = 2000
    dp = nm.DataProcessor(sfreq=2000, settings=s, channels=get_default_channels_from_data(x), line_noise=50, verbose=False)
    starts, lengths, _ = window_grid(x.shape[1], 2000, 10, 1000)
    cols, mat = dp.process_windows(x, starts, int(lengths[0]))
    assert mat.shape == (starts.size, 128 * 31) and np.isfinite(mat).all()
    wo = orc.WindowOracle(2000, s.model_dump(), n_channels=128, line_noise=50)
    xd = x.astype(np.float64)
    for k in range(6):  # stateful (bursts, normaliser): the oracle has to walk from window 0
        ref = wo.process(xd[:, starts[k] : starts[k] + 2000])
        r = np.array([float(v) for v in ref.values()])
        assert rel(mat[k], r) < 1e-7, k


def test_batch_equals_streaming_at_scale(gpu):
    x = neural_like(8, 16, 6000).astype(np.float64)
    s = nm.NMSettings.get_default()
    ch = get_default_channels_from_data(x)
    dp = nm.DataProcessor(sfreq=1000, settings=s, channels=ch, line_noise=50, verbose=False)
    starts, _, _ = window_grid(x.shape[1], 1000, 10, 1000)
    cols, mat = dp.process_windows(x, starts, 1000)
    dp2 = nm.DataProcessor(sfreq=1000, settings=s, channels=ch, line_noise=50, verbose=False)
    for k in range(starts.size):
        d = dp2.process(x[:, starts[k] : starts[k] + 1000])
        v = np.array([float(t) for t in d.values()])
        assert rel(v, mat[k]) < 1e-12, k


def test_burst_thresholds_long_run_incremental_vs_direct_and_oracle(gpu):
    """Sliding order statistics far beyond the ring fill: 64 ch x 120 s (1 191 windows) incremental == direct selection bit
    for bit; 8 ch x 100 s against the oracle's true-ring quantile on every window."""
    s = nm.NMSettings.get_default().reset()
    s.features.bursts = True
    s.postprocessing.feature_normalization = False
    x = neural_like(9, 64, 120_000).astype(np.float32)
    starts, _, _ = window_grid(x.shape[1], 1000, 10, 1000)
    outs = []
    for incremental in (True, False):
        dp = nm.DataProcessor(sfreq=1000, settings=s, channels=get_default_channels_from_data(x), line_noise=50, verbose=False)
        dp.plan(1000).pipe.set_burst_threshold_mode(incremental)
        cols, mat = dp.process_windows(x, starts, 1000)
        outs.append(mat)
        if incremental:
            rebuilds, direct = dp.plan(1000).pipe.burst_threshold_stats()
            assert rebuilds + direct < 0.1 * 128 * starts.size, (rebuilds, direct)
    assert np.array_equal(outs[0], outs[1])
    x8 = x[:8, :100_000]
    starts8, _, _ = window_grid(x8.shape[1], 1000, 10, 1000)
    dp = nm.DataProcessor(sfreq=1000, settings=s, channels=get_default_channels_from_data(x8), line_noise=50, verbose=False)
    cols, mat = dp.process_windows(x8, starts8, 1000)
    ref_cols, ref = orc.run_offline(x8.astype(np.float64), 1000, s.model_dump())
    assert ref_cols[: len(cols)] == cols
    ref = ref[:, : len(cols)]
    assert rel(mat, ref) < 1e-9
    for j, key in enumerate(cols):
        if key.endswith(("_in_burst", "_duration_max")):
            assert np.array_equal(mat[:, j], ref[:, j]), key


def test_next_rows_at_scale_prefilter_and_raw_normalizer(gpu):
    """SURVEY 8f-3 at 256 channels: PreprocessingFilter + notch + CAR + RawNormalizer (zscore-median) in front of the C3
    feature set; the oracle walks the stateful normaliser from window 0."""
    x = neural_like(10, 256, 4000).astype(np.float32)
    s = c3_settings()
    s.preprocessing = ["preprocessing_filter", "notch_filter", "re_referencing", "raw_normalization"]
    s.preprocessing_filter.bandstop_filter = False
    s.raw_normalization_settings.normalization_time_s = 1.5
    s.raw_normalization_settings.normalization_method = "zscore-median"
    s.postprocessing.feature_normalization = False
    dp = nm.DataProcessor(sfreq=1000, settings=s, channels=get_default_channels_from_data(x), line_noise=50, verbose=False)
    starts, _, _ = window_grid(x.shape[1], 1000, 10, 1000)
    cols, mat = dp.process_windows(x, starts, 1000)
    assert "nm_convx_kernel" in dp.plan(1000).pipe.describe_plan()
    wo = orc.WindowOracle(1000, s.model_dump(), n_channels=256, line_noise=50)
    xd = x.astype(np.float64)
    for k in range(starts.size):
        ref = wo.process(xd[:, starts[k] : starts[k] + 1000])
        if k == 0:
            assert list(ref.keys()) == cols
        assert rel(mat[k], np.array([float(v) for v in ref.values()])) < 1e-8, k
