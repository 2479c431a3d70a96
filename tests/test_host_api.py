"""CPU-only tests of the host side: settings surface and validation (mirroring the reference's own tests),
channel table helpers, window grid, FIR design, key ordering, and the C ABI (library loads, every symbol of
include/nmb200.h is exported; no compute call is made without a GPU)."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest
from pydantic import ValidationError

import py_neuromodulation_b200 as nm
from oracle import np_oracle as orc
from py_neuromodulation_b200 import _lib
from py_neuromodulation_b200.stream.generator import RawDataGenerator, window_grid
from py_neuromodulation_b200.utils.channels import get_default_channels_from_data, set_channels

ROOT = Path(__file__).resolve().parents[1]


# ----------------------------------------------------------------------------- C ABI
def test_cuda_library_loads_and_exports_every_declared_symbol():
    header = (ROOT / "include" / "nmb200.h").read_text()
    declared = set(re.findall(r"\b(nm_[a-z0-9_]+)\s*\(", header))
    declared -= {"nm_pipeline", "nm_spectral_cfg"}
    assert len(declared) >= 35
    assert _lib.LIB_PATH.is_file(), "build the CUDA library first (python -c 'import __graft_entry__ as g; g.build()')"
    lib = ctypes.CDLL(str(_lib.LIB_PATH))  # statically linked cudart: loads without a GPU
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/nmb200.h but not exported by libnmb200.so"
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    lib.nm_abi_version.restype = ctypes.c_int
    assert lib.nm_abi_version() == 1


def test_library_has_no_cpu_path_without_gpu():
    lib = _lib.declare(ctypes.CDLL(str(_lib.LIB_PATH)))
    n = ctypes.c_int(-1)
    rc = lib.nm_device_count(ctypes.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is visible here")
    handle = ctypes.c_void_p()
    assert lib.nm_pipeline_create(0, 2, 2, 100, 4, ctypes.byref(handle)) != 0
    assert lib.nm_last_error()  # loud, descriptive failure instead of a fallback


def test_sass_contains_sm100a_code():
    import shutil
    import subprocess

    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    out = subprocess.run(["cuobjdump", "-lelf", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in out


# ----------------------------------------------------------------------------- settings
def test_default_settings_match_reference_defaults():
    s = nm.NMSettings.get_default()
    assert s.features.get_enabled() == ["raw_hjorth", "return_raw", "fft", "welch", "sharpwave_analysis", "bursts", "linelength"]
    assert list(s.frequency_ranges_hz) == ["theta", "alpha", "low_beta", "high_beta"]
    assert s.preprocessing == ["raw_resampling", "notch_filter", "re_referencing"]
    assert s.sampling_rate_features_hz == 10 and s.segment_length_features_ms == 1000
    assert s.bursts_settings.frequency_bands == ["low_beta", "high_beta"] and s.bursts_settings.threshold == 75
    assert s.fft_settings.return_spectrum is False and nm.features.OscillatorySettings().return_spectrum is True
    f = nm.NMSettings.get_fast_compute()
    assert f.features.get_enabled() == ["fft"] and f.postprocessing.feature_normalization is True


def test_settings_validation_errors_like_reference_tests():
    s = nm.NMSettings.get_default()
    s.fft_settings.log_transform = "123"
    with pytest.raises(ValidationError):
        nm.FFT(s, ["a"], 1000)
    s = nm.NMSettings.get_default()
    s.bursts_settings.frequency_bands = ["wrong_band"]
    with pytest.raises(ValidationError):
        nm.Bursts(s, ["a"], 1000)
    for attr, val in (("threshold", -1), ("time_duration_s", -1)):
        s = nm.NMSettings.get_default()
        setattr(s.bursts_settings, attr, val)
        with pytest.raises(ValidationError):
            nm.Bursts(s, ["a"], 1000)
    s = nm.NMSettings.get_default()
    s.bursts_settings.burst_features.duration = -1
    with pytest.raises(ValidationError):
        nm.Bursts(s, ["a"], 1000)
    s = nm.NMSettings.get_default()
    s.features.disable_all()
    with pytest.raises(ValidationError):
        s.validate()
    s = nm.NMSettings.get_default()
    s.features.bandpass_filter = True
    s.bandpass_filter_settings.segment_lengths_ms["theta"] = 5000
    with pytest.raises(ValidationError):
        s.validate()
    s = nm.NMSettings.get_default()
    s.sharpwave_analysis_settings.sharpwave_features.width = True  # no estimator listed for it
    with pytest.raises((ValidationError, AssertionError)):
        s.validate()
    with pytest.raises(ValidationError):
        nm.NMSettings(frequency_ranges_hz={"theta": [8, 4]})


def test_nyquist_check_and_window_length_check():
    data = np.random.random([4, 1000])
    s = nm.NMSettings.get_default().reset()
    s.features.fft = True
    s.frequency_ranges_hz = {"theta": [4, 8], "broadband": [10, 600]}
    with pytest.raises(AssertionError):
        nm.Stream(sfreq=1000, data=data, settings=s)
    s = nm.NMSettings.get_default()
    s.fft_settings.windowlength_ms = 2000
    with pytest.raises(AssertionError):
        nm.FFT(s, ["a"], 1000)


def test_settings_round_trip_and_spaces_in_band_names(tmp_path):
    s = nm.NMSettings.get_default()
    s.frequency_ranges_hz = {"low beta": [13, 20], "theta": [4, 8]}
    s.bursts_settings.frequency_bands = ["low beta"]
    s = s.validate()
    assert list(s.frequency_ranges_hz) == ["low_beta", "theta"] and s.bursts_settings.frequency_bands == ["low_beta"]
    (tmp_path / "exp").mkdir()
    s.save(tmp_path, "exp")
    back = nm.NMSettings.from_file(tmp_path / "exp" / "exp_SETTINGS.yaml")
    assert back.model_dump() == s.model_dump()
    assert nm.NMSettings.from_file(tmp_path / "exp").model_dump() == s.model_dump()  # directory lookup


def test_reference_settings_dump_loads_unchanged():
    """A model_dump() of the REFERENCE's NMSettings (stored in the golden fixtures) validates here and dumps identically."""
    from tests.helpers import load_golden

    ref_dump = load_golden("dataprocessor_default")["settings"]
    assert nm.NMSettings(**ref_dump).model_dump() == ref_dump


def test_post_init_edits_take_effect_like_reference(tmp_path, monkeypatch):
    """reference tests/test_settings_change_after_init.py: run() rebuilds the processor from the current settings."""
    s = nm.NMSettings.get_fast_compute()
    stream = nm.Stream(sfreq=1000, data=np.random.random([2, 1500]), settings=s)
    stream.settings.features.fft = False
    stream.settings.features.raw_hjorth = True
    captured = {}

    def fake_run(self, data, out_dir, experiment_name):
        captured["enabled"] = self.data_processor.settings.features.get_enabled()
        import pandas as pd

        return pd.DataFrame()

    monkeypatch.setattr(nm.Stream, "_run_batched", fake_run)
    monkeypatch.setattr(nm.Stream, "_save_after_stream", lambda self: None)
    stream.run(out_dir=tmp_path, experiment_name="x")
    assert captured["enabled"] == ["raw_hjorth"]


def test_out_of_scope_requests_fail_loudly():
    x = np.random.random([2, 2000])
    s = nm.NMSettings.get_default()
    s.features.fooof = True
    with pytest.raises(NotImplementedError):
        nm.Stream(sfreq=1000, data=x, settings=s)
    s = nm.NMSettings.get_default()
    s.preprocessing = ["raw_normalization", "notch_filter"]
    s.raw_normalization_settings.normalization_method = "quantile"  # scikit-learn transformer: out of scope
    with pytest.raises(NotImplementedError):
        nm.Stream(sfreq=1000, data=x, settings=s)
    s = nm.NMSettings.get_default()  # default resampling to 1 kHz of a 2 kHz recording (row 8f-2) is served since round 2
    assert nm.Stream(sfreq=2000, data=x, settings=s).data_processor.resample_ratio == 0.5
    s.raw_resampling_settings.resample_freq_hz = 2000  # identity, like the reference
    assert nm.Stream(sfreq=2000, data=x, settings=s).data_processor.resample_ratio is None
    s = nm.NMSettings.get_default()
    s.postprocessing.project_cortex = True
    with pytest.raises(NotImplementedError):
        nm.Stream(sfreq=1000, data=x, settings=s)
    with pytest.raises(ValueError):
        nm.Stream(sfreq=1000)
    with pytest.raises(NotImplementedError):
        nm.Stream(sfreq=1000, data=x).run(is_stream_lsl=True)


# ----------------------------------------------------------------------------- channels / grid / filters
def test_channel_tables():
    ch = get_default_channels_from_data(np.zeros((3, 10)))
    assert list(ch.columns) == ["name", "rereference", "used", "target", "type", "status", "new_name"]
    assert list(ch["new_name"]) == ["ch0_avgref", "ch1_avgref", "ch2_avgref"]
    ch = set_channels(["ECOG_L_1", "ECOG_L_2", "LFP_L_1", "LFP_L_2", "LFP_L_3", "MOV_LEFT"],
                      ["ecog", "ecog", "dbs", "dbs", "dbs", "misc"], bads=["LFP_L_3"])
    assert list(ch["rereference"]) == ["average", "average", "LFP_L_3", "LFP_L_1", "LFP_L_2", "None"]
    assert list(ch["used"]) == [1, 1, 1, 1, 0, 0] and list(ch["target"]) == [0, 0, 0, 0, 0, 1]
    assert ch["new_name"][2] == "LFP_L_1_LFP_L_3" and ch["new_name"][0] == "ECOG_L_1_avgref"
    with pytest.raises(ValueError):
        set_channels(["a"], ["ecog", "ecog"])


def test_reference_matrix_matches_oracle():
    from py_neuromodulation_b200.processing.rereference import build_reference_matrix

    ch = set_channels(["e0", "e1", "e2", "e3", "s0", "s1"], ["ecog"] * 4 + ["seeg"] * 2, reference=["average", "average", "e0&e1", "None", "s1", "s0"],
                      bads=None)
    m = build_reference_matrix(ch)
    ref = orc.reref_matrix({k: list(v) for k, v in ch.to_dict(orient="list").items()})
    assert np.array_equal(m, ref)
    with pytest.raises(ValueError):
        build_reference_matrix(set_channels(["a", "b"], ["ecog", "ecog"], reference=["a", "None"]))


def test_reference_matrix_factorisation_is_exact():
    from py_neuromodulation_b200._pipeline import factor_reference_matrix
    from py_neuromodulation_b200.processing.rereference import build_reference_matrix

    names = [f"c{i}" for i in range(12)]
    types = ["ecog"] * 7 + ["seeg"] * 5
    refs = ["average"] * 6 + ["c0&c1"] + ["average"] * 4 + ["None"]
    m = build_reference_matrix(set_channels(names, types, reference=refs))
    g, group_of, gcoef, rem = factor_reference_matrix(m)
    assert g == 2
    rebuilt = rem.copy()
    for grp in range(g):
        rebuilt += np.outer(gcoef[:, grp], (group_of == grp).astype(float))
    assert np.array_equal(rebuilt, m)
    assert np.count_nonzero(rem) < 3 * 12  # the dense average rows collapsed


@pytest.mark.parametrize("n,fs,rate,seg", [(10000, 1000, 3, 1000), (300000, 1000, 10, 1000), (120000, 2000, 10, 1000),
                                           (5555, 1111.111, 200, 333), (5000, 1000, 200, 1000), (2500, 1000, 7, 450)])
def test_window_grid_matches_reference_generator(n, fs, rate, seg):
    starts, lengths, times = window_grid(n, fs, rate, seg)
    ref = orc.window_grid(n, fs, rate, seg)
    assert len(ref) == len(starts)
    assert [(int(a), int(a + b), t) for a, b, t in zip(starts, lengths, times)] == ref
    gen = RawDataGenerator(np.zeros((1, n)), fs, rate, seg)
    for k, (ts, batch) in enumerate(gen):
        assert batch.shape[1] == lengths[k] and np.ceil(ts[-1] * 1000 + 1) == times[k]
    assert k == len(starts) - 1


def test_fir_design_matches_oracle_restatement():
    from oracle.mne_filter_restated import create_filter
    from py_neuromodulation_b200.filter.fir_design import design_fir

    for fs in (150, 500, 1000, 1111.111, 2000):
        for lo, hi in ((4, 8), (13, 20), (5, 30), (2, 6)):
            if hi >= fs / 2:
                continue
            for kw in ({}, {"filter_length": int(fs - 1), "l_trans_bandwidth": 4, "h_trans_bandwidth": 4}):
                try:
                    a = create_filter(None, fs, lo, hi, **kw)
                except ValueError:
                    with pytest.raises(ValueError):
                        design_fir(fs, lo, hi, use_mne=False, **kw)
                    continue
                b = design_fir(fs, lo, hi, use_mne=False, **kw)
                assert a.shape == b.shape and np.max(np.abs(a - b)) < 1e-14
    h = nm.filter.NotchFilter(1000, 50).filter_bank
    assert len(h) == 999 and abs(h.sum() - 1) < 1e-12 and abs(h[499] - 0.884774434572) < 1e-11
    assert np.max(np.abs(h - orc.design_notch(1000, 50))) < 1e-14
    assert nm.filter.NotchFilter(90, 50).filter_bank is None  # no harmonic below Nyquist -> identity


def test_feature_keys_match_reference_order_without_gpu():
    from py_neuromodulation_b200.stream.data_processor import build_specs
    from tests.helpers import load_golden

    for name in ("dataprocessor_default", "dataprocessor_c3_nan", "dataprocessor_realdata"):
        g = load_golden(name)
        s = nm.NMSettings(**g["settings"])
        names = [f"ch{i}_avgref" for i in range(g["x"].shape[0])]
        assert build_specs(s, names, 1000, 1000)[2] == g["keys"]
    g = load_golden("plugins_allest_uniform")
    s = nm.NMSettings(**g["settings"])
    from py_neuromodulation_b200._pipeline import SharpwaveSpec, SpectralSpec, band_items

    assert SharpwaveSpec(s, g["ch_names"], 1000).keys() == g["sharpwave_keys"]
    for kind in ("fft", "welch", "stft"):
        assert SpectralSpec(kind, s[f"{kind}_settings"], band_items(s), g["ch_names"], 1000, 1000).keys() == g[f"{kind}_keys"]
