"""Benchmark of the per-window feature-extraction hot path (contract: see the task description / DESIGN.md section 6).

    python bench.py --gpus 1 --steps 10 --warmup 3            # this implementation on 1 B200
    torchrun ... bench.py --gpus N ...                        # channel-sharded over N B200s (weak scaling)
    python bench.py --impl reference ...                      # the UNMODIFIED reference (baseline/_ref through the shim) on the host cores
    python bench.py --config c4|c5|default [--gpus N]         # BASELINE.json configs[3..4] / the reference's default settings

Workload = BASELINE.json configs[2] ("C3"): 256 ch x 300 s @ 1 kHz, default preprocessing (notch + common average),
FFT + band-pass power + Hjorth + line length, 10 Hz feature rate -> 2 991 windows of 256 x 1000 samples per step.
With N GPUs every rank owns its own 256-channel shard of a 256*N-channel recording (common average over ALL channels).
A step = one pass of the hot path over the whole recording; metric = feature-windows/s where one unit is one
256 ch x 1000 samp window with all its features (BASELINE.json `metric`).

`value`  = resident step (recording already in HBM: re-reference + every window kernel), float64 arithmetic.
`e2e`    = the public call with HOST buffers: pinned recording in, feature matrix out (sliced asynchronous upload, rows of
           finished chunks copied back while the next chunk computes; N > 1: per-slice all-reduce of the common-average sums,
           every rank writes its block into one shared page-locked host matrix).
`roofline`, `cpu_baseline`, `clocks`, `gpu_launches`: see DESIGN.md section 5-6.  `f32_linear_mode` (N = 1 only) reports the
optional float32 mode of the FIR families beside the float64 headline.
"""

from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
# stdout carries exactly ONE JSON line: everything native libraries print there (NCCL's version banner / NCCL_DEBUG output) is
# sent to stderr at the file-descriptor level; emit() writes the line to the real stdout
_REAL_STDOUT = None


def claim_stdout() -> None:
    """Called by main(): from here on fd 1 is stderr for everybody but emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict) -> None:
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())

CH_PER_GPU = 256
SFREQ = 1000.0
DURATION_S = 300
LINE_NOISE = 50
METRIC = "feature-windows/sec (256ch x 1000samp fp32)"
UNIT = "windows/s"


def c3_settings():
    import py_neuromodulation_b200 as nm

    s = nm.NMSettings.get_default().reset()
    s.features.fft = True
    s.features.bandpass_filter = True
    s.features.raw_hjorth = True
    s.features.linelength = True
    return s


def synth(n_ch: int, n_samples: int, seed: int, out: np.ndarray | None = None) -> np.ndarray:
    """Uniform [0, 1) float32 -- what every reference example / test feeds (README.rst:83)."""
    rng = np.random.default_rng(seed)
    if out is None:
        out = np.empty((n_ch, n_samples), dtype=np.float32)
    rng.random(out=out, dtype=np.float32)
    return out


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int) -> None:
        self.device = device
        self.samples: list[list[str]] = []
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._loop, daemon=True)

    def _loop(self) -> None:
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-i", str(self.device)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([v.strip() for v in out.splitlines()[0].split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._thread.start()
        return self

    def __exit__(self, *exc) -> None:
        self._stop.set()
        self._thread.join(timeout=6)

    def summary(self) -> dict:
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 2 + i and s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------ pinned memory
def pinned_array(lib, shape, dtype) -> np.ndarray:
    from py_neuromodulation_b200 import _lib

    n_bytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    ptr = C.c_void_p()
    _lib.check(lib.nm_host_alloc(C.byref(ptr), n_bytes))
    buf = (C.c_char * n_bytes).from_address(ptr.value)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


def ncu_traffic(kernel: str, windows_per_launch: float) -> tuple[float | None, dict]:
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/r2_ncu_traffic.json), rescaled to this
    run's windows per launch; None if the capture does not cover the kernel."""
    f = ROOT / "profiles" / "r2_ncu_traffic.json"
    try:
        d = json.loads(f.read_text())
        k = d["kernels"][kernel]
        scale = windows_per_launch / k["windows_per_launch"]
        return (k["dram_bytes_read"] + k["dram_bytes_write"]) * scale, {"fp64_pipe_pct": k["fp64_pipe_pct"], "l1tex_pct": k["l1tex_pct"],
                                                                           "source": d["source"]}
    except Exception:
        return None, {}


def measured_peak_gbs() -> tuple[float, str]:
    f = ROOT / "MEASURED_PEAKS.json"
    if f.is_file():
        try:
            return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------ CPU legs
def _oracle_settings() -> dict:
    return c3_settings().model_dump()


# The CPU legs time the UNMODIFIED reference (`DataProcessor.process`, stream/data_processor.py:238-311) whenever its files are
# present -- /root/reference in the authoring container, the byte-identical copy under the git-ignored baseline/_ref on the GPU box
# (oracle/build_ref.py) -- through oracle/ref_shim.py (bare package objects + the restated mne.filter entry points: MNE itself is
# not installed).  Only when neither exists they fall back to the NumPy port (oracle/np_oracle.py, kind "port").
_REF_DP = None
_REF_X = None


def _reference_processor(n_ch: int):
    """Reference DataProcessor for the C3 settings on `n_ch` default channels, or None when the reference files are absent."""
    import logging

    from oracle import ref_shim

    if not ref_shim.reference_available():
        return None
    ref = ref_shim.load_reference()
    ref.logger.setLevel(logging.ERROR)
    settings = ref.NMSettings(**_oracle_settings())
    channels = ref.utils.channels.get_default_channels_from_data(np.empty((n_ch, 1)))
    return ref.DataProcessor(sfreq=SFREQ, settings=settings, channels=channels, line_noise=LINE_NOISE, verbose=False)


def _port_processor(n_ch: int):
    from oracle import np_oracle as orc

    return orc.WindowOracle(SFREQ, _oracle_settings(), n_channels=n_ch, line_noise=LINE_NOISE)


def cpu_baseline_single(budget_s: float = 12.0) -> dict:
    """The reference (else its port) in one process -- the way the reference runs -- on a bounded sample of the same workload."""
    n_win_max = 400
    x = synth(CH_PER_GPU, int(1000 + 100 * n_win_max), seed=0).astype(np.float64)
    proc = _reference_processor(CH_PER_GPU)
    kind = "reference" if proc is not None else "port"
    if proc is None:
        proc = _port_processor(CH_PER_GPU)
    proc.process(x[:, :1000])  # warm-up: FIR design, FFT plans
    done, t0 = 0, time.perf_counter()
    while done < n_win_max:
        proc.process(x[:, 100 * done : 100 * done + 1000])
        done += 1
        if time.perf_counter() - t0 > budget_s and done >= 8:
            break
    dt = time.perf_counter() - t0
    what = "unmodified reference DataProcessor.process via oracle/ref_shim.py" if kind == "reference" else "NumPy port (oracle/np_oracle.py)"
    return {"value": done / dt, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": f"{done} consecutive windows of the C3 workload (256 ch x 1000 samp, float64), {dt:.1f} s, single process, {what}"}


def _ref_init(n_ch: int, n_samples: int) -> None:
    """Pool initialiser: one processor per worker process, BLAS / OpenMP threads pinned to 1 (the pool supplies the parallelism)."""
    global _REF_DP, _REF_X
    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(1)
    except Exception:
        pass
    _REF_DP = _reference_processor(n_ch) or _port_processor(n_ch)
    _REF_X = synth(n_ch, n_samples, seed=0).astype(np.float64)
    _REF_DP.process(_REF_X[:, :1000])  # warm-up


def _ref_windows(ks) -> int:
    for k in ks:
        _REF_DP.process(_REF_X[:, 100 * k : 100 * k + 1000])
    return len(ks)


def run_reference_arm(args) -> None:
    """--impl reference: the reference's own CPU implementation of the path on all host cores.

    Every worker process runs the unmodified `DataProcessor.process` on whole 256-channel windows; the windows of a step are dealt
    round-robin to the workers (C3 has no state across windows, so this is the reference's best embarrassingly-parallel case and
    needs no change to its code).  A step is a bounded sample (4 windows per core, at least 64); `value` uses the MEDIAN step time."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp

    from oracle import ref_shim

    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[v] = "1"
    kind = "reference" if ref_shim.reference_available() else "port"
    cores = os.cpu_count() or 1
    n_win = max(64, 4 * cores)
    n_samples = 1000 + 100 * (n_win - 1)
    jobs = [list(range(w, n_win, cores)) for w in range(cores)]
    steps = max(args.steps, 5)
    times = []
    with mp.get_context("fork").Pool(cores, initializer=_ref_init, initargs=(CH_PER_GPU, n_samples)) as pool:
        for _ in range(max(1, args.warmup)):
            pool.map(_ref_windows, jobs, chunksize=1)
        for _ in range(steps):
            t0 = time.perf_counter()
            pool.map(_ref_windows, jobs, chunksize=1)
            times.append(time.perf_counter() - t0)
    dt = float(np.median(times))
    value = n_win / dt
    what = ("unmodified reference DataProcessor.process (baseline/_ref or /root/reference through oracle/ref_shim.py; mne.filter restated)"
            if kind == "reference" else "NumPy port of the reference (oracle/np_oracle.py)")
    sample = (f"{n_win} windows of the C3 workload per step (256 ch x 1000 samp, float64), dealt to {cores} worker processes, "
              f"BLAS/OMP threads = 1, median of {steps} steps (min {min(times):.2f} s, max {max(times):.2f} s); {what}")
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C3: 256ch x 1000samp windows, notch+CAR, FFT+bandpass+Hjorth+linelength (bounded sample: "
                               f"{n_win} windows/step)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


# ------------------------------------------------------------------------------------------ GPU arm
def c4_settings():
    import py_neuromodulation_b200 as nm

    s = nm.NMSettings.get_default().reset()
    for f in ("fft", "welch", "stft", "bursts", "sharpwave_analysis"):
        s.features[f] = True
    return s


def default_settings(sfreq: float = 1000.0):
    import py_neuromodulation_b200 as nm

    return nm.NMSettings.get_default()  # untouched: raw_resampling -> 1000 Hz, notch, common average, 7 plug-ins, z-score


# name -> (channels, weak scaling?, sfreq, seconds, settings factory, description).  weak: `channels` per GPU (the recording grows
# with N, common average over ALL channels); strong: `channels` in total, split N ways (BASELINE.json configs[3..4]).
CONFIGS = {
    "c3": (256, True, 1000.0, 300, c3_settings, "C3 (BASELINE.json configs[2]): notch+CAR, FFT+bandpass+Hjorth+linelength"),
    "c4": (256, False, 1000.0, 300, c4_settings, "C4 (configs[3]): 256 ch split over the GPUs, notch+CAR, oscillatory (FFT+Welch+STFT) + bursts + sharp waves"),
    "c5": (1024, False, 2000.0, 600, default_settings, "C5 (configs[4]): 1024 ch x 600 s @ 2 kHz split over the GPUs, untouched default settings (resample to 1 kHz, notch, CAR, "
                                                       "raw Hjorth, raw, FFT, Welch, sharp waves, bursts, line length, z-score)"),
    "default": (256, True, 1000.0, 300, default_settings, "reference default settings (all seven default plug-ins + feature normaliser)"),
}


def synth_rows(row0: int, n_rows: int, n_samples: int, seed: int, out: np.ndarray | None = None) -> np.ndarray:
    """Rows [row0, row0 + n_rows) of the GLOBAL synthetic recording (uniform [0, 1) float32, README.rst:83).  Every global channel
    has its own generator, so any rank -- and the parity check on rank 0 -- can reproduce any channel's first samples."""
    if out is None:
        out = np.empty((n_rows, n_samples), dtype=np.float32)
    for r in range(n_rows):
        np.random.default_rng([seed, row0 + r]).random(out=out[r], dtype=np.float32)
    return out


def run_gpu_arm(args) -> None:
    import py_neuromodulation_b200 as nm
    from py_neuromodulation_b200 import _lib
    from py_neuromodulation_b200.parallel import NativeComm, ShardedRun, car_shard_factorization, merge_permutation, shard_bounds
    from py_neuromodulation_b200.stream.generator import window_grid
    from py_neuromodulation_b200.utils.channels import get_default_channels_from_data

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE {world}"
    lib = _lib.load()
    # collectives live in libnmb200 (nm_comm_*: NCCL, dlopen'ed); torch only supplies the TCP store of the rendezvous
    comm = NativeComm.from_env(device=local_rank) if world > 1 else None

    ch_cfg, weak, sfreq, dur, make_settings, label = CONFIGS[args.config]
    if args.channels:  # (debugging / sweeps: not the BASELINE.json configuration any more)
        ch_cfg, label = args.channels, label + f" [channels overridden: {args.channels}]"
    settings = make_settings()
    n_samples = int(dur * sfreq)
    c_total = ch_cfg * world if weak else ch_cfg
    lo, hi = shard_bounds(c_total, world, rank)
    n_loc = hi - lo
    x = pinned_array(lib, (n_loc, n_samples), np.float32)
    synth_rows(lo, n_loc, n_samples, seed=args.seed, out=x)

    channels = get_default_channels_from_data(np.empty((c_total, 1)))
    local_channels = channels.iloc[lo:hi].reset_index(drop=True)
    reref = None
    if world > 1:
        reref = car_shard_factorization(list(channels["type"]), list(channels["status"]), list(channels["rereference"]), lo, hi)
    dp = nm.DataProcessor(sfreq=sfreq, settings=settings, channels=local_channels, line_noise=LINE_NOISE, verbose=False,
                          device=local_rank, reref_factored=reref)
    starts, lengths, _ = window_grid(n_samples, sfreq, settings.sampling_rate_features_hz, settings.segment_length_features_ms)
    W = int(lengths[0])
    plan = dp.plan(W)
    pipe = plan.pipe
    n_win, F = int(starts.size), pipe.F
    out = pinned_array(lib, (n_win, F), np.float64)
    sharded = ShardedRun(pipe, on_gpu=True, comm=comm) if world > 1 else None
    stateful = dp.stateful

    def barrier() -> None:
        pipe.synchronize()
        if comm is not None:
            comm.barrier()

    def step_e2e() -> None:
        """Public call with HOST buffers: H2D of the recording, all kernels, D2H of the feature matrix."""
        if stateful:
            pipe.reset_state()
        if sharded is None:
            pipe.upload(x)
            pipe.run(starts, out=out)
        else:
            sharded.upload(x)   # nm_upload_sharded_f32: sliced H2D, per-slice sums + ncclAllReduce inside the library
            sharded.run(starts)
            sharded.gather(n_win)

    def step_resident() -> None:
        """Recording already in HBM: window-independent preprocessing (N > 1: group sums + ncclAllReduce of the common average)
        + all window kernels; results stay on the device."""
        if stateful:
            pipe.reset_state()
        if sharded is None:
            pipe.prepare_resident()
        else:
            _lib.check(lib.nm_prepare_resident_sharded(pipe._h, comm._h))
        pipe.run(starts, download=False)

    def timed(fn, n: int) -> float:
        """CUDA events on the pipeline's stream around n steps (both sides behind a barrier + device sync); max over ranks."""
        barrier()
        t0 = time.perf_counter()
        pipe.timer_start()
        for _ in range(n):
            fn()
        ms = pipe.timer_stop()
        pipe.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        # e2e steps also spend time on the copy stream / in host-side waits after the last kernel: take the larger figure
        local = max(ms, wall) if fn is step_e2e else ms
        return comm.max(local) if comm is not None else local

    for _ in range(max(args.warmup, 3)):
        step_e2e()
    with ClockSampler(local_rank) as clocks:
        launches0 = pipe.kernel_launches
        coll0 = comm.collectives if comm is not None else 0
        ms_res = timed(step_resident, args.steps)
        launches = pipe.kernel_launches - launches0
        collectives = (comm.collectives - coll0) if comm is not None else 0
        ms_e2e = timed(step_e2e, args.steps)
    units = n_win * (world if weak else 1) * args.steps
    value = units / (ms_res * 1e-3)
    e2e_value = units / (ms_e2e * 1e-3)

    # ---- parity of the gathered matrix against the oracle on the first windows (after the timed region)
    parity = None
    n_chk = 4 if args.config != "c5" else 2
    gathered = sharded.gather(n_win) if sharded is not None else out
    if rank == 0 and not args.no_parity:
        from oracle import np_oracle as orc

        t_chk = int(starts[n_chk - 1] + W)
        xg = synth_rows(0, c_total, t_chk, seed=args.seed).astype(np.float64)
        ref_cols, ref = orc.run_offline(xg, sfreq, settings.model_dump(), line_noise=LINE_NOISE, max_windows=n_chk)
        if world > 1:
            cols, perm = merge_permutation(settings, list(channels["new_name"]), dp.sfreq_raw, dp.feature_window(W), world)
            got = np.asarray(gathered)[:n_chk][:, perm]
        else:
            cols, got = plan.columns, np.asarray(gathered)[:n_chk]
        assert ref_cols[: len(cols)] == list(cols), "column order differs from the oracle"
        ref = ref[:n_chk, : len(cols)]
        fin = np.isfinite(ref)
        err = float(np.max(np.abs(np.where(fin, got - ref, 0.0)) / np.maximum(np.abs(np.where(fin, ref, 1.0)), 1.0)))
        same_special = bool(np.array_equal(np.isnan(got), np.isnan(ref)) and np.array_equal(got[~fin & ~np.isnan(ref)], ref[~fin & ~np.isnan(ref)]))
        parity = {"parity_checked": bool(err < 1e-5 and same_special), "windows": n_chk, "columns": len(cols), "max_err": err,
                  "rule": "|got - oracle| <= 1e-5 * max(|oracle|, 1), NaN / inf patterns equal; oracle = oracle/np_oracle.py on the un-sharded recording"}

    # ---- N = 1 extras: per-kernel profile -> roofline of the dominant kernel, float32 mode, zero-overlap run, Stream.run e2e
    roof, f32_info, overlap_info, stream_info = None, None, None, None
    if world == 1:
        pipe.set_profiling(True)
        step_resident()
        pipe.synchronize()
        prof = pipe.profile()
        pipe.set_profiling(False)
        dominant = max(prof, key=lambda k: prof[k][0])
        dom_ms, dom_launches = prof[dominant]
        per_ch = F // max(n_loc, 1)
        feats_of = {"notch": 0, "scan": 4 * n_loc, "spectral": 4 * n_loc, "bandpower": 4 * n_loc, "prep": 0, "fused": F}
        unit_bytes = n_loc * W * 4 + feats_of.get(dominant, per_ch * n_loc) * 4  # fp32 tile in + fp32 features out
        bytes_per_launch = unit_bytes * n_win / max(dom_launches, 1)
        achieved = bytes_per_launch / (dom_ms / max(dom_launches, 1) * 1e-3) / 1e9
        peak, peak_src = measured_peak_gbs()
        traffic, ncu_info = ncu_traffic(dominant, n_win / max(dom_launches, 1))
        step_bytes_windowed = n_win * (n_loc * W * 4 + F * 4)
        step_bytes_unique = n_loc * n_samples * 4 + n_win * F * 4
        roof = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "algorithmic_bytes_per_launch": bytes_per_launch, "ncu": ncu_info, "peak_source": peak_src,
                "launches_per_step": int(dom_launches), "ms_per_launch": dom_ms / max(dom_launches, 1),
                "profile_ms_per_step": {k: round(v[0], 3) for k, v in prof.items()},
                "whole_step": {"windowed_bytes": step_bytes_windowed, "unique_bytes": step_bytes_unique,
                               "windowed_gbs": step_bytes_windowed / (ms_res / args.steps * 1e-3) / 1e9,
                               "unique_gbs": step_bytes_unique / (ms_res / args.steps * 1e-3) / 1e9,
                               "frac_of_measured_peak": step_bytes_windowed / (ms_res / args.steps * 1e-3) / 1e9 / peak,
                               "frac_of_nominal_8000": step_bytes_windowed / (ms_res / args.steps * 1e-3) / 1e9 / 8000.0},
                "note": "FFT-convolution kernels are FP64-pipe / shared-memory bound, not HBM bound (DESIGN.md section 5): "
                        "the honest utilisation figure is ncu.fp64_pipe_pct; windows overlap 90 %, so windowed bytes exceed unique bytes 10x"}
        if args.config == "c3" and not args.quick:
            # zero-overlap variant: feature rate 1 Hz -> stride = W, windowed bytes == unique bytes
            s1 = make_settings()
            s1.sampling_rate_features_hz = 1
            dp1 = nm.DataProcessor(sfreq=sfreq, settings=s1, channels=local_channels, line_noise=LINE_NOISE, verbose=False, device=local_rank)
            st1, _, _ = window_grid(n_samples, sfreq, 1, s1.segment_length_features_ms)
            pipe1 = dp1.plan(W).pipe
            pipe1.upload(x)
            for _ in range(3):
                pipe1.run(st1, download=False)
            pipe1.synchronize()
            pipe1.timer_start()
            for _ in range(args.steps):
                pipe1.prepare_resident()
                pipe1.run(st1, download=False)
            ms1 = pipe1.timer_stop() / args.steps
            b1 = int(st1.size) * (n_loc * W * 4 + F * 4)
            overlap_info = {"windows": int(st1.size), "ms_per_step": ms1, "value": st1.size / (ms1 * 1e-3), "unit": UNIT,
                            "gbs": b1 / (ms1 * 1e-3) / 1e9, "note": "feature rate 1 Hz: stride == window, no sample is read twice"}
            pipe1.close()
            # optional float32 modes of the linear FIR families (same workload, resident timing; reported next to the float64
            # headline): scalar float32 and packed float32 pairs (Blackwell f32x2 arithmetic, two channel pairs per item)
            ref64 = pipe.run(starts[:64])
            f32_info = {}
            for prec in ("f32", "f32x2"):
                dp32 = nm.DataProcessor(sfreq=sfreq, settings=settings, channels=local_channels, line_noise=LINE_NOISE, verbose=False,
                                        device=local_rank, precision=prec)
                pipe32 = dp32.plan(W).pipe
                pipe32.upload(x)
                for _ in range(3):
                    pipe32.run(starts, download=False)
                pipe32.synchronize()
                pipe32.timer_start()
                for _ in range(args.steps):
                    pipe32.prepare_resident()
                    pipe32.run(starts, download=False)
                ms32 = pipe32.timer_stop()
                got32 = pipe32.run(starts[:64])
                err32 = float(np.max(np.abs(got32 - ref64) / np.maximum(np.abs(ref64), 1.0)))
                f32_info[prec] = {"value": n_win * args.steps / (ms32 * 1e-3), "unit": UNIT, "ms_per_step": ms32 / args.steps,
                                  "max_err_vs_f64_64_windows": err32}
                pipe32.close()
            f32_info["note"] = ("nm_set_precision(1 | 2): float32 inside the notch / band-pass FFT convolutions, moments and outputs float64; "
                                "max_err = |f32 - f64| / max(|f64|, 1) over the first 64 windows of this recording (worst entries: log10 FFT band "
                                "amplitudes behind the float32 notch); fixtures gate: tests/test_parity_pipeline.py::"
                                "test_float32_linear_mode_within_north_star_tolerance")
        # the call a reference user makes: nm.Stream(...).run(data) -> DataFrame (window grid, upload, kernels, download, frame)
        import tempfile

        if not args.quick:
            reps = 3
            x_user = np.array(x)  # what a reference user holds: an ordinary (pageable) numpy array, not the page-locked bench buffer
            stream = nm.Stream(sfreq=sfreq, data=x_user, settings=settings, line_noise=LINE_NOISE, verbose=False)
            with tempfile.TemporaryDirectory() as td:
                stream.run(out_dir=td, experiment_name="bench", save_csv=False)  # first call: builds the GPU plans (warm-up)
                ts = []
                for _ in range(reps):
                    t0 = time.perf_counter()
                    df = stream.run(out_dir=td, experiment_name="bench", save_csv=False)
                    ts.append(time.perf_counter() - t0)
            t_med = float(np.median(ts))
            stream_info = {"value": n_win / t_med, "unit": UNIT, "ms_per_step": 1e3 * t_med, "frame_shape": list(df.shape),
                           "note": "nm.Stream.run(data, save_csv=False) wall clock, median of 3 after one warm-up call: window grid, H2D from the "
                                   "caller's PAGEABLE array (staged slice by slice through page-locked buffers by a host thread of the library, "
                                   "overlapping the kernels), kernels, rows of finished chunks through a page-locked buffer into the final "
                                   "(pageable) table, pandas DataFrame, side files; the processor of an unchanged configuration is kept across "
                                   "calls; CSV writing excluded"}

    if rank == 0:
        tile = f"{n_loc}x{W}"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak" if weak else "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"{label}; {c_total} ch x {dur} s @ {int(sfreq)} Hz -> {n_win} windows per step, {n_loc} ch per GPU "
                            f"(tile {tile} per GPU), F = {F} per GPU",
                "name": args.config,
                "unit_definition": f"one window of all {c_total} channels with all its features" if not weak else
                                   "one 256 ch x 1000 samp window with all its features (per GPU; the recording grows with N)",
                "sharding": "channels" if world > 1 else "none",
                "collectives": "libnmb200 nm_comm (NCCL): per-step ncclAllReduce of the common-average sums inside `value`; e2e adds the per-slice "
                               "all-reduce of the pipelined upload; results meet in one page-locked shared host matrix (no collective)" if world > 1 else "none",
                "l2": f"inputs larger than L2 (raw {x.nbytes / 1e6:.0f} MB f32 per GPU)",
                "timing": "CUDA events on the pipeline stream, barrier + device sync on both sides, max over ranks (nm_comm_allreduce_max)",
            },
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int((x.nbytes + starts.nbytes) * world),
                    "d2h_bytes_per_step": int(out.nbytes * world), "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "nccl_collectives_in_value": int(collectives),
            "clocks": clocks.summary(),
        }
        if parity is not None:
            line["parity"] = parity
            line["parity_checked"] = parity["parity_checked"]
        if roof is not None:
            line["roofline"] = roof
        if f32_info is not None:
            line["f32_linear_mode"] = f32_info
        if overlap_info is not None:
            line["zero_overlap"] = overlap_info
        if stream_info is not None:
            line["e2e_stream"] = stream_info
        if world == 1 and not args.no_cpu_baseline and args.config == "c3":
            line["cpu_baseline"] = cpu_baseline_single()
        emit(line)
    if comm is not None:
        comm.barrier()
        sharded.close()
        comm.close()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--channels", type=int, default=0, help="override the configuration's channel count (sweeps / debugging)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--quick", action="store_true", help="skip the N = 1 extras (zero-overlap run, float32 mode, Stream.run timing)")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
