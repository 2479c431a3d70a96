"""ctypes binding of ``libnmb200.so`` (C ABI declared in ``include/nmb200.h``).

There is exactly one compute backend: the CUDA library built in-tree under ``csrc/``.  If it
is missing, cannot be loaded, or no CUDA device is visible, every compute entry point raises
-- there is no CPU fallback of any kind in this package.
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "csrc" / "libnmb200.so"

_LIB: C.CDLL | None = None

c_int_p = C.POINTER(C.c_int)
c_double_p = C.POINTER(C.c_double)
c_float_p = C.POINTER(C.c_float)
c_ll_p = C.POINTER(C.c_longlong)


class SpectralCfg(C.Structure):
    """Mirror of ``nm_spectral_cfg``."""

    _fields_ = [
        ("nper", C.c_int), ("nseg", C.c_int), ("hop", C.c_int), ("start", C.c_int),
        ("ext_even", C.c_int), ("ext_len", C.c_int),
        ("detrend", C.c_int),
        ("power", C.c_int),
        ("scale", C.c_double),
        ("log", C.c_int),
        ("keep_segments", C.c_int),
        ("n_bands", C.c_int),
        ("est_mask", C.c_int),
        ("want_spectrum", C.c_int),
        ("win", c_double_p),
        ("band_lo", c_int_p),
        ("band_hi", c_int_p),
        ("colmap", c_int_p),
    ]


_SIGNATURES = {
    "nm_last_error": (C.c_char_p, []),
    "nm_abi_version": (C.c_int, []),
    "nm_device_count": (C.c_int, [c_int_p]),
    "nm_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_longlong]),
    "nm_host_free": (C.c_int, [C.c_void_p]),
    "nm_pipeline_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "nm_pipeline_destroy": (None, [C.c_void_p]),
    "nm_finalize": (C.c_int, [C.c_void_p]),
    "nm_reset_state": (C.c_int, [C.c_void_p]),
    "nm_set_pick": (C.c_int, [C.c_void_p, c_int_p]),
    "nm_set_reref": (C.c_int, [C.c_void_p, C.c_int, c_int_p, c_double_p, c_int_p, c_int_p, c_double_p]),
    "nm_set_notch": (C.c_int, [C.c_void_p, c_double_p, C.c_int]),
    "nm_set_precision": (C.c_int, [C.c_void_p, C.c_int]),
    "nm_set_fused": (C.c_int, [C.c_void_p, C.c_int]),
    "nm_set_raw_normalizer": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int]),
    "nm_add_prefilter": (C.c_int, [C.c_void_p, c_double_p, C.c_int]),
    "nm_set_nan_columns": (C.c_int, [C.c_void_p, c_int_p, c_int_p]),
    "nm_add_scan": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_int_p]),
    "nm_add_spectral": (C.c_int, [C.c_void_p, C.POINTER(SpectralCfg)]),
    "nm_add_bandpower": (C.c_int, [C.c_void_p, C.c_int, c_double_p, C.c_int, c_int_p, C.c_int, C.c_int, C.c_int, C.c_int, c_int_p]),
    "nm_add_bursts": (C.c_int, [C.c_void_p, C.c_int, c_double_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, c_int_p]),
    "nm_add_sharpwave": (C.c_int, [C.c_void_p, C.c_int, c_double_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                                   c_int_p, c_int_p, C.c_int, C.c_int, c_int_p]),
    "nm_add_feature_normalizer": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int, c_int_p]),
    "nm_upload_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong]),
    "nm_upload_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong]),
    "nm_run_windows": (C.c_int, [C.c_void_p, c_ll_p, C.c_int, C.c_void_p]),
    "nm_set_output_pitch": (C.c_int, [C.c_void_p, C.c_longlong]),
    "nm_host_register": (C.c_int, [C.c_void_p, C.c_longlong]),
    "nm_host_unregister": (C.c_int, [C.c_void_p]),
    "nm_download": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "nm_process_window": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "nm_preprocess_window": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "nm_fir_apply": (C.c_int, [C.c_int, c_double_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "nm_timer_start": (C.c_int, [C.c_void_p]),
    "nm_timer_stop": (C.c_int, [C.c_void_p, c_double_p]),
    "nm_kernel_launches": (C.c_longlong, [C.c_void_p]),
    "nm_prepare_resident": (C.c_int, [C.c_void_p]),
    "nm_synchronize": (C.c_int, [C.c_void_p]),
    "nm_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "nm_get_profile": (C.c_int, [C.c_void_p, c_double_p, c_ll_p, C.c_int]),
    "nm_chunk_windows": (C.c_int, [C.c_void_p]),
    "nm_set_burst_threshold_mode": (C.c_int, [C.c_void_p, C.c_int]),
    "nm_burst_threshold_stats": (C.c_int, [C.c_void_p, c_ll_p, c_ll_p]),
    "nm_describe_plan": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "nm_result_device_ptr": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), c_ll_p, c_int_p]),
    "nm_stream_handle": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "nm_upload_begin_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong]),
    "nm_group_sums_device_ptr": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), c_ll_p]),
    "nm_upload_slices": (C.c_int, [C.c_void_p, c_int_p, c_ll_p, c_int_p, c_ll_p]),
    "nm_side_stream_handle": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "nm_set_resampler": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_int]),
    "nm_upload_slice_sums": (C.c_int, [C.c_void_p, C.c_int]),
    "nm_upload_slice_reduced": (C.c_int, [C.c_void_p, C.c_int]),
    "nm_upload_finish": (C.c_int, [C.c_void_p]),
    "nm_stream_open": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "nm_stream_input": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_longlong)]),
    "nm_stream_submit": (C.c_int, [C.c_void_p, C.c_int]),
    "nm_stream_wait": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "nm_stream_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "nm_stream_close": (C.c_int, [C.c_void_p]),
    "nm_comm_unique_id": (C.c_int, [C.c_void_p]),
    "nm_comm_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "nm_comm_destroy": (None, [C.c_void_p]),
    "nm_comm_rank": (C.c_int, [C.c_void_p]),
    "nm_comm_size": (C.c_int, [C.c_void_p]),
    "nm_comm_collectives": (C.c_longlong, [C.c_void_p]),
    "nm_comm_barrier": (C.c_int, [C.c_void_p]),
    "nm_comm_allreduce_max": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "nm_upload_sharded_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong]),
    "nm_gather_results": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_void_p]),
    "nm_prepare_resident_sharded": (C.c_int, [C.c_void_p, C.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def declare(lib: C.CDLL) -> C.CDLL:
    """Attach restype/argtypes for every symbol of the C ABI (raises AttributeError if one is missing)."""
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


def load() -> C.CDLL:
    """Return the loaded CUDA library; raise loudly if it is not available."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not LIB_PATH.is_file():
        raise ImportError(
            f"{LIB_PATH} not found: build it with `make -C {LIB_PATH.parent}` (or `python -c 'import __graft_entry__ as g; "
            "g.build()'`).  py_neuromodulation_b200 has no CPU fallback."
        )
    try:
        lib = C.CDLL(str(LIB_PATH))
    except OSError as exc:  # e.g. libcudart missing
        raise ImportError(f"could not load {LIB_PATH}: {exc}.  py_neuromodulation_b200 has no CPU fallback.") from exc
    declare(lib)
    if lib.nm_abi_version() != 1:
        raise ImportError(f"{LIB_PATH} has ABI version {lib.nm_abi_version()}, expected 1: rebuild it")
    _LIB = lib
    return lib


def last_error() -> str:
    msg = load().nm_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError(f"libnmb200: {last_error()}")


def device_count() -> int:
    n = C.c_int(0)
    check(load().nm_device_count(C.byref(n)))
    return n.value
