"""Zero-phase windowed-sinc FIR design on the host.

The reference obtains its filters from ``mne.filter.create_filter(method="fir", phase="zero",
fir_window="hamming", fir_design="firwin")`` (call sites: ``filter/mne_filter.py:53-73``,
``filter/notch_filter.py:62-76``, ``features/sharpwaves.py:127-142``).  When MNE is importable it
is used directly; otherwise :func:`design_fir` builds the same filter from MNE's documented
recipe: a piecewise-constant gain specification is realised as a signed sum of Hamming-windowed
sinc low-passes, one per gain step, each as long as its own transition band requires.

Taps are computed once per pipeline, in float64, and handed to the GPU as data.
"""

from __future__ import annotations

import math

import numpy as np

_WINDOW_FACTOR = {"hann": 3.1, "hamming": 3.3, "blackman": 5.0}


def _windowed_sinc_lowpass(n_taps: int, cutoff: float) -> np.ndarray:
    """Unit-DC-gain Hamming low-pass; ``cutoff`` is relative to Nyquist (same as scipy.signal.firwin(fs=2))."""
    k = np.arange(n_taps) - (n_taps - 1) / 2.0
    taps = cutoff * np.sinc(cutoff * k)
    taps *= 0.54 - 0.46 * np.cos(2.0 * np.pi * np.arange(n_taps) / (n_taps - 1))
    return taps / taps.sum()


def _length_in_samples(filter_length, sfreq: float) -> int:
    if isinstance(filter_length, str):
        text = filter_length.strip().lower()
        if text.endswith("ms"):
            seconds = float(text[:-2]) * 1e-3
        elif text.endswith("s"):
            seconds = float(text[:-1])
        else:
            raise ValueError(f"filter_length must end in 's' or 'ms', got {filter_length!r}")
        n = max(int(math.ceil(seconds * sfreq)), 1)
    else:
        if int(filter_length) != filter_length:
            raise TypeError("filter_length must be an integer number of samples or a string")
        n = int(filter_length)
    return n if n % 2 == 1 else n + 1


def _auto_bandwidth(freq, upper):
    return np.minimum(np.maximum(0.25 * freq, 2.0), upper)


def _realise(n_taps: int, edges: list[tuple[float, float]], nyq: float, window: str) -> np.ndarray:
    """edges: ascending (frequency Hz, gain in {0,1}) break points from 0 to Nyquist."""
    taps = np.zeros(n_taps)
    if edges[-1][1] == 1:
        taps[n_taps // 2] = 1.0
    # Frequencies are normalised to Nyquist BEFORE differences are taken: the required low-pass length is a
    # round() of 3.3 / half-width and sits on .5 for common specifications (e.g. 2 Hz at sfreq 150), so the
    # order of the floating-point operations decides the filter length and must follow MNE's.
    rel = [(f / nyq, g) for f, g in edges]
    # walk from Nyquist towards DC; every gain change adds or removes one low-pass
    for (f_hi, g_hi), (f_lo, g_lo) in zip(rel[:0:-1], rel[-2::-1]):
        if g_lo == g_hi:
            continue
        half_width = (f_hi - f_lo) / 2.0
        need = int(round(_WINDOW_FACTOR[window] / half_width))
        need += 1 - need % 2
        if need > n_taps:
            raise ValueError(
                f"The requested filter length {n_taps} is too short for the requested {half_width * nyq:0.2f} Hz "
                f"transition band, which requires {need} samples"
            )
        lp = _windowed_sinc_lowpass(need, (f_hi + f_lo) / 2.0)
        pad = (n_taps - need) // 2
        if g_lo == 1:
            taps[pad : n_taps - pad] += lp
        else:
            taps[pad : n_taps - pad] -= lp
    return taps


def design_fir(sfreq: float, l_freq, h_freq, filter_length="auto", l_trans_bandwidth="auto", h_trans_bandwidth="auto",
               window: str = "hamming", use_mne: bool = True) -> np.ndarray:
    """Low-pass (``l_freq is None``), high-pass, band-pass (``l < h``) or multi-band band-stop (arrays, ``l > h``)."""
    if use_mne:
        try:
            from mne.filter import create_filter  # type: ignore

            return create_filter(None, sfreq, l_freq=l_freq, h_freq=h_freq, filter_length=filter_length,
                                 l_trans_bandwidth=l_trans_bandwidth, h_trans_bandwidth=h_trans_bandwidth, method="fir",
                                 phase="zero", fir_window=window, fir_design="firwin", verbose=False)
        except ImportError:
            pass

    sfreq = float(sfreq)
    nyq = sfreq / 2.0
    lo = None if l_freq is None else np.atleast_1d(np.asarray(l_freq, dtype=float))
    hi = None if h_freq is None else np.atleast_1d(np.asarray(h_freq, dtype=float))
    if hi is not None and (hi > nyq).any():
        raise ValueError(f"h_freq ({hi}) must be less than the Nyquist frequency {nyq}")
    if lo is not None and (lo == 0).all():
        lo = None
    band_stop = lo is not None and hi is not None and not (lo < hi).any()

    widths = []
    if band_stop:
        if len(lo) != len(hi):
            raise ValueError("l_freq and h_freq must be the same length")
        # for a band-stop the roles are swapped: `hi` are the lower corner frequencies, `lo` the upper ones
        low_corner, high_corner = hi, lo
        if np.any(low_corner <= 0) or np.any(high_corner >= nyq):
            raise ValueError("band-stop corner frequencies must lie strictly between 0 and Nyquist")
        tb_low = _auto_bandwidth(low_corner, low_corner) if isinstance(h_trans_bandwidth, str) else np.full_like(low_corner, float(h_trans_bandwidth))
        tb_high = _auto_bandwidth(high_corner, nyq - high_corner) if isinstance(l_trans_bandwidth, str) else np.full_like(high_corner, float(l_trans_bandwidth))
        if np.any(tb_low <= 0) or np.any(tb_high <= 0):
            raise ValueError("transition bandwidths must be positive")
        pass_lo, stop_lo = low_corner, low_corner + tb_low
        stop_hi, pass_hi = high_corner - tb_high, high_corner
        if np.any(pass_lo < 0) or np.any(pass_hi > nyq):
            raise ValueError("Filter specification invalid: band-stop edge outside [0, Nyquist]")
        points = sorted([(f, 1.0) for f in pass_lo] + [(f, 0.0) for f in stop_lo] + [(f, 0.0) for f in stop_hi]
                        + [(f, 1.0) for f in pass_hi])
        edges = [(0.0, 1.0)] + points + [(nyq, 1.0)]
        gains = np.array([g for _, g in edges])
        if np.any(np.abs(np.diff(gains, 2)) > 1):
            raise ValueError("Stop bands are not sufficiently separated.")
        widths = [float(tb_low.min()), float(tb_high.min())]
    else:
        edges = []
        l_stop = h_stop = None
        if lo is not None:
            l = float(lo.item())
            if l <= 0:
                raise ValueError(f"highpass frequency {l} must be greater than zero")
            tb = float(_auto_bandwidth(l, l)) if isinstance(l_trans_bandwidth, str) else float(l_trans_bandwidth)
            if tb <= 0:
                raise ValueError("l_trans_bandwidth must be positive")
            l_stop = l - tb
            if l_stop < 0:
                raise ValueError(f"Filter specification invalid: Lower stop frequency negative ({l_stop:0.2f} Hz).")
            widths.append(tb)
        if hi is not None:
            h = float(hi.item())
            if h >= nyq:
                raise ValueError(f"lowpass frequency {h} must be less than Nyquist ({nyq})")
            tb = float(_auto_bandwidth(h, nyq - h)) if isinstance(h_trans_bandwidth, str) else float(h_trans_bandwidth)
            if tb <= 0:
                raise ValueError("h_trans_bandwidth must be positive")
            h_stop = h + tb
            if h_stop > nyq:
                raise ValueError(f"Effective band-stop frequency ({h_stop}) is too high (maximum based on Nyquist is {nyq})")
            widths.append(tb)
        if lo is None and hi is None:
            edges = [(0.0, 1.0), (nyq, 1.0)]
        elif lo is None:
            edges = [(0.0, 1.0), (h, 1.0), (h_stop, 0.0)] + ([(nyq, 0.0)] if h_stop != nyq else [])
        elif hi is None:
            edges = ([(0.0, 0.0)] if l_stop != 0 else []) + [(l_stop, 0.0), (l, 1.0), (nyq, 1.0)]
        else:
            edges = ([(0.0, 0.0)] if l_stop != 0 else []) + [(l_stop, 0.0), (l, 1.0), (h, 1.0), (h_stop, 0.0)] \
                + ([(nyq, 0.0)] if h_stop != nyq else [])

    if isinstance(filter_length, str) and filter_length.lower() == "auto":
        filter_length = f"{_WINDOW_FACTOR[window] / min(widths) if widths else 0.0}s"
    n_taps = _length_in_samples(filter_length, sfreq)
    if edges[0][0] != 0 or edges[-1][0] != nyq:
        raise ValueError("gain specification must span 0 .. Nyquist")
    return _realise(n_taps, edges, nyq, window)
