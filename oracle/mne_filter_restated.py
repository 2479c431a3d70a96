"""TEST INFRASTRUCTURE (oracle) -- restatement of the subset of ``mne.filter`` the reference calls.

``mne`` is an *un-vendored, unpinned* third-party dependency of the reference
(``pyproject.toml:36``; not installed in this image, no network).  The reference's hot
path calls three private/public functions of it:

* ``mne.filter.create_filter``       -- ``filter/mne_filter.py:44,53-73``,
                                         ``filter/notch_filter.py:17,62-76``,
                                         ``features/sharpwaves.py:127-142``
* ``mne.filter._overlap_add_filter`` -- ``filter/notch_filter.py:82-93``

This file restates MNE's published algorithm for exactly the argument patterns
used there (``method="fir"``, ``phase="zero"``, ``fir_window="hamming"``,
``fir_design="firwin"``).  PARITY UNPINNED against a real MNE install: the
reference repository stores no filter coefficients or filtered samples
(SURVEY.md section 8c); the only anchors are structural (odd length, symmetry,
DC / pass-band gain) and the known-answer values of SURVEY.md section 8c which the
tests check.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU baseline may
import this module.  The shipped package has its own FIR designer
(``py_neuromodulation_b200/filter/fir_design.py``).
"""

from __future__ import annotations

import numpy as np

_LENGTH_FACTORS = {"hann": 3.1, "hamming": 3.3, "blackman": 5.0}


# --------------------------------------------------------------------------- design
def _parse_length(filter_length, sfreq: float) -> int:
    """MNE ``_to_samples``: "999ms" / "1.65s" / int -> sample count (not yet odd)."""
    if isinstance(filter_length, str):
        txt = filter_length.lower()
        if txt.endswith("ms"):
            scale, num = 1e-3, txt[:-2]
        elif txt.endswith("s"):
            scale, num = 1.0, txt[:-1]
        else:
            raise ValueError(f"filter_length string must end in 's' or 'ms', got {filter_length}")
        n = max(int(np.ceil(float(num) * scale * sfreq)), 1)
        n += (n - 1) % 2
        return n
    if int(filter_length) != filter_length:
        raise TypeError("filter_length must be an integer or a string")
    return int(filter_length)


def _triage(sfreq, l_freq, h_freq, l_tb, h_tb, filter_length, fir_window, *, arr=False, reverse=False):
    """MNE ``_triage_filter_params`` for method="fir"."""
    cast = (lambda v: np.array(v, float).ravel()) if arr else float
    l_stop = h_stop = None
    if l_freq is not None:
        l_freq = cast(l_freq)
        if np.any(l_freq <= 0):
            raise ValueError(f"highpass frequency {l_freq} must be greater than zero")
    if h_freq is not None:
        h_freq = cast(h_freq)
        if np.any(h_freq >= sfreq / 2.0):
            raise ValueError(f"lowpass frequency {h_freq} must be less than Nyquist ({sfreq / 2.})")

    if l_freq is not None:
        if isinstance(l_tb, str):
            if l_tb != "auto":
                raise ValueError("l_trans_bandwidth must be 'auto' if string")
            l_tb = np.minimum(np.maximum(0.25 * l_freq, 2.0), l_freq)
        l_tb = cast(l_tb)
        if np.any(l_tb <= 0):
            raise ValueError("l_trans_bandwidth must be positive")
        l_stop = l_freq - l_tb
        if reverse:
            l_stop = l_stop + l_tb
            l_freq = l_freq + l_tb
        if np.any(l_stop < 0):
            raise ValueError("Filter specification invalid: Lower stop frequency negative")
    if h_freq is not None:
        if isinstance(h_tb, str):
            if h_tb != "auto":
                raise ValueError("h_trans_bandwidth must be 'auto' if string")
            h_tb = np.minimum(np.maximum(0.25 * h_freq, 2.0), sfreq / 2.0 - h_freq)
        h_tb = cast(h_tb)
        if np.any(h_tb <= 0):
            raise ValueError("h_trans_bandwidth must be positive")
        h_stop = h_freq + h_tb
        if reverse:
            h_stop = h_stop - h_tb
            h_freq = h_freq - h_tb
        if np.any(h_stop > sfreq / 2.0):
            raise ValueError("Effective band-stop frequency is too high")

    if isinstance(filter_length, str) and filter_length.lower() == "auto":
        h_chk = l_chk = np.inf
        if h_freq is not None:
            h_chk = float(np.min(np.atleast_1d(h_tb)))
        if l_freq is not None:
            l_chk = float(np.min(np.atleast_1d(l_tb)))
        filter_length = f"{_LENGTH_FACTORS[fir_window] / float(min(h_chk, l_chk))}s"
    n = _parse_length(filter_length, sfreq)
    n += (n - 1) % 2  # firwin / zero phase -> odd
    return l_freq, h_freq, l_stop, h_stop, n


def _firwin_sum(n_taps: int, freq: np.ndarray, gain: np.ndarray, window: str) -> np.ndarray:
    """MNE ``_firwin_design``: build the filter as a signed sum of windowed-sinc low-passes."""
    from scipy.signal import firwin

    assert freq[0] == 0 and len(freq) == len(gain) and n_taps % 2 == 1
    h = np.zeros(n_taps)
    if gain[-1] == 1:
        h[n_taps // 2] = 1.0
    prev_f, prev_g = freq[-1], gain[-1]
    for this_f, this_g in zip(freq[::-1][1:], gain[::-1][1:]):
        if this_g != prev_g:
            transition = (prev_f - this_f) / 2.0
            this_n = int(round(_LENGTH_FACTORS[window] / transition))
            this_n += 1 - this_n % 2
            if this_n > n_taps:
                raise ValueError(
                    f"The requested filter length {n_taps} is too short for the requested "
                    f"transition band, which requires {this_n} samples"
                )
            lp = firwin(this_n, (prev_f + this_f) / 2.0, window=window, pass_zero=True, fs=freq[-1] * 2)
            off = (n_taps - this_n) // 2
            if this_g == 0:
                h[off : n_taps - off] -= lp
            else:
                h[off : n_taps - off] += lp
        prev_g, prev_f = this_g, this_f
    return h


def create_filter(
    data,
    sfreq,
    l_freq,
    h_freq,
    filter_length="auto",
    l_trans_bandwidth="auto",
    h_trans_bandwidth="auto",
    method="fir",
    iir_params=None,
    phase="zero",
    fir_window="hamming",
    fir_design="firwin",
    verbose=None,
):
    """Restated ``mne.filter.create_filter`` (FIR / zero-phase / firwin only)."""
    if method != "fir" or phase != "zero" or fir_design != "firwin":
        raise NotImplementedError("oracle restates only method='fir', phase='zero', fir_design='firwin'")
    sfreq = float(sfreq)
    if sfreq < 0:
        raise ValueError("sfreq must be positive")
    nyq = sfreq / 2.0
    if h_freq is not None:
        h_freq = np.array(h_freq, float).ravel()
        if (h_freq > nyq).any():
            raise ValueError(f"h_freq ({h_freq}) must be less than the Nyquist frequency {nyq}")
    if l_freq is not None:
        l_freq = np.array(l_freq, float).ravel()
        if (l_freq == 0).all():
            l_freq = None

    if l_freq is None and h_freq is None:
        _, _, _, _, n = _triage(sfreq, None, None, None, None, filter_length, fir_window)
        freq, gain = [0.0, nyq], [1.0, 1.0]
    elif l_freq is None:
        h = h_freq.item()
        _, f_p, _, f_s, n = _triage(sfreq, None, h, None, h_trans_bandwidth, filter_length, fir_window)
        freq, gain = [0.0, f_p, f_s], [1.0, 1.0, 0.0]
        if f_s != nyq:
            freq.append(nyq)
            gain.append(0.0)
    elif h_freq is None:
        lo = l_freq.item()
        pass_, _, stop, _, n = _triage(sfreq, lo, None, l_trans_bandwidth, None, filter_length, fir_window)
        freq, gain = [stop, pass_, nyq], [0.0, 1.0, 1.0]
        if stop != 0:
            freq.insert(0, 0.0)
            gain.insert(0, 0.0)
    elif (l_freq < h_freq).any():
        lo, hi = l_freq.item(), h_freq.item()
        f_p1, f_p2, f_s1, f_s2, n = _triage(
            sfreq, lo, hi, l_trans_bandwidth, h_trans_bandwidth, filter_length, fir_window
        )
        freq, gain = [f_s1, f_p1, f_p2, f_s2], [0.0, 1.0, 1.0, 0.0]
        if f_s2 != nyq:
            freq.append(nyq)
            gain.append(0.0)
        if f_s1 != 0:
            freq.insert(0, 0.0)
            gain.insert(0, 0.0)
    else:
        if len(l_freq) != len(h_freq):
            raise ValueError("l_freq and h_freq must be the same length")
        # band-stop: roles of l/h swapped on purpose, "reverse" triage
        f_s1, f_s2, f_p1, f_p2, n = _triage(
            sfreq, h_freq, l_freq, h_trans_bandwidth, l_trans_bandwidth, filter_length, fir_window,
            arr=True, reverse=True,
        )
        freq = np.r_[f_p1, f_s1, f_s2, f_p2]
        gain = np.r_[np.ones_like(f_p1), np.zeros_like(f_s1), np.zeros_like(f_s2), np.ones_like(f_p2)]
        order = np.argsort(freq)
        freq = np.r_[0.0, freq[order], nyq]
        gain = np.r_[1.0, gain[order], 1.0]
        if np.any(np.abs(np.diff(gain, 2)) > 1):
            raise ValueError("Stop bands are not sufficiently separated.")

    freq = np.asarray(freq, float) / nyq
    gain = np.asarray(gain, float)
    if freq[0] != 0 or freq[-1] != 1:
        raise ValueError("freq must start at 0 and end at Nyquist")
    if n % 2 == 0:
        raise RuntimeError(f'filter_length must be odd if phase="zero", got {n}')
    return _firwin_sum(n, freq, gain, fir_window)


# --------------------------------------------------------------------------- application
def _reflect_limited(x: np.ndarray, n_edge: int) -> np.ndarray:
    """MNE ``_smart_pad(..., pad="reflect_limited")``: odd reflection about both end points."""
    if n_edge == 0:
        return x
    lz = np.zeros(max(n_edge - len(x) + 1, 0), dtype=x.dtype)
    return np.concatenate([lz, 2 * x[0] - x[n_edge:0:-1], x, 2 * x[-1] - x[-2 : -n_edge - 2 : -1], lz])


def _overlap_add_filter(x, h, n_fft=None, phase="zero", picks=None, n_jobs=1, copy=True, pad="reflect_limited"):
    """Restated ``mne.filter._overlap_add_filter`` for phase="zero".

    Numerically this is the centred linear convolution of the reflect-limited
    extension.  Like MNE it is evaluated row by row with
    overlap-add block FFTs whose power-of-two length is picked by MNE's
    multiplication-count cost model, so its cost is representative when this
    module serves as the CPU baseline.
    """
    from scipy.fft import rfft, irfft, next_fast_len

    x = np.asarray(x)
    if x.dtype != np.float64:
        raise TypeError("Arrays passed for filtering must have a dtype of np.float64")
    if phase != "zero" or pad != "reflect_limited":
        raise NotImplementedError
    one_d = x.ndim == 1
    x2 = np.atleast_2d(x)
    out = x2.copy() if copy else x2
    n_h = len(h)
    if n_h % 2 == 0:
        raise RuntimeError('filter_length must be odd if phase="zero"')
    if n_h == 1:
        out[...] = x2 * h
        return out[0] if one_d else out
    w = x2.shape[1]
    n_edge = max(min(n_h, w) - 1, 0)
    shift = (n_h - 1) // 2 + n_edge
    n_x = w + 2 * n_edge
    min_fft = 2 * n_h - 1
    if n_fft is None:
        if n_x >= min_fft:
            cand = 2 ** np.arange(np.ceil(np.log2(min_fft)), np.ceil(np.log2(n_x)) + 1, dtype=int)
            cost = np.ceil(n_x / (cand - n_h + 1).astype(np.float64)) * cand * (np.log2(cand) + 1)
            cost += 4e-5 * cand * n_x
            n_fft = int(cand[np.argmin(cost)])
        else:
            n_fft = int(next_fast_len(min_fft))
    if n_fft < min_fft:
        raise ValueError("n_fft is too short for the filter")
    h_fft = rfft(h, n=n_fft)
    n_seg = n_fft - n_h + 1
    n_blocks = int(np.ceil(n_x / float(n_seg)))
    rows = range(x2.shape[0]) if picks is None else picks
    for r in rows:
        ext = _reflect_limited(x2[r], n_edge)
        acc = np.zeros_like(ext)
        for b in range(n_blocks):
            start = b * n_seg
            seg = ext[start : start + n_seg]
            prod = irfft(rfft(seg, n=n_fft) * h_fft, n=n_fft)
            lo = max(0, start - shift)
            hi = min(start - shift + n_fft, len(ext))
            p0 = max(0, shift - start)
            acc[lo:hi] += prod[p0 : p0 + hi - lo]
        out[r] = acc[:w]
    return out[0] if one_d else out


# ----------------------------------------------------------------------------- resample (processing/resample.py:58-60)
def _smart_pad2(x: np.ndarray, npads, pad: str = "reflect_limited") -> np.ndarray:
    """MNE ``_smart_pad`` with separate left / right pad counts (odd reflection about the end points, zero fill beyond)."""
    n0, n1 = int(npads[0]), int(npads[1])
    if n0 == 0 and n1 == 0:
        return x
    if pad != "reflect_limited":
        raise NotImplementedError(pad)
    lz = np.zeros(max(n0 - len(x) + 1, 0), dtype=x.dtype)
    rz = np.zeros(max(n1 - len(x) + 1, 0), dtype=x.dtype)
    return np.concatenate([lz, 2 * x[0] - x[n0:0:-1], x, 2 * x[-1] - x[-2 : -n1 - 2 : -1], rz])


def resample(x, up=1.0, down=1.0, *, axis=-1, window="auto", n_jobs=None, pad="auto", npad="auto", method="fft", verbose=None):
    """Restated ``mne.filter.resample`` for the one call pattern of the reference -- ``resample(x.astype(float64), up=ratio,
    down=1.0)`` with every other argument at its default (method="fft", npad="auto", pad="auto" -> "reflect_limited",
    window="auto" -> boxcar).  Written from MNE's documented behaviour (``_resamp_ratio_len`` / ``_resample_fft`` /
    ``_fft_resample``); MNE itself is not installed here, so this part is parity-unpinned like the FIR design above.

      ratio = up / down;  final_len = max(round(ratio * n), 1)
      npad "auto": min_add = min(n // 8, 100) * 2; pad to the next power of two >= n + min_add, split (floor, ceil)
      new_len = max(round(ratio * orig_len), 1); to_remove = [round(ratio * npad0), new_len - final_len - that]
      X = rfft(padded); if min(new_len, orig_len) is even: X[nyq] *= 2 (shorter) or 0.5 (longer)
      y = irfft(X * (new_len / orig_len), new_len)[to_remove0 : new_len - to_remove1]
    """
    from scipy.fft import irfft, rfft

    if method != "fft" or axis != -1 or not (isinstance(npad, str) and npad == "auto"):
        raise NotImplementedError("only the reference's call pattern is restated")
    x = np.asarray(x)
    if x.dtype != np.float64:
        raise TypeError("Arrays passed for resampling must have a dtype of np.float64")
    ratio = float(up) / down
    n = x.shape[-1]
    final_len = max(int(round(ratio * n)), 1)
    flat = x.reshape(-1, n)
    min_add = min(n // 8, 100) * 2
    npad_tot = 2 ** int(np.ceil(np.log2(n + min_add))) - n
    n0, extra = divmod(npad_tot, 2)
    npads = np.array([n0, n0 + extra], int)
    orig_len = n + int(npads.sum())
    new_len = max(int(round(ratio * orig_len)), 1)
    rem0 = int(round(ratio * npads[0]))
    rem1 = new_len - final_len - rem0
    shorter = new_len < orig_len
    use_len = new_len if shorter else orig_len
    # boxcar window, folded onto the rfft bins and scaled (MNE: W = ifftshift(boxcar) * new_len / orig_len)
    scale = float(new_len) / float(orig_len)
    out = np.zeros((flat.shape[0], new_len - rem0 - rem1), dtype=np.float64)
    for i, row in enumerate(flat):
        xp = _smart_pad2(row, npads)
        xf = rfft(xp)
        if use_len % 2 == 0:
            nyq = use_len // 2
            xf[nyq : nyq + 1] *= 2 if shorter else 0.5
        xf *= scale
        y = irfft(xf, new_len)
        out[i] = y[rem0 : y.shape[0] - rem1] if (rem0 > 0 or rem1 > 0) else y
    return out.reshape(x.shape[:-1] + (out.shape[-1],))
