"""Raw-data resampling (reference: ``processing/resample.py`` -> ``mne.filter.resample(x.astype(float64), up=ratio, down=1.0)``).

FFT resampling of one window is a LINEAR map of its samples: reflect-limited padding to a power-of-two length, forward real
transform, spectrum truncated (down-sampling) or zero-extended (up-sampling) with the Nyquist-bin correction, inverse transform
of the new length, padding cut off.  Like the FIR taps of the other preprocessors the map is *designed on the host* -- as the
dense ``(n_out, n_in)`` operator ``R`` -- and handed to the library as data (``nm_set_resampler``); on the GPU every window row
is ``y = R @ x`` (``csrc/nm_resample.cuh``, one float64 GEMM per chunk of windows, after the notch like in the reference's fixed
preprocessor order; the re-reference is hoisted in front of both -- it commutes with per-row linear maps).  Any ratio works, including the non-integer ones of float sampling rates.

MNE is not installed in this image, so the construction below is written from MNE's documented algorithm
(``_resample_fft`` / ``_fft_resample`` with the defaults npad="auto", pad="reflect_limited", window="boxcar") and is
parity-unpinned in the same sense as ``filter/fir_design.py``.

Quirk kept on purpose: everything downstream of the resampler is still built with the ORIGINAL sampling rate
(``stream/data_processor.py:55,77-81`` never update ``sfreq_raw``).
"""

from __future__ import annotations

from functools import lru_cache

import numpy as np

from ..utils.types import NMPreprocessor
from .settings_models import ResamplerSettings


def resample_geometry(n_in: int, ratio: float) -> dict:
    """Lengths MNE derives for an ``n_in``-sample row: pads (left, right), padded / resampled lengths, samples to cut, result length."""
    final_len = max(int(round(ratio * n_in)), 1)
    min_add = min(n_in // 8, 100) * 2
    npad_tot = 2 ** int(np.ceil(np.log2(n_in + min_add))) - n_in
    n0, extra = divmod(npad_tot, 2)
    pads = (n0, n0 + extra)
    orig_len = n_in + n0 + n0 + extra
    new_len = max(int(round(ratio * orig_len)), 1)
    rem0 = int(round(ratio * pads[0]))
    return {"final_len": final_len, "pads": pads, "orig_len": orig_len, "new_len": new_len, "remove": (rem0, new_len - final_len - rem0)}


@lru_cache(maxsize=8)
def resample_operator(n_in: int, ratio: float) -> np.ndarray:
    """Dense ``(n_out, n_in)`` float64 matrix ``R`` with ``mne.filter.resample(x, up=ratio) == R @ x`` for every row ``x``."""
    g = resample_geometry(n_in, ratio)
    n0, n1 = g["pads"]
    orig_len, new_len = g["orig_len"], g["new_len"]
    # padding operator (orig_len, n_in): odd reflection about both end samples, zeros beyond the reach of the reflection
    pad = np.zeros((orig_len, n_in))
    lz, rz = max(n0 - n_in + 1, 0), max(n1 - n_in + 1, 0)
    row = lz
    for j in range(min(n0, n_in - 1), 0, -1):  # 2 x[0] - x[j]
        pad[row, 0] += 2.0
        pad[row, j] -= 1.0
        row += 1
    pad[row : row + n_in, :] = np.eye(n_in)
    row += n_in
    for j in range(n_in - 2, n_in - 2 - min(n1, n_in - 1), -1):  # 2 x[-1] - x[j]
        pad[row, n_in - 1] += 2.0
        pad[row, j] -= 1.0
        row += 1
    assert row + rz == orig_len
    spec = np.fft.rfft(pad, axis=0)
    shorter = new_len < orig_len
    use_len = new_len if shorter else orig_len
    if use_len % 2 == 0:
        spec[use_len // 2] *= 2.0 if shorter else 0.5
    spec *= float(new_len) / float(orig_len)
    full = np.fft.irfft(spec, n=new_len, axis=0)
    r0, r1 = g["remove"]
    out = full[r0 : new_len - r1] if (r0 > 0 or r1 > 0) else full
    assert out.shape == (g["final_len"], n_in), (out.shape, g)
    return np.ascontiguousarray(out)


def integer_decimation(ratio: float) -> int:
    """D >= 2 when ``ratio == 1 / D`` exactly (2 kHz -> 1 kHz, 4 kHz -> 1 kHz, ...), else 0."""
    if not 0.0 < ratio < 1.0:
        return 0
    d = int(round(1.0 / ratio))
    return d if d >= 2 and 1.0 / d == ratio else 0


class Resampler(NMPreprocessor):
    """Same constructor / ``process`` contract as the reference class; the arithmetic runs on the GPU."""

    def __init__(self, sfreq: float, resample_freq_hz: float, **kwargs) -> None:
        self.settings = ResamplerSettings(resample_freq_hz=resample_freq_hz)
        ratio = float(resample_freq_hz / sfreq)
        self.up = 0.0 if ratio == 1.0 else ratio
        self._pipes: dict = {}

    def process(self, data: np.ndarray) -> np.ndarray:
        if not self.up:
            return data
        from .._pipeline import Pipeline, ScanSpec

        data = np.asarray(data, dtype=np.float64)
        key = data.shape
        if key not in self._pipes:
            names = [f"c{i}" for i in range(data.shape[0])]
            op = resample_operator(int(data.shape[1]), float(self.up))
            pipe = Pipeline(data.shape[0], data.shape[0], op.shape[0], [f"{n}_raw" for n in names])
            pipe.set_resampler(op, integer_decimation(float(self.up)))
            ScanSpec(names, raw=True).attach(pipe)
            pipe.finalize()
            self._pipes[key] = pipe
        return self._pipes[key].preprocess_window(data)
