"""Line-noise notch (reference: ``filter/notch_filter.py``): multi-band band-stop FIR applied with
reflect-limited padding, zero phase."""

from __future__ import annotations

import numpy as np

from ..utils.types import NMPreprocessor
from .fir_design import design_fir


class NotchFilter(NMPreprocessor):
    def __init__(
        self,
        sfreq: float,
        line_noise: float | None = None,
        freqs: np.ndarray | None = None,
        notch_widths: int | np.ndarray | None = 3,
        trans_bandwidth: float = 6.8,
    ) -> None:
        from .. import logger

        if line_noise is None and freqs is None:
            raise ValueError("Either line_noise or freqs must be defined if notch_filter isactivated.")
        if freqs is None:
            freqs = np.arange(line_noise, sfreq / 2, line_noise, dtype=int)
        freqs = np.asarray(freqs)
        if freqs.size > 0 and freqs[-1] >= sfreq / 2:
            freqs = freqs[:-1]
        self.sfreq = sfreq
        self._pipes: dict = {}
        if freqs.size == 0:
            self.filter_bank = None
            logger.warning(
                "WARNING: notch_filter is activated but data is not being filtered. This may be due to a low sampling "
                f"frequency or incorrect specifications. Make sure your settings are correct. Got: {sfreq = }, "
                f"{line_noise = }, {freqs = }."
            )
            return
        if notch_widths is None:
            widths = freqs / 200.0
        else:
            widths = np.atleast_1d(notch_widths)
            if np.any(widths < 0):
                raise ValueError("notch_widths must be >= 0")
            if len(widths) == 1:
                widths = widths[0] * np.ones_like(freqs)
            elif len(widths) != len(freqs):
                raise ValueError("notch_widths must be None, scalar, or the same length as freqs")
        half = trans_bandwidth / 2.0
        lows = [f - w / 2.0 - half for f, w in zip(freqs, widths)]
        highs = [f + w / 2.0 + half for f, w in zip(freqs, widths)]
        self.filter_bank = design_fir(sfreq, highs, lows, filter_length=int(sfreq - 1), l_trans_bandwidth=half,
                                      h_trans_bandwidth=half)

    def process(self, data: np.ndarray) -> np.ndarray:
        if self.filter_bank is None:
            return data
        from .._pipeline import filter_rows

        data = np.asarray(data)
        if data.dtype != np.float64:
            raise TypeError("Arrays passed for filtering must have a dtype of np.float64")
        return filter_rows(self, np.atleast_2d(data), mode="reflect").reshape(data.shape)
