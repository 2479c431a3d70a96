"""Band-pass FIR bank (reference: ``filter/mne_filter.py``).

Design happens on the host (:mod:`.fir_design`); ``filter_data`` runs the channels-batched
FFT-convolution kernel (``csrc/nm_fir.cuh``, 'same' mode) and returns the filtered rows.
"""

from __future__ import annotations

from collections.abc import Sequence

import numpy as np

from .fir_design import design_fir


class MNEFilter:
    def __init__(
        self,
        f_ranges: Sequence[tuple[float | None, float | None]],
        sfreq: float,
        filter_length: str | float = "999ms",
        l_trans_bandwidth: float | str = 4,
        h_trans_bandwidth: float | str = 4,
        verbose: bool | int | str | None = None,
    ) -> None:
        if isinstance(filter_length, float):
            filter_length = int(filter_length)
        bank = []
        for lo, hi in f_ranges:
            try:
                taps = design_fir(sfreq, lo, hi, filter_length=filter_length, l_trans_bandwidth=l_trans_bandwidth,
                                  h_trans_bandwidth=h_trans_bandwidth)
            except ValueError:
                # same fallback as the reference: automatic length and transition bands
                taps = design_fir(sfreq, lo, hi)
            bank.append(taps)
        self.num_filters = len(bank)
        self.filter_bank = np.vstack(bank)
        self.sfreq = sfreq
        self._pipes: dict = {}

    def filter_data(self, data: np.ndarray) -> np.ndarray:
        """(n_samples,) or (n_channels, n_samples) -> (n_channels, n_filters, n_samples)."""
        from .._pipeline import filter_rows

        data = np.asarray(data, dtype=np.float64)
        if data.ndim > 2:
            raise ValueError(f"Data must have one or two dimensions. Got: {data.ndim} dimensions.")
        if data.ndim == 1:
            data = data[None, :]
        return filter_rows(self, data, mode="same")
