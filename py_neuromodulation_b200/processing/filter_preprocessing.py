"""Window-wise FIR filter chain (reference: ``processing/filter_preprocessing.py``).

One single-filter :class:`MNEFilter` per enabled stage, ``filter_length = sfreq - 1``: the band stages first (in
``FilterSettings`` field order), then the low-pass, then the high-pass (reference lines 44-77); ``process`` applies them
one after the other with zero-padded 'same' FFT convolution (lines 79-94).  Inside a :class:`DataProcessor` the stages run
on the GPU in front of the notch (``nm_set_prefilters``); a stand-alone instance filters through ``nm_fir_apply``.

Reference behaviour reproduced on purpose: the ``bandstop_filter`` range is handed to ``create_filter`` as
``(l_freq, h_freq) = (low, high)``, which designs a band-PASS.
"""

from __future__ import annotations

from typing import TYPE_CHECKING

import numpy as np

from ..utils.types import NMPreprocessor
from .settings_models import FilterSettings  # noqa: F401  (re-exported like the reference module)

if TYPE_CHECKING:
    from ..stream.settings import NMSettings


class PreprocessingFilter(NMPreprocessor):
    def __init__(self, settings: "NMSettings", sfreq: float) -> None:
        from ..filter.mne_filter import MNEFilter

        fs = settings.preprocessing_filter
        enabled = fs.get_enabled()
        ranges = [fs.get_filter_tuple(name) for name in enabled if name not in ("lowpass_filter", "highpass_filter")]
        if "lowpass_filter" in enabled:
            ranges.append((None, fs.lowpass_filter_cutoff_hz))
        if "highpass_filter" in enabled:
            ranges.append((fs.highpass_filter_cutoff_hz, None))
        self.filters = [MNEFilter(f_ranges=[r], sfreq=sfreq, filter_length=sfreq - 1, verbose=False) for r in ranges]

    def stage_taps(self) -> list[np.ndarray] | None:
        """One tap vector per stage for the fused pipeline (lengths may differ: MNEFilter falls back to an automatic
        design when the requested length is too short for the transition band); None if no stage is enabled."""
        if not self.filters:
            return None
        return [np.ascontiguousarray(f.filter_bank[0]) for f in self.filters]

    def process(self, data: np.ndarray) -> np.ndarray:
        for f in self.filters:
            data = f.filter_data(data if data.ndim == 2 else data[:, 0, :])
        return data if data.ndim == 2 else data[:, 0, :]
