"""Re-referencing (reference: ``processing/rereference.py``)."""

from __future__ import annotations

import numpy as np

from ..utils.types import NMPreprocessor


def build_reference_matrix(channels) -> np.ndarray | None:
    """(n_good_used, n_good_used) matrix with ``y = M @ x`` or ``None`` when <= 1 channel is used.

    Row i: identity plus ``-1/len(ref)`` on the reference channels -- all *good* channels of the same
    ``type`` except i for "average", the named channels for "a&b", nothing for "none"/bad channels.
    """
    import pandas as pd

    ch = channels[channels["used"] == 1].reset_index(drop=True)
    n = ch.shape[0]
    if n in (0, 1):
        return None
    names = ch["name"].tolist()
    types = ch["type"].tolist()
    status = ch["status"].tolist()
    m = np.zeros((n, n))
    refs = ch["rereference"].tolist()
    good_mask = np.array([st == "good" for st in status])
    type_arr = np.array([str(t) for t in types], dtype=object)
    same_type_good: dict = {}  # type -> indices of the good channels of that type (computed once per type, not per row)
    for i in range(n):
        m[i, i] = 1
        ref = refs[i]
        if pd.isnull(ref) or str(ref).lower() == "none" or status[i] != "good":
            continue
        if str(ref).lower() == "average":
            key = str(types[i])
            if key not in same_type_good:
                same_type_good[key] = np.flatnonzero(good_mask & (type_arr == key))
            pool = same_type_good[key]
            idx = pool[pool != i]
        else:
            idx = []
            for other in str(ref).split("&"):
                if other not in names:
                    raise ValueError(
                        "One or more of the reference channels are not part of the recording channels. First missing"
                        f" channel: {other}."
                    )
                if other == names[i]:
                    raise ValueError(f"You cannot rereference to the same channel. Channel: {other}.")
                idx.append(names.index(other))
        m[i, idx] = -1 / len(idx)
    good = [i for i in range(n) if status[i] == "good"]
    return m[np.ix_(good, good)]


class ReReferencer(NMPreprocessor):
    def __init__(self, sfreq: float, channels) -> None:
        self.ref_matrix = build_reference_matrix(channels)
        self._pipes: dict = {}

    def process(self, data: np.ndarray) -> np.ndarray:
        if self.ref_matrix is None:
            return data
        from .._pipeline import reref_rows

        return reref_rows(self, np.asarray(data, dtype=np.float64))
