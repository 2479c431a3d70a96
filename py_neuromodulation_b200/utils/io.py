"""File helpers used by the stream (reference: ``utils/io.py``).  BIDS / MNE readers are out of scope."""

from __future__ import annotations

import json
from pathlib import Path

import numpy as np

from .types import _PathLike


def load_channels(channels):
    """DataFrame passes through, a path is read as CSV (reference ``utils/io.py:16-31``)."""
    import pandas as pd

    if isinstance(channels, pd.DataFrame):
        return channels
    if not Path(channels).is_file():
        raise ValueError(f"PATH_CHANNELS is not a valid file. Got: {channels}")
    return pd.read_csv(channels)


def write_csv(df, path_out) -> None:
    from pyarrow import Table, csv

    csv.write_csv(Table.from_pandas(df), path_out)


def save_channels(nmchannels, out_dir: _PathLike = "", prefix: str = "") -> None:
    out_dir = Path.cwd() if not out_dir else Path(out_dir)
    write_csv(nmchannels, out_dir / prefix / ("channels.csv" if not prefix else prefix + "_channels.csv"))


def save_features(df_features, out_dir: _PathLike = "", prefix: str = "") -> None:
    out_dir = Path.cwd() if not out_dir else Path(out_dir)
    write_csv(df_features, out_dir / (f"{prefix}_FEATURES.csv" if prefix else "_FEATURES.csv"))


def default_json_convert(obj):
    import pandas as pd

    if isinstance(obj, np.ndarray):
        return obj.tolist()
    if isinstance(obj, pd.DataFrame):
        return obj.to_numpy().tolist()
    if isinstance(obj, np.integer):
        return int(obj)
    if isinstance(obj, np.floating):
        return float(obj)
    raise TypeError("Not serializable")


def save_general_dict(dict_: dict, out_dir: _PathLike = "", prefix: str = "", str_add: str = "") -> None:
    out_dir = Path.cwd() if not out_dir else Path(out_dir)
    with open(out_dir / prefix / f"{prefix}{str_add}", "w") as f:
        json.dump(dict_, f, default=default_json_convert, indent=4, separators=(",", ": "))


def save_sidecar(sidecar: dict, out_dir: _PathLike = "", prefix: str = "") -> None:
    save_general_dict(sidecar, out_dir, prefix, "_SIDECAR.json")
