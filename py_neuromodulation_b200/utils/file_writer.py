"""Feature table writers (reference: ``utils/file_writer.py``).

``MsgPackFileWriter`` keeps the reference's per-window interface (used by the streaming path);
``FeatureTableWriter`` is the batched equivalent used by the offline fast path: it receives the
whole ``(n_windows, n_columns)`` float64 matrix at once and produces the same DataFrame / CSV
without a Python dict per window.
"""

from __future__ import annotations

from pathlib import Path

import numpy as np

from .types import _PathLike


class MsgPackFileWriter:
    def __init__(self, name: str = "sub", out_dir: _PathLike = "") -> None:
        self.out_dir = (Path.cwd() if not out_dir else Path(out_dir)) / name
        self.out_dir.mkdir(parents=True, exist_ok=True)
        self.idx = 0
        self.name = name
        self.csv_path = self.out_dir / f"{name}_FEATURES.csv"
        self.data_l: list[dict] = []

    def insert_data(self, feature_dict: dict) -> None:
        for key, value in feature_dict.items():
            feature_dict[key] = float(value) if value is not None else 0
        self.data_l.append(feature_dict)

    def save(self) -> None:
        import msgpack

        if not self.data_l:
            return
        with open(self.out_dir / f"{self.name}-{self.idx}.msgpack", "wb") as f:
            msgpack.pack(self.data_l, f)
        self.idx += 1
        self.data_l = []

    def load_all(self):
        import msgpack
        import pandas as pd

        rows: list[dict] = []
        for i in range(self.idx):
            with open(self.out_dir / f"{self.name}-{i}.msgpack", "rb") as f:
                rows.extend(msgpack.unpack(f))
        if not rows:
            raise ValueError("No data to load")
        return pd.DataFrame(rows)

    def save_as_csv(self, save_all_combined: bool = False) -> None:
        import pandas as pd

        if save_all_combined:
            try:
                self.load_all().to_csv(self.csv_path, index=False)
            except ValueError:
                return
        elif self.data_l:
            pd.DataFrame([self.data_l[-1]]).to_csv(self.csv_path, index=False)

    def delete_ind_files(self) -> None:
        for f in self.out_dir.glob(f"{self.name}-*.msgpack"):
            f.unlink()


class FeatureTableWriter:
    """Batched writer: one float64 matrix + column names -> DataFrame / ``{name}_FEATURES.csv``."""

    def __init__(self, name: str = "sub", out_dir: _PathLike = "") -> None:
        self.out_dir = (Path.cwd() if not out_dir else Path(out_dir)) / name
        self.out_dir.mkdir(parents=True, exist_ok=True)
        self.name = name
        self.csv_path = self.out_dir / f"{name}_FEATURES.csv"

    def to_frame(self, columns: list[str], matrix: np.ndarray):
        import pandas as pd

        # (the matrix is a fresh C-contiguous float64 block owned by the caller: wrap it, do not copy it)
        return pd.DataFrame(np.asarray(matrix, dtype=np.float64), columns=list(columns), copy=False)

    def save_csv(self, frame) -> None:
        frame.to_csv(self.csv_path, index=False)
