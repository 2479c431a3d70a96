"""Package logger (reference: ``utils/logging.py``): console at INFO by default, optional file log."""

from __future__ import annotations

import logging
from pathlib import Path


class NMLogger(logging.Logger):
    def __init__(self, name: str, level: int = logging.INFO) -> None:
        super().__init__(name, level)
        handler = logging.StreamHandler()
        handler.setFormatter(logging.Formatter("%(name)s:\t%(message)s"))
        self.addHandler(handler)
        self._console = handler

    def set_level(self, level) -> None:
        self.setLevel(level)
        for h in self.handlers:
            h.setLevel(level)

    def log_to_file(self, path, mode: str = "w") -> None:
        path = Path(path)
        path.mkdir(parents=True, exist_ok=True)
        fh = logging.FileHandler(path / "logging_file.log", mode=mode)
        fh.setFormatter(logging.Formatter("%(asctime)s:%(levelname)s:%(name)s:%(filename)s:%(funcName)s:%(lineno)d:\t%(message)s"))
        self.addHandler(fh)
