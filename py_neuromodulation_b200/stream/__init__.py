from .data_processor import DataProcessor  # noqa: F401
from .generator import RawDataGenerator, window_grid  # noqa: F401
from .settings import NMSettings  # noqa: F401
from .stream import Stream  # noqa: F401
