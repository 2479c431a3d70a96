from .fir_design import design_fir  # noqa: F401
from .mne_filter import MNEFilter  # noqa: F401
from .notch_filter import NotchFilter  # noqa: F401
from .kalman_settings import KalmanSettings  # noqa: F401
