from .weight_trans_base import weight_calibration  # noqa: F401
