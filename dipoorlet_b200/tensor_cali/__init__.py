from .basic_algorithm import find_clip_val_minmax_weight, tensor_cali_dispatcher  # noqa: F401
from .tensor_cali_base import tensor_calibration  # noqa: F401
