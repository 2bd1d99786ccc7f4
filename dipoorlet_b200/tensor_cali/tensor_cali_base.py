from .basic_algorithm import find_clip_val_minmax_weight, tensor_cali_dispatcher


def tensor_calibration(onnx_graph, args):
    """Entry of the calibration registry (dipoorlet/tensor_cali/tensor_cali_base.py:4-7):
    -> (act_clip_val {name: [lo, hi]}, weight_clip_val {name: [lo[C], hi[C]]})."""
    weight_clip_val = find_clip_val_minmax_weight(onnx_graph, args)
    act_clip_val = tensor_cali_dispatcher(args.act_quant, onnx_graph, args)
    return act_clip_val, weight_clip_val
