"""`tensor_calibration(onnx_graph, args)` — the one entry the CLI and the weight transforms call
(dipoorlet/tensor_cali/tensor_cali_base.py:4-7)."""
from . import basic_algorithm as _algo


def tensor_calibration(onnx_graph, args):
    """-> (act_clip_val, weight_clip_val)
         act_clip_val    {blob name: [lo, hi]} np.float32 scalars, blob order of the engine
         weight_clip_val {initializer name: [lo[C], hi[C]]} per output channel
    The activation calibrator is looked up by `args.act_quant` in the registry, so plugins added
    with `@tensor_cali_dispatcher.register(...)` are reachable from the CLI unchanged."""
    ranges_w = _algo.find_clip_val_minmax_weight(onnx_graph, args)
    ranges_a = _algo.tensor_cali_dispatcher(args.act_quant, onnx_graph, args)
    return ranges_a, ranges_w
