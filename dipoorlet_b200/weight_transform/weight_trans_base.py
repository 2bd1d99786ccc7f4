from .. import dist_helper
from ..graph import load_graph
from ..tensor_cali import find_clip_val_minmax_weight, tensor_calibration
from ..utils import ONNXGraph, logger, update_model_path
from .adaround import adaround
from .bias_correction import bias_correction
from .brecq import brecq
from .weight_equalization import weight_equalization


def weight_calibration(onnx_graph, act_clip_val, weight_clip_val, args):
    """Ordering of the weight transforms (dipoorlet/weight_transform/weight_trans_base.py:15-68):
    bias correction (rank 0, all images) -> weight equalisation (+ re-calibration) -> adaround ->
    brecq / qdrop. After it the model, args and clip values are identical on every rank.
      -> (graph_after_wt, graph_ori, act_clip_val, weight_clip_val)
    Not on the B200 hot path (no BASELINE.json config uses them): --update_bn,
    --sparse raise NotImplementedError instead of being silently ignored."""
    for flag in ("update_bn", "sparse"):
        if getattr(args, flag, False):
            raise NotImplementedError(f"--{flag} is outside the B200 hot path (see DESIGN.md, out of scope)")
    graph_after_wt = ONNXGraph()
    graph_after_wt.copy_from(onnx_graph)
    if args.bc:
        if dist_helper.get_rank() == 0:
            bias_correction(graph_after_wt, act_clip_val, weight_clip_val, args)
        dist_helper.barrier()
        update_model_path('update_bias_model', args)
        # weight_trans_base.py:27 reloads without `deploy`, which makes a later adaround /
        # brecq crash in the reference (no `_q` value types); keeping `deploy` is a superset.
        graph_after_wt = load_graph(args.model, args.output_dir, onnx_graph.deploy,
                                    onnx_graph.model_type, do_simplify=False)
        weight_clip_val = find_clip_val_minmax_weight(graph_after_wt, args)   # bias ranges changed
    if getattr(args, "we", False):
        if dist_helper.get_rank() == 0:
            weight_equalization(graph_after_wt, args)
        dist_helper.barrier()
        update_model_path('weight_equal_model', args)
        graph_after_wt = load_graph(args.model, args.output_dir, onnx_graph.deploy,
                                    onnx_graph.model_type, do_simplify=False)
        act_clip_val, weight_clip_val = tensor_calibration(graph_after_wt, args)   # weight_trans_base.py:36
    if args.adaround:
        args.acti_quant = False
        graph_after_wt = adaround(onnx_graph, graph_after_wt, act_clip_val, weight_clip_val, args)
    if args.brecq:
        args.acti_quant = bool(args.drop is True)
        graph_after_wt = brecq(onnx_graph, graph_after_wt, act_clip_val, weight_clip_val, args)
    return graph_after_wt, onnx_graph, act_clip_val, weight_clip_val
