"""Quantisation parameters and the Q/DQ rewrite of the graph — semantics of
dipoorlet/quantize.py:20-239 on the plain-container IR.

`get_qnode_by_param` turns a clip range into scale / zero point / integer limits;
`quant_graph` decides WHICH tensors are fake-quantised (weights of the platform's
quant_nodes, their activation inputs minus the merged-ReLU and TensorRT add-merge
exceptions) and inserts QuantizeLinear + DequantizeLinear pairs, which the engine then
executes as one fused K5 launch each.
"""
import copy

import numpy as np

from . import onnx_lite as ol
from .graph import ONNXGraph
from .platform_settings import LAYER_HAS_WEIGHT, platform_setting_table
from .utils import logger

QTENSORSUFFIX = '_q'
DQTENSORSUFFIX = '_dq'
QNODESUFFIX = '_fake_quant'
DQNODESUFFIX = '_fake_dequant'
QUANT_NODE_NAME_LIST = ['QuantizeLinear', 'DequantizeLinear']
MERGE_RELU = ['Conv', 'Gemm', 'Eltwise', 'Add']
RELU_TYPE = ['Relu', 'PRelu', 'Mul']
WEIGHT_TRANSPOSE_SUFFIX = '_transpose'
CLIP_SUFFIX = '_clip'


class QNodes:
    """The Q/DQ pair of one tensor: what the reference carries around as a small
    GraphProto (`q_nodes.node`, `.initializer`, `.output`)."""

    def __init__(self, nodes, initializers, outputs):
        self.node = nodes
        self.initializer = initializers      # [(name, ndarray)]
        self.output = outputs                # [ValueInfo]


def quant_graph(onnx_graph, clip_val, args):
    """quantize.py:20-37 -> (graph_q, quant_node_list)."""
    graph_q = ONNXGraph()
    graph_q.copy_from(onnx_graph)
    platform = platform_setting_table[args.deploy]
    quant_node_list = [n for n in graph_q.graph.node
                       if n.name not in args.skip_layers and n.op_type in platform['quant_nodes']]
    done = []
    for node in quant_node_list:
        insert_fake_quant_node(graph_q, node, done, clip_val, args)
    if platform['quantize_network_output']:
        insert_fake_quant_node_output(graph_q, clip_val, args)
    graph_q.update_model()
    graph_q.topologize_graph()
    return graph_q, quant_node_list


def insert_fake_quant_node(graph, node, act_quantized, data_range_list, args):
    """quantize.py:40-95. Inputs of `node` are rewired to `<tensor>_dq`; each tensor gets
    its Q/DQ pair once."""
    param = platform_setting_table[args.deploy]
    found_weight = False
    trt_merged_one_add_branch = False
    for idx, tensor in enumerate(list(node.input)):
        need_transpose = False
        shape = graph.tensor_name_shape_map.get(tensor)
        # a ReLU fed by Conv/Gemm/Add is executed fused with it: its input stays fp
        if node.op_type in RELU_TYPE:
            prev = graph.get_tensor_producer(node.input[0])
            if isinstance(prev, str):
                continue
            if len(node.input) == 1 and prev.op_type in MERGE_RELU:
                continue
        q_nodes = None
        if tensor in graph.initializer and node.op_type in LAYER_HAS_WEIGHT:
            if not found_weight:
                found_weight = True
                need_transpose = node.op_type == 'ConvTranspose'
                q_nodes, _, _ = get_qnode_by_param(param['qw_params'], tensor, shape,
                                                   data_range_list[tensor], need_transpose)
            elif 'qb_params' in param:
                q_nodes, _, _ = get_qnode_by_param(param['qb_params'], tensor, shape,
                                                   data_range_list[tensor], need_transpose)
        if tensor in graph.network_inputs or tensor not in graph.input:
            # TensorRT fuses the first Conv-produced branch of an Add into that Conv
            if args.deploy == 'trt' and node.op_type == 'Add' and not trt_merged_one_add_branch:
                prev = graph.get_tensor_producer(tensor)
                if not isinstance(prev, str) and prev.op_type == 'Conv':
                    trt_merged_one_add_branch = True
                    continue
            q_nodes, _, _ = get_qnode_by_param(param['qi_params'], tensor, shape,
                                               data_range_list[tensor])
        if q_nodes is not None:
            node.input[idx] = q_nodes.output[0].name
            if tensor in act_quantized:
                continue
            graph.insert_qnodes_purely(q_nodes=q_nodes, node=node)
            act_quantized.append(tensor)
    graph.topologize_graph()


def insert_fake_quant_node_output(graph, clip_val, args):
    """quantize.py:98-108 (platforms with quantize_network_output; not trt)."""
    param = platform_setting_table[args.deploy]
    for out_tensor in copy.deepcopy(graph.network_outputs):
        q_nodes, _, _ = get_qnode_by_param(param['qi_params'], out_tensor,
                                           graph.get_tensor_shape(out_tensor), clip_val[out_tensor])
        graph.insert_qnodes_purely(q_nodes=q_nodes,
                                   idx=graph.index(graph.get_tensor_producer(out_tensor)) + 1)
        graph.del_network_output(out_tensor)
        graph.add_network_output(q_nodes.output[0])
    graph.topologize_graph()


def get_qnode_by_param(param, in_tensor_name, tensor_shape, range, need_transpose=False):
    """quantize.py:111-194 -> (q_nodes, q_min, q_max). Like the reference this collapses
    `range` IN PLACE to scalars when the parameter set is not per-channel."""
    bit_width = param['bit_width']
    zero_point = [0]
    per_channel = bool(param.get('per_channel', False))
    if param['type'] != "Linear":
        raise NotImplementedError(f"quantisation type {param['type']}")
    symmetric = param['symmetric']
    if not per_channel:
        range[0] = np.min(range[0])
        range[1] = np.max(range[1])
        if param.get('dynamic_sym') and np.abs(range[0] - 0.0) < 1e-6:
            symmetric = False
    if symmetric:
        channels = len(range[0]) if isinstance(range[0], np.ndarray) else 1
        q_min = [-2 ** (bit_width - 1) + 1] * channels     # -127: "-128..127 is identical to -127..127"
        q_max = [2 ** (bit_width - 1) - 1] * channels
        data_max = np.max(np.abs(range), axis=0)
        scale = np.array(data_max) / q_max
        if np.any(scale == 0):
            scale = np.where(scale == 0, 1., scale)          # all-zero channel
        scale = scale.tolist()
    elif not isinstance(range[0], np.ndarray):
        data_min, data_max = min(0, range[0]), max(0, range[1])
        scale = (data_max - data_min) / (2 ** bit_width - 1)
        if scale == 0.0:
            scale += 1.
        zero_point = np.round(-data_min / scale)
        q_min = [int(-zero_point)]
        q_max = [int(2 ** bit_width - 1 - zero_point)]
        scale = [float(scale)]
    else:
        data_min, data_max = range[0], range[1]
        data_min[data_min > 0.] = 0.
        data_max[data_max < 0.] = 0.
        scale = (data_max - data_min) / (2 ** bit_width - 1)
        if np.any(scale == 0):
            logger.warning("Find {} channels all zero in {}, set scale to 1.".format(
                len(np.where(scale == 0)[0]), in_tensor_name))
            scale = np.where(scale == 0, 1., scale)
        zero_point = (-data_min / scale).round()
        q_max = (2 ** bit_width - 1 - zero_point).astype(np.int32).tolist()
        q_min = (-zero_point).astype(np.int32).tolist()
        scale = scale.tolist()
    if param.get('log_scale'):
        scale = 2 ** np.round(np.log2(scale))
    scale = np.array(scale, dtype=np.float32)
    zero_point = np.full(scale.shape, zero_point, dtype=np.int8)
    q_nodes = make_quant_dequant(in_tensor_name, tensor_shape, scale, zero_point, need_transpose,
                                 per_channel, symmetric)
    return q_nodes, q_min, q_max


def make_quant_dequant(tensor_name, tensor_shape, scale_val, zero_point_val, need_transpose=False,
                       per_channel=False, symmetric=True):
    """quantize.py:197-239: QuantizeLinear -> `<t>_q` -> DequantizeLinear -> `<t>_dq`;
    int8 zero point when symmetric else uint8; axis 0 (1 for ConvTranspose weights)."""
    scale = np.asarray(scale_val, dtype=np.float32).reshape(-1)
    zp = np.asarray(zero_point_val).reshape(-1).astype(np.int8 if symmetric else np.uint8)
    if scale.size == 1:
        scale, zp = scale.reshape(()), zp.reshape(())
    attrs = {"axis": 1 if need_transpose else 0} if per_channel else {}
    s_name, z_name = tensor_name + '_scale', tensor_name + '_zero_point'
    q = ol.Node("QuantizeLinear", [tensor_name, s_name, z_name], [tensor_name + QTENSORSUFFIX],
                tensor_name + "_QuantizeLinear", dict(attrs))
    dq = ol.Node("DequantizeLinear", [tensor_name + QTENSORSUFFIX, s_name, z_name],
                 [tensor_name + DQTENSORSUFFIX], tensor_name + "_DequantizeLinear", dict(attrs))
    out = ol.ValueInfo(tensor_name + DQTENSORSUFFIX, ol.FLOAT, tensor_shape)
    return QNodes([q, dq], [(s_name, scale), (z_name, zp)], [out])
