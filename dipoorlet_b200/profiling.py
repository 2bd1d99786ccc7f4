"""Quantisation-accuracy profiling: cosine similarity between the fp graph and the Q/DQ
graph per quantised layer and per network output (dipoorlet/profiling.py:34-99).

The reference rebuilds two ONNXRuntime sessions PER IMAGE (forward_net.py:467-485) and
reduces on the host; here both graphs run batched on the GPU, every cosine is three sums
from one K7b launch per tensor per batch, and ranks are combined with one all-reduce."""
import heapq
import os
import math

import numpy as np
import torch

from . import dist_helper
from . import kernels as K
from .engine import Engine
from .forward_net import _device_of, _per_image_shape, as_input_source
from .platform_settings import platform_setting_table
from .quantize import DQTENSORSUFFIX, quant_graph
from .utils import logger


def get_output_single_map(graph):
    """Outputs with <= 10 elements per image are compared stacked over all images
    (profiling.py:219-224)."""
    return {t: np.prod(graph.get_tensor_shape(t)[1:]) <= 10 for t in graph.network_outputs}


def _cosines(a, b):
    """Per-image cosine of two [B, ...] tensors -> float64[B] on the device."""
    sums = torch.zeros((a.shape[0], 3), dtype=torch.float64, device=a.device)
    K.cosine3(a.contiguous(), b.contiguous(), sums)
    cos = sums[:, 0] / torch.sqrt(sums[:, 1]) / torch.sqrt(sums[:, 2])
    return torch.where(sums[:, 0] == 0, torch.zeros_like(cos), cos)   # utils.py:275-276


def quantize_profiling_multipass(graph_after_wt, graph_ori, act_clip_val, weight_clip_val, args):
    """-> (layer_cosine_dict {tensor: mean cos}, model_cosine_dict {output: [mean, min]},
    quant_node_list). Identical on every rank."""
    clip_val = dict(act_clip_val)
    clip_val.update(weight_clip_val)
    graph_q, quant_node_list = quant_graph(graph_after_wt, clip_val, args)
    rank, world = dist_helper.get_rank(), dist_helper.get_world_size()
    if rank == 0:
        graph_q.save_onnx_model(name='quant_model')
    dev = _device_of(args)
    fp_eng = Engine(graph_ori, dev, _unit_test_cpu=dev.type != "cuda")
    q_eng = Engine(graph_q, dev, _unit_test_cpu=dev.type != "cuda")
    per = math.ceil(args.data_num / world)                      # profiling.py:48-51 (ceil rule)
    st, ed = min(rank * per, args.data_num), min(rank * per + per, args.data_num)
    source = as_input_source(args.input_dir)
    in_shapes = {n: _per_image_shape(graph_ori, n) for n in graph_ori.network_inputs}
    layer_names = [t for node in quant_node_list for t in node.output]
    outputs = list(graph_after_wt.network_outputs)
    single = get_output_single_map(graph_after_wt)
    q_out_name = {t: (t + DQTENSORSUFFIX if t + DQTENSORSUFFIX in graph_q.output_map else t) for t in outputs}
    want_fp = list(dict.fromkeys(layer_names + outputs))
    want_q = list(dict.fromkeys(layer_names + [q_out_name[t] for t in outputs]))
    layer_sum = torch.zeros(len(layer_names), dtype=torch.float64, device=dev)
    out_sum = torch.zeros(len(outputs), dtype=torch.float64, device=dev)
    out_min = torch.full((len(outputs),), float("inf"), dtype=torch.float64, device=dev)
    single_sums = torch.zeros((len(outputs), 3), dtype=torch.float64, device=dev)
    bs = int(getattr(args, "calib_bs", 0) or 32)
    for b0 in range(st, ed, bs):
        b1 = min(b0 + bs, ed)
        feeds = {n: source.fetch(n, b0, b1, shp).to(dev, non_blocking=True) for n, shp in in_shapes.items()}
        fp = fp_eng.run(feeds, want=want_fp)
        q = q_eng.run(feeds, want=want_q)
        for i, t in enumerate(layer_names):
            layer_sum[i] += _cosines(fp[t], q[t]).sum()
        for i, t in enumerate(outputs):
            a, b = fp[t], q[q_out_name[t]]
            if single[t]:   # one cosine over the stacked outputs of all images
                s = torch.zeros((1, 3), dtype=torch.float64, device=dev)
                K.cosine3(b.reshape(1, -1).contiguous(), a.reshape(1, -1).contiguous(), s)
                single_sums[i] += s[0]
            else:
                c = _cosines(a, b)
                out_sum[i] += c.sum()
                out_min[i] = torch.minimum(out_min[i], c.min())
            if getattr(args, "savefp", False) and dist_helper.get_rank() == 0:
                # --savefp (profiling.py:82-86): rank 0's fp network outputs, one raw float32 file per image
                save_path = os.path.join(args.output_dir, 'output', t)
                os.makedirs(save_path, exist_ok=True)
                host = a.detach().cpu().numpy().astype(np.float32)
                for j in range(host.shape[0]):
                    host[j].tofile(os.path.join(save_path, 'onnx-output-{}.bin'.format(b0 + j)))
        del fp, q
    n_local = torch.tensor([float(max(ed - st, 0))], dtype=torch.float64, device=dev)   # an empty ceil-rule shard counts 0
    if dist_helper.is_dist():
        for t in (layer_sum, out_sum, single_sums, n_local):
            dist_helper.allreduce_sum(t)
        torch.distributed.all_reduce(out_min, op=torch.distributed.ReduceOp.MIN)
    n_all = float(n_local.item())
    layer_mean = (layer_sum / n_all).cpu().numpy()
    layer_cosine_dict = {t: layer_mean[i] for i, t in enumerate(layer_names)}
    model_cosine_dict = {}
    out_mean, out_min_h, ss = (out_sum / n_all).cpu().numpy(), out_min.cpu().numpy(), single_sums.cpu().numpy()
    for i, t in enumerate(outputs):
        if single[t]:
            c = 0. if ss[i, 0] == 0 else ss[i, 0] / np.sqrt(ss[i, 1]) / np.sqrt(ss[i, 2])
            model_cosine_dict[t] = [c, c]
        else:
            model_cosine_dict[t] = [out_mean[i], out_min_h[i]]
    return layer_cosine_dict, model_cosine_dict, quant_node_list


def show_model_ranges(graph, act_clip_val, weight_clip_val, args):
    logger.info("Model ranges:")
    ranges_all = dict(act_clip_val)
    ranges_all.update(weight_clip_val)
    per_channel = platform_setting_table[args.deploy]['qw_params'].get('per_channel', False)
    for name, rng in ranges_all.items():
        shape = graph.tensor_name_shape_map.get(name)
        if isinstance(rng[0], np.ndarray):
            logger.info("{:<30} Shape: {:<20} Range: {}[{:<10f} {:<10f}]".format(
                name, str(shape), "per channel " if per_channel else "", rng[0].min(), rng[1].max()))
        else:
            logger.info("{:<30} Shape: {:<20} Range: [{:<10f} {:<10f}]".format(name, str(shape), rng[0], rng[1]))


def weight_need_perchannel(graph, args):
    if platform_setting_table[args.deploy]['qw_params'].get('per_channel'):
        return
    logger.info("Layer degradate by per layer: ")
    heap = []
    for node in graph.graph.node:
        if node.op_type == 'Conv':
            w = graph.get_initializer(node.input[1])
            flat = w.reshape((w.shape[0], -1))
            ratio = (flat.max(-1) - flat.min(-1)).mean() / (w.max() - w.min())
            heapq.heappush(heap, (ratio, node.name))
    for ratio, name in heapq.nsmallest(len(heap), heap):
        logger.info("{:40} ratio : {:<.5f}".format(name, ratio))


def show_model_profiling_res(graph_after_wt, layer_cosine_dict, model_cosine_dict, quant_node_list, args):
    single = get_output_single_map(graph_after_wt)
    if not args.skip_prof_layer:
        heap = []
        for node in quant_node_list:
            logger.info(node.name)
            for t in node.output:
                logger.info("Layer  cos: {:.5f}".format(layer_cosine_dict[t]))
                heapq.heappush(heap, (layer_cosine_dict[t], node.name + '-' + t))
        logger.info("The smallest cos value of 10 layers: ")
        for cos, name in heapq.nsmallest(10, heap):
            logger.info("{:40} cos : {:<.5f}".format(name, cos))
    logger.info("Quant model output cos: ")
    for name in graph_after_wt.network_outputs:
        if not single[name]:
            logger.info("{:40} avgcos : {:<.5f}    mincos : {:<.5f}".format(
                name, model_cosine_dict[name][0], model_cosine_dict[name][1]))
        else:
            logger.info("{:40} tolcos : {:<.5f}".format(name, model_cosine_dict[name][0]))
