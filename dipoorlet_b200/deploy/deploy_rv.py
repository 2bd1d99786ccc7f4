"""Rockchip deploy writer, `-D rv` (dipoorlet/deploy/deploy_rv.py:11-178): two parameter files per format,
rv_quantized_param.{yaml,json} (RV1126: per-tensor asymmetric u8 records keyed `@<node>:out<i>` /
`:weight` / `:bias`) and rk_quantized_param.{yaml,json} (RK3568: {min, max} records keyed by tensor name,
`<node>_W`, `<node>_b`). Host-side formatting only. Pinned byte for byte against the reference's files
(tests/golden/*/deploy_vendors.json) — including the YAML anchors, which come from the reference's way
of merging a Relu into its producer: both keys hold the SAME record object.
Rules restated from the reference: a Concat's inputs take its output range; every range is widened to
contain 0; a node whose only consumer is a Sigmoid gets no record; biases are i32 with scale =
activation scale x weight scale (RV1126) or a symmetric range (RK3568)."""
import json
import os

import numpy as np
import yaml

from ..platform_settings import LAYER_HAS_WEIGHT
from .deploy_default import deploy_dispatcher


def _affine_u8(clip):
    """8-bit step and zero point of a range widened to contain 0 (deploy_rv.py:11-21)."""
    lo, hi = min(0, np.min(clip[0])), max(0, np.max(clip[1]))
    step = (hi - lo) / 255.
    if step == 0.0:
        step = 1.0 / 255.
    return {'scale': [float(step)], 'zero_point': [int(round(-lo / step))]}


def _feeds_only_sigmoid(graph, node):
    nxt = graph.get_tensor_consumer(node.output[0])
    return len(nxt) == 1 and not isinstance(nxt[0], str) and nxt[0].op_type == 'Sigmoid'


def _spread_concat_ranges(graph, clip_val):
    for node in graph.graph.node:
        if node.op_type == 'Concat':
            for t in node.input:
                clip_val[t][0] = clip_val[node.output[0]][0]
                clip_val[t][1] = clip_val[node.output[0]][1]


def _write(res, args, stem):
    with open(os.path.join(args.output_dir, stem + '.yaml'), 'w') as f:
        f.write(yaml.dump(res))
    with open(os.path.join(args.output_dir, stem + '.json'), 'w') as f:
        json.dump(res, f, indent=4)


def _rv1126(graph, clip_val, args):
    table = {}

    def u8_record(clip):
        rec = {'dtype': 'asymmetric_affine', 'method': 'layer',
               'max_value': [max(0., float(np.max(clip[1])))], 'min_value': [min(0., float(np.min(clip[0])))],
               'qtype': 'u8'}
        rec.update(_affine_u8(clip))
        return rec

    _spread_concat_ranges(graph, clip_val)
    for name in graph.network_inputs:
        table[f'@{name}:out0'] = u8_record(clip_val[name])
    for node in graph.graph.node:
        if _feeds_only_sigmoid(graph, node):
            continue
        if node.op_type in LAYER_HAS_WEIGHT:
            for idx, t in enumerate(node.input[1:]):
                if idx == 0:
                    table[f'@{node.name}:weight'] = u8_record(clip_val[t])
                elif idx == 1:
                    act_scale = _affine_u8(clip_val[node.input[0]])['scale'][0]
                    w_scale = _affine_u8(clip_val[node.input[1]])['scale'][0]
                    table[f'@{node.name}:bias'] = {'dtype': 'asymmetric_affine', 'method': 'layer', 'max_value': [],
                                                   'min_value': [], 'zero_point': [0],
                                                   'scale': [w_scale * act_scale], 'qtype': 'i32'}
                else:
                    print("We meet unsupported node{}, skip.".format(node.name))
        last = None
        for idx, t in enumerate(node.output):
            last = f'@{node.name}:out{idx}'
            table[last] = u8_record(clip_val[t])
        if node.op_type == 'Relu' and last is not None:
            producer = graph.get_tensor_producer(node.input[0])
            if not isinstance(producer, str):
                # the producer's output records BECOME the Relu's record (one shared object); the reference
                # matches keys by substring, so "Conv_1" also captures "Conv_10:out0"
                for key in table:
                    if producer.name in key and 'out' in key:
                        table[key] = table[last]
    _write({'customized_quantize_layers': {}, 'quantize_parameters': table}, args, 'rv_quantized_param')


def _rk3568(graph, clip_val, args):
    table = {}

    def span(clip):
        return {'max': [max(0., float(np.max(clip[1])))], 'min': [min(0., float(np.min(clip[0])))]}

    _spread_concat_ranges(graph, clip_val)
    for name in graph.network_inputs:
        table[f'{name}'] = span(clip_val[name])
    for node in graph.graph.node:
        if _feeds_only_sigmoid(graph, node):
            continue
        if node.op_type in LAYER_HAS_WEIGHT:
            for idx, t in enumerate(node.input[1:]):
                if idx == 0:
                    table[f'{node.name}_W'] = span(clip_val[t])
                elif idx == 1:
                    bound = max(abs(np.max(clip_val[node.input[2]])), abs(np.min(clip_val[node.input[2]])))
                    table[f'{node.name}_b'] = {'max': [float(bound)], 'min': [float(-bound)]}
                else:
                    print("We meet unsupported node{}, skip.".format(node.name))
        last = None
        for t in node.output:
            last = f'{t}'
            table[last] = span(clip_val[t])
        if node.op_type == 'Relu' and last is not None:
            table[node.input[0]] = table[last]
    _write({'custom_quantize_layers': {}, 'quantize_parameters': table}, args, 'rk_quantized_param')


@deploy_dispatcher.register("rv")
def gen_rv_yaml(graph, clip_val, args, **kwargs):
    _rv1126(graph, clip_val, args)
    _rk3568(graph, clip_val, args)
