"""STPU deploy writer, `-D stpu` (dipoorlet/deploy/deploy_stpu.py:23-222) -> stpu_minmax.json.
Host-side formatting only; pinned byte for byte against the reference's file (tests/golden/*/deploy_vendors.json).
Records, in the reference's order: symmetric weight ranges `<node>_weights`; symmetric activation ranges per
tensor; a Relu / Clip input takes its output's range; `emin`, the smallest exponent the accumulator of a
Conv / ConvTranspose / Gemm (and Upsample / Corr) must represent, from the float32 exponent of
max(sqrt(n) * in_max * w_max, out_max); bias records with alpha = weight step x input step at 8 bits.
`--stpu_wg` (Winograd weight ranges) crashes in the reference (NodeProto has no get_attribute_value,
deploy_stpu.py:72-81); here it is implemented from the attributes."""
import json
import os

import numpy as np

from ..platform_settings import LAYER_HAS_WEIGHT
from .deploy_default import deploy_dispatcher


def _biased_exponent(v):
    """Biased float32 exponent e with 2**(e-127) <= v < 2**(e-126), clamped to [1, 254]; 0 for v == 0."""
    if abs(v) == 0:
        return 0
    for e in range(1, 254):
        if 2 ** (e - 127) <= v < 2 ** (e - 126):
            return e
    return 1 if v < 2 ** (-126) else 254


def _emin_accumulator(in_max, w_max, out_max, fan_in, r):
    return _biased_exponent(max(fan_in ** .5 * in_max * w_max, out_max)) - (12 - r)


def _sym(bound):
    return {'min': float(-bound), 'max': float(bound)}


def _winograd_weight_bound(ker):
    g = np.array([[2, 0, 0], [1, 1, 1], [1, -1, 1], [0, 0, 2]], dtype='float32')
    wu = np.einsum('ab,ijbc,dc->ijad', g, ker, g)
    return max(max(wu.max(), 0), -min(wu.min(), 0))


@deploy_dispatcher.register("stpu")
def gen_stpu_minmax(graph, clip_val, args, **kwargs):
    param = {}
    nodes = list(graph.graph.node)
    for node in nodes:                                   # weights
        if node.op_type in LAYER_HAS_WEIGHT:
            w = clip_val[node.input[1]]
            param[node.name + '_weights'] = _sym(max(np.abs(np.min(w[0])), np.max(w[1])))
    for t in graph.network_inputs:                       # activations
        param[t] = _sym(max(np.abs(clip_val[t][0]), clip_val[t][1]))
    for node in nodes:
        for t in node.output:
            param[t] = _sym(max(np.abs(clip_val[t][0]), clip_val[t][1]))
    for node in nodes:                                   # merged activations
        if node.op_type in ('Relu', 'Clip'):
            param[node.input[0]] = param[node.output[0]].copy()
    if getattr(args, 'stpu_wg', False):
        for node in nodes:
            if (node.op_type == 'Conv' and node.attrs.get('group', 1) == 1
                    and list(node.attrs.get('kernel_shape', [])) == [3, 3]
                    and list(node.attrs.get('strides', [1, 1])) == [1, 1] and 'layer_' + node.name not in param):
                param['layer_' + node.name] = {'wg': True}
                param[node.name + '_weights'] = _sym(_winograd_weight_bound(np.asarray(graph.get_initializer(node.input[1]))))
    for node in nodes:                                   # accumulator exponents
        out = node.output[0]
        if node.op_type in ('Upsample', 'DynamicUpsample'):
            param[out]['emin'] = _biased_exponent(param[out]['max']) - (22 - 2)
        if node.op_type in ('Conv', 'ConvTranspose'):
            ws = graph.get_tensor_shape(node.input[1])
            param[out]['emin'] = _emin_accumulator(param[node.input[0]]['max'], param[node.name + '_weights']['max'],
                                                   param[out]['max'], ws[1] * ws[2] * ws[3], 2)
        if node.op_type == 'Gemm':
            param[out]['emin'] = _emin_accumulator(param[node.input[0]]['max'], param[node.name + '_weights']['max'],
                                                   param[out]['max'], np.prod(graph.get_tensor_shape(node.input[0])), 2)
        if node.op_type == 'Corr':
            n = np.prod(graph.get_tensor_shape(node.input[0])) / node.attrs['groups']
            param[out]['emin'] = _biased_exponent(param[out]['max'] * n ** .5) - (12 - 4)
    for node in nodes:                                   # biases
        if node.op_type in ('Conv', 'ConvTranspose', 'Gemm') and len(node.input) == 3:
            w, x = param[node.name + '_weights'], param[node.input[0]]
            param[node.name + '_bias'] = {'alpha': (w['max'] - w['min']) / (2 ** 8 - 2) * ((x['max'] - x['min']) / (2 ** 8 - 2)),
                                          'zero_point': 0}
    with open(os.path.join(args.output_dir, 'stpu_minmax.json'), 'wt') as f:
        json.dump(param, f, indent=4)
