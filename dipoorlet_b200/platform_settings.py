"""Per-platform quantisation settings (data restated from
dipoorlet/platform_settings.py:1-184). Only 'trt' is exercised by the B200 hot path
(every BASELINE.json config deploys to TensorRT); the other rows are carried so that the
`-D` choices and registry dispatch of the reference CLI keep resolving."""

LAYER_HAS_WEIGHT = ['Conv', 'Gemm', 'ConvTranspose', 'PRelu', 'BatchNormalization']

_COMMON_NODES = ['Relu', 'Eltwise', 'MaxPool', 'Conv', 'Gemm', 'ConvTranspose', 'PRelu',
                 'AveragePool', 'Concat', 'Split', 'Add', 'Mul', 'Abs', 'Reciprocal', 'Sigmoid']
basic_quant_node = _COMMON_NODES


def _lin(symmetric, bit_width=8, **extra):
    d = {'bit_width': bit_width, 'type': 'Linear', 'symmetric': symmetric}
    d.update(extra)
    return d


def _platform(quant_nodes, qw, qi, quantize_network_output, deploy_weight, exclude=True):
    d = {'quant_nodes': quant_nodes, 'qw_params': qw, 'qi_params': qi,
         'quantize_network_output': quantize_network_output, 'deploy_weight': deploy_weight}
    if exclude:
        d['deploy_exclude_layers'] = []
    return d


platform_setting_table = {
    # TensorRT: symmetric int8, per-channel weights, activations-only clip file
    'trt': _platform(['Relu', 'MaxPool', 'Conv', 'Gemm', 'ConvTranspose', 'PRelu', 'AveragePool',
                      'Add', 'Sigmoid'],
                     _lin(True, per_channel=True), _lin(True), False, False),
    'stpu': _platform(_COMMON_NODES + ['Clip', 'HardSigmoid'],
                      _lin(True, per_channel=False), _lin(True), False, True),
    'magicmind': _platform(['Gemm', 'Conv', 'ConvTranspose', 'MatMul'],
                           _lin(False, log_scale=False, per_channel=True),
                           _lin(False, log_scale=False), False, False),
    'rv': _platform(_COMMON_NODES, _lin(False, per_channel=False), _lin(False), True, True),
    'atlas': _platform(['Conv', 'Gemm', 'AveragePool'], _lin(True, per_channel=True),
                       _lin(False), False, False, exclude=False),
    'snpe': _platform(_COMMON_NODES + ['Sigmoid'], _lin(False, per_channel=False), _lin(False),
                      True, False),
    'ti': _platform(_COMMON_NODES, _lin(True, per_channel=False, log_scale=False),
                    _lin(True, dynamic_sym=True, log_scale=True), False, False),
    'imx': _platform(_COMMON_NODES, _lin(True, per_channel=True, log_scale=True),
                     _lin(True, log_scale=True), True, True),
}

trt_platform_settings = platform_setting_table['trt']
