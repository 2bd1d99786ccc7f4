"""Both model families of BASELINE.json at full width (ResNet-50, MobileNetV2; 64 x 64 images to keep it short)
through adaround and brecq + drop with DPL_STRICT_NATIVE=1: every learnable layer must find a libdpl_b200
contraction for its forward, weight gradient and data gradient (a torch / cuDNN convolution raises), and every
rounded weight must lie on its quantisation grid (adaround.py:100-110, brecq.py:130-150)."""
import copy

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("family", ["r50", "mbv2"])
@pytest.mark.parametrize("algo", ["adaround", "brecq"])
def test_family_runs_on_native_contractions(dpl_built, family, algo, tmp_path, monkeypatch):
    import torch
    from dipoorlet_b200 import forward_net as fwd, kernels as K, workloads as W
    from dipoorlet_b200.cli_args import make_args
    from dipoorlet_b200.graph import ONNXGraph
    from dipoorlet_b200.tensor_cali import tensor_calibration
    from dipoorlet_b200.weight_transform import weight_calibration
    monkeypatch.setenv("DPL_STRICT_NATIVE", "1")
    n = 8
    model = (W.build_resnet50(seed=0, num_classes=40, image=64) if family == "r50"
             else W.build_mobilenetv2(seed=0, num_classes=40, image=64))
    graph = ONNXGraph(model, str(tmp_path), "trt")
    images = W.synthetic_images(n, (3, 64, 64), seed=2)
    kw = dict(adaround=True) if algo == "adaround" else dict(brecq=True, drop=True)
    args = make_args(input_dir=fwd.ArrayInput({"input": images[:, 0]}), data_num=n, deploy="trt", act_quant="minmax",
                     output_dir=str(tmp_path), ada_bs=4, ada_epoch=1, calib_bs=8, **kw)
    act, weight = tensor_calibration(graph, args)
    act = {k: [np.float64(v[0]), np.float64(v[1])] for k, v in act.items()}
    l0 = K.launches()
    graph_wt, _, _, _ = weight_calibration(graph, act, copy.deepcopy(weight), args)
    K.gemm_check_errors()
    assert K.launches() > l0
    n_layers = 0
    for node in graph.graph.node:
        if node.op_type in ("Conv", "Gemm"):
            w = graph_wt.get_initializer(node.input[1])
            lo, hi = weight[node.input[1]]
            s = (np.maximum(np.abs(lo), np.abs(hi)).astype(np.float64) / 127).astype(np.float32)
            s = np.where(s == 0, 1, s).reshape([-1] + [1] * (w.ndim - 1))
            q = w / s
            assert np.allclose(q, np.round(q), atol=1e-3), node.name
            assert np.abs(q).max() <= 127 + 1e-3
            n_layers += 1
    assert n_layers >= 53
    torch.cuda.synchronize()
