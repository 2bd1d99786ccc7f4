"""`--sparse` on the GPU (runs last among the GPU files — not the measured path). The finetune is
sensitive to the last bit of its inputs (tests/test_host_logic.py pins it against the reference on the CPU, where
the arithmetic is the reference's own), so here only what must hold on any device: the pattern, the grid, the file."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("pattern", ["unstruction", "nv24"])
def test_sparse_quant_pattern_and_grid(dpl_built, pattern, tmp_path):
    from dipoorlet_b200 import onnx_lite as ol
    from dipoorlet_b200.cli_args import make_args
    from dipoorlet_b200.forward_net import ArrayInput
    from dipoorlet_b200.graph import ONNXGraph
    from dipoorlet_b200.tensor_cali import tensor_calibration
    from dipoorlet_b200.weight_transform import weight_calibration
    d = os.path.join(GOLD, "tiny_r50")
    model = ol.load(os.path.join(d, "model.onnx"))
    images = np.load(os.path.join(d, "images.npy"))
    graph = ONNXGraph(model, str(tmp_path), "trt")
    args = make_args(input_dir=ArrayInput({"input": images[:, 0]}), data_num=images.shape[0], deploy="trt",
                     act_quant="minmax", output_dir=str(tmp_path), calib_bs=8, sparse=True, sparse_rate=0.5,
                     pattern=pattern, ada_bs=4, ada_epoch=4)
    act, weight = tensor_calibration(graph, args)
    g2, _, _, _ = weight_calibration(graph, act, weight, args)
    assert os.path.exists(os.path.join(str(tmp_path), "sparse_quant.onnx"))
    for node in graph.graph.node:
        if node.op_type not in ("Conv", "Gemm"):
            continue
        name = node.input[1]
        got = g2.get_initializer(name)
        step = (np.maximum(np.abs(weight[name][0]), np.abs(weight[name][1])) / 127).astype(np.float32)
        q = got / np.where(step == 0, 1, step).reshape([-1] + [1] * (got.ndim - 1))
        assert np.allclose(q, np.round(q), atol=1e-3) and np.abs(q).max() <= 127 + 1e-3, name
        if pattern == "nv24":
            groups = (np.transpose(got, (0, 2, 3, 1)) if got.ndim == 4 else got).reshape(-1, 4)
            assert ((groups == 0).sum(axis=1) >= 2).all(), name
        else:
            assert (got == 0).sum() >= int(0.5 * got.size), name
