"""The N > 1 path on CPU: two gloo ranks, each calibrating its shard with the NumPy
stand-in kernels, must produce the same clip values as one rank over all images — the
statistics (range, histogram counts, per-image OCTAV values) are combined, not the clip
values, so the result is world-size invariant."""
import json
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, tmp, algo, out_q):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import fake_kernels
    from dipoorlet_b200 import forward_net as fwd
    from dipoorlet_b200 import workloads as W
    from dipoorlet_b200.cli_args import make_args
    from dipoorlet_b200.graph import ONNXGraph
    from dipoorlet_b200.tensor_cali import tensor_calibration
    fwd.K = fake_kernels
    torch.set_num_threads(1)
    model = W.build_resnet50(seed=3, blocks=[1, 1], planes=(4, 8), stem=8, num_classes=5, image=32)
    graph = ONNXGraph(model, tmp, "trt")
    args = make_args(input_dir=os.path.join(tmp, "data"), data_num=8, deploy="trt", output_dir=tmp,
                     calib_bs=3, _test_device="cpu", act_quant=algo, rank=rank, local_rank=rank,
                     world_size=world)
    act, _ = tensor_calibration(graph, args)
    out_q.put((rank, {k: [float(v[0]), float(v[1])] for k, v in act.items()}))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("algo", ["minmax", "hist", "mse"])
def test_two_ranks_equal_one_rank(tmp_path, algo):
    sys.path.insert(0, HERE)
    import fake_kernels
    from dipoorlet_b200 import forward_net as fwd
    from dipoorlet_b200 import workloads as W
    from dipoorlet_b200.cli_args import make_args
    from dipoorlet_b200.graph import ONNXGraph
    from dipoorlet_b200.tensor_cali import tensor_calibration
    tmp = str(tmp_path)
    images = W.synthetic_images(8, (3, 32, 32), seed=9)
    W.write_input_dir(images, os.path.join(tmp, "data"), "input")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, tmp, algo, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results[0] == results[1]
    # single rank over the same 8 images
    old = fwd.K
    fwd.K = fake_kernels
    try:
        fwd._SESSIONS.clear()
        model = W.build_resnet50(seed=3, blocks=[1, 1], planes=(4, 8), stem=8, num_classes=5, image=32)
        graph = ONNXGraph(model, tmp, "trt")
        args = make_args(input_dir=os.path.join(tmp, "data"), data_num=8, deploy="trt", output_dir=tmp,
                         calib_bs=3, _test_device="cpu", act_quant=algo)
        act, _ = tensor_calibration(graph, args)
    finally:
        fwd.K = old
        fwd._SESSIONS.clear()
    one = {k: [float(v[0]), float(v[1])] for k, v in act.items()}
    assert list(one) == list(results[0])
    for k in one:
        # batches are split differently (3+1 per rank vs 3+3+2): allow the last-bit wobble of
        # the batched CPU conv, nothing more
        assert np.allclose(one[k], results[0][k], rtol=2e-6, atol=1e-7), (k, one[k], results[0][k])


def _wt_worker(rank, world, port, tmp, flags, out_q):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import fake_kernels
    from dipoorlet_b200 import engine as eng
    from dipoorlet_b200 import forward_net as fwd
    from dipoorlet_b200 import onnx_lite as ol
    from dipoorlet_b200.cli_args import make_args
    from dipoorlet_b200.graph import ONNXGraph
    from dipoorlet_b200.tensor_cali import tensor_calibration
    from dipoorlet_b200.weight_transform import weight_calibration
    fwd.K = fake_kernels
    eng.K = fake_kernels
    torch.set_num_threads(1)
    model = ol.load(os.path.join(tmp, "model.onnx"))
    graph = ONNXGraph(model, tmp, "trt")
    args = make_args(input_dir=os.path.join(tmp, "data"), data_num=8, deploy="trt", output_dir=tmp,
                     calib_bs=4, _test_device="cpu", act_quant="minmax", rank=rank, local_rank=rank,
                     world_size=world, ada_bs=2, ada_epoch=3, **flags)
    act, weight = tensor_calibration(graph, args)
    g2, _, act2, weight2 = weight_calibration(graph, act, weight, args)
    inits = {k: np.asarray(v).tobytes().hex() for k, v in sorted(g2.model.graph.initializers.items())}
    out_q.put((rank, inits, {k: [float(v[0]), float(v[1])] for k, v in act2.items()}))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("flags", [{"update_bn": True}, {"sparse": True, "pattern": "nv24"}],
                         ids=["update_bn", "sparse"])
def test_weight_transforms_leave_identical_replicas(tmp_path, flags):
    """weight_trans_base.py:15-18: "after weight calibration, model / args / clip_val must be exactly the same on
    every GPU". Two gloo ranks through weight_calibration: --update_bn (rank 0 updates and saves, everybody reloads
    and re-calibrates collectively) and --sparse (each rank finetunes on its shard, gradients averaged per step)."""
    from dipoorlet_b200 import onnx_lite as ol
    from dipoorlet_b200 import workloads as W
    tmp = str(tmp_path)
    gold = os.path.join(HERE, "golden", "tiny_preact" if "update_bn" in flags else "tiny_r50")
    ol.save(ol.load(os.path.join(gold, "model.onnx")), os.path.join(tmp, "model.onnx"))
    W.write_input_dir(np.load(os.path.join(gold, "images.npy")), os.path.join(tmp, "data"), "input")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_wt_worker, args=(r, 2, port, tmp, flags, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = {r: (i, a) for r, i, a in (q.get(timeout=600) for _ in range(2))}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results[0][0] == results[1][0]          # every initializer, bit for bit
    assert results[0][1] == results[1][1]
    before = ol.load(os.path.join(gold, "model.onnx")).graph.initializers
    changed = [k for k, v in before.items() if np.asarray(v).tobytes().hex() != results[0][0][k]]
    assert changed
