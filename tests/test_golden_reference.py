"""The oracle and the product's host logic against fixtures produced by the REFERENCE
ITSELF (tests/golden/*, written by oracle/gen_golden.py which runs /root/reference's
modules under stand-ins for onnx / onnxruntime). This is what pins the oracle."""
import copy
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODELS = ["tiny_r50", "tiny_mbv2"]


def _load(mname):
    from dipoorlet_b200 import onnx_lite as ol
    d = os.path.join(GOLD, mname)
    model = ol.load(os.path.join(d, "model.onnx"))
    images = np.load(os.path.join(d, "images.npy"))
    calib = json.load(open(os.path.join(d, "calibration.json")))
    return d, model, images, calib


@pytest.mark.parametrize("mname", MODELS)
@pytest.mark.parametrize("algo", ["minmax", "hist", "mse"])
def test_oracle_calibration_equals_reference(mname, algo):
    """oracle.pipeline (forward + statistics + clip search) == reference tensor_calibration,
    value for value, and the files written from it are byte-identical."""
    from oracle import pipeline as P
    from oracle import stats as O
    d, model, images, calib = _load(mname)
    got = P.calibrate(model, images, algo)
    want = calib[algo]["act"]
    assert list(got) == list(want)
    for k in want:
        assert float(got[k][0]) == want[k][0] and float(got[k][1]) == want[k][1], (k, got[k], want[k])
    # act_clip_val.json as save_clip_val writes it, and trt_clip_val.json
    text = json.dumps({k: [np.asarray(v[0]).tolist(), np.asarray(v[1]).tolist()] for k, v in got.items()}, indent=4)
    assert text == calib[algo]["act_clip_val_json"]
    trt = json.dumps({"blob_range": {k: float(v) for k, v in O.trt_blob_range(got).items()}}, indent=4)
    assert trt == calib[algo]["trt_clip_val_json"]


@pytest.mark.parametrize("mname", MODELS)
def test_product_files_byte_identical(mname, tmp_path):
    """save_clip_val / load_clip_val / to_deploy('trt') of the product reproduce the
    reference's files byte for byte when fed the reference's clip values."""
    from dipoorlet_b200.cli_args import make_args
    from dipoorlet_b200.deploy import to_deploy
    from dipoorlet_b200.graph import ONNXGraph
    from dipoorlet_b200.tensor_cali import find_clip_val_minmax_weight
    from dipoorlet_b200.utils import load_clip_val, save_clip_val
    d, model, images, calib = _load(mname)
    graph = ONNXGraph(model, str(tmp_path), "trt")
    args = make_args(input_dir="unused", data_num=8, deploy="trt", output_dir=str(tmp_path))
    weight = find_clip_val_minmax_weight(graph, args)
    gold_w = np.load(os.path.join(d, "weight_clip.npz"))
    assert sorted(f"{k}|{i}" for k in weight for i in (0, 1)) == sorted(gold_w.files)
    for k in weight:
        assert np.array_equal(weight[k][0], gold_w[f"{k}|0"]) and np.array_equal(weight[k][1], gold_w[f"{k}|1"])
    for algo in ("minmax", "hist", "mse"):
        act = {k: [np.float32(v[0]), np.float32(v[1])] for k, v in calib[algo]["act"].items()}
        save_clip_val(act, copy.deepcopy(weight), args)
        assert open(tmp_path / "act_clip_val.json").read() == calib[algo]["act_clip_val_json"]
        act2, w2 = load_clip_val(args)
        to_deploy(graph, act2, w2, args)
        assert open(tmp_path / "trt_clip_val.json").read() == calib[algo]["trt_clip_val_json"]


@pytest.mark.parametrize("mname", MODELS)
def test_quant_graph_equals_reference(mname, tmp_path):
    """Which tensors get Q/DQ pairs (merged ReLU, TensorRT add-merge, per-channel weights),
    node order, names, rewiring and the scale / zero-point initializers."""
    from dipoorlet_b200.cli_args import make_args
    from dipoorlet_b200.graph import ONNXGraph
    from dipoorlet_b200.quantize import quant_graph
    d, model, images, calib = _load(mname)
    graph = ONNXGraph(model, str(tmp_path), "trt")
    args = make_args(input_dir="unused", data_num=8, deploy="trt", output_dir=str(tmp_path))
    gold_w = np.load(os.path.join(d, "weight_clip.npz"))
    clip = {k: [np.float64(v[0]), np.float64(v[1])] for k, v in calib["minmax"]["act"].items()}
    for key in gold_w.files:
        name, i = key.rsplit("|", 1)
        clip.setdefault(name, [None, None])[int(i)] = gold_w[key]
    gq, qlist = quant_graph(graph, copy.deepcopy(clip), args)
    want = json.load(open(os.path.join(d, "quant_graph.json")))
    assert [n.name for n in qlist] == want["quant_node_list"]
    got_nodes = [[n.op_type, n.name, list(n.input), list(n.output),
                  {k: v for k, v in n.attrs.items() if k == "axis"}] for n in gq.graph.node]
    assert got_nodes == want["nodes"]
    qp = np.load(os.path.join(d, "quant_params.npz"))
    inits = gq.model.graph.initializers
    for name in qp.files:
        assert name in inits, name
        assert inits[name].dtype == qp[name].dtype and np.array_equal(inits[name].reshape(-1), qp[name].reshape(-1)), name


@pytest.mark.parametrize("mname", MODELS)
def test_oracle_qdq_forward_equals_reference(mname, tmp_path):
    """fp and Q/DQ graph activations (ActivationCache of the reference) == the oracle forward
    on the product's Q/DQ graph: pins the fake-quant semantics end to end."""
    from dipoorlet_b200.cli_args import make_args
    from dipoorlet_b200.graph import ONNXGraph
    from dipoorlet_b200.quantize import quant_graph
    from oracle import forward as OF
    d, model, images, calib = _load(mname)
    graph = ONNXGraph(model, str(tmp_path), "trt")
    args = make_args(input_dir="unused", data_num=8, deploy="trt", output_dir=str(tmp_path))
    gold_w = np.load(os.path.join(d, "weight_clip.npz"))
    clip = {k: [np.float64(v[0]), np.float64(v[1])] for k, v in calib["minmax"]["act"].items()}
    for key in gold_w.files:
        name, i = key.rsplit("|", 1)
        clip.setdefault(name, [None, None])[int(i)] = gold_w[key]
    gq, _ = quant_graph(graph, copy.deepcopy(clip), args)
    gold = np.load(os.path.join(d, "qforward.npz"))
    n = images.shape[0]
    fp = OF.blobs_for_images(model, {"input": images}, n)
    q = OF.blobs_for_images(gq.model, {"input": images}, n)
    for key in gold.files:
        kind, t = key.split("|", 1)
        src = fp[t] if kind == "fp" else q[t if t in q else t + "_dq"]
        assert np.array_equal(np.stack(src), gold[key]), key


def test_quant_params_equal_reference_on_all_platform_parameter_sets():
    """get_qnode_by_param (quantize.py:111-194) against the reference's own outputs on 208 seeded cases:
    the weight and activation parameter sets of all eight platforms (symmetric / asymmetric, per-tensor /
    per-channel, log_scale, dynamic_sym), scalar and per-channel ranges incl. zero ranges, one-sided ranges
    and all-zero channels (tests/golden/qparam_fuzz.json, oracle/gen_golden_qparams.py). Scales must be the
    same float32 bit patterns, zero points / q limits the same integers, dtypes and axis attributes equal;
    and the product's platform table must carry the same parameter sets."""
    from dipoorlet_b200.platform_settings import platform_setting_table
    from dipoorlet_b200.quantize import get_qnode_by_param
    gold = json.load(open(os.path.join(GOLD, "qparam_fuzz.json")))
    for platform, sets in gold["params"].items():
        for key, want in sets.items():
            assert platform_setting_table[platform][key] == want, (platform, key)
    assert len(gold["rows"]) == 208
    for row in gold["rows"]:
        param = gold["params"][row["platform"]][row["key"]]
        rr = [np.array(v, dtype=np.float64) if isinstance(v, list) else np.float64(v) for v in row["range"]]
        q_nodes, q_min, q_max = get_qnode_by_param(param, "t", row["shape"], copy.deepcopy(rr))
        inits = dict(q_nodes.initializer)
        scale, zp = np.asarray(inits["t_scale"]), np.asarray(inits["t_zero_point"])
        tag = (row["platform"], row["key"], row["range"])
        assert str(scale.dtype) == row["scale_dtype"] and str(zp.dtype) == row["zp_dtype"], tag
        assert np.array_equal(scale.reshape(-1), np.asarray(row["scale"], dtype=np.float32)), tag
        assert zp.reshape(-1).astype(int).tolist() == row["zero_point"], tag
        assert np.asarray(q_min).reshape(-1).astype(int).tolist() == row["q_min"], tag
        assert np.asarray(q_max).reshape(-1).astype(int).tolist() == row["q_max"], tag
        assert [n.attrs.get("axis") for n in q_nodes.node] == row["axis"], tag


@pytest.mark.parametrize("mname", MODELS)
@pytest.mark.parametrize("platform", ["trt", "stpu", "magicmind", "rv", "atlas", "snpe", "ti", "imx"])
def test_quant_graph_equals_reference_on_every_platform(mname, platform, tmp_path):
    """The Q/DQ rewrite (quantize.py:20-108) for each of the eight deploy platforms of the reference's table:
    same quantised-node list, node-for-node identical graph (names, wiring, axis attributes), same network
    outputs (platforms that quantise them) and bit-identical scale / zero-point initializers
    (tests/golden/*/quant_graph_platforms.json, oracle/gen_golden_quant_graphs.py)."""
    from dipoorlet_b200.cli_args import make_args
    from dipoorlet_b200.graph import ONNXGraph
    from dipoorlet_b200.quantize import quant_graph
    d, model, images, calib = _load(mname)
    want = json.load(open(os.path.join(d, "quant_graph_platforms.json")))[platform]
    graph = ONNXGraph(model, str(tmp_path), platform)
    args = make_args(input_dir="unused", data_num=8, deploy=platform, output_dir=str(tmp_path))
    gold_w = np.load(os.path.join(d, "weight_clip.npz"))
    clip = {k: [np.float64(v[0]), np.float64(v[1])] for k, v in calib["minmax"]["act"].items()}
    for key in gold_w.files:
        name, i = key.rsplit("|", 1)
        clip.setdefault(name, [None, None])[int(i)] = gold_w[key]
    gq, qlist = quant_graph(graph, copy.deepcopy(clip), args)
    assert [n.name for n in qlist] == want["quant_node_list"]
    got_nodes = [[n.op_type, n.name, list(n.input), list(n.output),
                  {k: v for k, v in n.attrs.items() if k == "axis"}] for n in gq.graph.node]
    assert got_nodes == want["nodes"]
    assert list(gq.network_outputs) == want["network_outputs"]
    inits = gq.model.graph.initializers
    assert sorted(k for k in inits if k.endswith("_scale") or k.endswith("_zero_point")) == sorted(want["qparams"])
    for name, (dtype, vals) in want["qparams"].items():
        assert str(inits[name].dtype) == dtype, name
        assert np.array_equal(np.asarray(inits[name], dtype=np.float64).reshape(-1), np.asarray(vals)), name


def test_clip_and_profiling_files_equal_reference_over_three_ranks(tmp_path):
    """The on-disk boundary (utils.py:313-412): per-rank act / weight clip files, the rank-0 combine
    (min / max for minmax, mean of clips otherwise), the reload types (np.float64 scalars; per-channel arrays
    only on per-channel platforms) and the per-rank profiling files with their mean / min combine - every
    file byte for byte, every value and type equal to what the reference produced on the same 3-rank inputs
    (tests/golden/persistence.json, oracle/gen_golden_persistence.py)."""
    import types
    from dipoorlet_b200 import utils as U
    gold = json.load(open(os.path.join(GOLD, "persistence.json")))
    ranks = gold["ranks"]
    for case, want in gold["cases"].items():
        deploy, algo = case.split("|")
        out = tmp_path / case.replace("|", "_")
        out.mkdir()
        args = types.SimpleNamespace(output_dir=str(out), act_quant=algo, deploy=deploy, model_type=None)
        for r in range(ranks):
            act = {k: [np.float32(v[0]), np.float32(v[1])] for k, v in gold["act_ranks"][r].items()}
            weight = {k: [np.asarray(v[0], np.float32), np.asarray(v[1], np.float32)] for k, v in gold["weight"].items()}
            U.save_clip_val(act, weight, args, act_fname=f"act_clip_val.json.rank{r}",
                            weight_fname=f"weight_clip_val.json.rank{r}")
        U.reduce_clip_val(ranks, args)
        act, w = U.load_clip_val(args)
        files = {f: open(os.path.join(str(out), f)).read() for f in sorted(os.listdir(str(out)))}
        assert sorted(files) == sorted(want["files"]), case
        for f in files:
            assert files[f] == want["files"][f], (case, f)
        assert list(act) == list(want["act"])
        for k, (lo, hi, tname) in want["act"].items():
            assert float(act[k][0]) == lo and float(act[k][1]) == hi and type(act[k][0]).__name__ == tname, (case, k)
        for k, (lo, hi, tname, shape) in want["weight"].items():
            assert np.asarray(w[k][0]).tolist() == lo and np.asarray(w[k][1]).tolist() == hi, (case, k)
            assert type(w[k][0]).__name__ == tname and list(np.asarray(w[k][0]).shape) == shape, (case, k)
    out = tmp_path / "prof"
    out.mkdir()
    args = types.SimpleNamespace(output_dir=str(out), model_type=None)
    for r in range(ranks):
        U.save_profiling_res(dict(gold["layer_ranks"][r]), {k: list(v) for k, v in gold["model_ranks"][r].items()},
                             args, rank=r)
    files = {f: open(os.path.join(str(out), f)).read() for f in sorted(os.listdir(str(out)))}
    assert files == gold["profiling"]["files"]
    layer, model = U.reduce_profiling_res(ranks, args)
    assert layer == gold["profiling"]["layer"] and model == gold["profiling"]["model"]


@pytest.mark.parametrize("mname", MODELS)
def test_graph_ir_equals_reference(mname, tmp_path):
    """ONNXGraph, the first argument of every registry function (utils.py:22-250): node names in order,
    network inputs / outputs, initializer names, tensor shapes, producer and consumer maps as the
    reference's own class reports them (tests/golden/*/graph_api.json, oracle/gen_golden_graph_api.py)."""
    from dipoorlet_b200.graph import ONNXGraph
    d, model, images, calib = _load(mname)
    want = json.load(open(os.path.join(d, "graph_api.json")))
    g = ONNXGraph(model, str(tmp_path), "trt")
    assert [[n.op_type, n.name, list(n.input), list(n.output)] for n in g.graph.node] == want["nodes"]
    assert list(g.network_inputs) == want["network_inputs"]
    assert list(g.network_outputs) == want["network_outputs"]
    assert sorted(g.initializer.keys()) == want["initializers"]
    for t, shape in want["shapes"].items():
        assert [int(v) for v in g.get_tensor_shape(t)] == shape, t
    for t, p in want["producer"].items():
        got = g.get_tensor_producer(t)
        assert (got if isinstance(got, str) else got.name) == p, t
    for t, cons in want["consumer"].items():
        got = [(c if isinstance(c, str) else c.name) for c in g.get_tensor_consumer(t)]
        assert got == cons, t


@pytest.mark.parametrize("mname", MODELS)
@pytest.mark.parametrize("platform", ["atlas", "imx", "magicmind", "snpe", "ti", "rv", "stpu"])
def test_vendor_deploy_files_byte_identical(mname, platform, tmp_path):
    """to_deploy for the seven vendor formats (rv: four files incl. the YAML anchors of merged records): fed the clip-value files the reference wrote, the
    product reloads them (load_clip_val: scalar vs per-channel weights per platform) and must write the same
    deploy files byte for byte (tests/golden/*/deploy_vendors.json, oracle/gen_golden_deploy.py)."""
    from dipoorlet_b200.cli_args import make_args
    from dipoorlet_b200.deploy import to_deploy
    from dipoorlet_b200.graph import ONNXGraph
    from dipoorlet_b200.utils import load_clip_val
    d, model, images, calib = _load(mname)
    gold = json.load(open(os.path.join(d, "deploy_vendors.json")))[platform]
    (tmp_path / "act_clip_val.json").write_text(gold["act_clip_val"])
    (tmp_path / "weight_clip_val.json").write_text(gold["weight_clip_val"])
    graph = ONNXGraph(model, str(tmp_path), platform)
    args = make_args(input_dir="unused", data_num=8, deploy=platform, output_dir=str(tmp_path))
    act, weight = load_clip_val(args)
    to_deploy(graph, act, weight, args)
    for fname, text in gold["files"].items():
        assert (tmp_path / fname).read_text() == text, fname


def test_stpu_winograd_weight_ranges(tmp_path):
    """`-D stpu --stpu_wg`: the reference crashes here (deploy_stpu.py:72-81 calls NodeProto.get_attribute_value), so
    there is no fixture; the product's output is checked against the rule itself: every 3x3 stride-1 group-1 Conv
    gets `layer_<name>: {wg: true}` and the symmetric range of G k G^T over its kernels, everything else is
    unchanged from the file without the flag."""
    from dipoorlet_b200.cli_args import make_args
    from dipoorlet_b200.deploy import to_deploy
    from dipoorlet_b200.graph import ONNXGraph
    from dipoorlet_b200.utils import load_clip_val
    d, model, images, calib = _load("tiny_r50")
    gold = json.load(open(os.path.join(d, "deploy_vendors.json")))["stpu"]
    (tmp_path / "act_clip_val.json").write_text(gold["act_clip_val"])
    (tmp_path / "weight_clip_val.json").write_text(gold["weight_clip_val"])
    graph = ONNXGraph(model, str(tmp_path), "stpu")
    args = make_args(input_dir="unused", data_num=8, deploy="stpu", output_dir=str(tmp_path), stpu_wg=True)
    act, weight = load_clip_val(args)
    to_deploy(graph, act, weight, args)
    got = json.load(open(tmp_path / "stpu_minmax.json"))
    plain = json.loads(gold["files"]["stpu_minmax.json"])
    G = np.array([[2, 0, 0], [1, 1, 1], [1, -1, 1], [0, 0, 2]], dtype=np.float32)
    wg_nodes = [n for n in graph.graph.node if n.op_type == "Conv" and n.attrs.get("group", 1) == 1
                and list(n.attrs["kernel_shape"]) == [3, 3] and list(n.attrs.get("strides", [1, 1])) == [1, 1]]
    assert wg_nodes
    for n in wg_nodes:
        assert got["layer_" + n.name] == {"wg": True}
        w = np.asarray(graph.get_initializer(n.input[1]))
        bound = max(np.abs(G.dot(w[i, j]).dot(G.T)).max() for i in range(w.shape[0]) for j in range(w.shape[1]))
        assert got[n.name + "_weights"]["max"] == pytest.approx(float(bound), rel=1e-6)
        assert got[n.name + "_weights"]["min"] == -got[n.name + "_weights"]["max"]
    touched = {n.name for n in wg_nodes}
    for k, v in plain.items():
        owner = k.rsplit("_", 1)[0]
        if owner in touched and k.endswith(("_weights", "_bias")):
            continue
        if isinstance(v, dict) and "emin" in v and any(k in n.output for n in wg_nodes):
            continue
        assert got[k] == v, k


@pytest.mark.parametrize("mname", MODELS)
def test_quant_graph_with_skip_layers_equals_reference(mname, tmp_path):
    """--skip_layers (quantize.py:29-31): the named layers get no Q/DQ on their weight or inputs and are not in the
    quantised-node list; the rewritten graph is the reference's node for node
    (tests/golden/*/quant_graph_skip_layers.json, oracle/gen_golden_skip_layers.py)."""
    from dipoorlet_b200.cli_args import make_args
    from dipoorlet_b200.graph import ONNXGraph
    from dipoorlet_b200.quantize import quant_graph
    d, model, images, calib = _load(mname)
    want = json.load(open(os.path.join(d, "quant_graph_skip_layers.json")))
    full = json.load(open(os.path.join(d, "quant_graph_platforms.json")))["trt"]
    graph = ONNXGraph(model, str(tmp_path), "trt")
    args = make_args(input_dir="unused", data_num=8, deploy="trt", output_dir=str(tmp_path),
                     skip_layers=want["skip_layers"])
    gold_w = np.load(os.path.join(d, "weight_clip.npz"))
    clip = {k: [np.float64(v[0]), np.float64(v[1])] for k, v in calib["minmax"]["act"].items()}
    for key in gold_w.files:
        name, i = key.rsplit("|", 1)
        clip.setdefault(name, [None, None])[int(i)] = gold_w[key]
    gq, qlist = quant_graph(graph, copy.deepcopy(clip), args)
    assert [n.name for n in qlist] == want["quant_node_list"]
    assert not set(want["skip_layers"]) & set(want["quant_node_list"])
    assert [[n.op_type, n.name, list(n.input), list(n.output)] for n in gq.graph.node] == want["nodes"]
    assert len(want["nodes"]) < len(full["nodes"])          # the fixture does exercise the flag


@pytest.mark.parametrize("mname", MODELS)
@pytest.mark.parametrize("platform", ["trt", "snpe"])
def test_report_lines_equal_reference(mname, platform, tmp_path):
    """The user-visible report — show_model_profiling_res, show_model_ranges, weight_need_perchannel
    (profiling.py:199-260) — line for line what the reference logs for the same numbers
    (tests/golden/*/reports.json, oracle/gen_golden_reports.py): per-layer cosines, the ten worst layers,
    output cosines, every range with its shape, and on a per-tensor platform the per-layer degradation ranking."""
    import logging
    from dipoorlet_b200 import profiling as prof
    from dipoorlet_b200.cli_args import make_args
    from dipoorlet_b200.graph import ONNXGraph
    from dipoorlet_b200.quantize import quant_graph
    from dipoorlet_b200.utils import logger
    d, model, images, calib = _load(mname)
    gold = json.load(open(os.path.join(d, "reports.json")))[platform]
    pb = json.load(open(os.path.join(d, "profiling_bc.json")))
    graph = ONNXGraph(model, str(tmp_path), platform)
    args = make_args(input_dir="unused", data_num=8, deploy=platform, output_dir=str(tmp_path))
    gold_w = np.load(os.path.join(d, "weight_clip.npz"))
    act = {k: [np.float64(v[0]), np.float64(v[1])] for k, v in calib["minmax"]["act"].items()}
    weight = {}
    for key in gold_w.files:
        name, i = key.rsplit("|", 1)
        weight.setdefault(name, [None, None])[int(i)] = gold_w[key]
    if platform == "snpe":
        weight = {k: [np.min(v[0]), np.max(v[1])] for k, v in weight.items()}
    clip = dict(act)
    clip.update(weight)
    _, qlist = quant_graph(graph, copy.deepcopy(clip), args)
    layer = {t: gold["layer"][t] for n in qlist for t in n.output}
    model_cos = {k: list(v) for k, v in pb["model"].items()}

    class Lines(logging.Handler):
        def __init__(self):
            super().__init__()
            self.lines = []

        def emit(self, record):
            self.lines.append(record.getMessage())

    h = Lines()
    logger.addHandler(h)
    old = logger.level
    logger.setLevel(logging.INFO)
    try:
        prof.show_model_profiling_res(graph, layer, model_cos, qlist, args)
        prof.show_model_ranges(graph, act, weight, args)
        prof.weight_need_perchannel(graph, args)
    finally:
        logger.removeHandler(h)
        logger.setLevel(old)
    assert h.lines == gold["lines"]
