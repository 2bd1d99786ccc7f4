"""Generate tests/golden/* by executing the REFERENCE ITSELF (/root/reference/dipoorlet,
imported unmodified) under the stand-ins of oracle/ref_shim on small seeded models.

    python oracle/gen_golden.py            # rewrites tests/golden/

Run in the build container only (the GPU box has no /root/reference); the fixtures it
writes are committed. What is reference code and what is stand-in is listed in
oracle/ref_shim/__init__.py: only the ONNX container classes and the operator arithmetic
of onnxruntime (restated with torch CPU fp32, oracle/forward.run_node) are ours.
"""
import copy
import json
import logging
import os
import shutil
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torchvision  # noqa: E402,F401  (must be imported before the stand-ins are installed)
import torch  # noqa: E402

from oracle import ref_shim  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
N_IMG = 8
ADA_BS = 4
ADA_EPOCH = 12


def jsonable_clip(d):
    return {k: [np.asarray(v[0]).tolist(), np.asarray(v[1]).tolist()] for k, v in d.items()}


def main():
    from dipoorlet_b200 import onnx_lite as ol, workloads as W
    ref_shim.install()
    import dipoorlet.tensor_cali as RTC
    import dipoorlet.utils as RU
    from dipoorlet.deploy import to_deploy
    from dipoorlet.forward_net import ActivationCache
    from dipoorlet.profiling import quantize_profiling_multipass
    from dipoorlet.quantize import quant_graph
    from dipoorlet.weight_transform import weight_calibration
    logging.getLogger("dipoorlet").setLevel(logging.WARNING)

    models = {
        "tiny_r50": W.build_resnet50(seed=11, blocks=[1, 2, 1], planes=(8, 16, 32), stem=16, num_classes=10,
                                     image=32),
        "tiny_mbv2": W.build_mobilenetv2(seed=12, width_mult=0.25, num_classes=10, image=32),
    }
    for mname, model in models.items():
        out = os.path.join(GOLD, mname)
        shutil.rmtree(out, ignore_errors=True)
        os.makedirs(out)
        ol.save(model, os.path.join(out, "model.onnx"))
        images = W.synthetic_images(N_IMG, (3, 32, 32), seed=21)
        np.save(os.path.join(out, "images.npy"), images)

        def fresh(**kw):
            tmp = tempfile.mkdtemp(prefix="dpl_gold_")
            W.write_input_dir(images, os.path.join(tmp, "data"), "input")
            g = RU.ONNXGraph(ref_shim.from_lite(model), tmp, "trt", None)
            args = types.SimpleNamespace(
                input_dir=os.path.join(tmp, "data"), output_dir=tmp, data_num=N_IMG, world_size=1, rank=0,
                local_rank=0, act_quant="minmax", deploy="trt", bins=2048, threshold=0.99999,
                optim_transformer=False, skip_layers=[], bc=False, we=False, update_bn=False, adaround=False,
                brecq=False, drop=False, sparse=False, ada_bs=ADA_BS, ada_epoch=ADA_EPOCH, model=None,
                model_type=None, savefp=False, skip_prof_layer=False)
            for k, v in kw.items():
                setattr(args, k, v)
            return g, args, tmp

        def calibrate(g, args):
            """__main__.py:119-128: calibrate, per-rank files, reduce, reload."""
            act, w = RTC.tensor_calibration(g, args)
            RU.save_clip_val(act, w, args, act_fname="act_clip_val.json.rank0",
                             weight_fname="weight_clip_val.json.rank0")
            RU.reduce_clip_val(1, args)
            return RU.load_clip_val(args)

        # ---- calibration + deploy, the three algorithms --------------------------------
        calib = {}
        for algo in ("minmax", "hist", "mse"):
            g, args, tmp = fresh(act_quant=algo)
            act, w = calibrate(g, args)
            entry = {"act_clip_val_json": open(os.path.join(tmp, "act_clip_val.json")).read(),
                     "act": jsonable_clip(act)}
            if algo == "minmax":
                np.savez_compressed(os.path.join(out, "weight_clip.npz"),
                                    **{f"{k}|{i}": np.asarray(v[i]) for k, v in w.items() for i in (0, 1)})
            to_deploy(g, act, w, args)
            entry["trt_clip_val_json"] = open(os.path.join(tmp, "trt_clip_val.json")).read()
            calib[algo] = entry
            shutil.rmtree(tmp)
        json.dump(calib, open(os.path.join(out, "calibration.json"), "w"), indent=1)

        # ---- which tensors get Q/DQ, their parameters, and the quantised forward ------------
        g, args, tmp = fresh()
        act, w = calibrate(g, args)
        clip = act.copy()
        clip.update(w)
        gq, qlist = quant_graph(g, copy.deepcopy(clip), args)
        qg = {"quant_node_list": [n.name for n in qlist],
              "nodes": [[n.op_type, n.name, list(n.input), list(n.output),
                         {a.name: a.value for a in n.attribute if a.name == "axis"}] for n in gq.graph.node]}
        json.dump(qg, open(os.path.join(out, "quant_graph.json"), "w"), indent=1)
        qparams = {}
        for t in gq.graph.initializer:
            if t.name.endswith("_scale") or t.name.endswith("_zero_point"):
                qparams[t.name] = t.array
        np.savez_compressed(os.path.join(out, "quant_params.npz"), **qparams)
        fp_cache = ActivationCache(g, args)
        q_cache = ActivationCache(gq, args)
        net_out = g.network_outputs[0]
        keep = {"fp|" + net_out: np.stack(fp_cache[net_out]), "q|" + net_out: np.stack(q_cache[net_out])}
        mids = [n.output[0] for n in g.graph.node if n.op_type in ("Conv", "Add")][2:8:2]
        for t in mids:
            keep["fp|" + t] = np.stack(fp_cache[t])
            keep["q|" + t] = np.stack(q_cache[t])
        np.savez_compressed(os.path.join(out, "qforward.npz"), **keep)
        shutil.rmtree(tmp)

        # ---- weight transforms (each alone: bc + adaround together crashes in the reference
        #      because weight_trans_base.py:27 reloads the graph without `deploy`) ------------
        def weights_of(graph):
            return {t.name: np.asarray(t.array) for t in graph.graph.initializer}

        for tag, kw in (("bc", dict(bc=True)), ("adaround", dict(adaround=True)),
                        ("brecq", dict(brecq=True))):
            g, args, tmp = fresh(**kw)
            act, w = calibrate(g, args)
            torch.manual_seed(0)
            graph, graph_ori, act2, w2 = weight_calibration(g, act, w, args)
            before = {k: v for k, v in model.graph.initializers.items()}
            changed = {k: v for k, v in weights_of(graph).items()
                       if k not in before or before[k].shape != v.shape or not np.array_equal(before[k], v)}
            np.savez_compressed(os.path.join(out, f"wt_{tag}.npz"), **changed)
            layer_cos, model_cos, _ = quantize_profiling_multipass(graph, graph_ori, act2, w2, args)
            json.dump({"layer": {k: float(v) for k, v in layer_cos.items()},
                       "model": {k: [float(v[0]), float(v[1])] for k, v in model_cos.items()}},
                      open(os.path.join(out, f"profiling_{tag}.json"), "w"), indent=1)
            shutil.rmtree(tmp)
        print("wrote", out, sorted(os.listdir(out)))


if __name__ == "__main__":
    main()
