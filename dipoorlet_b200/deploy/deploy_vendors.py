"""The five small vendor deploy writers: atlas, imx, magicmind, snpe, ti
(dipoorlet/deploy/deploy_{atlas,imx,magicmind,snpe,ti}.py). Host-side formatting of the clip values the
calibration produced — no GPU work. Every file is pinned byte for byte against what the reference itself
writes for the same clip values (tests/golden/*/deploy_vendors.json, oracle/gen_golden_deploy.py).
Like the reference's writers, imx and ti rewrite the `clip_val` dict they are given.
The two large vendor formats (rv, stpu) are not implemented."""
import json
import os

import numpy as np

from ..platform_settings import platform_setting_table
from .deploy_default import deploy_dispatcher


def _dump(obj, args, fname):
    with open(os.path.join(args.output_dir, fname), "w") as f:
        json.dump(obj, f, indent=4)


@deploy_dispatcher.register("atlas")
def gen_atlas_quant_param(graph, clip_val, args, **kwargs):
    """deploy_atlas.py:12-32 -> atlas_quant_param.json: for the first input of every quantised node type an
    8-bit affine step over the range widened to contain 0, offset shifted into [-128, 127]."""
    quant_types = platform_setting_table["atlas"]["quant_nodes"]
    res = {}
    for node in graph.graph.node:
        if node.op_type not in quant_types:
            continue
        name = node.input[0]
        lo, hi = min(0, clip_val[name][0]), max(0, clip_val[name][1])
        step = (hi - lo) / 255.
        if step == 0.0:
            step = 1.0
        res[name] = {"scale": step, "offset": int(round(-lo / step) - 128)}
    _dump(res, args, "atlas_quant_param.json")


@deploy_dispatcher.register("imx")
def gen_imx_range(graph, clip_val, args, **kwargs):
    """deploy_imx.py:9-26 -> imx_scale.json: power-of-two symmetric 8-bit scales (per channel where the range
    is per channel); bias ranges are dropped."""
    for k in [k for k in clip_val if k.endswith(".bias")]:
        del clip_val[k]
    for k in clip_val:
        scale = np.array(np.max(np.abs(clip_val[k]), axis=0)) / [127]
        if np.any(scale == 0):
            scale = np.where(scale == 0, 1., scale)
        clip_val[k] = (2 ** np.round(np.log2(scale))).tolist()
    _dump({"blob_range": clip_val}, args, "imx_scale.json")


@deploy_dispatcher.register("magicmind")
def gen_magicmind_proto(graph, clip_val, args, **kwargs):
    """deploy_magicmind.py:10-20 -> magicmind_quant_param.json: one {min, max} pair per tensor."""
    ranges = {k: {"min": float(np.min(v[0])), "max": float(np.max(v[1]))} for k, v in clip_val.items()}
    _dump({"blob_range": ranges}, args, "magicmind_quant_param.json")


@deploy_dispatcher.register("snpe")
def gen_snpe_encodings(graph, clip_val, args, **kwargs):
    """deploy_snpe.py:8-34 -> snpe_encodings.json: 8-bit activation encodings of every non-initializer node
    input and of the network outputs; max is at least 0 and at least min + 0.01; no parameter encodings."""
    def encoding(name):
        lo, hi = float(clip_val[name][0]), float(clip_val[name][1])
        return [{"bitwidth": 8, "min": lo, "max": max(max(0.0, hi), lo + 0.01)}]

    act = {}
    for node in graph.graph.node:
        for name in node.input:
            if name != "" and name not in graph.initializer:
                act[name] = encoding(name)
    for name in graph.network_outputs:
        act[name] = encoding(name)
    _dump({"activation_encodings": act, "param_encodings": {}}, args, "snpe_encodings.json")


@deploy_dispatcher.register("ti")
def gen_ti_json(graph, clip_val, args, **kwargs):
    """deploy_ti.py:8-19 -> ti_blob_range.txt ("name lo hi" per line, values as Python prints them) and
    ti_blob_range.json (the same ranges as floats)."""
    with open(os.path.join(args.output_dir, "ti_blob_range.txt"), "w") as f:
        for k, v in clip_val.items():
            f.write("{} {} {}\n".format(k, v[0], v[1]))
    for k, v in clip_val.items():
        clip_val[k] = [float(x) for x in v]
    _dump({"blob_range": clip_val}, args, "ti_blob_range.json")
