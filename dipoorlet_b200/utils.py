"""Host-side helpers of the plugin boundary: the registry decorator, the logger and the
clip-value / profiling JSON files (formats of dipoorlet/utils.py:281-412 kept so that
downstream tools reading `act_clip_val.json` / `trt_clip_val.json` keep working)."""
import json
import logging
import os
import sys
import time

import numpy as np

from .graph import ONNXGraph, load_graph, simplify  # noqa: F401  (re-exported, as in the reference)
from .platform_settings import platform_setting_table

logger = logging.getLogger("dipoorlet")


def setup_logger(args):
    """File + stdout handlers on rank 0 (dipoorlet/utils.py:253-270, without termcolor)."""
    fmt = logging.Formatter("[%(asctime)s %(name)s] (%(filename)s %(lineno)d): %(levelname)s %(message)s",
                            datefmt="%Y-%m-%d %H:%M:%S")
    logger.setLevel(logging.INFO)
    path = os.path.join(args.output_dir, "log-{}.txt".format(time.strftime("%Y-%m-%d-%H:%M:%S")))
    with open(path, "w") as f:
        f.write(str(args) + "\n")
    for handler in (logging.FileHandler(path), logging.StreamHandler(sys.stdout)):
        handler.setLevel(logging.INFO)
        handler.setFormatter(fmt)
        logger.addHandler(handler)


class _Registry:
    """`dispatch_functool` (dipoorlet/utils.py:281-303): call as
    `dispatcher(key, *args, **kw)`; unknown keys fall back to the decorated default."""

    def __init__(self, default):
        self._default = default
        self.registry = {}
        self.__name__ = getattr(default, "__name__", "dispatcher")
        self.__doc__ = default.__doc__

    def register(self, key, func=None):
        if func is None:
            return lambda f: self.register(key, f)
        self.registry[key] = func
        return func

    def dispatch(self, key):
        return self.registry.get(key, self._default)

    def __call__(self, *args, **kw):
        return self.dispatch(args[0])(*args[1:], **kw)


def dispatch_functool(func):
    return _Registry(func)


def cos_similarity(ta, tb):
    """dipoorlet/utils.py:273-278 for host arrays (the device path uses K7b)."""
    assert ta.shape == tb.shape
    dot = np.sum(ta * tb)
    if dot == 0:
        return 0.
    return dot / np.sqrt(np.square(ta).sum()) / np.sqrt(np.square(tb).sum())


def update_model_path(name, args):
    args.model = os.path.join(args.output_dir, f"{name}.onnx")


# ---- clip-value files --------------------------------------------------------------
def _as_jsonable(v):
    return v.tolist() if hasattr(v, "tolist") else v


def save_clip_val(act_clip_val, weight_clip_val, args, act_fname="act_clip_val.json",
                  weight_fname="weight_clip_val.json"):
    """utils.py:313-323. Like the reference this converts the dict entries in place."""
    for table in (act_clip_val, weight_clip_val):
        for k in table:
            table[k][0] = _as_jsonable(table[k][0])
            table[k][1] = _as_jsonable(table[k][1])
    os.makedirs(args.output_dir, exist_ok=True)
    for table, fname in ((act_clip_val, act_fname), (weight_clip_val, weight_fname)):
        with open(os.path.join(args.output_dir, fname), "w") as f:
            json.dump(table, f, indent=4)


def load_clip_val(args, act_fname="act_clip_val.json", weight_fname="weight_clip_val.json"):
    """utils.py:348-368: activations come back as np.float64 scalars, weights as arrays
    (collapsed to scalars when the platform quantises weights per layer)."""
    with open(os.path.join(args.output_dir, act_fname)) as f:
        act = json.load(f)
    for k in act:
        act[k] = [np.float64(act[k][0]), np.float64(act[k][1])]
    per_channel = platform_setting_table[args.deploy]["qw_params"].get("per_channel", False)
    with open(os.path.join(args.output_dir, weight_fname)) as f:
        weight = json.load(f)
    for k in weight:
        lo, hi = np.array(weight[k][0]), np.array(weight[k][1])
        weight[k] = [lo, hi] if per_channel else [np.float64(lo), np.float64(hi)]
    return act, weight


def reduce_clip_val(rank_size, args, act_fname="act_clip_val.json",
                    weight_fname="weight_clip_val.json"):
    """utils.py:326-345, kept for drop-in use of the per-rank files: min/max for
    'minmax', mean of the per-rank clips otherwise. The B200 path does NOT use it for
    hist/mse — statistics are all-reduced on device so the result is world-size
    invariant — but the files written stay compatible."""
    act, weight = load_clip_val(args, act_fname + ".rank0", weight_fname + ".rank0")
    mean = args.act_quant != "minmax"
    if mean:
        for k in act:
            act[k] = [act[k][0] / float(rank_size), act[k][1] / float(rank_size)]
    for r in range(1, rank_size):
        with open(os.path.join(args.output_dir, f"{act_fname}.rank{r}")) as f:
            other = json.load(f)
        for k, v in other.items():
            if mean:
                act[k][0] += v[0] / float(rank_size)
                act[k][1] += v[1] / float(rank_size)
            else:
                act[k] = [np.array(min(v[0], act[k][0])), np.array(max(v[1], act[k][1]))]
    save_clip_val(act, weight, args)


# ---- profiling files -------------------------------------------------------------
def save_profiling_res(layer_cosine_dict, model_cosine_dict, args, rank=0,
                       layer_res_fname="layer_res.json", model_res_fname="model_res.json"):
    layer = {k: float(v) for k, v in layer_cosine_dict.items()}
    model = {k: [float(v[0]), float(v[1])] for k, v in model_cosine_dict.items()}
    if layer:
        with open(os.path.join(args.output_dir, f"{layer_res_fname}.rank{rank}"), "w") as f:
            json.dump(layer, f, indent=4)
    with open(os.path.join(args.output_dir, f"{model_res_fname}.rank{rank}"), "w") as f:
        json.dump(model, f, indent=4)


def reduce_profiling_res(rank_size, args, layer_res_fname="layer_res.json",
                         model_res_fname="model_res.json"):
    """utils.py:386-412: mean of per-rank layer cosines, mean/min of model cosines."""
    layer = {}
    if args.model_type is None:
        for r in range(rank_size):
            with open(os.path.join(args.output_dir, f"{layer_res_fname}.rank{r}")) as f:
                for k, v in json.load(f).items():
                    layer[k] = layer.get(k, 0.) + v / float(rank_size)
    model = {}
    for r in range(rank_size):
        with open(os.path.join(args.output_dir, f"{model_res_fname}.rank{r}")) as f:
            for k, v in json.load(f).items():
                if k not in model:
                    model[k] = [v[0] / float(rank_size), v[1]]
                else:
                    model[k][0] += v[0] / float(rank_size)
                    model[k][1] = min(model[k][1], v[1])
    return layer, model
