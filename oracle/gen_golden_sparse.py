"""Golden fixture for `--sparse`: runs the REFERENCE's own weight_calibration(sparse=True)
(/root/reference/dipoorlet, unmodified, under oracle/ref_shim; torch CPU) on the two small seeded models of
gen_golden.py, for both patterns, and writes tests/golden/<model>/wt_sparse_<pattern>.npz: the pruned and
quantised weights of every learnable layer.

    python oracle/gen_golden_sparse.py        # build container only; the fixtures are committed
"""
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

import torchvision  # noqa: E402,F401
import torch  # noqa: E402,F401

from oracle import ref_shim  # noqa: E402
from oracle.gen_golden import GOLD, N_IMG, ADA_BS, ADA_EPOCH  # noqa: E402


def main():
    from dipoorlet_b200 import onnx_lite as ol, workloads as W
    ref_shim.install()
    import dipoorlet.tensor_cali as RTC
    import dipoorlet.utils as RU
    from dipoorlet.weight_transform import weight_calibration
    logging.getLogger("dipoorlet").setLevel(logging.WARNING)
    for mname in ("tiny_r50", "tiny_mbv2"):
        out = os.path.join(GOLD, mname)
        model = ol.load(os.path.join(out, "model.onnx"))
        images = np.load(os.path.join(out, "images.npy"))
        for pattern, rate in (("unstruction", 0.5), ("nv24", 0.5), ("unstruction", 0.3)):
            tmp = tempfile.mkdtemp(prefix="dpl_gold_sp_")
            W.write_input_dir(images, os.path.join(tmp, "data"), "input")
            g = RU.ONNXGraph(ref_shim.from_lite(model), tmp, "trt", None)
            args = types.SimpleNamespace(
                input_dir=os.path.join(tmp, "data"), output_dir=tmp, data_num=N_IMG, world_size=1, rank=0,
                local_rank=0, act_quant="minmax", deploy="trt", bins=2048, threshold=0.99999,
                optim_transformer=False, skip_layers=[], bc=False, we=False, update_bn=False, adaround=False,
                brecq=False, drop=False, sparse=True, sparse_rate=rate, pattern=pattern, ada_bs=ADA_BS,
                ada_epoch=ADA_EPOCH, model=None, model_type=None, savefp=False, skip_prof_layer=False)
            act, w = RTC.tensor_calibration(g, args)
            graph, graph_ori, act2, w2 = weight_calibration(g, act, w, args)
            before = dict(model.graph.initializers)
            changed = {t.name: np.asarray(t.array) for t in graph.graph.initializer
                       if t.name not in before or not np.array_equal(before[t.name], np.asarray(t.array))}
            tag = pattern if rate == 0.5 else f"{pattern}_{int(rate * 100)}"
            np.savez_compressed(os.path.join(out, f"wt_sparse_{tag}.npz"), **changed)
            shutil.rmtree(tmp)
            zeros = sum(int((v == 0).sum()) for v in changed.values()) / sum(v.size for v in changed.values())
            print(mname, tag, "changed initializers:", len(changed), "zero fraction %.3f" % zeros)


if __name__ == "__main__":
    main()
