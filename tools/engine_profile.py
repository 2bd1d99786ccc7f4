"""Where the calibration forward's time goes: CUDA-event time per node of the batched engine
(ResNet-50, batch 64), grouped by operator kind. Usage: python tools/engine_profile.py [batch]"""
import collections
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dipoorlet_b200 import workloads as W  # noqa: E402
from dipoorlet_b200.engine import Engine  # noqa: E402
from dipoorlet_b200.graph import ONNXGraph  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
graph = ONNXGraph(W.build_resnet50(), "/tmp/engine_profile", "trt")
eng = Engine(graph, torch.device("cuda", 0))
x = torch.randn((batch, 3, 224, 224), device="cuda")
feeds = {graph.network_inputs[0]: x}
for _ in range(3):
    eng.run(feeds, want="all")
torch.cuda.synchronize()

records = []
orig = eng._exec


def timed_exec(node, env):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = orig(node, env)
    b.record()
    kind = node.op_type
    if kind == "Conv":
        w = eng._val(node.input[1], env)
        kind = "Conv %dx%d s%d%s" % (w.shape[2], w.shape[3], node.attrs.get("strides", [1])[0],
                                      " dw" if node.attrs.get("group", 1) > 1 else "")
    records.append((kind, a, b))
    return out


eng._exec = timed_exec
reps = 5
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(reps):
    eng.run(feeds, want="all")
t1.record()
torch.cuda.synchronize()
agg = collections.OrderedDict()
for kind, a, b in records:
    d = agg.setdefault(kind, [0, 0.0])
    d[0] += 1
    d[1] += a.elapsed_time(b)
total = t0.elapsed_time(t1) / reps
rows = [{"kind": k, "nodes": v[0] // reps, "ms": v[1] / reps, "share": v[1] / reps / total} for k, v in agg.items()]
rows.sort(key=lambda r: -r["ms"])
out = {"batch": batch, "forward_ms": total, "images_per_s_forward_only": batch / total * 1e3,
       "tensor_cores": eng.tensor_cores, "conv3x3_on_tensor_cores": eng.tc_conv3x3, "rows": rows}
for r in rows:
    print("%-14s nodes %3d  %8.3f ms  %5.1f %%" % (r["kind"], r["nodes"], r["ms"], 100 * r["share"]))
print("forward %.3f ms / batch of %d" % (total, batch))
os.makedirs("gpurun_out", exist_ok=True)
tag = os.environ.get("PROFILE_TAG", "")
json.dump(out, open("gpurun_out/engine_profile%s.json" % tag, "w"), indent=1)
