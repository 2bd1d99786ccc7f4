"""Fixture for the graph IR boundary (dipoorlet/utils.py:22-250, ONNXGraph): the reference's own ONNXGraph
built over the two small seeded models -> tests/golden/<model>/graph_api.json (node names in order,
network inputs / outputs, initializer names, tensor shapes, producer and consumer maps).

    python oracle/gen_golden_graph_api.py      # build container only; the fixtures are committed
"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torchvision  # noqa: E402,F401
from oracle import ref_shim  # noqa: E402
from oracle.gen_golden import GOLD  # noqa: E402


def main():
    from dipoorlet_b200 import onnx_lite as ol
    ref_shim.install()
    import dipoorlet.utils as RU
    for mname in ("tiny_r50", "tiny_mbv2"):
        d = os.path.join(GOLD, mname)
        model = ol.load(os.path.join(d, "model.onnx"))
        g = RU.ONNXGraph(ref_shim.from_lite(model), tempfile.mkdtemp(), "trt", None)
        tensors = list(g.network_inputs) + [o for n in g.graph.node for o in n.output]
        out = {
            "nodes": [[n.op_type, n.name, list(n.input), list(n.output)] for n in g.graph.node],
            "network_inputs": list(g.network_inputs),
            "network_outputs": list(g.network_outputs),
            "initializers": sorted(g.initializer.keys()),
            "shapes": {t: [int(v) for v in g.get_tensor_shape(t)] for t in tensors},
            "producer": {t: (p if isinstance(p, str) else p.name) for t in tensors
                         for p in [g.get_tensor_producer(t)]},
            "consumer": {t: [(c if isinstance(c, str) else c.name) for c in g.get_tensor_consumer(t)]
                         for t in tensors},
        }
        json.dump(out, open(os.path.join(d, "graph_api.json"), "w"))
        print(mname, len(out["nodes"]), "nodes")


if __name__ == "__main__":
    main()
