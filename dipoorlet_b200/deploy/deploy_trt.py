import json
import os

from .deploy_default import deploy_dispatcher


@deploy_dispatcher.register("trt")
def gen_trt_range(graph, clip_val, args, **kwargs):
    """trt_clip_val.json: blob_range[name] = max(-lo, hi) as a double, key order = blob
    order (dipoorlet/deploy/deploy_trt.py:7-16). Mutates clip_val like the reference."""
    for name in clip_val:
        lo, hi = clip_val[name]
        clip_val[name] = max(-lo.astype(float), hi.astype(float))
    os.makedirs(args.output_dir, exist_ok=True)
    with open(os.path.join(args.output_dir, 'trt_clip_val.json'), 'w') as f:
        json.dump({'blob_range': clip_val}, f, indent=4)
