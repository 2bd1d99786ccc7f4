from ..platform_settings import platform_setting_table
from . import deploy_trt  # noqa: F401  (registers 'trt')
from .deploy_default import deploy_dispatcher


def to_deploy(graph, act_clip_val, weight_clip_val, args, **kwargs):
    """dipoorlet/deploy/deploy_base.py:13-19. Only the 'trt' writer is registered here —
    the other vendors' file formats are outside the B200 hot path."""
    clip_val = act_clip_val
    if platform_setting_table[args.deploy]['deploy_weight']:
        clip_val = dict(act_clip_val)
        clip_val.update(weight_clip_val)
    deploy_dispatcher(args.deploy, graph, clip_val, args, **kwargs)
