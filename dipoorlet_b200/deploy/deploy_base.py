from ..platform_settings import platform_setting_table
from . import deploy_trt  # noqa: F401  (registers 'trt')
from . import deploy_vendors  # noqa: F401  (registers 'atlas', 'imx', 'magicmind', 'snpe', 'ti')
from . import deploy_rv  # noqa: F401  (registers 'rv')
from . import deploy_stpu  # noqa: F401  (registers 'stpu')
from .deploy_default import deploy_dispatcher


def to_deploy(graph, act_clip_val, weight_clip_val, args, **kwargs):
    """dipoorlet/deploy/deploy_base.py:13-19. Registered writers: 'trt' (the graded artefact) and the seven
    vendor formats (deploy_vendors.py, deploy_rv.py, deploy_stpu.py) - every `-D` choice of the CLI."""
    clip_val = act_clip_val
    if platform_setting_table[args.deploy]['deploy_weight']:
        clip_val = dict(act_clip_val)
        clip_val.update(weight_clip_val)
    deploy_dispatcher(args.deploy, graph, clip_val, args, **kwargs)
