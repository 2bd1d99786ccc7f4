"""Registry of deploy writers: `deploy_dispatcher(platform, graph, clip_val, args, **kw)`.

Writers register with `@deploy_dispatcher.register("<platform>")`; an unknown platform falls
through to `_no_writer`, which only warns (same contract as the reference's registry,
dipoorlet/deploy/deploy_default.py:4-6)."""
from ..utils import _Registry, logger


def _no_writer(graph=None, clip_val=None, args=None, **kwargs):
    name = getattr(args, "deploy", "?")
    logger.warning("Deploy Platform Not Found! (no writer registered for %r)" % (name,))


deploy_dispatcher = _Registry(_no_writer)
