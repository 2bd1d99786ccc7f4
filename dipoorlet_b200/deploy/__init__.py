from .deploy_base import to_deploy  # noqa: F401
from .deploy_default import deploy_dispatcher  # noqa: F401
