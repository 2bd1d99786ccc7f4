"""Import-path drop-in: after `dipoorlet_b200.compat.install()`, code written against the reference's module
paths resolves to this package —

    from dipoorlet.tensor_cali.basic_algorithm import tensor_cali_dispatcher      # basic_algorithm.py:8
    from dipoorlet.tensor_cali import tensor_calibration, find_clip_val_minmax_weight
    from dipoorlet.weight_transform import weight_calibration
    from dipoorlet.deploy import to_deploy
    from dipoorlet.deploy.deploy_default import deploy_dispatcher

— so that third-party calibrator / deploy plugins (`@tensor_cali_dispatcher.register('name')`) and scripts that
call the entry points keep working unchanged. `dipoorlet.X` is the SAME module object as
`dipoorlet_b200.X` (one registry, not a copy). Nothing is aliased unless install() is called, and install()
refuses to shadow a real `dipoorlet` that is already imported."""
import importlib
import importlib.abc
import importlib.util
import sys

ALIAS, TARGET = "dipoorlet", "dipoorlet_b200"


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, real_name):
        self.real_name = real_name

    def create_module(self, spec):
        return importlib.import_module(self.real_name)      # the real module object itself

    def exec_module(self, module):
        pass                                                # already executed under its own name


class _AliasFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if fullname != ALIAS and not fullname.startswith(ALIAS + "."):
            return None
        real_name = TARGET + fullname[len(ALIAS):]
        try:
            real_spec = importlib.util.find_spec(real_name)
        except ModuleNotFoundError:
            return None
        if real_spec is None:
            return None
        spec = importlib.util.spec_from_loader(fullname, _AliasLoader(real_name),
                                               is_package=real_spec.submodule_search_locations is not None)
        return spec


def install():
    """Idempotent. Raises if a different `dipoorlet` package is already imported."""
    existing = sys.modules.get(ALIAS)
    target = importlib.import_module(TARGET)
    if existing is not None and existing is not target:
        raise RuntimeError("a different 'dipoorlet' package is already imported; refusing to shadow it")
    if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _AliasFinder())
    sys.modules[ALIAS] = target
    return target
