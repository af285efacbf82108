"""Alias package: `import xlb` resolves to `xlb_b200` so that scripts written for Autodesk/XLB run unchanged.

Every `xlb.<sub>` import is redirected to the already-imported `xlb_b200.<sub>` module object (one copy of every class,
one DefaultConfig, one BC registry).
"""

import importlib
import importlib.abc
import importlib.util
import sys

import xlb_b200


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, target):
        self._target = target

    def create_module(self, spec):
        return importlib.import_module(self._target)

    def exec_module(self, module):
        pass


class _AliasFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if fullname == "xlb" or not fullname.startswith("xlb."):
            return None
        real = "xlb_b200." + fullname[len("xlb.") :]
        try:
            importlib.import_module(real)
        except ModuleNotFoundError as e:
            if e.name == real:
                return None
            raise
        return importlib.util.spec_from_loader(fullname, _AliasLoader(real), is_package=hasattr(sys.modules[real], "__path__"))


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())
from xlb_b200.compat import install as _install_standins

_install_standins()  # `warp` / `jax` names for reference scripts when the real packages are absent
sys.modules[__name__] = xlb_b200
