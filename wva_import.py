"""Import helper: loads the hyphen-named package directory `inferno-autoscaler_b200/` as the
module `inferno_autoscaler_b200`."""
import importlib.util
import os
import sys

_NAME = "inferno_autoscaler_b200"
ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "inferno-autoscaler_b200")


def load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    spec = importlib.util.spec_from_file_location(_NAME, os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod
