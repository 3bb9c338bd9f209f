"""Import shim: the package directory is ``pl-viwo_b200/`` (not a valid Python identifier), so this module loads
it under the importable name ``plviwo_b200`` (and its ``synth`` / ``build`` sub-modules)."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pl-viwo_b200")


def _load(name, filename, submodule_search=None):
    spec = importlib.util.spec_from_file_location(name, os.path.join(_DIR, filename),
                                                  submodule_search_locations=submodule_search)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


_pkg = _load("plviwo_b200", "__init__.py", [_DIR])
synth = _load("plviwo_b200.synth", "synth.py")
build = _load("plviwo_b200.build", "build.py")
shard = _load("plviwo_b200.shard", "shard.py")
_pkg.synth = synth
_pkg.build = build
_pkg.shard = shard
