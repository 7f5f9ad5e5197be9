"""Loads the package directory `hydrograd.jl_b200/` (the dot makes it un-importable by name) as module
`hydrograd_jl_b200`."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
NAME = "hydrograd_jl_b200"


def load():
    if NAME in sys.modules:
        return sys.modules[NAME]
    d = os.path.join(ROOT, "hydrograd.jl_b200")
    spec = importlib.util.spec_from_file_location(NAME, os.path.join(d, "__init__.py"), submodule_search_locations=[d])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[NAME] = mod
    spec.loader.exec_module(mod)
    return mod
