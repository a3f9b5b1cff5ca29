"""Import shim: makes the package directory `llama3.np_b200/` importable as `llama3_np_b200`.

A directory name containing a dot cannot be named in an `import` statement, so this module
loads `llama3.np_b200/__init__.py` as the package `llama3_np_b200` and replaces itself in
`sys.modules` with it.  After `import llama3_np_b200`, submodules import normally
(`from llama3_np_b200.llama3 import Llama`).
"""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg_dir = os.path.join(_here, "llama3.np_b200")
_spec = importlib.util.spec_from_file_location(
    "llama3_np_b200", os.path.join(_pkg_dir, "__init__.py"),
    submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["llama3_np_b200"] = _mod
_spec.loader.exec_module(_mod)
