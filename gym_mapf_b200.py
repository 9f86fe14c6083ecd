"""Import shim: the package lives in the directory `gym-mapf_b200/`, whose name is not a Python identifier.
`import gym_mapf_b200` finds this file first and replaces itself in `sys.modules` with the real package."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gym-mapf_b200")
_spec = importlib.util.spec_from_file_location("gym_mapf_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["gym_mapf_b200"] = _mod
_spec.loader.exec_module(_mod)
