"""CPU-only checks of the boundary: the C-ABI library loads and exports every symbol include/mapf_b200.h declares, the
ctypes binding covers all of them, a context cannot be created without a CUDA device (no CPU fallback), and the host
mirror of the reference API (encodings, grid, parsers, factory) reproduces the reference's known answers
(tests/golden/misc.json was written by the unmodified reference)."""
import ctypes
import os
import re

import numpy as np
import pytest

import golden_util as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mapf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mapf_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from gym_mapf_b200 import _native
    lib = _native.lib()
    names = declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert getattr(lib, name) is not None, name
    assert sorted(_native.EXPORTS) == names  # the binding neither misses nor invents an entry point
    assert lib.mapf_version().decode().startswith("mapf_b200")


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from gym_mapf_b200 import _native
    from gym_mapf_b200.envs.mapf_env import OptimizationCriteria
    from gym_mapf_b200.envs.utils import create_mapf_env
    env = create_mapf_env("empty-8-8", 1, 2, 0.2, -1000.0, 100.0, -1.0, OptimizationCriteria.Makespan)
    assert env.s == 1856 and env.nS == 4096 and env.nA == 25  # host-side set-up works without a device
    with pytest.raises(RuntimeError):
        env.P[env.s][0]
    with pytest.raises(RuntimeError):
        env.step(0)
    # the C ABI itself refuses as well
    obst = np.zeros((2, 2), np.uint8)
    rc = np.zeros(2, np.int32)
    spec = _native.MapfSpec(2, 2, obst.ctypes.data, 1, rc.ctypes.data, rc.ctypes.data, 0.2, -1.0, 1.0, -1.0, 0)
    h = ctypes.c_void_p()
    assert _native.lib().mapf_ctx_create(ctypes.byref(spec), 0, ctypes.byref(h)) == _native.MAPF_ERR_NO_DEVICE
    assert b"no CPU fallback" in _native.lib().mapf_last_error()


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "gym_mapf_b200")
    for base, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f)).read()
                assert "oracle" not in text.replace("the oracle", "").replace("C oracle", ""), os.path.join(base, f)


def test_host_encodings_and_factory_match_reference():
    from gym_mapf_b200.envs import integer_to_vector, vector_to_integer
    from gym_mapf_b200.envs.mapf_env import (OptimizationCriteria, integer_action_to_vector, vector_action_to_integer)
    from gym_mapf_b200.envs.utils import create_mapf_env
    misc = G.misc()
    for e in misc["encodings"]:
        assert vector_to_integer(tuple(e["digits"]), [e["radix"]] * e["n"], lambda v: v) == int(e["value"])
        assert list(integer_to_vector(int(e["value"]), [e["radix"]] * e["n"], e["n"], lambda v: v)) == e["digits"]
    assert vector_action_to_integer(("DOWN", "STAY", "UP")) == 28          # reference utils_tests.py:37-72
    assert integer_action_to_vector(28, 3) == ("DOWN", "STAY", "UP")
    for f in misc["factory"]:
        if f.get("error"):
            with pytest.raises(Exception) as ei:
                create_mapf_env(f["map"], f["scen"], f["n_req"], 0.2, -1000.0, 100.0, -1.0, OptimizationCriteria.SoC)
            assert type(ei.value).__name__ == f["error"]
        elif "s0" in f:
            env = create_mapf_env(f["map"], f["scen"], f["n_req"], 0.2, -1000.0, 100.0, -1.0, OptimizationCriteria.SoC)
            assert env.n_agents == f["n"] and len(env.valid_locations) == f["L"]
            assert env.s == int(f["s0"]) and env.nS == int(f["nS"])
            assert [list(x) for x in env.agents_starts] == f["starts"]
            assert env.locations_to_state(env.agents_goals) == int(f["goal_state"])


def test_host_predecessors_match_reference():
    from gym_mapf_b200.envs.grid import MapfGrid
    from gym_mapf_b200.envs.mapf_env import MapfEnv, OptimizationCriteria
    from gym_mapf_b200.envs.utils import create_mapf_env
    misc = G.misc()
    c1 = create_mapf_env("empty-8-8", 1, 2, 0.2, -1000.0, 100.0, -1.0, OptimizationCriteria.Makespan)
    o = misc["obst4_spec"]
    obst4 = MapfEnv(MapfGrid(o["rows"]), o["n_agents"], tuple(map(tuple, o["starts"])), tuple(map(tuple, o["goals"])),
                    0.2, -1000.0, 100.0, -1.0, OptimizationCriteria.Makespan)
    for p in misc["predecessors"]:
        env = c1 if p["case"] == "c1" else obst4
        assert sorted(str(x) for x in env.predecessors(int(p["s"]))) == p["pred"]
