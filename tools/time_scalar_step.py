"""Where a scalar `env.step()` call goes (BASELINE configs[0]: the reference's own loop, one env at a time)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from gym_mapf_b200._native import _ptr, lib  # noqa: E402
from gym_mapf_b200.envs.mapf_env import OptimizationCriteria  # noqa: E402
from gym_mapf_b200.envs.utils import create_mapf_env  # noqa: E402

env = create_mapf_env("empty-8-8", 1, 2, 0.2, -1000.0, 100.0, -1.0, OptimizationCriteria.Makespan, device=0)
eng = env.engine
rng = np.random.default_rng(1)
acts = rng.integers(0, env.nA, 20000)
env.reset()
for a in acts[:200]:
    if env.step(int(a))[2]:
        env.reset()
t0 = time.perf_counter()
for a in acts:
    if env.step(int(a))[2]:
        env.reset()
t_env = (time.perf_counter() - t0) / len(acts)
# the C-ABI call alone on one env (numpy buffers)
s = np.array([env.s], np.int64); a = np.array([3], np.int32); u = np.array([[0.3, 0.6]], np.float64)
out = (np.zeros(1, np.int64), np.zeros(1), np.zeros(1), np.zeros(1, np.uint8), np.zeros(1, np.uint8))
for _ in range(200):
    eng.step_host(s, a, out, uniforms=u)
t0 = time.perf_counter()
for _ in range(20000):
    eng.step_host(s, a, out, uniforms=u)
t_host = (time.perf_counter() - t0) / 20000
L = lib(); h = eng._h
args = (h, _ptr(s), _ptr(a), 1, _ptr(u), 0, 0, 0, 0, _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _ptr(out[3]), _ptr(out[4]))
t0 = time.perf_counter()
for _ in range(20000):
    L.mapf_step_host(*args)
t_c = (time.perf_counter() - t0) / 20000
print("env.step() %.1f us   Engine.step_host(B=1) %.1f us   mapf_step_host through ctypes with prebuilt arguments %.1f us" % (
    t_env * 1e6, t_host * 1e6, t_c * 1e6))
