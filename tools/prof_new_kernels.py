"""One launch each of the grouped step (4 specs) and the lane-per-agent step on 2**20 room-32-32-4 envs, for an ncu capture:

    ncu --set full --clock-control none --import-source on -k regex:'k_step_group|k_step_lanes' -c 2 \\
        -o gpurun_out/new_kernels python tools/prof_new_kernels.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from gym_mapf_b200 import _native  # noqa: E402
from gym_mapf_b200.envs.mapf_env import OptimizationCriteria  # noqa: E402
from gym_mapf_b200.envs.utils import create_mapf_env  # noqa: E402

B = 1 << 20
envs = [create_mapf_env("room-32-32-4", 1, 4, 0.2, -1000.0 - k, 100.0, -1.0, OptimizationCriteria.SoC, device=0) for k in range(4)]
eng = envs[0].engine
rng = np.random.default_rng(3)
states = eng.encode(torch.from_numpy(rng.integers(0, eng.L, (B, 4)).astype(np.int32)).cuda())
actions = torch.from_numpy(rng.integers(0, 625, B).astype(np.int32)).cuda()
grp = _native.Group([e.engine for e in envs], [B // 4] * 4)
for _ in range(3):  # warm-up launches (ncu -c 2 with --launch-skip 6 captures the last pair)
    grp.step(states, actions, seed=1)
    eng.step(states, actions, seed=1, mapping="lanes")
torch.cuda.synchronize()
print("ok")
