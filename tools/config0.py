"""BASELINE.json configs[0] -- the reference's own CPU-runnable case -- through the drop-in scalar API and through the
batched API: empty-8-8 scen 1, 2 agents, fail_prob 0.2: the full env.P table (4096 x 25 rows, 669 808 transitions)
and 10 000 random-policy steps."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from gym_mapf_b200.envs.mapf_env import OptimizationCriteria  # noqa: E402
from gym_mapf_b200.envs.utils import create_mapf_env  # noqa: E402
from gym_mapf_b200.envs.vec_env import VecMapfEnv  # noqa: E402

env = create_mapf_env("empty-8-8", 1, 2, 0.2, -1000.0, 100.0, -1.0, OptimizationCriteria.Makespan, device=0)
env.P[0][0]  # context creation, first launch
t0 = time.perf_counter()
n = 0
for s in range(env.nS):
    Ps = env.P[s]
    for a in range(env.nA):
        n += len(Ps[a])
t_table = time.perf_counter() - t0
rng = np.random.default_rng(1)
acts = rng.integers(0, env.nA, 10000)
env.reset()
t0 = time.perf_counter()
for a in acts:
    _, _, done, _ = env.step(int(a))
    if done:
        env.reset()
t_steps = time.perf_counter() - t0
# the same work through the batched API
vec = VecMapfEnv(env, 10000, seed=1)
torch.cuda.synchronize()
t0 = time.perf_counter()
tr = vec.build_table(0, env.nS)
cs = vec.checksum(tr)
t_vtable = time.perf_counter() - t0
t0 = time.perf_counter()
out = vec.step(torch.from_numpy(acts.astype(np.int32)).cuda())
torch.cuda.synchronize()
t_vstep = time.perf_counter() - t0
print(json.dumps({"case": "configs[0] empty-8-8 n=2", "transitions": n,
                  "scalar_api": {"table_s": t_table, "transitions_per_s": n / t_table, "steps_s": t_steps,
                                 "steps_per_s": 10000 / t_steps},
                  "batched_api": {"table_s": t_vtable, "count": cs["count"], "ordered_checksum": cs["ordered"],
                                  "ten_thousand_steps_s": t_vstep},
                  "reference_cpython_1core (SURVEY 6)": {"table_s": 4.45, "transitions_per_s": 1.5e5, "steps_per_s": 3.6e4}}))
