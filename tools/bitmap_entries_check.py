"""Parity of the "moves derived from the shared-memory obstacle bitmap" experiment build against the C oracle on the maps
whose move table does not fit shared memory (the shipped library reads the table through L2 there).

    make -C gym_mapf_b200/csrc variant TAG=bm4 VN=4 EXTRA=-DMAPF_BITMAP_ENTRIES
    MAPF_B200_LIB=$PWD/gym_mapf_b200/csrc/libmapf_b200_bm4.so python tools/bitmap_entries_check.py
    MAPF_B200_LIB=... python tools/bench_configs.py big_maze-128-128-10_step big_Berlin_1_256_step ...   # timing

Measured (profiles/r02_ablations.txt): bit-exact; step 40.4 us vs 12.5 us per 2**20 envs, expand 56-63 % vs 66 %."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_mapf_b200.envs.mapf_env import OptimizationCriteria
from gym_mapf_b200.envs.utils import create_mapf_env
from oracle import c_oracle
for name, scen in (("maze-128-128-10", 2), ("Berlin_1_256", 11), ("maze-128-128-2", 1)):
    try:
        env = create_mapf_env(name, scen, 4, 0.2, -1000.0, 100.0, -1.0, OptimizationCriteria.SoC, device=0)
    except Exception as e:
        print(name, "skip", repr(e)[:80]); continue
    eng = env.engine
    rows = ["".join("@" if v else "." for v in r) for r in env.grid.obstacles]
    ora = c_oracle.COracle(rows, env.n_agents, env.agents_goals, 0.2, -1000.0, 100.0, -1.0, True)
    rng = np.random.default_rng(5)
    B = 200000
    cells = rng.integers(0, eng.L, (B, eng.n)).astype(np.int32)
    cells[: B // 4] = rng.integers(0, 60, (B // 4, eng.n))     # dense: conflicts, goals nearby
    gl = np.array([env.loc_to_int[g] for g in env.agents_goals], dtype=np.int32)
    cells[B // 4: B // 4 + 1000] = gl                            # on the goals (parked bits)
    lo, hi = ora.encode(cells)
    a = rng.integers(0, eng.nA, B).astype(np.int64)
    a[B // 4: B // 4 + 500] = 0
    uni = rng.random((B, eng.n))
    want = ora.step(lo, hi, a, uni)
    st = eng.encode(torch.from_numpy(cells).cuda())
    ns, reward, prob, done, coll = eng.step(st, torch.from_numpy(a.astype(np.int32)).cuda(), uniforms=torch.from_numpy(uni).cuda())
    ok = (np.array_equal(ns.cpu().numpy().view(np.uint64), want["next_lo"]) and
          np.array_equal(reward.cpu().numpy().view(np.uint64), want["reward"].view(np.uint64)) and
          np.array_equal(prob.cpu().numpy().view(np.uint64), want["prob"].view(np.uint64)) and
          np.array_equal(done.cpu().numpy().astype(np.uint8), want["done"]) and
          np.array_equal(coll.cpu().numpy().astype(np.uint8), want["collision"]))
    m = 3000
    tr = eng.transitions(st[:m], torch.from_numpy(a[:m].astype(np.int32)).cuda())
    w = ora.rows(lo[:m], hi[:m], a[:m])
    ok2 = (np.array_equal(tr[0].cpu().numpy(), w["row_ptr"]) and np.array_equal(tr[1].cpu().numpy().view(np.uint64), w["next_lo"]) and
           np.array_equal(tr[2].cpu().numpy().view(np.uint64), w["prob"].view(np.uint64)) and
           np.array_equal(tr[3].cpu().numpy().view(np.uint64), w["reward"].view(np.uint64)))
    print(name, "L", eng.L, "moves_in_smem", eng.info.moves_in_smem if hasattr(eng, "info") else "?", "step parity", ok, "rows parity", ok2,
          "collisions", int(want["collision"].sum()), "done", int(want["done"].sum()))
