"""TEST INFRASTRUCTURE ONLY -- generate `tests/golden/*.npz` by running the UNMODIFIED reference.

    python oracle/make_golden.py            # build container only; needs /root/reference

Every array in the fixtures is an output of `/root/reference/gym_mapf` itself (imported through
`oracle/ref_shim.py`), never of the oracle restatement or of the CUDA path.  Floating-point values are stored as
their IEEE-754 bit patterns (uint64) so that the parity tests compare bits, not rounded decimals.  Joint-state
indices can exceed 64 bits (room-64-64-8 with 8 agents needs 94), so every state is stored as two uint64 limbs
(`*_lo`, `*_hi`).

Fixture families (see `tests/golden/README.md` for the per-file description):
  rows_*      P[s][a] rows: CSR (row_ptr) of (next_state, prob, reward, done, collision)      mapf_env.py:448-479
  full_*      checksums of a complete table (every s, every a)                                 mapf_env.py:448-479
  steps_*     step() traces with the uniforms fed to categorical_sample recorded              mapf_env.py:237-266
  moves_*     single_agent_movements for every (cell, action)                                  mapf_env.py:163-184
  misc        encodings, factory start states, predecessors                          envs/__init__.py:50-79 ...
"""
import json
import os
import struct
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.load_reference()
from gym_mapf.envs import (integer_to_vector, vector_to_integer, ACTIONS)  # noqa: E402
from gym_mapf.envs.grid import MapfGrid, EmptyCell  # noqa: E402
from gym_mapf.envs.mapf_env import MapfEnv, OptimizationCriteria  # noqa: E402
from gym_mapf.envs.utils import create_mapf_env  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
M64 = (1 << 64) - 1


def f64_bits(x):
    return struct.unpack("<Q", struct.pack("<d", float(x)))[0]


def limbs(values):
    lo = np.array([v & M64 for v in values], dtype=np.uint64)
    hi = np.array([(v >> 64) & M64 for v in values], dtype=np.uint64)
    return lo, hi


def spec_json(env, criterion):
    h, w = len(env.grid), len(env.grid[0])
    rows = ["".join("." if env.grid[r, c] is EmptyCell else "@" for c in range(w)) for r in range(h)]
    return json.dumps({
        "rows": rows, "n_agents": env.n_agents,
        "starts": [list(x) for x in env.agents_starts], "goals": [list(x) for x in env.agents_goals],
        "fail_prob": env.fail_prob, "r_clash": env.reward_of_clash, "r_goal": env.reward_of_goal,
        "r_living": env.reward_of_living, "soc": criterion == OptimizationCriteria.SoC,
        "L": len(env.valid_locations), "nS": str(env.nS), "nA": env.nA, "s0": str(env.s)})


def rows_arrays(env, pairs):
    """Run env.P[s][a] for every (s, a) in `pairs`; return the CSR arrays."""
    row_ptr = [0]
    ns, pb, rb, dn, cl = [], [], [], [], []
    for s, a in pairs:
        for (prob, coll), nxt, reward, done in env.P[s][a]:
            ns.append(int(nxt))
            pb.append(f64_bits(prob))
            rb.append(f64_bits(reward))
            dn.append(1 if done else 0)
            cl.append(1 if coll else 0)
        row_ptr.append(len(ns))
    s_lo, s_hi = limbs([p[0] for p in pairs])
    n_lo, n_hi = limbs(ns)
    return dict(state_lo=s_lo, state_hi=s_hi, action=np.array([p[1] for p in pairs], dtype=np.int64),
                row_ptr=np.array(row_ptr, dtype=np.int64), next_lo=n_lo, next_hi=n_hi,
                prob_bits=np.array(pb, dtype=np.uint64), reward_bits=np.array(rb, dtype=np.uint64),
                done=np.array(dn, dtype=np.uint8), collision=np.array(cl, dtype=np.uint8))


def checksums(arr):
    """Whole-table checksums, all mod 2**64 (documented in tests/golden/README.md)."""
    count = int(arr["next_lo"].shape[0])
    idx = np.arange(1, count + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        tag = arr["next_lo"] + np.uint64(1) + np.uint64(2) * arr["collision"].astype(np.uint64) \
            + np.uint64(4) * arr["done"].astype(np.uint64)
        ordered = np.sum(idx * tag, dtype=np.uint64)
        return dict(count=np.uint64(count), n_collision=np.uint64(int(arr["collision"].sum())),
                    n_done=np.uint64(int(arr["done"].sum())),
                    sum_next_lo=np.sum(arr["next_lo"], dtype=np.uint64),
                    sum_next_hi=np.sum(arr["next_hi"], dtype=np.uint64),
                    sum_prob_bits=np.sum(arr["prob_bits"], dtype=np.uint64),
                    sum_reward_bits=np.sum(arr["reward_bits"], dtype=np.uint64), ordered=ordered)


def save(name, spec, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, spec=np.array(spec), **arrays)
    print("%-28s %8.1f KiB" % (name, os.path.getsize(path) / 1024.0), flush=True)


def random_state(env, rng, window=None, allow_dup=True):
    """A joint state: every agent on a uniformly random free cell, optionally inside a w x w window (so that
    agents interact) -- duplicates (already-clashed, hence terminal, states) are allowed."""
    cells = env.valid_locations
    if window is not None:
        h, w = len(env.grid), len(env.grid[0])
        for _ in range(200):
            r0 = int(rng.integers(0, max(1, h - window + 1)))
            c0 = int(rng.integers(0, max(1, w - window + 1)))
            sub = [rc for rc in cells if r0 <= rc[0] < r0 + window and c0 <= rc[1] < c0 + window]
            if len(sub) >= 2:
                cells = sub
                break
    picks = [cells[int(rng.integers(0, len(cells)))] for _ in range(env.n_agents)]
    if not allow_dup and len(set(picks)) != len(picks):
        return random_state(env, rng, window, allow_dup)
    return env.locations_to_state(tuple(picks))


def near_goal_state(env, rng):
    """Most agents already on their goal (exercises the goal reward and the SoC parked-agent discount)."""
    locs = list(env.agents_goals)
    k = int(rng.integers(0, env.n_agents))
    r, c = locs[k]
    for dr, dc in ((0, 1), (1, 0), (0, -1), (-1, 0)):
        if (r + dr, c + dc) in env.loc_to_int:
            locs[k] = (r + dr, c + dc)
            break
    return env.locations_to_state(tuple(locs))


def mixed_pairs(env, rng, n_uniform, n_dense, n_goal, window=3):
    pairs = []
    for _ in range(n_uniform):
        pairs.append((random_state(env, rng), int(rng.integers(0, env.nA))))
    for _ in range(n_dense):
        pairs.append((random_state(env, rng, window=window), int(rng.integers(0, env.nA))))
    for _ in range(n_goal):
        s = near_goal_state(env, rng)
        a = int(rng.integers(0, env.nA))
        if rng.random() < 0.5:  # bias towards STAY digits so SoC parks agents
            digits = [0 if rng.random() < 0.7 else int(rng.integers(0, 5)) for _ in range(env.n_agents)]
            a = sum(d * 5 ** i for i, d in enumerate(digits))
        pairs.append((s, a))
    return pairs


def steps_arrays(env, rng, n_steps, reset_prob=0.0):
    """A step() trace.  The env's RandomState is replaced by a tape of recorded uniforms so the exact
    draws `categorical_sample` consumed are part of the fixture (mapf_env.py:253-257)."""
    n = env.n_agents
    pre, act, uni, nxt, rb, pb, dn, cl, term = [], [], [], [], [], [], [], [], []
    env.reset()
    for _ in range(n_steps):
        a = int(rng.integers(0, env.nA))
        u = rng.random(n)
        tape = ref_shim.UniformTape(u)
        env.np_random = tape
        s_before = env.s
        s_after, reward, done, info = env.step(a)
        pre.append(int(s_before))
        act.append(a)
        uni.append([f64_bits(x) for x in u])
        nxt.append(int(s_after))
        rb.append(f64_bits(reward))
        pb.append(f64_bits(info["prob"]))
        dn.append(1 if done else 0)
        cl.append(1 if info.get("collision", False) else 0)
        term.append(1 if tape.pos == 0 else 0)  # terminal no-op: no draw consumed (mapf_env.py:238-240)
        assert tape.pos in (0, n)
        if done and (term[-1] or rng.random() < 0.9):
            env.reset()
        elif reset_prob and rng.random() < reset_prob:
            env.s = random_state(env, rng, window=3)
    p_lo, p_hi = limbs(pre)
    n_lo, n_hi = limbs(nxt)
    return dict(state_lo=p_lo, state_hi=p_hi, action=np.array(act, dtype=np.int64),
                uniform_bits=np.array(uni, dtype=np.uint64).reshape(n_steps, n), next_lo=n_lo, next_hi=n_hi,
                reward_bits=np.array(rb, dtype=np.uint64), prob_bits=np.array(pb, dtype=np.uint64),
                done=np.array(dn, dtype=np.uint8), collision=np.array(cl, dtype=np.uint8),
                terminal=np.array(term, dtype=np.uint8))


def moves_arrays(env):
    """single_agent_movements for every (cell, action): up to 3 (next_cell, prob) slots, -1 padded."""
    L = len(env.valid_locations)
    k = np.zeros((L, 5), dtype=np.uint8)
    dest = np.full((L, 5, 3), -1, dtype=np.int32)
    pb = np.zeros((L, 5, 3), dtype=np.uint64)
    for cell in range(L):
        for a in range(5):
            mv = env.single_agent_movements(cell, a)
            k[cell, a] = len(mv)
            for j, (src, nxt, p) in enumerate(mv):
                assert src == cell
                dest[cell, a, j] = nxt
                pb[cell, a, j] = f64_bits(p)
    cells = np.array(env.valid_locations, dtype=np.int32)
    return dict(k=k, dest=dest, prob_bits=pb, cells=cells)


CUSTOM_GRIDS = {
    "obst4": ["..@.", ".@..", "....", "@..@"],
    "corridor": ["....."],
    "column": [".", ".", "."],
    "pair": [".."],
    "ring": ["...", ".@.", "..."],
    "wide": ["..@...", "......"],
}


def main():
    os.makedirs(OUT, exist_ok=True)
    t0 = time.time()
    SoC, Mk = OptimizationCriteria.SoC, OptimizationCriteria.Makespan

    # ---- C1: empty-8-8 scen 1, 2 agents: the complete table, both criteria (SURVEY 8c known answers)
    for crit, tag in ((Mk, "makespan"), (SoC, "soc")):
        env = create_mapf_env("empty-8-8", 1, 2, 0.2, -1000.0, 100.0, -1.0, crit)
        pairs = [(s, a) for s in range(env.nS) for a in range(env.nA)]
        arr = rows_arrays(env, pairs)
        cs = checksums(arr)
        row_len = np.diff(arr["row_ptr"]).astype(np.uint8)
        save("full_c1_" + tag, spec_json(env, crit), row_len=row_len, **cs)
        if crit == Mk:
            # keep every record of the first 512 states (12 800 rows) verbatim
            keep = rows_arrays(env, pairs[:512 * 25])
            save("rows_c1_makespan_head", spec_json(env, crit), **keep)
        else:
            rng = np.random.default_rng(11)
            sel = sorted(set(int(x) for x in rng.integers(0, len(pairs), 4000)))
            save("rows_c1_soc_sample", spec_json(env, crit), **rows_arrays(env, [pairs[i] for i in sel]))
        rng = np.random.default_rng(1)
        env2 = create_mapf_env("empty-8-8", 1, 2, 0.2, -1000.0, 100.0, -1.0, crit)
        save("steps_c1_" + tag, spec_json(env2, crit), **steps_arrays(env2, rng, 10000, reset_prob=0.02))

    # ---- C2: room-32-32-4 scen 1, 4 agents, SoC
    env = create_mapf_env("room-32-32-4", 1, 4, 0.2, -1000.0, 100.0, -1.0, SoC)
    rng = np.random.default_rng(2)
    save("rows_c2", spec_json(env, SoC), **rows_arrays(env, mixed_pairs(env, rng, 700, 500, 300)))
    env = create_mapf_env("room-32-32-4", 1, 4, 0.2, -1000.0, 100.0, -1.0, SoC)
    save("steps_c2", spec_json(env, SoC), **steps_arrays(env, np.random.default_rng(22), 6000, reset_prob=0.05))
    save("moves_room-32-32-4", spec_json(env, SoC), **moves_arrays(env))
    envm = create_mapf_env("room-32-32-4", 1, 4, 0.2, -1000.0, 100.0, -1.0, Mk)
    save("rows_c2_makespan", spec_json(envm, Mk), **rows_arrays(envm, mixed_pairs(envm, rng, 100, 150, 100)))

    # ---- C3: maze-32-32-4 scen 10, 6 agents, Makespan: random rows + a slab of consecutive (s, a)
    env = create_mapf_env("maze-32-32-4", 10, 6, 0.2, -1000.0, 100.0, -1.0, Mk)
    rng = np.random.default_rng(3)
    pairs = mixed_pairs(env, rng, 60, 80, 40)
    s_slab = env.nS // 8 * 3 + 12345
    pairs += [(s_slab + i // 400, (7000 + i) % env.nA) for i in range(800)]
    save("rows_c3", spec_json(env, Mk), **rows_arrays(env, pairs))

    # ---- C4: room-64-64-8 scen 1, 8 agents, Makespan: 94-bit states
    env = create_mapf_env("room-64-64-8", 1, 8, 0.2, -1000.0, 100.0, -1.0, Mk)
    rng = np.random.default_rng(4)
    save("rows_c4", spec_json(env, Mk), **rows_arrays(env, mixed_pairs(env, rng, 8, 10, 6)))
    env = create_mapf_env("room-64-64-8", 1, 8, 0.2, -1000.0, 100.0, -1.0, Mk)
    save("steps_c4", spec_json(env, Mk), **steps_arrays(env, np.random.default_rng(44), 3000, reset_prob=0.05))
    save("moves_room-64-64-8", spec_json(env, Mk), **moves_arrays(env))

    # ---- C5: empty-32-32 scen 1, n = 2..10, conflict-density windows
    for n, (nu, nd, ng) in {2: (150, 200, 80), 3: (100, 150, 60), 4: (60, 100, 40), 5: (30, 60, 20),
                            6: (15, 30, 10), 7: (8, 14, 6), 8: (3, 5, 2), 10: (1, 1, 1)}.items():
        crit = SoC if n % 2 else Mk
        env = create_mapf_env("empty-32-32", 1, n, 0.2, -1000.0, 100.0, -1.0, crit)
        rng = np.random.default_rng(50 + n)
        pairs = mixed_pairs(env, rng, nu, nd // 2, ng, window=4) + mixed_pairs(env, rng, 0, nd - nd // 2, 0, window=2)
        save("rows_c5_n%d" % n, spec_json(env, crit), **rows_arrays(env, pairs))

    # ---- Berlin_1_256 (largest shipped map; move table does not fit shared memory)
    env = create_mapf_env("Berlin_1_256", 2, 3, 0.2, -1000.0, 100.0, -1.0, SoC)
    rng = np.random.default_rng(6)
    save("rows_berlin", spec_json(env, SoC), **rows_arrays(env, mixed_pairs(env, rng, 150, 200, 60)))
    save("moves_Berlin_1_256", spec_json(env, SoC), **moves_arrays(env))

    # ---- hand-made grids: complete tables, several noise levels and reward types
    variants = [(0.2, -1000.0, 100.0, -1.0), (0.0, -1000.0, 100.0, -1), (1.0, -10.0, 7.5, -0.25),
                (0.1, -1000, 100, -1), (0.5, -3.0, 2.0, -0.1), (0.3, -1000.0, 100.0, -1.0)]
    grid_cases = [("obst4", 2, ((0, 0), (3, 2)), ((2, 3), (0, 1))),
                  ("obst4", 3, ((0, 0), (3, 2), (2, 0)), ((2, 3), (0, 1), (1, 2))),
                  ("corridor", 2, ((0, 0), (0, 4)), ((0, 4), (0, 0))),
                  ("column", 2, ((0, 0), (2, 0)), ((2, 0), (1, 0))),
                  ("pair", 2, ((0, 0), (0, 1)), ((0, 1), (0, 0))),
                  ("ring", 3, ((0, 0), (2, 2), (0, 2)), ((2, 2), (0, 0), (2, 0))),
                  ("ring", 1, ((0, 0),), ((2, 2),)),
                  ("wide", 4, ((0, 0), (1, 5), (0, 3), (1, 2)), ((1, 5), (0, 0), (1, 1), (0, 5)))]
    for gname, n, starts, goals in grid_cases:
        for vi, (fp, rc, rg, rl) in enumerate(variants):
            if n >= 4 and vi not in (0, 2):
                continue
            if n == 3 and vi not in (0, 1, 2, 3):
                continue
            for crit, tag in ((SoC, "soc"), (Mk, "mk")):
                env = MapfEnv(MapfGrid(CUSTOM_GRIDS[gname]), n, starts, goals, fp, rc, rg, rl, crit)
                pairs = [(s, a) for s in range(env.nS) for a in range(env.nA)]
                if len(pairs) > 40000:
                    rng = np.random.default_rng(70 + vi)
                    sel = sorted(set(int(x) for x in rng.integers(0, len(pairs), 12000)))
                    pairs = [pairs[i] for i in sel]
                arr = rows_arrays(env, pairs)
                save("rows_%s_n%d_v%d_%s" % (gname, n, vi, tag), spec_json(env, crit), **arr)
        env = MapfEnv(MapfGrid(CUSTOM_GRIDS[gname]), n, starts, goals, 0.2, -1000.0, 100.0, -1.0, SoC)
        save("steps_%s_n%d" % (gname, n), spec_json(env, SoC),
             **steps_arrays(env, np.random.default_rng(80 + n), 3000, reset_prob=0.1))
        env = MapfEnv(MapfGrid(CUSTOM_GRIDS[gname]), n, starts, goals, 1.0, -10.0, 7.5, -0.25, Mk)
        save("steps_%s_n%d_fp1" % (gname, n), spec_json(env, Mk),
             **steps_arrays(env, np.random.default_rng(90 + n), 1500, reset_prob=0.1))

    # ---- misc: encodings, factory start states, predecessors
    rng = np.random.default_rng(9)
    enc = []
    for _ in range(300):
        n = int(rng.integers(1, 13))
        radix = int(rng.integers(1, 50000))
        digits = [int(rng.integers(0, radix)) for _ in range(n)]
        x = vector_to_integer(tuple(digits), [radix] * n, lambda v: v)
        back = integer_to_vector(x, [radix] * n, n, lambda v: v)
        assert list(back) == digits
        enc.append({"n": n, "radix": radix, "digits": digits, "value": str(x)})
    factory = []
    for mname, scen, n in [("empty-8-8", 1, 2), ("empty-48-48", 16, 2), ("room-32-32-4", 1, 4),
                           ("maze-32-32-4", 10, 6), ("maze-32-32-2", 1, 6), ("room-64-64-8", 1, 8),
                           ("empty-32-32", 1, 10), ("empty-16-16", 7, 5), ("room-64-64-16", 1, 3),
                           ("maze-128-128-2", 1, 2), ("maze-128-128-10", 1, 3), ("Berlin_1_256", 2, 3), ("Berlin_1_256", 1, 3),
                           ("sanity-2-8", None, 4), ("sanity-3-8", None, 5), ("empty-8-8", 3, 1000)]:
        try:
            env = create_mapf_env(mname, scen, n, 0.2, -1000.0, 100.0, -1.0, SoC)
            factory.append({"map": mname, "scen": scen, "n_req": n, "n": env.n_agents, "L": len(env.valid_locations),
                            "starts": [list(x) for x in env.agents_starts],
                            "goals": [list(x) for x in env.agents_goals], "s0": str(env.s), "nS": str(env.nS),
                            "goal_state": str(env.locations_to_state(env.agents_goals)),
                            "H": len(env.grid), "W": len(env.grid[0])})
        except Exception as e:  # noqa: BLE001 - the exception type is the fixture
            factory.append({"map": mname, "scen": scen, "n_req": n, "error": type(e).__name__})
    for mname, scen, n in [("room-32-32-4", 2, 4), ("maze-32-32-4", 1, 6)]:
        try:
            create_mapf_env(mname, scen, n, 0.2, -1000.0, 100.0, -1.0, SoC)
            factory.append({"map": mname, "scen": scen, "n_req": n, "error": None})
        except Exception as e:  # noqa: BLE001
            factory.append({"map": mname, "scen": scen, "n_req": n, "error": type(e).__name__})
    preds = []
    env = create_mapf_env("empty-8-8", 1, 2, 0.2, -1000.0, 100.0, -1.0, Mk)
    for s in [0, 1856, 3393, 4095, 777]:
        preds.append({"case": "c1", "s": str(s), "pred": sorted(str(x) for x in env.predecessors(s))})
    env = MapfEnv(MapfGrid(CUSTOM_GRIDS["obst4"]), 2, ((0, 0), (3, 2)), ((2, 3), (0, 1)), 0.2, -1000.0, 100.0, -1.0, Mk)
    for s in range(0, env.nS, 7):
        preds.append({"case": "obst4", "s": str(s), "pred": sorted(str(x) for x in env.predecessors(s))})
    with open(os.path.join(OUT, "misc.json"), "w") as f:
        json.dump({"encodings": enc, "factory": factory, "predecessors": preds,
                   "obst4_spec": {"rows": CUSTOM_GRIDS["obst4"], "n_agents": 2, "starts": [[0, 0], [3, 2]],
                                  "goals": [[2, 3], [0, 1]]}}, f)
    print("done in %.1f s" % (time.time() - t0))


if __name__ == "__main__":
    main()
