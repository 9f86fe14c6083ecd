"""LIVE pin of the oracle: the Python and C restatements against the UNMODIFIED reference imported in this process
(`oracle/ref_shim.py`: `$MAPF_REFERENCE_ROOT`, `/root/reference`, or the `baseline/_ref` install).  Skipped when no
copy of the reference is present -- the committed fixtures (`tests/test_oracle_golden.py`) pin the oracle then.

CPU only: no `gpu` marker, nothing here touches the CUDA path.
"""
import struct

import numpy as np
import pytest

from oracle import c_oracle, mapf_oracle, ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(),
                                reason="no copy of the reference (MAPF_REFERENCE_ROOT, /root/reference, baseline/_ref)")

M64 = (1 << 64) - 1
CASES = [  # map, scen, agents, fail_prob, soc
    ("empty-8-8", 1, 2, 0.2, False),
    ("empty-8-8", 3, 3, 0.3, True),
    ("room-32-32-4", 1, 4, 0.2, True),
    ("maze-32-32-4", 10, 6, 0.2, False),
    ("room-64-64-8", 1, 8, 0.2, False),     # 94-bit states
    ("empty-16-16", 2, 5, 0.0, True),       # no noise: single-outcome rows
    ("empty-16-16", 2, 3, 1.0, True),       # the intended move never happens (probability 0 candidate dropped)
]


def bits(x):
    return struct.unpack("<Q", struct.pack("<d", float(x)))[0]


@pytest.fixture(scope="module")
def ref():
    ref_shim.load_reference()
    import gym_mapf.envs.mapf_env as me
    import gym_mapf.envs.utils as ut
    return me, ut


def make(ref, case):
    me, ut = ref
    name, scen, n, fp, soc = case
    crit = me.OptimizationCriteria.SoC if soc else me.OptimizationCriteria.Makespan
    env = ut.create_mapf_env(name, scen, n, fp, -1000.0, 100.0, -1.0, crit)
    spec = mapf_oracle.spec_from_reference_env(env, soc)
    cora = c_oracle.COracle(spec.rows, spec.n, spec.goals, fp, -1000.0, 100.0, -1.0, soc)
    return env, spec, cora


def sample_states(env, spec, rng, count):
    """Random states, states packed into a small window (conflicts), the start state and the goal state."""
    out = [env.s, env.locations_to_state(env.agents_goals)]
    for _ in range(count):
        out.append(mapf_oracle.from_digits([int(x) for x in rng.integers(0, spec.L, spec.n)], spec.L))
    k = min(spec.L, 6)
    for _ in range(count):
        out.append(mapf_oracle.from_digits([int(x) for x in rng.integers(0, k, spec.n)], spec.L))
    return out


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%s-s%d-n%d-f%g-%s" % (c[0], c[1], c[2], c[3], "soc" if c[4] else "mk"))
def test_rows_live(ref, case):
    env, spec, cora = make(ref, case)
    rng = np.random.default_rng(11)
    assert spec.L == len(env.valid_locations) and spec.nS == env.nS and spec.nA == env.nA and spec.s0 == env.s
    n_rows = 24 if spec.n <= 6 else 4
    states = sample_states(env, spec, rng, n_rows)
    pairs = [(s, int(rng.integers(0, spec.nA))) for s in states]
    lo = np.array([s & M64 for s, _ in pairs], np.uint64)
    hi = np.array([s >> 64 for s, _ in pairs], np.uint64)
    got_c = cora.rows(lo, hi, np.array([a for _, a in pairs], np.int64))
    at = 0
    for b, (s, a) in enumerate(pairs):
        want = env.P[s][a]
        mine = spec.row(s, a)
        assert len(mine) == len(want)
        assert int(got_c["row_ptr"][b + 1] - got_c["row_ptr"][b]) == len(want)
        for ((p, coll), ns, r, done), (p2, coll2, ns2, r2, done2) in zip(want, mine):
            assert ns == ns2 and bits(p) == bits(p2) and bits(r) == bits(r2)
            assert bool(done) == bool(done2) and bool(coll) == bool(coll2)
            assert int(got_c["next_lo"][at]) == ns & M64 and int(got_c["next_hi"][at]) == ns >> 64
            assert bits(got_c["prob"][at]) == bits(p) and bits(got_c["reward"][at]) == bits(r)
            assert int(got_c["done"][at]) == int(bool(done)) and int(got_c["collision"][at]) == int(bool(coll))
            at += 1


@pytest.mark.parametrize("case", CASES[:5], ids=lambda c: "%s-n%d" % (c[0], c[2]))
def test_step_stream_live(ref, case):
    """The reference stepping with ITS OWN seeded stream (gym 0.13.0 seeding, seed 42) against the oracle fed with
    uniforms drawn from an identically seeded generator: same states, rewards, flags and probabilities, draw for draw."""
    env, spec, cora = make(ref, case)
    twin, _ = ref_shim._np_random(42)
    rng = np.random.default_rng(5)
    s = env.reset()
    for t in range(300):
        a = int(rng.integers(0, spec.nA))
        terminal = env.is_terminal(env.state_to_locations(s))
        u = [] if terminal else [twin.rand() for _ in range(spec.n)]
        ns, r, done, info = env.step(a)
        ns2, r2, done2, p2, coll2, used = spec.step(s, a, u)
        assert used == len(u)
        assert ns == ns2 and bits(r) == bits(r2) and bool(done) == bool(done2) and bits(info["prob"]) == bits(p2)
        if not terminal:
            assert bool(info["collision"]) == bool(coll2)
            got = cora.step(np.array([s & M64], np.uint64), np.array([s >> 64], np.uint64), np.array([a], np.int64),
                            np.array([u], np.float64))
            assert int(got["next_lo"][0]) == ns & M64 and int(got["next_hi"][0]) == ns >> 64
            assert bits(got["reward"][0]) == bits(r) and bits(got["prob"][0]) == bits(info["prob"])
            assert int(got["done"][0]) == int(bool(done)) and int(got["collision"][0]) == int(bool(info["collision"]))
        s = env.reset() if done else ns


def test_moves_and_encodings_live(ref):
    env, spec, cora = make(ref, CASES[2])
    for cell in range(0, spec.L, 7):
        for a in range(5):
            want = env.single_agent_movements(cell, a)
            mine = spec.agent_outcomes(cell, a)
            assert [(w[1], bits(w[2])) for w in want] == [(m[0], bits(m[1])) for m in mine]
    rng = np.random.default_rng(3)
    for _ in range(50):
        ids = [int(x) for x in rng.integers(0, spec.L, spec.n)]
        s = mapf_oracle.from_digits(ids, spec.L)
        locs = env.state_to_locations(s)
        assert [env.loc_to_int[loc] for loc in locs] == ids and env.locations_to_state(locs) == s
        assert env.predecessors(s) == spec.predecessors(s)
