"""Pin the oracle (oracle/mapf_oracle.py and oracle/mapf_oracle.c) to the reference: every fixture in
tests/golden/ was produced by the unmodified reference (oracle/make_golden.py)."""
import numpy as np
import pytest

import golden_util as G
from oracle import c_oracle, mapf_oracle

ROWS = G.names("rows_")
STEPS = G.names("steps_")
MOVES = G.names("moves_")


def c_env(spec):
    return c_oracle.COracle(spec["rows"], spec["n_agents"], spec["goals"], spec["fail_prob"], spec["r_clash"],
                            spec["r_goal"], spec["r_living"], spec["soc"])


def py_env(spec):
    return mapf_oracle.OracleSpec(spec["rows"], spec["n_agents"], spec["starts"], spec["goals"], spec["fail_prob"],
                                  spec["r_clash"], spec["r_goal"], spec["r_living"], spec["soc"])


@pytest.mark.parametrize("name", ROWS)
def test_c_oracle_rows(name):
    spec, d = G.load(name)
    env = c_env(spec)
    assert env.L == spec["L"]
    got = env.rows(d["state_lo"], d["state_hi"], d["action"])
    assert np.array_equal(got["row_ptr"], d["row_ptr"])
    assert np.array_equal(got["next_lo"], d["next_lo"]) and np.array_equal(got["next_hi"], d["next_hi"])
    assert np.array_equal(G.f64_to_bits(got["prob"]), d["prob_bits"])
    assert np.array_equal(G.f64_to_bits(got["reward"]), d["reward_bits"])
    assert np.array_equal(got["done"], d["done"]) and np.array_equal(got["collision"], d["collision"])


@pytest.mark.parametrize("name", ROWS)
def test_py_oracle_rows(name):
    spec, d = G.load(name)
    env = py_env(spec)
    assert env.L == spec["L"] and str(env.nS) == spec["nS"] and str(env.s0) == spec["s0"]
    B = len(d["action"])
    # the pure-Python port is slow: check a spread of rows per fixture, all records of each
    pick = range(B) if B <= 400 else sorted(set(np.random.default_rng(0).integers(0, B, 400).tolist()))
    budget = 60000
    for b in pick:
        lo, hi = int(d["row_ptr"][b]), int(d["row_ptr"][b + 1])
        if hi - lo > budget:
            continue
        budget -= hi - lo
        row = env.row(G.big(d["state_lo"][b], d["state_hi"][b]), int(d["action"][b]))
        assert len(row) == hi - lo
        for j, (p, coll, ns, r, done) in enumerate(row):
            i = lo + j
            assert ns == G.big(d["next_lo"][i], d["next_hi"][i])
            assert G.py_bits(p) == int(d["prob_bits"][i]) and G.py_bits(r) == int(d["reward_bits"][i])
            assert int(done) == int(d["done"][i]) and int(coll) == int(d["collision"][i])
        if budget <= 0:
            break


@pytest.mark.parametrize("name", G.names("full_"))
def test_c_oracle_full_table(name):
    spec, d = G.load(name)
    env = c_env(spec)
    cs = env.table_checksums(0, int(spec["nS"]))
    for key, val in cs.items():
        assert val == int(d[key]), key
    # SURVEY 8c known answers for C1
    assert cs["count"] == 669808 and cs["n_collision"] == 10104 and cs["n_done"] == 11894
    assert cs["sum_next_lo"] == 1371237816 and cs["ordered"] == 0x000226d3bff8a161
    nS, nA = int(spec["nS"]), spec["nA"]
    s = np.repeat(np.arange(nS, dtype=np.uint64), nA)
    a = np.tile(np.arange(nA, dtype=np.int64), nS)
    got = env.rows(s, np.zeros_like(s), a)
    assert np.array_equal(np.diff(got["row_ptr"]).astype(np.uint8), d["row_len"])
    again = G.checksums(got["next_lo"], got["next_hi"], G.f64_to_bits(got["prob"]), G.f64_to_bits(got["reward"]),
                        got["done"], got["collision"])
    assert again == cs


@pytest.mark.parametrize("name", STEPS)
def test_c_oracle_steps(name):
    spec, d = G.load(name)
    env = c_env(spec)
    got = env.step(d["state_lo"], d["state_hi"], d["action"], G.bits_to_f64(d["uniform_bits"]))
    assert np.array_equal(got["next_lo"], d["next_lo"]) and np.array_equal(got["next_hi"], d["next_hi"])
    assert np.array_equal(G.f64_to_bits(got["reward"]), d["reward_bits"])
    assert np.array_equal(G.f64_to_bits(got["prob"]), d["prob_bits"])
    assert np.array_equal(got["done"], d["done"]) and np.array_equal(got["collision"], d["collision"])
    assert np.array_equal(got["terminal"], d["terminal"])


@pytest.mark.parametrize("name", STEPS)
def test_py_oracle_steps(name):
    spec, d = G.load(name)
    env = py_env(spec)
    u = G.bits_to_f64(d["uniform_bits"])
    T = min(len(d["action"]), 1500)
    for t in range(T):
        ns, r, done, p, coll, used = env.step(G.big(d["state_lo"][t], d["state_hi"][t]), int(d["action"][t]),
                                              [float(x) for x in u[t]])
        assert ns == G.big(d["next_lo"][t], d["next_hi"][t])
        assert G.py_bits(r) == int(d["reward_bits"][t]) and G.py_bits(p) == int(d["prob_bits"][t])
        assert int(done) == int(d["done"][t]) and int(bool(coll)) == int(d["collision"][t])
        assert (used == 0) == bool(d["terminal"][t])


@pytest.mark.parametrize("name", MOVES)
def test_oracle_moves(name):
    spec, d = G.load(name)
    env = c_env(spec)
    k, dest, prob = env.moves()
    assert np.array_equal(k, d["k"]) and np.array_equal(dest, d["dest"])
    assert np.array_equal(G.f64_to_bits(prob), d["prob_bits"])
    pe = py_env(spec)
    assert np.array_equal(np.array(pe.cells, dtype=np.int32), d["cells"])
    rng = np.random.default_rng(1)
    for cell in rng.integers(0, pe.L, 300):
        for a in range(5):
            outs = pe.agent_outcomes(int(cell), a)
            assert len(outs) == d["k"][cell, a]
            for j, (nxt, p) in enumerate(outs):
                assert nxt == d["dest"][cell, a, j] and G.py_bits(p) == int(d["prob_bits"][cell, a, j])


def test_oracle_misc():
    m = G.misc()
    for e in m["encodings"]:
        assert mapf_oracle.from_digits(e["digits"], e["radix"]) == int(e["value"])
        assert mapf_oracle.to_digits(int(e["value"]), e["radix"], e["n"]) == e["digits"]
    o = m["obst4_spec"]
    env = mapf_oracle.OracleSpec(o["rows"], o["n_agents"], o["starts"], o["goals"], 0.2, -1000.0, 100.0, -1.0, False)
    for p in m["predecessors"]:
        if p["case"] == "obst4":
            assert sorted(str(x) for x in env.predecessors(int(p["s"]))) == p["pred"]


def test_fp_constants_known_answers():
    """the fp64 constants at fail_prob 0.2, in the reference's evaluation order."""
    rf = 0.2 / 2
    p = 1 - rf - rf
    assert p.hex() == "0x1.999999999999ap-1" and rf.hex() == "0x1.999999999999ap-4"
    assert (p + rf).hex() == "0x1.ccccccccccccdp-1" and (p + rf) + rf == 1.0
    assert (p * p).hex() == "0x1.47ae147ae147cp-1"


# ---- rows built after the hot path: the oracle against reference-generated fixtures -------------
@pytest.mark.parametrize("name", G.names("backup_"))
def test_oracle_backup(name):
    spec, d = G.load(name)
    ora = c_env(spec)
    V = G.bits_to_f64(d["v_bits"])
    gamma = float(G.bits_to_f64(np.array([d["gamma_bits"]]))[0])
    q = ora.backup(d["state_lo"], d["state_hi"], d["action"], V, gamma)
    assert np.array_equal(G.f64_to_bits(q), d["q_bits"])
    # the pure-Python restatement on a sample
    py = py_env(spec)
    for i in range(0, len(d["action"]), max(1, len(d["action"]) // 150)):
        got = py.backup(int(d["state_lo"][i]), int(d["action"][i]), V.tolist(), gamma)
        assert G.py_bits(got) == int(d["q_bits"][i])


@pytest.mark.parametrize("name", G.names("preds_"))
def test_oracle_predecessors(name):
    spec, d = G.load(name)
    ora = c_env(spec)
    got = ora.predecessors(d["state_lo"], d["state_hi"])
    assert np.array_equal(got["row_ptr"], d["row_ptr"])
    assert np.array_equal(got["pred_lo"], d["pred_lo"]) and np.array_equal(got["pred_hi"], d["pred_hi"])
    py = py_env(spec)
    for i in range(0, len(d["state_lo"]), max(1, len(d["state_lo"]) // 5)):
        if d["row_ptr"][i + 1] - d["row_ptr"][i] > 50000:
            continue
        want = sorted(G.big(lo, hi) for lo, hi in zip(d["pred_lo"][d["row_ptr"][i]:d["row_ptr"][i + 1]],
                                                      d["pred_hi"][d["row_ptr"][i]:d["row_ptr"][i + 1]]))
        assert sorted(py.predecessors(G.big(d["state_lo"][i], d["state_hi"][i]))) == want


@pytest.mark.parametrize("name", G.names("project_"))
def test_oracle_projection(name):
    spec, d = G.load(name)
    ora = c_env(spec)
    py = py_env(spec)
    k = 0
    while "agents_%d" % k in d:
        agents = d["agents_%d" % k]
        lo, hi = ora.project(d["state_lo"], d["state_hi"], agents)
        assert np.array_equal(lo, d["proj_lo_%d" % k]) and np.array_equal(hi, d["proj_hi_%d" % k])
        assert py.project(G.big(d["state_lo"][3], d["state_hi"][3]), [int(a) for a in agents]) == \
            G.big(d["proj_lo_%d" % k][3], d["proj_hi_%d" % k][3])
        k += 1
    assert k >= 5


def test_np_random_stream_pinned():
    """`MapfEnv.np_random` (gym 0.13.0 `seeding.np_random(42)`, mapf_env.py:40,139): the product's restatement and the
    reference shim's build the same RandomState; the first draws are pinned as bit patterns."""
    from gym_mapf_b200.envs.mapf_env import GYM_MAPF_SEED, _gym_np_random
    from oracle import ref_shim
    a, seed_a = _gym_np_random(GYM_MAPF_SEED)
    b, seed_b = ref_shim._np_random(GYM_MAPF_SEED)
    assert GYM_MAPF_SEED == 42 and seed_a == seed_b == 42
    xa = [a.rand() for _ in range(4)]
    xb = [b.rand() for _ in range(4)]
    assert xa == xb
    assert [v.hex() for v in xa] == ["0x1.7f1f7113f5ff8p-2", "0x1.eff671fe363dcp-2", "0x1.d76f45e56b83cp-1",
                                     "0x1.ed831da02bffcp-2"]
    assert np.array_equal(a.get_state()[1], b.get_state()[1])
