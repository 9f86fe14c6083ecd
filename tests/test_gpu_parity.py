"""GPU parity: the CUDA path (through the C ABI, include/mapf_b200.h) against the reference-generated golden
fixtures and against the C oracle on seeded inputs.  Integer fields and fp64 bit patterns must be identical."""
import os

import numpy as np
import pytest

import golden_util as G
from engine_util import (make_engine, make_oracle, states_tensor, split_states, u64, assert_rows_equal)

pytestmark = pytest.mark.gpu

ROWS = G.names("rows_")
STEPS = G.names("steps_")
MOVES = G.names("moves_")


@pytest.fixture(scope="module")
def torch():
    import torch as t
    return t


@pytest.mark.parametrize("name", ROWS)
def test_expand_matches_reference_rows(name, torch):
    spec, d = G.load(name)
    eng = make_engine(spec)
    assert eng.L == spec["L"] and str(eng.nS) == spec["nS"] and str(eng.s0) == spec["s0"] and eng.nA == spec["nA"]
    states = states_tensor(eng, d["state_lo"], d["state_hi"])
    actions = torch.from_numpy(d["action"].astype(np.int32)).to(eng.torch_device)
    got = eng.transitions(states, actions)
    assert_rows_equal(eng, got, d["row_ptr"], d["next_lo"], d["next_hi"], d["prob_bits"], d["reward_bits"], d["done"],
                      d["collision"])
    # checksum kernel agrees with the host definition
    cs = eng.checksum(got[1], got[2], got[3], got[4]).cpu().numpy().view(np.uint64)
    want = G.checksums(d["next_lo"], d["next_hi"], d["prob_bits"], d["reward_bits"], d["done"], d["collision"])
    keys = ["count", "n_collision", "n_done", "sum_next_lo", "sum_next_hi", "sum_prob_bits", "sum_reward_bits", "ordered"]
    assert {k: int(v) for k, v in zip(keys, cs)} == want


@pytest.mark.parametrize("name", G.names("full_"))
def test_full_table_c1(name, torch):
    spec, d = G.load(name)
    eng = make_engine(spec)
    nS = int(spec["nS"])
    row_ptr, ns, prob, reward, flags = eng.table_range(0, nS)
    assert np.array_equal(np.diff(row_ptr.cpu().numpy()).astype(np.uint8), d["row_len"])
    cs = eng.checksum(ns, prob, reward, flags).cpu().numpy().view(np.uint64)
    keys = ["count", "n_collision", "n_done", "sum_next_lo", "sum_next_hi", "sum_prob_bits", "sum_reward_bits", "ordered"]
    for k, v in zip(keys, cs):
        assert int(v) == int(d[k]), k
    assert int(cs[0]) == 669808
    # and record by record against the oracle
    ora = make_oracle(spec)
    s = np.repeat(np.arange(nS, dtype=np.uint64), spec["nA"])
    a = np.tile(np.arange(spec["nA"], dtype=np.int64), nS)
    w = ora.rows(s, np.zeros_like(s), a)
    assert_rows_equal(eng, (row_ptr, ns, prob, reward, flags), w["row_ptr"], w["next_lo"], w["next_hi"],
                      G.f64_to_bits(w["prob"]), G.f64_to_bits(w["reward"]), w["done"], w["collision"])
    # the slab split in two must concatenate to the same table
    half = nS // 2
    a_ptr, a_ns, *_ = eng.table_range(0, half)
    b_ptr, b_ns, *_ = eng.table_range(half, nS - half)
    assert int(a_ptr[-1]) + int(b_ptr[-1]) == 669808
    assert torch.equal(torch.cat([a_ns, b_ns]), ns)


@pytest.mark.parametrize("name", STEPS)
def test_step_matches_reference_traces(name, torch):
    spec, d = G.load(name)
    eng = make_engine(spec)
    states = states_tensor(eng, d["state_lo"], d["state_hi"])
    actions = torch.from_numpy(d["action"].astype(np.int32)).to(eng.torch_device)
    uniforms = torch.from_numpy(G.bits_to_f64(d["uniform_bits"]).copy()).to(eng.torch_device)
    ns, reward, prob, done, coll = eng.step(states, actions, uniforms=uniforms)
    lo, hi = split_states(eng, ns)
    assert np.array_equal(lo, d["next_lo"]) and np.array_equal(hi, d["next_hi"])
    assert np.array_equal(u64(reward), d["reward_bits"]) and np.array_equal(u64(prob), d["prob_bits"])
    assert np.array_equal(done.cpu().numpy().astype(np.uint8), d["done"])
    assert np.array_equal(coll.cpu().numpy().astype(np.uint8), d["collision"])
    # in place (next_states aliasing states) gives the same answer
    st2 = states.clone()
    out = (st2, torch.empty_like(reward), torch.empty_like(prob), torch.empty_like(done), torch.empty_like(coll))
    eng.step(st2, actions, uniforms=uniforms, out=out)
    assert torch.equal(st2, ns) and torch.equal(out[1], reward)


@pytest.mark.parametrize("name", MOVES)
def test_move_table_matches_reference(name):
    spec, d = G.load(name)
    eng = make_engine(spec)
    k, dest, prob, rc = eng.moves()
    assert np.array_equal(k, d["k"]) and np.array_equal(dest, d["dest"])
    assert np.array_equal(G.f64_to_bits(prob), d["prob_bits"])
    assert np.array_equal(rc, d["cells"])


@pytest.mark.parametrize("name", ["rows_c2", "rows_c4", "rows_c5_n10", "rows_berlin", "rows_obst4_n3_v0_soc"])
def test_encode_decode_roundtrip(name, torch):
    spec, _ = G.load(name)
    eng = make_engine(spec)
    ora = make_oracle(spec)
    rng = np.random.default_rng(5)
    B = 20000
    cells = rng.integers(0, eng.L, (B, eng.n)).astype(np.int32)
    lo, hi = ora.encode(cells)
    t = eng.encode(torch.from_numpy(cells).to(eng.torch_device))
    glo, ghi = split_states(eng, t)
    assert np.array_equal(glo, lo) and np.array_equal(ghi, hi)
    back = eng.decode(t).cpu().numpy()
    assert np.array_equal(back, cells)
    # extremes: state 0 and nS - 1
    ext = eng.states_from_ints([0, eng.nS - 1, eng.s0, eng.goal_state])
    dec = eng.decode(ext).cpu().numpy()
    assert np.all(dec[0] == 0) and np.all(dec[1] == eng.L - 1)
    assert eng.states_to_ints(eng.encode(torch.from_numpy(dec).to(eng.torch_device))) == [0, eng.nS - 1, eng.s0,
                                                                                         eng.goal_state]


@pytest.mark.parametrize("name,B", [("rows_c2", 1 << 20), ("rows_c4", 1 << 17), ("rows_c3", 1 << 18),
                                    ("rows_c5_n10", 1 << 15), ("rows_berlin", 1 << 17)])
def test_step_large_batch_vs_oracle(name, B, torch):
    """BASELINE.json sizes (C2: 2**20 envs): every env-step compared bit-exactly against the C oracle."""
    spec, _ = G.load(name)
    eng = make_engine(spec)
    ora = make_oracle(spec)
    rng = np.random.default_rng(17)
    n, L = eng.n, eng.L
    # a third uniformly random states, a third inside a small id window (dense conflicts), a third near the goal
    cells = rng.integers(0, L, (B, n)).astype(np.int32)
    base = rng.integers(0, max(1, L - 6), B)
    dense = (base[:, None] + rng.integers(0, 6, (B, n))).astype(np.int32)
    cells[B // 3: 2 * B // 3] = dense[B // 3: 2 * B // 3]
    goal = ora.decode(np.array([eng.goal_state & ((1 << 64) - 1)], np.uint64),
                      np.array([eng.goal_state >> 64], np.uint64))[0]
    near = np.tile(goal, (B, 1))
    jitter = rng.integers(0, n, B)
    near[np.arange(B), jitter] = rng.integers(0, L, B)
    cells[2 * B // 3:] = near[2 * B // 3:]
    lo, hi = ora.encode(cells)
    actions = rng.integers(0, eng.nA, B).astype(np.int64)
    actions[::7] = 0
    uniforms = rng.random((B, n))
    want = ora.step(lo, hi, actions, uniforms, threads=8)
    states = states_tensor(eng, lo, hi)
    ns, reward, prob, done, coll = eng.step(states, torch.from_numpy(actions.astype(np.int32)).to(eng.torch_device),
                                            uniforms=torch.from_numpy(uniforms).to(eng.torch_device))
    glo, ghi = split_states(eng, ns)
    assert np.array_equal(glo, want["next_lo"]) and np.array_equal(ghi, want["next_hi"])
    assert np.array_equal(u64(reward), G.f64_to_bits(want["reward"]))
    assert np.array_equal(u64(prob), G.f64_to_bits(want["prob"]))
    assert np.array_equal(done.cpu().numpy().astype(np.uint8), want["done"])
    assert np.array_equal(coll.cpu().numpy().astype(np.uint8), want["collision"])
    assert want["collision"].sum() > 0 and want["terminal"].sum() > 0 and (want["done"] & ~want["collision"]).sum() > 0


@pytest.mark.parametrize("name,B", [("rows_c2", 200000), ("rows_c4", 600), ("rows_c3", 20000), ("rows_c5_n7", 4000),
                                    ("rows_berlin", 100000)])
def test_expand_large_batch_vs_oracle(name, B, torch):
    spec, _ = G.load(name)
    eng = make_engine(spec)
    ora = make_oracle(spec)
    rng = np.random.default_rng(23)
    n, L = eng.n, eng.L
    cells = rng.integers(0, L, (B, n)).astype(np.int32)
    base = rng.integers(0, max(1, L - 5), B)
    dense = (base[:, None] + rng.integers(0, 5, (B, n))).astype(np.int32)
    cells[::2] = dense[::2]
    lo, hi = ora.encode(cells)
    actions = rng.integers(0, eng.nA, B).astype(np.int64)
    want = ora.rows(lo, hi, actions, threads=8)
    got = eng.transitions(states_tensor(eng, lo, hi), torch.from_numpy(actions.astype(np.int32)).to(eng.torch_device))
    assert_rows_equal(eng, got, want["row_ptr"], want["next_lo"], want["next_hi"], G.f64_to_bits(want["prob"]),
                      G.f64_to_bits(want["reward"]), want["done"], want["collision"])
    # size-independent property: every row's probabilities sum to 1 (within fp64 rounding of a <=3**n-term sum)
    row_ptr, _, prob, _, _ = got
    sums = torch.zeros(B, dtype=torch.float64, device=eng.torch_device)
    rows = torch.repeat_interleave(torch.arange(B, device=eng.torch_device), row_ptr[1:] - row_ptr[:-1])
    sums.index_add_(0, rows, prob)
    assert float((sums - 1.0).abs().max()) < 1e-9


def test_table_range_slab_c3(torch):
    """C3-style slab: consecutive joint states x all 15 625 actions of maze-32-32-4 (6 agents), vs the oracle."""
    spec, _ = G.load("rows_c3")
    eng = make_engine(spec)
    ora = make_oracle(spec)
    s_begin = int(spec["nS"]) // 8 * 3 + 12345
    n_states = 3
    got = eng.table_range(s_begin, n_states)
    want = ora.table_checksums(s_begin, n_states)
    cs = eng.checksum(got[1], got[2], got[3], got[4]).cpu().numpy().view(np.uint64)
    keys = ["count", "n_collision", "n_done", "sum_next_lo", "sum_next_hi", "sum_prob_bits", "sum_reward_bits", "ordered"]
    assert {k: int(v) for k, v in zip(keys, cs)} == want


def test_rollout_equals_repeated_steps(torch):
    spec, _ = G.load("rows_c2")
    eng = make_engine(spec)
    ora = make_oracle(spec)
    rng = np.random.default_rng(3)
    B, T, n = 5000, 12, eng.n
    actions = rng.integers(0, eng.nA, (T, B)).astype(np.int32)
    uniforms = rng.random((T, B, n))
    for auto_reset in (False, True):
        states = eng.states_from_ints([eng.s0] * B)
        out = eng.rollout(states, torch.from_numpy(actions).to(eng.torch_device), T,
                          uniforms=torch.from_numpy(uniforms).to(eng.torch_device), auto_reset=auto_reset)
        cur_lo = np.full(B, eng.s0, dtype=np.uint64)
        zeros = np.zeros(B, np.uint64)
        for t in range(T):
            w = ora.step(cur_lo, zeros, actions[t].astype(np.int64), uniforms[t])
            nxt = w["next_lo"].copy()
            if auto_reset:
                nxt[w["done"] == 1] = eng.s0
            assert np.array_equal(u64(out[0][t]), nxt), (auto_reset, t)
            assert np.array_equal(u64(out[1][t]), G.f64_to_bits(w["reward"]))
            assert np.array_equal(u64(out[2][t]), G.f64_to_bits(w["prob"]))
            assert np.array_equal(out[3][t].cpu().numpy().astype(np.uint8), w["done"])
            assert np.array_equal(out[4][t].cpu().numpy().astype(np.uint8), w["collision"])
            cur_lo = nxt
        assert np.array_equal(u64(states), cur_lo)


def philox4x32_10(c, k):
    """numpy restatement of Philox4x32-10 (Salmon et al. 2011) for checking the device stream."""
    c = [np.asarray(x, dtype=np.uint64) for x in c]
    k0, k1 = np.uint64(k[0]), np.uint64(k[1])
    m32 = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c[0]
        p1 = np.uint64(0xCD9E8D57) * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & m32, p1 >> np.uint64(32), p1 & m32
        c = [(hi1 ^ c[1] ^ k0) & m32, lo1, (hi0 ^ c[3] ^ k1) & m32, lo0]
        k0 = (k0 + np.uint64(0x9E3779B9)) & m32
        k1 = (k1 + np.uint64(0xBB67AE85)) & m32
    return c


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    out = philox4x32_10([0, 0, 0, 0], [0, 0])
    assert [int(x) for x in out] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    out = philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2)
    assert [int(x) for x in out] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    out = philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])
    assert [int(x) for x in out] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


@pytest.mark.parametrize("name", ["rows_c2", "rows_c5_n7"])
def test_device_sampling_stream(name, torch):
    """Perf mode (uniforms=None): the device draws agent i's uniform as word i%4 of Philox block i//4 at counter
    (env_offset + env, step).  Replaying those uniforms through the oracle must give the same step."""
    spec, _ = G.load(name)
    eng = make_engine(spec)
    ora = make_oracle(spec)
    B, n = 50000, eng.n
    seed, step_index, env_offset = 0x123456789abcdef, (5 << 32) | 77, 1 << 33
    rng = np.random.default_rng(9)
    cells = rng.integers(0, eng.L, (B, n)).astype(np.int32)
    lo, hi = ora.encode(cells)
    actions = rng.integers(0, eng.nA, B).astype(np.int64)
    env = np.arange(B, dtype=np.uint64) + np.uint64(env_offset)
    uniforms = np.zeros((B, n))
    for blk in range((n + 3) // 4):
        words = philox4x32_10([env & np.uint64(0xFFFFFFFF), env >> np.uint64(32),
                               np.full(B, step_index & 0xFFFFFFFF, np.uint64),
                               np.full(B, ((step_index >> 32) << 8) | blk, np.uint64)],
                              [seed & 0xFFFFFFFF, seed >> 32])
        for q in range(4):
            if blk * 4 + q < n:
                uniforms[:, blk * 4 + q] = words[q].astype(np.float64) * 2.0 ** -32
    want = ora.step(lo, hi, actions, uniforms)
    ns, reward, prob, done, coll = eng.step(states_tensor(eng, lo, hi),
                                            torch.from_numpy(actions.astype(np.int32)).to(eng.torch_device),
                                            seed=seed, step_index=step_index, env_offset=env_offset)
    glo, ghi = split_states(eng, ns)
    assert np.array_equal(glo, want["next_lo"]) and np.array_equal(ghi, want["next_hi"])
    assert np.array_equal(u64(prob), G.f64_to_bits(want["prob"]))
    assert np.array_equal(u64(reward), G.f64_to_bits(want["reward"]))
    # slip frequencies: with fail_prob 0.2 an unobstructed agent keeps its intended move 80% of the time
    frac = float((want["prob"] > 0).mean())
    assert frac > 0.5


def test_errors_and_edge_cases(torch):
    from gym_mapf_b200 import _native
    spec, _ = G.load("rows_obst4_n2_v0_soc")
    obst = np.array([[1 if ch == "@" else 0 for ch in row] for row in spec["rows"]], dtype=np.uint8)
    with pytest.raises(KeyError):  # start on an obstacle (reference: KeyError from loc_to_int, mapf_env.py:369)
        _native.Engine(obst, 2, [[0, 2], [3, 2]], spec["goals"], 0.2, -1000.0, 100.0, -1.0, False)
    with pytest.raises(KeyError):  # goal off the grid
        _native.Engine(obst, 2, spec["starts"], [[9, 9], [0, 1]], 0.2, -1000.0, 100.0, -1.0, False)
    with pytest.raises(_native.NativeError):
        _native.Engine(obst, 14, [[0, 0]] * 14, [[0, 0]] * 14, 0.2, -1000.0, 100.0, -1.0, False)
    eng = make_engine(spec)
    # empty batches
    empty_s = eng.new_states(0)
    empty_a = torch.empty(0, dtype=torch.int32, device=eng.torch_device)
    row_ptr, ns, prob, reward, flags = eng.transitions(empty_s, empty_a)
    assert row_ptr.tolist() == [0] and ns.numel() == 0
    out = eng.step(empty_s, empty_a)
    assert out[0].numel() == 0
    # ragged sizes around warp / block boundaries
    ora = make_oracle(spec)
    for B in (1, 31, 32, 33, 255, 257, 2049):
        rng = np.random.default_rng(B)
        lo = rng.integers(0, eng.nS, B).astype(np.uint64)
        a = rng.integers(0, eng.nA, B).astype(np.int64)
        want = ora.rows(lo, np.zeros_like(lo), a)
        got = eng.transitions(states_tensor(eng, lo, np.zeros_like(lo)),
                              torch.from_numpy(a.astype(np.int32)).to(eng.torch_device))
        assert_rows_equal(eng, got, want["row_ptr"], want["next_lo"], want["next_hi"], G.f64_to_bits(want["prob"]),
                          G.f64_to_bits(want["reward"]), want["done"], want["collision"])


@pytest.mark.parametrize("name", ["rows_c2", "rows_c4"])
def test_step_host_zero_copy_and_staged(name, torch):
    """mapf_step_host: pinned host buffers take the zero-copy path (the kernel reads/writes host memory over PCIe),
    pageable numpy buffers the staged-copy path; both must equal the oracle on the same uniforms, and the Philox mode
    must equal the device-buffer step."""
    spec, _ = G.load(name)
    eng = make_engine(spec)
    ora = make_oracle(spec)
    B, n = 70001, eng.n
    rng = np.random.default_rng(21)
    cells = rng.integers(0, eng.L, (B, n)).astype(np.int32)
    lo, hi = ora.encode(cells)
    actions = rng.integers(0, eng.nA, B).astype(np.int32)
    uniforms = rng.random((B, n))
    want = ora.step(lo, hi, actions.astype(np.int64), uniforms)
    host_states = states_tensor(eng, lo, hi).cpu()

    def outputs(pinned):
        out = (torch.empty(eng.state_shape(B), dtype=torch.int64), torch.empty(B, dtype=torch.float64),
               torch.empty(B, dtype=torch.float64), torch.empty(B, dtype=torch.bool), torch.empty(B, dtype=torch.bool))
        return tuple(t.pin_memory() for t in out) if pinned else out

    for pinned in (True, False):
        st = host_states.pin_memory() if pinned else host_states.clone()
        ac = torch.from_numpy(actions)
        un = torch.from_numpy(uniforms)
        if pinned:
            ac, un = ac.pin_memory(), un.pin_memory()
        out = outputs(pinned)
        eng.step_host(st, ac, out, uniforms=un)
        ns = out[0].numpy().view(np.uint64)
        glo, ghi = (ns, np.zeros_like(ns)) if eng.words == 1 else (ns[:, 0], ns[:, 1])
        assert np.array_equal(glo, want["next_lo"]) and np.array_equal(ghi, want["next_hi"]), pinned
        assert np.array_equal(out[1].numpy().view(np.uint64), G.f64_to_bits(want["reward"]))
        assert np.array_equal(out[2].numpy().view(np.uint64), G.f64_to_bits(want["prob"]))
        assert np.array_equal(out[3].numpy().astype(np.uint8), want["done"])
        assert np.array_equal(out[4].numpy().astype(np.uint8), want["collision"])
        # device-side sampling: host path == device path
        out2 = outputs(pinned)
        eng.step_host(st, ac, out2, seed=99, step_index=5, env_offset=123)
        dev = eng.step(states_tensor(eng, lo, hi), torch.from_numpy(actions).to(eng.torch_device), seed=99, step_index=5,
                       env_offset=123)
        assert np.array_equal(out2[0].numpy(), dev[0].cpu().numpy())
        assert np.array_equal(out2[2].numpy().view(np.uint64), u64(dev[2]))


# ---- rows next to the hot path ---------------------------------------------------------------------
@pytest.mark.parametrize("name", G.names("backup_"))
def test_backup_matches_reference(name, torch):
    """k_backup vs the reference's own loop over env.P (fixtures from oracle/make_golden_next.py), bit for bit."""
    spec, d = G.load(name)
    eng = make_engine(spec)
    V = torch.from_numpy(G.bits_to_f64(d["v_bits"]).copy()).to(eng.torch_device)
    gamma = float(G.bits_to_f64(np.array([d["gamma_bits"]]))[0])
    Q = eng.backup(states_tensor(eng, d["state_lo"], d["state_hi"]),
                   torch.from_numpy(d["action"].astype(np.int32)).to(eng.torch_device), V, gamma)
    assert np.array_equal(u64(Q), d["q_bits"])
    if name.startswith("backup_c1"):  # the slab entry point on the complete table + greedy
        Qr = eng.backup_range(0, eng.nS, V, gamma)
        assert np.array_equal(u64(Qr).reshape(-1), d["q_bits"])
        v, pi = eng.greedy(Qr)
        want = G.bits_to_f64(d["q_bits"]).reshape(eng.nS, eng.nA)
        assert np.array_equal(u64(v), G.f64_to_bits(want.max(axis=1)))
        assert np.array_equal(pi.cpu().numpy(), want.argmax(axis=1).astype(np.int32))


def test_backup_large_vs_oracle(torch):
    """2 agents on empty-32-32 (2**20 states): a whole sweep against the C oracle, then value iteration end to end."""
    spec, _ = G.load("rows_c5_n2")
    eng = make_engine(spec)
    ora = make_oracle(spec)
    rng = np.random.default_rng(77)
    V = rng.normal(0, 30, eng.nS)
    n_states, s0 = 3000, 517000
    s = np.repeat(np.arange(s0, s0 + n_states, dtype=np.uint64), eng.nA)
    a = np.tile(np.arange(eng.nA, dtype=np.int64), n_states)
    want = ora.backup(s, np.zeros_like(s), a, V, 0.97)
    got = eng.backup_range(s0, n_states, torch.from_numpy(V).to(eng.torch_device), 0.97)
    assert np.array_equal(u64(got).reshape(-1), G.f64_to_bits(want))


def test_value_iteration_c1(torch):
    from gym_mapf_b200.envs.mapf_env import OptimizationCriteria
    from gym_mapf_b200.envs.utils import create_mapf_env
    from gym_mapf_b200.envs.vec_env import VecMapfEnv
    env = create_mapf_env("empty-8-8", 1, 2, 0.2, -1000.0, 100.0, -1.0, OptimizationCriteria.Makespan, device=0)
    vec = VecMapfEnv(env, 4)
    V, pi, iters = vec.value_iteration(gamma=1.0, eps=1e-6, max_iter=200)
    # the same sweeps with the C oracle
    spec = dict(rows=["".join("@" if v else "." for v in r) for r in env.grid.obstacles], n_agents=2,
                goals=[list(g) for g in env.agents_goals], fail_prob=0.2, r_clash=-1000.0, r_goal=100.0, r_living=-1.0,
                soc=False)
    ora = make_oracle(spec)
    s = np.repeat(np.arange(env.nS, dtype=np.uint64), env.nA)
    a = np.tile(np.arange(env.nA, dtype=np.int64), env.nS)
    Vh = np.zeros(env.nS)
    for _ in range(iters):
        Vh = ora.backup(s, np.zeros_like(s), a, Vh, 1.0, threads=4).reshape(env.nS, env.nA).max(axis=1)
    assert np.array_equal(u64(V), G.f64_to_bits(Vh))
    assert iters < 200 and float(V[env.s]) > 0  # the start state is worth reaching the goal


@pytest.mark.parametrize("name", G.names("preds_"))
def test_predecessors_match_reference(name, torch):
    spec, d = G.load(name)
    eng = make_engine(spec)
    row_ptr, pred = eng.predecessors(states_tensor(eng, d["state_lo"], d["state_hi"]))
    assert np.array_equal(row_ptr.cpu().numpy(), d["row_ptr"])
    lo, hi = split_states(eng, pred)
    for b in range(len(d["state_lo"])):  # the reference returns a set: compare sorted
        sl = slice(int(d["row_ptr"][b]), int(d["row_ptr"][b + 1]))
        got = np.sort((hi[sl].astype(object) << 64) | lo[sl].astype(object)) if eng.words == 2 else np.sort(lo[sl])
        want = ((d["pred_hi"][sl].astype(object) << 64) | d["pred_lo"][sl].astype(object)) if eng.words == 2 \
            else d["pred_lo"][sl]
        assert np.array_equal(got, want), b


@pytest.mark.parametrize("name", G.names("project_"))
def test_projection_matches_reference(name, torch):
    spec, d = G.load(name)
    eng = make_engine(spec)
    st = states_tensor(eng, d["state_lo"], d["state_hi"])
    k = 0
    while "agents_%d" % k in d:
        out = eng.project(st, d["agents_%d" % k]).cpu().numpy().view(np.uint64)
        if out.ndim == 1:
            assert np.array_equal(out, d["proj_lo_%d" % k]) and not d["proj_hi_%d" % k].any()
        else:
            assert np.array_equal(out[:, 0], d["proj_lo_%d" % k]) and np.array_equal(out[:, 1], d["proj_hi_%d" % k])
        k += 1
    with pytest.raises(Exception):
        eng.project(st, [0, 0])


def test_greedy_bcast_single_peer(torch):
    """The fused exchange step with one 'peer' (the local vector): values land at their global positions."""
    spec, d = G.load("backup_obst4_n2")
    eng = make_engine(spec)
    rng = np.random.default_rng(5)
    n_states, s_begin = 57, 40
    Q = torch.from_numpy(rng.normal(0, 5, (n_states, eng.nA))).to(eng.torch_device)
    Q[3, 7] = Q[3, 2] = Q[3].max() + 1.0  # a tie: the first maximum wins
    v, pi = eng.greedy(Q)
    full = torch.full((eng.nS,), -1.0, dtype=torch.float64, device=eng.torch_device)
    pi2 = eng.greedy_bcast(Q, s_begin, [full.data_ptr()])
    assert torch.equal(pi, pi2) and int(pi[3]) == 2
    assert torch.equal(full[s_begin:s_begin + n_states], v)
    assert bool((full[:s_begin] == -1).all()) and bool((full[s_begin + n_states:] == -1).all())
    assert torch.equal(v, Q.max(dim=1).values)


# ---- BASELINE.json's full sizes through size-independent properties ---------------------------------------------------
def _c4_engine():
    from gym_mapf_b200.envs.mapf_env import OptimizationCriteria
    from gym_mapf_b200.envs.utils import create_mapf_env
    env = create_mapf_env("room-64-64-8", 1, 8, 0.2, -1000.0, 100.0, -1.0, OptimizationCriteria.Makespan, device=0)
    return env, env.engine


def test_c4_full_size_step_properties(torch):
    """16 M envs, 8 agents, 128-bit states (BASELINE configs[3]): the vectorised and the scalar kernel variants agree,
    a batch stepped in two shards with the right env offsets equals the batch stepped whole (disjoint Philox
    streams), every next state is inside the state space, terminal states are fixed points with reward 0 / prob 0."""
    env, eng = _c4_engine()
    B = 1 << 24
    dev = eng.torch_device
    g = torch.Generator(device=dev)
    g.manual_seed(3)
    cells = torch.randint(0, eng.L, (B, eng.n), generator=g, device=dev, dtype=torch.int32)
    cells[:1000, 1] = cells[:1000, 0]  # some already-clashed (terminal) states
    states = eng.encode(cells)
    actions = torch.randint(0, eng.nA, (B,), generator=g, device=dev, dtype=torch.int32)
    whole = eng.step(states, actions, seed=11, step_index=4)
    ns, reward, prob, done, coll = whole
    # (a) scalar variant: an odd batch on misaligned views takes the one-env-per-thread kernel
    part = eng.step(states[1:B - 2], actions[1:B - 2], seed=11, step_index=4, env_offset=1)
    for a, b in zip(whole, part):
        assert torch.equal(a[1:B - 2], b)
    # (b) two shards == whole
    h = B // 2
    lo_half = eng.step(states[:h], actions[:h], seed=11, step_index=4, env_offset=0)
    hi_half = eng.step(states[h:], actions[h:], seed=11, step_index=4, env_offset=h)
    for a, b, c in zip(whole, lo_half, hi_half):
        assert torch.equal(a[:h], b) and torch.equal(a[h:], c)
    # (c) range and terminal fixed points
    back = eng.decode(ns)
    assert int(back.min()) >= 0 and int(back.max()) < eng.L
    assert torch.equal(eng.encode(back), ns)
    assert torch.equal(ns[:1000], states[:1000]) and bool(done[:1000].all()) and not bool(coll[:1000].any())
    assert float(reward[:1000].abs().max()) == 0.0 and float(prob[:1000].abs().max()) == 0.0
    # (d) a collision always ends the episode; probabilities are products of {0.8, 0.1, 0.9, 1.0, 0.2}
    assert bool(done[coll].all())
    assert float(prob.min()) >= 0.0 and float(prob.max()) <= 1.0
    zero = prob == 0.0  # exactly the steps from a terminal state (two agents on one cell): no-ops
    srt = cells.sort(dim=1).values
    dup = (srt[:, 1:] == srt[:, :-1]).any(dim=1)
    assert torch.equal(zero, dup)
    assert torch.equal(ns[zero], states[zero]) and bool(done[zero].all()) and float(reward[zero].abs().max()) == 0.0


def test_c2_full_size_expand_properties(torch):
    """2**20 random (s, a) rows of C2 (36 M records): every row's probabilities add up to 1, row lengths are
    products of 1/2/3, collision implies done, and the checksums equal the C oracle's on the same rows."""
    spec, _ = G.load("rows_c2")
    eng = make_engine(spec)
    ora = make_oracle(spec)
    B = 1 << 20
    rng = np.random.default_rng(8)
    cells = rng.integers(0, eng.L, (B, eng.n)).astype(np.int32)
    lo, hi = ora.encode(cells)
    a = rng.integers(0, eng.nA, B).astype(np.int64)
    row_ptr, ns, prob, reward, flags = eng.transitions(states_tensor(eng, lo, hi),
                                                       torch.from_numpy(a.astype(np.int32)).to(eng.torch_device))
    lens = (row_ptr[1:] - row_ptr[:-1])
    assert int(lens.min()) >= 1 and int(lens.max()) <= 81
    seg = torch.repeat_interleave(torch.arange(B, device=eng.torch_device), lens)
    sums = torch.zeros(B, dtype=torch.float64, device=eng.torch_device).index_add_(0, seg, prob)
    assert float((sums - 1.0).abs().max()) < 1e-12
    assert bool(((flags & 2) == 0).logical_or((flags & 1) == 1).all())
    got = eng.checksum(ns, prob, reward, flags).cpu().numpy().view(np.uint64)
    want = ora.rows(lo, hi, a, threads=8)
    cs = G.checksums(want["next_lo"], want["next_hi"], G.f64_to_bits(want["prob"]), G.f64_to_bits(want["reward"]),
                     want["done"], want["collision"])
    assert [int(x) for x in got] == [cs[k] for k in ("count", "n_collision", "n_done", "sum_next_lo", "sum_next_hi",
                                                     "sum_prob_bits", "sum_reward_bits", "ordered")]


def test_thirteen_agents(torch):
    """The largest supported agent count (5**13 joint actions, rows of up to 3**13 records) on a small grid."""
    rows = ["....", ".@..", "...."]
    starts = [[r, c] for r in range(3) for c in range(4) if not (r == 1 and c == 1)] + [[0, 0], [2, 3]]
    goals = list(reversed(starts))
    spec = dict(rows=rows, n_agents=13, starts=starts, goals=goals, fail_prob=0.2, r_clash=-1000.0, r_goal=100.0,
                r_living=-1.0, soc=True)
    eng = make_engine(spec)
    ora = make_oracle(spec)
    rng = np.random.default_rng(13)
    cells = np.stack([rng.permutation(11)[:11].tolist() + rng.integers(0, 11, 2).tolist() for _ in range(6)]).astype(np.int32)
    cells[0] = np.arange(13) % 11          # duplicates: terminal
    cells[1, :11] = np.arange(11)          # 11 distinct + 2 random
    lo, hi = ora.encode(cells)
    a = rng.integers(0, eng.nA, 6).astype(np.int64)
    a[2] = 0                               # everybody stays: a single outcome
    want = ora.rows(lo, hi, a)
    got = eng.transitions(states_tensor(eng, lo, hi), torch.from_numpy(a.astype(np.int32)).to(eng.torch_device))
    assert_rows_equal(eng, got, want["row_ptr"], want["next_lo"], want["next_hi"], G.f64_to_bits(want["prob"]),
                      G.f64_to_bits(want["reward"]), want["done"], want["collision"])
    uni = rng.random((6, 13))
    ws = ora.step(lo, hi, a, uni)
    ns, reward, prob, done, coll = eng.step(states_tensor(eng, lo, hi), torch.from_numpy(a.astype(np.int32)).to(eng.torch_device),
                                            uniforms=torch.from_numpy(uni).to(eng.torch_device))
    glo, ghi = split_states(eng, ns)
    assert np.array_equal(glo, ws["next_lo"]) and np.array_equal(ghi, ws["next_hi"])
    assert np.array_equal(u64(prob), G.f64_to_bits(ws["prob"])) and np.array_equal(u64(reward), G.f64_to_bits(ws["reward"]))


def _random_spec(rng):
    H, W = int(rng.integers(1, 8)), int(rng.integers(1, 8))
    while True:
        grid = rng.random((H, W)) < rng.choice([0.0, 0.15, 0.4])
        if (~grid).sum() >= 1:
            break
    free = [(r, c) for r in range(H) for c in range(W) if not grid[r, c]]
    n = int(rng.integers(1, 6))
    pick = lambda: [list(free[int(rng.integers(0, len(free)))]) for _ in range(n)]  # noqa: E731
    fp = float(rng.choice([0.0, 0.1, 0.2, 0.37, 0.5, 1.0]))
    return dict(rows=["".join("@" if v else "." for v in row) for row in grid], n_agents=n, starts=pick(), goals=pick(),
                fail_prob=fp, r_clash=float(rng.choice([-1000.0, -3.5, 0.0])), r_goal=float(rng.choice([100.0, 7.25, 1.0])),
                r_living=float(rng.choice([-1.0, -0.25, 0.0])), soc=bool(rng.integers(0, 2)))


def test_fuzz_random_specs(torch):
    """Differential test on 80 random env specs (tiny grids down to a single free cell, 1-5 agents, every noise level
    including 0 and 1, both criteria, starts/goals that may coincide): rows, steps with given uniforms, the fused
    backup, predecessors and projections must equal the C oracle bit for bit."""
    rng = np.random.default_rng(int(os.environ.get("MAPF_FUZZ_SEED", "2026")))
    for case in range(80):
        spec = _random_spec(rng)
        eng = make_engine(spec)
        ora = make_oracle(spec)
        nS, nA = eng.nS, eng.nA
        B = int(min(4000, nS * nA))
        if nS * nA <= 4000:
            s = np.repeat(np.arange(nS, dtype=np.uint64), nA)
            a = np.tile(np.arange(nA, dtype=np.int64), nS)
        else:
            s = rng.integers(0, nS, B).astype(np.uint64)
            a = rng.integers(0, nA, B).astype(np.int64)
        z = np.zeros_like(s)
        st = states_tensor(eng, s, z)
        at = torch.from_numpy(a.astype(np.int32)).to(eng.torch_device)
        want = ora.rows(s, z, a)
        got = eng.transitions(st, at)
        assert_rows_equal(eng, got, want["row_ptr"], want["next_lo"], want["next_hi"], G.f64_to_bits(want["prob"]),
                          G.f64_to_bits(want["reward"]), want["done"], want["collision"])
        uni = rng.random((len(s), eng.n))
        ws = ora.step(s, z, a, uni)
        ns, reward, prob, done, coll = eng.step(st, at, uniforms=torch.from_numpy(uni).to(eng.torch_device))
        assert np.array_equal(u64(ns), ws["next_lo"]), (case, spec)
        assert np.array_equal(u64(reward), G.f64_to_bits(ws["reward"])) and np.array_equal(u64(prob), G.f64_to_bits(ws["prob"]))
        assert np.array_equal(done.cpu().numpy().astype(np.uint8), ws["done"])
        assert np.array_equal(coll.cpu().numpy().astype(np.uint8), ws["collision"])
        V = rng.normal(0, 20, nS)
        q = eng.backup(st, at, torch.from_numpy(V).to(eng.torch_device), 0.9)
        assert np.array_equal(u64(q), G.f64_to_bits(ora.backup(s, z, a, V, 0.9))), (case, spec)
        wp = ora.predecessors(s[:200], z[:200])
        row_ptr, pred = eng.predecessors(st[:200])
        assert np.array_equal(row_ptr.cpu().numpy(), wp["row_ptr"])
        plo = u64(pred)
        for b in range(min(200, len(s))):
            sl = slice(int(wp["row_ptr"][b]), int(wp["row_ptr"][b + 1]))
            assert np.array_equal(np.sort(plo[sl]), wp["pred_lo"][sl])
        sub = [int(x) for x in rng.permutation(eng.n)[:int(rng.integers(1, eng.n + 1))]]
        assert np.array_equal(u64(eng.project(st, sub)), ora.project(s, z, sub)[0])
        # the device-side sampling mode needs slip probabilities that add up to 1 (always true here)
        eng.step(st, at, seed=case)
        eng.close()


def test_fuzz_medium_agent_counts(torch):
    """Differential test for 6-9 agents on random 8..16 x 8..16 grids (one- and two-word states; the conflict pair lists
    from 8 agents on, the head cache from 9): agents drawn from a small window so that vertex and swap conflicts are
    frequent; rows of up to a few thousand records and steps with given uniforms against the C oracle."""
    rng = np.random.default_rng(int(os.environ.get("MAPF_FUZZ_SEED", "91")))
    words_seen = set()
    for case in range(16):
        n = 6 + case % 4
        lo_side, hi_side = (13, 17) if n == 9 else (8, 13)   # 9 agents on >= 128 free cells need two words
        H, W = int(rng.integers(lo_side, hi_side)), int(rng.integers(lo_side, hi_side))
        grid = rng.random((H, W)) < 0.12
        free = [(r, c) for r in range(H) for c in range(W) if not grid[r, c]]
        pick = lambda: [list(free[int(rng.integers(0, len(free)))]) for _ in range(n)]  # noqa: E731
        spec = dict(rows=["".join("@" if v else "." for v in row) for row in grid], n_agents=n, starts=pick(), goals=pick(),
                    fail_prob=float(rng.choice([0.2, 0.5, 1.0])), r_clash=-1000.0, r_goal=100.0, r_living=-1.0,
                    soc=bool(rng.integers(0, 2)))
        eng = make_engine(spec)
        ora = make_oracle(spec)
        words_seen.add(eng.words)
        B = 2000
        window = min(eng.L, 3 * n)
        cells = rng.integers(0, window, (B, n)).astype(np.int32)
        cells[::2] = np.stack([rng.permutation(window)[:n] for _ in range(len(cells[::2]))])  # no duplicates: not terminal
        lo, hi = ora.encode(cells)
        st = states_tensor(eng, lo, hi)
        a = rng.integers(0, eng.nA, B).astype(np.int64)
        at = torch.from_numpy(a.astype(np.int32)).to(eng.torch_device)
        uni = rng.random((B, n))
        ws = ora.step(lo, hi, a, uni)
        ns, reward, prob, done, coll = eng.step(st, at, uniforms=torch.from_numpy(uni).to(eng.torch_device))
        glo, ghi = split_states(eng, ns)
        assert np.array_equal(glo, ws["next_lo"]) and np.array_equal(ghi, ws["next_hi"]), (case, eng.L, n)
        assert np.array_equal(u64(prob), G.f64_to_bits(ws["prob"])) and np.array_equal(u64(reward), G.f64_to_bits(ws["reward"]))
        assert np.array_equal(coll.cpu().numpy().astype(np.uint8), ws["collision"])
        assert int(ws["collision"].sum()) > 0
        # rows: about half of the agents move, so a row has up to 3**(n/2) records
        m = 40
        digits = (rng.random((m, n)) < 0.5) * rng.integers(1, 5, (m, n))
        ar = (digits * (5 ** np.arange(n))).sum(axis=1).astype(np.int64)
        want = ora.rows(lo[:m], hi[:m], ar)
        got = eng.transitions(st[:m], torch.from_numpy(ar.astype(np.int32)).to(eng.torch_device))
        assert_rows_equal(eng, got, want["row_ptr"], want["next_lo"], want["next_hi"], G.f64_to_bits(want["prob"]),
                          G.f64_to_bits(want["reward"]), want["done"], want["collision"])
        assert int(want["collision"].sum()) > 0, (case, n)
        eng.close()
    assert words_seen == {1, 2}


def test_fuzz_two_word_specs(torch):
    """The same differential test where the joint state needs two words (10-13 agents on 40-64 free cells): both the
    split decode/encode (L**k groups) and the limb long division are exercised, depending on L and n."""
    rng = np.random.default_rng(77)
    seen_split = set()
    for case in range(10):
        H, W = int(rng.integers(6, 9)), int(rng.integers(7, 9))
        grid = rng.random((H, W)) < 0.1
        free = [(r, c) for r in range(H) for c in range(W) if not grid[r, c]]
        n = int(rng.integers(10, 14))
        if len(free) ** n < 2 ** 63:
            continue
        pick = lambda: [list(free[int(rng.integers(0, len(free)))]) for _ in range(n)]  # noqa: E731
        spec = dict(rows=["".join("@" if v else "." for v in row) for row in grid], n_agents=n, starts=pick(), goals=pick(),
                    fail_prob=float(rng.choice([0.2, 0.5])), r_clash=-1000.0, r_goal=100.0, r_living=-1.0,
                    soc=bool(rng.integers(0, 2)))
        eng = make_engine(spec)
        assert eng.words == 2
        ora = make_oracle(spec)
        B = 3000
        cells = rng.integers(0, eng.L, (B, n)).astype(np.int32)
        cells[::3] = np.stack([rng.permutation(eng.L)[:n] for _ in range(len(cells[::3]))])  # no duplicates: not terminal
        lo, hi = ora.encode(cells)
        st = states_tensor(eng, lo, hi)
        a = rng.integers(0, eng.nA, B).astype(np.int64)
        at = torch.from_numpy(a.astype(np.int32)).to(eng.torch_device)
        # decode / encode round trip and steps on the whole batch
        assert np.array_equal(eng.decode(st).cpu().numpy(), cells)
        uni = rng.random((B, n))
        ws = ora.step(lo, hi, a, uni)
        ns, reward, prob, done, coll = eng.step(st, at, uniforms=torch.from_numpy(uni).to(eng.torch_device))
        glo, ghi = split_states(eng, ns)
        assert np.array_equal(glo, ws["next_lo"]) and np.array_equal(ghi, ws["next_hi"]), (case, eng.L, n)
        assert np.array_equal(u64(prob), G.f64_to_bits(ws["prob"])) and np.array_equal(u64(reward), G.f64_to_bits(ws["reward"]))
        # a few rows (each up to 3**n records): mostly-STAY actions keep them small
        digits = (rng.random((6, n)) < 0.35) * rng.integers(1, 5, (6, n))
        ar = (digits * (5 ** np.arange(n))).sum(axis=1).astype(np.int64)
        want = ora.rows(lo[:6], hi[:6], ar)
        got = eng.transitions(st[:6], torch.from_numpy(ar.astype(np.int32)).to(eng.torch_device))
        assert_rows_equal(eng, got, want["row_ptr"], want["next_lo"], want["next_hi"], G.f64_to_bits(want["prob"]),
                          G.f64_to_bits(want["reward"]), want["done"], want["collision"])
        seen_split.add((eng.L, n))
        eng.close()
    assert len(seen_split) >= 4


def test_batch_validation(torch):
    spec, _ = G.load("rows_c2")
    eng = make_engine(spec)
    st = eng.states_from_ints([eng.s0] * 8)
    good = torch.zeros(8, dtype=torch.int32, device=eng.torch_device)
    eng.step(st, good)
    for bad in (good.to(torch.int64), good[:7], good.cpu(), torch.zeros(16, dtype=torch.int32, device=eng.torch_device)[::2]):
        with pytest.raises(ValueError):
            eng.step(st, bad)
        with pytest.raises(ValueError):
            eng.transitions(st, bad)
    with pytest.raises(ValueError):
        eng.step(st.to(torch.int32), good)
    with pytest.raises(ValueError):
        eng.step(st, good, uniforms=torch.zeros(8, 3, dtype=torch.float64, device=eng.torch_device))


def test_next_row_error_paths(torch):
    from gym_mapf_b200 import _native
    spec, _ = G.load("rows_c2")
    eng = make_engine(spec)
    st = eng.states_from_ints([eng.s0] * 4)
    ac = torch.zeros(4, dtype=torch.int32, device=eng.torch_device)
    small_v = torch.zeros(10, dtype=torch.float64, device=eng.torch_device)
    with pytest.raises(_native.NativeError):  # V does not cover the state space
        eng.backup(st, ac, small_v, 0.9)
    with pytest.raises(_native.NativeError):  # the slab leaves the state space
        eng.backup_range(eng.nS - 1, 2, small_v, 0.9)
    with pytest.raises(_native.NativeError):
        eng.project(st, [0, 9])
    with pytest.raises(_native.NativeError):
        eng.greedy_bcast(torch.zeros((2, eng.nA), dtype=torch.float64, device=eng.torch_device), 0, [])
    # two-word contexts have no value vector to back up against
    spec4, _ = G.load("rows_c4")
    eng4 = make_engine(spec4)
    st4 = eng4.states_from_ints([eng4.s0] * 2)
    with pytest.raises(_native.NativeError) as ei:
        eng4.backup(st4, ac[:2], small_v, 0.9)
    assert ei.value.code == _native.MAPF_ERR_UNSUPPORTED
    # empty batches are fine everywhere
    e_s, e_a = eng.new_states(0), ac[:0]
    big_v = torch.zeros(8, dtype=torch.float64, device=eng.torch_device)
    assert eng.predecessors(e_s)[0].tolist() == [0]
    assert eng.project(e_s, [1, 0]).numel() == 0
