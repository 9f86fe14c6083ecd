"""GPU tests added in round 2: contexts that share kernel functions, the host/device state mirror of VecMapfEnv,
host-side validation of every batched entry point, and the bench's own oracle checker on small cases."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch as t
    return t


def _env(name, scen, n, soc=True):
    from gym_mapf_b200.envs.mapf_env import OptimizationCriteria
    from gym_mapf_b200.envs.utils import create_mapf_env
    crit = OptimizationCriteria.SoC if soc else OptimizationCriteria.Makespan
    return create_mapf_env(name, scen, n, 0.2, -1000.0, 100.0, -1.0, crit, device=0)


def _oracle(env, soc=True):
    from oracle import c_oracle
    rows = ["".join("@" if v else "." for v in r) for r in env.grid.obstacles]
    return c_oracle.COracle(rows, env.n_agents, env.agents_goals, 0.2, -1000.0, 100.0, -1.0, soc)


def _check_step(torch, env, B=20000, seed=3):
    eng, ora = env.engine, _oracle(env)
    rng = np.random.default_rng(seed)
    cells = rng.integers(0, eng.L, (B, eng.n)).astype(np.int32)
    lo, hi = ora.encode(cells)
    actions = rng.integers(0, eng.nA, B).astype(np.int32)
    uniforms = rng.random((B, eng.n))
    want = ora.step(lo, hi, actions.astype(np.int64), uniforms)
    st = eng.encode(torch.from_numpy(cells).cuda())
    ns, reward, prob, done, coll = eng.step(st, torch.from_numpy(actions).cuda(), uniforms=torch.from_numpy(uniforms).cuda())
    arr = ns.cpu().numpy().view(np.uint64)
    glo = arr if eng.words == 1 else arr.reshape(-1, 2)[:, 0]
    assert np.array_equal(glo, want["next_lo"])
    assert np.array_equal(reward.cpu().numpy().view(np.uint64), want["reward"].view(np.uint64))
    assert np.array_equal(prob.cpu().numpy().view(np.uint64), want["prob"].view(np.uint64))
    tr = eng.transitions(st[:500], torch.from_numpy(actions[:500]).cuda())
    w = ora.rows(lo[:500], hi[:500], actions[:500].astype(np.int64))
    assert np.array_equal(tr[0].cpu().numpy(), w["row_ptr"])
    assert np.array_equal(tr[2].cpu().numpy().view(np.uint64), w["prob"].view(np.uint64))


def test_contexts_sharing_kernels_in_decreasing_table_size(torch):
    """Two contexts with the same agent count share the kernel functions; the second one's smaller staged table must not
    lower the dynamic shared-memory cap under the first (cudaFuncSetAttribute is per function)."""
    big = _env("room-64-64-16", 1, 4)     # 3648 cells: 146 KB move table
    _check_step(torch, big)
    small = _env("room-64-64-8", 1, 4)    # 3232 cells: 129 KB
    _check_step(torch, small)
    smaller = _env("room-32-32-4", 1, 4)  # 27 KB
    _check_step(torch, smaller)
    _check_step(torch, big, seed=4)       # launches of the FIRST context still fit
    _check_step(torch, small, seed=5)
    from gym_mapf_b200 import _native
    grp_big = _native.Group([big.engine, small.engine], [3000, 3000])
    grp_small = _native.Group([smaller.engine], [1000])
    st = torch.cat([big.engine.states_from_ints([big.engine.s0] * 3000), small.engine.states_from_ints([small.engine.s0] * 3000)])
    ac = torch.zeros(6000, dtype=torch.int32, device="cuda")
    grp_small.step(smaller.engine.states_from_ints([smaller.engine.s0] * 1000), ac[:1000])
    out = grp_big.step(st, ac, seed=1)  # the larger group, after a smaller one was created
    torch.cuda.synchronize()
    assert out[0].shape[0] == 6000


def test_vec_env_host_and_device_steps_interleave(torch):
    """step_host continues from wherever the last step / reset / set_states / rollout left the states, and the device
    path continues from the host mirror (the mirror went stale)."""
    from gym_mapf_b200.envs.vec_env import VecMapfEnv
    env = _env("room-32-32-4", 1, 4)
    B = 3000
    for reuse in (False, True):
        a = VecMapfEnv(env, B, seed=11, reuse_outputs=reuse)   # mixed host / device steps
        b = VecMapfEnv(env, B, seed=11)                         # device only
        rng = np.random.default_rng(1)
        plan = ["dev", "host", "host", "dev", "rollout", "host", "reset", "host", "dev", "set", "host", "dev"]
        for what in plan:
            acts = rng.integers(0, env.nA, B).astype(np.int32)
            if what == "reset":
                a.reset(); b.reset()
                continue
            if what == "set":
                st = b.states.clone()
                st[: B // 2] = b.states_from_ints([env.s] * (B // 2))
                a.set_states(st); b.set_states(st)
                continue
            if what == "rollout":
                ra = a.rollout(3); rb = b.rollout(3)
                assert torch.equal(ra.next_state, rb.next_state)
                continue
            want = b.step(torch.from_numpy(acts).cuda())
            if what == "dev":
                got = a.step(torch.from_numpy(acts).cuda())
                assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
            else:
                got = a.step_host(torch.from_numpy(acts).pin_memory())
                assert np.array_equal(got[0].numpy(), want[0].cpu().numpy()), what
                assert np.array_equal(got[1].numpy().view(np.uint64), want[1].cpu().numpy().view(np.uint64))
                assert np.array_equal(got[3].numpy(), want[2].cpu().numpy())


def test_validation_of_every_batched_entry_point(torch):
    env = _env("room-32-32-4", 1, 4)
    eng = env.engine
    st = eng.states_from_ints([eng.s0] * 8)
    with pytest.raises(ValueError):
        eng.rollout(st.to(torch.int32), None, 4)
    with pytest.raises(ValueError):
        eng.rollout(st, torch.zeros((8, 4), dtype=torch.int32, device="cuda"), 4)   # [B, T] instead of [T, B]
    with pytest.raises(ValueError):
        eng.rollout(st, torch.zeros((4, 8), dtype=torch.int64, device="cuda"), 4)
    with pytest.raises(ValueError):
        eng.rollout(st, None, 4, uniforms=torch.zeros((4, 8, 3), dtype=torch.float64, device="cuda"))
    with pytest.raises(ValueError):
        eng.decode(st.cpu())
    with pytest.raises(ValueError):
        eng.encode(torch.zeros((8, 3), dtype=torch.int32, device="cuda"))
    with pytest.raises(ValueError):
        eng.predecessors(st.to(torch.float64))
    with pytest.raises(ValueError):
        eng.project(st[::2], [0, 1])
    small = _env("empty-8-8", 1, 2, soc=False)
    e2 = small.engine
    V = torch.zeros(e2.nS, dtype=torch.float64, device="cuda")
    s2 = e2.states_from_ints([e2.s0] * 4)
    a2 = torch.zeros(4, dtype=torch.int32, device="cuda")
    with pytest.raises(ValueError):
        e2.backup(s2, a2, V.to(torch.float32), 0.9)
    with pytest.raises(ValueError):
        e2.backup(s2, a2.to(torch.int64), V, 0.9)
    with pytest.raises(ValueError):
        e2.backup_range(0, 4, V.cpu(), 0.9)
    assert e2.backup(s2, a2, V, 0.9).shape == (4,)


def test_bench_checker_on_small_cases(torch):
    """bench.py's `other_configs` machinery end to end on reduced sizes: every case must report parity, and the checker
    must catch a corrupted sample."""
    sys.path.insert(0, ROOT)
    import bench
    from tools import bench_configs
    cases = bench_configs.run_all(0, 1, 0, only={"c2_step_strong", "c3_table", "c4_step", "c2_expand", "c2_rollout",
                                                 "c2_rollout_random", "c5_n2", "c5_n8", "c5_density_w2"}, quick=True)
    assert len(cases) == 9
    for name, entry, payload in cases:
        ok, what = bench.verify_payload(payload)
        assert ok, (name, what)
        assert entry["frac"] > 0 and entry["value"] > 0
    name, entry, payload = cases[0]
    payload["reward"] = payload["reward"].copy()
    payload["reward"][17] += 1.0
    assert not bench.verify_payload(payload)[0]


def test_greedy_bcast_sharded_sweep_three_ranks_one_device(torch):
    """The fused exchange step of sharded value iteration with THREE ranks' value vectors (all on this device, one
    stream per rank): every rank's greedy launch stores its shard's new values into every rank's vector, so after the
    sweep all three vectors equal the single-GPU sweep.  (On a multi-GPU box the pointers are peer mappings; the kernel
    and the ABI call are the same.)"""
    from gym_mapf_b200 import sharding
    env = _env("empty-8-8", 1, 2, soc=False)
    eng = env.engine
    nS, nA = int(eng.nS), int(eng.nA)
    rng = np.random.default_rng(8)
    V0 = torch.from_numpy(rng.normal(0, 20, nS)).cuda()
    world = 3
    shards = [sharding.table_shard(0, nS, world, r) for r in range(world)]
    assert sum(s.count for s in shards) == nS and shards[0].count != shards[2].count  # uneven shards
    # reference: one sweep on one "rank"
    Q_all = eng.backup_range(0, nS, V0, 0.95)
    V_ref, pi_ref = eng.greedy(Q_all)
    vectors = [torch.full((nS,), float("nan"), dtype=torch.float64, device="cuda") for _ in range(world)]
    ptrs = [v.data_ptr() for v in vectors]
    streams = [torch.cuda.Stream() for _ in range(world)]
    policies = []
    torch.cuda.synchronize()
    for r, sh in enumerate(shards):
        with torch.cuda.stream(streams[r]):
            Q = eng.backup_range(sh.begin, sh.count, V0, 0.95)
            policies.append(eng.greedy_bcast(Q, sh.begin, ptrs))
    torch.cuda.synchronize()
    for v in vectors:
        assert torch.equal(v, V_ref)
    assert torch.equal(torch.cat(policies), pi_ref)


def test_compact_result_layout(torch):
    """MAPF_OPT_COMPACT: reward codes + one flag byte decode to exactly the default results (device buffers, pinned host
    buffers, pageable host buffers; Philox and replayed-uniform modes; odd batch = scalar kernel, even = 128-bit kernel)."""
    for name, scen, n, soc in (("room-32-32-4", 1, 4, True), ("room-64-64-8", 1, 8, False)):
        env = _env(name, scen, n, soc)
        eng = env.engine
        table = torch.from_numpy(eng.reward_table()).cuda()
        for B in (70001, 70000):
            rng = np.random.default_rng(B)
            cells = torch.from_numpy(rng.integers(0, min(eng.L, 40), (B, eng.n)).astype(np.int32)).cuda()  # dense: clashes
            st = eng.encode(cells)
            st[: B // 3] = eng.states_from_ints([eng.s0] * (B // 3))
            ac = torch.from_numpy(rng.integers(0, eng.nA, B).astype(np.int32)).cuda()
            un = torch.from_numpy(rng.random((B, eng.n))).cuda()
            for uniforms in (None, un):
                want = eng.step(st, ac, uniforms=uniforms, seed=5, step_index=9, auto_reset=True)
                ns, code, prob, flags = eng.step(st, ac, uniforms=uniforms, seed=5, step_index=9, auto_reset=True, compact=True)
                assert torch.equal(ns, want[0]) and torch.equal(prob, want[2])
                assert torch.equal(table[code.long()].view(torch.int64), want[1].view(torch.int64))
                assert torch.equal((flags & 1).bool(), want[3]) and torch.equal((flags >> 1).bool(), want[4])
                assert int(want[4].sum()) > 0 and int(want[3].sum()) > int(want[4].sum())
            for pinned in (True, False):
                mk = (lambda t: t.pin_memory()) if pinned else (lambda t: t)
                out = (mk(torch.empty(eng.state_shape(B), dtype=torch.int64)), mk(torch.empty(B, dtype=torch.uint8)),
                       mk(torch.empty(B, dtype=torch.float64)), mk(torch.empty(B, dtype=torch.uint8)))
                eng.step_host(mk(st.cpu()), mk(ac.cpu()), out, seed=5, step_index=9, auto_reset=True, compact=True)
                ref = eng.step(st, ac, seed=5, step_index=9, auto_reset=True, compact=True)
                for a, b in zip(out, ref):
                    assert torch.equal(a, b.cpu()), pinned


def test_step_host_resident_matches_the_oracle_and_step_host(torch):
    """mapf_step_host_resident (states stay on the device, only the actions come from the host): replayed uniforms equal
    the C oracle, Philox draws equal mapf_step_host, and the device-resident states end up equal to the returned next
    states -- one- and two-word states; pinned (zero-copy launch), pageable (staged copies) and small (pinned scratch)
    batches; even (128-bit kernel) and odd (scalar kernel) sizes; default and compact result layouts; two chained steps."""
    for name, scen, n, soc in (("room-32-32-4", 1, 4, True), ("room-64-64-8", 1, 8, False)):
        env = _env(name, scen, n, soc)
        eng, ora = env.engine, _oracle(env, soc)
        for B in (70000, 70001, 600):
            rng = np.random.default_rng(B)
            cells = rng.integers(0, min(eng.L, 60), (B, eng.n)).astype(np.int32)  # dense: clashes and terminal states
            lo, hi = ora.encode(cells)
            st0 = eng.encode(torch.from_numpy(cells).cuda())
            st0[: B // 3] = eng.states_from_ints([eng.s0] * (B // 3))
            h0 = st0.cpu().numpy().view(np.uint64)
            lo, hi = (h0, np.zeros_like(h0)) if eng.words == 1 else (h0[:, 0].copy(), h0[:, 1].copy())
            acts = rng.integers(0, eng.nA, (2, B)).astype(np.int32)
            unis = rng.random((2, B, eng.n))
            for pinned in (True, False):
                mk = (lambda t: t.pin_memory()) if pinned else (lambda t: t)

                def outs(compact=False):
                    if compact:
                        return (mk(torch.empty(eng.state_shape(B), dtype=torch.int64)), mk(torch.empty(B, dtype=torch.uint8)),
                                mk(torch.empty(B, dtype=torch.float64)), mk(torch.empty(B, dtype=torch.uint8)))
                    return (mk(torch.empty(eng.state_shape(B), dtype=torch.int64)), mk(torch.empty(B, dtype=torch.float64)),
                            mk(torch.empty(B, dtype=torch.float64)), mk(torch.empty(B, dtype=torch.bool)),
                            mk(torch.empty(B, dtype=torch.bool)))
                # replayed uniforms, reference semantics (no auto-reset), two chained steps vs the oracle
                dev = st0.clone()
                clo, chi = lo, hi
                for t in range(2):
                    out = outs()
                    eng.step_host_resident(dev, mk(torch.from_numpy(acts[t])), out, uniforms=mk(torch.from_numpy(unis[t])))
                    want = ora.step(clo, chi, acts[t].astype(np.int64), unis[t])
                    ns = out[0].numpy().view(np.uint64)
                    glo, ghi = (ns, np.zeros_like(ns)) if eng.words == 1 else (ns[:, 0], ns[:, 1])
                    assert np.array_equal(glo, want["next_lo"]) and np.array_equal(ghi, want["next_hi"]), (name, B, pinned, t)
                    assert np.array_equal(out[1].numpy().view(np.uint64), want["reward"].view(np.uint64))
                    assert np.array_equal(out[2].numpy().view(np.uint64), want["prob"].view(np.uint64))
                    assert np.array_equal(out[3].numpy().astype(np.uint8), want["done"])
                    assert np.array_equal(out[4].numpy().astype(np.uint8), want["collision"])
                    assert torch.equal(dev.cpu(), out[0]), "resident states == returned next states"
                    clo, chi = want["next_lo"], want["next_hi"]
                    if t == 0:
                        assert int(want["collision"].sum()) > 0 and int(want["done"].sum()) > int(want["collision"].sum())
                # device Philox draws with auto-reset: resident == stateless host step, default and compact layouts
                for compact in (False, True):
                    dev = st0.clone()
                    host = mk(st0.cpu())
                    for t in range(2):
                        a = mk(torch.from_numpy(acts[t]))
                        got, want = outs(compact), outs(compact)
                        eng.step_host_resident(dev, a, got, seed=77, step_index=t, env_offset=5, auto_reset=True, compact=compact)
                        eng.step_host(host, a, want, seed=77, step_index=t, env_offset=5, auto_reset=True, compact=compact)
                        for x, y in zip(got, want):
                            assert torch.equal(x, y), (name, B, pinned, compact, t)
                        assert torch.equal(dev.cpu(), got[0])
                        host = want[0]
    # the resident states must be device memory of the context's GPU
    from gym_mapf_b200 import _native
    bad = torch.zeros(eng.state_shape(600), dtype=torch.int64).pin_memory()
    with pytest.raises(ValueError):
        eng.step_host_resident(bad, torch.from_numpy(acts[0]), outs())
    rc = _native.lib().mapf_step_host_resident(eng._h, bad.data_ptr(), acts[0].ctypes.data, 600, None, 0, 0, 0, 0,
                                               *(t.data_ptr() for t in outs()))
    assert rc == _native.MAPF_ERR_INVALID


def test_share_sm_and_pools_give_identical_results(torch):
    """MAPF_OPT_SHARE_SM only changes the launch shape; two pools stepped on two streams reproduce the single launch."""
    env = _env("room-32-32-4", 1, 4)
    eng = env.engine
    B = 1 << 17
    rng = np.random.default_rng(2)
    st = eng.encode(torch.from_numpy(rng.integers(0, eng.L, (B, eng.n)).astype(np.int32)).cuda())
    ac = torch.from_numpy(rng.integers(0, eng.nA, B).astype(np.int32)).cuda()
    want = eng.step(st, ac, seed=3, step_index=4, env_offset=1000, auto_reset=True)
    got = eng.step(st, ac, seed=3, step_index=4, env_offset=1000, auto_reset=True, share_sm=True)
    for a, b in zip(want, got):
        assert torch.equal(a, b)
    out = tuple(torch.empty_like(t) for t in want)
    H = B // 2
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    for k, s in enumerate(streams):
        with torch.cuda.stream(s):
            sl = slice(k * H, (k + 1) * H)
            eng.step(st[sl], ac[sl], seed=3, step_index=4, env_offset=1000 + k * H, auto_reset=True, share_sm=True,
                     out=tuple(t[sl] for t in out))
    torch.cuda.synchronize()
    for a, b in zip(want, out):
        assert torch.equal(a, b)


def test_count_scan_without_row_len_at_chunk_boundaries(torch):
    """row_ptr from the fused count + scan called with row_len = NULL (compact lengths in the scratch) equals the cumulative
    sum of mapf_count_rows at sizes around the chunk edges, for one- and two-word states, batch and slab mode, and with an
    only 8-byte aligned row_ptr."""
    import ctypes as C
    from gym_mapf_b200._native import _ptr, check, lib
    sizes = (1, 2, 511, 2047, 2048, 2049, 8191, 8192, 8193, 3 * 8192 + 77, 1_000_003)
    for name, scen, n, soc in (("empty-32-32", 1, 2, True), ("room-32-32-4", 1, 4, True), ("room-64-64-8", 1, 8, False)):
        env = _env(name, scen, n, soc)
        eng = env.engine
        for B in sizes:
            rng = np.random.default_rng(B + n)
            cells = torch.from_numpy(rng.integers(0, min(eng.L, 200), (B, eng.n)).astype(np.int32)).cuda()
            st = eng.encode(cells)
            ac = torch.from_numpy(rng.integers(0, eng.nA, B).astype(np.int32)).cuda()
            row_len = torch.empty(B, dtype=torch.int64, device="cuda")
            check(lib().mapf_count_rows(eng._h, _ptr(st), _ptr(ac), B, _ptr(row_len), eng._stream()))
            want = torch.zeros(B + 1, dtype=torch.int64, device="cuda")
            want[1:] = torch.cumsum(row_len, 0)
            for shift in (0, 1):
                buf = torch.full((B + 4,), -7, dtype=torch.int64, device="cuda")
                row_ptr = buf[shift:shift + B + 1]
                scratch = torch.empty(int(lib().mapf_scan_scratch_bytes(B)) // 8 + 1, dtype=torch.int64, device="cuda")
                check(lib().mapf_count_scan_rows(eng._h, _ptr(st), _ptr(ac), B, None, _ptr(row_ptr), _ptr(scratch),
                                                 eng._stream()))
                assert torch.equal(row_ptr, want), (name, B, shift)
                assert int(buf[shift + B + 1]) == -7 and (shift == 0 or int(buf[0]) == -7)
        # slab mode: n_states x nA rows
        n_states = max(1, 20000 // eng.nA + 1)
        s_begin = eng.s0
        B = n_states * eng.nA
        sb = (C.c_uint64 * 2)(s_begin & ((1 << 64) - 1), s_begin >> 64)
        row_len = torch.empty(B, dtype=torch.int64, device="cuda")
        check(lib().mapf_count_range(eng._h, C.byref(sb), n_states, _ptr(row_len), eng._stream()))
        row_ptr = eng.table_range(s_begin, n_states)[0]
        assert torch.equal(row_ptr[1:], torch.cumsum(row_len, 0)) and int(row_ptr[0]) == 0
