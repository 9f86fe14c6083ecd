"""The N>1 path on CPU: world_size-2 `gloo` processes shard a table build and an env batch exactly as the GPU ranks do
(gym_mapf_b200.sharding), produce their shards with the C oracle standing in for the kernels, gather the eight checksum
words and must reproduce the reference-generated whole-job checksums (tests/golden/full_c1_*.npz)."""
import os
import socket

import numpy as np
import pytest

import golden_util as G
from gym_mapf_b200 import sharding


def test_split_range_covers_without_overlap():
    for total in (0, 1, 7, 8, 4096, 10 ** 30 + 3):
        for world in (1, 2, 3, 8):
            shards = [sharding.split_range(total, world, r, begin=5) for r in range(world)]
            assert shards[0].begin == 5
            assert sum(s.count for s in shards) == total
            for a, b in zip(shards, shards[1:]):
                assert a.begin + a.count == b.begin
            assert max(s.count for s in shards) - min(s.count for s in shards) <= 1
    with pytest.raises(ValueError):
        sharding.split_range(10, 2, 2)


def test_combine_is_mod_2_64():
    a = np.array([1, 2, 3, (1 << 64) - 1, 0, 5, 6, 7], dtype=np.uint64)
    b = np.array([1, 0, 0, 2, 0, 0, 0, (1 << 64) - 7], dtype=np.uint64)
    got = sharding.combine([a, b])
    assert got["count"] == 2 and got["sum_next_lo"] == 1 and got["ordered"] == 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _table_worker(rank, world, port, name, q):
    import torch
    import torch.distributed as dist
    from engine_util import make_oracle
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        spec, _ = G.load(name)
        ora = make_oracle(spec)
        nS, nA = int(spec["nS"]), int(spec["nA"])
        sh = sharding.table_shard(0, nS, world, rank)
        s = np.repeat(np.arange(sh.begin, sh.begin + sh.count, dtype=np.uint64), nA)
        a = np.tile(np.arange(nA, dtype=np.int64), sh.count)
        rows = ora.rows(s, np.zeros_like(s), a)
        # phase 1: record counts, so that every shard knows the global index of its first record
        n_rec = torch.tensor([int(rows["row_ptr"][-1])], dtype=torch.int64)
        counts = [torch.zeros_like(n_rec) for _ in range(world)]
        dist.all_gather(counts, n_rec)
        base = sharding.record_index_base(int(c) for c in counts[:rank])
        cs = G.checksums(rows["next_lo"], rows["next_hi"], G.f64_to_bits(rows["prob"]), G.f64_to_bits(rows["reward"]),
                         rows["done"], rows["collision"])
        # order-sensitive word with the GLOBAL record index (what mapf_checksum does with index_base)
        with np.errstate(over="ignore"):
            tag = (rows["next_lo"] + np.uint64(1) + np.uint64(2) * rows["collision"].astype(np.uint64)
                   + np.uint64(4) * rows["done"].astype(np.uint64))
            idx = np.arange(1, len(tag) + 1, dtype=np.uint64) + np.uint64(base)
            cs["ordered"] = int(np.sum(idx * tag, dtype=np.uint64))
        words = torch.from_numpy(np.array([cs[k] for k in sharding.CHECKSUM_KEYS], dtype=np.uint64).view(np.int64))
        per_rank = sharding.gather_words(words)
        if rank == 0:
            q.put(sharding.combine(per_rank))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["full_c1_makespan", "full_c1_soc"])
def test_sharded_table_checksums_match_reference(name):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    world, port = 2, _free_port()
    procs = [ctx.Process(target=_table_worker, args=(r, world, port, name, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = q.get()
    _, want = G.load(name)
    for k in sharding.CHECKSUM_KEYS:
        assert got[k] == int(want[k]), k


def _step_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from engine_util import make_oracle
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        spec, _ = G.load("rows_c2")
        ora = make_oracle(spec)
        B = 4001  # odd on purpose: the shards differ in size
        rng = np.random.default_rng(11)  # same global batch on every rank; each rank takes its slice
        lo = rng.integers(0, int(spec["nS"]), B).astype(np.uint64)
        act = rng.integers(0, int(spec["nA"]), B).astype(np.int64)
        uni = rng.random((B, spec["n_agents"]))
        sh = sharding.env_shard(B, world, rank)
        sl = slice(sh.begin, sh.begin + sh.count)
        out = ora.step(lo[sl], np.zeros(sh.count, np.uint64), act[sl], uni[sl])
        cs = G.checksums(out["next_lo"], out["next_hi"], G.f64_to_bits(out["prob"]), G.f64_to_bits(out["reward"]),
                         out["done"], out["collision"])
        with np.errstate(over="ignore"):
            tag = (out["next_lo"] + np.uint64(1) + np.uint64(2) * out["collision"].astype(np.uint64)
                   + np.uint64(4) * out["done"].astype(np.uint64))
            idx = np.arange(1, sh.count + 1, dtype=np.uint64) + np.uint64(sh.begin)  # index_base = env_offset
            cs["ordered"] = int(np.sum(idx * tag, dtype=np.uint64))
        words = torch.from_numpy(np.array([cs[k] for k in sharding.CHECKSUM_KEYS], dtype=np.uint64).view(np.int64))
        per_rank = sharding.gather_words(words)
        if rank == 0:
            whole = ora.step(lo, np.zeros(B, np.uint64), act, uni)
            want = G.checksums(whole["next_lo"], whole["next_hi"], G.f64_to_bits(whole["prob"]),
                               G.f64_to_bits(whole["reward"]), whole["done"], whole["collision"])
            q.put((sharding.combine(per_rank), want))
    finally:
        dist.destroy_process_group()


def test_sharded_step_checksums_equal_unsharded():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    world, port = 2, _free_port()
    procs = [ctx.Process(target=_step_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got, want = q.get()
    assert got == want


def _vi_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from engine_util import make_oracle
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        spec, _ = G.load("full_c1_makespan")
        ora = make_oracle(spec)
        nS, nA = int(spec["nS"]), int(spec["nA"])

        def sweep(begin, count, V, gamma):  # the C oracle stands in for k_backup + k_greedy
            s = np.repeat(np.arange(begin, begin + count, dtype=np.uint64), nA)
            a = np.tile(np.arange(nA, dtype=np.int64), count)
            Q = ora.backup(s, np.zeros_like(s), a, V.numpy(), gamma).reshape(count, nA)
            return torch.from_numpy(Q.max(axis=1)), torch.from_numpy(Q.argmax(axis=1).astype(np.int32))

        V, pi, iters = sharding.sharded_value_iteration(sweep, nS, world, rank, gamma=1.0, eps=1e-3, max_iter=60)
        if rank == 0:
            V1, pi1, it1 = sharding.sharded_value_iteration(sweep, nS, 1, 0, gamma=1.0, eps=1e-3, max_iter=60)
            q.put((iters, it1, bool(np.array_equal(V.numpy().view(np.uint64), V1.numpy().view(np.uint64))),
                   float(V[int(spec["s0"])])))
    finally:
        dist.destroy_process_group()


def test_sharded_value_iteration_equals_single_rank():
    """The exchange step of the sharded sweep (one all-gather of the new values per iteration) reproduces the
    single-rank run bit for bit, with uneven shards (world_size 3 over 4096 states)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    world, port = 3, _free_port()
    procs = [ctx.Process(target=_vi_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    iters, it1, same, v0 = q.get()
    assert iters == it1 and same and v0 > 0


def test_segment_shard_covers_the_heterogeneous_batch():
    """sharding.segment_shard: the shards of a concatenation of per-spec segments are disjoint, ordered and complete,
    for even and uneven splits, empty segments and more ranks than envs."""
    from gym_mapf_b200 import sharding
    for counts in ([3000, 1, 2049, 0, 777], [5], [0, 0, 7], [1, 1, 1], [10, 0, 0, 10]):
        total = sum(counts)
        for world in (1, 2, 3, 8, 16):
            seen = []
            at = 0
            for rank in range(world):
                sh, parts = sharding.segment_shard(counts, world, rank)
                assert sh.begin == at and sum(c for _, c in parts) == sh.count
                assert [i for i, _ in parts] == sorted({i for i, _ in parts})     # specs in order, each once
                at += sh.count
                pos = sh.begin
                for i, c in parts:
                    seg_lo = sum(counts[:i])
                    assert seg_lo <= pos and pos + c <= seg_lo + counts[i] and c > 0
                    seen.append((i, pos - seg_lo, c))
                    pos += c
            assert at == total
            per_spec = {}
            for i, off, c in seen:
                assert per_spec.get(i, 0) == off          # consecutive pieces of a segment
                per_spec[i] = off + c
            assert all(per_spec.get(i, 0) == c for i, c in enumerate(counts))
