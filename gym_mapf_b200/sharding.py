"""Sharding of the joint-transition path over the GPUs of one box (one process per GPU).

The path has no exchange step: a table row depends only on (s, a), an env-step only on that env's
state, action and draw.  So a shard is just a contiguous slice -- of the env batch in step mode, of the joint-state
index range in table mode -- and the only communication is a gather of eight 64-bit checksum words per shard after
the work is done (NCCL on the GPU box, gloo in the CPU tests).  Nothing here touches the data path.
"""
import collections

import numpy as np

CHECKSUM_KEYS = ("count", "n_collision", "n_done", "sum_next_lo", "sum_next_hi", "sum_prob_bits", "sum_reward_bits",
                 "ordered")
M64 = (1 << 64) - 1

Shard = collections.namedtuple("Shard", "rank world begin count")


def split_range(total, world, rank, begin=0):
    """Contiguous shard `rank` of `world` over [begin, begin + total): the first `total % world` shards hold one
    more element.  `total` and `begin` may be Python big ints (joint-state indices go up to 2**127)."""
    total, world, rank = int(total), int(world), int(rank)
    if world < 1 or not 0 <= rank < world or total < 0:
        raise ValueError("bad shard request: total=%d world=%d rank=%d" % (total, world, rank))
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return Shard(rank, world, int(begin) + start, base + (1 if rank < extra else 0))


def env_shard(global_envs, world, rank):
    """Step mode: rank owns envs [begin, begin + count) of the global batch; `begin` is the `env_offset` that keys
    its Philox stream, so the union of the shards draws exactly what one GPU would draw for the whole batch."""
    return split_range(global_envs, world, rank)


def segment_shard(counts, world, rank):
    """Heterogeneous batch (MultiMapVecEnv: counts[i] envs of spec i, concatenated): the contiguous shard of rank
    `rank` as (Shard, [(spec index, envs of that spec inside the shard), ...]).  `Shard.begin` is the `env_offset` of
    the rank's MultiMapVecEnv, so the union of the shards draws exactly what one GPU draws for the whole batch."""
    counts = [int(c) for c in counts]
    if any(c < 0 for c in counts):
        raise ValueError("negative env count")
    sh = split_range(sum(counts), world, rank)
    parts, at = [], 0
    for i, c in enumerate(counts):
        lo, hi = max(at, sh.begin), min(at + c, sh.begin + sh.count)
        if hi > lo:
            parts.append((i, hi - lo))
        at += c
    return sh, parts


def table_shard(s_begin, n_states, world, rank):
    """Table mode: rank owns joint states [begin, begin + count) x all actions of the slab [s_begin, s_begin + n_states)."""
    return split_range(n_states, world, rank, begin=s_begin)


def record_index_base(row_counts_before):
    """Index of a shard's first record in the global (s, a, outcome) order = records emitted by lower ranks."""
    return int(sum(int(c) for c in row_counts_before))


def gather_words(words, group=None):
    """All-gather one int64[8] tensor of checksum words per rank -> list of per-rank numpy uint64[8] arrays.
    `words` lives on the device of the group's backend (cuda for nccl, cpu for gloo).  Single process: no-op."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return [words.detach().cpu().numpy().view(np.uint64).copy()]
    bucket = [torch.zeros_like(words) for _ in range(dist.get_world_size(group))]
    dist.all_gather(bucket, words.contiguous(), group=group)
    return [b.detach().cpu().numpy().view(np.uint64).copy() for b in bucket]


def combine(per_rank):
    """Sum the per-shard checksum words mod 2**64.  Every word is additive over disjoint shards (the order-sensitive
    one because each shard weights its records with their GLOBAL index, see mapf_checksum's index_base)."""
    total = [0] * len(CHECKSUM_KEYS)
    for words in per_rank:
        for i, w in enumerate(words):
            total[i] = (total[i] + int(w)) & M64
    return dict(zip(CHECKSUM_KEYS, total))


# ---- sharded value iteration: the one place with a real exchange step ------------------------------------------
def all_gather_values(v_shard, shards, group=None):
    """Every rank contributes the values of its own states; returns the full vector on every rank.  Shards may
    differ in size by one (split_range), so the gather runs on shards padded to the largest and trims afterwards.
    Backend-agnostic (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist
    if len(shards) == 1 or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return v_shard
    width = max(sh.count for sh in shards)
    padded = torch.zeros(width, dtype=v_shard.dtype, device=v_shard.device)
    padded[:v_shard.shape[0]] = v_shard
    out = torch.empty(width * len(shards), dtype=v_shard.dtype, device=v_shard.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * width:r * width + sh.count] for r, sh in enumerate(shards)])


def sharded_value_iteration(backup_and_greedy, n_states, world, rank, gamma=1.0, eps=1e-2, max_iter=1000, device="cpu",
                            group=None):
    """Synchronous value iteration with the state space cut into one contiguous shard per rank.
    `backup_and_greedy(s_begin, count, V, gamma) -> (V_new_shard, policy_shard)` is the per-shard sweep (on the GPU:
    `Engine.backup_range` + `Engine.greedy`).  Per sweep: local backup of the own shard against the full V, then ONE
    all-gather of the new values (8 bytes per state) and one all-reduce of the change.  Returns (V, policy_shard,
    iterations); every rank ends with the same, bit-identical V as a single-rank run."""
    import torch
    import torch.distributed as dist
    shards = [split_range(n_states, world, r) for r in range(world)]
    mine = shards[rank]
    V = torch.zeros(n_states, dtype=torch.float64, device=device)
    policy = None
    it = 0
    for it in range(1, max_iter + 1):
        v_new, policy = backup_and_greedy(mine.begin, mine.count, V, gamma)
        delta = (v_new - V[mine.begin:mine.begin + mine.count]).abs().sum().reshape(1)
        V = all_gather_values(v_new, shards, group)
        if world > 1:
            dist.all_reduce(delta, op=dist.ReduceOp.SUM, group=group)
        if float(delta.item()) <= eps:
            break
    return V, policy, it


def bind_host_to_device(device_index):
    """Host-buffer paths (`mapf_step_host`, pinned staging buffers): run the calling process on the CPU cores next to
    GPU `device_index` (its NUMA node), so that page-locked buffers allocated afterwards are first-touched in the memory
    the GPU reaches without crossing the socket interconnect.  Returns the previous affinity set (pass it to
    `os.sched_setaffinity(0, ...)` to undo), or None when the topology cannot be read -- then nothing is changed."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(visible.split(",")[device_index]) if visible and visible.replace(",", "").isdigit() else device_index
        handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        before = os.sched_getaffinity(0)
        cpus &= before
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return before
    except Exception:  # noqa: BLE001 - no NVML, no permission, unknown topology: leave the affinity alone
        return None
