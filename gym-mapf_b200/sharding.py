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
