"""Sharded value iteration on N GPUs (one process per GPU, torchrun): each rank sweeps its contiguous shard of the
state space with k_backup + k_greedy; the new values are exchanged either by an NCCL all-gather after the sweep
("nccl") or INSIDE the greedy kernel, which writes every value straight into all ranks' value vectors through
peer-mapped symmetric memory over NVLink ("fused").  Both must end bit-identical to a single-GPU run.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/vi_multi_gpu.py [--sweeps 30]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from gym_mapf_b200 import sharding  # noqa: E402
from gym_mapf_b200.envs.mapf_env import OptimizationCriteria  # noqa: E402
from gym_mapf_b200.envs.utils import create_mapf_env  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--map", default="empty-32-32")
    ap.add_argument("--agents", type=int, default=2)
    ap.add_argument("--sweeps", type=int, default=30)
    ap.add_argument("--gamma", type=float, default=0.99)
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    env = create_mapf_env(args.map, 1, args.agents, 0.2, -1000.0, 100.0, -1.0, OptimizationCriteria.Makespan, device=local)
    eng = env.engine
    nS = int(eng.nS)
    shards = [sharding.split_range(nS, world, r) for r in range(world)]
    mine = shards[rank]
    Q = torch.empty((mine.count, eng.nA), dtype=torch.float64, device=dev)

    def timed_sweeps(step):
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.sweeps):
            step()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.sweeps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- exchange by NCCL all-gather after the sweep
    state = {"V": torch.zeros(nS, dtype=torch.float64, device=dev)}

    def step_nccl():
        eng.backup_range(mine.begin, mine.count, state["V"], args.gamma, out=Q)
        v, _pi = eng.greedy(Q)
        state["V"] = sharding.all_gather_values(v, shards)
    ms_nccl = timed_sweeps(step_nccl)
    state["V"] = torch.zeros(nS, dtype=torch.float64, device=dev)
    for _ in range(args.sweeps):
        step_nccl()
    v_nccl = state["V"].clone()

    # ---- exchange fused into the greedy kernel: peer stores into symmetric memory
    ms_fused, v_fused, fused_note = None, None, None
    if world > 1:
        try:
            import torch.distributed._symmetric_memory as symm
            bufs, hdls = [], []
            for _ in range(2):
                t = symm.empty(nS, dtype=torch.float64, device=dev)
                bufs.append(t)
                hdls.append(symm.rendezvous(t, dist.group.WORLD))
            cur = {"i": 0}

            def reset():
                for t in bufs:
                    t.zero_()
                torch.cuda.synchronize()
                dist.barrier()
                cur["i"] = 0

            def step_fused():
                i = cur["i"]
                eng.backup_range(mine.begin, mine.count, bufs[i], args.gamma, out=Q)
                eng.greedy_bcast(Q, mine.begin, hdls[1 - i].buffer_ptrs)
                hdls[1 - i].barrier()  # every rank's stores into every vector have landed before anyone reads them
                cur["i"] = 1 - i
            reset()
            ms_fused = timed_sweeps(step_fused)
            reset()
            for _ in range(args.sweeps):
                step_fused()
            torch.cuda.synchronize()
            v_fused = bufs[cur["i"]].clone()
        except Exception as e:  # noqa: BLE001 - report, do not hide
            fused_note = "%s: %s" % (type(e).__name__, e)

    # ---- single-GPU reference run on rank 0 (same sweeps, whole state space)
    same_nccl = same_fused = None
    if rank == 0:
        V1 = torch.zeros(nS, dtype=torch.float64, device=dev)
        Q1 = torch.empty((nS, eng.nA), dtype=torch.float64, device=dev)
        for _ in range(args.sweeps):
            eng.backup_range(0, nS, V1, args.gamma, out=Q1)
            V1, _ = eng.greedy(Q1)
        same_nccl = bool(torch.equal(V1.view(torch.int64), v_nccl.view(torch.int64)))
        if v_fused is not None:
            same_fused = bool(torch.equal(V1.view(torch.int64), v_fused.view(torch.int64)))
        print(json.dumps({"case": "sharded value iteration %s n=%d" % (args.map, args.agents), "n_gpus": world,
                          "states": nS, "rows_per_sweep": nS * int(eng.nA), "sweeps": args.sweeps,
                          "ms_per_sweep_nccl_allgather": ms_nccl, "ms_per_sweep_fused_peer_stores": ms_fused,
                          "bit_identical_to_single_gpu": {"nccl": same_nccl, "fused": same_fused},
                          "fused_note": fused_note, "exchange_bytes_per_sweep_per_rank": mine.count * 8 * (world - 1)}),
              flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
