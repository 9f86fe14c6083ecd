"""BASELINE.json configs[2]: the joint-state transition table of a 6-agent maze, sharded over the GPUs of one box.
The full table (L**6 = 2.4e17 states x 15 625 actions) cannot exist; each rank streams ONE slab of consecutive
states [s_g, s_g + N) x all actions through a reused record buffer (count -> scan -> expand -> checksum per chunk),
keeps only counts and checksums, and the ranks' eight words are gathered with NCCL at the end (no collective on the
data path).

    python tools/table_multi_gpu.py [--states 4096] [--chunk 32]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/table_multi_gpu.py --states 65536
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from gym_mapf_b200 import sharding  # noqa: E402
from gym_mapf_b200._native import _ptr, check, lib  # noqa: E402
from gym_mapf_b200.envs.mapf_env import OptimizationCriteria  # noqa: E402
from gym_mapf_b200.envs.utils import create_mapf_env  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--map", default="maze-32-32-4")
    ap.add_argument("--scen", type=int, default=10)
    ap.add_argument("--agents", type=int, default=6)
    ap.add_argument("--states", type=int, default=4096, help="consecutive states per GPU")
    ap.add_argument("--chunk", type=int, default=32, help="states per launch sequence")
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    env = create_mapf_env(args.map, args.scen, args.agents, 0.2, -1000.0, 100.0, -1.0, OptimizationCriteria.Makespan, device=local)
    eng = env.engine
    nA = int(eng.nA)
    # slab of rank g: offset g * floor(nS / world) moved onto the start state's neighbourhood so that rows are not
    # all terminal (consecutive states sweep agent 0's cell, then agent 1's, ...)
    s_begin = (eng.s0 + rank * (eng.nS // max(world, 1))) % (eng.nS - args.states)
    rows_per_chunk = args.chunk * nA
    row_len = torch.empty(rows_per_chunk, dtype=torch.int64, device=dev)
    row_ptr = torch.empty(rows_per_chunk + 1, dtype=torch.int64, device=dev)
    scratch = torch.empty(int(lib().mapf_scan_scratch_bytes(rows_per_chunk)) // 8 + 1, dtype=torch.int64, device=dev)
    cap = rows_per_chunk * int(eng.max_row_len) // 2 + 1024  # record buffer, reused by every chunk
    ns, prob, reward, flags = eng._alloc_records(cap)
    words = torch.zeros(8, dtype=torch.int64, device=dev)
    stream = eng._stream()

    def sweep(n_states, base_records):
        done_records = base_records
        for c0 in range(0, n_states, args.chunk):
            n = min(args.chunk, n_states - c0)
            s = s_begin + c0
            sb = (C.c_uint64 * 2)(s & ((1 << 64) - 1), s >> 64)
            check(lib().mapf_count_range(eng._h, C.byref(sb), n, _ptr(row_len), stream))
            check(lib().mapf_scan_rows(eng._h, _ptr(row_len), n * nA, _ptr(row_ptr), _ptr(scratch), stream))
            total = int(row_ptr[n * nA].item())
            if total > cap:
                raise SystemExit("record buffer too small: %d > %d (lower --chunk)" % (total, cap))
            check(lib().mapf_expand_range(eng._h, C.byref(sb), n, _ptr(row_ptr), _ptr(ns), _ptr(prob), _ptr(reward),
                                          _ptr(flags), stream))
            check(lib().mapf_checksum(eng._h, total, done_records, _ptr(ns), _ptr(prob), _ptr(reward), _ptr(flags),
                                      _ptr(words), stream))
            done_records += total
        return done_records

    sweep(min(args.states, 2 * args.chunk), 0)  # warm-up
    words.zero_()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    records = sweep(args.states, 0)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=dev)
    rec = torch.tensor([records], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(rec, op=dist.ReduceOp.SUM)
    per_rank = sharding.gather_words(words)
    if rank == 0:
        secs, total = float(t.item()), int(rec.item())
        print(json.dumps({"case": "configs[2] table slabs %s scen %d n=%d" % (args.map, args.scen, args.agents),
                          "n_gpus": world, "states_per_gpu": args.states, "rows": world * args.states * nA,
                          "records": total, "seconds": secs, "records_per_s": total / secs,
                          "table_bytes_streamed": total * 25,
                          "per_gpu_checksums": [[int(x) for x in w] for w in per_rank],
                          "combined": sharding.combine(per_rank)}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
