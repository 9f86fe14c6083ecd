"""PCIe copy bandwidth of the box (pinned host memory) with 1, 2, 4 and 8 GPUs copying AT THE SAME TIME: one direction at
a time, both at once, and in the 12 : 26 byte mix of the end-to-end step (12 B/env in, 26 B/env out) -- the ceiling of
`mapf_step_host`.  One process per GPU (spawned here), a barrier before every measurement, wall-clock of the slowest rank.

    python tools/pcie_peak.py [--gpus 1,2,4,8] [--out profiles/r02_pcie_ceiling.json]
"""
import argparse
import json
import os
import sys
import time


def worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 256 << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_out = torch.ones(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(f_in, f_out, reps=6):
        """copy f_in * n bytes host->device and f_out * n bytes device->host per repetition, concurrently"""
        a, b = int(n * f_in), int(n * f_out)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            if a:
                with torch.cuda.stream(s1):
                    d_in[:a].copy_(h_in[:a], non_blocking=True)
            if b:
                with torch.cuda.stream(s2):
                    h_out[:b].copy_(d_out[:b], non_blocking=True)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return a * reps / dt.item() / 1e9, b * reps / dt.item() / 1e9

    run(1, 1, 2)
    res = {"h2d_gbs_per_gpu": run(1, 0)[0], "d2h_gbs_per_gpu": run(0, 1)[1]}
    i, o = run(1, 1)
    res["both_h2d_gbs_per_gpu"], res["both_d2h_gbs_per_gpu"] = i, o
    i, o = run(12 / 26, 1)
    res["step_mix_h2d_gbs_per_gpu"], res["step_mix_d2h_gbs_per_gpu"] = i, o
    if rank == 0:
        res["gpus"] = world
        res["both_total_gbs"] = (res["both_h2d_gbs_per_gpu"] + res["both_d2h_gbs_per_gpu"]) * world
        res["step_mix_total_gbs"] = (res["step_mix_h2d_gbs_per_gpu"] + res["step_mix_d2h_gbs_per_gpu"]) * world
        res["step_ceiling_env_steps_per_s"] = res["step_mix_total_gbs"] * 1e9 / 38
        q.put(res)
    if world > 1:
        dist.destroy_process_group()


def main():
    import torch
    import torch.multiprocessing as mp
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", default="1,2,4,8")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    have = torch.cuda.device_count()
    out = {}
    ctx = mp.get_context("spawn")
    for k, world in enumerate(int(x) for x in args.gpus.split(",")):
        if world > have:
            continue
        q = ctx.Queue()
        procs = [ctx.Process(target=worker, args=(r, world, 29700 + k, q)) for r in range(world)]
        for p in procs:
            p.start()
        res = q.get(timeout=300)
        for p in procs:
            p.join()
        out[str(world)] = res
        print(json.dumps(res), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    sys.exit(main())
