#!/usr/bin/env python
"""bench.py -- joint transitions/s of the gym-mapf hot path on 1..8 B200 (one process per GPU).

Workload (BASELINE.json configs[1], SURVEY.md 8d "C2"): room-32-32-4 scen 1, 4 agents, fail_prob 0.2, SoC,
batched step over 2**20 parallel envs PER GPU (weak scaling; shards are independent, no data-path collective).

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # our arm
    python bench.py --impl reference [--steps K] [--warmup W]           # the reference's CPU path (oracle port)

A "step" = one launch of the step kernel over this rank's 2**20 envs: it reads 8 B state + 4 B action and writes
8 B next state + 8 B reward + 8 B probability + 1 B done + 1 B collision per env (38 B), drawing the
slip uniforms on the device (Philox4x32-10).  Steps cycle over a ring of pre-filled (state, action, output) slots
whose total footprint (>= 8x the 126 MB L2) keeps every timed launch reading from and writing to HBM.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MAP, SCEN, N_AGENTS, FAIL_PROB = "room-32-32-4", 1, 4, 0.2
R_CLASH, R_GOAL, R_LIVING = -1000.0, 100.0, -1.0
ENVS_PER_GPU = 1 << 20
STEP_BYTES = 38           # 2W + 22 with W = 8
RING_SLOTS = 32           # 32 x (12 MB in + 26 MB out) = 1.2 GB >> 126 MB L2
NCU_DRAM_BYTES_PER_ENV = (100711936 + 162370048) / (1 << 23)   # profiles/r01_h_step_8m_raw.csv
METRIC = "joint transitions/sec"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples of one GPU while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.12)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6 or not (t0 - 0.05 <= ts <= t1 + 0.1):
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, val in zip(names, f[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_env(device=None):
    from gym_mapf_b200.envs.mapf_env import OptimizationCriteria
    from gym_mapf_b200.envs.utils import create_mapf_env
    return create_mapf_env(MAP, SCEN, N_AGENTS, FAIL_PROB, R_CLASH, R_GOAL, R_LIVING, OptimizationCriteria.SoC,
                           device=device)


def oracle_env(env):
    from oracle import c_oracle
    rows = ["".join("@" if v else "." for v in r) for r in env.grid.obstacles]
    return c_oracle.COracle(rows, env.n_agents, env.agents_goals, FAIL_PROB, R_CLASH, R_GOAL, R_LIVING, True)


# ------------------------------------------------------------------------------------------------------------
# CPU legs (the only places bench.py touches oracle/)
# ------------------------------------------------------------------------------------------------------------
def cpu_c_port(env, seconds=4.0):
    """C oracle, all host threads, on the bench batch (2**20 envs at the start state, random actions/uniforms)."""
    import numpy as np
    ora = oracle_env(env)
    cores = os.cpu_count() or 1
    B = ENVS_PER_GPU
    rng = np.random.default_rng(1)
    lo = np.full(B, env.s, dtype=np.uint64)
    hi = np.zeros(B, np.uint64)
    actions = rng.integers(0, env.nA, B).astype(np.int64)
    uniforms = rng.random((B, env.n_agents))
    ora.step(lo, hi, actions, uniforms, threads=cores)  # warm-up
    t0, reps = time.perf_counter(), 0
    while time.perf_counter() - t0 < seconds:
        out = ora.step(lo, hi, actions, uniforms, threads=cores)
        lo = np.where(out["done"] == 1, np.uint64(env.s), out["next_lo"])
        reps += 1
    dt = time.perf_counter() - t0
    return B * reps / dt, cores, "%d x 2**20-env batches, C port of the reference, %d threads" % (reps, cores)


def _py_worker(args):
    seconds, seed = args
    import numpy as np
    from oracle import mapf_oracle
    env = make_env()
    rows = ["".join("@" if v else "." for v in r) for r in env.grid.obstacles]
    spec = mapf_oracle.OracleSpec(rows, env.n_agents, env.agents_starts, env.agents_goals, FAIL_PROB, R_CLASH, R_GOAL,
                                  R_LIVING, True)
    rng = np.random.default_rng(seed)
    s, n = spec.s0, 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        for _ in range(200):
            s, _, done, _, _, _ = spec.step(s, int(rng.integers(0, spec.nA)), rng.random(spec.n))
            if done:
                s = spec.s0
        n += 200
    return n, time.perf_counter() - t0


def cpu_python_port(seconds=4.0, pool=None):
    """Pure-Python port (the reference itself is pure Python), one process per host core."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    own = pool is None
    if own:
        pool = mp.get_context("spawn").Pool(cores)
    try:
        res = pool.map(_py_worker, [(seconds, 100 + i) for i in range(cores)])
    finally:
        if own:
            pool.close()
            pool.join()
    total = sum(r[0] for r in res)
    dt = max(r[1] for r in res)
    return total / dt, cores, "%d sequential env-steps in %d processes, pure-Python port, %.2f s window" % (
        total, cores, seconds)


def run_reference(args):
    """The reference's own CPU implementation of the path on the host cores.  gym-mapf is pure Python (nothing to
    compile into oracle/_ref), so this times the pure-Python port (oracle/mapf_oracle.py, pinned to the reference by
    tests/golden) with one process per core; a step = one bounded window of sequential env-steps in every process,
    sized so that the whole run stays within about two minutes."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_all = time.time()
    env = make_env()
    n = max(1, args.steps + args.warmup)
    per_step = max(0.05, min(6.0, 100.0 / n))
    cores = os.cpu_count() or 1
    vals, sample = [], ""
    with mp.get_context("spawn").Pool(cores) as pool:
        pool.map(_py_worker, [(0.05, i) for i in range(cores)])  # start the workers before anything is timed
        for i in range(args.warmup + args.steps):
            v, cores, sample = cpu_python_port(per_step, pool)
            if i >= args.warmup:
                vals.append(v)
    value = sum(vals) / len(vals)
    c_val, c_cores, c_sample = cpu_c_port(env, 3.0)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "transitions/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int64+f64", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": "transitions/s", "cores": cores, "kind": "port",
                             "sample": sample + " per step (the reference is pure Python and cannot be compiled into "
                                                "oracle/_ref; this is oracle/mapf_oracle.py)",
                             "c_port_value": c_val, "c_port_sample": c_sample},
            "e2e": {"value": value, "unit": "transitions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.time() - t_all}
    print(json.dumps(line), flush=True)


def workload_config(n_gpus):
    return {"workload": "C2: %s scen %d, %d agents, fail_prob %.1f, SoC; batched step over 2**20 envs per GPU" % (
        MAP, SCEN, N_AGENTS, FAIL_PROB), "envs_per_gpu": ENVS_PER_GPU, "global_envs": ENVS_PER_GPU * n_gpus,
        "parallelism": "env-sharded x%d, no data-path collective" % n_gpus, "sampling": "device Philox4x32-10",
        "l2": "ring of %d input/output slots (%.1f GB per GPU) cycled so every launch misses the 126 MB L2" % (
            RING_SLOTS, RING_SLOTS * ENVS_PER_GPU * STEP_BYTES / 1e9)}


# ------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL writes its version / debug lines to stdout when NCCL_DEBUG is set: keep stdout for the one JSON line by
        # pointing file descriptor 1 at stderr while the communicator comes up (first collective included)
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    env = make_env(device=local)
    eng = env.engine
    B, K, W = ENVS_PER_GPU, args.steps, args.warmup
    from gym_mapf_b200 import sharding
    shard = sharding.env_shard(world * B, world, rank)   # contiguous slice of the global env batch, no exchange step
    assert shard.count == B
    env_offset = shard.begin
    seed = 20261018

    # ---- ring of pre-filled slots: slot j holds the states after j random steps from reset (auto-reset on)
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    states = [eng.states_from_ints([eng.s0]).expand(B).contiguous()]
    actions = [torch.randint(0, env.nA, (B,), generator=g, device=dev, dtype=torch.int32) for _ in range(RING_SLOTS)]
    outs = []
    for j in range(RING_SLOTS):
        out = (eng.new_states(B), torch.empty(B, dtype=torch.float64, device=dev),
               torch.empty(B, dtype=torch.float64, device=dev), torch.empty(B, dtype=torch.bool, device=dev),
               torch.empty(B, dtype=torch.bool, device=dev))
        outs.append(out)
        eng.step(states[j], actions[j], seed=seed, step_index=j, env_offset=env_offset, auto_reset=True, out=out)
        if j + 1 < RING_SLOTS:
            states.append(out[0].clone())
    torch.cuda.synchronize()

    def one_step(i):
        j = i % RING_SLOTS
        eng.step(states[j], actions[j], seed=seed, step_index=1000 + i, env_offset=env_offset, auto_reset=True,
                 out=outs[j])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(W, 3)):
        one_step(i)
    barrier()

    # The K steps of one repetition are captured once into a CUDA graph (K kernel nodes joined by programmatic
    # dependent-launch edges) and replayed: the host launch path (~13 us per call through Python/ctypes) would
    # otherwise be as long as the kernel itself.  --no-graph times plain stream launches instead.
    graph = None
    if not args.no_graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            one_step(0)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for i in range(K):
                one_step(i)
        barrier()

    # ---- timed region: exactly K steps per repetition, CUDA events on the launching stream; repetitions until the
    # GPU has been busy long enough for nvidia-smi to sample clocks under load.  The median repetition is reported.
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.15)
    reps_ms = []
    t_wall0 = time.time()
    while True:
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if graph is not None:
            graph.replay()
        else:
            for i in range(K):
                one_step(i)
        e1.record()
        barrier()
        reps_ms.append(e0.elapsed_time(e1))
        if args.reps > 0 and len(reps_ms) >= args.reps:
            break
        if args.reps == 0 and len(reps_ms) >= 3 and time.time() - t_wall0 > args.min_seconds:
            break
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1)
    reps_sorted = sorted(reps_ms)
    ms_total = reps_sorted[len(reps_sorted) // 2]
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / K
    value = world * B * K / (ms_total * 1e-3)

    # ---- per-shard counts + checksums, gathered with NCCL (the only collective; outside the timed region)
    j = (K - 1) % RING_SLOTS
    ns, reward, prob, done, coll = outs[j]
    flags = done.to(torch.uint8) + 2 * coll.to(torch.uint8)
    cs = eng.checksum(ns, prob, reward, flags, index_base=env_offset)
    shard_sums = [[int(x) for x in words] for words in sharding.gather_words(cs)]

    # ---- end to end through the host-buffer C-ABI call: pinned host inputs -> H2D -> step -> D2H, every step
    e2e = None
    if not args.no_e2e:
        # the pinned buffers are allocated (first-touched) on the GPU's own NUMA node; undone after the leg so that the
        # cpu_baseline leg sees every host core again
        prev_affinity = None if os.environ.get("BENCH_NO_NUMA_BIND") else sharding.bind_host_to_device(local)
        hs = torch.empty(eng.state_shape(B), dtype=torch.int64).pin_memory()
        hs.copy_(states[0])
        ha = actions[0].cpu().pin_memory()
        hout = (torch.empty(eng.state_shape(B), dtype=torch.int64).pin_memory(),
                torch.empty(B, dtype=torch.float64).pin_memory(), torch.empty(B, dtype=torch.float64).pin_memory(),
                torch.empty(B, dtype=torch.bool).pin_memory(), torch.empty(B, dtype=torch.bool).pin_memory())
        n_e2e = max(3, min(K, 30))
        for i in range(3):
            eng.step_host(hs, ha, hout, seed=seed, step_index=i, env_offset=env_offset, auto_reset=True)
        barrier()
        t0 = time.perf_counter()
        for i in range(n_e2e):
            eng.step_host(hs, ha, hout, seed=seed, step_index=10 + i, env_offset=env_offset, auto_reset=True)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * B * n_e2e / float(dt.item()), "unit": "transitions/s",
               "h2d_bytes_per_step": B * 12 * world, "d2h_bytes_per_step": B * 26 * world, "steps": n_e2e,
               "api": "mapf_step_host (C ABI, pinned host buffers in and out)",
               "host_numa_bind": prev_affinity is not None}
        if prev_affinity is not None:
            os.sched_setaffinity(0, prev_affinity)

    if rank == 0:
        peak, peak_src = load_peaks()
        achieved = B * STEP_BYTES / (ms_per_step * 1e-3) / 1e9
        line = {"metric": METRIC, "value": value, "unit": "transitions/s", "n_gpus": world, "steps": K, "warmup": max(W, 3),
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "int64+f64", "data": "synthetic", "config": workload_config(world),
                "reps": len(reps_ms), "rep_ms_min_med_max": [reps_sorted[0], reps_sorted[len(reps_sorted) // 2],
                                                             reps_sorted[-1]],
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": NCU_DRAM_BYTES_PER_ENV * B,
                             "traffic_source": "ncu --set full, profiles/r01_h_step_8m_raw.csv: dram__bytes_read 100.7 MB "
                                               "+ dram__bytes_write 162.4 MB for a 2**23-env launch = 31.4 B/env (reads = "
                                               "the algorithmic 12 B/env; 19.4 of the 26 B/env written had reached DRAM "
                                               "when the kernel ended, the rest was still in the 126 MB L2)",
                             "peak_source": peak_src,
                             "kernel": "k_step<4,smem move table>", "bytes_per_unit": STEP_BYTES,
                             "units_per_launch": B},
                "e2e": e2e, "gpu_launches": K, "clocks": clocks,
                "launch": ("one CUDA graph of K step kernels (programmatic dependent launch)" if graph is not None
                           else "K stream launches"),
                "shard_checksums": {"keys": ["count", "n_collision", "n_done", "sum_next_lo", "sum_next_hi",
                                             "sum_prob_bits", "sum_reward_bits", "ordered"], "per_gpu": shard_sums}}
        if world == 1 and not args.no_cpu:
            c_val, c_cores, c_sample = cpu_c_port(env, 4.0)
            p_val, p_cores, p_sample = cpu_python_port(3.0)
            line["cpu_baseline"] = {"value": c_val, "unit": "transitions/s", "cores": c_cores, "kind": "port",
                                    "sample": c_sample, "python_port_value": p_val, "python_port_sample": p_sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reps", type=int, default=0, help="timed repetitions of the K-step region (0 = auto, >= ~1.5 s)")
    ap.add_argument("--min-seconds", type=float, default=1.5)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg (profiling runs)")
    ap.add_argument("--no-graph", action="store_true", help="plain stream launches instead of a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
