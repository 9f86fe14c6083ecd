#!/usr/bin/env python
"""bench.py -- joint transitions/s of the gym-mapf hot path on 1..8 B200 (one process per GPU).

Workload (BASELINE.json configs[1], SURVEY.md 8d "C2"): room-32-32-4 scen 1, 4 agents, fail_prob 0.2, SoC,
batched step over 2**20 parallel envs PER GPU (weak scaling; shards are independent, no data-path collective).

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # our arm
    python bench.py --impl reference [--steps K] [--warmup W]           # the reference's own CPU path on the host cores

A "step" = one pass of the step kernel over this rank's 2**20 envs: it reads 8 B state + 4 B action and writes
8 B next state + 8 B reward + 8 B probability + 1 B done + 1 B collision per env (38 B), drawing the
slip uniforms on the device (Philox4x32-10).  Steps cycle over a ring of pre-filled (state, action, output) slots
whose total footprint (>= 8x the 126 MB L2) keeps every timed launch reading from and writing to HBM.

The rank's envs are held as `--pools` (default 2) independent pools of 2**20 / pools envs, each stepped on its own CUDA
stream with one resident CTA per SM (MAPF_OPT_SHARE_SM): every pool is a chain of launches, each ordered after the
previous launch of ITS pool (programmatic dependent launch), and the two chains share every SM -- while one pool's
launch drains and the next one waits for it, the other pool computes.  One step = `pools` launches; `gpu_launches`
counts them all.  `--pools 1` is the single chain of full-GPU launches (also reported, as `single_chain`).

Besides the headline the line carries `other_configs`: every other BASELINE.json config (C1 full table + 10k steps,
C2 expand / rollout / strong scaling, C3 table slab per rank, C4 2**24-env 128-bit step, C5 agent-count and
conflict-density sweeps, the maps whose move table does not fit shared memory), each with its fraction of the HBM
roofline and a `parity` flag: a sample of what the GPU produced, compared bit for bit with the CPU oracle outside the
timed region.  The headline's own last launch is checked the same way (`parity`).
"""
import argparse
import datetime
import json
import math
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
NCU_DRAM_BYTES_PER_ENV = (100711936 + 162091776) / (1 << 23)   # profiles/r02_step_8m_raw.csv
METRIC = "joint transitions/sec"
MAX_REPS = 20000


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


# ------------------------------------------------------------------------------------------------------------
# CPU legs (the only places bench.py touches oracle/): the cpu_baseline / --impl reference timings, and the
# oracle as the CHECKER of GPU output samples (never as the thing measured on our arm)
# ------------------------------------------------------------------------------------------------------------
def oracle_of(spec):
    from oracle import c_oracle
    return c_oracle.COracle(spec["rows"], spec["n_agents"], spec["goals"], spec["fail_prob"], spec["r_clash"],
                            spec["r_goal"], spec["r_living"], spec["soc"])


def oracle_env(env):
    return oracle_of({"rows": ["".join("@" if v else "." for v in r) for r in env.grid.obstacles],
                      "n_agents": env.n_agents, "goals": env.agents_goals, "fail_prob": FAIL_PROB, "r_clash": R_CLASH,
                      "r_goal": R_GOAL, "r_living": R_LIVING, "soc": True})


def philox4x32_10(ctr, key):
    """numpy Philox4x32-10 (Salmon et al., SC'11): the device-side sampling stream, regenerated on the host so that the
    oracle can replay a Philox-mode launch.  ctr: four uint64 arrays holding 32-bit words; key: two 32-bit ints."""
    import numpy as np
    c = [np.asarray(x, dtype=np.uint64) for x in ctr]
    k0, k1 = np.uint64(key[0]), np.uint64(key[1])
    m32 = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c[0]
        p1 = np.uint64(0xCD9E8D57) * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & m32, p1 >> np.uint64(32), p1 & m32
        c = [(hi1 ^ c[1] ^ k0) & m32, lo1, (hi0 ^ c[3] ^ k1) & m32, lo0]
        k0 = (k0 + np.uint64(0x9E3779B9)) & m32
        k1 = (k1 + np.uint64(0xBB67AE85)) & m32
    return c


def philox_block(env_ids, step_index, block, seed):
    import numpy as np
    n = len(env_ids)
    return philox4x32_10([env_ids & np.uint64(0xFFFFFFFF), env_ids >> np.uint64(32),
                          np.full(n, step_index & 0xFFFFFFFF, np.uint64),
                          np.full(n, (((step_index >> 32) << 8) | block) & 0xFFFFFFFF, np.uint64)],
                         [seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF])


def device_uniforms(env_ids, step_index, seed, n_agents):
    """Agent i's uniform = word i % 4 of block i // 4 at counter (env, step), times 2**-32 (mapf_device.cuh)."""
    import numpy as np
    u = np.zeros((len(env_ids), n_agents))
    for blk in range((n_agents + 3) // 4):
        words = philox_block(env_ids, step_index, blk, seed)
        for q in range(4):
            if blk * 4 + q < n_agents:
                u[:, blk * 4 + q] = words[q].astype(np.float64) * 2.0 ** -32
    return u


def verify_payload(p):
    """-> (ok, description).  The CPU oracle recomputes the sample; every field is compared bit for bit."""
    import numpy as np
    M64 = (1 << 64) - 1
    threads = os.cpu_count() or 1
    kind = p["kind"]
    if kind == "rows":
        ora = oracle_of(p["spec"])
        want = ora.rows(p["s_lo"], p["s_hi"], p["action"], threads=threads)
        ok = (np.array_equal(want["row_ptr"], p["row_ptr"]) and np.array_equal(want["next_lo"], p["next_lo"])
              and np.array_equal(want["next_hi"], p["next_hi"])
              and np.array_equal(want["prob"].view(np.uint64), p["prob"].view(np.uint64))
              and np.array_equal(want["reward"].view(np.uint64), p["reward"].view(np.uint64))
              and np.array_equal(want["done"] | (want["collision"] << 1), p["flags"]))
        return bool(ok), "%d rows / %d records vs the C oracle" % (len(p["action"]), len(p["prob"]))
    if kind == "table":
        ora = oracle_of(p["spec"])
        want = ora.table_checksums(p["s_begin"], p["n_states"])
        got = dict(zip(want.keys(), p["words"]))
        return want == got, "8 checksum words of %d states x all actions (%d records) vs the C oracle's table walk" % (
            p["n_states"], want["count"])
    if kind == "step":
        ora = oracle_of(p["spec"])
        n = p["spec"]["n_agents"]
        env_ids = np.arange(len(p["action"]), dtype=np.uint64) + np.uint64(p["env_offset"])
        u = device_uniforms(env_ids, p["step_index"], p["seed"], n)
        want = ora.step(p["s_lo"], p["s_hi"], p["action"], u, threads=threads)
        lo, hi = want["next_lo"], want["next_hi"]
        if p["auto_reset"]:
            lo = np.where(want["done"] == 1, np.uint64(p["s0"] & M64), lo)
            hi = np.where(want["done"] == 1, np.uint64(p["s0"] >> 64), hi)
        ok = (np.array_equal(lo, p["next_lo"]) and np.array_equal(hi, p["next_hi"])
              and np.array_equal(want["reward"].view(np.uint64), p["reward"].view(np.uint64))
              and np.array_equal(want["prob"].view(np.uint64), p["prob"].view(np.uint64))
              and np.array_equal(want["done"], p["done"]) and np.array_equal(want["collision"], p["collision"]))
        return bool(ok), "%d env-steps (device Philox draws regenerated on the host) vs the C oracle; %d done, %d collisions" % (
            len(p["action"]), int(want["done"].sum()), int(want["collision"].sum()))
    if kind == "rollout":
        ora = oracle_of(p["spec"])
        n, T = p["spec"]["n_agents"], p["T"]
        lo, hi = p["s_lo"].copy(), p["s_hi"].copy()
        env_ids = np.arange(len(lo), dtype=np.uint64) + np.uint64(p["env_offset"])
        ok = True
        for t in range(T):
            stp = p["step_index"] + t
            if p.get("actions") is not None:
                act = p["actions"][t]
            else:
                w = philox_block(env_ids, stp, 15, p["seed"])
                frac = (w[0] << np.uint64(32)) | w[1]
                act = np.array([(int(f) * p["nA"]) >> 64 for f in frac], dtype=np.int64)  # umul64hi(frac, nA)
            u = device_uniforms(env_ids, stp, p["seed"], n)
            want = ora.step(lo, hi, act, u, threads=threads)
            nlo = np.where(want["done"] == 1, np.uint64(p["s0"] & M64), want["next_lo"])
            nhi = np.where(want["done"] == 1, np.uint64(p["s0"] >> 64), want["next_hi"])
            ok = ok and (np.array_equal(nlo, p["next_lo"][t]) and np.array_equal(nhi, p["next_hi"][t])
                         and np.array_equal(want["reward"].view(np.uint64), p["reward"][t].view(np.uint64))
                         and np.array_equal(want["prob"].view(np.uint64), p["prob"][t].view(np.uint64))
                         and np.array_equal(want["done"], p["done"][t]) and np.array_equal(want["collision"], p["collision"][t]))
            lo, hi = nlo, nhi
        return bool(ok), "%d envs x %d steps (device-drawn actions and slips regenerated on the host) vs the C oracle" % (len(lo), T)
    if kind == "c1":
        from gym_mapf_b200.envs.mapf_env import GYM_MAPF_SEED, _gym_np_random
        from oracle import mapf_oracle
        ora = oracle_of(p["spec"])
        want = ora.table_checksums(0, p["nS"])
        ok = want == dict(zip(want.keys(), p["words"])) and want["count"] == 669808 == p["scalar_transitions"]
        sp = p["spec"]
        spec = mapf_oracle.OracleSpec(sp["rows"], sp["n_agents"], sp["starts"], sp["goals"], sp["fail_prob"], sp["r_clash"],
                                      sp["r_goal"], sp["r_living"], sp["soc"])
        twin, _ = _gym_np_random(GYM_MAPF_SEED)  # the stream env.step() consumed
        for i, a in enumerate(p["acts"]):
            s = p["trace_s"][i]
            term = spec.is_terminal(mapf_oracle.to_digits(s, spec.L, spec.n))
            u = [] if term else [twin.rand() for _ in range(spec.n)]
            ns, r, done, prob, _, _ = spec.step(s, int(a), u)
            ok = ok and ns == p["trace_ns"][i] and float(r) == float(p["trace_r"][i]) and bool(done) == bool(p["trace_d"][i]) \
                and float(prob) == float(p["trace_p"][i])
        return bool(ok), "all 669 808 records by checksum vs the C oracle + 10 000 scalar env.step() calls vs the Python oracle"
    return False, "unknown payload kind %r" % kind


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


_WORKER = {}


def _py_worker(args):
    """One host process stepping ONE env sequentially for `seconds` (random policy, reset on done).  kind "reference":
    the unmodified reference's MapfEnv.step (mapf_env.py:237-266); kind "port": oracle/mapf_oracle.py."""
    seconds, seed, kind = args
    import numpy as np
    rng = np.random.default_rng(seed)
    n = 0
    if kind == "reference":
        if "ref_env" not in _WORKER:
            from oracle import ref_shim
            ref_shim.load_reference()
            import gym_mapf.envs.mapf_env as me
            import gym_mapf.envs.utils as ut
            _WORKER["ref_env"] = ut.create_mapf_env(MAP, SCEN, N_AGENTS, FAIL_PROB, R_CLASH, R_GOAL, R_LIVING,
                                                    me.OptimizationCriteria.SoC)
        env = _WORKER["ref_env"]
        env.reset()
        nA = env.nA
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            for _ in range(100):
                _, _, done, _ = env.step(int(rng.integers(0, nA)))
                if done:
                    env.reset()
            n += 100
        return n, time.perf_counter() - t0
    if "spec" not in _WORKER:
        from oracle import mapf_oracle
        env = make_env()
        rows = ["".join("@" if v else "." for v in r) for r in env.grid.obstacles]
        _WORKER["spec"] = mapf_oracle.OracleSpec(rows, env.n_agents, env.agents_starts, env.agents_goals, FAIL_PROB,
                                                 R_CLASH, R_GOAL, R_LIVING, True)
    spec = _WORKER["spec"]
    s = spec.s0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        for _ in range(200):
            s, _, done, _, _, _ = spec.step(s, int(rng.integers(0, spec.nA)), rng.random(spec.n))
            if done:
                s = spec.s0
        n += 200
    return n, time.perf_counter() - t0


def reference_kind():
    from oracle import ref_shim
    return "reference" if ref_shim.reference_available() else "port"


def cpu_python(seconds, pool, kind, cores):
    res = pool.map(_py_worker, [(seconds, 100 + i, kind) for i in range(cores)])
    total = sum(r[0] for r in res)
    dt = max(r[1] for r in res)
    what = ("the UNMODIFIED reference's MapfEnv.step (mapf_env.py:237-266, imported through oracle/ref_shim.py)"
            if kind == "reference" else "pure-Python port oracle/mapf_oracle.py (no copy of the reference present)")
    return total / dt, "%d sequential env-steps in %d processes (one env each, random policy, reset on done), %s, %.2f s window" % (
        total, cores, what, seconds)


def run_reference(args):
    """The reference's own CPU implementation of the path on the host cores: gym-mapf is single-threaded pure Python,
    so one process per core each steps its own env through the UNMODIFIED `MapfEnv.step` (kind "reference": a copy of
    the reference is found at $MAPF_REFERENCE_ROOT, /root/reference or baseline/_ref; the uninstalled gym / colorama
    imports are stubbed by oracle/ref_shim.py) -- else through the pure-Python port (kind "port").  A step = one bounded
    window of sequential env-steps in every process, sized so that the whole run stays within about two minutes."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_all = time.time()
    env = make_env()
    n = max(1, args.steps + args.warmup)
    per_step = max(0.05, min(6.0, 100.0 / n))
    cores = os.cpu_count() or 1
    kind = reference_kind()
    vals, sample = [], ""
    with mp.get_context("spawn").Pool(cores) as pool:
        pool.map(_py_worker, [(0.05, i, kind) for i in range(cores)])  # start the workers before anything is timed
        for i in range(args.warmup + args.steps):
            v, sample = cpu_python(per_step, pool, kind, cores)
            if i >= args.warmup:
                vals.append(v)
    value = sum(vals) / len(vals)
    c_val, c_cores, c_sample = cpu_c_port(env, 3.0)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "transitions/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int64+f64", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": "transitions/s", "cores": cores, "kind": kind,
                             "sample": sample + " per step", "c_port_value": c_val, "c_port_sample": c_sample},
            "e2e": {"value": value, "unit": "transitions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.time() - t_all}
    print(json.dumps(line), flush=True)


def workload_config(n_gpus):
    return {"workload": "C2: %s scen %d, %d agents, fail_prob %.1f, SoC; batched step over 2**20 envs per GPU" % (
        MAP, SCEN, N_AGENTS, FAIL_PROB), "envs_per_gpu": ENVS_PER_GPU, "global_envs": ENVS_PER_GPU * n_gpus,
        "parallelism": "env-sharded x%d, no data-path collective; per GPU the shard is held as env pools stepped on "
                       "separate streams (see `pools`)" % n_gpus, "sampling": "device Philox4x32-10",
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
        # pointing file descriptor 1 at stderr while the communicator comes up (first collective included).  A short
        # timeout turns any mismatch of collectives into an error within two minutes instead of a spinning GPU.
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    env = make_env(device=local)
    eng = env.engine
    B, K, W = ENVS_PER_GPU, args.steps, max(args.warmup, 3)
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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_region(pools, reps_fixed, sample_clocks):
        """K steps per repetition with the envs in `pools` pools; -> (ms per repetition [median, max over ranks], reps,
        sorted per-repetition ms of this rank, clocks)."""
        H = B // pools
        streams = [torch.cuda.Stream() for _ in range(pools)] if pools > 1 else []

        def one_step(i):
            j = i % RING_SLOTS
            if pools == 1:
                eng.step(states[j], actions[j], seed=seed, step_index=1000 + i, env_offset=env_offset, auto_reset=True,
                         out=outs[j])
                return
            for k, st in enumerate(streams):
                sl = slice(k * H, (k + 1) * H)
                with torch.cuda.stream(st):
                    eng.step(states[j][sl], actions[j][sl], seed=seed, step_index=1000 + i, env_offset=env_offset + k * H,
                             auto_reset=True, out=tuple(t[sl] for t in outs[j]), share_sm=True)

        def k_steps_eager():
            cur = torch.cuda.current_stream()
            for st in streams:
                st.wait_stream(cur)
            for i in range(K):
                one_step(i)
            for st in streams:
                cur.wait_stream(st)

        for i in range(W):
            one_step(i)
        barrier()
        # The K steps of one repetition are captured once into a CUDA graph (kernel nodes joined by programmatic
        # dependent-launch edges, one chain per pool) and replayed: the host launch path (~13 us per call through
        # Python/ctypes) would otherwise be as long as the kernel itself.  --no-graph times plain stream launches.
        graph = None
        if not args.no_graph:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                k_steps_eager()
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                k_steps_eager()
        k_steps = graph.replay if graph is not None else k_steps_eager

        # ---- number of repetitions: decided ONCE and COLLECTIVELY (every rank derives it from the same all-reduced
        # calibration time), so that all ranks issue exactly the same sequence of collectives.  A rank-local wall-clock
        # exit from the loop let ranks disagree by one repetition and hang in mismatched NCCL calls (round 1, N = 8).
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        k_steps()
        e1.record()
        torch.cuda.synchronize()
        cal_ms = max_ranks(e0.elapsed_time(e1))
        reps = reps_fixed if reps_fixed > 0 else int(min(MAX_REPS, max(3, math.ceil(args.min_seconds * 1e3 / max(cal_ms, 1e-3)))))
        # ---- timed region: `reps` repetitions of exactly K steps, each bracketed by its own pair of CUDA events on
        # the launching stream; one barrier + synchronize before the first and after the last (none in between: the
        # GPU runs the repetitions back to back, long enough for nvidia-smi to sample clocks under load).  The median
        # repetition is reported, max over ranks.
        sampler = ClockSampler(local) if sample_clocks else None
        if sampler:
            sampler.start()
            time.sleep(0.15)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        barrier()
        t_wall0 = time.time()
        for a, b in ev:
            a.record()
            k_steps()
            b.record()
        barrier()
        t_wall1 = time.time()
        clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
        reps_sorted = sorted(a.elapsed_time(b) for a, b in ev)
        return max_ranks(reps_sorted[len(reps_sorted) // 2]), reps, reps_sorted, clocks, graph is not None

    pools = max(1, args.pools)
    if B % (2 * pools):
        raise SystemExit("--pools must divide the env batch into even pools")
    single = None
    if pools > 1:  # the single chain of full-GPU launches, for comparison (short: a quarter of the timed window)
        ms1, reps1, _, _, _ = timed_region(1, args.reps, False)
        single = {"ms_per_step": ms1 / K, "value": world * B * K / (ms1 * 1e-3), "reps": reps1,
                  "frac": B * STEP_BYTES / (ms1 / K * 1e-3) / 1e9 / load_peaks()[0],
                  "launch": "one chain of K full-GPU launches (two CTAs per SM), programmatic dependent launch"}
    ms_total, reps, reps_sorted, clocks, graphed = timed_region(pools, args.reps, True)
    ms_per_step = ms_total / K
    value = world * B * K / (ms_total * 1e-3)

    # ---- per-shard counts + checksums, gathered with NCCL (the only collective; outside the timed region)
    j = (K - 1) % RING_SLOTS
    ns, reward, prob, done, coll = outs[j]
    flags = done.to(torch.uint8) + 2 * coll.to(torch.uint8)
    cs = eng.checksum(ns, prob, reward, flags, index_base=env_offset)
    shard_sums = [[int(x) for x in words] for words in sharding.gather_words(cs)]
    # sample of the timed variant's last launch (Philox, two envs per thread, auto-reset) for the oracle check below
    m = 1 << 18
    head_payload = None
    if rank == 0 and not args.no_cpu:
        head_payload = {"kind": "step", "spec": None, "s_lo": states[j][:m].cpu().numpy().view(np.uint64).copy(),
                        "s_hi": np.zeros(m, np.uint64), "action": actions[j][:m].cpu().numpy().astype(np.int64),
                        "seed": seed, "step_index": 1000 + K - 1, "env_offset": env_offset, "auto_reset": True,
                        "s0": int(eng.s0), "next_lo": ns[:m].cpu().numpy().view(np.uint64).copy(), "next_hi": np.zeros(m, np.uint64),
                        "reward": reward[:m].cpu().numpy(), "prob": prob[:m].cpu().numpy(),
                        "done": done[:m].cpu().numpy().astype(np.uint8), "collision": coll[:m].cpu().numpy().astype(np.uint8)}

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

        def host_leg(call):
            for i in range(3):
                call(i)
            barrier()
            t0 = time.perf_counter()
            for i in range(n_e2e):
                call(10 + i)
            torch.cuda.synchronize()
            return max_ranks(time.perf_counter() - t0)

        # Headline: the batched MapfEnv.step(action) -- like the reference's env object (self.s, mapf_env.py:237, 264) the
        # engine keeps the envs' states, on the device; every step the host sends the actions (4 B/env) and receives
        # next_state, reward, done, prob and collision (26 B/env) in pinned host memory.
        ds = states[0].clone()
        dt = host_leg(lambda i: eng.step_host_resident(ds, ha, hout, seed=seed, step_index=i, env_offset=env_offset,
                                                       auto_reset=True))
        e2e = {"value": world * B * n_e2e / dt, "unit": "transitions/s",
               "h2d_bytes_per_step": B * 4 * world, "d2h_bytes_per_step": B * 26 * world, "steps": n_e2e,
               "api": "mapf_step_host_resident (C ABI: states resident on the device as in the reference's env object, "
                      "actions from and all five results to pinned host buffers)",
               "host_numa_bind": prev_affinity is not None,
               "host_link_gbs": world * B * 30 * n_e2e / dt / 1e9}
        resident_ok = bool(torch.equal(ds.cpu(), hout[0]))  # the resident states ARE the returned next states
        # sample of the same call for the oracle check below: 2**16 envs from known dense states (outside the timed region)
        e2e_payload = None
        if rank == 0 and not args.no_cpu:
            me = 1 << 16
            gs = torch.Generator(device=dev)
            gs.manual_seed(11)  # agents packed into 40 cells: terminal states, clashes and goals all occur in the sample
            ds_s = eng.encode(torch.randint(0, 40, (me, N_AGENTS), generator=gs, device=dev, dtype=torch.int32))
            s_before = ds_s.cpu().numpy().view(np.uint64).copy()
            ha_s = ha[:me].clone().pin_memory()
            out_s = tuple(t[:me].clone().pin_memory() for t in hout)
            eng.step_host_resident(ds_s, ha_s, out_s, seed=seed, step_index=7, env_offset=env_offset, auto_reset=True)
            resident_ok = resident_ok and bool(torch.equal(ds_s.cpu(), out_s[0]))
            e2e_payload = {"kind": "step", "spec": None, "s_lo": s_before,
                           "s_hi": np.zeros(me, np.uint64), "action": ha_s.numpy().astype(np.int64), "seed": seed,
                           "step_index": 7, "env_offset": env_offset, "auto_reset": True, "s0": int(eng.s0),
                           "next_lo": out_s[0].numpy().view(np.uint64).copy(), "next_hi": np.zeros(me, np.uint64),
                           "reward": out_s[1].numpy().copy(), "prob": out_s[2].numpy().copy(),
                           "done": out_s[3].numpy().astype(np.uint8), "collision": out_s[4].numpy().astype(np.uint8)}
        e2e["resident_states_match"] = resident_ok
        # the stateless call (round-1 headline): the states cross the link too, 12 B/env in
        dts = host_leg(lambda i: eng.step_host(hs, ha, hout, seed=seed, step_index=i, env_offset=env_offset, auto_reset=True))
        e2e["stateless"] = {"value": world * B * n_e2e / dts, "unit": "transitions/s", "h2d_bytes_per_step": B * 12 * world,
                            "d2h_bytes_per_step": B * 26 * world, "host_link_gbs": world * B * 38 * n_e2e / dts / 1e9,
                            "api": "mapf_step_host (states and actions from pinned host buffers)"}
        # opt-in compact result layout (MAPF_OPT_COMPACT: reward code + one flag byte, 18 instead of 26 B/env back)
        cout = (hout[0], torch.empty(B, dtype=torch.uint8).pin_memory(), hout[2], torch.empty(B, dtype=torch.uint8).pin_memory())
        dtc = host_leg(lambda i: eng.step_host_resident(ds, ha, cout, seed=seed, step_index=i, env_offset=env_offset,
                                                        auto_reset=True, compact=True))
        e2e["compact"] = {"value": world * B * n_e2e / dtc, "unit": "transitions/s", "h2d_bytes_per_step": B * 4 * world,
                          "d2h_bytes_per_step": B * 18 * world, "host_link_gbs": world * B * 22 * n_e2e / dtc / 1e9,
                          "api": "mapf_step_host_resident with MAPF_OPT_COMPACT (reward = reward_table[code], flags = done | "
                                 "collision << 1)"}
        # The memcpy ceiling of THIS box for the same traffic, measured live: every rank copies the step's 4 B/env in and
        # 26 B/env out between the same pinned buffers and device memory with cudaMemcpyAsync on two streams at once
        # (what tools/pcie_peak.py does; boxes of the pool differ by 1.5x in host fabric, so a committed figure misleads).
        d_in = (torch.empty_like(hs, device=dev), torch.empty_like(ha, device=dev))
        d_out = tuple(torch.empty_like(t, device=dev) for t in hout)
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

        def copy_mix(n):
            for _ in range(n):
                with torch.cuda.stream(s_in):
                    d_in[1].copy_(ha, non_blocking=True)
                with torch.cuda.stream(s_out):
                    for h, d in zip(hout, d_out):
                        h.copy_(d, non_blocking=True)
            torch.cuda.synchronize()

        copy_mix(3)
        barrier()
        t0 = time.perf_counter()
        copy_mix(n_e2e)
        dtm = max_ranks(time.perf_counter() - t0)
        mix_gbs = world * B * 30 * n_e2e / dtm / 1e9
        e2e["pcie_ceiling"] = {"step_mix_total_gbs": mix_gbs, "env_steps_per_s": world * B * n_e2e / dtm,
                               "source": "measured live on this box: cudaMemcpyAsync of the step's buffers (4 B/env H2D, 26 B/env "
                                         "D2H, pinned memory, two streams, %d GPU(s) at the same time)" % world}
        e2e["pcie_frac"] = e2e["host_link_gbs"] / mix_gbs
        del d_in, d_out
        if prev_affinity is not None:
            os.sched_setaffinity(0, prev_affinity)

    # ---- every other BASELINE.json config (device-timed, then checked against the oracle on rank 0)
    others = []
    if not args.no_other:
        from tools import bench_configs
        others = bench_configs.run_all(local, world, rank, only=set(args.only.split(",")) if args.only else None,
                                       quick=args.quick)

    if rank == 0:
        peak, peak_src = load_peaks()
        achieved = B * STEP_BYTES / (ms_per_step * 1e-3) / 1e9
        line = {"metric": METRIC, "value": value, "unit": "transitions/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "int64+f64", "data": "synthetic", "config": workload_config(world),
                "reps": reps, "rep_ms_min_med_max": [reps_sorted[0], reps_sorted[len(reps_sorted) // 2], reps_sorted[-1]],
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": NCU_DRAM_BYTES_PER_ENV * B,
                             "traffic_source": "ncu --set full, profiles/r02_step_8m_raw.csv: dram__bytes_read 100.7 MB "
                                               "+ dram__bytes_write 162.1 MB for a 2**23-env launch = 31.3 B/env (reads = "
                                               "the algorithmic 12 B/env; 19.4 of the 26 B/env written had reached DRAM "
                                               "when the kernel ended, the rest was still in the 126 MB L2)",
                             "peak_source": peak_src,
                             "traffic_definition": "per step of 2**20 envs (all pools' launches together), like `achieved`",
                             "kernel": "k_step<4,smem move table>", "bytes_per_unit": STEP_BYTES,
                             "units_per_launch": B // pools, "launches_per_step": pools,
                             "achieved_definition": "38 B x 2**20 envs / (timed region / K steps): the %d launches of a step "
                                                    "run concurrently, so a step's duration, not a launch's, is what the "
                                                    "events measure" % pools},
                "e2e": e2e, "gpu_launches": K * pools, "clocks": clocks, "pools": pools, "single_chain": single,
                "launch": ("one CUDA graph of K x %d step kernels: %d env pool(s), one stream and one chain of programmatic "
                           "dependent launches per pool" % (pools, pools) if graphed else "K x %d stream launches" % pools),
                "shard_checksums": {"keys": ["count", "n_collision", "n_done", "sum_next_lo", "sum_next_hi",
                                             "sum_prob_bits", "sum_reward_bits", "ordered"], "per_gpu": shard_sums}}
        if not args.no_cpu:
            head_payload["spec"] = {"rows": ["".join("@" if v else "." for v in r) for r in env.grid.obstacles],
                                    "n_agents": env.n_agents, "goals": env.agents_goals, "fail_prob": FAIL_PROB,
                                    "r_clash": R_CLASH, "r_goal": R_GOAL, "r_living": R_LIVING, "soc": True}
            ok, what = verify_payload(head_payload)
            line["parity"] = ok
            line["parity_sample"] = "last timed launch of the bench kernel: " + what
            if e2e is not None and e2e_payload is not None:
                e2e_payload["spec"] = head_payload["spec"]
                ok, what = verify_payload(e2e_payload)
                e2e["parity"] = bool(ok and e2e["resident_states_match"])
                e2e["parity_sample"] = "one mapf_step_host_resident call outside the timed region: " + what
            block = {}
            for name, entry, payload in others:
                ok, what = verify_payload(payload)
                entry["parity"] = ok
                entry["parity_sample"] = what
                block[name] = entry
            if block:
                line["other_configs"] = block
                line["other_configs_all_parity"] = all(e["parity"] for e in block.values())
        elif others:
            line["other_configs"] = {name: entry for name, entry, _ in others}
        if world == 1 and not args.no_cpu:
            import multiprocessing as mp
            c_val, c_cores, c_sample = cpu_c_port(env, 4.0)
            kind = reference_kind()
            with mp.get_context("spawn").Pool(c_cores) as pool:
                pool.map(_py_worker, [(0.05, i, kind) for i in range(c_cores)])
                p_val, p_sample = cpu_python(3.0, pool, kind, c_cores)
            line["cpu_baseline"] = {"value": c_val, "unit": "transitions/s", "cores": c_cores, "kind": "port",
                                    "sample": c_sample, "python_value": p_val, "python_kind": kind,
                                    "python_sample": p_sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reps", type=int, default=0, help="timed repetitions of the K-step region (0 = auto, >= ~1.5 s)")
    ap.add_argument("--min-seconds", type=float, default=1.5)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg and the oracle checks")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg (profiling runs)")
    ap.add_argument("--no-graph", action="store_true", help="plain stream launches instead of a CUDA graph")
    ap.add_argument("--pools", type=int, default=2, help="independent env pools per GPU, one stream each (1 = single chain)")
    ap.add_argument("--no-other", action="store_true", help="skip the other_configs block")
    ap.add_argument("--only", default="", help="comma-separated other_configs cases to run")
    ap.add_argument("--quick", action="store_true", help="smaller other_configs cases (smoke runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
