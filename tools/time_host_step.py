"""Time the end-to-end host-buffer step (mapf_step_host) on the C2 workload: pinned host buffers in and out.
Usage: [MAPF_B200_LIB=...libmapf_b200_tune.so MAPF_HOST_MODE=0|1] python tools/time_host_step.py [B]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else bench.ENVS_PER_GPU
env = bench.make_env(device=0)
eng = env.engine
hs = torch.empty(eng.state_shape(B), dtype=torch.int64).pin_memory()
hs.copy_(eng.states_from_ints([eng.s0]).expand(B))
ha = torch.randint(0, env.nA, (B,), dtype=torch.int32).pin_memory()
hout = (torch.empty(eng.state_shape(B), dtype=torch.int64).pin_memory(), torch.empty(B, dtype=torch.float64).pin_memory(),
        torch.empty(B, dtype=torch.float64).pin_memory(), torch.empty(B, dtype=torch.bool).pin_memory(),
        torch.empty(B, dtype=torch.bool).pin_memory())
for i in range(3):
    eng.step_host(hs, ha, hout, seed=1, step_index=i, auto_reset=True)
best = 1e9
for rep in range(5):
    t0 = time.perf_counter()
    for i in range(20):
        eng.step_host(hs, ha, hout, seed=1, step_index=10 + i, auto_reset=True)
    torch.cuda.synchronize()
    best = min(best, (time.perf_counter() - t0) / 20)
dev = eng.step(hs.cuda(), ha.cuda(), seed=1, step_index=29, auto_reset=True)
same = all(torch.equal(a.cpu(), b) for a, b in zip(dev, hout))
print("mode=%s B=%d  %.3f ms per call  %.3e env-steps/s  %.1f GB/s over the host link  same_as_device_step=%s" % (
    os.environ.get("MAPF_HOST_MODE", "default"), B, best * 1e3, B / best, B * 38 / best / 1e9, same))
