"""Host-side call rate of VecMapfEnv.step (no CUDA graph): how much of the kernel's throughput survives the Python /
ctypes launch path."""
import sys, time, torch
sys.path.insert(0,'/root/repo')
from gym_mapf_b200.envs.mapf_env import OptimizationCriteria
from gym_mapf_b200.envs.utils import create_mapf_env
from gym_mapf_b200.envs.vec_env import VecMapfEnv
env = create_mapf_env("room-32-32-4", 1, 4, 0.2, -1000.0, 100.0, -1.0, OptimizationCriteria.SoC, device=0)
for B, reuse in ((1<<20, False), (1<<20, True), (1<<16, True)):
    vec = VecMapfEnv(env, B, seed=1, reuse_outputs=reuse)
    acts = torch.randint(0, env.nA, (B,), device='cuda', dtype=torch.int32)
    for _ in range(20): vec.step(acts)
    torch.cuda.synchronize(); t0=time.perf_counter()
    n=2000
    for _ in range(n): vec.step(acts)
    torch.cuda.synchronize(); dt=time.perf_counter()-t0
    print("B=%d reuse_outputs=%s: %.2f us per vec.step call, %.3g env-steps/s" % (B, reuse, dt/n*1e6, B*n/dt))
