"""Time the step kernel on the C2 workload (2**20 envs, room-32-32-4, 4 agents) -- tuning helper, prints one line.
Usage: [MAPF_B200_LIB=...] [MAPF_THREADS=..] [MAPF_BLOCKS_PER_SM=..] [MAPF_STEP_EPT=1] python tools/time_step.py [tag]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else ""
    mode = os.environ.get("TIME_MODE", "step")
    env = bench.make_env(device=0)
    eng = env.engine
    B, dev = int(os.environ.get("TIME_B", bench.ENVS_PER_GPU)), torch.device("cuda", 0)
    R = max(2, min(16, (1 << 24) // B))
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    states = [eng.states_from_ints([eng.s0]).expand(B).contiguous()]
    actions = [torch.randint(0, env.nA, (B,), generator=g, device=dev, dtype=torch.int32) for _ in range(R)]
    outs = []
    for j in range(R):
        out = (eng.new_states(B), torch.empty(B, dtype=torch.float64, device=dev),
               torch.empty(B, dtype=torch.float64, device=dev), torch.empty(B, dtype=torch.bool, device=dev),
               torch.empty(B, dtype=torch.bool, device=dev))
        outs.append(out)
        eng.step(states[j], actions[j], seed=3, step_index=j, auto_reset=True, out=out)
        if j + 1 < R:
            states.append(out[0].clone())
    torch.cuda.synchronize()
    best = 1e9
    K = 96
    use_graph = os.environ.get("TIME_GRAPH", "0") == "1"

    n_streams = int(os.environ.get("TIME_STREAMS", "1"))
    pools = [torch.cuda.Stream() for _ in range(n_streams)] if n_streams > 1 else []
    H = B // max(n_streams, 1)

    def run_k():
        if not pools:
            for i in range(K):
                j = i % R
                eng.step(states[j], actions[j], seed=3, step_index=100 + i, auto_reset=True, out=outs[j])
            return
        # the batch as `n_streams` independent env pools, each stepped on its own stream (a chain of dependent launches)
        cur = torch.cuda.current_stream()
        for st in pools:
            st.wait_stream(cur)
        for i in range(K):
            j = i % R
            for k, st in enumerate(pools):
                with torch.cuda.stream(st):
                    sl = slice(k * H, (k + 1) * H)
                    eng.step(states[j][sl], actions[j][sl], seed=3, step_index=100 + i, env_offset=k * H, auto_reset=True,
                             out=tuple(t[sl] for t in outs[j]), share_sm=os.environ.get("TIME_SHARE_SM", "1") == "1")
        for st in pools:
            cur.wait_stream(st)

    graph = None
    if use_graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            run_k()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            run_k()
        torch.cuda.synchronize()
    for rep in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if graph is not None:
            graph.replay()
        else:
            run_k()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / K * 1e3)
    cs = eng.checksum(outs[0][0], outs[0][2], outs[0][1], outs[0][3].to(torch.uint8) + 2 * outs[0][4].to(torch.uint8))
    print("%-8s graph=%s B=%d thr=%s bps=%s ept=%s  %7.2f us/step  %5.1f%% roofline  cs=%d" % (
        tag, os.environ.get("TIME_GRAPH", "0"), B, os.environ.get("MAPF_THREADS", "-"), os.environ.get("MAPF_BLOCKS_PER_SM", "-"),
        os.environ.get("MAPF_STEP_EPT", "-"), best, 100 * B * 38 / (best * 1e-6) / 1e9 / 6436.1,
        int(cs.cpu()[7].item()) & 0xffffffff), flush=True)


if __name__ == "__main__":
    main()
