"""Timeline of the bench's step launches without a profiler: a -DMAPF_TRACE build of the library stamps %globaltimer at
the phase boundaries of every CTA (entry, image staging issued, before / after the grid-dependency wait, tables ready,
end of every iteration, exit).  Replays one CUDA graph of K launches exactly as bench.py does and prints, per launch,
when its CTAs started, when the dependency wait released them and when they finished, relative to the first stamp.

    make -C gym_mapf_b200/csrc variant TAG=trace VN=4 EXTRA="-DMAPF_TRACE -DMAPF_TUNING"
    MAPF_B200_LIB=gym_mapf_b200/csrc/libmapf_b200_trace.so python tools/trace_step.py [K] > profiles/r02_step_timeline.txt
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    B = int(os.environ.get("TIME_B", bench.ENVS_PER_GPU))
    dev = torch.device("cuda", 0)
    buf = torch.zeros(2 + 2 * (1 << 20), dtype=torch.int64, device=dev)
    os.environ["MAPF_TRACE_PTR"] = str(buf.data_ptr())
    env = bench.make_env(device=0)
    eng = env.engine
    R = 16
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    states = [eng.states_from_ints([eng.s0]).expand(B).contiguous()]
    actions = [torch.randint(0, env.nA, (B,), generator=g, device=dev, dtype=torch.int32) for _ in range(R)]
    outs = []
    for j in range(R):
        out = (eng.new_states(B), torch.empty(B, dtype=torch.float64, device=dev), torch.empty(B, dtype=torch.float64, device=dev),
               torch.empty(B, dtype=torch.bool, device=dev), torch.empty(B, dtype=torch.bool, device=dev))
        outs.append(out)
        eng.step(states[j], actions[j], seed=3, step_index=j, auto_reset=True, out=out)
        if j + 1 < R:
            states.append(out[0].clone())

    n_pools = int(os.environ.get("TIME_STREAMS", "1"))
    pools = [torch.cuda.Stream() for _ in range(n_pools)] if n_pools > 1 else []
    H = B // max(n_pools, 1)

    def run_k():
        if not pools:
            for i in range(K):
                j = i % R
                eng.step(states[j], actions[j], seed=3, step_index=100 + i, auto_reset=True, out=outs[j])
            return
        cur = torch.cuda.current_stream()
        for st in pools:
            st.wait_stream(cur)
        for i in range(K):
            j = i % R
            for k, st in enumerate(pools):  # pool k stamps its launches with step index 100 + i + 1000 * k
                with torch.cuda.stream(st):
                    sl = slice(k * H, (k + 1) * H)
                    eng.step(states[j][sl], actions[j][sl], seed=3, step_index=100 + i + 1000 * k, env_offset=k * H,
                             auto_reset=True, out=tuple(t[sl] for t in outs[j]), share_sm=True)
        for st in pools:
            cur.wait_stream(st)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        run_k()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        run_k()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    buf.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    raw = buf.cpu().numpy().view(np.uint64)
    n = int(raw[0])
    ev = raw[2:2 + 2 * n].reshape(n, 2)
    tag, step, smid, cta, t = (ev[:, 0] >> 48).astype(int), ((ev[:, 0] >> 32) & 0xffff).astype(int), \
        ((ev[:, 0] >> 16) & 0xffff).astype(int), (ev[:, 0] & 0xffff).astype(int), ev[:, 1].astype(np.int64)
    t0 = t.min()
    t = (t - t0) / 1e3  # us
    print("# graph of %d steps of k_step<4,...> over %d envs, %d pool(s): %.2f us per step by CUDA events (with the stamps; "
          "the stamps' atomics cost ~40 %%)" % (K, B, max(n_pools, 1), e0.elapsed_time(e1) * 1e3 / K))
    if n_pools > 1:
        print("# launch ids: pool 0 = 0..%d, pool 1 = 1000.. (sorted by id, not by time: compare the two pools' columns)" % (K - 1))
    print("# %d stamps; columns are microseconds since the first stamp: min / median / max over the launch's CTAs" % n)
    print("# launch | CTA entry            | wait released        | tables ready (1st it) | end of iteration 1   | CTA exit             | "
          "span first entry -> last exit | next launch's first entry - this launch's last exit")
    names = {0: "entry", 3: "released", 4: "tables", 8: "iter1", 7: "exit"}
    prev_exit = None
    rows = []
    for s in sorted(set(step)):
        m = step == s
        cells = []
        for tg in (0, 3, 4, 8, 7):
            x = np.sort(t[m & (tag == tg)])
            cells.append("%6.2f %6.2f %6.2f" % (x[0], x[len(x) // 2], x[-1]) if len(x) else " " * 20)
        first, last = t[m & (tag == 0)].min(), t[m & (tag == 7)].max()
        rows.append((s, cells, first, last))
    for i, (s, cells, first, last) in enumerate(rows):
        nxt = rows[i + 1][2] - last if i + 1 < len(rows) else float("nan")
        print("%7d | %s | %s | %s | %s | %s | %6.2f | %+6.2f" % (s - 100, cells[0], cells[1], cells[2], cells[3], cells[4],
                                                                 last - first, nxt))
    # per-iteration durations of one middle launch
    mid = sorted(set(step))[len(set(step)) // 2]
    m = step == mid
    print("# launch %d: iterations per CTA and duration of each (us, median over CTAs)" % (mid - 100))
    per = {}
    for c in set(cta[m]):
        mc = m & (cta == c)
        rel = t[mc & (tag == 3)]
        its = sorted((tg, tt) for tg, tt in zip(tag[mc], t[mc]) if tg >= 8)
        prev = rel[0] if len(rel) else None
        for tg, tt in its:
            per.setdefault(tg - 8, []).append(tt - prev)
            prev = tt
    for k in sorted(per):
        v = np.sort(per[k])
        print("#   iteration %d: %d CTAs, %.2f / %.2f / %.2f us (min / median / max)" % (k + 1, len(v), v[0], v[len(v) // 2], v[-1]))


if __name__ == "__main__":
    main()
