"""Time the batched step on any (map, scenario, agents) spec -- tuning helper (bench_configs.step_case on one spec).
Usage: [MAPF_B200_LIB=...] python tools/time_step_spec.py <map> <scen> <agents> [log2 envs, default 22] [soc 0|1]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tools import bench_configs as bc  # noqa: E402


def main():
    name, scen, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    lg = int(sys.argv[4]) if len(sys.argv) > 4 else 22
    soc = bool(int(sys.argv[5])) if len(sys.argv) > 5 else True
    dev = torch.device("cuda", 0)
    bc.warm_up_clocks()
    env = bc.make(name, scen, n, soc, 0)
    entry, _ = bc.step_case("%s scen %d, %d agents" % (name, scen, n), env, soc, dev, bc.load_peak(), 1 << lg, 1, 0,
                            ring=2, K=4, graph=False)
    print(json.dumps({"spec": [name, scen, n], "envs": 1 << lg, "words": env.engine.words,
                      "us_per_step": entry["us_per_step"], "frac": entry["frac"]}), flush=True)


if __name__ == "__main__":
    main()
