"""Pack the MovingAI benchmark data the reference ships (`gym_mapf/maps/<map>/<map>.map` and
`<map>-even-<1..25>.scen`, reference `MANIFEST.in:1`) into one compressed bundle, `movingai.npz`.

Run once in the build container (the reference tree is not available on the GPU box):

    python gym_mapf_b200/maps/build_bundle.py [/root/reference/gym_mapf/maps]

Bundle layout (numpy `.npz`, deflate):
    names                      unicode array of map names
    <map>/hw                   int32[2]  (height, width)
    <map>/bits                 uint8[]   np.packbits of the row-major obstacle mask (1 = '@')
    <map>/scen<k>              int32[m,5] per scenario k: bucket, x_start, y_start, x_goal, y_goal
    <map>/scen<k>_len          float64[m] the scenario's optimal-length column
`gym_mapf_b200.envs.maps` re-materialises standard MovingAI text files from it on demand, so
`map_name_to_files()` keeps returning real file paths (reference `envs/__init__.py:6-10`).
"""
import os
import sys

import numpy as np


def pack(maps_root, out_path):
    arrays = {}
    names = sorted(d for d in os.listdir(maps_root) if os.path.isdir(os.path.join(maps_root, d)))
    for name in names:
        with open(os.path.join(maps_root, name, name + ".map")) as f:
            lines = f.read().split("\n")
        assert lines[0].startswith("type octile") and lines[3].strip() == "map", name
        h = int(lines[1].split()[1])
        w = int(lines[2].split()[1])
        rows = [ln.strip() for ln in lines[4:4 + h]]
        assert all(len(r) == w and set(r) <= {".", "@"} for r in rows), name
        mask = np.array([[ch == "@" for ch in r] for r in rows], dtype=np.uint8)
        arrays[name + "/hw"] = np.array([h, w], dtype=np.int32)
        arrays[name + "/bits"] = np.packbits(mask.reshape(-1))
        for k in range(1, 26):
            path = os.path.join(maps_root, name, "%s-even-%d.scen" % (name, k))
            recs, lens = [], []
            with open(path) as f:
                header = f.readline()
                assert header.strip() == "version 1", path
                for ln in f:
                    if not ln.strip():
                        continue
                    bucket, mname, mw, mh, xs, ys, xg, yg, opt = ln.rstrip("\n").split("\t")
                    assert mname == name + ".map" and int(mw) == w and int(mh) == h, path
                    recs.append((int(bucket), int(xs), int(ys), int(xg), int(yg)))
                    lens.append(float(opt))
                    assert "%.8f" % float(opt) == opt, (path, opt)
            arrays["%s/scen%d" % (name, k)] = np.array(recs, dtype=np.int32).reshape(-1, 5)
            arrays["%s/scen%d_len" % (name, k)] = np.array(lens, dtype=np.float64)
    arrays["names"] = np.array(names)
    np.savez_compressed(out_path, **arrays)
    return names


if __name__ == "__main__":
    root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/gym_mapf/maps"
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "movingai.npz")
    print(pack(root, out), os.path.getsize(out))
