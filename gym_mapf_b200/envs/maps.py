"""The MovingAI benchmark data gym-mapf ships (12 maps x 25 "even" scenarios), stored as one packed bundle
(`maps/movingai.npz`, built by `maps/build_bundle.py`) and written out as standard `.map` / `.scen` text files on
first use, so `map_name_to_files` keeps returning real paths (reference envs/__init__.py:3-10, MANIFEST.in:1)."""
import os
import tempfile
import threading

import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUNDLE = os.path.join(_PKG, "maps", "movingai.npz")


def _default_root():
    root = os.environ.get("GYM_MAPF_B200_MAPS")
    if root:
        return root
    cand = os.path.join(_PKG, "maps", "_files")
    try:
        os.makedirs(cand, exist_ok=True)
        probe = os.path.join(cand, ".w")
        with open(probe, "w"):
            pass
        os.remove(probe)
        return cand
    except OSError:
        return os.path.join(tempfile.gettempdir(), "gym_mapf_b200_maps")


MAPS_PATH = _default_root()
_lock = threading.Lock()
_bundle = None


def bundle():
    global _bundle
    if _bundle is None:
        _bundle = np.load(BUNDLE)
    return _bundle


def map_names():
    return [str(x) for x in bundle()["names"]]


def obstacle_mask(name):
    """uint8[H, W], 1 = obstacle."""
    b = bundle()
    h, w = (int(x) for x in b[name + "/hw"])
    return np.unpackbits(b[name + "/bits"])[:h * w].reshape(h, w)


def ensure_map_files(name):
    """Write `<MAPS_PATH>/<name>/` (the .map file and its 25 .scen files) if the bundle knows the map."""
    folder = os.path.join(MAPS_PATH, name)
    marker = os.path.join(folder, ".complete")
    if os.path.exists(marker):
        return
    with _lock:
        if os.path.exists(marker) or name not in map_names():
            return
        os.makedirs(folder, exist_ok=True)
        b = bundle()
        mask = obstacle_mask(name)
        h, w = mask.shape
        lines = ["type octile", "height %d" % h, "width %d" % w, "map"]
        lines += ["".join("@" if v else "." for v in row) for row in mask]
        _atomic_write(os.path.join(folder, name + ".map"), "\n".join(lines) + "\n")
        for k in range(1, 26):
            recs = b["%s/scen%d" % (name, k)]
            lens = b["%s/scen%d_len" % (name, k)]
            out = ["version 1"]
            for (bucket, xs, ys, xg, yg), opt in zip(recs, lens):
                out.append("%d\t%s.map\t%d\t%d\t%d\t%d\t%d\t%d\t%.8f" % (bucket, name, w, h, xs, ys, xg, yg, opt))
            _atomic_write(os.path.join(folder, "%s-even-%d.scen" % (name, k)), "\n".join(out) + "\n")
        _atomic_write(marker, "ok\n")


def _atomic_write(path, text):
    tmp = "%s.%d.tmp" % (path, os.getpid())
    with open(tmp, "w") as f:
        f.write(text)
    os.replace(tmp, path)
