"""Helpers shared by the parity tests: load the reference-generated fixtures in tests/golden/."""
import glob
import json
import os
import struct

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
M64 = (1 << 64) - 1


def names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = {k: z[k] for k in z.files if k != "spec"}
    spec = json.loads(str(z["spec"]))
    return spec, d


def misc():
    with open(os.path.join(GOLDEN, "misc.json")) as f:
        return json.load(f)


def bits_to_f64(bits):
    return np.asarray(bits, dtype=np.uint64).view(np.float64)


def f64_to_bits(x):
    return np.ascontiguousarray(x, dtype=np.float64).view(np.uint64)


def py_bits(x):
    return struct.unpack("<Q", struct.pack("<d", float(x)))[0]


def big(lo, hi):
    return int(lo) | (int(hi) << 64)


def checksums(next_lo, next_hi, prob_bits, reward_bits, done, collision):
    """Same definition as oracle/make_golden.py:checksums (all mod 2**64)."""
    count = int(next_lo.shape[0])
    idx = np.arange(1, count + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        tag = next_lo + np.uint64(1) + np.uint64(2) * collision.astype(np.uint64) + np.uint64(4) * done.astype(np.uint64)
        return dict(count=count, n_collision=int(collision.sum()), n_done=int(done.sum()),
                    sum_next_lo=int(np.sum(next_lo, dtype=np.uint64)), sum_next_hi=int(np.sum(next_hi, dtype=np.uint64)),
                    sum_prob_bits=int(np.sum(prob_bits, dtype=np.uint64)),
                    sum_reward_bits=int(np.sum(reward_bits, dtype=np.uint64)),
                    ordered=int(np.sum(idx * tag, dtype=np.uint64)))
