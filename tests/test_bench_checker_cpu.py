"""CPU-only checks of bench.py's oracle checker: the host Philox against the Random123 known answers, and
`verify_payload` on payloads built from the oracle itself (must pass) and on corrupted copies (must fail)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import c_oracle  # noqa: E402

SPEC = {"rows": ["....", ".@..", "....", "...."], "n_agents": 2, "goals": [[3, 3], [0, 0]], "starts": [[0, 0], [3, 3]],
        "fail_prob": 0.2, "r_clash": -1000.0, "r_goal": 100.0, "r_living": -1.0, "soc": True}


def test_host_philox_known_answers():
    out = bench.philox4x32_10([[0], [0], [0], [0]], [0, 0])
    assert [int(x[0]) for x in out] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    out = bench.philox4x32_10([[0x243f6a88], [0x85a308d3], [0x13198a2e], [0x03707344]], [0xa4093822, 0x299f31d0])
    assert [int(x[0]) for x in out] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    u = bench.device_uniforms(np.arange(5, dtype=np.uint64), 7, 11, 6)
    assert u.shape == (5, 6) and float(u.min()) >= 0.0 and float(u.max()) < 1.0


def _step_payload():
    ora = c_oracle.COracle(SPEC["rows"], 2, SPEC["goals"], 0.2, -1000.0, 100.0, -1.0, True)
    rng = np.random.default_rng(0)
    B = 500
    lo, hi = ora.encode(rng.integers(0, ora.L, (B, 2)).astype(np.int32))
    act = rng.integers(0, 25, B).astype(np.int64)
    seed, step_index, env_offset, s0 = 99, (3 << 32) | 5, 1 << 34, 7
    u = bench.device_uniforms(np.arange(B, dtype=np.uint64) + np.uint64(env_offset), step_index, seed, 2)
    w = ora.step(lo, hi, act, u)
    nlo = np.where(w["done"] == 1, np.uint64(s0), w["next_lo"])
    return {"kind": "step", "spec": SPEC, "s_lo": lo, "s_hi": hi, "action": act, "seed": seed, "step_index": step_index,
            "env_offset": env_offset, "auto_reset": True, "s0": s0, "next_lo": nlo, "next_hi": np.zeros(B, np.uint64),
            "reward": w["reward"], "prob": w["prob"], "done": w["done"], "collision": w["collision"]}


def test_verify_payload_step_and_rows():
    p = _step_payload()
    ok, what = bench.verify_payload(p)
    assert ok, what
    bad = dict(p)
    bad["prob"] = p["prob"].copy()
    bad["prob"][3] = np.nextafter(bad["prob"][3], 2.0)  # one ulp off must be caught
    assert not bench.verify_payload(bad)[0]
    ora = c_oracle.COracle(SPEC["rows"], 2, SPEC["goals"], 0.2, -1000.0, 100.0, -1.0, True)
    w = ora.rows(p["s_lo"][:50], p["s_hi"][:50], p["action"][:50])
    rows = {"kind": "rows", "spec": SPEC, "s_lo": p["s_lo"][:50], "s_hi": p["s_hi"][:50], "action": p["action"][:50],
            "row_ptr": w["row_ptr"], "next_lo": w["next_lo"], "next_hi": w["next_hi"], "prob": w["prob"], "reward": w["reward"],
            "flags": w["done"] | (w["collision"] << 1)}
    assert bench.verify_payload(rows)[0]
    rows["flags"] = rows["flags"].copy()
    rows["flags"][0] ^= 2
    assert not bench.verify_payload(rows)[0]
    table = {"kind": "table", "spec": SPEC, "s_begin": 10, "n_states": 20, "words": list(ora.table_checksums(10, 20).values())}
    assert bench.verify_payload(table)[0]
    table["words"][7] ^= 1
    assert not bench.verify_payload(table)[0]
