"""Shared helpers of the GPU parity tests: build a device Engine / C oracle from a golden fixture's spec."""
import numpy as np

import golden_util as G
from oracle import c_oracle


def obstacles_of(spec):
    return np.array([[1 if ch == "@" else 0 for ch in row] for row in spec["rows"]], dtype=np.uint8)


def make_engine(spec, device=0):
    from gym_mapf_b200 import _native
    return _native.Engine(obstacles_of(spec), spec["n_agents"], spec["starts"], spec["goals"], spec["fail_prob"],
                          spec["r_clash"], spec["r_goal"], spec["r_living"], not spec["soc"], device=device)


def make_oracle(spec):
    return c_oracle.COracle(spec["rows"], spec["n_agents"], spec["goals"], spec["fail_prob"], spec["r_clash"],
                            spec["r_goal"], spec["r_living"], spec["soc"])


def states_tensor(eng, lo, hi):
    import torch
    lo = np.ascontiguousarray(lo, dtype=np.uint64)
    if eng.words == 1:
        assert not np.any(hi)
        return torch.from_numpy(lo.view(np.int64).copy()).to(eng.torch_device)
    both = np.stack([lo, np.ascontiguousarray(hi, dtype=np.uint64)], axis=1).view(np.int64)
    return torch.from_numpy(both.copy()).to(eng.torch_device)


def split_states(eng, t):
    arr = t.detach().cpu().numpy().view(np.uint64)
    if eng.words == 1:
        return arr.copy(), np.zeros_like(arr)
    arr = arr.reshape(-1, 2)
    return arr[:, 0].copy(), arr[:, 1].copy()


def u64(t):
    return t.detach().cpu().numpy().view(np.uint64)


def assert_rows_equal(eng, got, want_ptr, want_lo, want_hi, want_prob_bits, want_reward_bits, want_done, want_coll):
    row_ptr, ns, prob, reward, flags = got
    assert np.array_equal(row_ptr.cpu().numpy(), want_ptr)
    lo, hi = split_states(eng, ns)
    assert np.array_equal(lo, want_lo) and np.array_equal(hi, want_hi)
    assert np.array_equal(u64(prob), want_prob_bits)
    assert np.array_equal(u64(reward), want_reward_bits)
    f = flags.cpu().numpy()
    assert np.array_equal(f & 1, want_done) and np.array_equal((f >> 1) & 1, want_coll)
    assert not np.any(f & ~np.uint8(3))
