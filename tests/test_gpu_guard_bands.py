"""Out-of-bounds writes: every output buffer of the hot entry points is carved out of a larger allocation whose padding
on both sides holds a canary pattern; after the call the padding must be untouched.  (compute-sanitizer is not available
on the GPU pool these tests run on; this is the memory-safety check that does not depend on it.)  Ragged sizes on
purpose: 1, odd (scalar kernels), even (128-bit kernels), around a warp / a CTA / the grid stride."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

PAD = 256        # bytes on each side
CANARY = 0xA5
SIZES = (1, 2, 31, 33, 255, 258, 2049, 70001)


@pytest.fixture(scope="module")
def torch():
    import torch as t
    return t


def _env(name, scen, n, soc=True):
    from gym_mapf_b200.envs.mapf_env import OptimizationCriteria
    from gym_mapf_b200.envs.utils import create_mapf_env
    crit = OptimizationCriteria.SoC if soc else OptimizationCriteria.Makespan
    return create_mapf_env(name, scen, n, 0.2, -1000.0, 100.0, -1.0, crit, device=0)


class Guarded:
    """Device tensors carved out of canary-filled allocations."""

    def __init__(self, torch):
        self.torch, self.raw = torch, []

    def new(self, shape, dtype):
        torch = self.torch
        shape = tuple(int(x) for x in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        item = torch.empty(0, dtype=dtype).element_size()
        nbytes = item * int(np.prod(shape))
        body = (nbytes + 15) // 16 * 16  # the view starts 256-byte aligned like a fresh allocation
        buf = torch.full((PAD + body + PAD,), CANARY, dtype=torch.uint8, device="cuda:0")
        self.raw.append((buf, nbytes))
        view = buf[PAD:PAD + nbytes].view(dtype).view(shape)
        assert view.is_contiguous() and view.data_ptr() % 16 == 0
        return view

    def check(self, what):
        self.torch.cuda.synchronize()
        for k, (buf, nbytes) in enumerate(self.raw):
            head, tail = buf[:PAD], buf[PAD + nbytes:]
            assert bool((head == CANARY).all()) and bool((tail == CANARY).all()), "%s: write outside buffer %d" % (what, k)
        self.raw = []


SPECS = (("room-32-32-4", 1, 4, True), ("room-64-64-8", 1, 8, False), ("empty-8-8", 1, 2, True))


@pytest.mark.parametrize("spec", SPECS, ids=lambda s: "%s-%d" % (s[0], s[2]))
def test_step_and_rollout_write_only_their_outputs(spec, torch):
    env = _env(*spec)
    eng = env.engine
    g = Guarded(torch)
    rng = np.random.default_rng(5)
    for B in SIZES:
        cells = torch.from_numpy(rng.integers(0, min(eng.L, 50), (B, eng.n)).astype(np.int32)).cuda()
        st = eng.encode(cells)
        ac = torch.from_numpy(rng.integers(0, eng.nA, B).astype(np.int32)).cuda()
        un = torch.from_numpy(rng.random((B, eng.n))).cuda()
        for uniforms in (None, un):
            out = (g.new(eng.state_shape(B), torch.int64), g.new(B, torch.float64), g.new(B, torch.float64),
                   g.new(B, torch.bool), g.new(B, torch.bool))
            eng.step(st, ac, uniforms=uniforms, seed=1, step_index=2, auto_reset=True, out=out)
            g.check("step B=%d tape=%s" % (B, uniforms is not None))
            outc = (g.new(eng.state_shape(B), torch.int64), g.new(B, torch.uint8), g.new(B, torch.float64),
                    g.new(B, torch.uint8))
            eng.step(st, ac, uniforms=uniforms, seed=1, step_index=2, auto_reset=True, out=outc, compact=True)
            g.check("compact step B=%d tape=%s" % (B, uniforms is not None))
        # host step with the states resident on the device: the in-place state array is an output too
        keep = g.new(eng.state_shape(B), torch.int64)
        keep.copy_(st)
        hout = (torch.empty(eng.state_shape(B), dtype=torch.int64).pin_memory(), torch.empty(B, dtype=torch.float64).pin_memory(),
                torch.empty(B, dtype=torch.float64).pin_memory(), torch.empty(B, dtype=torch.bool).pin_memory(),
                torch.empty(B, dtype=torch.bool).pin_memory())
        eng.step_host_resident(keep, ac.cpu().pin_memory(), hout, seed=1, step_index=2, auto_reset=True)
        g.check("step_host_resident B=%d" % B)
        # rollout: T x B result slabs and the state array advanced in place
        T = 3
        for given in (True, False):
            acts = torch.from_numpy(rng.integers(0, eng.nA, (T, B)).astype(np.int32)).cuda() if given else None
            sio = g.new(eng.state_shape(B), torch.int64)
            sio.copy_(st)
            out = (g.new((T,) + eng.state_shape(B), torch.int64), g.new((T, B), torch.float64), g.new((T, B), torch.float64),
                   g.new((T, B), torch.bool), g.new((T, B), torch.bool))
            eng.rollout(sio, acts, T, seed=3, step_index=0, auto_reset=True, out=out)
            g.check("rollout B=%d given=%s" % (B, given))


@pytest.mark.parametrize("spec", SPECS, ids=lambda s: "%s-%d" % (s[0], s[2]))
def test_table_path_writes_only_its_outputs(spec, torch):
    from gym_mapf_b200._native import _ptr, check, lib
    env = _env(*spec)
    eng = env.engine
    g = Guarded(torch)
    rng = np.random.default_rng(6)
    s = eng._stream()
    for B in SIZES[:-1] + (5001,):
        cells = torch.from_numpy(rng.integers(0, min(eng.L, 50), (B, eng.n)).astype(np.int32)).cuda()
        st = eng.encode(cells)
        ac = torch.from_numpy(rng.integers(0, eng.nA, B).astype(np.int32)).cuda()
        scratch_bytes = int(lib().mapf_scan_scratch_bytes(B))
        for with_len in (False, True):
            row_ptr = g.new(B + 1, torch.int64)
            scratch = g.new((scratch_bytes + 7) // 8, torch.int64)
            row_len = g.new(B, torch.int64) if with_len else None
            check(lib().mapf_count_scan_rows(eng._h, _ptr(st), _ptr(ac), B, _ptr(row_len), _ptr(row_ptr), _ptr(scratch), s))
            g.check("count_scan_rows B=%d row_len=%s" % (B, with_len))
        total = int(row_ptr[-1].item())
        rp = row_ptr.clone()
        out = (g.new(eng.state_shape(total), torch.int64), g.new(total, torch.float64), g.new(total, torch.float64),
               g.new(total, torch.uint8))
        check(lib().mapf_expand(eng._h, _ptr(st), _ptr(ac), B, _ptr(rp), _ptr(out[0]), _ptr(out[1]), _ptr(out[2]),
                                _ptr(out[3]), s))
        g.check("expand B=%d (%d records)" % (B, total))
    # a slab of consecutive states x all actions
    n_states = 3
    s_begin = int(eng.s0)
    sb = (C.c_uint64 * 2)(s_begin & ((1 << 64) - 1), s_begin >> 64)
    Bt = n_states * int(eng.nA)
    if int(eng.nA) <= 15625:  # (8 agents: 390625 rows of up to 6561 records per state)
        row_ptr = g.new(Bt + 1, torch.int64)
        scratch = g.new((int(lib().mapf_scan_scratch_bytes(Bt)) + 7) // 8, torch.int64)
        check(lib().mapf_count_scan_range(eng._h, C.byref(sb), n_states, None, _ptr(row_ptr), _ptr(scratch), s))
        g.check("count_scan_range")
        total = int(row_ptr[-1].item())
        rp = row_ptr.clone()
        out = (g.new(eng.state_shape(total), torch.int64), g.new(total, torch.float64), g.new(total, torch.float64),
               g.new(total, torch.uint8))
        check(lib().mapf_expand_range(eng._h, C.byref(sb), n_states, _ptr(rp), _ptr(out[0]), _ptr(out[1]), _ptr(out[2]),
                                      _ptr(out[3]), s))
        g.check("expand_range (%d records)" % total)
