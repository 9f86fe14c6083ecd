"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference (`/root/reference/gym_mapf`).

Used by `oracle/make_golden*.py`, `tests/test_oracle_vs_reference.py` (skipped when no reference is present) and
`bench.py --impl reference` (the CPU arm times the unmodified reference when a copy is present).  The reference is
looked for at `$MAPF_REFERENCE_ROOT`, `/root/reference` (build container only) and `<repo>/baseline/_ref` (the
offline `pip install --no-deps --target baseline/_ref` of the unmodified reference: git-ignored, shipped to the GPU
box by gpurun).  Nothing in the `-m gpu` tests or `smoke()` reads `/root/reference`.

The reference imports five third-party names that are not installed here (no network):

    colorama.Fore                                   mapf_env.py:7      (render colours only)
    gym.spaces.Discrete, gym.Env                    mapf_env.py:8-9    (inert holders)
    gym.envs.toy_text.discrete.categorical_sample   mapf_env.py:10,255 (step() sampling)
    gym.utils.seeding.np_random                     mapf_env.py:11,139 (RandomState, seed 42)

`install_stubs()` registers minimal stand-ins in `sys.modules`.  `categorical_sample` restates the published
gym 0.13.0 behaviour (`requirements.txt:7` pins gym==0.13.0): `(np.cumsum(p) > rng.rand()).argmax()`.
`np_random` restates gym 0.13.0's `seeding.np_random`: a `RandomState` seeded with the 32-bit words of the first
8 bytes of sha512(str(seed)) -- the same stream the product's `envs/mapf_env.py:_gym_np_random` builds
(`tests/test_oracle_golden.py::test_np_random_stream_pinned` pins both).  The reference's own tests never pin that
stream (every `step()` test uses `fail_prob=0`); step traces are additionally checked through *explicit uniforms*
(see `UniformTape`).
"""
import hashlib
import os
import struct
import sys
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_reference_root():
    cands = [os.environ.get("MAPF_REFERENCE_ROOT"), "/root/reference",
             os.path.join(os.path.dirname(_HERE), "baseline", "_ref")]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "gym_mapf", "envs")):
            return c
    return cands[0] or "/root/reference"


REFERENCE_ROOT = _find_reference_root()


class UniformTape:
    """A stand-in for `RandomState` whose `rand()` replays a given list of uniforms (and records how many
    were consumed), so a `step()` trace of the reference can be replayed bit-exactly elsewhere."""

    def __init__(self, values):
        self.values = [float(v) for v in values]
        self.pos = 0

    def rand(self):
        v = self.values[self.pos]
        self.pos += 1
        return v


def _categorical_sample(prob_n, np_random):
    prob_n = np.asarray(prob_n)
    csprob_n = np.cumsum(prob_n)
    return (csprob_n > np_random.rand()).argmax()


def _np_random(seed=None):
    """gym 0.13.0 `seeding.np_random` (create_seed -> hash_seed -> _bigint_from_bytes -> _int_list_from_bigint)."""
    seed = int(seed) % 2 ** 64
    digest = hashlib.sha512(str(seed).encode("utf8")).digest()[:8] + b"\0" * 4  # gym pads to a multiple of 4 + 4
    big = sum(2 ** (32 * i) * v for i, v in enumerate(struct.unpack("3I", digest)))
    words = []
    while big > 0:
        big, mod = divmod(big, 2 ** 32)
        words.append(mod)
    rng = np.random.RandomState()
    rng.seed(words if words else [0])
    return rng, seed


def install_stubs():
    if "gym" in sys.modules and getattr(sys.modules["gym"], "__mapf_stub__", False):
        return
    gym = types.ModuleType("gym")
    gym.__mapf_stub__ = True

    class Env:  # gym.Env: inert base class
        pass

    gym.Env = Env
    spaces = types.ModuleType("gym.spaces")

    class Discrete:
        def __init__(self, n):
            self.n = n

    spaces.Discrete = Discrete
    gym.spaces = spaces
    envs = types.ModuleType("gym.envs")
    toy_text = types.ModuleType("gym.envs.toy_text")
    discrete = types.ModuleType("gym.envs.toy_text.discrete")
    discrete.categorical_sample = _categorical_sample
    toy_text.discrete = discrete
    envs.toy_text = toy_text
    gym.envs = envs
    utils = types.ModuleType("gym.utils")
    seeding = types.ModuleType("gym.utils.seeding")
    seeding.np_random = _np_random
    utils.seeding = seeding
    gym.utils = utils
    colorama = types.ModuleType("colorama")

    class _Fore:
        RED = GREEN = YELLOW = BLUE = RESET = ""

    colorama.Fore = _Fore
    for name, mod in [("gym", gym), ("gym.spaces", spaces), ("gym.envs", envs),
                      ("gym.envs.toy_text", toy_text), ("gym.envs.toy_text.discrete", discrete),
                      ("gym.utils", utils), ("gym.utils.seeding", seeding), ("colorama", colorama)]:
        sys.modules[name] = mod


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "gym_mapf", "envs"))


def load_reference():
    """Import the unmodified reference package and return the module `gym_mapf`."""
    if not reference_available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True  # /root/reference is read-only
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import gym_mapf  # noqa: F401
    import gym_mapf.envs.mapf_env  # noqa: F401
    import gym_mapf.envs.utils  # noqa: F401
    return sys.modules["gym_mapf"]
