"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference (`/root/reference/gym_mapf`).

Only `oracle/make_golden.py` and `tests/test_oracle_vs_reference.py` use this file, and only in the build
container: `/root/reference` does not exist on the GPU box, so nothing in the `-m gpu` tests, `smoke()` or
`bench.py` may import it.

The reference imports five third-party names that are not installed here (no network):

    colorama.Fore                                   mapf_env.py:7      (render colours only)
    gym.spaces.Discrete, gym.Env                    mapf_env.py:8-9    (inert holders)
    gym.envs.toy_text.discrete.categorical_sample   mapf_env.py:10,255 (step() sampling)
    gym.utils.seeding.np_random                     mapf_env.py:11,139 (RandomState, seed 42)

`install_stubs()` registers minimal stand-ins in `sys.modules`.  `categorical_sample` restates the published
gym 0.13.0 behaviour (`requirements.txt:7` pins gym==0.13.0): `(np.cumsum(p) > rng.rand()).argmax()`.
The seeding stub returns a `RandomState(seed)`; the reference's tests never pin that stream (every `step()`
test uses `fail_prob=0`), so sampling parity is only ever checked through *explicit uniforms* (see
`UniformTape`), never through the stream itself.
"""
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("MAPF_REFERENCE_ROOT", "/root/reference")


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
    return np.random.RandomState(seed), seed


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
