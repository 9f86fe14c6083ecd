"""MapfEnv -- gym-mapf's environment class with its transition model on the GPU.

Drop-in for reference `gym_mapf/envs/mapf_env.py`: same constructor, attributes, methods, return shapes and
exceptions.  What differs is where the work happens:

    env.P[s][a]   -> rows come from the `k_count` / `k_expand` CUDA kernels (csrc/), fetched a whole state
                     (all nA actions) at a time and cached on the host        (reference mapf_env.py:448-483)
    env.step(a)   -> one call of the `k_step` kernel with the per-agent uniforms drawn from `env.np_random`
                     on the host, exactly where the reference draws them      (reference mapf_env.py:237-266)

Host-side Python below is argument marshalling and the O(n) scalar helpers of the public API
(`state_to_locations`, `locations_to_state`, `is_terminal`, ...); it never computes a transition.  Without the
built library or a CUDA device `P` and `step` raise -- there is no CPU fallback.
"""
import collections
import enum
import functools
import hashlib
import itertools
import struct

import numpy as np

from . import (ACTIONS, ACTIONS_TO_INT, POSSIBILITIES, STAY, UP, RIGHT, DOWN, LEFT,  # noqa: F401 (re-exported)
               ALL_STAY_JOINT_ACTION, integer_to_vector, vector_to_integer,
               integer_to_vector_multiple_numbers, vector_to_integer_multiple_numbers, map_name_to_files, MAPS_PATH)
from .grid import MapfGrid, EmptyCell, ObstacleCell

try:  # the reference derives from gym.Env and exposes gym spaces; both are optional here
    import gym as _gym
    from gym import spaces as _spaces
    _EnvBase = _gym.Env
    _Discrete = _spaces.Discrete
except Exception:  # noqa: BLE001 - gym is not a dependency of the engine
    _EnvBase = object

    class _Discrete:
        def __init__(self, n):
            self.n = n

        def __repr__(self):
            return "Discrete(%d)" % self.n

CELL_TO_CHAR = {EmptyCell: ".", ObstacleCell: "@"}
ACTION_TO_CHAR = {UP: "^", RIGHT: ">", DOWN: "V", LEFT: "<", STAY: "S"}
GYM_MAPF_SEED = 42


class OptimizationCriteria(enum.Enum):
    SoC = "SoC"
    Makespan = "Makespan"


# ---- single-cell motion on the host (public helper of the reference, mapf_env.py:43-94) -----------------------
_DELTA = {UP: (-1, 0), DOWN: (1, 0), RIGHT: (0, 1), LEFT: (0, -1)}


def _move(grid, loc, action):
    if action == STAY:
        return loc
    dr, dc = _DELTA[action]
    target = (min(max(loc[0] + dr, 0), len(grid) - 1), min(max(loc[1] + dc, 0), len(grid[0]) - 1))
    return loc if grid[target] is ObstacleCell else target


def execute_action(grid, s, noised_action):
    """Move every agent one cell: clamp at the border, stay in place on an obstacle (mapf_env.py:87-94)."""
    return tuple(_move(grid, loc, act) for loc, act in zip(s, noised_action))


def vector_action_to_integer(a):
    return vector_to_integer(a, [len(ACTIONS)] * len(a), lambda x: ACTIONS.index(x))


def integer_action_to_vector(a, n_agents):
    return integer_to_vector(a, [len(ACTIONS)] * n_agents, n_agents, lambda n: ACTIONS[n])


def function_to_get_item_of_object(func):
    """An object whose `obj[item]` is `func(item)` (mapf_env.py:105-112)."""

    class ret_type:
        def __getitem__(self, item):
            return func(item)

    return ret_type()


def _gym_np_random(seed):
    """`gym.utils.seeding.np_random` of gym 0.13.0 (the reference's pin), restated: a RandomState seeded with the
    32-bit words of the first 8 bytes of sha512(str(seed)).  Not pinned by any reference test."""
    seed = int(seed) % 2 ** 64
    digest = hashlib.sha512(str(seed).encode("utf8")).digest()[:8]
    big = sum(v << (32 * i) for i, v in enumerate(struct.unpack("2I", digest)))
    words = []
    while big > 0:
        big, low = divmod(big, 2 ** 32)
        words.append(low)
    rng = np.random.RandomState()
    rng.seed(words)
    return rng, seed


class MapfEnv(_EnvBase):
    # rows of P fetched per state when nA * 3**n stays below this many records, else one (s, a) row at a time
    _PREFETCH_RECORDS = 1 << 21
    _PREFETCH_STATES = 64
    _CACHE_STATES = 2048
    _CACHE_RECORDS = 1 << 23   # records held by the host-side row cache at most

    def __init__(self, grid: MapfGrid, n_agents: int, start_locations: tuple, goal_locations: tuple, fail_prob: float,
                 reward_of_collision: float, reward_of_goal: float, reward_of_living: float,
                 optimization_criteria: OptimizationCriteria, *, device=None):
        self.grid = grid
        self.agents_starts, self.agents_goals = start_locations, goal_locations
        self.n_agents = n_agents
        self.fail_prob = fail_prob
        self.right_fail = self.fail_prob / 2
        self.left_fail = self.fail_prob / 2
        self.reward_of_clash = reward_of_collision
        self.reward_of_goal = reward_of_goal
        self.reward_of_living = reward_of_living
        self.optimization_criteria = optimization_criteria
        self.np_random, self.seed = _gym_np_random(GYM_MAPF_SEED)
        self.device = device

        # cell numbering: free cells in the grid's (column-major) iteration order (mapf_env.py:142-143)
        self.valid_locations = [tuple(rc) for rc in self.grid.free_cells().tolist()]
        self.loc_to_int = {loc: i for i, loc in enumerate(self.valid_locations)}
        self.nS = len(self.valid_locations) ** self.n_agents
        self.nA = len(ACTIONS) ** self.n_agents

        self.P = function_to_get_item_of_object(self._partial_get_transitions)
        self.action_space = _Discrete(self.nA)
        self.observation_space = _Discrete(self.nS)

        self._engine_obj = None
        self._rows = collections.OrderedDict()  # state -> host CSR of all its actions (or (s, a) -> one row)
        self.reset()                                   # KeyError when a start is not a free cell
        self.locations_to_state(self.agents_goals)     # KeyError when a goal is not a free cell
        self.lastaction = None

    # ---- engine --------------------------------------------------------------------------------------------
    @property
    def engine(self):
        """The device context; created on first use.  Raises without CUDA / without the built library."""
        if self._engine_obj is None:
            from .. import _native
            self._engine_obj = _native.Engine(
                self.grid.obstacles, self.n_agents, self.agents_starts, self.agents_goals, self.fail_prob,
                self.reward_of_clash, self.reward_of_goal, self.reward_of_living,
                self.optimization_criteria == OptimizationCriteria.Makespan,
                device=0 if self.device is None else self.device)
        return self._engine_obj

    def __copy__(self):
        # copy(env) must give a steppable env (mapf_env_tests.py:92-105); the immutable device context is shared
        new = object.__new__(type(self))
        new.__dict__.update(self.__dict__)
        new.P = function_to_get_item_of_object(new._partial_get_transitions)
        new._rows = collections.OrderedDict()
        new._rows_records = 0
        new.__dict__.pop("_step_buf", None)
        return new

    # ---- encodings (host scalars; the bulk versions are VecMapfEnv.state_to_cells / cells_to_state) ---------
    def state_to_locations(self, state):
        return integer_to_vector(state, [len(self.valid_locations)] * self.n_agents, self.n_agents,
                                 lambda x: self.valid_locations[x])

    def locations_to_state(self, locs):
        if self.n_agents != len(locs):
            raise AssertionError(f'{locs} locations number is different than the number of agents {self.n_agents}')
        cells = tuple(self.loc_to_int[loc] for loc in locs)
        return vector_to_integer(cells, [len(self.valid_locations)] * len(cells), lambda x: x)

    def is_terminal(self, s):
        """`s` is a tuple of locations: two agents share a cell, or everyone is on its goal (mapf_env.py:210-223)."""
        if len(set(s)) != len(s):
            return True
        return all(loc == self.agents_goals[i] for i, loc in enumerate(s))

    # ---- per-agent outcomes, read back from the device move table --------------------------------------------
    def single_agent_movements(self, local_state, a):
        """[(local_state, next_local_state, prob), ...] for one agent (mapf_env.py:163-184)."""
        k, dest, prob, _ = self.engine.moves()
        return [(local_state, int(dest[local_state, a, j]), float(prob[local_state, a, j]))
                for j in range(int(k[local_state, a]))]

    def get_possible_actions(self, a):
        """All noised joint actions of `a` with their probabilities (mapf_env.py:186-208; unused by the env)."""
        keep = 1.0 - self.right_fail - self.left_fail
        right, left = POSSIBILITIES[a[-1]]
        out = [(self.right_fail, (right,)), (self.left_fail, (left,)), (keep, (a[-1],))]
        for head in reversed(a[:-1]):
            right, left = POSSIBILITIES[head]
            nxt = []
            for prob, tail in out:
                nxt += [(self.right_fail * prob, (right,) + tail), (self.left_fail * prob, (left,) + tail),
                        (keep * prob, (head,) + tail)]
            out = nxt
        return out

    def calc_transition_reward_from_local_states(self, prev_local_states, action: int, next_local_states):
        """(reward, done, collision) of ONE given joint transition (mapf_env.py:225-235, 378-389, 436-446).
        Scalar public helper kept for API compatibility: it accepts arbitrary (even unreachable) cell tuples, so
        it is evaluated here on the host; `P` and `step` never call it -- their rewards come from the kernels."""
        n = self.n_agents
        goals = [self.loc_to_int[g] for g in self.agents_goals]
        if self.optimization_criteria == OptimizationCriteria.Makespan:
            living = self.reward_of_living
        else:
            acts = integer_action_to_vector(action, n)
            parked = sum(1 for i in range(n) if prev_local_states[i] == goals[i] and acts[i] == STAY)
            living = (n - parked) * self.reward_of_living
        for i, j in itertools.combinations(range(n), 2):
            swap = prev_local_states[i] == next_local_states[j] and prev_local_states[j] == next_local_states[i]
            if swap or next_local_states[i] == next_local_states[j]:
                return self.reward_of_clash + living, True, True
        if all(goals[i] == next_local_states[i] for i in range(n)):
            return self.reward_of_goal + living, True, False
        return living, False, False

    # ---- the transition table -------------------------------------------------------------------------------
    def _partial_get_transitions(self, s):
        return function_to_get_item_of_object(functools.partial(self._get_transitions, s))

    def _fetch(self, key, producer):
        got = self._rows.get(key)
        if got is None:
            got = producer()
            self._rows[key] = got
            self._rows_records = getattr(self, "_rows_records", 0) + len(got[2])
            # bounded by entries AND by records held (a block of C2 rows is ~1e6 records)
            while len(self._rows) > 1 and (len(self._rows) > self._CACHE_STATES or self._rows_records > self._CACHE_RECORDS):
                _, old = self._rows.popitem(last=False)
                self._rows_records -= len(old[2])
        else:
            self._rows.move_to_end(key)
        return got

    def _host_csr(self, csr):
        """Device CSR -> host arrays ready for row slicing: (row_ptr, next states, prob, reward, done, collision)."""
        row_ptr, ns, prob, reward, flags = csr
        flags = flags.cpu().numpy()
        if self.engine.words == 1:
            ns = ns.cpu().numpy().view(np.uint64)
        else:
            ns = np.array(self.engine.states_to_ints(ns), dtype=object)
        return (row_ptr.cpu().numpy().tolist(), ns, prob.cpu().numpy(), reward.cpu().numpy(), (flags & 1).astype(bool),
                (flags & 2).astype(bool))

    def _get_transitions(self, s, a):
        """P[s][a]: list of ((prob, collision), next_state, reward, done) in itertools.product order
        (mapf_env.py:448-479), produced by the CUDA expand kernel."""
        s, a = int(s), int(a)
        if not 0 <= s < self.nS or not 0 <= a < self.nA:
            raise IndexError("state %d / action %d outside [0, %d) x [0, %d)" % (s, a, self.nS, self.nA))
        eng = self.engine
        if self.nA * eng.max_row_len <= self._PREFETCH_RECORDS:
            # a block of consecutive states per GPU call (planners sweep s in order): one launch sequence and one
            # read-back serve up to _PREFETCH_STATES states
            k = max(1, min(self._PREFETCH_STATES, self._PREFETCH_RECORDS // (self.nA * eng.max_row_len)))
            s0 = s - s % k
            n_states = min(k, self.nS - s0)
            row_ptr, ns, prob, reward, done, coll = self._fetch(("block", s0),
                                                                lambda: self._host_csr(eng.table_range(s0, n_states)))
            row = (s - s0) * self.nA + a
            lo, hi = row_ptr[row], row_ptr[row + 1]
        else:
            def one_row():
                import torch
                acts = torch.tensor([a], dtype=torch.int32, device=eng.torch_device)
                return self._host_csr(eng.transitions(eng.states_from_ints([s]), acts))
            row_ptr, ns, prob, reward, done, coll = self._fetch((s, a), one_row)
            lo, hi = 0, row_ptr[1]
        # slices -> Python scalars in C (tolist), then one zip: a few hundred nanoseconds per record
        return list(zip(zip(prob[lo:hi].tolist(), coll[lo:hi].tolist()), ns[lo:hi].tolist(), reward[lo:hi].tolist(),
                        done[lo:hi].tolist()))

    # ---- sampled step ----------------------------------------------------------------------------------------
    def step(self, a: int):
        """One sampled joint transition (mapf_env.py:237-266) computed by the step kernel.  The per-agent uniforms
        are drawn here from `self.np_random`, one per agent in agent order and none for a terminal state, so the
        random stream is consumed exactly as the reference consumes it."""
        buf = self.__dict__.get("_step_buf")
        if buf is None:  # reused across calls: the scalar path is dominated by fixed costs (one launch, one sync)
            eng = self.engine
            from .._native import check, lib
            state, action = np.zeros(eng.words, np.uint64), np.zeros(1, np.int32)
            uniforms = np.zeros(self.n_agents, np.float64)
            out = (np.zeros(eng.words, np.uint64), np.zeros(1, np.float64), np.zeros(1, np.float64),
                   np.zeros(1, np.uint8), np.zeros(1, np.uint8))
            # the C-ABI call with its arguments marshalled once (the arrays above never move)
            call = functools.partial(lib().mapf_step_host, eng._h, state.ctypes.data, action.ctypes.data, 1,
                                     uniforms.ctypes.data, 0, 0, 0, 0, *(o.ctypes.data for o in out))
            buf = self._step_buf = dict(state=state, action=action, uniforms=uniforms, out=out, call=call, check=check,
                                        words=eng.words, known=(None, False))
        # is_terminal(current state): known from the previous step unless the caller moved the env (env.s = ..., reset)
        # or that step ended in a clash (a swap leaves a non-terminal state, a vertex clash a terminal one)
        known_s, known_flag = buf["known"]
        if known_s == self.s:
            terminal = known_flag
        else:
            terminal = self.is_terminal(self.state_to_locations(self.s))
            buf["known"] = (self.s, terminal)
        uniforms = buf["uniforms"]
        if terminal:
            uniforms[:] = 0.0
        else:
            uniforms[:] = self.np_random.random_sample(self.n_agents)  # n successive rand() draws, in agent order
        state = buf["state"]
        state[0] = self.s & 0xFFFFFFFFFFFFFFFF
        if buf["words"] == 2:
            state[1] = self.s >> 64
        buf["action"][0] = a % self.nA
        rc = buf["call"]()
        if rc:
            buf["check"](rc)
        ns, reward, prob, done, coll = buf["out"]
        new_state = int(ns[0]) | (int(ns[1]) << 64 if buf["words"] == 2 else 0)
        self.lastaction = a
        if terminal:
            return self.s, 0, True, {"prob": 0}
        self.s = new_state
        is_done, is_coll = bool(done[0]), bool(coll[0])
        # not done -> distinct cells, not all on their goals; goal reached -> terminal; clash -> decided on the next call
        buf["known"] = (None, False) if is_coll else (new_state, is_done)
        return new_state, float(reward[0]), is_done, {"prob": float(prob[0]), "collision": is_coll}

    def reset(self):
        self.lastaction = None
        self.s = self.locations_to_state(self.agents_starts)
        return self.s

    # ---- reverse neighbours (host helper; not on the accelerated path -- SURVEY 8f) --------------------------
    def _single_location_predecessors(self, loc):
        # cells from which `loc` is reached by DOWN, UP, LEFT, RIGHT, STAY respectively (mapf_env.py:414-425)
        return [_move(self.grid, loc, act) for act in (DOWN, UP, LEFT, RIGHT, STAY)]

    def predecessors(self, s: int):
        per_agent = [self._single_location_predecessors(loc) for loc in self.state_to_locations(s)]
        return set(self.locations_to_state(combo) for combo in itertools.product(*per_agent))

    # ---- text rendering (mapf_env.py:295-356) -----------------------------------------------------------------
    def render(self, mode="human"):
        red, green, yellow, blue, off = "\033[31m", "\033[32m", "\033[33m", "\033[34m", "\033[39m"
        agents = self.state_to_locations(self.s)
        for r in range(len(self.grid)):
            cells = []
            for c in range(len(self.grid[0])):
                here = [i for i, loc in enumerate(agents) if loc == (r, c)]
                if len(here) > 1:
                    cells.append(red + "*" + off)
                elif here:
                    on_goal = self.agents_goals[here[0]] == (r, c)
                    cells.append((green if on_goal else yellow) + str(here[0]) + off)
                elif (r, c) in self.agents_goals:
                    cells.append(blue + str(self.agents_goals.index((r, c))) + off)
                else:
                    cells.append(CELL_TO_CHAR[self.grid[r, c]])
            print(" ".join(cells) + " ")

    def render_with_policy(self, agent: int, policy):
        green, yellow, blue, off = "\033[32m", "\033[33m", "\033[34m", "\033[39m"
        agents = self.state_to_locations(self.s)
        here, goal = agents[agent], self.agents_goals[agent]
        print("")
        for r in range(len(self.grid)):
            print("")
            for c in range(len(self.grid[0])):
                if (r, c) == here:
                    print((green if here == goal else yellow) + str(agent) + off, end=" ")
                elif (r, c) == goal:
                    print(blue + str(agent) + off, end=" ")
                else:
                    moved = agents[:agent] + ((r, c),) + agents[agent + 1:]
                    joint = policy(self.locations_to_state(moved))
                    print(ACTION_TO_CHAR[integer_action_to_vector(joint, self.n_agents)[agent]], end=" ")
        print("")
