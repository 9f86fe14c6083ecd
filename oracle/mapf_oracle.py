"""TEST INFRASTRUCTURE ONLY -- pure-Python CPU restatement of gym-mapf's joint-transition path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may import
this module, and only as the checker.  Nothing under `gym_mapf_b200/` imports it.

Parity status: PINNED.  `tests/test_oracle_golden.py` checks every function below against fixtures in
`tests/golden/` that `oracle/make_golden.py` produced by running the *unmodified* reference
(`/root/reference/gym_mapf`, through `oracle/ref_shim.py`) in the build container, and against the
reference's own known-answer tests.

All arithmetic is Python int (arbitrary precision) and Python float (IEEE binary64), evaluated in the
reference's order, so the outputs are bit-identical to the reference's.  Each function cites the reference
file:line it restates; paths are relative to `/root/reference/gym_mapf/envs/`.
"""
from itertools import product

# action ids: `__init__.py:26` ACTIONS = [STAY, UP, RIGHT, DOWN, LEFT]
STAY, UP, RIGHT, DOWN, LEFT = 0, 1, 2, 3, 4
N_ACTIONS = 5
# (right-slip, left-slip) per intended action: `__init__.py:19-25` POSSIBILITIES
SLIP = {UP: (RIGHT, LEFT), DOWN: (LEFT, RIGHT), LEFT: (UP, DOWN), RIGHT: (DOWN, UP), STAY: (STAY, STAY)}


def to_digits(x, radix, n):
    """Little-endian fixed-radix digits, agent 0 least significant (`__init__.py:50-67`)."""
    out = []
    for _ in range(n):
        out.append(x % radix)
        x //= radix
    return out


def from_digits(digits, radix):
    """Inverse of `to_digits` (`__init__.py:70-79`)."""
    total, weight = 0, 1
    for d in digits:
        total += d * weight
        weight *= radix
    return total


class OracleSpec:
    """Immutable description of one env: grid, agents, noise and rewards (`mapf_env.py:116-161`)."""

    def __init__(self, rows, n_agents, starts, goals, fail_prob, r_clash, r_goal, r_living, soc):
        self.rows = [ln.strip() for ln in rows]
        for ln in self.rows:
            for ch in ln:
                if ch not in ".@":
                    raise KeyError(ch)  # `grid.py:9-13,21`
        self.H, self.W = len(self.rows), len(self.rows[0])
        self.n = n_agents
        self.fail_prob = fail_prob
        self.right_fail = fail_prob / 2  # `mapf_env.py:131-132`
        self.left_fail = fail_prob / 2
        self.r_clash, self.r_goal, self.r_living, self.soc = r_clash, r_goal, r_living, bool(soc)
        # cell numbering: free cells in COLUMN-major order (`grid.py:37-40`, `mapf_env.py:142-143`)
        self.cells = [(r, c) for c in range(self.W) for r in range(self.H) if self.rows[r][c] == "."]
        self.cell_id = {rc: i for i, rc in enumerate(self.cells)}
        self.L = len(self.cells)
        self.nS = self.L ** self.n  # `mapf_env.py:145-146`
        self.nA = N_ACTIONS ** self.n
        self.starts = tuple(tuple(x) for x in starts)
        self.goals = tuple(tuple(x) for x in goals)
        if len(self.starts) != self.n or len(self.goals) != self.n:
            raise AssertionError("location count differs from agent count")  # `mapf_env.py:366-367`
        self.start_ids = [self.cell_id[rc] for rc in self.starts]  # KeyError on an obstacle, `:155,369`
        self.goal_ids = [self.cell_id[rc] for rc in self.goals]  # `:158`
        self.s0 = from_digits(self.start_ids, self.L)
        self._moves = {}

    # -- single-cell motion: `mapf_env.py:43-84`
    def shift(self, rc, d):
        r, c = rc
        if d == UP:
            t = (max(0, r - 1), c)
        elif d == DOWN:
            t = (min(self.H - 1, r + 1), c)
        elif d == RIGHT:
            t = (r, min(self.W - 1, c + 1))
        elif d == LEFT:
            t = (r, max(0, c - 1))
        else:
            return rc  # STAY has no obstacle test (`:74-75`)
        return rc if self.rows[t[0]][t[1]] == "@" else t

    # -- per-agent stochastic outcomes: `mapf_env.py:163-184`
    def agent_outcomes(self, cell, a):
        """-> list of (next_cell_id, prob): [intended, right-slip, left-slip], zero-probability candidates
        dropped, equal destinations merged into their first occurrence (probabilities added in list order)."""
        key = (cell, a)
        got = self._moves.get(key)
        if got is not None:
            return got
        slip_r, slip_l = SLIP[a]
        cand = [(1 - self.right_fail - self.left_fail, a), (self.right_fail, slip_r), (self.left_fail, slip_l)]
        dest, prob = [], []
        for p, d in cand:
            if not p > 0:
                continue
            nxt = self.cell_id[self.shift(self.cells[cell], d)]
            if nxt in dest:
                j = dest.index(nxt)
                prob[j] = prob[j] + p
            else:
                dest.append(nxt)
                prob.append(p)
        got = list(zip(dest, prob))
        self._moves[key] = got
        return got

    # -- `mapf_env.py:210-223`
    def is_terminal(self, ids):
        if len(set(ids)) != len(ids):
            return True
        return all(ids[i] == self.goal_ids[i] for i in range(self.n))

    # -- `mapf_env.py:436-446`
    def living(self, prev, acts):
        if not self.soc:
            return self.r_living
        parked = sum(1 for i in range(self.n) if prev[i] == self.goal_ids[i] and acts[i] == STAY)
        return (self.n - parked) * self.r_living

    # -- `mapf_env.py:378-389`
    def clash(self, prev, nxt):
        for i in range(self.n):
            for j in range(i + 1, self.n):
                if prev[i] == nxt[j] and prev[j] == nxt[i]:
                    return True
                if nxt[i] == nxt[j]:
                    return True
        return False

    # -- `mapf_env.py:225-235` -> (reward, done, collision)
    def judge(self, prev, acts, nxt):
        live = self.living(prev, acts)
        if self.clash(prev, nxt):
            return self.r_clash + live, True, True
        if all(nxt[i] == self.goal_ids[i] for i in range(self.n)):
            return self.r_goal + live, True, False
        return live, False, False

    # -- `mapf_env.py:448-479`; records are (prob, collision, next_state, reward, done)
    def row(self, s, a):
        prev = to_digits(s, self.L, self.n)
        if self.is_terminal(prev):
            return [(1.0, False, s, 0, True)]
        acts = to_digits(a, N_ACTIONS, self.n)
        per_agent = [self.agent_outcomes(prev[i], acts[i]) for i in range(self.n)]
        out = []
        for combo in product(*per_agent):  # agent 0 varies slowest (`:467`)
            p = combo[0][1]
            for m in combo[1:]:
                p = p * m[1]  # left-to-right (`:468`)
            nxt = [m[0] for m in combo]
            reward, done, coll = self.judge(prev, acts, nxt)
            out.append((p, coll, from_digits(nxt, self.L), reward, done))
        return out

    # -- `mapf_env.py:237-266`; `uniforms` supplies the draws `categorical_sample` would make (`:255`)
    def step(self, s, a, uniforms):
        """-> (next_state, reward, done, prob, collision, draws_used); a terminal state is a no-op that
        returns prob 0 and consumes no draw (`:238-240`)."""
        prev = to_digits(s, self.L, self.n)
        if self.is_terminal(prev):
            return s, 0, True, 0, None, 0
        acts = to_digits(a, N_ACTIONS, self.n)
        nxt, total = [], 1
        for i in range(self.n):
            outs = self.agent_outcomes(prev[i], acts[i])
            u = uniforms[i]
            acc, pick, found = 0.0, 0, False
            for j, (_, p) in enumerate(outs):  # np.cumsum, then first index whose cumsum > u, else 0
                acc = p if j == 0 else acc + p
                if not found and acc > u:
                    pick, found = j, True
            nxt.append(outs[pick][0])
            total = total * outs[pick][1]
        reward, done, coll = self.judge(prev, acts, nxt)
        return from_digits(nxt, self.L), reward, done, total, coll, self.n

    # -- `mapf_env.py:373-376, 414-434`
    def predecessors(self, s):
        ids = to_digits(s, self.L, self.n)
        per_agent = []
        for cid in ids:
            rc = self.cells[cid]
            # reverse moves tried in the order DOWN, UP, LEFT, RIGHT, STAY (`:416-420`); the free-cell filter
            # of `:422-423` never drops anything because a blocked move already stays in place
            per_agent.append([self.cell_id[self.shift(rc, d)] for d in (DOWN, UP, LEFT, RIGHT, STAY)])
        return {from_digits(combo, self.L) for combo in product(*per_agent)}


    # -- the consumer loop of a planner over P (gym value-iteration backup on `mapf_env.py:448-479`'s rows)
    def backup(self, s, a, V, gamma):
        acc = 0  # explicit left-to-right adds: the builtin sum() of floats is compensated since Python 3.12
        for (p, _c, s2, r, _d) in self.row(s, a):
            acc = acc + p * (r + gamma * V[s2])
        return acc

    # -- `utils.py:138-157`: the state as the sub-env of `agents` numbers it
    def project(self, s, agents):
        ids = to_digits(s, self.L, self.n)
        return from_digits([ids[i] for i in agents], self.L)


def spec_from_reference_env(env, soc):
    """Build an OracleSpec from a live reference `MapfEnv` (golden generation only)."""
    h = len(env.grid)
    w = len(env.grid[0])
    free = set(env.valid_locations)
    rows = ["".join("." if (r, c) in free else "@" for c in range(w)) for r in range(h)]
    return OracleSpec(rows, env.n_agents, env.agents_starts, env.agents_goals, env.fail_prob,
                      env.reward_of_clash, env.reward_of_goal, env.reward_of_living, soc)


# ---- on-disk formats; pinned by tests/golden/formats.json (oracle/make_golden_formats.py) -------
def parse_map_text(data: bytes):
    """Rows ('.'/'@' strings) of a MovingAI .map file's contents: text-mode `readlines()[4:]` (`utils.py:33-37`, universal
    newlines), every line `strip()`ped and every character looked up in CHAR_TO_CELL (`grid.py:9-13,19-22`: KeyError for
    anything but '.' and '@')."""
    text = data.decode("latin-1").replace("\r\n", "\n").replace("\r", "\n")
    lines = text.split("\n")
    if lines and lines[-1] == "":
        lines.pop()  # readlines() has no empty line behind a final terminator
    rows = []
    for line in lines[4:]:
        line = line.strip()
        for ch in line:
            if ch not in ".@":
                raise KeyError(ch)
        rows.append(line)
    return rows


def parse_scen_text(data: bytes, n_agents: int):
    """(starts, goals) of a .scen file's contents (`utils.py:8-30`): the first line is skipped, every further line is
    unpacked into exactly nine tab-separated fields (ValueError otherwise), fields 4..7 are `int()`ed and used as
    (row, col) pairs; reading stops after `n_agents` lines."""
    text = data.decode("latin-1").replace("\r\n", "\n").replace("\r", "\n")
    lines = text.split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    if not lines:
        raise StopIteration
    starts, goals = [], []
    for i, line in enumerate(lines[1:]):
        fields = line.split("\t")
        if len(fields) != 9:
            raise ValueError("expected 9 fields, got %d" % len(fields))
        starts.append((int(fields[4]), int(fields[5])))
        goals.append((int(fields[6]), int(fields[7])))
        if i == n_agents - 1:
            break
    return tuple(starts), tuple(goals)
