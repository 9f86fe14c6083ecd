"""TEST INFRASTRUCTURE ONLY -- ctypes binding of the plain-C oracle (oracle/mapf_oracle.c).

Used by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.  Never imported by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libmapf_oracle.so")
_lib = None

u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build(force=False):
    src = os.path.join(HERE, "mapf_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s"])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.c_int, C.c_int, u8p, C.c_int, i32p, C.c_double, C.c_double, C.c_double,
                                    C.c_double, C.c_int]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_num_cells.argtypes = [C.c_void_p]
        L.oracle_moves.argtypes = [C.c_void_p, u8p, i32p, f64p]
        L.oracle_decode_states.argtypes = [C.c_void_p, C.c_int64, u64p, u64p, i32p]
        L.oracle_encode_states.argtypes = [C.c_void_p, C.c_int64, i32p, u64p, u64p]
        L.oracle_count_rows.restype = C.c_int64
        L.oracle_count_rows.argtypes = [C.c_void_p, C.c_int64, u64p, u64p, i64p, i64p]
        L.oracle_expand.argtypes = [C.c_void_p, C.c_int64, u64p, u64p, i64p, i64p, u64p, u64p, f64p, f64p, u8p, u8p]
        L.oracle_expand_mt.argtypes = L.oracle_expand.argtypes + [C.c_int]
        L.oracle_table_checksums.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int64, u64p]
        L.oracle_step.argtypes = [C.c_void_p, C.c_int64, u64p, u64p, i64p, f64p, u64p, u64p, f64p, f64p, u8p, u8p,
                                  u8p]
        L.oracle_step_mt.argtypes = L.oracle_step.argtypes + [C.c_int]
        L.oracle_backup.argtypes = [C.c_void_p, C.c_int64, u64p, u64p, i64p, f64p, C.c_double, f64p]
        L.oracle_backup_mt.argtypes = L.oracle_backup.argtypes + [C.c_int]
        L.oracle_count_predecessors.restype = C.c_int64
        L.oracle_count_predecessors.argtypes = [C.c_void_p, C.c_int64, u64p, u64p, i64p]
        L.oracle_predecessors.argtypes = [C.c_void_p, C.c_int64, u64p, u64p, i64p, u64p, u64p]
        L.oracle_project_states.argtypes = [C.c_void_p, C.c_int64, u64p, u64p, C.c_int, i32p, u64p, u64p]
        _lib = L
    return _lib


class COracle:
    """One env spec held by the C oracle.  `rows` are '.'/'@' strings; goals are (row, col) pairs."""

    def __init__(self, rows, n_agents, goals, fail_prob, r_clash, r_goal, r_living, soc):
        rows = [r.strip() for r in rows]
        self.H, self.W, self.n = len(rows), len(rows[0]), n_agents
        obst = np.array([[1 if ch == "@" else 0 for ch in r] for r in rows], dtype=np.uint8)
        goal_rc = np.array(goals, dtype=np.int32).reshape(-1)
        self._h = lib().oracle_create(self.H, self.W, np.ascontiguousarray(obst), n_agents, goal_rc, float(fail_prob),
                                      float(r_clash), float(r_goal), float(r_living), 1 if soc else 0)
        if not self._h:
            raise KeyError("goal on an obstacle / off the grid, or unsupported agent count")
        self.L = lib().oracle_num_cells(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_destroy(self._h)
            self._h = None

    def moves(self):
        k = np.zeros((self.L, 5), np.uint8)
        dest = np.zeros((self.L, 5, 3), np.int32)
        prob = np.zeros((self.L, 5, 3), np.float64)
        lib().oracle_moves(self._h, k, dest, prob)
        return k, dest, prob

    def decode(self, lo, hi):
        ids = np.zeros((len(lo), self.n), np.int32)
        lib().oracle_decode_states(self._h, len(lo), np.ascontiguousarray(lo), np.ascontiguousarray(hi), ids)
        return ids

    def encode(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        lo = np.zeros(len(ids), np.uint64)
        hi = np.zeros(len(ids), np.uint64)
        lib().oracle_encode_states(self._h, len(ids), ids, lo, hi)
        return lo, hi

    def rows(self, s_lo, s_hi, action, threads=1):
        """CSR expansion of P[s][a] for every pair -> dict like the golden `rows_*` fixtures."""
        s_lo = np.ascontiguousarray(s_lo, np.uint64)
        s_hi = np.ascontiguousarray(s_hi, np.uint64)
        action = np.ascontiguousarray(action, np.int64)
        B = len(s_lo)
        row_len = np.zeros(B, np.int64)
        total = lib().oracle_count_rows(self._h, B, s_lo, s_hi, action, row_len)
        row_ptr = np.zeros(B + 1, np.int64)
        np.cumsum(row_len, out=row_ptr[1:])
        out = dict(row_ptr=row_ptr, next_lo=np.zeros(total, np.uint64), next_hi=np.zeros(total, np.uint64),
                   prob=np.zeros(total, np.float64), reward=np.zeros(total, np.float64),
                   done=np.zeros(total, np.uint8), collision=np.zeros(total, np.uint8))
        args = (self._h, B, s_lo, s_hi, action, row_ptr, out["next_lo"], out["next_hi"], out["prob"], out["reward"],
                out["done"], out["collision"])
        if threads > 1:
            lib().oracle_expand_mt(*args, threads)
        else:
            lib().oracle_expand(*args)
        return out

    def table_checksums(self, s_begin, n_states):
        out = np.zeros(8, np.uint64)
        lib().oracle_table_checksums(self._h, s_begin & ((1 << 64) - 1), s_begin >> 64, n_states, out)
        keys = ["count", "n_collision", "n_done", "sum_next_lo", "sum_next_hi", "sum_prob_bits", "sum_reward_bits",
                "ordered"]
        return dict(zip(keys, (int(x) for x in out)))

    def step(self, s_lo, s_hi, action, uniforms, threads=1):
        s_lo = np.ascontiguousarray(s_lo, np.uint64)
        s_hi = np.ascontiguousarray(s_hi, np.uint64)
        action = np.ascontiguousarray(action, np.int64)
        uniforms = np.ascontiguousarray(uniforms, np.float64)
        B = len(s_lo)
        out = dict(next_lo=np.zeros(B, np.uint64), next_hi=np.zeros(B, np.uint64), reward=np.zeros(B, np.float64),
                   prob=np.zeros(B, np.float64), done=np.zeros(B, np.uint8), collision=np.zeros(B, np.uint8),
                   terminal=np.zeros(B, np.uint8))
        args = (self._h, B, s_lo, s_hi, action, uniforms, out["next_lo"], out["next_hi"], out["reward"], out["prob"],
                out["done"], out["collision"], out["terminal"])
        if threads > 1:
            lib().oracle_step_mt(*args, threads)
        else:
            lib().oracle_step(*args)
        return out

    # ---- rows built after the hot path
    def backup(self, s_lo, s_hi, action, V, gamma, threads=1):
        """Q[b] = sum over P[s_b][a_b] of p * (r + gamma * V[s2]) in the row's order (states must fit one word)."""
        s_lo = np.ascontiguousarray(s_lo, np.uint64)
        s_hi = np.ascontiguousarray(s_hi, np.uint64)
        action = np.ascontiguousarray(action, np.int64)
        V = np.ascontiguousarray(V, np.float64)
        Q = np.zeros(len(s_lo), np.float64)
        if threads > 1:
            lib().oracle_backup_mt(self._h, len(s_lo), s_lo, s_hi, action, V, float(gamma), Q, threads)
        else:
            lib().oracle_backup(self._h, len(s_lo), s_lo, s_hi, action, V, float(gamma), Q)
        return Q

    def predecessors(self, s_lo, s_hi):
        """CSR of env.predecessors(s) for every state, each row sorted ascending."""
        s_lo = np.ascontiguousarray(s_lo, np.uint64)
        s_hi = np.ascontiguousarray(s_hi, np.uint64)
        B = len(s_lo)
        row_len = np.zeros(B, np.int64)
        total = lib().oracle_count_predecessors(self._h, B, s_lo, s_hi, row_len)
        row_ptr = np.zeros(B + 1, np.int64)
        np.cumsum(row_len, out=row_ptr[1:])
        p_lo, p_hi = np.zeros(total, np.uint64), np.zeros(total, np.uint64)
        lib().oracle_predecessors(self._h, B, s_lo, s_hi, row_ptr, p_lo, p_hi)
        return dict(row_ptr=row_ptr, pred_lo=p_lo, pred_hi=p_hi)

    def project(self, s_lo, s_hi, agents):
        s_lo = np.ascontiguousarray(s_lo, np.uint64)
        s_hi = np.ascontiguousarray(s_hi, np.uint64)
        agents = np.ascontiguousarray(agents, np.int32)
        o_lo, o_hi = np.zeros(len(s_lo), np.uint64), np.zeros(len(s_lo), np.uint64)
        lib().oracle_project_states(self._h, len(s_lo), s_lo, s_hi, len(agents), agents, o_lo, o_hi)
        return o_lo, o_hi
