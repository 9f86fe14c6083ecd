/* TEST INFRASTRUCTURE ONLY -- plain-C CPU restatement of gym-mapf's joint-transition path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library (oracle/_build/libmapf_oracle.so), and only as the checker or as the timed CPU baseline.  Nothing under
 * gym-mapf_b200/ links or loads it; the product path has no CPU fallback.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every entry point against the .npz fixtures in tests/golden/,
 * which oracle/make_golden.py produced by running the unmodified reference (/root/reference/gym_mapf) in the
 * build container.  The reference is pure Python and cannot be compiled into oracle/_ref (there is no C/C++
 * source in it), so this port is what runs on the GPU box.
 *
 * Each function cites the reference file:line it restates (paths relative to /root/reference/gym_mapf/envs/).
 * Floating point: every operation is a single IEEE binary64 add or multiply in the reference's order; build
 * with -ffp-contract=off (see oracle/Makefile) so no multiply-add is fused.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

enum { STAY = 0, UP = 1, RIGHT = 2, DOWN = 3, LEFT = 4 }; /* __init__.py:26 */
#define MAX_AGENTS 16

typedef struct {
    int k;          /* merged outcomes, 1..3 (0 never happens for fail_prob in [0,1]) */
    int32_t dest[3];
    double prob[3];
} agent_moves;

typedef struct oracle_env {
    int H, W, n, L, soc;
    uint8_t *obst;       /* row-major, 1 = '@' */
    int32_t *cell_of;    /* (r*W+c) -> id or -1 */
    int32_t *rc_of;      /* id -> r*W+c */
    int32_t goal[MAX_AGENTS];
    double right_fail, left_fail, r_clash, r_goal, r_living;
    agent_moves *moves;  /* [L*5] */
} oracle_env;

/* mapf_env.py:43-84 -- one-cell move: clamp to the grid, then an obstacle means "stay" */
static int shift_cell(const oracle_env *e, int r, int c, int d) {
    int tr = r, tc = c;
    if (d == UP) tr = r - 1 < 0 ? 0 : r - 1;
    else if (d == DOWN) tr = r + 1 > e->H - 1 ? e->H - 1 : r + 1;
    else if (d == RIGHT) tc = c + 1 > e->W - 1 ? e->W - 1 : c + 1;
    else if (d == LEFT) tc = c - 1 < 0 ? 0 : c - 1;
    else return r * e->W + c;
    if (e->obst[tr * e->W + tc]) return r * e->W + c;
    return tr * e->W + tc;
}

/* mapf_env.py:163-184 with __init__.py:19-25 */
static void build_moves(oracle_env *e) {
    static const int slip_r[5] = {STAY, RIGHT, DOWN, LEFT, UP};
    static const int slip_l[5] = {STAY, LEFT, UP, RIGHT, DOWN};
    for (int cell = 0; cell < e->L; ++cell) {
        int r = e->rc_of[cell] / e->W, c = e->rc_of[cell] % e->W;
        for (int a = 0; a < 5; ++a) {
            agent_moves *m = &e->moves[cell * 5 + a];
            double p[3] = {1 - e->right_fail - e->left_fail, e->right_fail, e->left_fail};
            int d[3] = {a, slip_r[a], slip_l[a]};
            m->k = 0;
            for (int j = 0; j < 3; ++j) {
                if (!(p[j] > 0)) continue;
                int nxt = e->cell_of[shift_cell(e, r, c, d[j])];
                int at = -1;
                for (int q = 0; q < m->k; ++q)
                    if (m->dest[q] == nxt) { at = q; break; }
                if (at >= 0) m->prob[at] = m->prob[at] + p[j];
                else { m->dest[m->k] = nxt; m->prob[m->k] = p[j]; m->k++; }
            }
        }
    }
}

oracle_env *oracle_create(int H, int W, const uint8_t *obstacles, int n, const int32_t *goal_rc,
                          double fail_prob, double r_clash, double r_goal, double r_living, int soc) {
    if (n < 1 || n > MAX_AGENTS) return NULL;
    oracle_env *e = (oracle_env *)calloc(1, sizeof(*e));
    e->H = H; e->W = W; e->n = n; e->soc = soc;
    e->right_fail = fail_prob / 2; e->left_fail = fail_prob / 2; /* mapf_env.py:131-132 */
    e->r_clash = r_clash; e->r_goal = r_goal; e->r_living = r_living;
    e->obst = (uint8_t *)malloc((size_t)H * W);
    for (int i = 0; i < H * W; ++i) e->obst[i] = obstacles[i] ? 1 : 0;
    e->cell_of = (int32_t *)malloc(sizeof(int32_t) * H * W);
    e->rc_of = (int32_t *)malloc(sizeof(int32_t) * H * W);
    int L = 0;
    for (int c = 0; c < W; ++c)       /* column-major numbering: grid.py:37-40, mapf_env.py:142-143 */
        for (int r = 0; r < H; ++r) {
            if (e->obst[r * W + c]) e->cell_of[r * W + c] = -1;
            else { e->cell_of[r * W + c] = L; e->rc_of[L] = r * W + c; L++; }
        }
    e->L = L;
    for (int i = 0; i < n; ++i) {
        int r = goal_rc[2 * i], c = goal_rc[2 * i + 1];
        if (r < 0 || r >= H || c < 0 || c >= W || e->cell_of[r * W + c] < 0) { /* KeyError, mapf_env.py:158 */
            free(e->obst); free(e->cell_of); free(e->rc_of); free(e);
            return NULL;
        }
        e->goal[i] = e->cell_of[r * W + c];
    }
    e->moves = (agent_moves *)malloc(sizeof(agent_moves) * (size_t)L * 5);
    build_moves(e);
    return e;
}

void oracle_destroy(oracle_env *e) {
    if (!e) return;
    free(e->obst); free(e->cell_of); free(e->rc_of); free(e->moves); free(e);
}

int oracle_num_cells(const oracle_env *e) { return e->L; }

int32_t oracle_cell_id(const oracle_env *e, int r, int c) {
    if (r < 0 || r >= e->H || c < 0 || c >= e->W) return -1;
    return e->cell_of[r * e->W + c];
}

void oracle_moves(const oracle_env *e, uint8_t *k, int32_t *dest, double *prob) {
    for (int i = 0; i < e->L * 5; ++i) {
        k[i] = (uint8_t)e->moves[i].k;
        for (int j = 0; j < 3; ++j) {
            dest[i * 3 + j] = j < e->moves[i].k ? e->moves[i].dest[j] : -1;
            prob[i * 3 + j] = j < e->moves[i].k ? e->moves[i].prob[j] : 0.0;
        }
    }
}

/* __init__.py:50-67 */
static void decode_state(const oracle_env *e, u128 s, int32_t *ids) {
    for (int i = 0; i < e->n; ++i) { ids[i] = (int32_t)(s % (u128)e->L); s /= (u128)e->L; }
}
/* __init__.py:70-79 */
static u128 encode_state(const oracle_env *e, const int32_t *ids) {
    u128 sum = 0, mul = 1;
    for (int i = 0; i < e->n; ++i) { sum += (u128)ids[i] * mul; mul *= (u128)e->L; }
    return sum;
}
static void decode_action(const oracle_env *e, int64_t a, int *acts) {
    for (int i = 0; i < e->n; ++i) { acts[i] = (int)(a % 5); a /= 5; }
}

void oracle_decode_states(const oracle_env *e, int64_t B, const uint64_t *lo, const uint64_t *hi, int32_t *ids) {
    for (int64_t b = 0; b < B; ++b) decode_state(e, ((u128)hi[b] << 64) | lo[b], ids + b * e->n);
}
void oracle_encode_states(const oracle_env *e, int64_t B, const int32_t *ids, uint64_t *lo, uint64_t *hi) {
    for (int64_t b = 0; b < B; ++b) {
        u128 s = encode_state(e, ids + b * e->n);
        lo[b] = (uint64_t)s; hi[b] = (uint64_t)(s >> 64);
    }
}

/* mapf_env.py:210-223 */
static int is_terminal(const oracle_env *e, const int32_t *ids) {
    int all_goal = 1;
    for (int i = 0; i < e->n; ++i) {
        for (int j = i + 1; j < e->n; ++j)
            if (ids[i] == ids[j]) return 1;
        if (ids[i] != e->goal[i]) all_goal = 0;
    }
    return all_goal;
}

/* mapf_env.py:436-446 */
static double living(const oracle_env *e, const int32_t *prev, const int *acts) {
    if (!e->soc) return e->r_living;
    int parked = 0;
    for (int i = 0; i < e->n; ++i)
        if (prev[i] == e->goal[i] && acts[i] == STAY) parked++;
    return (double)(e->n - parked) * e->r_living;
}

/* mapf_env.py:378-389 */
static int clash(const oracle_env *e, const int32_t *prev, const int32_t *nxt) {
    for (int i = 0; i < e->n; ++i)
        for (int j = i + 1; j < e->n; ++j) {
            if (prev[i] == nxt[j] && prev[j] == nxt[i]) return 1;
            if (nxt[i] == nxt[j]) return 1;
        }
    return 0;
}

/* mapf_env.py:225-235 */
static void judge(const oracle_env *e, const int32_t *prev, const int *acts, const int32_t *nxt, double live,
                  double *reward, uint8_t *done, uint8_t *coll) {
    (void)acts;
    if (clash(e, prev, nxt)) { *reward = e->r_clash + live; *done = 1; *coll = 1; return; }
    int all_goal = 1;
    for (int i = 0; i < e->n; ++i)
        if (nxt[i] != e->goal[i]) { all_goal = 0; break; }
    if (all_goal) { *reward = e->r_goal + live; *done = 1; *coll = 0; return; }
    *reward = live; *done = 0; *coll = 0;
}

/* Length of P[s][a] without materialising it (product of the per-agent merged-outcome counts). */
static int64_t row_length(const oracle_env *e, const int32_t *prev, const int *acts) {
    if (is_terminal(e, prev)) return 1;
    int64_t len = 1;
    for (int i = 0; i < e->n; ++i) len *= e->moves[prev[i] * 5 + acts[i]].k;
    return len;
}

typedef struct { uint64_t v[8]; } checks; /* count, n_coll, n_done, sum_lo, sum_hi, sum_prob, sum_rew, ordered */

/* mapf_env.py:448-479 -- one row.  When `next_lo` is NULL only the checksums are accumulated. */
static int64_t emit_row(const oracle_env *e, u128 s, int64_t a, int64_t base, uint64_t *next_lo, uint64_t *next_hi,
                        double *prob, double *reward, uint8_t *done, uint8_t *coll, checks *cs) {
    int32_t prev[MAX_AGENTS] = {0}, nxt[MAX_AGENTS];
    int acts[MAX_AGENTS], digit[MAX_AGENTS];
    const agent_moves *mv[MAX_AGENTS];
    decode_state(e, s, prev);
    int64_t at = base;
    if (is_terminal(e, prev)) { /* :455-456 */
        double p = 1.0, r = 0.0;
        if (next_lo) {
            next_lo[at] = (uint64_t)s; next_hi[at] = (uint64_t)(s >> 64);
            prob[at] = p; reward[at] = r; done[at] = 1; coll[at] = 0;
        }
        if (cs) {
            uint64_t pb, rb; memcpy(&pb, &p, 8); memcpy(&rb, &r, 8);
            cs->v[0] += 1; cs->v[2] += 1; cs->v[3] += (uint64_t)s; cs->v[4] += (uint64_t)(s >> 64);
            cs->v[5] += pb; cs->v[6] += rb;
            cs->v[7] += (uint64_t)(at + 1) * ((uint64_t)s + 1 + 4);
        }
        return 1;
    }
    decode_action(e, a, acts);
    for (int i = 0; i < e->n; ++i) { mv[i] = &e->moves[prev[i] * 5 + acts[i]]; digit[i] = 0; }
    double live = living(e, prev, acts);
    for (;;) { /* itertools.product: the LAST agent's digit moves fastest (:467) */
        double p = mv[0]->prob[digit[0]];
        nxt[0] = mv[0]->dest[digit[0]];
        for (int i = 1; i < e->n; ++i) { p = p * mv[i]->prob[digit[i]]; nxt[i] = mv[i]->dest[digit[i]]; } /* :468 */
        double r; uint8_t d, c;
        judge(e, prev, acts, nxt, live, &r, &d, &c);
        u128 ns = encode_state(e, nxt);
        if (next_lo) {
            next_lo[at] = (uint64_t)ns; next_hi[at] = (uint64_t)(ns >> 64);
            prob[at] = p; reward[at] = r; done[at] = d; coll[at] = c;
        }
        if (cs) {
            uint64_t pb, rb; memcpy(&pb, &p, 8); memcpy(&rb, &r, 8);
            cs->v[0] += 1; cs->v[1] += c; cs->v[2] += d; cs->v[3] += (uint64_t)ns; cs->v[4] += (uint64_t)(ns >> 64);
            cs->v[5] += pb; cs->v[6] += rb;
            cs->v[7] += (uint64_t)(at + 1) * ((uint64_t)ns + 1 + 2 * (uint64_t)c + 4 * (uint64_t)d);
        }
        at++;
        int i = e->n - 1;
        while (i >= 0 && ++digit[i] == mv[i]->k) { digit[i] = 0; --i; }
        if (i < 0) break;
    }
    return at - base;
}

/* Row lengths for B (state, action) pairs; returns their sum. */
int64_t oracle_count_rows(const oracle_env *e, int64_t B, const uint64_t *s_lo, const uint64_t *s_hi,
                          const int64_t *action, int64_t *row_len) {
    int64_t total = 0;
    for (int64_t b = 0; b < B; ++b) {
        int32_t prev[MAX_AGENTS]; int acts[MAX_AGENTS];
        decode_state(e, ((u128)s_hi[b] << 64) | s_lo[b], prev);
        decode_action(e, action[b], acts);
        row_len[b] = row_length(e, prev, acts);
        total += row_len[b];
    }
    return total;
}

/* CSR expansion of B rows; row_ptr[B+1] must already hold the exclusive scan of the row lengths. */
void oracle_expand(const oracle_env *e, int64_t B, const uint64_t *s_lo, const uint64_t *s_hi, const int64_t *action,
                   const int64_t *row_ptr, uint64_t *next_lo, uint64_t *next_hi, double *prob, double *reward,
                   uint8_t *done, uint8_t *coll) {
    for (int64_t b = 0; b < B; ++b)
        emit_row(e, ((u128)s_hi[b] << 64) | s_lo[b], action[b], row_ptr[b], next_lo, next_hi, prob, reward, done,
                 coll, NULL);
}

/* Checksums (mod 2^64) over the table slab [s_begin, s_begin + n_states) x [0, nA), rows in (s, a) order. */
void oracle_table_checksums(const oracle_env *e, uint64_t s_lo, uint64_t s_hi, int64_t n_states, uint64_t *out8) {
    checks cs; memset(&cs, 0, sizeof(cs));
    int64_t nA = 1;
    for (int i = 0; i < e->n; ++i) nA *= 5;
    u128 s = ((u128)s_hi << 64) | s_lo;
    int64_t at = 0;
    for (int64_t i = 0; i < n_states; ++i, ++s)
        for (int64_t a = 0; a < nA; ++a) at += emit_row(e, s, a, at, NULL, NULL, NULL, NULL, NULL, NULL, &cs);
    memcpy(out8, cs.v, sizeof(cs.v));
}

/* mapf_env.py:237-266 -- one sampled step per env; uniforms[b*n + i] is the draw categorical_sample would make
 * for agent i (mapf_env.py:255; gym 0.13.0: (cumsum(p) > u).argmax()).  terminal[b]=1 marks the no-op branch
 * (:238-240), which returns reward 0, prob 0 and consumes no draw. */
void oracle_step(const oracle_env *e, int64_t B, const uint64_t *s_lo, const uint64_t *s_hi, const int64_t *action,
                 const double *uniforms, uint64_t *next_lo, uint64_t *next_hi, double *reward, double *prob,
                 uint8_t *done, uint8_t *coll, uint8_t *terminal) {
    for (int64_t b = 0; b < B; ++b) {
        int32_t prev[MAX_AGENTS], nxt[MAX_AGENTS]; int acts[MAX_AGENTS];
        u128 s = ((u128)s_hi[b] << 64) | s_lo[b];
        decode_state(e, s, prev);
        if (is_terminal(e, prev)) {
            next_lo[b] = s_lo[b]; next_hi[b] = s_hi[b]; reward[b] = 0.0; prob[b] = 0.0; done[b] = 1; coll[b] = 0;
            terminal[b] = 1;
            continue;
        }
        decode_action(e, action[b], acts);
        double total = 1.0;
        for (int i = 0; i < e->n; ++i) {
            const agent_moves *m = &e->moves[prev[i] * 5 + acts[i]];
            double u = uniforms[b * e->n + i], acc = 0.0;
            int pick = 0, found = 0;
            for (int j = 0; j < m->k; ++j) {
                acc = j == 0 ? m->prob[0] : acc + m->prob[j];
                if (!found && acc > u) { pick = j; found = 1; }
            }
            nxt[i] = m->dest[pick];
            total = total * m->prob[pick]; /* :257, starts from 1 */
        }
        double live = living(e, prev, acts);
        judge(e, prev, acts, nxt, live, &reward[b], &done[b], &coll[b]);
        u128 ns = encode_state(e, nxt);
        next_lo[b] = (uint64_t)ns; next_hi[b] = (uint64_t)(ns >> 64); prob[b] = total; terminal[b] = 0;
    }
}

/* ---- rows built after the hot path ------------------------------------------------------------ */

/* The consumer loop of a planner over the table (the classic gym value-iteration backup that gym-mapf's downstream
 * planners run on env.P, mapf_env.py:448-479):
 *     q = sum([p * (r + gamma * V[s2]) for ((p, collision), s2, r, done) in env.P[s][a]])
 * Python's sum() adds left to right starting from the int 0; every operation is one IEEE binary64 operation. */
static double backup_row(const oracle_env *e, u128 s, int64_t a, const double *V, double gamma, uint64_t *t_lo,
                         uint64_t *t_hi, double *t_p, double *t_r, uint8_t *t_d, uint8_t *t_c) {
    int64_t len = emit_row(e, s, a, 0, t_lo, t_hi, t_p, t_r, t_d, t_c, NULL);
    double q = 0.0;
    for (int64_t j = 0; j < len; ++j) {
        double gv = gamma * V[t_lo[j]];
        double inner = t_r[j] + gv;
        double term = t_p[j] * inner;
        q = q + term;
    }
    return q;
}

static int64_t max_row(const oracle_env *e) {
    int64_t m = 1;
    for (int i = 0; i < e->n; ++i) m *= 3;
    return m;
}

void oracle_backup(const oracle_env *e, int64_t B, const uint64_t *s_lo, const uint64_t *s_hi, const int64_t *action,
                   const double *V, double gamma, double *Q) {
    int64_t m = max_row(e);
    uint64_t *t_lo = malloc(m * 8), *t_hi = malloc(m * 8);
    double *t_p = malloc(m * 8), *t_r = malloc(m * 8);
    uint8_t *t_d = malloc(m), *t_c = malloc(m);
    for (int64_t b = 0; b < B; ++b)
        Q[b] = backup_row(e, ((u128)s_hi[b] << 64) | s_lo[b], action[b], V, gamma, t_lo, t_hi, t_p, t_r, t_d, t_c);
    free(t_lo); free(t_hi); free(t_p); free(t_r); free(t_d); free(t_c);
}

/* mapf_env.py:414-425 -- the cells `_single_location_predecessors` returns for one cell: the results of moving
 * DOWN, UP, LEFT, RIGHT, STAY from it (in that order), with duplicates kept out (the caller builds a set). */
static int pred_cells(const oracle_env *e, int32_t cell, int32_t *out) {
    static const int order[5] = {DOWN, UP, LEFT, RIGHT, STAY};
    int r = e->rc_of[cell] / e->W, c = e->rc_of[cell] % e->W, k = 0;
    for (int j = 0; j < 5; ++j) {
        int32_t id = e->cell_of[shift_cell(e, r, c, order[j])];
        int seen = 0;
        for (int q = 0; q < k; ++q) seen |= out[q] == id;
        if (!seen) out[k++] = id;
    }
    return k;
}

/* mapf_env.py:373-376, 426-434 -- |predecessors(s)| for B states */
int64_t oracle_count_predecessors(const oracle_env *e, int64_t B, const uint64_t *s_lo, const uint64_t *s_hi,
                                  int64_t *row_len) {
    int64_t total = 0;
    for (int64_t b = 0; b < B; ++b) {
        int32_t ids[MAX_AGENTS], tmp[5];
        decode_state(e, ((u128)s_hi[b] << 64) | s_lo[b], ids);
        int64_t len = 1;
        for (int i = 0; i < e->n; ++i) len *= pred_cells(e, ids[i], tmp);
        row_len[b] = len;
        total += len;
    }
    return total;
}

static int cmp_u128(const void *a, const void *b) {
    u128 x = *(const u128 *)a, y = *(const u128 *)b;
    return x < y ? -1 : x > y;
}

/* The predecessor sets as CSR, each row sorted ascending (the reference returns an unordered Python set). */
void oracle_predecessors(const oracle_env *e, int64_t B, const uint64_t *s_lo, const uint64_t *s_hi,
                         const int64_t *row_ptr, uint64_t *p_lo, uint64_t *p_hi) {
    for (int64_t b = 0; b < B; ++b) {
        int32_t ids[MAX_AGENTS], opt[MAX_AGENTS][5], pick[MAX_AGENTS];
        int k[MAX_AGENTS], digit[MAX_AGENTS];
        decode_state(e, ((u128)s_hi[b] << 64) | s_lo[b], ids);
        for (int i = 0; i < e->n; ++i) { k[i] = pred_cells(e, ids[i], opt[i]); digit[i] = 0; }
        int64_t len = row_ptr[b + 1] - row_ptr[b], at = 0;
        u128 *buf = malloc(len * sizeof(u128));
        for (;;) {
            for (int i = 0; i < e->n; ++i) pick[i] = opt[i][digit[i]];
            buf[at++] = encode_state(e, pick);
            int i = e->n - 1;
            while (i >= 0 && ++digit[i] == k[i]) { digit[i] = 0; --i; }
            if (i < 0) break;
        }
        qsort(buf, len, sizeof(u128), cmp_u128);
        for (int64_t j = 0; j < len; ++j) { p_lo[row_ptr[b] + j] = (uint64_t)buf[j]; p_hi[row_ptr[b] + j] = (uint64_t)(buf[j] >> 64); }
        free(buf);
    }
}

/* utils.py:138-157 -- a joint state as the sub-env of `agents` (get_local_view) numbers it: the chosen agents'
 * cells, in the sub-env's agent order, re-encoded with the same radix L. */
void oracle_project_states(const oracle_env *e, int64_t B, const uint64_t *s_lo, const uint64_t *s_hi, int n_sub,
                           const int32_t *agents, uint64_t *o_lo, uint64_t *o_hi) {
    for (int64_t b = 0; b < B; ++b) {
        int32_t ids[MAX_AGENTS];
        decode_state(e, ((u128)s_hi[b] << 64) | s_lo[b], ids);
        u128 x = 0, w = 1;
        for (int j = 0; j < n_sub; ++j) { x += (u128)ids[agents[j]] * w; w *= (u128)e->L; }
        o_lo[b] = (uint64_t)x; o_hi[b] = (uint64_t)(x >> 64);
    }
}

/* ---- multi-threaded drivers, used only as the timed CPU baseline (bench.py) -------------------------------- */
typedef struct {
    const oracle_env *e; int64_t b0, b1;
    const uint64_t *s_lo, *s_hi; const int64_t *action; const double *uniforms; const int64_t *row_ptr;
    uint64_t *next_lo, *next_hi; double *reward, *prob; uint8_t *done, *coll, *terminal;
} job;

static void *step_worker(void *arg) {
    job *j = (job *)arg;
    int64_t o = j->b0;
    oracle_step(j->e, j->b1 - j->b0, j->s_lo + o, j->s_hi + o, j->action + o, j->uniforms + o * j->e->n,
                j->next_lo + o, j->next_hi + o, j->reward + o, j->prob + o, j->done + o, j->coll + o, j->terminal + o);
    return NULL;
}

void oracle_step_mt(const oracle_env *e, int64_t B, const uint64_t *s_lo, const uint64_t *s_hi, const int64_t *action,
                    const double *uniforms, uint64_t *next_lo, uint64_t *next_hi, double *reward, double *prob,
                    uint8_t *done, uint8_t *coll, uint8_t *terminal, int threads) {
    if (threads < 1) threads = 1;
    pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * threads);
    job *jobs = (job *)malloc(sizeof(job) * threads);
    for (int t = 0; t < threads; ++t) {
        job j = {e, B * t / threads, B * (t + 1) / threads, s_lo, s_hi, action, uniforms, NULL,
                 next_lo, next_hi, reward, prob, done, coll, terminal};
        jobs[t] = j;
        pthread_create(&tid[t], NULL, step_worker, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(tid[t], NULL);
    free(tid); free(jobs);
}

static void *expand_worker(void *arg) {
    job *j = (job *)arg;
    for (int64_t b = j->b0; b < j->b1; ++b)
        emit_row(j->e, ((u128)j->s_hi[b] << 64) | j->s_lo[b], j->action[b], j->row_ptr[b], j->next_lo, j->next_hi,
                 j->prob, j->reward, j->done, j->coll, NULL);
    return NULL;
}

void oracle_expand_mt(const oracle_env *e, int64_t B, const uint64_t *s_lo, const uint64_t *s_hi,
                      const int64_t *action, const int64_t *row_ptr, uint64_t *next_lo, uint64_t *next_hi,
                      double *prob, double *reward, uint8_t *done, uint8_t *coll, int threads) {
    if (threads < 1) threads = 1;
    pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * threads);
    job *jobs = (job *)malloc(sizeof(job) * threads);
    for (int t = 0; t < threads; ++t) {
        job j = {e, B * t / threads, B * (t + 1) / threads, s_lo, s_hi, action, NULL, row_ptr,
                 next_lo, next_hi, reward, prob, done, coll, NULL};
        jobs[t] = j;
        pthread_create(&tid[t], NULL, expand_worker, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(tid[t], NULL);
    free(tid); free(jobs);
}

typedef struct {
    const oracle_env *e; int64_t b0, b1;
    const uint64_t *s_lo, *s_hi; const int64_t *action; const double *V; double gamma; double *Q;
} backup_job;

static void *backup_worker(void *arg) {
    backup_job *j = arg;
    oracle_backup(j->e, j->b1 - j->b0, j->s_lo + j->b0, j->s_hi + j->b0, j->action + j->b0, j->V, j->gamma, j->Q + j->b0);
    return NULL;
}

void oracle_backup_mt(const oracle_env *e, int64_t B, const uint64_t *s_lo, const uint64_t *s_hi, const int64_t *action,
                      const double *V, double gamma, double *Q, int threads) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t th[256];
    backup_job jobs[256];
    for (int t = 0; t < threads; ++t) {
        jobs[t] = (backup_job){e, B * t / threads, B * (t + 1) / threads, s_lo, s_hi, action, V, gamma, Q};
        pthread_create(&th[t], NULL, backup_worker, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
}
