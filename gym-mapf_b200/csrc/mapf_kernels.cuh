// mapf_kernels.cuh -- the sm_100a kernels of the joint-transition engine.
//
// Mapping ("T" family): one thread per emitted unit -- one env-step in step/rollout mode, one P[s][a] record in
// expand mode.  All per-agent loops are unrolled over the template parameter N, so an agent's cell, move-table
// entry and outcome digit live in registers.  The per-(cell, action) move table is staged in shared memory by one
// bulk asynchronous copy (TMA engine) per CTA.
//
// Reference citations are file:line relative to /root/reference/gym_mapf/envs/.
#pragma once
#include "mapf_device.cuh"

#ifndef MAPF_MAX_THREADS
#define MAPF_MAX_THREADS 512  // largest CTA the hot kernels are launched with
#endif
// Register budget of the step kernel: up to 6 agents everything fits 64 registers (2 CTAs of 512 or 4 of 256
// threads per SM, measured best on the 4-agent workload); more agents get the full 128.
#ifndef MAPF_MIN_BLOCKS
#define MAPF_MIN_BLOCKS(N) ((N) <= 6 ? 2 : 1)
#endif

// =====================================================================================================
// Move-table construction from the obstacle bitmap (ctx creation; not a hot kernel)
// =====================================================================================================
// colbits: per column c, wpc 32-bit words; bit r of the column is 1 when cell (r, c) is FREE.
// colbase: number of free cells in columns < c, i.e. the id of the first free cell of column c
//          (column-major numbering: grid.py:37-40, mapf_env.py:142-143).
// The whole bitmap is staged in shared memory; one thread per grid position.
struct BitmapView {
    const u32 *bits;
    const u32 *base;
    int H, W, wpc;
};

__device__ __forceinline__ bool bm_free(const BitmapView &bm, int r, int c) {
    return (bm.bits[c * bm.wpc + (r >> 5)] >> (r & 31)) & 1u;
}

__device__ __forceinline__ int bm_rank(const BitmapView &bm, int r, int c) {
    int id = (int)bm.base[c];
    const u32 *col = bm.bits + c * bm.wpc;
    for (int w = 0; w < (r >> 5); ++w) id += __popc(col[w]);
    id += __popc(col[r >> 5] & ((1u << (r & 31)) - 1u));
    return id;
}

// execute_up/down/right/left/stay + stay_if_hit_obstacle (mapf_env.py:43-84): clamp to the grid, obstacle = stay
__device__ __forceinline__ void bm_move(const BitmapView &bm, int r, int c, int d, int &nr, int &nc) {
    int tr = r, tc = c;
    if (d == 1) tr = max(0, r - 1);
    else if (d == 3) tr = min(bm.H - 1, r + 1);
    else if (d == 2) tc = min(bm.W - 1, c + 1);
    else if (d == 4) tc = max(0, c - 1);
    bool ok = (d == 0) || bm_free(bm, tr, tc);
    nr = ok ? tr : r;
    nc = ok ? tc : c;
}

// The merge patterns that can occur, as 9-bit triples (mask of slot 0 | mask of slot 1 << 3 | mask of slot 2 << 6)
// of candidate bits {intended = 1, right = 2, left = 4}; the index of a triple is its pattern id.
struct PatternList {
    u32 triple[MAPF_MAX_PATTERNS];
    int count;
};

// single_agent_movements (mapf_env.py:163-184) for one (cell, action): candidates [intended, right, left],
// zero-probability candidates dropped (cand_mask), equal destinations merged into the first occurrence.
__device__ __forceinline__ u64 bm_entry(const BitmapView &bm, int r, int c, int a, int cand_mask,
                                        const PatternList &pats) {
    // POSSIBILITIES (__init__.py:19-25): right/left slip of STAY,UP,RIGHT,DOWN,LEFT
    const int slip_r[5] = {0, 2, 3, 4, 1};
    const int slip_l[5] = {0, 4, 1, 2, 3};
    int dir[3] = {a, slip_r[a], slip_l[a]};
    u32 dest[3] = {0, 0, 0}, mask[3] = {0, 0, 0};
    int k = 0;
    for (int j = 0; j < 3; ++j) {
        if (!((cand_mask >> j) & 1)) continue;
        int nr, nc;
        bm_move(bm, r, c, dir[j], nr, nc);
        u32 id = (u32)bm_rank(bm, nr, nc);
        int at = -1;
        for (int q = 0; q < k; ++q)
            if (dest[q] == id && at < 0) at = q;
        if (at >= 0) mask[at] |= 1u << j;
        else { dest[k] = id; mask[k] = 1u << j; ++k; }
    }
    for (int q = k; q < 3; ++q) dest[q] = dest[0];
    const u32 triple = mask[0] | (mask[1] << 3) | (mask[2] << 6);
    u32 pid = 0;
    for (int q = 0; q < pats.count; ++q)
        if (pats.triple[q] == triple) pid = (u32)q;
    return (u64)dest[0] | ((u64)dest[1] << 16) | ((u64)dest[2] << 32) | ((u64)(pid * 32u) << 48) | ((u64)k << 56);
}

static __global__ void k_build_moves(const u32 *__restrict__ colbits, const u32 *__restrict__ colbase, int H, int W, int wpc,
                              int cand_mask, PatternList pats, u64 *__restrict__ lut, u32 *__restrict__ cell_rc) {
    extern __shared__ u32 bm_smem[];
    u32 *s_bits = bm_smem;
    u32 *s_base = bm_smem + W * wpc;
    for (int i = threadIdx.x; i < W * wpc; i += blockDim.x) s_bits[i] = colbits[i];
    for (int i = threadIdx.x; i < W; i += blockDim.x) s_base[i] = colbase[i];
    __syncthreads();
    BitmapView bm = {s_bits, s_base, H, W, wpc};
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < H * W; p += gridDim.x * blockDim.x) {
        int c = p / H, r = p - c * H;
        if (!bm_free(bm, r, c)) continue;
        int id = bm_rank(bm, r, c);
        cell_rc[id] = ((u32)r << 16) | (u32)c;
        for (int a = 0; a < 5; ++a) lut[id * 5 + a] = bm_entry(bm, r, c, a, cand_mask, pats);
    }
}

// =====================================================================================================
// Bulk state <-> cells (state_to_locations / locations_to_state, mapf_env.py:358-371)
// =====================================================================================================
template <int WORDS>
__device__ __forceinline__ void load_state(const u64 *states, i64 b, u64 &lo, u64 &hi) {
    if (WORDS == 1) { lo = states[b]; hi = 0; }
    else { ulonglong2 v = reinterpret_cast<const ulonglong2 *>(states)[b]; lo = v.x; hi = v.y; }
}
template <int WORDS>
__device__ __forceinline__ void store_state(u64 *states, i64 b, u64 lo, u64 hi) {
    if (WORDS == 1) states[b] = lo;
    else reinterpret_cast<ulonglong2 *>(states)[b] = make_ulonglong2(lo, hi);
}

template <int N, int WORDS>
__global__ void __launch_bounds__(256) k_decode(DevSpec sp, const u64 *__restrict__ states, i64 B, int *__restrict__ cells) {
    for (i64 b = (i64)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (i64)gridDim.x * blockDim.x) {
        u64 lo, hi;
        load_state<WORDS>(states, b, lo, hi);
        int cell[N];
        decode_state<N, WORDS>(sp, lo, hi, cell);
#pragma unroll
        for (int i = 0; i < N; ++i) cells[b * N + i] = cell[i];
    }
}

template <int N, int WORDS>
__global__ void __launch_bounds__(256) k_encode(DevSpec sp, const int *__restrict__ cells, i64 B, u64 *__restrict__ states) {
    for (i64 b = (i64)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (i64)gridDim.x * blockDim.x) {
        int cell[N];
#pragma unroll
        for (int i = 0; i < N; ++i) cell[i] = min(max(cells[b * N + i], 0), sp.L - 1);
        u64 lo, hi;
        encode_state<N, WORDS>(sp, cell, lo, hi);
        store_state<WORDS>(states, b, lo, hi);
    }
}

// =====================================================================================================
// Row sources: explicit (state, action) pairs, or a slab of the full table generated from the row index
// =====================================================================================================
template <int WORDS, bool RANGE>
__device__ __forceinline__ void row_input(const DevSpec &sp, const u64 *states, const int *actions, u64 sb_lo, u64 sb_hi,
                                          i64 b, u64 &lo, u64 &hi, u32 &a) {
    if (RANGE) {
        u64 off = (u64)b / sp.nA;  // nA is a kernel-uniform constant; one 64-bit division per ROW
        a = (u32)((u64)b - off * sp.nA);
        lo = sb_lo + off;
        hi = sb_hi + (lo < sb_lo ? 1ull : 0ull);
    } else {
        load_state<WORDS>(states, b, lo, hi);
        a = (u32)actions[b];
    }
}

// len(P[s][a]) (mapf_env.py:448-479): 1 for a terminal state, else the product of merged-outcome counts
template <int N, int WORDS, bool RANGE>
__global__ void __launch_bounds__(256) k_count(DevSpec sp, const u64 *__restrict__ states, const int *__restrict__ actions,
                                               u64 sb_lo, u64 sb_hi, i64 B, i64 *__restrict__ row_len) {
    for (i64 b = (i64)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (i64)gridDim.x * blockDim.x) {
        u64 lo, hi;
        u32 a;
        row_input<WORDS, RANGE>(sp, states, actions, sb_lo, sb_hi, b, lo, hi, a);
        int cell[N], act[N];
        decode_state<N, WORDS>(sp, lo, hi, cell);
        decode_action<N>(a, act);
        i64 len = 1;
        if (!is_terminal<N>(sp, cell, lo, hi)) {
#pragma unroll
            for (int i = 0; i < N; ++i) len *= (i64)ENT_K(__ldg(sp.lut + cell[i] * 5 + act[i]));
        }
        row_len[b] = len;
    }
}

// =====================================================================================================
// Exclusive scan of row lengths (three small kernels; chunk = 2048 elements per block)
// =====================================================================================================
#define SCAN_CHUNK 2048
__device__ __forceinline__ i64 block_scan_256(i64 v, i64 *warp_sums, i64 &block_total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    i64 x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        i64 y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
        i64 s = lane < 8 ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            i64 y = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += y;
        }
        if (lane < 8) warp_sums[lane] = s;
    }
    __syncthreads();
    block_total = warp_sums[7];
    i64 incl = x + (wid > 0 ? warp_sums[wid - 1] : 0);
    __syncthreads();
    return incl - v;  // exclusive
}

static __global__ void __launch_bounds__(256) k_scan_partials(const i64 *__restrict__ in, i64 B, i64 *__restrict__ partial) {
    __shared__ i64 ws[8];
    i64 base = (i64)blockIdx.x * SCAN_CHUNK;
    i64 s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_CHUNK / 256; ++j) {
        i64 i = base + j * 256 + threadIdx.x;
        s += i < B ? in[i] : 0;
    }
    i64 total;
    block_scan_256(s, ws, total);
    if (threadIdx.x == 0) partial[blockIdx.x] = total;
}

// single block: exclusive scan of the per-chunk totals in place; partial[n_chunks] = grand total
static __global__ void __launch_bounds__(256) k_scan_spine(i64 *partial, i64 n_chunks) {
    __shared__ i64 ws[8];
    i64 carry = 0;
    for (i64 base = 0; base < n_chunks; base += 256) {
        i64 i = base + threadIdx.x;
        i64 v = i < n_chunks ? partial[i] : 0;
        i64 total;
        i64 ex = block_scan_256(v, ws, total);
        if (i < n_chunks) partial[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) partial[n_chunks] = carry;
}

static __global__ void __launch_bounds__(256) k_scan_final(const i64 *__restrict__ in, i64 B, const i64 *__restrict__ partial,
                                                    i64 *__restrict__ row_ptr) {
    __shared__ i64 ws[8];
    i64 base = (i64)blockIdx.x * SCAN_CHUNK;
    i64 carry = partial[blockIdx.x];
    // thread t owns 8 consecutive elements of the chunk
    i64 v[SCAN_CHUNK / 256];
    i64 s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_CHUNK / 256; ++j) {
        i64 i = base + (i64)threadIdx.x * (SCAN_CHUNK / 256) + j;
        v[j] = i < B ? in[i] : 0;
        s += v[j];
    }
    i64 total;
    i64 ex = block_scan_256(s, ws, total) + carry;
#pragma unroll
    for (int j = 0; j < SCAN_CHUNK / 256; ++j) {
        i64 i = base + (i64)threadIdx.x * (SCAN_CHUNK / 256) + j;
        if (i < B) row_ptr[i] = ex;
        ex += v[j];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) row_ptr[B] = partial[gridDim.x];
}

// =====================================================================================================
// Expand: P[s][a] rows as CSR records (mapf_env.py:448-479)
// =====================================================================================================
// A warp takes 32 consecutive rows at a time.
//   phase A (lane = row):     decode (s, a), terminal test, fetch the N move-table entries, row length;
//                             the row descriptor goes to the warp's shared-memory slab
//   phase B (lane = record):  the rows' records are consecutive in the output (row_ptr is their exclusive scan),
//                             so record j of the 32-row batch goes to base + j: every store of the warp is one
//                             contiguous, fully coalesced segment
template <int N>
struct ExpandSlab {
    u64 ent[N][32];   // move-table entry of agent i for row r
    u64 st[2][32];    // the row's own state (terminal rows re-emit it)
    u32 pref[33];     // exclusive scan of the 32 row lengths
    u32 pad;
    u16 prev[N][32];  // current cell of agent i
    u8 parked[32];    // SoC: agents parked on their goal choosing STAY
    u8 term[32];
};

template <int N, int WORDS, bool LUTS, bool RANGE>
__global__ void __launch_bounds__(MAPF_MAX_THREADS)
k_expand(DevSpec sp, const u64 *__restrict__ states, const int *__restrict__ actions, u64 sb_lo, u64 sb_hi, i64 B,
         const i64 *__restrict__ row_ptr, u64 *__restrict__ next_state, double *__restrict__ prob,
         double *__restrict__ reward, u8 *__restrict__ flags) {
    extern __shared__ __align__(16) unsigned char smem[];
    SmemTables tb = tables_begin<LUTS>(sp, smem);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    ExpandSlab<N> &sl = reinterpret_cast<ExpandSlab<N> *>(smem + MAPF_SMEM_LUT + (LUTS ? sp.lut_bytes : 0))[wid];
    const i64 n_batches = (B + 31) >> 5;
    const i64 warps_total = (i64)gridDim.x * (blockDim.x >> 5);
    bool ready = false;
    for (i64 batch = (i64)blockIdx.x * (blockDim.x >> 5) + wid; batch < n_batches; batch += warps_total) {
        // ---------------- phase A
        const i64 b = batch * 32 + lane;
        u32 len = 0;
        int cell[N], act[N];
        u64 lo = 0, hi = 0;
        if (b < B) {
            u32 a;
            row_input<WORDS, RANGE>(sp, states, actions, sb_lo, sb_hi, b, lo, hi, a);
            decode_state<N, WORDS>(sp, lo, hi, cell);
            decode_action<N>(a, act);
        }
        if (!ready) { tables_wait<LUTS>(smem); ready = true; }
        if (b < B) {
            const bool term = is_terminal<N>(sp, cell, lo, hi);
            len = 1;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                u64 e = lut_entry<LUTS>(tb, (u32)cell[i], (u32)act[i] * 8u + (LUTS ? tb.lut : 0u));
                sl.ent[i][lane] = e;
                sl.prev[i][lane] = (u16)cell[i];
                len *= ENT_K(e);
            }
            if (term) len = 1;
            sl.st[0][lane] = lo;
            sl.st[1][lane] = hi;
            sl.parked[lane] = (u8)parked_agents<N>(sp, cell, act);
            sl.term[lane] = term ? 1 : 0;
        }
        u32 incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        sl.pref[lane + 1] = incl;
        if (lane == 0) sl.pref[0] = 0;
        const i64 out0 = row_ptr[batch * 32];
        __syncwarp();
        // ---------------- phase B
        const u32 total = sl.pref[32];
        int r = 0;
        for (u32 j = lane; j < total; j += 32) {
            while (j >= sl.pref[r + 1]) ++r;
            u32 o = j - sl.pref[r];
            const i64 idx = out0 + j;
            if (sl.term[r]) {  // [((1.0, False), s, 0, True)]  (mapf_env.py:455-456)
                store_state<WORDS>(next_state, idx, sl.st[0][r], sl.st[1][r]);
                prob[idx] = 1.0;
                reward[idx] = 0.0;
                flags[idx] = 1;
                continue;
            }
            // outcome digits: itertools.product, the LAST agent's digit moves fastest (mapf_env.py:467)
            int nxt[N], prv[N];
            u32 pj[N];
#pragma unroll
            for (int i = N - 1; i >= 0; --i) {
                const u64 e = sl.ent[i][r];
                const u32 k = ENT_K(e);
                u32 d;
                if (k == 1) { d = 0; }
                else if (k == 2) { d = o & 1u; o >>= 1; }
                else { u32 q = __umulhi(o, 0xAAAAAAABu) >> 1; d = o - 3u * q; o = q; }
                nxt[i] = (int)ent_dest(e, d);
                pj[i] = ENT_POFF(e) + d * 8u;
                prv[i] = (int)sl.prev[i][r];
            }
            // probability: left-to-right product (mapf_env.py:468)
            double p = lds_f64<MAPF_SMEM_PP>(tb.base + pj[0]);
#pragma unroll
            for (int i = 1; i < N; ++i) p = __dmul_rn(p, lds_f64<MAPF_SMEM_PP>(tb.base + pj[i]));
            // reward / done / collision (mapf_env.py:225-235): clash beats goal
            const bool clash = has_clash<N>(prv, nxt);
            bool goal = true;
#pragma unroll
            for (int i = 0; i < N; ++i) goal = goal && (nxt[i] == (int)sp.goal[i]);
            const int kind = clash ? 1 : (goal ? 2 : 0);
            u64 nlo, nhi;
            encode_state<N, WORDS>(sp, nxt, nlo, nhi);
            store_state<WORDS>(next_state, idx, nlo, nhi);
            prob[idx] = p;
            reward[idx] = lds_f64<MAPF_SMEM_REW>(tb.base + (u32)(kind * MAPF_REW_STRIDE + sl.parked[r]) * 8u);
            flags[idx] = (u8)((kind != 0 ? 1 : 0) | (clash ? 2 : 0));
        }
        __syncwarp();
    }
    if (!ready) tables_wait<LUTS>(smem);  // never leave a CTA while its bulk copy is in flight
}

// =====================================================================================================
// Checksums of a record array (mod 2**64), accumulated into out8 with atomics
// =====================================================================================================
static __global__ void __launch_bounds__(256)
k_checksum(int words, i64 n, i64 index_base, const u64 *__restrict__ next_state, const double *__restrict__ prob,
           const double *__restrict__ reward, const u8 *__restrict__ flags, u64 *out8) {
    u64 acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        u64 lo = next_state[i * words], hi = words == 2 ? next_state[i * 2 + 1] : 0;
        u32 f = flags[i];
        u64 d = f & 1u, c = (f >> 1) & 1u;
        acc[0] += 1; acc[1] += c; acc[2] += d; acc[3] += lo; acc[4] += hi;
        acc[5] += (u64)__double_as_longlong(prob[i]);
        acc[6] += (u64)__double_as_longlong(reward[i]);
        acc[7] += (u64)(index_base + i + 1) * (lo + 1 + 2 * c + 4 * d);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) atomicAdd(out8 + k, acc[k]);
    }
}

// =====================================================================================================
// Step / rollout (MapfEnv.step, mapf_env.py:237-266)
// =====================================================================================================
// Everything about one env that does not need the move table: its decoded cells, action digits and draws.
template <int N>
struct EnvIn {
    int cell[N];
    int act[N];
    u32 w[((N + 3) / 4) * 4];  // Philox words, one per agent
    u64 lo, hi;
};

struct EnvOut {
    u64 lo, hi;
    double reward, prob;
    u32 done, coll;
};

template <int N>
__device__ __forceinline__ void env_draws(const PhiloxKeys &K, u64 env, u64 step, EnvIn<N> &in) {
#pragma unroll
    for (int b = 0; b < (N + 3) / 4; ++b) {
        Philox4 x = philox_block(K, env, step, (u32)b);
#pragma unroll
        for (int q = 0; q < 4; ++q) in.w[b * 4 + q] = x.v[q];
    }
}

// One sampled joint transition.  TAPE: agent i's uniform is u[i] (a replayed reference draw) and the choice is
// `(cumsum > u).argmax()` in fp64 (mapf_env.py:255).  Otherwise the draw is the 32-bit Philox word w, u = w * 2**-32,
// and the same comparison is made on integers: cumsum_j > u  <=>  w <= T_j (the table holds ~T_j).  The host guarantees that the last
// threshold of every pattern is 2**32 - 1 (the probabilities of a pattern add up to 1), so the index is simply the
// number of thresholds below w.
template <int N, int WORDS, bool LUTS, bool TAPE>
__device__ __forceinline__ EnvOut env_step(const DevSpec &sp, const SmemTables &tb, const EnvIn<N> &in,
                                           const double *__restrict__ u, u32 opts, int (&nxt)[N]) {
    double total;
    const bool term = is_terminal<N>(sp, in.cell, in.lo, in.hi);
    const u32 act_base = LUTS ? tb.lut : 0u;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const u64 e = lut_entry<LUTS>(tb, (u32)in.cell[i], (u32)in.act[i] * 8u + act_base);
        const u32 row = tb.base + ENT_POFF(e);
        u32 pick;
        if (TAPE) {
            const double ui = u[i];
            const double c0 = lds_f64<MAPF_SMEM_CUM>(row), c1 = lds_f64<MAPF_SMEM_CUM + 8>(row),
                         c2 = lds_f64<MAPF_SMEM_CUM + 16>(row);
            pick = c0 > ui ? 0u : (c1 > ui ? 1u : (c2 > ui ? 2u : 0u));
        } else {
            const uint2 t = lds_u32x2<MAPF_SMEM_THR>(row);
            const u32 w = in.w[i];
            pick = count_below(w, t.x, t.y);
        }
        nxt[i] = (int)ent_dest(e, pick);
        const double pi = lds_f64<MAPF_SMEM_PP>(row + pick * 8u);
        total = i == 0 ? pi : __dmul_rn(total, pi);  // 1 * p0 * p1 * ... (mapf_env.py:250,257)
    }
    const bool clash = has_clash<N>(in.cell, nxt);
    EnvOut out;
    encode_state<N, WORDS>(sp, nxt, out.lo, out.hi);
    const bool goal = out.lo == sp.sgoal[0] && out.hi == sp.sgoal[1];  // every agent on its goal
    const int kind = clash ? 1 : (goal ? 2 : 0);
    out.reward = lds_f64<MAPF_SMEM_REW>(tb.base + (u32)(kind * MAPF_REW_STRIDE + parked_agents<N>(sp, in.cell, in.act)) * 8u);
    out.prob = total;
    out.done = kind != 0 ? 1u : 0u;
    out.coll = clash ? 1u : 0u;
    if (term) {  // (s, 0, True, {"prob": 0})  (mapf_env.py:238-240): a no-op that consumes no draw
        out.lo = in.lo; out.hi = in.hi; out.reward = 0.0; out.prob = 0.0; out.done = 1u; out.coll = 0u;
#pragma unroll
        for (int i = 0; i < N; ++i) nxt[i] = in.cell[i];
    }
    if ((opts & 1u) && out.done) {  // MAPF_OPT_AUTO_RESET
        out.lo = sp.s0[0]; out.hi = sp.s0[1];
#pragma unroll
        for (int i = 0; i < N; ++i) nxt[i] = (int)sp.start[i];
    }
    return out;
}

__device__ __forceinline__ u32 random_action(const DevSpec &sp, const PhiloxKeys &K, u64 env, u64 step) {
    Philox4 x = philox_block(K, env, step, 15u);
    return (u32)__umul64hi(((u64)x.v[0] << 32) | x.v[1], sp.nA);
}

// EPT = envs per thread per iteration.  EPT == 2 uses 128-bit loads/stores for the 8-byte fields (and 16-bit
// stores for the two flag bytes); the launcher picks it only when B is even and every pointer is 16-byte aligned.
// B < 2**31 (the launcher splits larger batches), so every index is 32-bit and an address is one wide multiply-add.
// The inputs of the thread's next iteration are loaded before the current one is computed (software prefetch).
template <int WORDS, int EPT>
struct RawIn {
    u64 lo[EPT], hi[EPT];
    u32 a[EPT];
};

template <int WORDS, int EPT>
__device__ __forceinline__ void load_raw(const u64 *states, const int *__restrict__ actions, u32 it, RawIn<WORDS, EPT> &r) {
    if (EPT == 2 && WORDS == 1) {
        const ulonglong2 s2 = reinterpret_cast<const ulonglong2 *>(states)[it];
        const int2 a2 = reinterpret_cast<const int2 *>(actions)[it];
        r.lo[0] = s2.x; r.lo[EPT - 1] = s2.y; r.hi[0] = 0; r.hi[EPT - 1] = 0;
        r.a[0] = (u32)a2.x; r.a[EPT - 1] = (u32)a2.y;
    } else {
#pragma unroll
        for (int q = 0; q < EPT; ++q) {
            load_state<WORDS>(states, it * EPT + q, r.lo[q], r.hi[q]);
            r.a[q] = (u32)actions[it * EPT + q];
        }
    }
}

template <int N, int WORDS, bool LUTS, bool TAPE, int EPT>
__global__ void __launch_bounds__(MAPF_MAX_THREADS, MAPF_MIN_BLOCKS(N))
k_step(DevSpec sp, PhiloxKeys keys, const u64 *states, const int *__restrict__ actions, u32 B,
       const double *__restrict__ uniforms, u64 step, u64 env0, u32 opts, u64 *next_states,
       double *__restrict__ reward, double *__restrict__ prob, u8 *__restrict__ done, u8 *__restrict__ coll) {
    extern __shared__ __align__(16) unsigned char smem[];
    // Programmatic dependent launch: let the next kernel of the stream start its prologue (table staging) while
    // this grid drains, and do our own prologue before waiting for the previous grid's results to be visible.
    asm volatile("griddepcontrol.launch_dependents;");
    SmemTables tb = tables_begin<LUTS>(sp, smem);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    bool ready = false;
    const u32 n_items = B / EPT;
    const u32 stride = gridDim.x * blockDim.x;
    u32 it = blockIdx.x * blockDim.x + threadIdx.x;
    RawIn<WORDS, EPT> raw;
    if (it < n_items) load_raw<WORDS, EPT>(states, actions, it, raw);
    while (it < n_items) {
        EnvIn<N> in[EPT];
        const u32 b = it * EPT;
#pragma unroll
        for (int q = 0; q < EPT; ++q) {
            in[q].lo = raw.lo[q];
            in[q].hi = raw.hi[q];
            decode_action<N>(raw.a[q], in[q].act);
        }
        const u32 it_next = it + stride;
        if (it_next < n_items) load_raw<WORDS, EPT>(states, actions, it_next, raw);  // in flight during the compute below
#pragma unroll
        for (int q = 0; q < EPT; ++q) {
            decode_state<N, WORDS>(sp, in[q].lo, in[q].hi, in[q].cell);
            if (!TAPE) env_draws<N>(keys, env0 + (u64)(b + q), step, in[q]);
        }
        if (!ready) { tables_wait<LUTS>(smem); ready = true; }
        EnvOut o[EPT];
#pragma unroll
        for (int q = 0; q < EPT; ++q) {
            int nxt[N];
            o[q] = env_step<N, WORDS, LUTS, TAPE>(sp, tb, in[q], TAPE ? uniforms + (size_t)(b + q) * N : nullptr, opts,
                                                  nxt);
        }
        if (EPT == 2) {
            if (WORDS == 1) reinterpret_cast<ulonglong2 *>(next_states)[it] = make_ulonglong2(o[0].lo, o[EPT - 1].lo);
            else {
                store_state<WORDS>(next_states, b, o[0].lo, o[0].hi);
                store_state<WORDS>(next_states, b + 1, o[EPT - 1].lo, o[EPT - 1].hi);
            }
            reinterpret_cast<double2 *>(reward)[it] = make_double2(o[0].reward, o[EPT - 1].reward);
            reinterpret_cast<double2 *>(prob)[it] = make_double2(o[0].prob, o[EPT - 1].prob);
            reinterpret_cast<u16 *>(done)[it] = (u16)(o[0].done | (o[EPT - 1].done << 8));
            reinterpret_cast<u16 *>(coll)[it] = (u16)(o[0].coll | (o[EPT - 1].coll << 8));
        } else {
            store_state<WORDS>(next_states, b, o[0].lo, o[0].hi);
            reward[b] = o[0].reward;
            prob[b] = o[0].prob;
            done[b] = (u8)o[0].done;
            coll[b] = (u8)o[0].coll;
        }
        it = it_next;
    }
    if (!ready) tables_wait<LUTS>(smem);
}

// T steps per launch; the env's cells stay in registers between steps, each step's results go to slab t.
template <int N, int WORDS, bool LUTS, bool TAPE>
__global__ void __launch_bounds__(MAPF_MAX_THREADS)
k_rollout(DevSpec sp, PhiloxKeys keys, u64 *states, const int *__restrict__ actions, i64 T, u32 B,
          const double *__restrict__ uniforms, u64 step0, u64 env0, u32 opts, u64 *__restrict__ next_states,
          double *__restrict__ reward, double *__restrict__ prob, u8 *__restrict__ done, u8 *__restrict__ coll) {
    extern __shared__ __align__(16) unsigned char smem[];
    SmemTables tb = tables_begin<LUTS>(sp, smem);
    bool ready = false;
    for (u32 b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
        EnvIn<N> in;
        load_state<WORDS>(states, b, in.lo, in.hi);
        decode_state<N, WORDS>(sp, in.lo, in.hi, in.cell);
        for (i64 t = 0; t < T; ++t) {
            const i64 o = t * (i64)B + b;
            const u64 env = env0 + (u64)b, stp = step0 + (u64)t;
            const u32 a = actions ? (u32)actions[o] : random_action(sp, keys, env, stp);
            decode_action<N>(a, in.act);
            if (!TAPE) env_draws<N>(keys, env, stp, in);
            if (!ready) { tables_wait<LUTS>(smem); ready = true; }
            int nxt[N];
            EnvOut r = env_step<N, WORDS, LUTS, TAPE>(sp, tb, in, TAPE ? uniforms + o * N : nullptr, opts, nxt);
            store_state<WORDS>(next_states, o, r.lo, r.hi);
            reward[o] = r.reward;
            prob[o] = r.prob;
            done[o] = (u8)r.done;
            coll[o] = (u8)r.coll;
            // carry the env forward in registers (cells of the possibly reset next state)
            in.lo = r.lo;
            in.hi = r.hi;
#pragma unroll
            for (int i = 0; i < N; ++i) in.cell[i] = nxt[i];
        }
        store_state<WORDS>(states, b, in.lo, in.hi);
    }
    if (!ready) tables_wait<LUTS>(smem);
}
