// mapf_kernels.cuh -- the sm_100a kernels of the joint-transition engine.
//
// Mapping ("T" family): one thread per emitted unit -- one env-step in step/rollout mode, one P[s][a] record in
// expand mode.  All per-agent loops are unrolled over the template parameter N, so an agent's cell, move-table
// entry and outcome digit live in registers.  The per-(cell, action) move table is staged in shared memory.
//
// Reference citations are file:line relative to /root/reference/gym_mapf/envs/.
#pragma once
#include "mapf_device.cuh"

#define MAPF_MAX_THREADS 512  // largest CTA the hot kernels are launched with

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

// single_agent_movements (mapf_env.py:163-184) for one (cell, action): candidates [intended, right, left],
// zero-probability candidates dropped (cand_mask), equal destinations merged into the first occurrence.
__device__ __forceinline__ u64 bm_entry(const BitmapView &bm, int r, int c, int a, int cand_mask) {
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
    return (u64)dest[0] | ((u64)dest[1] << 16) | ((u64)dest[2] << 32) | ((u64)mask[0] << 48) | ((u64)mask[1] << 51) |
           ((u64)mask[2] << 54) | ((u64)k << 57);
}

__global__ void k_build_moves(const u32 *__restrict__ colbits, const u32 *__restrict__ colbase, int H, int W, int wpc,
                              int cand_mask, u64 *__restrict__ lut, u32 *__restrict__ cell_rc) {
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
        for (int a = 0; a < 5; ++a) lut[id * 5 + a] = bm_entry(bm, r, c, a, cand_mask);
    }
}

// =====================================================================================================
// Bulk state <-> cells (state_to_locations / locations_to_state, mapf_env.py:358-371)
// =====================================================================================================
__device__ __forceinline__ void load_state(const DevSpec &sp, const u64 *states, i64 b, u64 &lo, u64 &hi) {
    if (sp.words == 1) { lo = states[b]; hi = 0; }
    else { ulonglong2 v = reinterpret_cast<const ulonglong2 *>(states)[b]; lo = v.x; hi = v.y; }
}
__device__ __forceinline__ void store_state(const DevSpec &sp, u64 *states, i64 b, u64 lo, u64 hi) {
    if (sp.words == 1) states[b] = lo;
    else reinterpret_cast<ulonglong2 *>(states)[b] = make_ulonglong2(lo, hi);
}

template <int N>
__global__ void __launch_bounds__(256) k_decode(DevSpec sp, const u64 *__restrict__ states, i64 B, int *__restrict__ cells) {
    for (i64 b = (i64)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (i64)gridDim.x * blockDim.x) {
        u64 lo, hi;
        load_state(sp, states, b, lo, hi);
        int cell[N];
        decode_state<N>(sp, lo, hi, cell);
#pragma unroll
        for (int i = 0; i < N; ++i) cells[b * N + i] = cell[i];
    }
}

template <int N>
__global__ void __launch_bounds__(256) k_encode(DevSpec sp, const int *__restrict__ cells, i64 B, u64 *__restrict__ states) {
    for (i64 b = (i64)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (i64)gridDim.x * blockDim.x) {
        int cell[N];
#pragma unroll
        for (int i = 0; i < N; ++i) cell[i] = cells[b * N + i];
        u64 lo, hi;
        encode_state<N>(sp, cell, lo, hi);
        store_state(sp, states, b, lo, hi);
    }
}

// =====================================================================================================
// Row sources: explicit (state, action) pairs, or a slab of the full table generated from the row index
// =====================================================================================================
template <bool RANGE>
__device__ __forceinline__ void row_input(const DevSpec &sp, const u64 *states, const int *actions, u64 sb_lo, u64 sb_hi,
                                          i64 b, u64 &lo, u64 &hi, u32 &a) {
    if (RANGE) {
        u64 off = (u64)b / sp.nA;  // nA is a kernel-uniform constant; one 64-bit division per ROW
        a = (u32)((u64)b - off * sp.nA);
        lo = sb_lo + off;
        hi = sb_hi + (lo < sb_lo ? 1ull : 0ull);
    } else {
        load_state(sp, states, b, lo, hi);
        a = (u32)actions[b];
    }
}

// len(P[s][a]) (mapf_env.py:448-479): 1 for a terminal state, else the product of merged-outcome counts
template <int N, bool RANGE>
__global__ void __launch_bounds__(256) k_count(DevSpec sp, const u64 *__restrict__ states, const int *__restrict__ actions,
                                               u64 sb_lo, u64 sb_hi, i64 B, i64 *__restrict__ row_len) {
    for (i64 b = (i64)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (i64)gridDim.x * blockDim.x) {
        u64 lo, hi;
        u32 a;
        row_input<RANGE>(sp, states, actions, sb_lo, sb_hi, b, lo, hi, a);
        int cell[N], act[N];
        decode_state<N>(sp, lo, hi, cell);
        decode_action<N>(a, act);
        i64 len = 1;
        if (!is_terminal<N>(sp, cell)) {
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

__global__ void __launch_bounds__(256) k_scan_partials(const i64 *__restrict__ in, i64 B, i64 *__restrict__ partial) {
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
__global__ void __launch_bounds__(256) k_scan_spine(i64 *partial, i64 n_chunks) {
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

__global__ void __launch_bounds__(256) k_scan_final(const i64 *__restrict__ in, i64 B, const i64 *__restrict__ partial,
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
    u16 prev[N][32];  // current cell of agent i
    u8 parked[32];    // SoC: agents parked on their goal choosing STAY
    u8 term[32];
};

template <int N, bool LUTS, bool RANGE>
__global__ void __launch_bounds__(MAPF_MAX_THREADS)
k_expand(DevSpec sp, const u64 *__restrict__ states, const int *__restrict__ actions, u64 sb_lo, u64 sb_hi, i64 B,
         const i64 *__restrict__ row_ptr, u64 *__restrict__ next_state, double *__restrict__ prob,
         double *__restrict__ reward, u8 *__restrict__ flags) {
    extern __shared__ __align__(16) unsigned char smem[];
    SmemTables tb = stage_tables<LUTS>(sp, smem);
    const int lut_bytes = LUTS ? ((sp.L * 5 * 8 + 15) & ~15) : 0;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    ExpandSlab<N> &sl = reinterpret_cast<ExpandSlab<N> *>(smem + MAPF_SMEM_SMALL_BYTES + lut_bytes)[wid];
    const i64 n_batches = (B + 31) >> 5;
    const i64 warps_total = (i64)gridDim.x * (blockDim.x >> 5);
    for (i64 batch = (i64)blockIdx.x * (blockDim.x >> 5) + wid; batch < n_batches; batch += warps_total) {
        // ---------------- phase A
        const i64 b = batch * 32 + lane;
        u32 len = 0;
        if (b < B) {
            u64 lo, hi;
            u32 a;
            row_input<RANGE>(sp, states, actions, sb_lo, sb_hi, b, lo, hi, a);
            int cell[N], act[N];
            decode_state<N>(sp, lo, hi, cell);
            decode_action<N>(a, act);
            const bool term = is_terminal<N>(sp, cell);
            len = 1;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                u64 e = lut_get<LUTS>(tb.lut, cell[i] * 5 + act[i]);
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
                if (sp.words == 1) next_state[idx] = sl.st[0][r];
                else reinterpret_cast<ulonglong2 *>(next_state)[idx] = make_ulonglong2(sl.st[0][r], sl.st[1][r]);
                prob[idx] = 1.0;
                reward[idx] = 0.0;
                flags[idx] = 1;
                continue;
            }
            // outcome digits: itertools.product, the LAST agent's digit moves fastest (mapf_env.py:467)
            int nxt[N], prv[N];
            u32 pm[N];
#pragma unroll
            for (int i = N - 1; i >= 0; --i) {
                const u64 e = sl.ent[i][r];
                const u32 k = ENT_K(e);
                u32 d;
                if (k == 1) { d = 0; }
                else if (k == 2) { d = o & 1u; o >>= 1; }
                else { u32 q = __umulhi(o, 0xAAAAAAABu) >> 1; d = o - 3u * q; o = q; }
                nxt[i] = (int)((u32)(e >> (16 * d)) & 0xffffu);
                pm[i] = (u32)(e >> (48 + 3 * d)) & 7u;
                prv[i] = (int)sl.prev[i][r];
            }
            // probability: left-to-right product (mapf_env.py:468)
            double p = tb.probtab[pm[0]];
#pragma unroll
            for (int i = 1; i < N; ++i) p = __dmul_rn(p, tb.probtab[pm[i]]);
            // reward / done / collision (mapf_env.py:225-235): clash beats goal
            const bool clash = has_clash<N>(prv, nxt);
            bool goal = true;
#pragma unroll
            for (int i = 0; i < N; ++i) goal = goal && (nxt[i] == (int)sp.goal[i]);
            const int kind = clash ? 1 : (goal ? 2 : 0);
            u64 nlo, nhi;
            encode_state<N>(sp, nxt, nlo, nhi);
            if (sp.words == 1) next_state[idx] = nlo;
            else reinterpret_cast<ulonglong2 *>(next_state)[idx] = make_ulonglong2(nlo, nhi);
            prob[idx] = p;
            reward[idx] = tb.reward[kind * MAPF_REW_STRIDE + sl.parked[r]];
            flags[idx] = (u8)((kind != 0 ? 1 : 0) | (clash ? 2 : 0));
        }
        __syncwarp();
    }
}

// =====================================================================================================
// Checksums of a record array (mod 2**64), accumulated into out8 with atomics
// =====================================================================================================
__global__ void __launch_bounds__(256)
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
struct StepOut {
    u64 lo, hi;
    double reward, prob;
    u32 done, coll;
};

// One env-step with the cells already decoded.  `draw(i)` supplies agent i's uniform.
template <int N, bool LUTS, class Draw>
__device__ __forceinline__ void step_cells(const DevSpec &sp, const SmemTables &tb, int (&cell)[N], u32 a, Draw draw,
                                           double &reward, double &prob, u32 &done, u32 &coll, bool &terminal) {
    terminal = is_terminal<N>(sp, cell);
    if (terminal) {  // (s, 0, True, {"prob": 0})  (mapf_env.py:238-240); no draw is consumed
        reward = 0.0; prob = 0.0; done = 1; coll = 0;
        return;
    }
    int act[N], nxt[N];
    decode_action<N>(a, act);
    double total = 1.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const u64 e = lut_get<LUTS>(tb.lut, cell[i] * 5 + act[i]);
        const u32 k = ENT_K(e);
        const double u = draw(i);
        // categorical_sample: first index whose cumulative sum exceeds u, else 0 (mapf_env.py:255)
        const double c0 = tb.probtab[ENT_MASK(e, 0)];
        u32 pick = 0;
        if (k >= 2 && !(c0 > u)) {
            const double c1 = __dadd_rn(c0, tb.probtab[ENT_MASK(e, 1)]);
            if (c1 > u) pick = 1;
            else if (k >= 3) {
                const double c2 = __dadd_rn(c1, tb.probtab[ENT_MASK(e, 2)]);
                if (c2 > u) pick = 2;
            }
        }
        nxt[i] = (int)((u32)(e >> (16 * pick)) & 0xffffu);
        total = __dmul_rn(total, tb.probtab[(u32)(e >> (48 + 3 * pick)) & 7u]);  // mapf_env.py:257
    }
    const bool clash = has_clash<N>(cell, nxt);
    bool goal = true;
#pragma unroll
    for (int i = 0; i < N; ++i) goal = goal && (nxt[i] == (int)sp.goal[i]);
    const int kind = clash ? 1 : (goal ? 2 : 0);
    reward = tb.reward[kind * MAPF_REW_STRIDE + parked_agents<N>(sp, cell, act)];
    prob = total;
    done = kind != 0 ? 1 : 0;
    coll = clash ? 1 : 0;
#pragma unroll
    for (int i = 0; i < N; ++i) cell[i] = nxt[i];
}

template <int N>
struct PhiloxDraw {
    u32 w[((N + 3) / 4) * 4];
    __device__ __forceinline__ PhiloxDraw(u64 seed, u64 env, u64 step) {
#pragma unroll
        for (int b = 0; b < (N + 3) / 4; ++b) {
            Philox4 x = philox_block(seed, env, step, (u32)b);
#pragma unroll
            for (int q = 0; q < 4; ++q) w[b * 4 + q] = x.v[q];
        }
    }
    __device__ __forceinline__ double operator()(int i) const { return u32_to_uniform(w[i]); }
};

struct TapeDraw {
    const double *u;
    __device__ __forceinline__ double operator()(int i) const { return u[i]; }
};

__device__ __forceinline__ u32 random_action(const DevSpec &sp, u64 seed, u64 env, u64 step) {
    Philox4 x = philox_block(seed, env, step, 15u);
    return (u32)__umul64hi(((u64)x.v[0] << 32) | x.v[1], sp.nA);
}

template <int N, bool LUTS>
__global__ void __launch_bounds__(MAPF_MAX_THREADS)
k_step(DevSpec sp, const u64 *states, const int *__restrict__ actions, i64 B, const double *__restrict__ uniforms,
       u64 seed, u64 step, u64 env0, u32 opts, u64 *next_states, double *__restrict__ reward, double *__restrict__ prob,
       u8 *__restrict__ done, u8 *__restrict__ coll) {
    extern __shared__ __align__(16) unsigned char smem[];
    SmemTables tb = stage_tables<LUTS>(sp, smem);
    for (i64 b = (i64)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (i64)gridDim.x * blockDim.x) {
        u64 lo, hi;
        load_state(sp, states, b, lo, hi);
        const u32 a = (u32)actions[b];
        int cell[N];
        decode_state<N>(sp, lo, hi, cell);
        double r, p;
        u32 d, c;
        bool term;
        if (uniforms) {
            TapeDraw draw = {uniforms + b * N};
            step_cells<N, LUTS>(sp, tb, cell, a, draw, r, p, d, c, term);
        } else {
            // the draws are generated even for a terminal env (they are simply not used)
            PhiloxDraw<N> draw(seed, env0 + (u64)b, step);
            step_cells<N, LUTS>(sp, tb, cell, a, draw, r, p, d, c, term);
        }
        if (!term) encode_state<N>(sp, cell, lo, hi);
        if ((opts & 1u) && d) { lo = sp.s0[0]; hi = sp.s0[1]; }  // MAPF_OPT_AUTO_RESET
        store_state(sp, next_states, b, lo, hi);
        reward[b] = r;
        prob[b] = p;
        done[b] = (u8)d;
        coll[b] = (u8)c;
    }
}

// T steps per launch; the env's cells stay in registers between steps, each step's results go to slab t.
template <int N, bool LUTS>
__global__ void __launch_bounds__(MAPF_MAX_THREADS)
k_rollout(DevSpec sp, u64 *states, const int *__restrict__ actions, i64 T, i64 B, const double *__restrict__ uniforms,
          u64 seed, u64 step0, u64 env0, u32 opts, u64 *__restrict__ next_states, double *__restrict__ reward,
          double *__restrict__ prob, u8 *__restrict__ done, u8 *__restrict__ coll) {
    extern __shared__ __align__(16) unsigned char smem[];
    SmemTables tb = stage_tables<LUTS>(sp, smem);
    for (i64 b = (i64)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (i64)gridDim.x * blockDim.x) {
        u64 lo, hi;
        load_state(sp, states, b, lo, hi);
        int cell[N];
        decode_state<N>(sp, lo, hi, cell);
        for (i64 t = 0; t < T; ++t) {
            const i64 o = t * B + b;
            const u32 a = actions ? (u32)actions[o] : random_action(sp, seed, env0 + (u64)b, step0 + (u64)t);
            double r, p;
            u32 d, c;
            bool term;
            if (uniforms) {
                TapeDraw draw = {uniforms + o * N};
                step_cells<N, LUTS>(sp, tb, cell, a, draw, r, p, d, c, term);
            } else {
                PhiloxDraw<N> draw(seed, env0 + (u64)b, step0 + (u64)t);
                step_cells<N, LUTS>(sp, tb, cell, a, draw, r, p, d, c, term);
            }
            if (!term) encode_state<N>(sp, cell, lo, hi);
            if ((opts & 1u) && d) {
                lo = sp.s0[0]; hi = sp.s0[1];
#pragma unroll
                for (int i = 0; i < N; ++i) cell[i] = (int)sp.start[i];
            }
            store_state(sp, next_states, o, lo, hi);
            reward[o] = r;
            prob[o] = p;
            done[o] = (u8)d;
            coll[o] = (u8)c;
        }
        store_state(sp, states, b, lo, hi);
    }
}
