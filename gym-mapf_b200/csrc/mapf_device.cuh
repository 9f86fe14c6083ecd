// mapf_device.cuh -- device-side building blocks of the joint-transition engine (sm_100a).
//
// Reference citations are file:line relative to /root/reference/gym_mapf/envs/.
#pragma once
#include <stdint.h>

typedef unsigned long long u64;
typedef long long i64;
typedef unsigned int u32;
typedef unsigned short u16;
typedef unsigned char u8;

#define MAPF_MAXN 13
#define MAPF_REW_STRIDE 16  // reward table row: parked-agent count 0..13

// Exact unsigned 64-bit division by a run-time constant (Granlund-Montgomery round-up method).
struct FastDiv {
    u64 magic;
    u32 shift;
    u32 add;  // 1: the 65-bit-magic fix-up path
};

// Move-table entry for one (cell, intended action): what `single_agent_movements` returns (mapf_env.py:163-184).
//   bits  0..15 / 16..31 / 32..47  destination cell of merged outcome 0 / 1 / 2
//   bits 48..50 / 51..53 / 54..56  which of the candidates {intended=1, right-slip=2, left-slip=4} merged into it;
//                                  the mask indexes probtab[], whose entries are the candidates' probabilities
//                                  added in list order (mapf_env.py:177-179)
//   bits 57..58                    k = number of merged outcomes (1..3)
#define ENT_DEST(e, j) ((u32)((e) >> (16 * (j))) & 0xffffu)
#define ENT_MASK(e, j) ((u32)((e) >> (48 + 3 * (j))) & 7u)
#define ENT_K(e) ((u32)((e) >> 57) & 3u)

// Everything a hot kernel needs about one env spec; passed by value as a kernel parameter (constant bank).
struct DevSpec {
    int n;         // agents
    int L;         // free cells
    int words;     // 64-bit words per joint state (1 or 2)
    int soc;       // 1: sum-of-costs living reward (mapf_env.py:440-446)
    int H, Wd;     // grid height / width
    int lut_smem;  // 1: the move table is staged in shared memory
    int cand_mask; // bit j set: candidate j (intended, right, left) has probability > 0 (mapf_env.py:172)
    u64 nA;        // 5**n
    FastDiv divL;  // division by L
    u64 s0[2];     // start state
    const u64 *lut;  // [L*5] move table, global memory
    u16 goal[16];    // goal cell per agent (mapf_env.py:158)
    u16 start[16];
    double probtab[8];                      // indexed by candidate mask
    double reward[3 * MAPF_REW_STRIDE];     // [0: living, 1: clash + living, 2: goal + living][parked agents]
};

__device__ __forceinline__ u64 fastdiv(u64 x, const FastDiv &d) {
    if (d.magic == 0) return x >> d.shift;
    u64 q = __umul64hi(x, d.magic);
    if (d.add) {
        u64 t = ((x - q) >> 1) + q;
        return t >> d.shift;
    }
    return q >> d.shift;
}

// ---- joint state <-> per-agent cells: little-endian radix L, agent 0 least significant (__init__.py:50-79) ----
template <int N>
__device__ __forceinline__ void decode_state(const DevSpec &sp, u64 lo, u64 hi, int (&cell)[N]) {
    const u32 L = (u32)sp.L;
    if (sp.words == 1) {
        u64 x = lo;
#pragma unroll
        for (int i = 0; i < N - 1; ++i) {
            u64 q = fastdiv(x, sp.divL);
            cell[i] = (int)(x - q * L);
            x = q;
        }
        cell[N - 1] = (int)min(x, (u64)(L - 1));  // out-of-range states are rejected on the host; stay in bounds
    } else {
        // 128-bit / 32-bit long division over four 32-bit limbs (partial dividends stay below 2**48)
        u32 limb[4] = {(u32)lo, (u32)(lo >> 32), (u32)hi, (u32)(hi >> 32)};
#pragma unroll
        for (int i = 0; i < N; ++i) {
            u64 rem = 0;
#pragma unroll
            for (int w = 3; w >= 0; --w) {
                u64 cur = (rem << 32) | limb[w];
                u64 q = fastdiv(cur, sp.divL);
                rem = cur - q * L;
                limb[w] = (u32)q;
            }
            cell[i] = (int)rem;
        }
    }
}

template <int N>
__device__ __forceinline__ void encode_state(const DevSpec &sp, const int (&cell)[N], u64 &lo, u64 &hi) {
    const u64 L = (u64)sp.L;
    if (sp.words == 1) {
        u64 acc = (u64)cell[N - 1];
#pragma unroll
        for (int i = N - 2; i >= 0; --i) acc = acc * L + (u64)cell[i];
        lo = acc;
        hi = 0;
    } else {
        u64 alo = (u64)cell[N - 1], ahi = 0;
#pragma unroll
        for (int i = N - 2; i >= 0; --i) {
            u64 carry = __umul64hi(alo, L);
            ahi = ahi * L + carry;
            alo = alo * L;
            u64 t = alo + (u64)cell[i];
            ahi += (t < alo) ? 1ull : 0ull;
            alo = t;
        }
        lo = alo;
        hi = ahi;
    }
}

template <int N>
__device__ __forceinline__ void decode_action(u32 a, int (&act)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
        u32 q = a / 5u;
        act[i] = (int)(a - q * 5u);
        a = q;
    }
}

// is_terminal (mapf_env.py:210-223): two agents on one cell, or every agent on its own goal
template <int N>
__device__ __forceinline__ bool is_terminal(const DevSpec &sp, const int (&cell)[N]) {
    bool dup = false, all_goal = true;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        all_goal = all_goal && (cell[i] == (int)sp.goal[i]);
#pragma unroll
        for (int j = i + 1; j < N; ++j) dup = dup || (cell[i] == cell[j]);
    }
    return dup || all_goal;
}

// _is_collision_transition_from_local_states (mapf_env.py:378-389): swap or vertex conflict over all pairs
template <int N>
__device__ __forceinline__ bool has_clash(const int (&prev)[N], const int (&nxt)[N]) {
    bool c = false;
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = i + 1; j < N; ++j)
            c = c || (nxt[i] == nxt[j]) || (prev[i] == nxt[j] && prev[j] == nxt[i]);
    return c;
}

// number of agents parked on their goal that chose STAY (mapf_env.py:441-446); 0 under Makespan
template <int N>
__device__ __forceinline__ int parked_agents(const DevSpec &sp, const int (&prev)[N], const int (&act)[N]) {
    int k = 0;
    if (sp.soc) {
#pragma unroll
        for (int i = 0; i < N; ++i) k += (prev[i] == (int)sp.goal[i] && act[i] == 0) ? 1 : 0;
    }
    return k;
}

// ---- Philox4x32-10 (Salmon et al., SC'11), the counter-based generator of the device-side sampling mode -------
struct Philox4 {
    u32 v[4];
};
__device__ __forceinline__ Philox4 philox4x32_10(u32 c0, u32 c1, u32 c2, u32 c3, u32 k0, u32 k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        u32 hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        u32 hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        u32 n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    Philox4 out;
    out.v[0] = c0; out.v[1] = c1; out.v[2] = c2; out.v[3] = c3;
    return out;
}

// Counter layout of the sampling stream: (env low, env high, step low, step high<<8 | block); key = seed.
// block b < 8 supplies the slip draws of agents 4b..4b+3 (one 32-bit word w each, u = w * 2**-32);
// block 15 supplies the random-policy action (words 0,1 as a 64-bit fraction of nA).
__device__ __forceinline__ Philox4 philox_block(u64 seed, u64 env, u64 step, u32 block) {
    return philox4x32_10((u32)env, (u32)(env >> 32), (u32)step, ((u32)(step >> 32) << 8) | block, (u32)seed,
                         (u32)(seed >> 32));
}

__device__ __forceinline__ double u32_to_uniform(u32 w) { return (double)w * 2.3283064365386963e-10; }  // 2**-32

// ---- shared-memory staging of the move table and the small constant tables -----------------------------------
struct SmemTables {
    const u64 *lut;        // smem or global
    const double *probtab; // smem [8]
    const double *reward;  // smem [3*MAPF_REW_STRIDE]
};

// Layout of the dynamic shared memory of every hot kernel: [probtab 8 f64][reward 48 f64][lut L*5 u64 (if staged)]
#define MAPF_SMEM_SMALL_BYTES ((8 + 3 * MAPF_REW_STRIDE) * 8)

template <bool LUTS>
__device__ __forceinline__ SmemTables stage_tables(const DevSpec &sp, unsigned char *smem) {
    double *pt = reinterpret_cast<double *>(smem);
    double *rw = pt + 8;
    u64 *lut_s = reinterpret_cast<u64 *>(smem + MAPF_SMEM_SMALL_BYTES);
    for (int i = threadIdx.x; i < 8; i += blockDim.x) pt[i] = sp.probtab[i];
    for (int i = threadIdx.x; i < 3 * MAPF_REW_STRIDE; i += blockDim.x) rw[i] = sp.reward[i];
    SmemTables t;
    t.probtab = pt;
    t.reward = rw;
    if (LUTS) {
        const int n_ent = sp.L * 5;
        // 128-bit copies of the table (global -> shared); the table base is 16-byte aligned
        const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(sp.lut);
        ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(lut_s);
        for (int i = threadIdx.x; i < n_ent / 2; i += blockDim.x) dst[i] = __ldg(src + i);
        if ((n_ent & 1) && threadIdx.x == 0) lut_s[n_ent - 1] = __ldg(sp.lut + n_ent - 1);
        t.lut = lut_s;
    } else {
        t.lut = sp.lut;
    }
    __syncthreads();
    return t;
}

template <bool LUTS>
__device__ __forceinline__ u64 lut_get(const u64 *lut, int idx) {
    if (LUTS) return lut[idx];
    return __ldg(lut + idx);
}
