// mapf_capi.cu -- C ABI (include/mapf_b200.h) over the sm_100a kernels: context creation, host-side derivation of
// the env constants in the reference's floating-point order, kernel dispatch.
//
// Reference citations are file:line relative to /root/reference/gym_mapf/envs/.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "../../include/mapf_b200.h"
#include "mapf_kernels.cuh"

typedef unsigned __int128 u128;

// ---------------------------------------------------------------------------------------------------------------
// error reporting
// ---------------------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                                         \
    do {                                                                                                       \
        cudaError_t _e = (expr);                                                                               \
        if (_e != cudaSuccess) return fail(MAPF_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                                           __FILE__, __LINE__);                                                \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};

// ---------------------------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------------------------
struct KernelSet {
    const void *step, *rollout, *expand, *expand_range, *count, *count_range, *decode, *encode;
};

struct mapf_ctx {
    DevSpec sp;
    mapf_info info;
    int device = 0;
    int threads = 256;           // CTA size of the hot kernels
    size_t smem_base = 0;        // small tables + staged move table
    size_t smem_expand = 0;      // smem_base + per-warp expand slabs
    int grid_step = 0, grid_rollout = 0, grid_expand = 0, grid_expand_range = 0, grid_plain = 0;
    KernelSet ks;
    u64 *d_lut = nullptr;
    u32 *d_cell_rc = nullptr, *d_colbits = nullptr, *d_colbase = nullptr;
    std::vector<u64> h_lut;
    std::vector<u32> h_cell_rc;
    // mapf_step_host staging
    std::mutex mu;
    cudaStream_t hs[2] = {nullptr, nullptr};
    unsigned char *d_stage = nullptr;
    size_t d_stage_bytes = 0;
};

template <int N>
static KernelSet kernels_for(bool luts) {
    KernelSet k;
    if (luts) {
        k.step = (const void *)k_step<N, true>;
        k.rollout = (const void *)k_rollout<N, true>;
        k.expand = (const void *)k_expand<N, true, false>;
        k.expand_range = (const void *)k_expand<N, true, true>;
    } else {
        k.step = (const void *)k_step<N, false>;
        k.rollout = (const void *)k_rollout<N, false>;
        k.expand = (const void *)k_expand<N, false, false>;
        k.expand_range = (const void *)k_expand<N, false, true>;
    }
    k.count = (const void *)k_count<N, false>;
    k.count_range = (const void *)k_count<N, true>;
    k.decode = (const void *)k_decode<N>;
    k.encode = (const void *)k_encode<N>;
    return k;
}

template <int N>
static size_t expand_slab_bytes() { return sizeof(ExpandSlab<N>); }

static KernelSet pick_kernels(int n, bool luts, size_t *slab) {
    switch (n) {
#define CASE(N) case N: *slab = expand_slab_bytes<N>(); return kernels_for<N>(luts);
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13)
#undef CASE
    }
    *slab = 0;
    return KernelSet();
}

static FastDiv make_fastdiv(u64 d) {
    FastDiv f;
    f.magic = 0; f.shift = 0; f.add = 0;
    int fl = 63 - __builtin_clzll(d);
    if ((d & (d - 1)) == 0) { f.shift = (u32)fl; return f; }
    u128 num = (u128)1 << (64 + fl);
    u64 proposed = (u64)(num / d);
    u64 rem = (u64)(num % d);
    u64 e = d - rem;
    if (e < (1ull << fl)) {
        f.shift = (u32)fl;
    } else {
        proposed += proposed;
        u64 twice = rem + rem;
        if (twice >= d || twice < rem) proposed += 1;
        f.shift = (u32)fl;
        f.add = 1;
    }
    f.magic = proposed + 1;
    return f;
}

extern "C" const char *mapf_last_error(void) { return g_err; }
extern "C" const char *mapf_version(void) { return "mapf_b200 0.1 (sm_100a)"; }

extern "C" void mapf_ctx_destroy(mapf_ctx *ctx) {
    if (!ctx) return;
    {
        DeviceGuard g(ctx->device);
        cudaFree(ctx->d_lut);
        cudaFree(ctx->d_cell_rc);
        cudaFree(ctx->d_colbits);
        cudaFree(ctx->d_colbase);
        cudaFree(ctx->d_stage);
        for (int i = 0; i < 2; ++i)
            if (ctx->hs[i]) cudaStreamDestroy(ctx->hs[i]);
    }
    delete ctx;
}

static int occupancy_grid(const void *fn, int threads, size_t smem, int sm_count, int *grid) {
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, smem));
    if (per_sm < 1) return fail(MAPF_ERR_CUDA, "kernel does not fit on an SM (threads=%d smem=%zu)", threads, smem);
    *grid = per_sm * sm_count;
    return MAPF_OK;
}

extern "C" int mapf_ctx_create(const mapf_spec *spec, int device, mapf_ctx **out) {
    if (!spec || !out || !spec->obstacles || !spec->start_rc || !spec->goal_rc)
        return fail(MAPF_ERR_INVALID, "mapf_ctx_create: NULL argument");
    *out = nullptr;
    const int H = spec->height, W = spec->width, n = spec->n_agents;
    if (H < 1 || W < 1) return fail(MAPF_ERR_INVALID, "empty grid");
    if (n < 1 || n > MAPF_MAX_AGENTS)
        return fail(MAPF_ERR_UNSUPPORTED, "%d agents: this build supports 1..%d", n, MAPF_MAX_AGENTS);
    if (spec->criterion != MAPF_SOC && spec->criterion != MAPF_MAKESPAN)
        return fail(MAPF_ERR_INVALID, "unknown optimisation criterion %d", spec->criterion);
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev < 1 || device < 0 || device >= n_dev)
        return fail(MAPF_ERR_NO_DEVICE, "CUDA device %d is not available (%d devices): there is no CPU fallback",
                    device, n_dev);

    // ---- column-major free-cell numbering (grid.py:37-40, mapf_env.py:142-143) and the column bitmap
    const int wpc = (H + 31) / 32;
    std::vector<u32> colbits((size_t)W * wpc, 0u), colbase(W, 0u);
    std::vector<int> id_of((size_t)H * W, -1);
    int L = 0;
    for (int c = 0; c < W; ++c) {
        colbase[c] = (u32)L;
        for (int r = 0; r < H; ++r)
            if (!spec->obstacles[(size_t)r * W + c]) {
                colbits[(size_t)c * wpc + (r >> 5)] |= 1u << (r & 31);
                id_of[(size_t)r * W + c] = L++;
            }
    }
    if (L < 1) return fail(MAPF_ERR_INVALID, "the grid has no free cell");
    if (L > MAPF_MAX_CELLS) return fail(MAPF_ERR_UNSUPPORTED, "%d free cells: this build supports up to %d", L,
                                        MAPF_MAX_CELLS);

    mapf_ctx *ctx = new (std::nothrow) mapf_ctx();
    if (!ctx) return fail(MAPF_ERR_INVALID, "out of host memory");
    ctx->device = device;
    DevSpec &sp = ctx->sp;
    memset(&sp, 0, sizeof(sp));
    sp.n = n; sp.L = L; sp.H = H; sp.Wd = W;
    sp.soc = spec->criterion == MAPF_SOC ? 1 : 0;

    // ---- starts / goals must be free cells: the reference raises KeyError (mapf_env.py:155,158,369)
    int start_id[MAPF_MAXN], goal_id[MAPF_MAXN];
    for (int i = 0; i < n; ++i) {
        const int sr = spec->start_rc[2 * i], sc = spec->start_rc[2 * i + 1];
        const int gr = spec->goal_rc[2 * i], gc = spec->goal_rc[2 * i + 1];
        if (sr < 0 || sr >= H || sc < 0 || sc >= W || id_of[(size_t)sr * W + sc] < 0) {
            delete ctx;
            return fail(MAPF_ERR_KEY, "(%d, %d)", sr, sc);
        }
        if (gr < 0 || gr >= H || gc < 0 || gc >= W || id_of[(size_t)gr * W + gc] < 0) {
            delete ctx;
            return fail(MAPF_ERR_KEY, "(%d, %d)", gr, gc);
        }
        start_id[i] = id_of[(size_t)sr * W + sc];
        goal_id[i] = id_of[(size_t)gr * W + gc];
        sp.start[i] = (u16)start_id[i];
        sp.goal[i] = (u16)goal_id[i];
    }

    // ---- nS = L**n, nA = 5**n (mapf_env.py:145-146); state width
    u128 nS = 1;
    for (int i = 0; i < n; ++i) {
        if (nS > (((u128)1 << 127) / (u128)L)) {
            delete ctx;
            return fail(MAPF_ERR_UNSUPPORTED, "L**n = %d**%d does not fit 127 bits", L, n);
        }
        nS *= (u128)L;
    }
    sp.words = nS < ((u128)1 << 63) ? 1 : 2;
    u64 nA = 1;
    for (int i = 0; i < n; ++i) nA *= 5;
    sp.nA = nA;
    sp.divL = make_fastdiv((u64)L);
    u128 s0 = 0, sg = 0, mul = 1;
    for (int i = 0; i < n; ++i) {  // vector_to_integer (__init__.py:70-79)
        s0 += (u128)start_id[i] * mul;
        sg += (u128)goal_id[i] * mul;
        mul *= (u128)L;
    }
    sp.s0[0] = (u64)s0; sp.s0[1] = (u64)(s0 >> 64);

    // ---- probabilities in the reference's order (mapf_env.py:131-132,168-170,177-179)
    const double rf = spec->fail_prob / 2, lf = spec->fail_prob / 2;
    const double cand[3] = {1 - rf - lf, rf, lf};
    sp.cand_mask = 0;
    for (int j = 0; j < 3; ++j)
        if (cand[j] > 0) sp.cand_mask |= 1 << j;
    if (sp.cand_mask == 0) {
        delete ctx;
        return fail(MAPF_ERR_INVALID, "fail_prob %g leaves no outcome with positive probability", spec->fail_prob);
    }
    for (int m = 1; m < 8; ++m) {
        double s = 0.0;
        bool first = true;
        for (int j = 0; j < 3; ++j)
            if ((m >> j) & 1) { s = first ? cand[j] : s + cand[j]; first = false; }
        sp.probtab[m] = s;
    }
    // ---- rewards (mapf_env.py:225-235, 436-446), one entry per number of parked agents
    for (int k = 0; k <= n; ++k) {
        const double live = sp.soc ? (double)(n - k) * spec->reward_of_living : spec->reward_of_living;
        sp.reward[0 * MAPF_REW_STRIDE + k] = live;
        sp.reward[1 * MAPF_REW_STRIDE + k] = spec->reward_of_clash + live;
        sp.reward[2 * MAPF_REW_STRIDE + k] = spec->reward_of_goal + live;
    }

    // ---- device side
    DeviceGuard guard(device);
    if (!guard.ok) { delete ctx; return fail(MAPF_ERR_CUDA, "cannot select device %d", device); }
    cudaDeviceProp prop;
    cudaError_t ce = cudaGetDeviceProperties(&prop, device);
    if (ce != cudaSuccess) { delete ctx; return fail(MAPF_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(ce)); }
    if (prop.major < 10) {
        delete ctx;
        return fail(MAPF_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                    prop.minor);
    }
    const int sm_count = prop.multiProcessorCount;
    const size_t lut_bytes = (size_t)L * 5 * sizeof(u64);
#define CTX_TRY(expr)                                                                                  \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess) {                                                                       \
            int _rc = fail(MAPF_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            mapf_ctx_destroy(ctx);                                                                     \
            return _rc;                                                                                \
        }                                                                                              \
    } while (0)
    CTX_TRY(cudaMalloc(&ctx->d_lut, (lut_bytes + 15) & ~(size_t)15));
    CTX_TRY(cudaMalloc(&ctx->d_cell_rc, (size_t)L * sizeof(u32)));
    CTX_TRY(cudaMalloc(&ctx->d_colbits, colbits.size() * sizeof(u32)));
    CTX_TRY(cudaMalloc(&ctx->d_colbase, colbase.size() * sizeof(u32)));
    CTX_TRY(cudaMemcpy(ctx->d_colbits, colbits.data(), colbits.size() * sizeof(u32), cudaMemcpyHostToDevice));
    CTX_TRY(cudaMemcpy(ctx->d_colbase, colbase.data(), colbase.size() * sizeof(u32), cudaMemcpyHostToDevice));
    {
        const size_t bm_smem = ((size_t)W * wpc + W) * sizeof(u32);
        if (bm_smem > (size_t)prop.sharedMemPerBlockOptin) {
            mapf_ctx_destroy(ctx);
            return fail(MAPF_ERR_UNSUPPORTED, "the %dx%d obstacle bitmap (%zu B) does not fit shared memory", H, W, bm_smem);
        }
        if (bm_smem > 48 * 1024)
            CTX_TRY(cudaFuncSetAttribute(k_build_moves, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bm_smem));
        const int blocks = (H * W + 255) / 256 < sm_count * 4 ? (H * W + 255) / 256 : sm_count * 4;
        k_build_moves<<<blocks, 256, bm_smem>>>(ctx->d_colbits, ctx->d_colbase, H, W, wpc, sp.cand_mask, ctx->d_lut,
                                                ctx->d_cell_rc);
        CTX_TRY(cudaGetLastError());
        CTX_TRY(cudaDeviceSynchronize());
    }
    ctx->h_lut.resize((size_t)L * 5);
    ctx->h_cell_rc.resize(L);
    CTX_TRY(cudaMemcpy(ctx->h_lut.data(), ctx->d_lut, lut_bytes, cudaMemcpyDeviceToHost));
    CTX_TRY(cudaMemcpy(ctx->h_cell_rc.data(), ctx->d_cell_rc, (size_t)L * sizeof(u32), cudaMemcpyDeviceToHost));
    sp.lut = ctx->d_lut;

    // ---- launch geometry: stage the move table in shared memory when it leaves room for >= 1 CTA per SM
    size_t slab = 0;
    const size_t lut_pad = (lut_bytes + 15) & ~(size_t)15;
    const size_t smem_limit = (size_t)prop.sharedMemPerBlockOptin;
    bool luts = MAPF_SMEM_SMALL_BYTES + lut_pad + 16 * expand_slab_bytes<MAPF_MAXN>() <= smem_limit;
    ctx->threads = (luts && lut_pad > 64 * 1024) ? 512 : 256;
    sp.lut_smem = luts ? 1 : 0;
    ctx->ks = pick_kernels(n, luts, &slab);
    ctx->smem_base = MAPF_SMEM_SMALL_BYTES + (luts ? lut_pad : 0);
    ctx->smem_expand = ctx->smem_base + (size_t)(ctx->threads / 32) * slab;
    const void *big[4] = {ctx->ks.step, ctx->ks.rollout, ctx->ks.expand, ctx->ks.expand_range};
    for (int i = 0; i < 4; ++i) {
        const size_t need = i < 2 ? ctx->smem_base : ctx->smem_expand;
        if (need > 48 * 1024) CTX_TRY(cudaFuncSetAttribute(big[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    }
    int rc;
    if ((rc = occupancy_grid(ctx->ks.step, ctx->threads, ctx->smem_base, sm_count, &ctx->grid_step)) ||
        (rc = occupancy_grid(ctx->ks.rollout, ctx->threads, ctx->smem_base, sm_count, &ctx->grid_rollout)) ||
        (rc = occupancy_grid(ctx->ks.expand, ctx->threads, ctx->smem_expand, sm_count, &ctx->grid_expand)) ||
        (rc = occupancy_grid(ctx->ks.expand_range, ctx->threads, ctx->smem_expand, sm_count, &ctx->grid_expand_range))) {
        mapf_ctx_destroy(ctx);
        return rc;
    }
    ctx->grid_plain = sm_count * 8;

    mapf_info &inf = ctx->info;
    memset(&inf, 0, sizeof(inf));
    inf.n_agents = n; inf.n_cells = L; inf.state_words = sp.words; inf.moves_in_smem = sp.lut_smem;
    inf.n_actions = (int64_t)nA;
    inf.n_states[0] = (u64)nS; inf.n_states[1] = (u64)(nS >> 64);
    inf.start_state[0] = (u64)s0; inf.start_state[1] = (u64)(s0 >> 64);
    inf.goal_state[0] = (u64)sg; inf.goal_state[1] = (u64)(sg >> 64);
    int64_t mr = 1;
    for (int i = 0; i < n; ++i) mr *= 3;
    inf.max_row_len = mr;
    inf.device = device; inf.sm_count = sm_count;
    *out = ctx;
    return MAPF_OK;
}

extern "C" int mapf_ctx_info(const mapf_ctx *ctx, mapf_info *out) {
    if (!ctx || !out) return fail(MAPF_ERR_INVALID, "mapf_ctx_info: NULL argument");
    *out = ctx->info;
    return MAPF_OK;
}

extern "C" int mapf_ctx_moves(const mapf_ctx *ctx, uint8_t *k, int32_t *dest, double *prob, int32_t *cells_rc) {
    if (!ctx) return fail(MAPF_ERR_INVALID, "mapf_ctx_moves: NULL context");
    const int L = ctx->sp.L;
    for (int i = 0; i < L * 5; ++i) {
        const u64 e = ctx->h_lut[i];
        const int kk = (int)ENT_K(e);
        if (k) k[i] = (uint8_t)kk;
        for (int j = 0; j < 3; ++j) {
            if (dest) dest[i * 3 + j] = j < kk ? (int32_t)ENT_DEST(e, j) : -1;
            if (prob) prob[i * 3 + j] = j < kk ? ctx->sp.probtab[ENT_MASK(e, j)] : 0.0;
        }
    }
    if (cells_rc)
        for (int i = 0; i < L; ++i) {
            cells_rc[2 * i] = (int32_t)(ctx->h_cell_rc[i] >> 16);
            cells_rc[2 * i + 1] = (int32_t)(ctx->h_cell_rc[i] & 0xffffu);
        }
    return MAPF_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// launches
// ---------------------------------------------------------------------------------------------------------------
static int grid_for(int64_t work_items, int threads, int grid_max) {
    int64_t need = (work_items + threads - 1) / threads;
    if (need < 1) need = 1;
    return (int)(need < grid_max ? need : grid_max);
}

#define LAUNCH(fn, grid, threads, smem, stream, args)                                               \
    do {                                                                                            \
        cudaError_t _e = cudaLaunchKernel(fn, dim3(grid), dim3(threads), args, smem, (cudaStream_t)(stream)); \
        if (_e != cudaSuccess) return fail(MAPF_ERR_CUDA, "launch %s: %s", #fn, cudaGetErrorString(_e));      \
    } while (0)

extern "C" int mapf_decode_states(const mapf_ctx *ctx, const void *states, int64_t B, int32_t *cells, void *stream) {
    if (!ctx || B < 0 || (B > 0 && (!states || !cells))) return fail(MAPF_ERR_INVALID, "mapf_decode_states: bad argument");
    if (B == 0) return MAPF_OK;
    DeviceGuard g(ctx->device);
    DevSpec sp = ctx->sp;
    void *args[] = {&sp, &states, &B, &cells};
    LAUNCH(ctx->ks.decode, grid_for(B, 256, ctx->grid_plain), 256, 0, stream, args);
    return MAPF_OK;
}

extern "C" int mapf_encode_states(const mapf_ctx *ctx, const int32_t *cells, int64_t B, void *states, void *stream) {
    if (!ctx || B < 0 || (B > 0 && (!states || !cells))) return fail(MAPF_ERR_INVALID, "mapf_encode_states: bad argument");
    if (B == 0) return MAPF_OK;
    DeviceGuard g(ctx->device);
    DevSpec sp = ctx->sp;
    void *args[] = {&sp, &cells, &B, &states};
    LAUNCH(ctx->ks.encode, grid_for(B, 256, ctx->grid_plain), 256, 0, stream, args);
    return MAPF_OK;
}

static int count_impl(const mapf_ctx *ctx, bool range, const void *states, const int32_t *actions, u64 sb_lo, u64 sb_hi,
                      int64_t B, int64_t *row_len, void *stream) {
    if (B == 0) return MAPF_OK;
    DeviceGuard g(ctx->device);
    DevSpec sp = ctx->sp;
    void *args[] = {&sp, &states, &actions, &sb_lo, &sb_hi, &B, &row_len};
    LAUNCH(range ? ctx->ks.count_range : ctx->ks.count, grid_for(B, 256, ctx->grid_plain), 256, 0, stream, args);
    return MAPF_OK;
}

extern "C" int mapf_count_rows(const mapf_ctx *ctx, const void *states, const int32_t *actions, int64_t B,
                               int64_t *row_len, void *stream) {
    if (!ctx || B < 0 || (B > 0 && (!states || !actions || !row_len)))
        return fail(MAPF_ERR_INVALID, "mapf_count_rows: bad argument");
    return count_impl(ctx, false, states, actions, 0, 0, B, row_len, stream);
}

static int range_rows(const mapf_ctx *ctx, int64_t n_states, int64_t *B) {
    if (n_states < 0) return fail(MAPF_ERR_INVALID, "negative state count");
    if ((u128)n_states * (u128)ctx->sp.nA > (u128)0x7fffffffffffffffLL / 4)
        return fail(MAPF_ERR_INVALID, "slab of %lld states x %llu actions is too large for one call", (long long)n_states,
                    (unsigned long long)ctx->sp.nA);
    *B = n_states * (int64_t)ctx->sp.nA;
    return MAPF_OK;
}

extern "C" int mapf_count_range(const mapf_ctx *ctx, const uint64_t s_begin[2], int64_t n_states, int64_t *row_len,
                                void *stream) {
    if (!ctx || !s_begin || (n_states > 0 && !row_len)) return fail(MAPF_ERR_INVALID, "mapf_count_range: bad argument");
    int64_t B = 0;
    int rc = range_rows(ctx, n_states, &B);
    if (rc) return rc;
    return count_impl(ctx, true, nullptr, nullptr, s_begin[0], s_begin[1], B, row_len, stream);
}

extern "C" int64_t mapf_scan_scratch_bytes(int64_t B) {
    int64_t chunks = (B + SCAN_CHUNK - 1) / SCAN_CHUNK;
    if (chunks < 1) chunks = 1;
    return (chunks + 1) * (int64_t)sizeof(int64_t);
}

extern "C" int mapf_scan_rows(const mapf_ctx *ctx, const int64_t *row_len, int64_t B, int64_t *row_ptr, void *scratch,
                              void *stream) {
    if (!ctx || B < 0 || !row_ptr || (B > 0 && (!row_len || !scratch)))
        return fail(MAPF_ERR_INVALID, "mapf_scan_rows: bad argument");
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (B == 0) {
        CUDA_TRY(cudaMemsetAsync(row_ptr, 0, sizeof(int64_t), st));
        return MAPF_OK;
    }
    const int64_t chunks = (B + SCAN_CHUNK - 1) / SCAN_CHUNK;
    if (chunks > 0x7fffffff) return fail(MAPF_ERR_INVALID, "too many rows for one scan");
    i64 *partial = (i64 *)scratch;
    k_scan_partials<<<(int)chunks, 256, 0, st>>>((const i64 *)row_len, B, partial);
    k_scan_spine<<<1, 256, 0, st>>>(partial, chunks);
    k_scan_final<<<(int)chunks, 256, 0, st>>>((const i64 *)row_len, B, partial, (i64 *)row_ptr);
    CUDA_TRY(cudaGetLastError());
    return MAPF_OK;
}

static int expand_impl(const mapf_ctx *ctx, bool range, const void *states, const int32_t *actions, u64 sb_lo, u64 sb_hi,
                       int64_t B, const int64_t *row_ptr, void *next_state, double *prob, double *reward, uint8_t *flags,
                       void *stream) {
    if (B == 0) return MAPF_OK;
    if (!row_ptr || !next_state || !prob || !reward || !flags) return fail(MAPF_ERR_INVALID, "expand: NULL output");
    DeviceGuard g(ctx->device);
    DevSpec sp = ctx->sp;
    void *args[] = {&sp, &states, &actions, &sb_lo, &sb_hi, &B, &row_ptr, &next_state, &prob, &reward, &flags};
    const int64_t warps_needed = (B + 31) / 32;
    const int grid = grid_for(warps_needed * 32, ctx->threads, range ? ctx->grid_expand_range : ctx->grid_expand);
    LAUNCH(range ? ctx->ks.expand_range : ctx->ks.expand, grid, ctx->threads, ctx->smem_expand, stream, args);
    return MAPF_OK;
}

extern "C" int mapf_expand(const mapf_ctx *ctx, const void *states, const int32_t *actions, int64_t B,
                           const int64_t *row_ptr, void *next_state, double *prob, double *reward, uint8_t *flags,
                           void *stream) {
    if (!ctx || B < 0 || (B > 0 && (!states || !actions))) return fail(MAPF_ERR_INVALID, "mapf_expand: bad argument");
    return expand_impl(ctx, false, states, actions, 0, 0, B, row_ptr, next_state, prob, reward, flags, stream);
}

extern "C" int mapf_expand_range(const mapf_ctx *ctx, const uint64_t s_begin[2], int64_t n_states, const int64_t *row_ptr,
                                 void *next_state, double *prob, double *reward, uint8_t *flags, void *stream) {
    if (!ctx || !s_begin) return fail(MAPF_ERR_INVALID, "mapf_expand_range: bad argument");
    int64_t B = 0;
    int rc = range_rows(ctx, n_states, &B);
    if (rc) return rc;
    return expand_impl(ctx, true, nullptr, nullptr, s_begin[0], s_begin[1], B, row_ptr, next_state, prob, reward, flags,
                       stream);
}

extern "C" int mapf_checksum(const mapf_ctx *ctx, int64_t n_records, int64_t index_base, const void *next_state,
                             const double *prob, const double *reward, const uint8_t *flags, uint64_t *out8, void *stream) {
    if (!ctx || n_records < 0 || !out8 || (n_records > 0 && (!next_state || !prob || !reward || !flags)))
        return fail(MAPF_ERR_INVALID, "mapf_checksum: bad argument");
    if (n_records == 0) return MAPF_OK;
    DeviceGuard g(ctx->device);
    k_checksum<<<grid_for(n_records, 256, ctx->grid_plain), 256, 0, (cudaStream_t)stream>>>(
        ctx->sp.words, n_records, index_base, (const u64 *)next_state, prob, reward, flags, (u64 *)out8);
    CUDA_TRY(cudaGetLastError());
    return MAPF_OK;
}

extern "C" int mapf_step(const mapf_ctx *ctx, const void *states, const int32_t *actions, int64_t B,
                         const double *uniforms, uint64_t seed, uint64_t step_index, int64_t env_offset,
                         uint32_t options, void *next_states, double *reward, double *prob, uint8_t *done,
                         uint8_t *collision, void *stream) {
    if (!ctx || B < 0 || (B > 0 && (!states || !actions || !next_states || !reward || !prob || !done || !collision)))
        return fail(MAPF_ERR_INVALID, "mapf_step: bad argument");
    if (B == 0) return MAPF_OK;
    DeviceGuard g(ctx->device);
    DevSpec sp = ctx->sp;
    u64 sd = seed, st = step_index, e0 = (u64)env_offset;
    u32 op = options;
    void *args[] = {&sp, &states, &actions, &B, &uniforms, &sd, &st, &e0, &op, &next_states, &reward, &prob, &done,
                    &collision};
    LAUNCH(ctx->ks.step, grid_for(B, ctx->threads, ctx->grid_step), ctx->threads, ctx->smem_base, stream, args);
    return MAPF_OK;
}

extern "C" int mapf_rollout(const mapf_ctx *ctx, void *states_inout, const int32_t *actions, int64_t T, int64_t B,
                            const double *uniforms, uint64_t seed, uint64_t step_index0, int64_t env_offset,
                            uint32_t options, void *next_states, double *reward, double *prob, uint8_t *done,
                            uint8_t *collision, void *stream) {
    if (!ctx || B < 0 || T < 0 ||
        (B > 0 && T > 0 && (!states_inout || !next_states || !reward || !prob || !done || !collision)))
        return fail(MAPF_ERR_INVALID, "mapf_rollout: bad argument");
    if (B == 0 || T == 0) return MAPF_OK;
    DeviceGuard g(ctx->device);
    DevSpec sp = ctx->sp;
    u64 sd = seed, st = step_index0, e0 = (u64)env_offset;
    u32 op = options;
    void *args[] = {&sp, &states_inout, &actions, &T, &B, &uniforms, &sd, &st, &e0, &op, &next_states, &reward, &prob,
                    &done, &collision};
    LAUNCH(ctx->ks.rollout, grid_for(B, ctx->threads, ctx->grid_rollout), ctx->threads, ctx->smem_base, stream, args);
    return MAPF_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// host-buffer step: H2D, step, D2H in two pipelined halves on the context's own streams
// ---------------------------------------------------------------------------------------------------------------
extern "C" int mapf_step_host(mapf_ctx *ctx, const void *states, const int32_t *actions, int64_t B,
                              const double *uniforms, uint64_t seed, uint64_t step_index, int64_t env_offset,
                              uint32_t options, void *next_states, double *reward, double *prob, uint8_t *done,
                              uint8_t *collision) {
    if (!ctx || B < 0 || (B > 0 && (!states || !actions || !next_states || !reward || !prob || !done || !collision)))
        return fail(MAPF_ERR_INVALID, "mapf_step_host: bad argument");
    if (B == 0) return MAPF_OK;
    std::lock_guard<std::mutex> lock(ctx->mu);
    DeviceGuard g(ctx->device);
    const int n = ctx->sp.n;
    const size_t sw = (size_t)ctx->sp.words * 8;
    // per-env device bytes: state in, action, uniforms, state out, reward, prob, done, collision
    const size_t per_env = sw + 4 + (uniforms ? (size_t)n * 8 : 0) + sw + 8 + 8 + 1 + 1;
    const size_t need = per_env * (size_t)B + 8 * 256;
    if (need > ctx->d_stage_bytes) {
        if (ctx->d_stage) cudaFree(ctx->d_stage);
        ctx->d_stage = nullptr;
        ctx->d_stage_bytes = 0;
        CUDA_TRY(cudaMalloc(&ctx->d_stage, need));
        ctx->d_stage_bytes = need;
    }
    for (int i = 0; i < 2; ++i)
        if (!ctx->hs[i]) CUDA_TRY(cudaStreamCreateWithFlags(&ctx->hs[i], cudaStreamNonBlocking));
    auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
    unsigned char *p = ctx->d_stage;
    unsigned char *d_s = p; p += align(sw * B);
    unsigned char *d_a = p; p += align(4 * (size_t)B);
    unsigned char *d_u = p; p += uniforms ? align((size_t)n * 8 * B) : 0;
    unsigned char *d_ns = p; p += align(sw * B);
    unsigned char *d_r = p; p += align(8 * (size_t)B);
    unsigned char *d_p = p; p += align(8 * (size_t)B);
    unsigned char *d_d = p; p += align((size_t)B);
    unsigned char *d_c = p;
    const int parts = B >= (1 << 16) ? 2 : 1;
    for (int h = 0; h < parts; ++h) {
        const int64_t b0 = B * h / parts, b1 = B * (h + 1) / parts, nb = b1 - b0;
        cudaStream_t st = ctx->hs[h];
        CUDA_TRY(cudaMemcpyAsync(d_s + sw * b0, (const unsigned char *)states + sw * b0, sw * nb, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_a + 4 * b0, (const unsigned char *)actions + 4 * b0, 4 * nb, cudaMemcpyHostToDevice, st));
        if (uniforms)
            CUDA_TRY(cudaMemcpyAsync(d_u + (size_t)n * 8 * b0, (const unsigned char *)uniforms + (size_t)n * 8 * b0,
                                     (size_t)n * 8 * nb, cudaMemcpyHostToDevice, st));
        DevSpec sp = ctx->sp;
        const void *a_states = d_s + sw * b0;
        const int32_t *a_actions = (const int32_t *)(d_a + 4 * b0);
        const double *a_u = uniforms ? (const double *)(d_u + (size_t)n * 8 * b0) : nullptr;
        void *a_ns = d_ns + sw * b0;
        double *a_r = (double *)(d_r + 8 * b0), *a_p = (double *)(d_p + 8 * b0);
        uint8_t *a_d = d_d + b0, *a_c = d_c + b0;
        int64_t nbv = nb;
        u64 sd = seed, stp = step_index, e0 = (u64)env_offset + (u64)b0;
        u32 op = options;
        void *args[] = {&sp, &a_states, &a_actions, &nbv, &a_u, &sd, &stp, &e0, &op, &a_ns, &a_r, &a_p, &a_d, &a_c};
        LAUNCH(ctx->ks.step, grid_for(nb, ctx->threads, ctx->grid_step), ctx->threads, ctx->smem_base, st, args);
        CUDA_TRY(cudaMemcpyAsync((unsigned char *)next_states + sw * b0, d_ns + sw * b0, sw * nb, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync((unsigned char *)reward + 8 * b0, d_r + 8 * b0, 8 * nb, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync((unsigned char *)prob + 8 * b0, d_p + 8 * b0, 8 * nb, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(done + b0, d_d + b0, nb, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(collision + b0, d_c + b0, nb, cudaMemcpyDeviceToHost, st));
    }
    for (int h = 0; h < parts; ++h) CUDA_TRY(cudaStreamSynchronize(ctx->hs[h]));
    return MAPF_OK;
}
