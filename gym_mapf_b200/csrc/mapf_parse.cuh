// mapf_parse.cuh -- the on-disk MovingAI `.map` format read on the device.
//
// Reference: parse_map_file (utils.py:33-37) returns `f.readlines()[4:]`, MapfGrid.__init__ (grid.py:17-25) strips
// every one of those lines and maps its characters through CHAR_TO_CELL ('.' free, '@' obstacle, anything else
// raises KeyError).  k_parse_map does the same on the raw file bytes: line starts, per-line strip(), character
// classification, row-major obstacle bytes.  Context creation only; one CTA (the largest shipped map is 66 KB).
#pragma once
#include "mapf_device.cuh"

struct ParsedMapHeader {
    int n_lines;     // lines of the file (a last line without a terminator counts)
    int H, W;        // grid rows = n_lines - 4, columns = stripped length of the first grid row
    int ragged;      // 1 + index of the first grid row whose stripped length differs from W (0: none)
    u32 bad_pos;     // byte offset of the first character (in reading order) that is neither '.' nor '@'; ~0u: none
    u32 bad_char;    // that character
};

#define PARSE_THREADS 1024

// Python's str.strip() without arguments, restricted to one-byte characters
__device__ __forceinline__ bool py_space(unsigned char ch) { return (ch >= 9 && ch <= 13) || (ch >= 28 && ch <= 32); }

// Text-mode universal newlines: "\n", "\r\n" and a lone "\r" all end a line.  A line ends AT position p when:
__device__ __forceinline__ bool line_ends_at(const unsigned char *t, u32 len, u32 p) {
    return t[p] == '\n' || (t[p] == '\r' && (p + 1 >= len || t[p + 1] != '\n'));
}

// line_start: u32[len + 2] scratch (start offset of every line, then the end sentinel)
// row_begin/row_len: u32[len + 1] scratch (stripped extent of every grid row)
// obstacles: u8[len] (at most one cell per input byte), row-major H x W, 1 = '@'
static __global__ void __launch_bounds__(PARSE_THREADS)
k_parse_map(const unsigned char *__restrict__ text, u32 len, u32 *__restrict__ line_start, u32 *__restrict__ row_begin,
            u32 *__restrict__ row_len, u8 *__restrict__ obstacles, ParsedMapHeader *__restrict__ hdr) {
    __shared__ u32 s_warp[PARSE_THREADS / 32];
    __shared__ u32 s_total;
    __shared__ int s_W, s_rag;
    __shared__ u32 s_bad;
    const u32 tid = threadIdx.x;
    const u32 chunk = (len + PARSE_THREADS - 1) / PARSE_THREADS;
    const u32 lo = min(len, tid * chunk), hi = min(len, lo + chunk);
    // ---- pass 1: line terminators per thread chunk, exclusive scan over the CTA
    u32 cnt = 0;
    for (u32 p = lo; p < hi; ++p) cnt += line_ends_at(text, len, p) ? 1u : 0u;
    u32 incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((tid & 31) >= (u32)o) incl += v;
    }
    if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32) {
        u32 w = tid < PARSE_THREADS / 32 ? s_warp[tid] : 0u, wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 v = __shfl_up_sync(0xffffffffu, wi, o);
            if (tid >= (u32)o) wi += v;
        }
        if (tid < PARSE_THREADS / 32) s_warp[tid] = wi - w;
        if (tid == 31) s_total = wi;
    }
    __syncthreads();
    u32 idx = s_warp[tid >> 5] + incl - cnt;  // terminators before this chunk
    // ---- pass 2: line k + 1 starts behind the k-th terminator
    if (tid == 0) line_start[0] = 0;
    for (u32 p = lo; p < hi; ++p)
        if (line_ends_at(text, len, p)) line_start[++idx] = p + 1;
    __syncthreads();
    const u32 n_term = s_total;
    // readlines(): a trailing piece without terminator is a line of its own
    const bool tail = len > 0 && !line_ends_at(text, len, len - 1);
    const int n_lines = (int)n_term + (tail ? 1 : 0);
    if (tid == 0 && tail) line_start[n_term + 1] = len;
    __syncthreads();
    // ---- grid rows: lines[4:], each strip()ped (grid.py:20)
    const int H = n_lines - 4;
    for (int r = (int)tid; r < H; r += PARSE_THREADS) {
        u32 b = line_start[r + 4], e = line_start[r + 5];
        while (b < e && py_space(text[b])) ++b;
        while (e > b && py_space(text[e - 1])) --e;
        row_begin[r] = b;
        row_len[r] = e - b;
    }
    __syncthreads();
    if (tid == 0) {
        s_W = H >= 1 ? (int)row_len[0] : 0;  // max_col = len(self._map[0]) - 1 (grid.py:25)
        s_bad = 0xffffffffu;
        s_rag = 0x7fffffff;
    }
    __syncthreads();
    const int W = s_W;
    // ---- characters -> cells (CHAR_TO_CELL, grid.py:9-13,21); the first offender in reading order is reported
    u32 my_bad = 0xffffffffu;
    int my_ragged = 0x7fffffff;
    for (int r = (int)tid; r < H; r += PARSE_THREADS)
        if ((int)row_len[r] != W) my_ragged = min(my_ragged, r);
    for (int r = 0; r < H; ++r) {
        const u32 b = row_begin[r], n = row_len[r];
        for (u32 c = tid; c < n; c += PARSE_THREADS) {
            const unsigned char ch = text[b + c];
            if (ch != '.' && ch != '@') my_bad = min(my_bad, b + c);
            if ((int)n == W) obstacles[(size_t)r * W + c] = ch == '@' ? 1 : 0;
        }
    }
    if (my_bad != 0xffffffffu) atomicMin(&s_bad, my_bad);
    if (my_ragged != 0x7fffffff) atomicMin(&s_rag, my_ragged);
    __syncthreads();
    if (tid == 0) {
        hdr->n_lines = n_lines;
        hdr->H = H;
        hdr->W = W;
        hdr->ragged = s_rag == 0x7fffffff ? 0 : s_rag + 1;
        hdr->bad_pos = s_bad;
        hdr->bad_char = s_bad != 0xffffffffu ? (u32)text[s_bad] : 0u;
    }
}
