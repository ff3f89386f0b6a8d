// radix.cuh -- per-block (segmented) LSD radix sort passes used by the BWT stage.
//
// MODE 0: initial sort of rotation indices; the digit of a pass is gathered from the block text
//         (T[(sa+off) mod n]), so only the 4-byte index moves.
// MODE 1: unresolved-list sort; element = (key64, val32), digit = 8 bits of the key.
// A pass = k_radix_hist (per-tile digit counts) -> k_radix_scan (per block: counts -> scatter offsets)
// -> k_radix_scatter (stable in-tile ranking with warp match, staged through shared memory so that
// every digit run leaves the CTA as consecutive addresses).
// Tile = 2048 elements (256 threads x 8): small enough for 4 CTAs per SM (the first version used 4096
// elements and 150 registers, i.e. one CTA per SM; ncu showed 12% warp occupancy).
#pragma once
#include "common.cuh"
#include <stdlib.h>

namespace radix {

constexpr int R_IPT = 8;
constexpr int R_TILE = BZ_THREADS * R_IPT;   // 2048

struct RadixArgs {
    const u8 *T; const u32 *len;   // text and block lengths
    const u32 *cnt;                // element count per block (MODE 0: len, MODE 1: list count)
    const u32 *sa_in; u32 *sa_out; // MODE 0 (sa_in == nullptr => identity)
    const u64 *key_in; u64 *key_out; const u32 *val_in; u32 *val_out;   // MODE 1
    u32 *thist;                    // [nblk][rtiles][256]
    u32 stride, rtiles;
    int off;                       // MODE 0: byte offset of this digit within the rotation
    int shift;                     // MODE 1: bit shift of this digit
};

template <int MODE>
__device__ __forceinline__ int radix_digit(const RadixArgs &a, u32 b, u32 n, u32 idx, u32 &sa, u64 &key) {
    if (MODE == 0) {
        sa = a.sa_in ? a.sa_in[(size_t)b * a.stride + idx] : idx;
        u32 p = sa + (u32)a.off;
        if (p >= n) { p -= n; if (p >= n) p %= n; }    // off < 8: one subtraction except for blocks shorter than 8
        return a.T[(size_t)b * a.stride + p];
    } else {
        key = a.key_in[(size_t)b * a.stride + idx];
        return (int)((key >> a.shift) & 255);
    }
}

template <int MODE>
__global__ void __launch_bounds__(BZ_THREADS, 4) k_radix_hist(RadixArgs a) {
    u32 b = blockIdx.y, t = blockIdx.x;
    u32 cnt = a.cnt[b];
    u32 base = t * R_TILE;
    if (base >= cnt) return;
    u32 n = a.len[b];
    __shared__ u32 h[8][256];
    for (int i = threadIdx.x; i < 8 * 256; i += BZ_THREADS) (&h[0][0])[i] = 0;
    __syncthreads();
    int w = threadIdx.x >> 5;
#pragma unroll
    for (int r = 0; r < R_IPT; r++) {
        u32 idx = base + r * BZ_THREADS + threadIdx.x;
        if (idx < cnt) {
            u32 sa; u64 key;
            int d = radix_digit<MODE>(a, b, n, idx, sa, key);
            atomicAdd(&h[w][d], 1u);
        }
    }
    __syncthreads();
    u32 s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += h[k][threadIdx.x];
    a.thist[((size_t)b * a.rtiles + t) * 256 + threadIdx.x] = s;
}

// per block: turn per-tile digit counts into global scatter offsets (in place)
static __global__ void __launch_bounds__(256) k_radix_scan(u32 *thist, const u32 *cntp, u32 rtiles_stride) {
    u32 b = blockIdx.x, d = threadIdx.x;
    u32 cnt = cntp[b];
    u32 tiles = (cnt + R_TILE - 1) / R_TILE;
    __shared__ u32 ws[8];
    u32 *h = thist + (size_t)b * rtiles_stride * 256;
    u32 total = 0;
#pragma unroll 8
    for (u32 t = 0; t < tiles; t++) total += h[(size_t)t * 256 + d];
    u32 all;
    u32 run = block_excl_sum(total, ws, all);
#pragma unroll 8
    for (u32 t = 0; t < tiles; t++) {
        u32 v = h[(size_t)t * 256 + d];
        h[(size_t)t * 256 + d] = run;
        run += v;
    }
}

template <int MODE>
__global__ void __launch_bounds__(BZ_THREADS, 4) k_radix_scatter(RadixArgs a) {
    u32 b = blockIdx.y, t = blockIdx.x;
    u32 cnt = a.cnt[b];
    u32 base = t * R_TILE;
    if (base >= cnt) return;
    u32 n = a.len[b];
    u32 tile_n = min((u32)R_TILE, cnt - base);

    __shared__ u32 wh[8 * 256];
    __shared__ u32 lbase[256];
    __shared__ u32 toff[256];
    __shared__ u32 ws[8];
    __shared__ u8 sdig[R_TILE];
    __shared__ u32 sval[R_TILE];
    __shared__ u64 skey[MODE == 1 ? R_TILE : 1];

    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 8 * 256; i += BZ_THREADS) wh[i] = 0;
    toff[threadIdx.x] = a.thist[((size_t)b * a.rtiles + t) * 256 + threadIdx.x];
    __syncthreads();

    int dig[R_IPT]; u32 val[R_IPT]; u64 key[R_IPT]; u32 rnk[R_IPT];
    // warp w owns elements [w*256, w*256+256) of the tile; round r covers 32 consecutive elements
#pragma unroll
    for (int r = 0; r < R_IPT; r++) {           // loads first: all independent
        u32 e = w * (32 * R_IPT) + r * 32 + lane;
        u32 idx = base + e;
        dig[r] = 0x7fff; val[r] = 0; key[r] = 0;
        if (e < tile_n) {
            u32 sa = 0; u64 k = 0;
            dig[r] = radix_digit<MODE>(a, b, n, idx, sa, k);
            if (MODE == 0) val[r] = sa; else { key[r] = k; val[r] = a.val_in[(size_t)b * a.stride + idx]; }
        }
    }
#pragma unroll
    for (int r = 0; r < R_IPT; r++) {           // stable ranking inside the warp's segment
        int d = dig[r];
        bool valid = d != 0x7fff;
        unsigned peers = __match_any_sync(0xffffffffu, d);
        int leader = __ffs(peers) - 1;
        u32 old = 0;
        if (lane == leader && valid) { old = wh[w * 256 + d]; wh[w * 256 + d] = old + __popc(peers); }
        old = __shfl_sync(0xffffffffu, old, leader);
        rnk[r] = old + __popc(peers & ((1u << lane) - 1));
        __syncwarp();
    }
    __syncthreads();
    {   // exclusive scan over warps for digit = threadIdx.x, then over digits
        u32 run = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) { u32 v = wh[k * 256 + threadIdx.x]; wh[k * 256 + threadIdx.x] = run; run += v; }
        u32 all;
        u32 ex = block_excl_sum(run, ws, all);
        lbase[threadIdx.x] = ex;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < R_IPT; r++) {
        int d = dig[r];
        if (d != 0x7fff) {
            u32 pos = lbase[d] + wh[w * 256 + d] + rnk[r];
            sdig[pos] = (u8)d;
            sval[pos] = val[r];
            if (MODE == 1) skey[pos] = key[r];
        }
    }
    __syncthreads();
    size_t ob = (size_t)b * a.stride;
#pragma unroll
    for (int r = 0; r < R_IPT; r++) {
        u32 p = r * BZ_THREADS + threadIdx.x;
        if (p < tile_n) {
            int d = sdig[p];
            u32 dst = toff[d] + (p - lbase[d]);
            if (MODE == 0) a.sa_out[ob + dst] = sval[p];
            else { a.key_out[ob + dst] = skey[p]; a.val_out[ob + dst] = sval[p]; }
        }
    }
}

// --------------------------------------------------------------------------------------------------------
// One-kernel pass ("onesweep"): the scatter kernel computes its own tile histogram (it needs it for the
// in-tile ranking anyway) and obtains the sum over the preceding tiles of the same bzip2 block by decoupled
// look-back, so the separate k_radix_hist / k_radix_scan launches disappear.  Needs, per block and pass, the
// exclusive digit offsets `dbase` (from the block's total digit histogram: for MODE 0 that is the block's byte
// histogram and is the same for all 8 passes; for MODE 1 one kernel reads the keys once for all passes).
// Tile state word: [31:30] flag (1 = tile aggregate, 2 = inclusive prefix) [29:22] epoch [21:0] count.
// A word whose epoch is not the current pass is "not ready", so the array is cleared once per batch, not per pass.
// Tiles are handed out by an atomic ticket so that every tile a CTA waits for has already started.
// --------------------------------------------------------------------------------------------------------
struct SweepArgs {
    RadixArgs r;
    const u32 *dbase;      // [nblk][dbase_stride] exclusive digit offsets of this pass (first 256 words of each row)
    u32 dbase_stride;
    u32 *tstate;           // [nblk][rtiles][256]
    u32 *ticket;           // one counter for this launch (zeroed before)
    u32 epoch;             // 1..255
    u32 tiles_x;           // tiles per block covered by the grid
    u32 nblk;
    u32 group;             // blocks interleaved per ticket group
};
// blocks whose tiles are interleaved (tuning knob: env BZ2B200_SWEEP_GROUP, default 32)
static inline u32 sweep_group() {
    static u32 g = 0;
    if (!g) { const char *e = getenv("BZ2B200_SWEEP_GROUP"); g = e ? (u32)atoi(e) : 32u; if (g < 1) g = 1; }
    return g;
}
// grid size for a sweep launch: whole groups of blocks
static inline u32 sweep_grid(u32 tiles_x, u32 nblk) { u32 G = sweep_group(); return ((nblk + G - 1) / G) * G * tiles_x; }

__device__ __forceinline__ u32 ld_volatile_u32(const u32 *p) {
    u32 v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_volatile_u32(u32 *p, u32 v) {
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int MODE>
// 6 CTAs/SM (40 registers) for the 4-byte payload, 4 (64 registers) for the 12-byte one; 8/5 spill and are slower
__global__ void __launch_bounds__(BZ_THREADS, MODE == 0 ? 6 : 4) k_radix_onesweep(SweepArgs s) {
    const RadixArgs &a = s.r;
    __shared__ u32 s_ticket;
    if (threadIdx.x == 0) s_ticket = atomicAdd(s.ticket, 1u);
    __syncthreads();
    u32 ticket = s_ticket;
    // Tickets walk the tiles of SWEEP_GROUP blocks in lock step (tile 0 of each block, then tile 1, ...): only a few
    // tiles of any one block are in flight, so the look-back is short, while the text of only SWEEP_GROUP blocks
    // (tens of MB, L2 resident) is being gathered from at a time.
    const u32 SWEEP_GROUP = s.group;
    u32 per_group = SWEEP_GROUP * s.tiles_x;
    u32 g = ticket / per_group, r = ticket % per_group;
    u32 b = g * SWEEP_GROUP + r % SWEEP_GROUP, t = r / SWEEP_GROUP;
    if (b >= s.nblk) return;
    u32 cnt = a.cnt[b];
    u32 base = t * R_TILE;
    if (base >= cnt) return;
    u32 n = a.len[b];
    u32 tile_n = min((u32)R_TILE, cnt - base);

    __shared__ u32 wh[8 * 256];
    __shared__ u32 lbase[256];
    __shared__ u32 toff[256];
    __shared__ u32 ws[8];
    __shared__ u8 sdig[R_TILE];
    __shared__ u32 sval[R_TILE];
    __shared__ u64 skey[MODE == 1 ? R_TILE : 1];

    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 8 * 256; i += BZ_THREADS) wh[i] = 0;
    __syncthreads();

    int dig[R_IPT]; u32 val[R_IPT]; u64 key[R_IPT]; u32 rnk[R_IPT];
#pragma unroll
    for (int r = 0; r < R_IPT; r++) {
        u32 e = w * (32 * R_IPT) + r * 32 + lane;
        u32 idx = base + e;
        dig[r] = 0x7fff; val[r] = 0; key[r] = 0;
        if (e < tile_n) {
            u32 sa = 0; u64 k = 0;
            dig[r] = radix_digit<MODE>(a, b, n, idx, sa, k);
            if (MODE == 0) val[r] = sa; else { key[r] = k; val[r] = a.val_in[(size_t)b * a.stride + idx]; }
        }
    }
#pragma unroll
    for (int r = 0; r < R_IPT; r++) {
        int d = dig[r];
        bool valid = d != 0x7fff;
        unsigned peers = __match_any_sync(0xffffffffu, d);
        int leader = __ffs(peers) - 1;
        u32 old = 0;
        if (lane == leader && valid) { old = wh[w * 256 + d]; wh[w * 256 + d] = old + __popc(peers); }
        old = __shfl_sync(0xffffffffu, old, leader);
        rnk[r] = old + __popc(peers & ((1u << lane) - 1));
        __syncwarp();
    }
    __syncthreads();
    {
        // digit = threadIdx.x: tile count, exclusive scan over warps and digits, then the look-back
        u32 run = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) { u32 v = wh[k * 256 + threadIdx.x]; wh[k * 256 + threadIdx.x] = run; run += v; }
        const u32 tag = s.epoch << 22;
        u32 *my = s.tstate + ((size_t)b * a.rtiles + t) * 256 + threadIdx.x;
        u32 excl = 0;
        if (t == 0) {
            st_volatile_u32(my, (2u << 30) | tag | run);
        } else {
            st_volatile_u32(my, (1u << 30) | tag | run);
            const u32 *p = my - 256;
            for (int tt = (int)t - 1; tt >= 0;) {
                u32 v = ld_volatile_u32(p);
                if (((v >> 22) & 255u) != s.epoch || (v >> 30) == 0) continue;      // predecessor not published yet
                excl += v & 0x3fffffu;
                if ((v >> 30) == 2u) break;
                tt--; p -= 256;
            }
            st_volatile_u32(my, (2u << 30) | tag | (excl + run));
        }
        toff[threadIdx.x] = s.dbase[(size_t)b * s.dbase_stride + threadIdx.x] + excl;
        u32 all;
        u32 ex = block_excl_sum(run, ws, all);
        lbase[threadIdx.x] = ex;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < R_IPT; r++) {
        int d = dig[r];
        if (d != 0x7fff) {
            u32 pos = lbase[d] + wh[w * 256 + d] + rnk[r];
            sdig[pos] = (u8)d;
            sval[pos] = val[r];
            if (MODE == 1) skey[pos] = key[r];
        }
    }
    __syncthreads();
    size_t ob = (size_t)b * a.stride;
#pragma unroll
    for (int r = 0; r < R_IPT; r++) {
        u32 p = r * BZ_THREADS + threadIdx.x;
        if (p < tile_n) {
            int d = sdig[p];
            u32 dst = toff[d] + (p - lbase[d]);
            if (MODE == 0) a.sa_out[ob + dst] = sval[p];
            else { a.key_out[ob + dst] = skey[p]; a.val_out[ob + dst] = sval[p]; }
        }
    }
}

// byte histogram of every block (MODE 0 digit totals): counts[b][256] += ...
static __global__ void __launch_bounds__(BZ_THREADS) k_byte_hist(const u8 *T, const u32 *len, u32 *counts, u32 stride,
                                                                  u32 cstride) {
    u32 b = blockIdx.y, n = len[b];
    u32 base = blockIdx.x * BZ_TILE;
    if (base >= n) return;
    __shared__ u32 h[8][256];
    for (int i = threadIdx.x; i < 8 * 256; i += BZ_THREADS) (&h[0][0])[i] = 0;
    __syncthreads();
    const u8 *t = T + (size_t)b * stride;
    int w = threadIdx.x >> 5;
#pragma unroll 4
    for (int r = 0; r < BZ_IPT; r++) {
        u32 i = base + r * BZ_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[w][t[i]], 1u);
    }
    __syncthreads();
    u32 s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += h[k][threadIdx.x];
    if (s) atomicAdd(&counts[(size_t)b * cstride + threadIdx.x], s);
}

// digit histograms of all `npass` passes of a list sort in one read of the keys: counts[b][p][256]
static __global__ void __launch_bounds__(BZ_THREADS) k_list_hist(const u64 *KEY, const u32 *cntp, u32 *counts, u32 stride,
                                                                  int npass) {
    u32 b = blockIdx.y, cnt = cntp[b];
    u32 base = blockIdx.x * BZ_TILE;
    if (base >= cnt) return;
    __shared__ u32 h[8][256];                   // up to 8 passes
    for (int i = threadIdx.x; i < 8 * 256; i += BZ_THREADS) (&h[0][0])[i] = 0;
    __syncthreads();
    const u64 *key = KEY + (size_t)b * stride;
#pragma unroll 4
    for (int r = 0; r < BZ_IPT; r++) {
        u32 i = base + r * BZ_THREADS + threadIdx.x;
        if (i < cnt) {
            u64 k = key[i];
            for (int p = 0; p < npass; p++) atomicAdd(&h[p][(k >> (8 * p)) & 255], 1u);
        }
    }
    __syncthreads();
    for (int p = 0; p < npass; p++) {
        u32 v = h[p][threadIdx.x];
        if (v) atomicAdd(&counts[((size_t)b * 8 + p) * 256 + threadIdx.x], v);
    }
}

// exclusive scan over the 256 digits of every histogram row (in place): row = blockIdx.x
static __global__ void __launch_bounds__(256) k_digit_scan(u32 *counts) {
    __shared__ u32 ws[8];
    u32 v = counts[(size_t)blockIdx.x * 256 + threadIdx.x], all;
    u32 ex = block_excl_sum(v, ws, all);
    counts[(size_t)blockIdx.x * 256 + threadIdx.x] = ex;
}

}  // namespace radix
