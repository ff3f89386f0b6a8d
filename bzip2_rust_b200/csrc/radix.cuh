// radix.cuh -- three-kernel stable counting-sort pass (histogram -> scan -> scatter), one bzip2 block per blockIdx.y.
// The inverse BWT (decode.cu) uses one MODE 0 pass; the forward BWT uses the one-kernel passes in sweep.cuh.
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

}  // namespace radix
