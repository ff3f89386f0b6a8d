// radix.cuh -- three-kernel stable counting-sort pass (histogram -> scan -> scatter), one bzip2 block per blockIdx.y.
// Used by the inverse BWT (decode.cu): P = rows sorted stably by their byte, i.e. one pass with digit = L[row].
// (The forward BWT sorts with the one-kernel passes in sweep.cuh.)
// Tile = 2048 elements (256 threads x 8): stable in-tile ranking with warp match, staged through shared memory so
// that every digit run leaves the CTA as consecutive addresses.
#pragma once
#include "common.cuh"

namespace radix {

constexpr int R_IPT = 8;
constexpr int R_TILE = BZ_THREADS * R_IPT;   // 2048

struct RadixArgs {
    const u8 *T; const u32 *len;   // digit source (one byte per row) and block lengths
    u32 *sa_out;                   // [nblk][stride] sorted row indices
    u32 *thist;                    // [nblk][rtiles][256]
    u32 stride, rtiles;
};

__global__ void __launch_bounds__(BZ_THREADS, 4) k_radix_hist(RadixArgs a) {
    u32 b = blockIdx.y, t = blockIdx.x;
    u32 n = a.len[b];
    u32 base = t * R_TILE;
    if (base >= n) return;
    __shared__ u32 h[8][256];
    for (int i = threadIdx.x; i < 8 * 256; i += BZ_THREADS) (&h[0][0])[i] = 0;
    __syncthreads();
    const u8 *T = a.T + (size_t)b * a.stride;
    int w = threadIdx.x >> 5;
#pragma unroll
    for (int r = 0; r < R_IPT; r++) {
        u32 idx = base + r * BZ_THREADS + threadIdx.x;
        if (idx < n) atomicAdd(&h[w][T[idx]], 1u);
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

__global__ void __launch_bounds__(BZ_THREADS, 4) k_radix_scatter(RadixArgs a) {
    u32 b = blockIdx.y, t = blockIdx.x;
    u32 n = a.len[b];
    u32 base = t * R_TILE;
    if (base >= n) return;
    u32 tile_n = min((u32)R_TILE, n - base);

    __shared__ u32 wh[8 * 256];
    __shared__ u32 lbase[256];
    __shared__ u32 toff[256];
    __shared__ u32 ws[8];
    __shared__ u8 sdig[R_TILE];
    __shared__ u32 sval[R_TILE];

    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 8 * 256; i += BZ_THREADS) wh[i] = 0;
    toff[threadIdx.x] = a.thist[((size_t)b * a.rtiles + t) * 256 + threadIdx.x];
    __syncthreads();

    const u8 *T = a.T + (size_t)b * a.stride;
    int dig[R_IPT]; u32 val[R_IPT]; u32 rnk[R_IPT];
    // warp w owns elements [w*256, w*256+256) of the tile; round r covers 32 consecutive elements
#pragma unroll
    for (int r = 0; r < R_IPT; r++) {
        u32 e = w * (32 * R_IPT) + r * 32 + lane;
        dig[r] = 0x7fff; val[r] = base + e;
        if (e < tile_n) dig[r] = T[base + e];
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
        }
    }
    __syncthreads();
    size_t ob = (size_t)b * a.stride;
#pragma unroll
    for (int r = 0; r < R_IPT; r++) {
        u32 p = r * BZ_THREADS + threadIdx.x;
        if (p < tile_n) {
            int d = sdig[p];
            a.sa_out[ob + toff[d] + (p - lbase[d])] = sval[p];
        }
    }
}

}  // namespace radix
