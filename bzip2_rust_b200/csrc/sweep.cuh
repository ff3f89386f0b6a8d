// sweep.cuh -- one-kernel stable LSD radix pass per bzip2 block ("onesweep"), second generation.
//
// The BWT stage (bwt.cu) is a suffix sort by prefix doubling; every step of it is a stable 8-bit counting
// sort of a per-block array.  A pass is ONE kernel: a CTA ranks a tile of 256*IPT elements in shared memory
// (warp match + per-warp digit counters), obtains the sum over the preceding tiles of the same block by
// decoupled look-back, and writes every digit run to consecutive addresses.
//
// Initial sort of the rotation indices by their first 8 bytes (keys are implicit, only the 4-byte index moves):
//   M_GATHER  digit = T[(sa+off) mod n]; the same aligned 8-byte load also yields T[(sa+off-1) mod n], the digit
//             of the NEXT pass, which is written into the spare top byte of the output word (indices < 2^20).
//   M_CARRY   digit = top byte of the input word; no gather at all.  Output = plain index.
//   ncu on the first version (gather in every pass) showed the L1TEX tag stage as the limiter: a fully divergent
//   byte gather costs 32 tag look-ups per warp.  Carrying one byte halves the gathers.
// Unresolved-list sort:
//   M_LIST    element = one packed u64 (group head | key2 | rotation index), digit = 8 bits of it.
#pragma once
#include "common.cuh"
#include <stdlib.h>
#include <type_traits>

namespace sweep {

enum { M_GATHER = 0, M_LIST = 1, M_CARRY = 2 };

struct Args {
    const u8 *T; const u32 *len;   // text and block lengths
    const u32 *cnt;                // element count per block (initial sort: len, M_LIST: list count)
    const void *in; void *out;     // initial sort: u32 (in == nullptr => identity); M_LIST: u64
    const u32 *dbase;              // [nblk][dbase_stride] exclusive digit offsets of this pass
    u32 dbase_stride;
    u32 *tstate;                   // [nblk][rtiles][256] look-back words
    u32 *ticket;                   // one counter for this launch (zeroed before)
    u32 stride;                    // elements reserved per block in T / in / out
    u32 rtiles;                    // tiles reserved per block in tstate
    u32 epoch;                     // 1..255; a state word of another epoch is "not ready"
    u32 tiles_x;                   // tiles per block covered by the grid
    u32 nblk;
    u32 group;                     // blocks whose tiles are interleaved in ticket order
    int off;                       // M_GATHER: byte offset of this digit within the rotation (>= 1)
    int shift;                     // M_LIST: bit shift of this digit
};

__device__ __forceinline__ u32 ld_vol(const u32 *p) {
    u32 v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_vol(u32 *p, u32 v) { asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// ---- 1-D bulk copy global -> shared (TMA engine, `cp.async.bulk` + mbarrier; SASS: UBLKCP / SYNCS) ----
// A full tile of a streaming pass (M_CARRY, M_LIST) is one contiguous, 16-byte aligned span: instead of 16 LDG per
// thread, one thread posts a bulk copy into the staging buffer the pass needs anyway and the threads pick their
// elements up with LDS once the mbarrier's transaction count is complete.
#ifndef BZ_SWEEP_TMA
#define BZ_SWEEP_TMA 0
#endif
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_load(void *dst, const void *src, u32 bytes, u64 *bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity) {
    u32 ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}

template <int MODE> struct Types { typedef u32 In; typedef u32 Stage; typedef u32 Out; };
template <> struct Types<M_GATHER> { typedef u32 In; typedef u64 Stage; typedef u32 Out; };
template <> struct Types<M_LIST> { typedef u64 In; typedef u64 Stage; typedef u64 Out; };

template <int MODE, int IPT, bool FULL>
__device__ __forceinline__ void tile_body(const Args &a, u32 b, u32 t, u32 base, u32 tile_n, u32 n, u32 *wh, u32 *toff,
                                          u32 *ws, typename Types<MODE>::Stage *sbuf, u64 *bar) {
    typedef typename Types<MODE>::Stage S;
    typedef typename Types<MODE>::Out O;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const u32 lt = (1u << lane) - 1u;
    const u32 e0 = (u32)w * (32 * IPT) + lane;
    u32 *whw = wh + w * 256;
    // v: M_GATHER / M_CARRY: index | digit << 24;  M_LIST: the packed element.  cp: carried bytes, 4 per word.
    typename std::conditional<MODE == M_LIST, u64, u32>::type v[IPT];
    u32 cp[MODE == M_GATHER ? (IPT + 3) / 4 : 1];
    u32 rk[IPT];

    if (MODE == M_GATHER) {
        const u32 *in = a.in ? (const u32 *)a.in + (size_t)b * a.stride + base : nullptr;
        const u8 *Tb = a.T + (size_t)b * a.stride;
        const u32 offm1 = (u32)a.off - 1u;
#pragma unroll
        for (int r = 0; r < IPT; r++) {
            u32 e = e0 + r * 32;
            v[r] = 0;
            if (FULL || e < tile_n) v[r] = in ? __ldg(in + e) : base + e;
        }
#pragma unroll
        for (int q = 0; q < (IPT + 3) / 4; q++) cp[q] = 0;
#pragma unroll
        for (int r = 0; r < IPT; r++) {
            u32 q0 = (u32)v[r] + offm1;                              // position of the carried byte; the digit is at q0+1
            if (q0 >= n) { q0 -= n; while (q0 >= n) q0 -= n; }      // off <= 7: one subtraction unless the block is tiny
            // one aligned 8-byte gather serves both bytes (two byte gathers were measured 23% slower: the cost is per
            // divergent load instruction, not per byte)
            u64 word = __ldg((const u64 *)(Tb + (q0 & ~7u)));
            u32 sh = (q0 & 7u) * 8u;
            u32 two = (u32)(word >> sh);
            u32 c = two & 255u, d = (two >> 8) & 255u;
            if ((q0 & 7u) == 7u || q0 + 1 >= n) d = __ldg(Tb + (q0 + 1 >= n ? 0u : q0 + 1));
            v[r] = (u32)v[r] | (d << 24);
            cp[r >> 2] |= c << (8 * (r & 3));
        }
    } else if (BZ_SWEEP_TMA && FULL) {                          // the tile was posted as one bulk copy into sbuf (k_sweep)
        mbar_wait(bar, 0);
#pragma unroll
        for (int r = 0; r < IPT; r++) v[r] = sbuf[e0 + r * 32];
        // (the barrier before the scatter below also orders these reads before sbuf is overwritten)
    } else if (MODE == M_CARRY) {
        const u32 *in = (const u32 *)a.in + (size_t)b * a.stride + base;
#pragma unroll
        for (int r = 0; r < IPT; r++) {
            u32 e = e0 + r * 32;
            v[r] = 0;
            if (FULL || e < tile_n) v[r] = __ldg(in + e);
        }
    } else {
        const u64 *in = (const u64 *)a.in + (size_t)b * a.stride + base;
#pragma unroll
        for (int r = 0; r < IPT; r++) {
            u32 e = e0 + r * 32;
            v[r] = 0;
            if (FULL || e < tile_n) v[r] = __ldg(in + e);
        }
    }
#define SWEEP_DIGIT(x) (MODE == M_LIST ? ((u32)((u64)(x) >> a.shift) & 255u) : ((u32)(x) >> 24))
    // stable ranking inside the warp's segment: rk = number of earlier elements of the warp with the same digit
#pragma unroll
    for (int r = 0; r < IPT; r++) {
        bool valid = FULL || (e0 + r * 32 < tile_n);
        u32 d = SWEEP_DIGIT(v[r]);
        u32 peers = __match_any_sync(0xffffffffu, valid ? d : 256u);
        u32 before = __popc(peers & lt);
        u32 old = 0;
        if (before == 0 && valid) { old = whw[d]; whw[d] = old + __popc(peers); }
        old = __shfl_sync(0xffffffffu, old, __ffs(peers) - 1);
        rk[r] = old + before;
        __syncwarp();
    }
    __syncthreads();
    {   // digit = tid: tile count, look-back over the block's earlier tiles, exclusive bases folded into wh
        u32 c[8], total = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) { c[k] = wh[k * 256 + tid]; total += c[k]; }
        const u32 tag = a.epoch << 22;
        u32 *my = a.tstate + ((size_t)b * a.rtiles + t) * 256 + tid;
        if (t == 0) st_vol(my, (2u << 30) | tag | total);
        else st_vol(my, (1u << 30) | tag | total);
        u32 all;
        u32 ex = block_excl_sum(total, ws, all);
        u32 run = ex;
#pragma unroll
        for (int k = 0; k < 8; k++) { wh[k * 256 + tid] = run; run += c[k]; }
        u32 excl = 0;
        if (t != 0) {
            const u32 *p = my - 256;
            for (int tt = (int)t - 1; tt >= 0;) {
                u32 x = ld_vol(p);
                if (((x >> 22) & 255u) != a.epoch || (x >> 30) == 0) continue;      // predecessor not published yet
                excl += x & 0x3fffffu;
                if ((x >> 30) == 2u) break;
                tt--; p -= 256;
            }
            st_vol(my, (2u << 30) | tag | (excl + total));
        }
        toff[tid] = a.dbase[(size_t)b * a.dbase_stride + tid] + excl - ex;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < IPT; r++) {
        if (FULL || (e0 + r * 32 < tile_n)) {
            u32 d = SWEEP_DIGIT(v[r]);
            S x = (S)v[r];
            if (MODE == M_GATHER) x = (S)((u64)v[r] | ((u64)((cp[r >> 2] >> (8 * (r & 3))) & 255u) << 32));
            sbuf[whw[d] + rk[r]] = x;
        }
    }
    __syncthreads();
    O *out = (O *)a.out + (size_t)b * a.stride;
#pragma unroll
    for (int r = 0; r < IPT; r++) {
        u32 p = r * BZ_THREADS + tid;
        if (FULL || p < tile_n) {
            S x = sbuf[p];
            u32 d = SWEEP_DIGIT(x);
            O y;
            if (MODE == M_GATHER) y = (O)(((u32)x & 0xffffffu) | ((u32)((u64)x >> 32) << 24));
            else if (MODE == M_CARRY) y = (O)((u32)x & 0xffffffu);
            else y = (O)x;
            out[toff[d] + p] = y;
        }
    }
#undef SWEEP_DIGIT
}

#ifndef BZ_SWEEP_MINB_G
#define BZ_SWEEP_MINB_G 3
#endif
#ifndef BZ_SWEEP_MINB_L
#define BZ_SWEEP_MINB_L 3
#endif
constexpr int min_ctas(int MODE, int IPT) {
    return MODE == M_CARRY ? (IPT <= 8 ? 6 : 4) : (IPT <= 8 ? 4 : (MODE == M_GATHER ? BZ_SWEEP_MINB_G : BZ_SWEEP_MINB_L));
}

template <int MODE, int IPT>
__global__ void __launch_bounds__(BZ_THREADS, min_ctas(MODE, IPT)) k_sweep(Args a) {
    typedef typename Types<MODE>::Stage S;
    constexpr int TILE = BZ_THREADS * IPT;
    __shared__ u32 wh[8 * 256];
    __shared__ u32 toff[256];
    __shared__ u32 ws[8];
    __shared__ __align__(16) S sbuf[TILE];
    __shared__ __align__(8) u64 s_bar;
    __shared__ u32 s_ticket;
    const int tid = threadIdx.x;
    if (tid == 0) {
        s_ticket = atomicAdd(a.ticket, 1u);
        if (BZ_SWEEP_TMA && MODE != M_GATHER) mbar_init(&s_bar, 1);
    }
#pragma unroll
    for (int k = 0; k < 8; k++) wh[k * 256 + tid] = 0;
    __syncthreads();
    // Tickets walk the tiles of `group` blocks in lock step (tile 0 of each block, then tile 1, ...): only a few tiles
    // of any one block are in flight, so the look-back is short, and every tile a CTA waits for has already started.
    u32 ticket = s_ticket;
    u32 per_group = a.group * a.tiles_x;
    u32 g = ticket / per_group, r = ticket - g * per_group;
    u32 t = r / a.group;
    u32 b = g * a.group + (r - t * a.group);
    if (b >= a.nblk) return;
    u32 cnt = a.cnt[b];
    u32 base = t * TILE;
    if (base >= cnt) return;
    u32 n = a.len[b];
    u32 tile_n = min((u32)TILE, cnt - base);
    if (tile_n == (u32)TILE) {
        if (BZ_SWEEP_TMA && MODE != M_GATHER && tid == 0) {
            typedef typename Types<MODE>::In I;
            bulk_load(sbuf, (const I *)a.in + (size_t)b * a.stride + base, (u32)(TILE * sizeof(I)), &s_bar);
        }
        tile_body<MODE, IPT, true>(a, b, t, base, tile_n, n, wh, toff, ws, sbuf, &s_bar);
    } else tile_body<MODE, IPT, false>(a, b, t, base, tile_n, n, wh, toff, ws, sbuf, &s_bar);
}

// tuning knobs (environment, read once): elements per thread of the initial / list passes, interleave group
struct Knobs { int ipt0, ipt1; u32 group, group_stream; };
static inline const Knobs &knobs() {
    static Knobs k = [] {
        Knobs q;
        const char *e;
        q.ipt0 = (e = getenv("BZ2B200_SWEEP_IPT0")) ? atoi(e) : 16;
        q.ipt1 = (e = getenv("BZ2B200_SWEEP_IPT1")) ? atoi(e) : 16;
        // gather passes: few blocks in flight keep the gathered text inside the L2; passes that only stream prefer
        // short look-back chains (many blocks in flight).  Measured on text100m: 32 / 112.
        q.group = (e = getenv("BZ2B200_SWEEP_GROUP")) ? (u32)atoi(e) : 32u;
        q.group_stream = (e = getenv("BZ2B200_SWEEP_GROUP_STREAM")) ? (u32)atoi(e) : 128u;
        if (q.group_stream < 1) q.group_stream = 1;
        if (q.ipt0 != 8 && q.ipt0 != 16) q.ipt0 = 16;
        if (q.ipt1 != 8 && q.ipt1 != 16) q.ipt1 = 16;
        if (q.group < 1) q.group = 1;
        return q;
    }();
    return k;
}
static inline u32 tile_elems(int mode) { return (u32)(BZ_THREADS * (mode == M_LIST ? knobs().ipt1 : knobs().ipt0)); }

// Launches one pass.  `max_cnt` = upper bound of the element count of any block (sizes the grid).
template <int MODE>
static inline void launch(Args a, u32 max_cnt, cudaStream_t st) {
    const Knobs &k = knobs();
    int ipt = MODE == M_LIST ? k.ipt1 : k.ipt0;
    u32 tile = (u32)(BZ_THREADS * ipt);
    a.tiles_x = (max_cnt + tile - 1) / tile;
    a.rtiles = a.stride / tile;
    a.group = MODE == M_GATHER ? k.group : k.group_stream;
    if (a.group > a.nblk) a.group = a.nblk;
    u32 grid = ((a.nblk + a.group - 1) / a.group) * a.group * a.tiles_x;
    if (grid == 0) return;
    if (ipt == 8) k_sweep<MODE, 8><<<grid, BZ_THREADS, 0, st>>>(a);
    else k_sweep<MODE, 16><<<grid, BZ_THREADS, 0, st>>>(a);
}

// byte histogram of every block (digit totals of the initial sort: every byte is digit p of exactly one rotation,
// so all 8 passes share it): counts[b][256] += ...
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

// exclusive scan over the 256 digits of every histogram row (in place): row = blockIdx.x
static __global__ void __launch_bounds__(256) k_digit_scan(u32 *counts) {
    __shared__ u32 ws[8];
    u32 v = counts[(size_t)blockIdx.x * 256 + threadIdx.x], all;
    u32 ex = block_excl_sum(v, ws, all);
    counts[(size_t)blockIdx.x * 256 + threadIdx.x] = ex;
}

}  // namespace sweep
