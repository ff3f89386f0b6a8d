// mtf.cu -- move-to-front + RUNA/RUNB zero-run coding of the BWT output, batched.
//
// Replaces rle2_mtf_encode (reference src/tools/rle2_mtf.rs:23-177): initial MTF list = used
// bytes ascending (:26-39), MTF rank per byte (:61-62), zero runs as bijective base-2 RUNA/RUNB
// (:68-100), symbol = rank+1 (:106), EOB = nused+1 appended (:42,:166), the quirky freq table
// (:72,:79,:90,:104: RUNA->freq[0], RUNB->freq[1], rank p>=1 -> freq[p]).
//
// The MTF recurrence is made parallel by 1024-byte chunks:
//   k_used        used-byte bitmap of the block (256 bits).
//   k_mtf_summary one warp per chunk: last position of every byte value inside the chunk, and
//                 the chunk's zero-run bookkeeping.  rank 0 <=> L[i] == L[i-1] (i > 0), so the
//                 RLE2 structure is a local property of the BWT string.
//   k_mtf_scan    one CTA per block: running max of last positions over chunks (thread = byte
//                 value) gives, for every chunk start, each byte's last occurrence before it --
//                 the MTF list at the chunk start is the byte values sorted by that, descending.
//                 A short sequential pass turns the zero-run bookkeeping into per-chunk output
//                 offsets and pending-zero counts.
//   k_mtf_emit3   one warp per chunk: ranks the used values into the start list and replays the chunk
//                 32 positions per step on the inverse list (mtf_emit3.cuh), writing final u16 symbols
//                 straight at their output offsets.
#include "common.cuh"
#include <stdlib.h>

namespace {

#ifndef BZ_MTF_CH
#define BZ_MTF_CH 2048
#endif
constexpr int CH = BZ_MTF_CH;         // bytes per MTF chunk (one warp)
constexpr int WPB = BZ_THREADS / 32;  // warps (= chunks) per CTA

struct ChunkAgg {
    u32 lead;        // leading positions with rank 0 (== CH' when the whole chunk is zeros)
    u32 count_rest;  // symbols emitted for positions at/after the first non-zero rank, runs ending inside
    u32 trail;       // trailing zeros whose run continues into the next chunk
    u32 flags;       // bit0: all zero, bit1: an all-zero chunk's run ends at the chunk end
};

__device__ __forceinline__ bool is_nz(const u8 *L, u32 i, u32 min_used) {
    return i == 0 ? (L[0] != min_used) : (L[i] != L[i - 1]);
}

__global__ void __launch_bounds__(BZ_THREADS) k_used(const u8 *T, const u32 *len, u32 *usedbits, u32 stride) {
    u32 b = blockIdx.y, n = len[b];
    u32 base = blockIdx.x * BZ_TILE;
    if (base >= n) return;
    __shared__ u32 bits[8];
    if (threadIdx.x < 8) bits[threadIdx.x] = 0;
    __syncthreads();
    const u8 *t = T + (size_t)b * stride;
    u32 loc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 4
    for (int r = 0; r < BZ_IPT; r++) {
        u32 i = base + r * BZ_THREADS + threadIdx.x;
        if (i < n) { u32 c = t[i]; loc[c >> 5] |= 1u << (c & 31); }
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
        u32 v = loc[k];
        v = __reduce_or_sync(0xffffffffu, v);
        if ((threadIdx.x & 31) == 0 && v) atomicOr(&bits[k], v);
    }
    __syncthreads();
    if (threadIdx.x < 8 && bits[threadIdx.x]) atomicOr(&usedbits[b * 8 + threadIdx.x], bits[threadIdx.x]);
}

__device__ __forceinline__ u32 min_used_of(const u32 *ub) {
    for (int k = 0; k < 8; k++) if (ub[k]) return k * 32 + __ffs(ub[k]) - 1;
    return 0;
}

// one warp per chunk
__global__ void __launch_bounds__(BZ_THREADS) k_mtf_summary(const u8 *Lall, const u32 *len, const u32 *usedbits,
                                                            int *lp, ChunkAgg *agg, u32 stride, u32 nch_stride) {
    u32 b = blockIdx.y, n = len[b];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    u32 c = blockIdx.x * WPB + w;
    u32 a = c * CH;
    __shared__ int slast[WPB][256];
    if (a >= n) return;     // whole warp exits together; no block-level barrier below
    u32 e = min(a + CH, n);
    const u8 *L = Lall + (size_t)b * stride;
    u32 min_used = min_used_of(usedbits + b * 8);
    for (int k = lane; k < 256; k += 32) slast[w][k] = -1;
    __syncwarp();
    int first_nz = -1;      // position of the first non-zero rank in the chunk
    int last_nz = -1;       // running: last non-zero position seen so far (in chunk)
    u32 nnz = 0, digits = 0;
    // the bytes of a step are loaded two steps ahead (a load used in the step that issues it costs its full latency on
    // every one of the chunk's 64 steps); neighbours come from the lanes next door, the previous step's last byte and
    // the next step's first one -- 0x200 stands for "no byte" (before the block / after its end: never equal)
    auto ld = [&](u32 pos) { return pos < n ? (u32)L[pos] : 0x200u; };
    u32 c0 = ld(a + lane), c1 = ld(a + 32 + lane);
    u32 prev_last = a > 0 ? (u32)L[a - 1] : 0x200u;
    for (u32 i0 = a; i0 < e; i0 += 32) {
        u32 i = i0 + lane;
        bool in = i < e;
        bool nz = false, run_end = false;
        const u32 raw = c0;
        c0 = c1; c1 = ld(i0 + 64 + lane);
        const u32 chm = in ? raw : (0x100u | (u32)lane);
        {   // last occurrence of every byte value: positions grow with the lane, so the highest lane of each value wins
            const u32 peers = __match_any_sync(0xffffffffu, chm);
            if (in && (peers >> lane) == 1u) slast[w][chm] = (int)i;
        }
        {
            u32 prev = __shfl_up_sync(0xffffffffu, raw, 1), next = __shfl_down_sync(0xffffffffu, raw, 1);
            const u32 nfirst = __shfl_sync(0xffffffffu, c0, 0);
            if (lane == 0) prev = prev_last;
            if (lane == 31) next = nfirst;
            prev_last = __shfl_sync(0xffffffffu, raw, 31);
            if (in) {
                nz = i == 0 ? (raw != min_used) : (raw != prev);
                run_end = !nz && next != raw;
            }
        }
        unsigned mnz = __ballot_sync(0xffffffffu, nz);
        if (first_nz < 0 && mnz) first_nz = (int)(i0 + __ffs(mnz) - 1);
        // last non-zero position strictly before this lane's position
        unsigned below = mnz & ((1u << lane) - 1);
        int prev_nz = below ? (int)(i0 + 31 - __clz(below)) : last_nz;
        u32 d = 0;
        if (run_end && prev_nz >= 0) {          // a run that started inside the chunk, after first_nz
            u32 z = i - (u32)prev_nz;
            d = 31 - __clz(z + 1);
        }
        d = __reduce_add_sync(0xffffffffu, d);
        digits += d;
        nnz += __popc(mnz);
        if (mnz) last_nz = (int)(i0 + 31 - __clz(mnz));
    }
    __syncwarp();
    int *out = lp + ((size_t)b * nch_stride + c) * 256;
    for (int k = lane; k < 256; k += 32) out[k] = slast[w][k];
    if (lane == 0) {
        ChunkAgg g;
        bool next_nz = (e >= n) ? true : (L[e] != L[e - 1]);
        if (first_nz < 0) {
            g.lead = e - a; g.count_rest = 0; g.trail = 0; g.flags = 1u | (next_nz ? 2u : 0u);
        } else {
            g.lead = (u32)first_nz - a;
            g.count_rest = nnz + digits;
            // trailing zeros continue into the next chunk unless the run ends at e-1
            u32 tz = (e - 1) - (u32)last_nz;
            g.trail = (tz > 0 && !next_nz) ? tz : 0;
            g.flags = 0;
        }
        agg[(size_t)b * nch_stride + c] = g;
    }
}

// one CTA per block
__global__ void __launch_bounds__(256) k_mtf_scan(const u32 *len, const u32 *usedbits, const int *__restrict__ lp,
                                                  int *__restrict__ pm, const ChunkAgg *agg, u32 *zbefore,
                                                  u32 *ooff, u32 *m_out, u32 nch_stride) {
    u32 b = blockIdx.x, n = len[b];
    u32 nch = (n + CH - 1) / CH;
    u32 s = threadIdx.x;
    const u32 *ub = usedbits + b * 8;
    bool used = (ub[s >> 5] >> (s & 31)) & 1;
    u32 idx0 = 0;
    for (u32 k = 0; k < (s >> 5); k++) idx0 += __popc(ub[k]);
    idx0 += __popc(ub[s >> 5] & ((1u << (s & 31)) - 1));
    int run = used ? -(int)(idx0 + 1) : -100000 - (int)s;
    const int *src = lp + (size_t)b * nch_stride * 256;
    int *dst = pm + (size_t)b * nch_stride * 256;
#pragma unroll 16
    for (u32 c = 0; c < nch; c++) {
        int v = src[(size_t)c * 256 + s];
        dst[(size_t)c * 256 + s] = run;
        if (v >= 0) run = v;
    }
    // zero-run bookkeeping -> output offset and pending zeros of every chunk: a sequential pass, but over shared memory
    // (as one thread reading global memory it paid a DRAM round trip per chunk: 0.1 ms however small the batch)
    constexpr u32 TILE = 512;
    __shared__ ChunkAgg sagg[TILE];
    __shared__ u32 szb[TILE], ssum[TILE];
    __shared__ u32 c_zb, c_sum;
    if (threadIdx.x == 0) { c_zb = 0; c_sum = 0; }
    const ChunkAgg *g = agg + (size_t)b * nch_stride;
    for (u32 c0 = 0; c0 < nch; c0 += TILE) {
        const u32 cn = min(TILE, nch - c0);
        for (u32 i = threadIdx.x; i < cn; i += 256) sagg[i] = g[c0 + i];
        __syncthreads();
        if (threadIdx.x == 0) {
            u32 zb = c_zb, sum = c_sum;
            for (u32 i = 0; i < cn; i++) {
                ChunkAgg x = sagg[i];
                szb[i] = zb; ssum[i] = sum;
                if (x.flags & 1u) {
                    zb += x.lead;
                    if (x.flags & 2u) { sum += 31 - __clz(zb + 1); zb = 0; }
                } else {
                    u32 z = zb + x.lead;
                    if (z) sum += 31 - __clz(z + 1);
                    sum += x.count_rest;
                    zb = x.trail;
                }
            }
            c_zb = zb; c_sum = sum;
        }
        __syncthreads();
        for (u32 i = threadIdx.x; i < cn; i += 256) {
            zbefore[(size_t)b * nch_stride + c0 + i] = szb[i];
            ooff[(size_t)b * nch_stride + c0 + i] = ssum[i];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) m_out[b] = c_sum + 1;     // + EOB
}

#include "mtf_emit3.cuh"   // k_mtf_emit3: position-parallel replay, the version that is launched

}  // namespace

#define LAUNCH_OK()                                                  \
    do {                                                             \
        ctx->prof_end();                                             \
        cudaError_t e_ = cudaGetLastError();                         \
        if (e_ != cudaSuccess) { ctx->fail("kernel launch", e_, __FILE__, __LINE__); return BZ2B200_E_CUDA; } \
    } while (0)

// d_used receives the 256-bit used bitmap per block as 8 u32 words ([nblk*8] u32 = 32 bytes/block).
int bz_mtf_batch(bz2b200_ctx *ctx, const Batch &B, const u8 *d_bwt, u16 *d_sym, u32 *d_m, u32 *d_freq,
                 u8 *d_used, bool used_ready) {
    if (B.nblk == 0) return BZ2B200_OK;
    cudaStream_t st = ctx->stream;
    const u64 ne_act = B.total_n;
    u32 nch_stride = B.stride / CH;
    size_t nchunks = (size_t)B.nblk * nch_stride;
    BZ_CHECK(ctx->d_mtfstate.ensure(nchunks * 256 * 4 * 2));
    BZ_CHECK(ctx->d_chunkrec.ensure(nchunks * (sizeof(ChunkAgg) + 8)));
    int *lp = ctx->d_mtfstate.as<int>();
    int *pm = lp + nchunks * 256;
    ChunkAgg *agg = ctx->d_chunkrec.as<ChunkAgg>();
    u32 *zbefore = (u32 *)(agg + nchunks);
    u32 *ooff = zbefore + nchunks;
    u32 *usedbits = (u32 *)d_used;
    if (!used_ready) BZ_CHECK(cudaMemsetAsync(usedbits, 0, (size_t)B.nblk * 32, st));
    BZ_CHECK(cudaMemsetAsync(d_freq, 0, (size_t)B.nblk * 256 * 4, st));
    dim3 gfull((B.max_n + BZ_TILE - 1) / BZ_TILE, B.nblk);
    u32 maxch = (B.max_n + CH - 1) / CH;
    dim3 gch((maxch + WPB - 1) / WPB, B.nblk);
    if (!used_ready) { ctx->prof_begin(K_USED, ne_act); k_used<<<gfull, BZ_THREADS, 0, st>>>(d_bwt, B.len, usedbits, B.stride); LAUNCH_OK(); }
    ctx->prof_begin(K_MTF_SUMMARY, ne_act * 2); k_mtf_summary<<<gch, BZ_THREADS, 0, st>>>(d_bwt, B.len, usedbits, lp, agg, B.stride, nch_stride); LAUNCH_OK();
    ctx->prof_begin(K_MTF_SCAN, ne_act * 2); k_mtf_scan<<<B.nblk, 256, 0, st>>>(B.len, usedbits, lp, pm, agg, zbefore, ooff, d_m, nch_stride); LAUNCH_OK();
    ctx->prof_begin(K_MTF_EMIT, ne_act * 3);
    dim3 gem((maxch + EWPB - 1) / EWPB, B.nblk);
    k_mtf_emit3<<<gem, 32 * EWPB, 0, st>>>(d_bwt, B.len, usedbits, pm, zbefore, ooff, d_m, d_sym, d_freq, B.stride, nch_stride);
    LAUNCH_OK();
    return BZ2B200_OK;
}
