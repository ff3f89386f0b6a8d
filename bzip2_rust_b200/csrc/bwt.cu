// bwt.cu -- cyclic-rotation BWT of a batch of RLE1 blocks as a GPU suffix sort.
//
// Replaces bwt_encode (reference src/bwt_algorithms/bwt_sort.rs:27-58): sort all n cyclic
// rotations of the block, key = row of rotation 0, bwt[i] = x[(idx[i]-1) mod n].  The reference
// does this with a comparison sort of full rotations (block_compare, bwt_sort.rs:61-86); here it
// is prefix doubling over packed rank keys:
//   1. LSD radix sort of all rotations by their first 8 bytes.  Keys are implicit: the digit of
//      pass p is T[(sa+7-p) mod n], gathered from the (L2 resident) block text, so a pass moves
//      only the 4-byte rotation index.
//   2. head flags -> rank[s] = first row of the group of rotations sharing the 8-byte prefix;
//      rotations in groups of size > 1 are compacted into an "unresolved" list.
//   3. doubling rounds h = 8,16,..: key2 = rank[(s+h) mod n]; the list is radix sorted by the
//      packed key (group head << nbits | key2), groups are split where key2 changes, ranks of
//      the refined groups are rewritten and the still unresolved rotations are re-compacted.
//      Stops when every list is empty or h >= n (rotations still tied are equal: periodic block).
//   4. key = rank[0] (first row of the class of rotation 0), bwt[j] = T[(SA[j]-1) mod n].
// One bzip2 block per blockIdx.y, tiles of 4096 elements along blockIdx.x.
#include "common.cuh"
#include "radix.cuh"

namespace {

struct BwtWs {
    u32 *SA, *SA2, *RANK;
    u8 *F;
    u64 *KEYA, *KEYB;
    u32 *VALA, *VALB;
    u32 *thist;     // [nblk][tiles][256]
    int4 *tagg;     // [nblk][tiles] tile aggregates {maxA, maxB, count, 0}
    u32 *cnt;       // [nblk] current list length; [nblk..2nblk) next list length
};

// --------------------------------------------------------------------------------------
// radix sort passes (hist -> scan -> scatter).  MODE 0: initial sort, element = rotation index,
// digit gathered from the text.  MODE 1: list sort, element = (key64, val32), digit from key.
// Now in radix.cuh (2048-element tiles, 4 CTAs/SM); the first version below is kept for reference only.
// --------------------------------------------------------------------------------------
#if 0
struct RadixArgs {
    const u8 *T; const u32 *len;   // text and block lengths
    const u32 *cnt;                // element count per block (MODE 0: len, MODE 1: list count)
    const u32 *sa_in; u32 *sa_out; // MODE 0 (sa_in == nullptr => identity)
    const u64 *key_in; u64 *key_out; const u32 *val_in; u32 *val_out;   // MODE 1
    u32 *thist;
    u32 stride, tiles;
    int off;                       // MODE 0: byte offset of this digit within the rotation
    int shift;                     // MODE 1: bit shift of this digit
};

template <int MODE>
__device__ __forceinline__ int radix_digit(const RadixArgs &a, u32 b, u32 n, u32 idx, u32 &sa, u64 &key) {
    if (MODE == 0) {
        sa = a.sa_in ? a.sa_in[(size_t)b * a.stride + idx] : idx;
        u32 p = sa + (u32)a.off;
        if (p >= n) p %= n;
        return a.T[(size_t)b * a.stride + p];
    } else {
        key = a.key_in[(size_t)b * a.stride + idx];
        return (int)((key >> a.shift) & 255);
    }
}

template <int MODE>
__global__ void __launch_bounds__(BZ_THREADS) k_radix_hist(RadixArgs a) {
    u32 b = blockIdx.y, t = blockIdx.x;
    u32 cnt = a.cnt[b];
    u32 base = t * BZ_TILE;
    if (base >= cnt) return;
    u32 n = a.len[b];
    __shared__ u32 h[8][256];
    for (int i = threadIdx.x; i < 8 * 256; i += BZ_THREADS) (&h[0][0])[i] = 0;
    __syncthreads();
    int w = threadIdx.x >> 5;
#pragma unroll 4
    for (int r = 0; r < BZ_IPT; r++) {
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
    a.thist[((size_t)b * a.tiles + t) * 256 + threadIdx.x] = s;
}

// per block: turn per-tile digit counts into global scatter offsets (in place)
__global__ void __launch_bounds__(256) k_radix_scan(u32 *thist, const u32 *cntp, u32 tiles_stride) {
    u32 b = blockIdx.x, d = threadIdx.x;
    u32 cnt = cntp[b];
    u32 tiles = (cnt + BZ_TILE - 1) / BZ_TILE;
    __shared__ u32 ws[8];
    u32 *h = thist + (size_t)b * tiles_stride * 256;
    u32 total = 0;
    for (u32 t = 0; t < tiles; t++) total += h[(size_t)t * 256 + d];
    u32 all;
    u32 run = block_excl_sum(total, ws, all);
    for (u32 t = 0; t < tiles; t++) {
        u32 v = h[(size_t)t * 256 + d];
        h[(size_t)t * 256 + d] = run;
        run += v;
    }
}

template <int MODE>
__global__ void __launch_bounds__(BZ_THREADS) k_radix_scatter(RadixArgs a) {
    u32 b = blockIdx.y, t = blockIdx.x;
    u32 cnt = a.cnt[b];
    u32 base = t * BZ_TILE;
    if (base >= cnt) return;
    u32 n = a.len[b];
    u32 tile_n = min((u32)BZ_TILE, cnt - base);

    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: wh[8][256] u32 | lbase[256] u32 | toff[256] u32 | ws[8] u32 | sdig[TILE] u8 | sval[TILE] u32 | skey[TILE] u64 (MODE 1)
    u32 *wh = (u32 *)smem_raw;
    u32 *lbase = wh + 8 * 256;
    u32 *toff = lbase + 256;
    u32 *ws = toff + 256;
    u8 *sdig = (u8 *)(ws + 8);
    u32 *sval = (u32 *)(sdig + BZ_TILE);
    u64 *skey = (u64 *)(sval + BZ_TILE);

    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 8 * 256; i += BZ_THREADS) wh[i] = 0;
    toff[threadIdx.x] = a.thist[((size_t)b * a.tiles + t) * 256 + threadIdx.x];
    __syncthreads();

    int dig[BZ_IPT]; u32 val[BZ_IPT]; u64 key[BZ_IPT]; u32 rnk[BZ_IPT];
    // warp w owns elements [w*512, w*512+512) of the tile; round r covers 32 consecutive elements
#pragma unroll
    for (int r = 0; r < BZ_IPT; r++) {
        u32 e = w * (32 * BZ_IPT) + r * 32 + lane;
        u32 idx = base + e;
        bool valid = e < tile_n;
        int d = 0x7fff;
        val[r] = 0; key[r] = 0;
        if (valid) {
            u32 sa = 0; u64 k = 0;
            d = radix_digit<MODE>(a, b, n, idx, sa, k);
            if (MODE == 0) val[r] = sa; else { key[r] = k; val[r] = a.val_in[(size_t)b * a.stride + idx]; }
        }
        dig[r] = d;
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
    for (int r = 0; r < BZ_IPT; r++) {
        u32 e = w * (32 * BZ_IPT) + r * 32 + lane;
        if (e < tile_n) {
            int d = dig[r];
            u32 pos = lbase[d] + wh[w * 256 + d] + rnk[r];
            sdig[pos] = (u8)d;
            sval[pos] = val[r];
            if (MODE == 1) skey[pos] = key[r];
        }
    }
    __syncthreads();
    size_t ob = (size_t)b * a.stride;
    for (u32 p = threadIdx.x; p < tile_n; p += BZ_THREADS) {
        int d = sdig[p];
        u32 dst = toff[d] + (p - lbase[d]);
        if (MODE == 0) a.sa_out[ob + dst] = sval[p];
        else { a.key_out[ob + dst] = skey[p]; a.val_out[ob + dst] = sval[p]; }
    }
}
#endif

// --------------------------------------------------------------------------------------
// step 2: head flags after the 8-byte sort
// --------------------------------------------------------------------------------------
// 8 bytes of rotation s, big endian.  T is 4-byte aligned and padded so that aligned word loads
// up to offset n+11 stay inside the block's stride.
__device__ __forceinline__ u64 rot_key8(const u8 *T, u32 n, u32 s) {
    if (s + 8 <= n) {
        const u32 *w = (const u32 *)(T + (s & ~3u));
        u32 w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
        int sh = (s & 3) * 8;
        u32 a = __funnelshift_r(w0, w1, sh), c = __funnelshift_r(w1, w2, sh);
        return ((u64)__byte_perm(a, 0, 0x0123) << 32) | __byte_perm(c, 0, 0x0123);
    }
    u64 k = 0;
    u32 p = s;
    for (int q = 0; q < 8; q++) { k = (k << 8) | T[p]; p++; if (p >= n) p = 0; }
    return k;
}

// F[j] = 1 iff rotation SA[j] starts a new 8-byte-prefix group.  Tile aggregates:
// x = last head position in tile (or -1), z = number of unresolved rows in tile.
__global__ void __launch_bounds__(BZ_THREADS) k_init_flags(const u8 *T, const u32 *len, const u32 *SA, u8 *F,
                                                           u32 stride) {
    u32 b = blockIdx.y, n = len[b];
    u32 base = blockIdx.x * BZ_TILE;
    if (base >= n) return;
    const u8 *Tb = T + (size_t)b * stride;
    const u32 *sa = SA + (size_t)b * stride;
    __shared__ u64 sk[BZ_TILE + 1];
    for (u32 e = threadIdx.x; e < BZ_TILE + 1; e += BZ_THREADS) {
        // sk[e] = key of row base + e - 1
        u32 j = base + e;
        if (j >= 1 && j - 1 < n) sk[e] = rot_key8(Tb, n, sa[j - 1]);
    }
    __syncthreads();
    for (u32 e = threadIdx.x; e < BZ_TILE; e += BZ_THREADS) {
        u32 j = base + e;
        if (j < n) F[(size_t)b * stride + j] = (j == 0 || sk[e + 1] != sk[e]) ? 1 : 0;
    }
}

// Generic tile aggregate over head flags F (rows) : used after the initial sort.
__global__ void __launch_bounds__(BZ_THREADS) k_flags_agg(const u8 *F, const u32 *len, int4 *tagg, u32 stride,
                                                          u32 tiles) {
    u32 b = blockIdx.y, n = len[b];
    u32 base = blockIdx.x * BZ_TILE;
    if (base >= n) return;
    const u8 *f = F + (size_t)b * stride;
    int last = -1; u32 unres = 0;
    u32 j0 = base + threadIdx.x * BZ_IPT;
#pragma unroll
    for (int r = 0; r < BZ_IPT; r++) {
        u32 j = j0 + r;
        if (j < n) {
            bool h = f[j] != 0;
            bool hn = (j + 1 >= n) ? true : (f[j + 1] != 0);
            if (h) last = (int)j;
            if (!(h && hn)) unres++;
        }
    }
    __shared__ int wsi[8];
    __shared__ u32 wsu[8];
    int tot_last; u32 tot_un;
    block_excl_max(last, wsi, tot_last);
    block_excl_sum(unres, wsu, tot_un);
    if (threadIdx.x == 0) tagg[(size_t)b * tiles + blockIdx.x] = make_int4(tot_last, -1, (int)tot_un, 0);
}

// per block: exclusive scan of the tile aggregates (max, max, sum).  Writes the next list count.
// Lists are dropped (count 0) once the sorted depth `depth_after` covers the whole block.
__global__ void __launch_bounds__(256) k_tile_scan(int4 *tagg, const u32 *cntp, const u32 *len, u32 *cnt_out,
                                                   u32 tiles_stride, u32 depth_after) {
    u32 b = blockIdx.x;
    u32 cnt = cntp[b];
    u32 tiles = (cnt + BZ_TILE - 1) / BZ_TILE;
    int4 *a = tagg + (size_t)b * tiles_stride;
    __shared__ int wsi[8];
    __shared__ u32 wsu[8];
    int carry_a = -1, carry_b = -1; u32 carry_c = 0;
    for (u32 t0 = 0; t0 < tiles; t0 += 256) {
        u32 t = t0 + threadIdx.x;
        int4 v = (t < tiles) ? a[t] : make_int4(-1, -1, 0, 0);
        int ta, tb; u32 tc;
        int ea = block_excl_max(v.x, wsi, ta);
        int eb = block_excl_max(v.y, wsi, tb);
        u32 ec = block_excl_sum((u32)v.z, wsu, tc);
        if (t < tiles) a[t] = make_int4(max(ea, carry_a), max(eb, carry_b), (int)(ec + carry_c), 0);
        carry_a = max(carry_a, ta); carry_b = max(carry_b, tb); carry_c += tc;
        __syncthreads();
    }
    if (threadIdx.x == 0) cnt_out[b] = (depth_after >= len[b]) ? 0u : carry_c;
}

// rank[SA[j]] = head(j); unresolved rows appended to the list as (key = head << nbits, val = SA[j]).
__global__ void __launch_bounds__(BZ_THREADS) k_init_apply(const u8 *F, const u32 *len, const u32 *SA, u32 *RANK,
                                                           const int4 *tagg, u64 *KEY, u32 *VAL, u32 stride,
                                                           u32 tiles, int nbits) {
    u32 b = blockIdx.y, n = len[b];
    u32 base = blockIdx.x * BZ_TILE;
    if (base >= n) return;
    size_t ob = (size_t)b * stride;
    const u8 *f = F + ob;
    int4 carry = tagg[(size_t)b * tiles + blockIdx.x];
    u32 j0 = base + threadIdx.x * BZ_IPT;
    bool h[BZ_IPT], un[BZ_IPT];
    int last = -1; u32 unres = 0;
#pragma unroll
    for (int r = 0; r < BZ_IPT; r++) {
        u32 j = j0 + r;
        h[r] = false; un[r] = false;
        if (j < n) {
            h[r] = f[j] != 0;
            bool hn = (j + 1 >= n) ? true : (f[j + 1] != 0);
            un[r] = !(h[r] && hn);
            if (h[r]) last = (int)j;
            if (un[r]) unres++;
        }
    }
    __shared__ int wsi[8];
    __shared__ u32 wsu[8];
    int tl; u32 tu;
    int head = max(block_excl_max(last, wsi, tl), carry.x);
    u32 k = block_excl_sum(unres, wsu, tu) + (u32)carry.z;
#pragma unroll
    for (int r = 0; r < BZ_IPT; r++) {
        u32 j = j0 + r;
        if (j < n) {
            if (h[r]) head = (int)j;
            u32 s = SA[ob + j];
            RANK[ob + s] = (u32)head;
            if (un[r]) { KEY[ob + k] = (u64)(u32)head << nbits; VAL[ob + k] = s; k++; }
        }
    }
}

// --------------------------------------------------------------------------------------
// step 3: doubling round
// --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BZ_THREADS) k_gather(const u32 *cntp, const u32 *len, const u32 *RANK, u64 *KEY,
                                                       const u32 *VAL, u32 stride, u32 h, int nbits) {
    u32 b = blockIdx.y, cnt = cntp[b], n = len[b];
    size_t ob = (size_t)b * stride;
    u32 base = blockIdx.x * BZ_TILE;
#pragma unroll 4
    for (int r = 0; r < BZ_IPT; r++) {
        u32 k = base + r * BZ_THREADS + threadIdx.x;
        if (k < cnt) {
            u32 s = VAL[ob + k];
            u32 p = s + h;
            if (p >= n) p %= n;
            u64 key = KEY[ob + k];
            KEY[ob + k] = ((key >> nbits) << nbits) | RANK[ob + p];
        }
    }
}

// tile aggregates over the sorted list: x = last k starting a group (head field changes),
// y = last k starting a subgroup (whole key changes), z = entries whose subgroup has size > 1.
__global__ void __launch_bounds__(BZ_THREADS) k_list_agg(const u32 *cntp, const u64 *KEY, int4 *tagg, u32 stride,
                                                         u32 tiles, int nbits) {
    u32 b = blockIdx.y, cnt = cntp[b];
    u32 base = blockIdx.x * BZ_TILE;
    if (base >= cnt) return;
    const u64 *key = KEY + (size_t)b * stride;
    u32 k0 = base + threadIdx.x * BZ_IPT;
    int la = -1, lb = -1; u32 unres = 0;
    u64 prev = (k0 > 0 && k0 - 1 < cnt) ? key[k0 - 1] : 0;
    u64 cur = (k0 < cnt) ? key[k0] : 0;
#pragma unroll
    for (int r = 0; r < BZ_IPT; r++) {
        u32 k = k0 + r;
        if (k < cnt) {
            u64 nxt = (k + 1 < cnt) ? key[k + 1] : 0;
            bool hs = (k == 0) || cur != prev;
            bool hl = (k == 0) || (cur >> nbits) != (prev >> nbits);
            bool hs_next = (k + 1 >= cnt) || nxt != cur;
            if (hl) la = (int)k;
            if (hs) lb = (int)k;
            if (!(hs && hs_next)) unres++;
            prev = cur; cur = nxt;
        }
    }
    __shared__ int wsi[8];
    __shared__ u32 wsu[8];
    int ta, tb; u32 tu;
    block_excl_max(la, wsi, ta);
    block_excl_max(lb, wsi, tb);
    block_excl_sum(unres, wsu, tu);
    if (threadIdx.x == 0) tagg[(size_t)b * tiles + blockIdx.x] = make_int4(ta, tb, (int)tu, 0);
}

// write back refined order and ranks; re-compact the still unresolved entries into the next list
__global__ void __launch_bounds__(BZ_THREADS) k_list_apply(const u32 *cntp, const u64 *KEY, const u32 *VAL,
                                                           const int4 *tagg, u32 *SA, u32 *RANK, u64 *KEYN,
                                                           u32 *VALN, u32 stride, u32 tiles, int nbits) {
    u32 b = blockIdx.y, cnt = cntp[b];
    u32 base = blockIdx.x * BZ_TILE;
    if (base >= cnt) return;
    size_t ob = (size_t)b * stride;
    const u64 *key = KEY + ob;
    int4 carry = tagg[(size_t)b * tiles + blockIdx.x];
    u32 k0 = base + threadIdx.x * BZ_IPT;
    bool hsv[BZ_IPT], hlv[BZ_IPT], unv[BZ_IPT];
    u32 headv[BZ_IPT];
    int la = -1, lb = -1; u32 unres = 0;
    u64 prev = (k0 > 0 && k0 - 1 < cnt) ? key[k0 - 1] : 0;
    u64 cur = (k0 < cnt) ? key[k0] : 0;
#pragma unroll
    for (int r = 0; r < BZ_IPT; r++) {
        u32 k = k0 + r;
        hsv[r] = hlv[r] = unv[r] = false; headv[r] = 0;
        if (k < cnt) {
            u64 nxt = (k + 1 < cnt) ? key[k + 1] : 0;
            bool hs = (k == 0) || cur != prev;
            bool hl = (k == 0) || (cur >> nbits) != (prev >> nbits);
            bool hs_next = (k + 1 >= cnt) || nxt != cur;
            hsv[r] = hs; hlv[r] = hl; unv[r] = !(hs && hs_next);
            headv[r] = (u32)(cur >> nbits);
            if (hl) la = (int)k;
            if (hs) lb = (int)k;
            if (unv[r]) unres++;
            prev = cur; cur = nxt;
        }
    }
    __shared__ int wsi[8];
    __shared__ u32 wsu[8];
    int ta, tb; u32 tu;
    int kg = max(block_excl_max(la, wsi, ta), carry.x);     // list index of the current group's first entry
    int ks = max(block_excl_max(lb, wsi, tb), carry.y);     // list index of the current subgroup's first entry
    u32 ko = block_excl_sum(unres, wsu, tu) + (u32)carry.z; // output slot in the next list
#pragma unroll
    for (int r = 0; r < BZ_IPT; r++) {
        u32 k = k0 + r;
        if (k < cnt) {
            if (hlv[r]) kg = (int)k;
            if (hsv[r]) ks = (int)k;
            u32 g = headv[r];                 // row of the group's first member
            u32 row = g + (k - (u32)kg);
            u32 nh = g + ((u32)ks - (u32)kg); // row of the subgroup's first member = new rank
            u32 s = VAL[ob + k];
            SA[ob + row] = s;
            RANK[ob + s] = nh;
            if (unv[r]) { KEYN[ob + ko] = (u64)nh << nbits; VALN[ob + ko] = s; ko++; }
        }
    }
}

// --------------------------------------------------------------------------------------
// step 4
// --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BZ_THREADS) k_bwt_out(const u8 *T, const u32 *len, const u32 *SA, const u32 *RANK,
                                                        u8 *bwt, u32 *keyout, u32 stride) {
    u32 b = blockIdx.y, n = len[b];
    size_t ob = (size_t)b * stride;
    u32 base = blockIdx.x * BZ_TILE;
    if (blockIdx.x == 0 && threadIdx.x == 0) keyout[b] = n ? RANK[ob] : 0;
#pragma unroll 4
    for (int r = 0; r < BZ_IPT; r++) {
        u32 j = base + r * BZ_THREADS + threadIdx.x;
        if (j < n) {
            u32 s = SA[ob + j];
            bwt[ob + j] = T[ob + (s == 0 ? n - 1 : s - 1)];
        }
    }
}

}  // namespace

using radix::RadixArgs;
using radix::R_TILE;

#define LAUNCH_OK()                                                  \
    do {                                                             \
        ctx->prof_end();                                             \
        cudaError_t e_ = cudaGetLastError();                         \
        if (e_ != cudaSuccess) { ctx->fail("kernel launch", e_, __FILE__, __LINE__); return BZ2B200_E_CUDA; } \
    } while (0)

int bz_bwt_batch(bz2b200_ctx *ctx, const Batch &B, u8 *d_bwt, u32 *d_key) {
    if (B.nblk == 0) return BZ2B200_OK;
    cudaStream_t st = ctx->stream;
    const u64 ne_act = B.total_n;
    u64 lsum = 0;
    size_t ne = (size_t)B.nblk * B.stride;
    BZ_CHECK(ctx->d_SA.ensure(ne * 4));
    BZ_CHECK(ctx->d_SA2.ensure(ne * 4));
    BZ_CHECK(ctx->d_RANK.ensure(ne * 4));
    BZ_CHECK(ctx->d_F.ensure(ne + 16));
    BZ_CHECK(ctx->d_KEYA.ensure(ne * 8));
    BZ_CHECK(ctx->d_KEYB.ensure(ne * 8));
    BZ_CHECK(ctx->d_VALA.ensure(ne * 4));
    BZ_CHECK(ctx->d_VALB.ensure(ne * 4));
    BZ_CHECK(ctx->d_thist.ensure((size_t)B.nblk * (B.stride / radix::R_TILE) * 256 * 4));
    BZ_CHECK(ctx->d_tagg.ensure((size_t)B.nblk * B.tiles * sizeof(int4)));
    BZ_CHECK(ctx->d_cnt.ensure((size_t)B.nblk * 2 * 4));
    BZ_CHECK(ctx->h_small.ensure((size_t)B.nblk * 4 + 64));
    BwtWs W;
    W.SA = ctx->d_SA.as<u32>(); W.SA2 = ctx->d_SA2.as<u32>(); W.RANK = ctx->d_RANK.as<u32>();
    W.F = ctx->d_F.as<u8>();
    W.KEYA = ctx->d_KEYA.as<u64>(); W.KEYB = ctx->d_KEYB.as<u64>();
    W.VALA = ctx->d_VALA.as<u32>(); W.VALB = ctx->d_VALB.as<u32>();
    W.thist = ctx->d_thist.as<u32>(); W.tagg = ctx->d_tagg.as<int4>(); W.cnt = ctx->d_cnt.as<u32>();

    dim3 gfull((B.max_n + BZ_TILE - 1) / BZ_TILE, B.nblk);
    dim3 gfull_r((B.max_n + R_TILE - 1) / R_TILE, B.nblk);
    const u32 rtiles = B.stride / R_TILE;

    // one-kernel radix passes (decoupled look-back): digit totals, tile state (aliases thist), tickets
    const u32 DSTRIDE = 8 * 256;                                 // per block: up to 8 passes x 256 digit offsets
    BZ_CHECK(ctx->d_R.ensure(((size_t)B.nblk * DSTRIDE + 256) * 4));
    u32 *dcounts = ctx->d_R.as<u32>();
    u32 *tickets = dcounts + (size_t)B.nblk * DSTRIDE;          // one counter per epoch
    u32 *tstate = W.thist;
    BZ_CHECK(cudaMemsetAsync(dcounts, 0, ((size_t)B.nblk * DSTRIDE + 256) * 4, st));
    BZ_CHECK(cudaMemsetAsync(tstate, 0, (size_t)B.nblk * rtiles * 256 * 4, st));
    u32 epoch = 0;

    // ---- 1. initial 8-byte LSD sort (implicit keys) ----
    u32 *cur = nullptr, *src = nullptr;
    u32 *bufs[2] = {W.SA, W.SA2};
    // every pass has the same digit totals: the block's byte histogram (each byte is digit p of exactly one rotation)
    ctx->prof_begin(K_RADIX_HIST0, ne_act); radix::k_byte_hist<<<gfull, BZ_THREADS, 0, st>>>(B.T, B.len, dcounts, B.stride, DSTRIDE); LAUNCH_OK();
    ctx->prof_begin(K_RADIX_SCAN, (u64)B.nblk * DSTRIDE * 4); radix::k_digit_scan<<<B.nblk * 8, 256, 0, st>>>(dcounts); LAUNCH_OK();
    for (int p = 0; p < 8; p++) {
        radix::SweepArgs s{};
        RadixArgs &a = s.r;
        a.T = B.T; a.len = B.len; a.cnt = B.len; a.sa_in = src; a.sa_out = bufs[p & 1];
        a.thist = W.thist; a.stride = B.stride; a.rtiles = rtiles; a.off = 7 - p;
        epoch++;
        s.dbase = dcounts; s.dbase_stride = DSTRIDE; s.tstate = tstate; s.ticket = tickets + epoch; s.epoch = epoch;
        s.tiles_x = gfull_r.x; s.nblk = (u32)B.nblk; s.group = radix::sweep_group();
        ctx->prof_begin(K_RADIX_SCATTER0, ne_act * 9); radix::k_radix_onesweep<0><<<radix::sweep_grid(gfull_r.x, B.nblk), BZ_THREADS, 0, st>>>(s); LAUNCH_OK();
        src = bufs[p & 1];
    }
    cur = src;   // after 8 passes: bufs[1] = SA2
    u32 *SA = cur;

    // ---- 2. heads, ranks, first unresolved list ----
    ctx->prof_begin(K_INIT_FLAGS, ne_act * 13); k_init_flags<<<gfull, BZ_THREADS, 0, st>>>(B.T, B.len, SA, W.F, B.stride); LAUNCH_OK();
    ctx->prof_begin(K_FLAGS_AGG, ne_act); k_flags_agg<<<gfull, BZ_THREADS, 0, st>>>(W.F, B.len, W.tagg, B.stride, B.tiles); LAUNCH_OK();
    ctx->prof_begin(K_TILE_SCAN, (u64)B.nblk * B.tiles * 32); k_tile_scan<<<B.nblk, 256, 0, st>>>(W.tagg, B.len, B.len, W.cnt, B.tiles, 8u); LAUNCH_OK();
    ctx->prof_begin(K_INIT_APPLY, ne_act * 21); k_init_apply<<<gfull, BZ_THREADS, 0, st>>>(W.F, B.len, SA, W.RANK, W.tagg, W.KEYA, W.VALA, B.stride, B.tiles, B.nbits);
    LAUNCH_OK();

    // ---- 3. doubling rounds ----
    u64 *K0 = W.KEYA, *K1 = W.KEYB; u32 *V0 = W.VALA, *V1 = W.VALB;
    u32 *cnt_cur = W.cnt, *cnt_nxt = W.cnt + B.nblk;
    u32 *h_cnt = ctx->h_small.as<u32>();
    int passes = (2 * B.nbits + 7) / 8;
    u64 rounds = 0, listsum = 0;
    for (u32 h = 8;; h *= 2) {
        BZ_CHECK(cudaMemcpyAsync(h_cnt, cnt_cur, (size_t)B.nblk * 4, cudaMemcpyDeviceToHost, st));
        BZ_CHECK(cudaStreamSynchronize(st));
        u32 maxc = 0;
        lsum = 0;
        for (int b = 0; b < B.nblk; b++) { if (h_cnt[b] > maxc) maxc = h_cnt[b]; lsum += h_cnt[b]; }
        listsum += lsum;
        if (maxc == 0) break;
        if (h >= (1u << 30)) { ctx->err = "bwt: doubling did not terminate"; return BZ2B200_E_CUDA; }
        rounds++;
        dim3 gl((maxc + BZ_TILE - 1) / BZ_TILE, B.nblk);
        dim3 gl_r((maxc + R_TILE - 1) / R_TILE, B.nblk);
        ctx->prof_begin(K_GATHER, lsum * 24); k_gather<<<gl, BZ_THREADS, 0, st>>>(cnt_cur, B.len, W.RANK, K0, V0, B.stride, h, B.nbits); LAUNCH_OK();
        // digit totals of all passes from one read of the keys
        BZ_CHECK(cudaMemsetAsync(dcounts, 0, (size_t)B.nblk * DSTRIDE * 4, st));
        ctx->prof_begin(K_RADIX_HIST1, lsum * 8); radix::k_list_hist<<<gl, BZ_THREADS, 0, st>>>(K0, cnt_cur, dcounts, B.stride, passes); LAUNCH_OK();
        ctx->prof_begin(K_RADIX_SCAN, (u64)B.nblk * DSTRIDE * 4); radix::k_digit_scan<<<B.nblk * 8, 256, 0, st>>>(dcounts); LAUNCH_OK();
        if (epoch + (u32)passes > 254) {                        // epochs are 8 bits: start over with a clean state array
            BZ_CHECK(cudaMemsetAsync(tstate, 0, (size_t)B.nblk * rtiles * 256 * 4, st));
            BZ_CHECK(cudaMemsetAsync(tickets, 0, 256 * 4, st));
            epoch = 0;
        }
        for (int p = 0; p < passes; p++) {
            radix::SweepArgs s{};
            RadixArgs &a = s.r;
            a.T = B.T; a.len = B.len; a.cnt = cnt_cur; a.key_in = K0; a.key_out = K1; a.val_in = V0; a.val_out = V1;
            a.thist = W.thist; a.stride = B.stride; a.rtiles = rtiles; a.shift = 8 * p;
            epoch++;
            s.dbase = dcounts + p * 256; s.dbase_stride = DSTRIDE; s.tstate = tstate; s.ticket = tickets + epoch; s.epoch = epoch;
            s.tiles_x = gl_r.x; s.nblk = (u32)B.nblk; s.group = radix::sweep_group();
            ctx->prof_begin(K_RADIX_SCATTER1, lsum * 24); radix::k_radix_onesweep<1><<<radix::sweep_grid(gl_r.x, B.nblk), BZ_THREADS, 0, st>>>(s); LAUNCH_OK();
            u64 *tk = K0; K0 = K1; K1 = tk;
            u32 *tv = V0; V0 = V1; V1 = tv;
        }
        // sorted list now in (K0, V0); the next list is written to (K1, V1)
        ctx->prof_begin(K_LIST_AGG, lsum * 8); k_list_agg<<<gl, BZ_THREADS, 0, st>>>(cnt_cur, K0, W.tagg, B.stride, B.tiles, B.nbits); LAUNCH_OK();
        ctx->prof_begin(K_TILE_SCAN, (u64)B.nblk * B.tiles * 32); k_tile_scan<<<B.nblk, 256, 0, st>>>(W.tagg, cnt_cur, B.len, cnt_nxt, B.tiles, 2 * h); LAUNCH_OK();
        ctx->prof_begin(K_LIST_APPLY, lsum * 32); k_list_apply<<<gl, BZ_THREADS, 0, st>>>(cnt_cur, K0, V0, W.tagg, SA, W.RANK, K1, V1, B.stride, B.tiles, B.nbits);
        LAUNCH_OK();
        { u64 *tk = K0; K0 = K1; K1 = tk; u32 *tv = V0; V0 = V1; V1 = tv; }
        { u32 *tc = cnt_cur; cnt_cur = cnt_nxt; cnt_nxt = tc; }
    }
    // ---- 4. output ----
    ctx->prof_begin(K_BWT_OUT, ne_act * 6); k_bwt_out<<<gfull, BZ_THREADS, 0, st>>>(B.T, B.len, SA, W.RANK, d_bwt, d_key, B.stride); LAUNCH_OK();
    ctx->bwt_stats[0] = (u64)B.nblk; ctx->bwt_stats[1] = ne_act; ctx->bwt_stats[2] = rounds; ctx->bwt_stats[3] = listsum;
    return BZ2B200_OK;
}
