// mtf_emit3.cuh -- third version of the MTF/RLE2 replay kernel (included by mtf.cu inside its namespace).
//
// One warp replays one 1024-byte chunk of the BWT string (rle2_mtf.rs:61-131), 32 positions per step, all
// lanes busy.  The earlier versions walked the chunk one byte at a time with the whole warp serving that one
// byte (~50 warp instructions per non-zero rank: the kernel was bound by instruction issue, 107 GB/s).
// Here a lane owns a position and the MTF list is kept as its INVERSE, pos[v] = list position of byte value v
// (256 bytes of shared memory per warp).  For the 32 bytes c_0..c_31 of a step, with j_i = the latest earlier
// lane holding the same byte as lane i:
//   rank_i = #distinct values in lanes (j_i, i)                               if j_i exists
//          = pos[c_i] + #distinct values v in lanes [0, i) with pos[v] > pos[c_i]   otherwise
// (every distinct value touched after the last access of c_i pushes it down by one).  Both counts are taken by
// looping over the step's DISTINCT values (few in a BWT string) and testing each value's lane mask against the
// lane's window.  The same loop moves the list: a value not in the step goes down by the number of step values
// that were behind it; a value in the step lands at the number of distinct values accessed after its last
// occurrence.  A lane keeps 8 list positions as bytes; "position < that of the touched value" is evaluated for all
// of them with 16-bit-lane arithmetic (the carry out of 255 - pos + rv).
// rank 0 <=> byte equals its predecessor, so zero runs, RUNA/RUNB digits and output offsets inside the step are
// plain lane arithmetic plus one warp scan.

// The register copy of the list positions covers only what a block uses: with nused <= 32 * SPL byte values a lane owns
// the SPL used values of slots [SPL * lane, SPL * lane + SPL) (compact, ascending) instead of the 8 byte values
// [8 * lane, 8 * lane + 8) -- text has 30-100 distinct bytes, so the per-distinct-value update of the inner loop works
// on 1, 1 or 2 words instead of 4.  SPL = 8 keeps the value-indexed 8-byte load / store (binary data: every value used).
template <int SPL>
__device__ __forceinline__ void mtf_emit_chunk(const u8 *L, u16 *so, u32 a, u32 e, int nused, const u8 *sused, u8 *sposw,
                                               u32 front, u32 z, u32 o, u32 *sfreq, u32 &runa, u32 &runb, u32 &o_out, u32 &z_out) {
    const int lane = threadIdx.x & 31;
    const u32 lt = (1u << lane) - 1u;
    constexpr int NW = SPL == 8 ? 4 : (SPL + 1) / 2;            // position words per lane (two 16-bit lanes each)
    const u32 M2 = 0x00ff00ffu;
    u32 prev_last = front;                                      // the list front stands in for "previous byte"
    u64 *pos8 = (u64 *)sposw + lane;                            // SPL == 8: values 8*lane .. 8*lane+7
    u32 myv[SPL == 8 ? 1 : SPL];                                // SPL < 8: the byte values this lane owns (256 = none)
    if (SPL < 8) {
#pragma unroll
        for (int q = 0; q < (SPL == 8 ? 1 : SPL); q++) {
            int slot = lane * SPL + q;
            myv[q] = slot < nused ? (u32)sused[slot] : 256u;
        }
    }
    u32 chn = a + lane < e ? (u32)L[a + lane] : 0u;             // the next step's byte is loaded one step ahead
    for (u32 i0 = a; i0 < e; i0 += 32) {
        const int cntk = (int)min(32u, e - i0);
        const bool valid = lane < cntk;
        const u32 ch = valid ? chn : (0x100u | (u32)lane);
        if (i0 + 32 + lane < e) chn = (u32)L[i0 + 32 + lane];
        u32 pb = __shfl_up_sync(0xffffffffu, ch, 1);
        if (lane == 0) pb = prev_last;
        const bool nz = valid && ch != pb;                  // rank != 0  <=>  differs from the previous byte
        const u32 nzmask = __ballot_sync(0xffffffffu, nz);
        prev_last = __shfl_sync(0xffffffffu, ch, cntk - 1);
        if (nzmask == 0) { z += (u32)cntk; continue; }      // the whole step continues one run
        const u32 peers = __match_any_sync(0xffffffffu, ch);
        const u32 r0 = valid ? (u32)sposw[ch] : 0u;
        const u32 earlier = peers & lt;
        const bool has_j = earlier != 0;
        // lanes strictly between the previous occurrence (or the step start) and this lane
        const u32 wmask = has_j ? (lt & ~((2u << (31 - __clz(earlier))) - 1u)) : lt;
        const bool is_last = valid && (peers >> lane) == 1u;    // no later lane holds the same byte
        u32 firstmask = __ballot_sync(0xffffffffu, valid && !has_j);
        const u32 lastmask = __ballot_sync(0xffffffffu, is_last);
        // positions as words of two 16-bit lanes: cm = 255 - pos (compare operand), np = new position
        u32 np[NW], cm[NW];
        if (SPL == 8) {
            const u64 op = *pos8;
            np[0] = (u32)op & M2; np[1 % NW] = ((u32)op >> 8) & M2; np[2 % NW] = (u32)(op >> 32) & M2; np[3 % NW] = ((u32)(op >> 32) >> 8) & M2;
        } else {
#pragma unroll
            for (int j = 0; j < NW; j++) {
                u32 lo = myv[(2 * j) % (SPL == 8 ? 1 : SPL)] < 256u ? (u32)sposw[myv[(2 * j) % (SPL == 8 ? 1 : SPL)]] : 255u;
                u32 hi = 255u;
                if (2 * j + 1 < SPL) hi = myv[(2 * j + 1) % (SPL == 8 ? 1 : SPL)] < 256u ? (u32)sposw[myv[(2 * j + 1) % (SPL == 8 ? 1 : SPL)]] : 255u;
                np[j] = lo | (hi << 16);
            }
        }
#pragma unroll
        for (int q = 0; q < NW; q++) cm[q] = M2 - np[q];
        const u32 thr = has_j ? 0u : r0 + 1u;               // "pos[v] > pos[c]" only matters when c has no earlier occurrence
        u32 rank = has_j ? 0u : r0;
#pragma unroll 1
        while (firstmask) {                                 // one iteration per distinct value of the step
            int k = __ffs(firstmask) - 1;
            firstmask &= firstmask - 1;
            u32 P = __shfl_sync(0xffffffffu, peers, k);
            u32 rv = __shfl_sync(0xffffffffu, r0, k);
            if ((P & wmask) != 0 && rv >= thr) rank++;
            const u32 rv2 = rv * 0x00010001u;
#pragma unroll
            for (int q = 0; q < NW; q++) np[q] += ((cm[q] + rv2) >> 8) & 0x00010001u;    // +1 where pos < rv: rv's value moves ahead
        }
        if (SPL == 8) {
            u32 lo = (np[0] & M2) | ((np[1 % NW] & M2) << 8), hi = (np[2 % NW] & M2) | ((np[3 % NW] & M2) << 8);
            *pos8 = ((u64)hi << 32) | lo;
        } else {
#pragma unroll
            for (int q = 0; q < (SPL == 8 ? 1 : SPL); q++)
                if (myv[q] < 256u) sposw[myv[q]] = (u8)(np[q >> 1] >> (16 * (q & 1)));
        }
        __syncwarp();
        if (is_last) sposw[ch] = (u8)__popc(lastmask & ~lt & ~(1u << lane));
        __syncwarp();
        // ---- zero runs and output slots ----
        u32 below = nzmask & lt;
        u32 zrun = 0, nd = 0;
        if (nz) {
            zrun = below ? (u32)(lane - (31 - __clz(below)) - 1) : z + (u32)lane;
            nd = zrun ? (u32)(31 - __clz(zrun + 1)) : 0u;
        }
        u32 emits = nz ? nd + 1u : 0u;
        u32 inc = emits;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u32 t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        u32 total = __shfl_sync(0xffffffffu, inc, 31);
        if (nz) {
            u32 off = o + inc - emits;
            u32 zz = zrun + 1;
#pragma unroll 1
            for (u32 q = 0; q < nd; q++) so[off + q] = (u16)((zz >> q) & 1u);     // RUNA = 0, RUNB = 1 (rle2_mtf.rs:68-100)
            u32 ones = __popc(zz & ((1u << nd) - 1u));
            runb += ones; runa += nd - ones;
            so[off + nd] = (u16)(rank + 1);
        }
        {   // freq[rank] (rle2_mtf.rs:104), one shared-memory atomic per distinct rank of the step
            u32 key = nz ? rank : 0xffffu;
            u32 same = __match_any_sync(0xffffffffu, key);
            if (nz && (same & lt) == 0) atomicAdd(&sfreq[rank], (u32)__popc(same));
        }
        o += total;
        z = (u32)(cntk - 1 - (31 - __clz(nzmask)));        // zeros after the last non-zero rank of the step
    }
    o_out = o; z_out = z;
}

#ifndef BZ_MTF_COMPACT
#define BZ_MTF_COMPACT 1
#endif
// EWPB warps (= chunks) per CTA: chunks differ in cost (distinct values per step), and a CTA keeps its slot until its
// slowest warp is done -- with 8 warps 17% of the kernel's stall samples sat at the barrier in front of the freq flush.
#ifndef BZ_MTF_EWPB
#define BZ_MTF_EWPB 2
#endif
constexpr int EWPB = BZ_MTF_EWPB;
__global__ void __launch_bounds__(32 * EWPB) k_mtf_emit3(const u8 *Lall, const u32 *len, const u32 *usedbits,
                                                          const int *pm, const u32 *zbefore, const u32 *ooff,
                                                          const u32 *m_in, u16 *sym, u32 *freq, u32 stride,
                                                          u32 nch_stride) {
    u32 b = blockIdx.y, n = len[b];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    u32 c = blockIdx.x * EWPB + w;
    u32 a = c * CH;
    __shared__ int sval[EWPB][256];
    __shared__ __align__(8) u8 spos[EWPB][256];     // list position of every byte value
    __shared__ u32 sfreq[256];
    __shared__ u8 sused[256];
    __shared__ int s_nused;
    __shared__ u32 s_front[EWPB];
    for (int k = threadIdx.x; k < 256; k += 32 * EWPB) sfreq[k] = 0;
    {   // compact list of the block's used byte values (ascending)
        const u32 *ub = usedbits + b * 8;
        for (u32 t = threadIdx.x; t < 256; t += 32 * EWPB) {
            u32 before = 0;
            for (u32 k = 0; k < (t >> 5); k++) before += __popc(ub[k]);
            before += __popc(ub[t >> 5] & ((1u << (t & 31)) - 1));
            if ((ub[t >> 5] >> (t & 31)) & 1) sused[before] = (u8)t;
            if (t == 255) s_nused = (int)(before + ((ub[7] >> 31) & 1));
        }
    }
    __syncthreads();
    const int nused = s_nused;
    const bool active = a < n;
    u32 runa = 0, runb = 0;
    if (active) {
        const u32 e = min(a + CH, n);
        const u8 *L = Lall + (size_t)b * stride;
        u16 *so = sym + (size_t)b * stride;
        const int *v = pm + ((size_t)b * nch_stride + c) * 256;
        for (int k = lane; k < 256; k += 32) { sval[w][k] = v[k]; spos[w][k] = 255; }
        __syncwarp();
        // start state: pos[s] = number of used values seen more recently than s before the chunk
        for (int j = lane; j < nused; j += 32) {
            int s = sused[j];
            int mine = sval[w][s], rk = 0;
            for (int t = 0; t < nused; t++) rk += (sval[w][sused[t]] > mine);
            spos[w][s] = (u8)rk;
            if (rk == 0) s_front[w] = (u32)s;
        }
        __syncwarp();
        u32 z = zbefore[(size_t)b * nch_stride + c];            // pending zero run
        u32 o = ooff[(size_t)b * nch_stride + c];               // next output slot
        const u32 front = s_front[w];
        if (BZ_MTF_COMPACT && nused <= 32) mtf_emit_chunk<1>(L, so, a, e, nused, sused, spos[w], front, z, o, sfreq, runa, runb, o, z);
        else if (BZ_MTF_COMPACT && nused <= 64) mtf_emit_chunk<2>(L, so, a, e, nused, sused, spos[w], front, z, o, sfreq, runa, runb, o, z);
        else if (BZ_MTF_COMPACT && nused <= 128) mtf_emit_chunk<4>(L, so, a, e, nused, sused, spos[w], front, z, o, sfreq, runa, runb, o, z);
        else mtf_emit_chunk<8>(L, so, a, e, nused, sused, spos[w], front, z, o, sfreq, runa, runb, o, z);
        if (lane == 0) {
            bool next_nz = (e >= n) ? true : (L[e] != L[e - 1]);
            if (z && next_nz) {
                u32 zz = z + 1; u32 nd = (u32)(31 - __clz(zz));
                for (u32 q = 0; q < nd; q++) so[o + q] = (u16)((zz >> q) & 1u);
                u32 ones = __popc(zz & ((1u << nd) - 1u));
                runb += ones; runa += nd - ones;
            }
            if (e >= n) so[m_in[b] - 1] = (u16)(nused + 1);     // EOB = nused + 1 (rle2_mtf.rs:42,:166)
        }
    }
    if (active) {
        runa = __reduce_add_sync(0xffffffffu, runa);
        runb = __reduce_add_sync(0xffffffffu, runb);
        if (lane == 0) {
            if (runa) atomicAdd(&sfreq[0], runa);
            if (runb) atomicAdd(&sfreq[1], runb);
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 256; k += 32 * EWPB) if (sfreq[k]) atomicAdd(&freq[b * 256 + k], sfreq[k]);
}
