// mtf_emit.cuh -- second version of the MTF/RLE2 replay kernel (included by mtf.cu inside its namespace).
//
// One warp replays one 1024-byte chunk of the BWT string (rle2_mtf.rs:61-131).  Differences from the
// first version (kept in mtf.cu for reference): the first 32 list positions live one per lane ("front
// window": a hit is a ballot + one shuffle), positions 32..255 live as 7 bytes per lane in a u64, and
// rank-0 positions (L[i] == L[i-1]) are never visited: a ballot per 32 bytes gives the non-zero
// positions and the loop jumps from one to the next, adding the gap to the pending zero run.

__global__ void __launch_bounds__(BZ_THREADS) k_mtf_emit2(const u8 *Lall, const u32 *len, const u32 *usedbits,
                                                          const int *pm, const u32 *zbefore, const u32 *ooff,
                                                          const u32 *m_in, u16 *sym, u32 *freq, u32 stride,
                                                          u32 nch_stride) {
    u32 b = blockIdx.y, n = len[b];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    u32 c = blockIdx.x * WPB + w;
    u32 a = c * CH;
    __shared__ int sval[WPB][256];
    __shared__ __align__(8) u8 slist[WPB][256];
    __shared__ u32 sfreq[256];
    __shared__ u8 sused[256];
    __shared__ int s_nused;
    sfreq[threadIdx.x] = 0;
    {   // compact list of the block's used byte values (ascending)
        const u32 *ub = usedbits + b * 8;
        u32 t = threadIdx.x;
        u32 before = 0;
        for (u32 k = 0; k < (t >> 5); k++) before += __popc(ub[k]);
        before += __popc(ub[t >> 5] & ((1u << (t & 31)) - 1));
        if ((ub[t >> 5] >> (t & 31)) & 1) sused[before] = (u8)t;
        if (t == 255) s_nused = (int)(before + ((ub[7] >> 31) & 1));
    }
    __syncthreads();
    const int nused = s_nused;
    bool active = a < n;
    u32 runa = 0, runb = 0;
    if (active) {
        u32 e = min(a + CH, n);
        const u8 *L = Lall + (size_t)b * stride;
        u16 *so = sym + (size_t)b * stride;
        const int *v = pm + ((size_t)b * nch_stride + c) * 256;
        for (int k = lane; k < 256; k += 32) { sval[w][k] = v[k]; slist[w][k] = 0; }
        __syncwarp();
        // start list: USED byte values sorted by last occurrence before the chunk, most recent first
        for (int j = lane; j < nused; j += 32) {
            int s = sused[j];
            int mine = sval[w][s], rk = 0;
            for (int t = 0; t < nused; t++) rk += (sval[w][sused[t]] > mine);
            slist[w][rk] = (u8)s;
        }
        __syncwarp();
        u32 fw = slist[w][lane];                               // list position `lane`
        u64 tl = 0;                                            // list positions 32 + 7*lane + k in byte k (k < 7)
#pragma unroll
        for (int k = 0; k < 7; k++) tl |= (u64)slist[w][32 + 7 * lane + k] << (8 * k);
        const u64 LOW7 = 0x00FFFFFFFFFFFFFFull;
        u32 z = zbefore[(size_t)b * nch_stride + c];
        u32 o = ooff[(size_t)b * nch_stride + c];
        const u32 o_start = o;
        u32 obase = o & ~31u;
        u32 staged = 0;
#define EMIT2(SYMV)                                                                   \
        do {                                                                          \
            if ((o & 31u) == (u32)lane) staged = (SYMV);                              \
            o++;                                                                      \
            if ((o & 31u) == 0) {                                                     \
                if (obase + lane >= o_start) so[obase + lane] = (u16)staged;          \
                obase = o;                                                            \
            }                                                                         \
        } while (0)
#define FLUSH_ZEROS2()                                                                \
        do {                                                                          \
            u32 zz = z + 1; int nd = 31 - __clz(zz);                                  \
            for (int q = 0; q < nd; q++) { u32 bit = (zz >> q) & 1u; if (bit) runb++; else runa++; EMIT2(bit); } \
            z = 0;                                                                    \
        } while (0)
        u32 prev_last = __shfl_sync(0xffffffffu, fw, 0);       // the list front stands in for "previous byte" at the chunk start
        for (u32 i0 = a; i0 < e; i0 += 32) {
            int cntk = (int)min(32u, e - i0);
            u32 my = (lane < cntk) ? L[i0 + lane] : 0;
            u32 pb = __shfl_up_sync(0xffffffffu, my, 1);
            if (lane == 0) pb = prev_last;
            unsigned nzmask = __ballot_sync(0xffffffffu, lane < cntk && my != pb);   // rank != 0  <=>  differs from the previous byte
            prev_last = __shfl_sync(0xffffffffu, my, cntk - 1);
            int lastk = -1;
            while (nzmask) {
                int k = __ffs(nzmask) - 1;
                nzmask &= nzmask - 1;
                z += (u32)(k - lastk - 1);
                lastk = k;
                u32 ch = __shfl_sync(0xffffffffu, my, k);
                if (z) FLUSH_ZEROS2();
                u32 pos;
                unsigned hit = __ballot_sync(0xffffffffu, fw == ch);
                u32 up = __shfl_up_sync(0xffffffffu, fw, 1);
                if (hit) {                                     // rank < 32
                    int p = __ffs(hit) - 1;
                    pos = (u32)p;
                    if (lane <= p) fw = lane ? up : ch;
                } else {                                       // rank >= 32: search the packed tail
                    u64 x = tl ^ (0x0001010101010101ull * ch);
                    u64 zm = (x - 0x0001010101010101ull) & ~x & 0x0080808080808080ull;
                    unsigned bal = __ballot_sync(0xffffffffu, zm != 0);
                    int Lh = __ffs(bal) - 1;
                    int kb = (__ffsll((long long)zm) - 1) >> 3;
                    kb = __shfl_sync(0xffffffffu, kb, Lh);
                    pos = 32u + 7u * (u32)Lh + (u32)kb;
                    u32 carry = __shfl_sync(0xffffffffu, fw, 31);          // old position 31 moves to 32
                    u32 top = (u32)(tl >> 48) & 0xffu;                      // byte 6 of this lane's tail
                    u32 incoming = __shfl_up_sync(0xffffffffu, top, 1);
                    if (lane == 0) incoming = carry;
                    if (lane < Lh) tl = ((tl << 8) | incoming) & LOW7;
                    else if (lane == Lh) {
                        u64 lowmask = kb ? ((1ull << (8 * kb)) - 1) : 0ull;
                        u64 keepmask = (~((1ull << (8 * (kb + 1))) - 1)) & LOW7;
                        tl = (tl & keepmask) | (((tl & lowmask) << 8) | incoming);
                    }
                    fw = lane ? up : ch;
                }
                if (lane == 0) atomicAdd(&sfreq[pos], 1u);
                EMIT2(pos + 1);
            }
            z += (u32)(cntk - 1 - lastk);
        }
        bool next_nz = (e >= n) ? true : (L[e] != L[e - 1]);
        if (z && next_nz) FLUSH_ZEROS2();
        if (obase + lane >= o_start && obase + lane < o) so[obase + lane] = (u16)staged;
        if (e >= n && lane == 0) {
            u32 nu = 0;
            for (int k = 0; k < 8; k++) nu += __popc(usedbits[b * 8 + k]);
            so[m_in[b] - 1] = (u16)(nu + 1);                   // EOB = nused + 1 (rle2_mtf.rs:42,:166)
        }
#undef EMIT2
#undef FLUSH_ZEROS2
    }
    if (lane == 0 && active) {
        if (runa) atomicAdd(&sfreq[0], runa);
        if (runb) atomicAdd(&sfreq[1], runb);
    }
    __syncthreads();
    if (sfreq[threadIdx.x]) atomicAdd(&freq[b * 256 + threadIdx.x], sfreq[threadIdx.x]);
}
