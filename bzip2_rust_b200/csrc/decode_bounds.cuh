// decode_bounds.cuh -- bit offset of every 50-symbol group of every block (included by decode.cu after decode4.cuh).
//
// Where a Huffman code starts depends on every code before it (decompress.rs:293-358): the one serial dependence of
// the entropy stage.  The first versions walked it with one lane per block, code by code (two per probe with a pair
// table): a chain of ~260 000 dependent shared-memory probes per block, 18.5 ms for 100 MB however many SMs idle.
// Here the chain is shortened to 7 dependent look-ups per GROUP:
//   k_dec_jumps   all SMs, every bit offset o of a block's data, every table t of the block: len_t(o) from the 10-bit
//                 LUT, then three pointer-doubling steps in shared memory -> J8_t(o) = bits spanned by the 8 codes
//                 that start at o under table t, and J2_t(o) likewise for 2 codes (0 = "special": an EOB, a malformed
//                 code or the end of the range among them).  One byte each, written window by window.
//   k_dec_bounds  one lane per block: a group of 50 codes under its selector's table is 6 J8 jumps + 1 J2 jump.
//                 The tables stream through a ring of shared-memory windows filled by 1-D bulk copies
//                 (cp.async.bulk + mbarrier) that the lane posts four windows ahead, so every look-up is an LDS.
//                 At a special entry it finishes code by code (the last group of a block; corrupt data; a range cut
//                 short by a chance magic inside the block's data).
constexpr int JW = 2048;              // bit offsets per window
constexpr int JHALO = 160;            // 8 codes of at most 20 bits
constexpr int JRING = 4;              // windows resident in the walker's ring
constexpr int JTILE = JW + JHALO;
constexpr size_t JRING_BYTES = (size_t)6 * 2 * JRING * JW;

// windows produced for a range of R bits: one more than needed, so that the window after the one holding any offset
// below R exists (all special) and the walker never indexes past what was produced
__host__ __device__ __forceinline__ u32 jump_windows(u64 R) { return (u32)((R + JW - 1) / JW) + 1u; }

// grid = (blocks of the slice, windows): CTAs are dealt block-fastest, so window w of EVERY block is produced before
// window w + 1 of any -- the walkers of all blocks follow the producer at the same distance.  flag[block][w] = 1
// (release) once window w is in global memory.
__global__ void __launch_bounds__(256) k_dec_jumps(const u8 *in, size_t n, const DecTables *tabs, const u64 *rend,
                                                   u8 *jt, size_t jstride, u32 *flags, u32 fstride, u32 b0) {
    const u32 b = b0 + blockIdx.x, w = blockIdx.y;
    const DecTables *tb = tabs + b;
    if (tb->status) return;
    const int T = (int)tb->T, alpha = (int)tb->alpha;
    const u64 base_bit = (tb->data_bit >> 5) << 5;
    const u64 end_bit = min(rend[b], (u64)n * 8);
    const u64 R = end_bit > base_bit ? end_bit - base_bit : 0;
    if (w >= jump_windows(R)) return;
    __shared__ u32 sbits[JTILE / 32 + 3];
    __shared__ u16 lut[6][1 << LUTBITS];
    __shared__ u8 A[6][JTILE], B[6][JTILE];
    const int tid = threadIdx.x;
    for (int i = tid; i < T * (1 << LUTBITS) / 2; i += 256) ((u32 *)&lut[0][0])[i] = ((const u32 *)&tb->lut[0][0])[i];
    {
        const u32 *wsrc = (const u32 *)in;
        const u64 last = (u64)((n + 3) / 4) - 1;
        const u64 w0 = (base_bit >> 5) + (u64)w * (JW / 32);
        for (int i = tid; i < JTILE / 32 + 3; i += 256) sbits[i] = __byte_perm(__ldg(wsrc + min(w0 + (u64)i, last)), 0, 0x0123);
    }
    __syncthreads();
    const u64 o_base = (u64)w * JW;
    for (int o = tid; o < JTILE; o += 256) {
        const u64 two = ((u64)sbits[o >> 5] << 32) | sbits[(o >> 5) + 1];
        const u32 b20 = (u32)((two << (o & 31)) >> 44);
        const bool inside = o_base + (u64)o < R;
        for (int t = 0; t < T; t++) {
            u32 len = 0;
            if (inside) {
                const u32 e = lut[t][b20 >> (20 - LUTBITS)];
                int sym = -1;
                if (e) { len = e >> 9; sym = (int)(e & 511u); }
                else {                                                   // longer than the LUT: canonical decode (decompress.rs:306-340)
                    int l = LUTBITS + 1;
                    int code = (int)(b20 >> (20 - l));
                    while (l <= 20 && code > tb->limit[t][l]) { l++; code = (int)(b20 >> (20 - min(l, 20))); }
                    if (l <= 20) {
                        int pi = code + tb->base[t][l];
                        if (pi >= 0 && pi < alpha) { len = (u32)l; sym = (int)tb->perm[t][pi]; }
                    }
                }
                if (sym == alpha - 1) len = 0;                           // EOB: the walker finishes code by code
            }
            A[t][o] = (u8)len;
        }
    }
    __syncthreads();
    for (int o = tid; o < JW + 140; o += 256)
        for (int t = 0; t < T; t++) { u32 a = A[t][o]; u32 c = a ? A[t][o + a] : 0u; B[t][o] = (u8)((a && c) ? a + c : 0u); }
    __syncthreads();
    for (int o = tid; o < JW + 100; o += 256)
        for (int t = 0; t < T; t++) { u32 a = B[t][o]; u32 c = a ? B[t][o + a] : 0u; A[t][o] = (u8)((a && c) ? a + c : 0u); }   // 4 codes (lengths no longer needed)
    __syncthreads();
    u8 *dst = jt + (size_t)blockIdx.x * jstride + (size_t)w * ((size_t)T * 2 * JW);
    for (int o = tid; o < JW; o += 256)
        for (int t = 0; t < T; t++) {
            u32 a = A[t][o]; u32 c = a ? A[t][o + a] : 0u;
            dst[(size_t)(2 * t) * JW + o] = (u8)((a && c) ? a + c : 0u);
            dst[(size_t)(2 * t + 1) * JW + o] = B[t][o];
        }
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + (size_t)blockIdx.x * fstride + w), "r"(1u) : "memory");
    }
}

// ---- mbarrier / bulk-copy helpers (SASS: SYNCS.*, UBLKCP) ----
__device__ __forceinline__ u32 jb_smem(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void jb_init(u64 *bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(jb_smem(bar)) : "memory");
}
__device__ __forceinline__ void jb_expect(u64 *bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(jb_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void jb_copy(void *dst, const void *src, u32 bytes, u64 *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(jb_smem(dst)), "l"(src), "r"(bytes), "r"(jb_smem(bar)) : "memory");
}
__device__ __forceinline__ void jb_wait(u64 *bar, u32 parity) {
    u32 ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(jb_smem(bar)), "r"(parity) : "memory");
    } while (!ok);
}

// one Huffman symbol with the block's tables in global memory (the walker's code-by-code path); -1 = malformed
__device__ __forceinline__ int decode_one_g(BitBuf &br, const DecTables *g, int t, int alpha) {
    br.refill();
    u16 e = g->lut[t][br.peek(LUTBITS)];
    if (e) { br.skip(e >> 9); return (int)(e & 511u); }
    int l = LUTBITS + 1;
    int code = (int)br.peek(l);
    while (l <= 20 && code > g->limit[t][l]) { l++; code = (int)br.peek(l); }
    if (l > 20) return -1;
    int pi = code + g->base[t][l];
    if (pi < 0 || pi >= alpha) return -1;
    br.skip(l);
    return (int)g->perm[t][pi];
}

__global__ void __launch_bounds__(32) k_dec_bounds(const u8 *in, size_t n, const u8 *sel_all, u32 sel_stride,
                                                   DecTables *tabs, const u64 *rend, const u8 *jt, size_t jstride,
                                                   const u32 *flags, u32 fstride, u32 *gbit_all, u32 max_sym, u32 b0) {
    const u32 b = b0 + blockIdx.x;
    extern __shared__ __align__(128) u8 ring[];                 // [table][J8 | J2][JRING * JW]
    __shared__ __align__(8) u64 bar[JRING];
    DecTables *tb = tabs + b;
    if (tb->status || threadIdx.x != 0) return;                 // one lane walks; the copies land in the CTA's shared memory
    const int alpha = (int)tb->alpha, T = (int)tb->T;
    const u32 G = tb->G;
    const u64 data_bit = tb->data_bit;
    const u64 base_bit = (data_bit >> 5) << 5;
    const u64 end_bit = min(rend[b], (u64)n * 8);
    const u64 R = end_bit > base_bit ? end_bit - base_bit : 0;
    const u32 nwin = jump_windows(R);
    const u32 win_bytes = (u32)T * 2u * JW;
    const u8 *jb = jt + (size_t)blockIdx.x * jstride;
    const u8 *sel = sel_all + (size_t)b * sel_stride;
    u32 *gbit = gbit_all + (size_t)b * sel_stride;
    for (int s = 0; s < JRING; s++) jb_init(&bar[s]);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    u32 loaded = 0, ready = 0;                                  // windows [0, loaded) are posted, [0, ready) have arrived
    const u32 *fl = flags + (size_t)blockIdx.x * fstride;
    bool timeout = false;
    auto post = [&](u32 wi) {
        // the producer runs concurrently (other stream), normally well ahead; the wait is bounded so that a producer
        // that never ran (launch failure) cannot hang the device
        u32 f = 0;
        for (long long t0 = clock64();;) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(f) : "l"(fl + wi) : "memory");
            if (f || clock64() - t0 > (8ll << 30)) break;
            __nanosleep(200);
        }
        if (!f) { timeout = true; return; }
        asm volatile("fence.proxy.async;" ::: "memory");        // the window was written through the generic proxy
        const u32 slot = wi % JRING;
        jb_expect(&bar[slot], win_bytes);
        const u8 *src = jb + (size_t)wi * win_bytes;
        for (int q = 0; q < 2 * T; q++) jb_copy(ring + ((size_t)q * JRING + slot) * JW, src + (size_t)q * JW, JW, &bar[slot]);
    };
    while (loaded < JRING && loaded < nwin && !timeout) { post(loaded); if (!timeout) loaded++; }
    const u32 p0 = (u32)(data_bit - base_bit);
    u64 p = p0;                                                 // bit offset from base_bit
    u32 g = 0, nsym = 0, status = timeout ? 9u : 0u;
    int rem = 0;                                                // codes of the current group still to do (code-by-code path)
    bool special = false;
    const u32 mask = JRING * JW - 1;
    u32 tnext = G ? sel[0] : 0;
    while (!special && status == 0) {
        if (g >= G) { status = 6; break; }                      // ran out of selectors before EOB
        const u32 wi = (u32)(p / JW);
        if (wi + 1 >= nwin) { special = true; rem = 0; break; } // cannot happen below R (see jump_windows); past it: code by code
        while (loaded < wi + JRING && loaded < nwin && !timeout) {   // windows below wi are done with: their slots take the next ones
            post(loaded);
            if (!timeout) loaded++;
        }
        if (timeout) { status = 9; break; }
        while (ready < wi + 2) { jb_wait(&bar[ready % JRING], (ready / JRING) & 1u); ready++; }
        const u32 t = tnext;
        gbit[g] = (u32)(p - p0);
        g++;
        if (g < G) tnext = sel[g];
        if (t >= (u32)T) { status = 7; break; }
        const u8 *r8 = ring + (size_t)(2 * t) * JRING * JW, *r2 = r8 + (size_t)JRING * JW;
        u32 q = (u32)p;                                         // ranges are far below 2^32 bits
        int left = 50;
#pragma unroll
        for (int j = 0; j < 6; j++) {
            u32 d = r8[q & mask];
            if (d == 0) { special = true; break; }
            q += d; left -= 8;
        }
        if (!special) {
            u32 d = r2[q & mask];
            if (d == 0) special = true; else { q += d; left -= 2; }
        }
        p = q;
        if (special) { rem = left; break; }
        nsym += 50;
        if (nsym > max_sym) { status = 7; break; }
    }
    u64 endpos = base_bit + p;
    if (special && status == 0) {                               // code by code from p: `rem` codes of group g - 1 are left (0: a new group starts)
        BitBuf br;
        br.init(in, n, base_bit + p);
        int t = g ? (int)sel[g - 1] : 0;
        nsym += (u32)(rem ? 50 - rem : 0);
        bool done = false;
        while (!done) {
            if (rem == 0) {
                if (g >= G) { status = 6; break; }
                t = sel[g];
                if (t >= T) { status = 7; break; }
                gbit[g] = (u32)(br.bitpos() - data_bit);
                g++;
                rem = 50;
            }
            int s = decode_one_g(br, tb, t, alpha);
            if (s < 0) { status = 7; break; }
            rem--; nsym++;
            if (s == alpha - 1) done = true;
            if (nsym > max_sym || br.bitpos() > (u64)n * 8 + 64) { status = 7; break; }
        }
        endpos = br.bitpos();
    }
    while (ready < loaded) { jb_wait(&bar[ready % JRING], (ready / JRING) & 1u); ready++; }   // no copy may outlive the CTA
    tb->status = status; tb->end_bit = endpos; tb->nsym = nsym; tb->ngroups = g;
}
