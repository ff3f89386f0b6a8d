// decode2.cuh -- second versions of the two slow decode kernels (included by decode.cu inside its namespace).
//
// k_dec_block2: same job as k_dec_block (decompress.rs:98-358 for one block) but with a 64-bit bit buffer refilled
//   four bytes at a time and a 10-bit lookup table per Huffman table (one probe decodes any code of <= 10 bits;
//   longer codes fall back to the canonical limit/base walk), and word-wide stores of the decoded BWT string.
//   Huffman decoding of a bzip2 block cannot be split: the table in force depends on the symbol index (50 per
//   selector), so one lane walks a block and the parallelism is across blocks.
// k_rle1_inv: inverse RLE1 (rle1.rs:267-316 semantics, but exact at block tails) in parallel.  The sequential rule
//   "the byte after four equal bytes is a repeat count" is a 5-state automaton driven only by eq[i] = (r[i]==r[i-1]):
//   state = run length so far (0 = just consumed a count).  Transition functions compose associatively, so a
//   scan over 15-bit packed functions gives every byte's state; a second scan of the output sizes gives offsets.
//   One CTA walks one block tile by tile (carry in registers), blocks run concurrently.

constexpr int LUTBITS = 10;

struct BitBuf {
    const u8 *p; size_t n; size_t byte;      // next byte to load
    u64 buf; int nb;                         // `nb` valid bits, MSB aligned
    __device__ void init(const u8 *in, size_t len, u64 bitpos) {
        p = in; n = len; byte = (size_t)(bitpos >> 3); buf = 0; nb = 0;
        refill();
        int sk = (int)(bitpos & 7);
        buf <<= sk; nb -= sk;
    }
    __device__ __forceinline__ void refill() {
        while (nb <= 32) {
            u32 w;
            if (byte + 4 <= n) {
                w = ((u32)p[byte] << 24) | ((u32)p[byte + 1] << 16) | ((u32)p[byte + 2] << 8) | (u32)p[byte + 3];
            } else {
                w = 0;
                for (int k = 0; k < 4; k++) w = (w << 8) | (byte + k < n ? p[byte + k] : 0);
            }
            byte += 4;
            buf |= (u64)w << (32 - nb);
            nb += 32;
        }
    }
    __device__ __forceinline__ u32 peek(int k) const { return (u32)(buf >> (64 - k)); }     // 1 <= k <= 32
    __device__ __forceinline__ void skip(int k) { buf <<= k; nb -= k; }
    __device__ __forceinline__ u32 get(int k) { if (k == 0) return 0; refill(); u32 v = peek(k); skip(k); return v; }
    __device__ u64 bitpos() const { return (u64)byte * 8 - (u64)nb; }
};

__global__ void __launch_bounds__(32) k_dec_block2(const u8 *in, size_t n, const u64 *start_bits, u32 max_block,
                                                   u8 *tt_all, u32 stride, u8 *sel_all, u32 sel_stride, DecBlock *out) {
    u32 b = blockIdx.x;
    __shared__ u8 len[6][258];
    __shared__ u16 perm[6][258];
    __shared__ int limit[6][22], base[6][22];
    __shared__ u16 lut[6][1 << LUTBITS];      // (code length << 9) | symbol; 0 = longer than LUTBITS
    __shared__ u8 seq[256];
    __shared__ int s_T, s_alpha, s_status;
    __shared__ u32 s_G;
    __shared__ u64 s_pos;
    DecBlock r; r.status = 0; r.nblock = 0; r.end_bit = 0; r.crc = 0; r.key = 0;
    u8 *tt = tt_all + (size_t)b * stride;
    u8 *sel = sel_all + (size_t)b * sel_stride;
    BitBuf br;
    if (threadIdx.x == 0) {
        s_status = 0;
        br.init(in, n, start_bits[b] + 48);
        { u32 hi16 = br.get(16); u32 lo16 = br.get(16); r.crc = (hi16 << 16) | lo16; }
        if (br.get(1)) s_status = 1;                                  // randomised blocks: not produced by this encoder
        r.key = br.get(24);
        u32 l1 = br.get(16);
        int nused = 0;
        for (int i = 0; i < 16; i++) if (l1 & (0x8000u >> i)) {
            u32 l2 = br.get(16);
            for (int j = 0; j < 16; j++) if (l2 & (0x8000u >> j)) seq[nused++] = (u8)(i * 16 + j);
        }
        if (nused == 0 && !s_status) s_status = 2;
        int alpha = nused + 2;
        int T = (int)br.get(3);
        u32 G = br.get(15);
        if (!s_status && (T < 2 || T > 6 || G < 1 || G > sel_stride)) s_status = 3;
        if (!s_status) {   // selectors (decompress.rs:140-203)
            u8 l6[6] = {0, 1, 2, 3, 4, 5};
            for (u32 g = 0; g < G && !s_status; g++) {
                int j = 0;
                while (br.get(1)) { j++; if (j >= T) { s_status = 4; break; } }
                if (s_status) break;
                u8 v = l6[j];
                for (int k = j; k > 0; k--) l6[k] = l6[k - 1];
                l6[0] = v;
                sel[g] = v;
            }
        }
        for (int t = 0; t < T && !s_status; t++) {        // code lengths (decompress.rs:216-260)
            int c = (int)br.get(5);
            for (int s = 0; s < alpha && !s_status; s++) {
                for (;;) {
                    if (c < 1 || c > 20) { s_status = 5; break; }
                    if (!br.get(1)) break;
                    c += br.get(1) ? -1 : 1;
                }
                len[t][s] = (u8)c;
            }
        }
        s_T = T; s_alpha = alpha; s_G = G; s_pos = br.bitpos();
    }
    __syncthreads();
    int T = s_T, alpha = s_alpha;
    if (s_status == 0) {
        // decode tables, built by the whole warp: lane = table for the canonical part, all lanes fill the LUTs
        if ((int)threadIdx.x < T) {
            int t = threadIdx.x;
            int mn = 32, mx = 0;
            for (int s = 0; s < alpha; s++) { int l = len[t][s]; mn = min(mn, l); mx = max(mx, l); }
            int pp = 0;
            for (int l = mn; l <= mx; l++) for (int s = 0; s < alpha; s++) if (len[t][s] == l) perm[t][pp++] = (u16)s;
            int cnt[22];
            for (int l = 0; l < 22; l++) cnt[l] = 0;
            for (int s = 0; s < alpha; s++) cnt[len[t][s]]++;
            int code = 0, idx = 0;
            for (int l = 1; l <= 20; l++) {
                base[t][l] = idx - code; code += cnt[l]; idx += cnt[l]; limit[t][l] = code - 1; code <<= 1;
            }
            for (int l = 1; l <= 20; l++) if (l > mx) limit[t][l] = 0x7fffffff;
        }
        for (int i = threadIdx.x; i < 6 * (1 << LUTBITS); i += 32) (&lut[0][0])[i] = 0;
        __syncwarp();
        // canonical code of symbol s in table t = limit/base arithmetic inverted: walk the sorted order
        for (int t = 0; t < T; t++) {
            // lane-parallel over sorted symbol positions
            for (int pi = threadIdx.x; pi < alpha; pi += 32) {
                int s = perm[t][pi];
                int l = len[t][s];
                if (l <= LUTBITS) {
                    int code = pi - base[t][l];                 // canonical code value of this symbol
                    int lo = code << (LUTBITS - l), hi = lo + (1 << (LUTBITS - l));
                    u16 e = (u16)((l << 9) | s);
                    for (int k = lo; k < hi; k++) lut[t][k] = e;
                }
            }
        }
    }
    __syncwarp();
    if (threadIdx.x != 0) return;
    if (s_status) { r.status = (u32)s_status; out[b] = r; return; }
    // ---- Huffman decode + inverse MTF/RLE2 (decompress.rs:293-358, rle2_mtf.rs:191-287) ----
    u32 G = s_G;
    u32 nblk = 0, runlen = 0, runbit = 1, g = 0, gpos = 50;
    int t = 0;
    u32 wacc = 0; int wn = 0;                 // up to 4 pending output bytes (tt is 4-byte aligned)
    u32 *tt32 = (u32 *)tt;
#define PUT_BYTE(V)                                                        \
    do {                                                                   \
        wacc |= (u32)(V) << (8 * wn);                                      \
        if (++wn == 4) { tt32[nblk >> 2] = wacc; wacc = 0; wn = 0; }      \
        nblk++;                                                            \
    } while (0)
    for (;;) {
        if (gpos == 50) { if (g >= G) { r.status = 6; break; } t = sel[g++]; gpos = 0; }
        gpos++;
        br.refill();
        u32 s;
        u16 e = lut[t][br.peek(LUTBITS)];
        if (e) { s = e & 511u; br.skip(e >> 9); }
        else {
            int l = LUTBITS + 1;
            int code = (int)br.peek(l);
            while (l <= 20 && code > limit[t][l]) { l++; code = (int)br.peek(l); }
            if (l > 20) { r.status = 7; break; }
            int pi = code + base[t][l];
            if (pi < 0 || pi >= alpha) { r.status = 8; break; }
            s = perm[t][pi];
            br.skip(l);
        }
        if ((br.byte >> 0) > n + 16) { r.status = 7; break; }
        if (s <= 1) { runlen += runbit << s; runbit <<= 1; if (runlen > max_block) { r.status = 9; break; } continue; }
        if (runlen) {
            if (nblk + runlen > max_block) { r.status = 9; break; }
            u32 c = seq[0];
            while (runlen && wn) { PUT_BYTE(c); runlen--; }
            u32 c4 = c * 0x01010101u;
            while (runlen >= 4) { tt32[nblk >> 2] = c4; nblk += 4; runlen -= 4; }
            while (runlen) { PUT_BYTE(c); runlen--; }
        }
        runbit = 1;
        if ((int)s == alpha - 1) break;                            // EOB
        u32 pos = s - 1;
        u8 v = seq[pos];
        for (u32 k = pos; k > 0; k--) seq[k] = seq[k - 1];
        seq[0] = v;
        if (nblk + 1 > max_block) { r.status = 9; break; }
        PUT_BYTE(v);
    }
    if (wn) { for (int k = 0; k < wn; k++) tt[(nblk - wn) + k] = (u8)(wacc >> (8 * k)); }
#undef PUT_BYTE
    r.nblock = nblk;
    r.end_bit = br.bitpos();
    if (r.status == 0 && (nblk == 0 || r.key >= nblk)) r.status = 10;
    out[b] = r;
}

// ---- parallel inverse RLE1 ------------------------------------------------------------------------------
// packed transition function: 5 states x 3 bits; F(s) = (f >> (3 s)) & 7
__device__ __forceinline__ u32 fn_compose(u32 f, u32 g) {       // apply f, then g
    u32 r = 0;
#pragma unroll
    for (int s = 0; s < 5; s++) { u32 m = (f >> (3 * s)) & 7u; r |= ((g >> (3 * m)) & 7u) << (3 * s); }
    return r;
}
constexpr u32 FN_ID = (0u) | (1u << 3) | (2u << 6) | (3u << 9) | (4u << 12);
constexpr u32 FN_EQ = (1u) | (2u << 3) | (3u << 6) | (4u << 9) | (0u << 12);    // 0->1 1->2 2->3 3->4 4->0(count)
constexpr u32 FN_NE = (1u) | (1u << 3) | (1u << 6) | (1u << 9) | (0u << 12);    // 0..3->1 4->0(count)

// WRITE = 0: outlen[b] = decoded length.  WRITE = 1: bytes written at out + outoff[b].
template <int WRITE>
__global__ void __launch_bounds__(256) k_rle1_inv(const u8 *blk, const u32 *len, u32 stride, u64 *outlen, const u64 *outoff,
                                                  u8 *out) {
    u32 b = blockIdx.x;
    const u8 *r = blk + (size_t)b * stride;
    u32 n = len[b];
    u8 *o = WRITE ? out + outoff[b] : nullptr;
    __shared__ u32 wfn[8];
    __shared__ u32 wsum[8];
    __shared__ u32 s_state;      // automaton state after the previous tile
    __shared__ u64 s_off;        // output bytes before this tile
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) { s_state = 0; s_off = 0; }
    __syncthreads();
    constexpr int PT = 8;                                  // bytes per thread
    for (u32 base = 0; base < n; base += 256 * PT) {
        u32 i0 = base + threadIdx.x * PT;
        u8 c[PT + 1];
        c[0] = (i0 > 0 && i0 - 1 < n) ? r[i0 - 1] : 0;
#pragma unroll
        for (int k = 0; k < PT; k++) c[k + 1] = (i0 + k < n) ? r[i0 + k] : 0;
        // thread function over its bytes
        u32 f = FN_ID;
#pragma unroll
        for (int k = 0; k < PT; k++) {
            if (i0 + k < n) {
                bool eq = (i0 + k > 0) && c[k + 1] == c[k];
                f = fn_compose(f, eq ? FN_EQ : FN_NE);
            }
        }
        // exclusive scan of functions across the CTA
        u32 inc = f;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u32 tpre = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc = fn_compose(tpre, inc);
        }
        if (lane == 31) wfn[w] = inc;
        __syncthreads();
        u32 pre = FN_ID;
        for (int k = 0; k < w; k++) pre = fn_compose(pre, wfn[k]);
        u32 excl = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) excl = FN_ID;
        u32 before = fn_compose(pre, excl);                // function from tile start to just before this thread
        u32 st = (before >> (3 * s_state)) & 7u;           // state before this thread's first byte
        // walk own bytes: classify, size
        u32 sz = 0;
        u32 st0 = st;
#pragma unroll
        for (int k = 0; k < PT; k++) {
            if (i0 + k < n) {
                bool eq = (i0 + k > 0) && c[k + 1] == c[k];
                if (st == 4) { sz += c[k + 1]; st = 0; }   // repeat count
                else { sz += 1; st = (st >= 1 && eq) ? st + 1 : 1; }
            }
        }
        // exclusive sum of sizes
        u32 sinc = sz;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { u32 tv = __shfl_up_sync(0xffffffffu, sinc, d); if (lane >= d) sinc += tv; }
        if (lane == 31) wsum[w] = sinc;
        __syncthreads();
        u32 wpre = 0, tot = 0;
        for (int k = 0; k < 8; k++) { if (k < w) wpre += wsum[k]; tot += wsum[k]; }
        u64 off = s_off + wpre + (sinc - sz);
        if (WRITE) {
            st = st0;
#pragma unroll
            for (int k = 0; k < PT; k++) {
                if (i0 + k < n) {
                    bool eq = (i0 + k > 0) && c[k + 1] == c[k];
                    if (st == 4) {
                        u32 cntv = c[k + 1]; u8 v = c[k];
                        // v is the run byte only if the previous byte was a run byte, which it is when st == 4
                        for (u32 q = 0; q < cntv; q++) o[off + q] = v;
                        off += cntv; st = 0;
                    } else { o[off++] = c[k + 1]; st = (st >= 1 && eq) ? st + 1 : 1; }
                }
            }
        }
        // carry to the next tile: state after the tile's last byte, total size
        u32 tile_fn = FN_ID;
        for (int k = 0; k < 8; k++) tile_fn = fn_compose(tile_fn, wfn[k]);
        __syncthreads();
        if (threadIdx.x == 0) { s_state = (tile_fn >> (3 * s_state)) & 7u; s_off += tot; }
        __syncthreads();
    }
    if (!WRITE && threadIdx.x == 0) outlen[b] = s_off;
}
