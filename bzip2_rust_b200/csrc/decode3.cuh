// decode3.cuh -- third version of the per-block entropy decoder (included by decode.cu after decode2.cuh).
//
// k_dec_block2 spent most of its time shifting the MTF list one shared-memory byte at a time on a single lane.
// Here the warp splits the work: lane 0 Huffman-decodes a batch of up to 64 symbols (the only inherently serial
// part: the table in force depends on the symbol index), then all 32 lanes replay the batch -- the MTF list lives
// in registers across the warp (positions 0..31 one per lane, 32..255 as 7 bytes per lane, exactly the layout of
// the encoder's k_mtf_emit2), zero runs are filled by the whole warp, and decoded bytes leave as coalesced stores.

__global__ void __launch_bounds__(32) k_dec_block3(const u8 *in, size_t n, const u64 *start_bits, u32 max_block,
                                                   u8 *tt_all, u32 stride, u8 *sel_all, u32 sel_stride, DecBlock *out) {
    u32 b = blockIdx.x;
    __shared__ u8 len[6][258];
    __shared__ u16 perm[6][258];
    __shared__ int limit[6][22], base[6][22];
    __shared__ u16 lut[6][1 << LUTBITS];      // (code length << 9) | symbol; 0 = longer than LUTBITS
    __shared__ __align__(8) u8 seq[256];
    __shared__ u16 sbuf[64];
    __shared__ int s_T, s_alpha, s_status, s_cnt, s_done;
    __shared__ u32 s_G;
    const int lane = threadIdx.x;
    DecBlock r; r.status = 0; r.nblock = 0; r.end_bit = 0; r.crc = 0; r.key = 0;
    u8 *tt = tt_all + (size_t)b * stride;
    u8 *sel = sel_all + (size_t)b * sel_stride;
    BitBuf br;
    for (int i = lane; i < 256; i += 32) seq[i] = 0;
    __syncwarp();
    if (lane == 0) {
        s_status = 0;
        br.init(in, n, start_bits[b] + 48);
        { u32 hi16 = br.get(16); u32 lo16 = br.get(16); r.crc = (hi16 << 16) | lo16; }
        if (br.get(1)) s_status = 1;
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
        if (!s_status) {
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
        for (int t = 0; t < T && !s_status; t++) {
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
        s_T = T; s_alpha = alpha; s_G = G;
    }
    __syncwarp();
    const int T = s_T, alpha = s_alpha;
    if (s_status == 0) {
        if (lane < T) {
            int t = lane;
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
        for (int i = lane; i < 6 * (1 << LUTBITS); i += 32) (&lut[0][0])[i] = 0;
        __syncwarp();
        for (int t = 0; t < T; t++) {
            for (int pi = lane; pi < alpha; pi += 32) {
                int s = perm[t][pi];
                int l = len[t][s];
                if (l <= LUTBITS) {
                    int code = pi - base[t][l];
                    int lo = code << (LUTBITS - l), hi = lo + (1 << (LUTBITS - l));
                    u16 e = (u16)((l << 9) | s);
                    for (int k = lo; k < hi; k++) lut[t][k] = e;
                }
            }
        }
    }
    __syncwarp();
    if (s_status) { if (lane == 0) { r.status = (u32)s_status; out[b] = r; } return; }

    // MTF list across the warp
    u32 fw = seq[lane];
    u64 tl = 0;
#pragma unroll
    for (int k = 0; k < 7; k++) tl |= (u64)seq[32 + 7 * lane + k] << (8 * k);
    const u64 LOW7 = 0x00FFFFFFFFFFFFFFull;

    const u32 G = s_G;
    u32 g = 0, gpos = 50; int t = 0;          // lane 0 only
    u32 nblk = 0, runlen = 0, runbit = 1;     // warp uniform
    u32 stage = 0;                            // byte staged by lane (nblk & 31)
    u32 status = 0;
    bool done = false;
    while (!done && status == 0) {
        // ---- lane 0: Huffman-decode a batch ----
        if (lane == 0) {
            int cnt = 0; int fin = 0; int st = 0;
            while (cnt < 64) {
                if (gpos == 50) { if (g >= G) { st = 6; break; } t = sel[g++]; gpos = 0; }
                gpos++;
                br.refill();
                u32 s;
                u16 e = lut[t][br.peek(LUTBITS)];
                if (e) { s = e & 511u; br.skip(e >> 9); }
                else {
                    int l = LUTBITS + 1;
                    int code = (int)br.peek(l);
                    while (l <= 20 && code > limit[t][l]) { l++; code = (int)br.peek(l); }
                    if (l > 20) { st = 7; break; }
                    int pi = code + base[t][l];
                    if (pi < 0 || pi >= alpha) { st = 8; break; }
                    s = perm[t][pi];
                    br.skip(l);
                }
                if (br.byte > n + 16) { st = 7; break; }
                sbuf[cnt++] = (u16)s;
                if ((int)s == alpha - 1) { fin = 1; break; }
            }
            s_cnt = cnt; s_done = fin; if (st) s_status = st;
        }
        __syncwarp();
        int cnt = s_cnt;
        done = s_done != 0;
        status = (u32)s_status;
        // ---- all lanes: replay the batch ----
        for (int q = 0; q < cnt; q++) {
            u32 s = sbuf[q];
            if (s <= 1) { runlen += runbit << s; runbit <<= 1; if (runlen > max_block) { status = 9; break; } continue; }
            if (runlen) {
                if (nblk + runlen > max_block) { status = 9; break; }
                u32 c = __shfl_sync(0xffffffffu, fw, 0);
                // finish the partially staged 32-byte line, then whole lines, then stage the remainder
                u32 head = min(runlen, (32u - (nblk & 31u)) & 31u);
                if (head) {
                    u32 slot = nblk & 31u;
                    if ((u32)lane >= slot && (u32)lane < slot + head) stage = c;
                    nblk += head; runlen -= head;
                    if ((nblk & 31u) == 0) tt[nblk - 32 + lane] = (u8)stage;
                }
                while (runlen >= 32) { tt[nblk + lane] = (u8)c; nblk += 32; runlen -= 32; }
                if (runlen) { if ((u32)lane < runlen) stage = c; nblk += runlen; runlen = 0; }
            }
            runbit = 1;
            if ((int)s == alpha - 1) break;                    // EOB
            u32 pos = s - 1;
            u32 v;
            u32 up = __shfl_up_sync(0xffffffffu, fw, 1);
            if (pos < 32) {
                v = __shfl_sync(0xffffffffu, fw, (int)pos);
                if ((u32)lane <= pos) fw = lane ? up : v;
            } else {
                int Lh = (int)((pos - 32) / 7), kb = (int)((pos - 32) % 7);
                u32 mine = (u32)(tl >> (8 * kb)) & 0xffu;
                v = __shfl_sync(0xffffffffu, mine, Lh);
                u32 carry = __shfl_sync(0xffffffffu, fw, 31);
                u32 top = (u32)(tl >> 48) & 0xffu;
                u32 incoming = __shfl_up_sync(0xffffffffu, top, 1);
                if (lane == 0) incoming = carry;
                if (lane < Lh) tl = ((tl << 8) | incoming) & LOW7;
                else if (lane == Lh) {
                    u64 lowmask = kb ? ((1ull << (8 * kb)) - 1) : 0ull;
                    u64 keepmask = (~((1ull << (8 * (kb + 1))) - 1)) & LOW7;
                    tl = (tl & keepmask) | (((tl & lowmask) << 8) | incoming);
                }
                fw = lane ? up : v;
            }
            if (nblk + 1 > max_block) { status = 9; break; }
            if ((nblk & 31u) == (u32)lane) stage = v;
            nblk++;
            if ((nblk & 31u) == 0) tt[nblk - 32 + lane] = (u8)stage;
        }
        __syncwarp();
    }
    // flush the partially staged line
    if ((nblk & 31u) && (u32)lane < (nblk & 31u)) tt[(nblk & ~31u) + lane] = (u8)stage;
    if (lane == 0) {
        r.status = status;
        r.nblock = nblk;
        r.end_bit = br.bitpos();
        if (r.status == 0 && (nblk == 0 || r.key >= nblk)) r.status = 10;
        out[b] = r;
    }
}
