// decode.cu -- block-parallel bzip2 decoder (the round-trip side of the path).
//
// Replaces decompress (reference src/compression/decompress.rs:38-404), huf_decode_map (:426-486),
// rle2_mtf_decode_fast (src/tools/rle2_mtf.rs:191-287), decode_sym_map (symbol_map.rs:20-42),
// bwt_decode (src/bwt_algorithms/bwt_sort.rs:91-130) and rle1_decode (src/tools/rle1.rs:267-316).
// The reference decodes one block after another on one thread; blocks are independent once their
// start bits are known, so:
//   k_dec_find_magic  every bit offset is tested for the 48-bit block / footer magic
//   decode4.cuh       entropy decode, parallel inside a block: header -> tables, a lengths-only walk for the
//                     bit offset of every 50-symbol group, one thread per group for the symbols, inverse MTF +
//                     RUNA/RUNB by chunks whose start lists come from composing per-chunk permutations
//   inverse BWT       stable counting sort of rows by byte = one pass of the BWT stage's radix kernels
//                     (P[j] = row of the j-th smallest byte), then the n-step pointer chase of
//                     bwt_decode is cut into ~n/256 segments at splitter rows: k_ibwt_chase measures
//                     every segment in parallel, k_ibwt_rank orders them from the origin pointer,
//                     k_ibwt_write replays each segment at its output offset
//   k_dec_rle1_*      inverse RLE1 (count pass, then write pass at the block's output offset)
//   CRC               block CRCs with the encoder's CRC kernels; combined CRC on the host.
// The standard run semantics are used (a count byte follows every 4 equal bytes); the reference's
// rle1_decode drops run expansion in a block's last 5 bytes -- libbz2 is the authority.
#include "common.cuh"
#include "radix.cuh"
#include <algorithm>
#include <stddef.h>
#include <string.h>

int bz_crc_spans_dev(bz2b200_ctx *ctx, const u8 *d_x, const u32 *d_se /* [nb][2] */, u32 nb, u32 max_span,
                     u32 *d_crc_out);

namespace {

constexpr u64 MAGIC_BLOCK = 0x314159265359ull;
constexpr u64 MAGIC_END = 0x177245385090ull;
constexpr int SPLIT = 256;               // inverse BWT: one splitter row every SPLIT rows

__device__ __forceinline__ u64 load_be64(const u8 *p, size_t n, size_t byte) {
    u64 v = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) v = (v << 8) | (byte + k < n ? p[byte + k] : 0);
    return v;
}

// candidates[*ncand] = bit positions where a block or footer magic starts (bit 63 set = footer)
__global__ void __launch_bounds__(256) k_dec_find_magic(const u8 *in, size_t n, u64 *cand, u32 *ncand, u32 cap) {
    size_t byte = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (byte >= n) return;
    u64 hi = load_be64(in, n, byte);
    u64 lo = load_be64(in, n, byte + 8);
#pragma unroll
    for (int sh = 0; sh < 8; sh++) {
        u64 w = sh ? ((hi << sh) | (lo >> (64 - sh))) : hi;
        u64 m = w >> 16;
        if (m == MAGIC_BLOCK || m == MAGIC_END) {
            u32 k = atomicAdd(ncand, 1u);
            if (k < cap) cand[k] = ((u64)byte * 8 + sh) | (m == MAGIC_END ? (1ull << 63) : 0);
        }
    }
}

// ---- inverse BWT ------------------------------------------------------------------------------
__device__ __forceinline__ bool is_split(u32 row, u32 key) { return (row % SPLIT) == 0 || row == key; }
__device__ __forceinline__ u32 split_index(u32 row, u32 key, u32 nreg) { return (row % SPLIT) == 0 ? row / SPLIT : nreg; }

// thread per splitter: length of its segment and the splitter that ends it
__global__ void __launch_bounds__(256) k_ibwt_chase(const u32 *P, const u32 *len, const u32 *keys, u32 stride,
                                                    u32 *seglen, u32 *segnext, u32 sstride) {
    u32 b = blockIdx.y, n = len[b], key = keys[b];
    u32 nreg = (n + SPLIT - 1) / SPLIT;
    u32 j = blockIdx.x * 256 + threadIdx.x;
    if (j > nreg) return;
    u32 row = j < nreg ? j * SPLIT : key;
    if (j == nreg && (key % SPLIT) == 0) { seglen[(size_t)b * sstride + j] = 0; return; }   // origin is a regular splitter
    const u32 *p = P + (size_t)b * stride;
    u32 steps = 0;
    do { row = p[row]; steps++; } while (!is_split(row, key));
    seglen[(size_t)b * sstride + j] = steps;
    segnext[(size_t)b * sstride + j] = split_index(row, key, nreg);
}

// one thread per block: visit splitters in text order starting at the origin pointer until n rows are covered
__global__ void __launch_bounds__(32) k_ibwt_rank(const u32 *len, const u32 *keys, const u32 *seglen, const u32 *segnext,
                                                  u32 sstride, u32 *visit_split, u32 *visit_off, u32 vstride, u32 *nvisit) {
    u32 b = blockIdx.x;
    if (threadIdx.x != 0) return;
    u32 n = len[b], key = keys[b];
    u32 nreg = (n + SPLIT - 1) / SPLIT;
    u32 j = (key % SPLIT) == 0 ? key / SPLIT : nreg;
    u32 off = 0, v = 0;
    while (off < n && v < vstride) {
        visit_split[(size_t)b * vstride + v] = j;
        visit_off[(size_t)b * vstride + v] = off;
        off += seglen[(size_t)b * sstride + j];
        j = segnext[(size_t)b * sstride + j];
        v++;
    }
    nvisit[b] = v;
}

// thread per visit: replay the segment; text position i (0-based) receives L[P^(i+1)(key)]  (bwt_sort.rs:118-128)
__global__ void __launch_bounds__(256) k_ibwt_write(const u32 *P, const u8 *L, const u32 *len, const u32 *keys, u32 stride,
                                                    const u32 *visit_split, const u32 *visit_off, u32 vstride,
                                                    const u32 *nvisit, u8 *out) {
    u32 b = blockIdx.y, n = len[b], key = keys[b];
    u32 v = blockIdx.x * 256 + threadIdx.x;
    if (v >= nvisit[b]) return;
    u32 nreg = (n + SPLIT - 1) / SPLIT;
    u32 j = visit_split[(size_t)b * vstride + v];
    u32 off = visit_off[(size_t)b * vstride + v];
    u32 row = j < nreg ? j * SPLIT : key;
    const u32 *p = P + (size_t)b * stride;
    const u8 *l = L + (size_t)b * stride;
    u8 *o = out + (size_t)b * stride;
    // row = P^off(key); position off-1 holds L[row] for off >= 1 and position n-1 holds L[key]
    do {
        u32 pos = off == 0 ? n - 1 : off - 1;
        if (off < n) o[pos] = l[row];
        row = p[row];
        off++;
    } while (!is_split(row, key) && off < n);
}

#include "decode4.cuh"       // entropy decode: header, group boundaries, symbols, chunked inverse MTF
#include "decode_rle1.cuh"   // k_rle1_inv

}  // namespace

#define LAUNCH_OK()                                                \
    do {                                                             \
        ctx->prof_end();                                             \
        cudaError_t e_ = cudaGetLastError();                         \
        if (e_ != cudaSuccess) { ctx->fail("kernel launch", e_, __FILE__, __LINE__); return BZ2B200_E_CUDA; } \
    } while (0)

// Inverse BWT of a batch already on the device: L = last column, keys = origin pointers.
static int ibwt_batch(bz2b200_ctx *ctx, const Batch &B, const u32 *d_keys, u8 *d_out) {
    cudaStream_t st = ctx->stream;
    size_t ne = (size_t)B.nblk * B.stride;
    const u32 rtiles = B.stride / radix::R_TILE;
    BZ_CHECK(ctx->d_SA.ensure(ne * 4));
    BZ_CHECK(ctx->d_thist.ensure((size_t)B.nblk * rtiles * 256 * 4));
    u32 sstride = B.stride / SPLIT + 2;
    u32 vstride = B.stride + sstride;
    BZ_CHECK(ctx->d_dec2.ensure((size_t)B.nblk * sstride * 8 + (size_t)B.nblk * vstride * 8 + (size_t)B.nblk * 4 + 64));
    u32 *seglen = ctx->d_dec2.as<u32>();
    u32 *segnext = seglen + (size_t)B.nblk * sstride;
    u32 *vsplit = segnext + (size_t)B.nblk * sstride;
    u32 *voff = vsplit + (size_t)B.nblk * vstride;
    u32 *nvisit = voff + (size_t)B.nblk * vstride;
    u32 *P = ctx->d_SA.as<u32>();
    // P = rows sorted stably by their byte: one radix pass with digit = L[row]
    radix::RadixArgs a{};
    a.T = B.T; a.len = B.len; a.sa_out = P; a.thist = ctx->d_thist.as<u32>(); a.stride = B.stride; a.rtiles = rtiles;
    dim3 gr((B.max_n + radix::R_TILE - 1) / radix::R_TILE, B.nblk);
    ctx->prof_begin(K_RADIX_HIST0, B.total_n * 5); radix::k_radix_hist<<<gr, BZ_THREADS, 0, st>>>(a); LAUNCH_OK();
    ctx->prof_begin(K_RADIX_SCAN, 0); radix::k_radix_scan<<<B.nblk, 256, 0, st>>>(a.thist, B.len, rtiles); LAUNCH_OK();
    ctx->prof_begin(K_RADIX_SCATTER0, B.total_n * 9); radix::k_radix_scatter<<<gr, BZ_THREADS, 0, st>>>(a); LAUNCH_OK();
    dim3 gs((B.max_n / SPLIT + 2 + 255) / 256, B.nblk);
    ctx->prof_begin(K_IBWT_CHASE, B.total_n * 4); k_ibwt_chase<<<gs, 256, 0, st>>>(P, B.len, d_keys, B.stride, seglen, segnext, sstride); LAUNCH_OK();
    ctx->prof_begin(K_IBWT_RANK, 0); k_ibwt_rank<<<B.nblk, 32, 0, st>>>(B.len, d_keys, seglen, segnext, sstride, vsplit, voff, vstride, nvisit); LAUNCH_OK();
    dim3 gv((vstride + 255) / 256, B.nblk);
    ctx->prof_begin(K_IBWT_WRITE, B.total_n * 6); k_ibwt_write<<<gv, 256, 0, st>>>(P, B.T, B.len, d_keys, B.stride, vsplit, voff, vstride, nvisit, d_out); LAUNCH_OK();
    return BZ2B200_OK;
}

extern "C" int bz2b200_bwt_decode(bz2b200_ctx *ctx, uint32_t key, const uint8_t *bwt, uint32_t n, uint8_t *out) {
    BZ_API_TRY
    if (!ctx || !bwt || !out || n == 0 || key >= n) return BZ2B200_E_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (cudaSetDevice(ctx->device) != cudaSuccess) return BZ2B200_E_CUDA;
    Batch B;
    const u8 *ins[1] = {bwt};
    int rc = bz_stage_blocks(ctx, 1, ins, &n, B);
    if (rc) return rc;
    BZ_CHECK(ctx->d_key.ensure(4));
    BZ_CHECK(ctx->d_bwt.ensure(B.stride));
    BZ_CHECK(cudaMemcpyAsync(ctx->d_key.p, &key, 4, cudaMemcpyHostToDevice, ctx->stream));
    rc = ibwt_batch(ctx, B, ctx->d_key.as<u32>(), ctx->d_bwt.as<u8>());
    if (rc) return rc;
    BZ_CHECK(cudaMemcpyAsync(out, ctx->d_bwt.p, n, cudaMemcpyDeviceToHost, ctx->stream));
    BZ_CHECK(cudaStreamSynchronize(ctx->stream));
    return BZ2B200_OK;
    BZ_API_CATCH
}

extern "C" int bz2b200_decompress_stream(bz2b200_ctx *ctx, const uint8_t *in, size_t n, uint8_t *out, size_t out_cap,
                                         size_t *out_len) {
    BZ_API_TRY
    if (!ctx || !in || !out_len || (!out && out_cap) || n < 14) return BZ2B200_E_ARG;
    if (in[0] != 'B' || in[1] != 'Z' || in[2] != 'h' || in[3] < '1' || in[3] > '9') return BZ2B200_E_FORMAT;   // decompress.rs:46-62
    int level = in[3] - '0';
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (cudaSetDevice(ctx->device) != cudaSuccess) return BZ2B200_E_CUDA;
    cudaStream_t st = ctx->stream;
    BZ_CHECK(ctx->d_in.ensure(n + 64));
    BZ_CHECK(cudaMemcpyAsync(ctx->d_in.p, in, n, cudaMemcpyHostToDevice, st));
    // ---- 1. block starts ----
    u32 cap = (u32)(n / 32 + 64);
    BZ_CHECK(ctx->d_dec1.ensure((size_t)cap * 8 + 64));
    u64 *d_cand = ctx->d_dec1.as<u64>();
    u32 *d_ncand = (u32 *)(d_cand + cap);
    BZ_CHECK(cudaMemsetAsync(d_ncand, 0, 4, st));
    ctx->prof_begin(K_DEC_MAGIC, n); k_dec_find_magic<<<(u32)((n + 255) / 256), 256, 0, st>>>(ctx->d_in.as<u8>(), n, d_cand, d_ncand, cap); LAUNCH_OK();
    u32 ncand = 0;
    BZ_CHECK(cudaMemcpyAsync(&ncand, d_ncand, 4, cudaMemcpyDeviceToHost, st));
    BZ_CHECK(cudaStreamSynchronize(st));
    if (ncand == 0 || ncand > cap) return BZ2B200_E_FORMAT;
    std::vector<u64> cand(ncand);
    BZ_CHECK(cudaMemcpy(cand.data(), d_cand, (size_t)ncand * 8, cudaMemcpyDeviceToHost));
    std::sort(cand.begin(), cand.end(), [](u64 a, u64 b) { return (a & ~(1ull << 63)) < (b & ~(1ull << 63)); });
    if ((cand[0] & ~(1ull << 63)) != 32) return BZ2B200_E_FORMAT;
    // blocks = candidates up to the first footer magic that is followed by the end of the stream
    std::vector<u64> starts;
    u64 footer_bit = 0; bool have_footer = false;
    for (u64 c : cand) {
        u64 bit = c & ~(1ull << 63);
        if (c >> 63) { if ((bit + 80 + 7) / 8 == (u64)n) { footer_bit = bit; have_footer = true; break; } continue; }
        starts.push_back(bit);
    }
    if (!have_footer) return BZ2B200_E_FORMAT;
    u32 nb = (u32)starts.size();
    u32 stored_combined = 0;
    {
        size_t byte = (size_t)((footer_bit + 48) >> 3); int sh = (int)((footer_bit + 48) & 7);
        u64 w = 0;
        for (int k = 0; k < 5; k++) w = (w << 8) | (byte + k < n ? in[byte + k] : 0);
        stored_combined = (u32)((w >> (8 - sh)) & 0xffffffffull);
    }
    if (nb == 0) { *out_len = 0; return stored_combined == 0 ? BZ2B200_OK : BZ2B200_E_CRC; }
    // ---- 2. entropy decode (decode4.cuh) ----
    u32 max_block = (u32)level * 100000u;
    u32 stride = ((max_block + 64 + BZ_TILE - 1) / BZ_TILE) * BZ_TILE;
    u32 sel_stride = max_block / 50 + 8;
    u32 max_sym = max_block + 2;                                 // every symbol but EOB yields at least one byte
    u32 sym_stride = ((max_sym + 64 + DCH - 1) / DCH) * DCH;
    u32 ch_stride = sym_stride / DCH + 2;
    BZ_CHECK(ctx->d_T.ensure((size_t)nb * stride + 64));
    BZ_CHECK(ctx->d_bwt.ensure((size_t)nb * stride + 64));
    BZ_CHECK(ctx->d_sel.ensure((size_t)nb * sel_stride));
    BZ_CHECK(ctx->d_gbits.ensure((size_t)nb * sel_stride * 4));
    BZ_CHECK(ctx->d_sym.ensure((size_t)nb * sym_stride * 2));
    BZ_CHECK(ctx->d_mtfstate.ensure((size_t)nb * ch_stride * 256));
    BZ_CHECK(ctx->d_chunkrec.ensure((size_t)nb * ch_stride * 8));
    BZ_CHECK(ctx->d_hdr.ensure((size_t)nb * sizeof(DecTables)));
    BZ_CHECK(ctx->d_dec3.ensure((size_t)nb * (8 + 4 + 4 + 8 + 8 + 8) + 256));
    DecTables *d_tabs = ctx->d_hdr.as<DecTables>();
    u32 *d_gbit = ctx->d_gbits.as<u32>();
    u16 *d_sym = ctx->d_sym.as<u16>();
    u8 *d_lists = ctx->d_mtfstate.as<u8>();
    u32 *d_ccount = ctx->d_chunkrec.as<u32>();
    u32 *d_coff = d_ccount + (size_t)nb * ch_stride;
    u64 *d_starts = ctx->d_dec3.as<u64>();
    u32 *d_len = (u32 *)(d_starts + nb);
    const u32 nbp = (nb + 1) & ~1u;                              // keep the u64 arrays 8-byte aligned
    u32 *d_keys = d_len + nbp;
    u64 *d_olen = (u64 *)(d_keys + nbp);
    u64 *d_ooff = d_olen + nb;
    u32 *d_se = (u32 *)(d_ooff + nb);
    const u8 *d_bits = ctx->d_in.as<u8>();
    BZ_CHECK(cudaMemcpyAsync(d_starts, starts.data(), (size_t)nb * 8, cudaMemcpyHostToDevice, st));
    ctx->prof_begin(K_DEC_HEADER, 0); k_dec_header<<<nb, 32, 0, st>>>(d_bits, n, d_starts, ctx->d_sel.as<u8>(), sel_stride, d_tabs); LAUNCH_OK();
    if (!ctx->dec_attr_done) { BZ_CHECK(cudaFuncSetAttribute(k_dec_bounds, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BoundsSmem))); ctx->dec_attr_done = true; }
    ctx->prof_begin(K_DEC_BOUNDS, n); k_dec_bounds<<<nb, 32, sizeof(BoundsSmem), st>>>(d_bits, n, ctx->d_sel.as<u8>(), sel_stride, d_tabs, d_gbit, max_sym); LAUNCH_OK();
    dim3 gs((sel_stride + 127) / 128, nb);
    ctx->prof_begin(K_DEC_SYMS, n); k_dec_syms<<<gs, 128, 0, st>>>(d_bits, n, ctx->d_sel.as<u8>(), sel_stride, d_tabs, d_gbit, d_sym, sym_stride); LAUNCH_OK();
    dim3 gch((ch_stride + 7) / 8, nb);
    ctx->prof_begin(K_DEC_CHUNKS, 0); k_dec_chunks<0><<<gch, 256, 0, st>>>(d_tabs, d_sym, sym_stride, d_lists, d_ccount, d_coff, ch_stride, ctx->d_T.as<u8>(), stride, max_block); LAUNCH_OK();
    ctx->prof_begin(K_DEC_CHUNK_SCAN, 0); k_dec_chunk_scan<<<nb, 256, 0, st>>>(d_tabs, d_lists, d_ccount, d_coff, ch_stride, max_block); LAUNCH_OK();
    ctx->prof_begin(K_DEC_CHUNKS, 0); k_dec_chunks<1><<<gch, 256, 0, st>>>(d_tabs, d_sym, sym_stride, d_lists, d_ccount, d_coff, ch_stride, ctx->d_T.as<u8>(), stride, max_block); LAUNCH_OK();
    // per-block results: the tail of DecTables (T .. pad)
    struct DecTail { u32 T, alpha, G, status, crc, key; u64 data_bit, end_bit; u32 nsym, ngroups, nblock, pad; };
    static_assert(sizeof(DecTail) == sizeof(DecTables) - offsetof(DecTables, T), "DecTail mirrors the tail of DecTables");
    std::vector<DecTail> db(nb);
    BZ_CHECK(cudaMemcpy2DAsync(db.data(), sizeof(DecTail), (const u8 *)d_tabs + offsetof(DecTables, T), sizeof(DecTables),
                               sizeof(DecTail), nb, cudaMemcpyDeviceToHost, st));
    BZ_CHECK(cudaStreamSynchronize(st));
    std::vector<u32> hlen(nb), hkey(nb);
    u32 combined = 0, max_n = 0; u64 total_n = 0;
    for (u32 k = 0; k < nb; k++) {
        if (db[k].status != 0) { ctx->err = "decode: block " + std::to_string(k) + " status " + std::to_string(db[k].status); return BZ2B200_E_FORMAT; }
        u64 next = k + 1 < nb ? starts[k + 1] : footer_bit;
        if (db[k].end_bit != next) { ctx->err = "decode: block " + std::to_string(k) + " does not end at the next block magic"; return BZ2B200_E_FORMAT; }
        hlen[k] = db[k].nblock; hkey[k] = db[k].key;
        max_n = std::max(max_n, hlen[k]); total_n += hlen[k];
        combined = ((combined << 1) | (combined >> 31)) ^ db[k].crc;
    }
    if (combined != stored_combined) return BZ2B200_E_CRC;
    BZ_CHECK(cudaMemcpyAsync(d_len, hlen.data(), (size_t)nb * 4, cudaMemcpyHostToDevice, st));
    BZ_CHECK(cudaMemcpyAsync(d_keys, hkey.data(), (size_t)nb * 4, cudaMemcpyHostToDevice, st));
    // ---- 3. inverse BWT ----
    Batch B;
    B.nblk = (int)nb; B.stride = stride; B.tiles = stride / BZ_TILE; B.max_n = max_n; B.nbits = 20; B.T = ctx->d_T.as<u8>();
    B.len = d_len; B.total_n = total_n;
    int rc = ibwt_batch(ctx, B, d_keys, ctx->d_bwt.as<u8>());
    if (rc) return rc;
    // ---- 4. inverse RLE1 + CRC ----
    ctx->prof_begin(K_DEC_RLE1_COUNT, total_n); k_rle1_inv<0><<<nb, 256, 0, st>>>(ctx->d_bwt.as<u8>(), d_len, stride, d_olen, nullptr, nullptr); LAUNCH_OK();
    std::vector<u64> olen(nb), ooff(nb);
    BZ_CHECK(cudaMemcpyAsync(olen.data(), d_olen, (size_t)nb * 8, cudaMemcpyDeviceToHost, st));
    BZ_CHECK(cudaStreamSynchronize(st));
    u64 total = 0, max_span = 0;
    std::vector<u32> se(2 * (size_t)nb);
    for (u32 k = 0; k < nb; k++) { ooff[k] = total; total += olen[k]; max_span = std::max(max_span, olen[k]); }
    if (total > 0xFFFFFF00ull) { ctx->err = "decode: output larger than 4 GiB is not supported in one call"; return BZ2B200_E_ARG; }
    for (u32 k = 0; k < nb; k++) { se[2 * k] = (u32)ooff[k]; se[2 * k + 1] = (u32)(ooff[k] + olen[k]); }
    *out_len = (size_t)total;
    if (total > out_cap) return BZ2B200_E_CAP;
    BZ_CHECK(ctx->d_stream.ensure((size_t)total + 64));
    BZ_CHECK(cudaMemcpyAsync(d_ooff, ooff.data(), (size_t)nb * 8, cudaMemcpyHostToDevice, st));
    BZ_CHECK(cudaMemcpyAsync(d_se, se.data(), (size_t)nb * 8, cudaMemcpyHostToDevice, st));
    ctx->prof_begin(K_DEC_RLE1_WRITE, total); k_rle1_inv<1><<<nb, 256, 0, st>>>(ctx->d_bwt.as<u8>(), d_len, stride, nullptr, d_ooff, ctx->d_stream.as<u8>()); LAUNCH_OK();
    BZ_CHECK(ctx->d_crc.ensure((size_t)nb * 4));
    rc = bz_crc_spans_dev(ctx, ctx->d_stream.as<u8>(), d_se, nb, (u32)max_span, ctx->d_crc.as<u32>());
    if (rc) return rc;
    std::vector<u32> crcs(nb);
    BZ_CHECK(cudaMemcpyAsync(crcs.data(), ctx->d_crc.p, (size_t)nb * 4, cudaMemcpyDeviceToHost, st));
    if (total) BZ_CHECK(cudaMemcpyAsync(out, ctx->d_stream.p, (size_t)total, cudaMemcpyDeviceToHost, st));
    BZ_CHECK(cudaStreamSynchronize(st));
    for (u32 k = 0; k < nb; k++)
        if (crcs[k] != db[k].crc) { ctx->err = "decode: CRC mismatch in block " + std::to_string(k); return BZ2B200_E_CRC; }   // enforced, unlike decompress.rs:379-386
    return BZ2B200_OK;
    BZ_API_CATCH
}
