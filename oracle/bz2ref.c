/*
 * bz2ref.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the per-block compression path of ohsnyt/bzip2-rust
 * (mounted read-only at /root/reference while this was written).  See bz2ref.h
 * for the parity-pin status.  Citations are reference file:line.
 *
 * Nothing under bzip2_rust_b200/ may include, link or call this file.
 */
#define _GNU_SOURCE
#include "bz2ref.h"
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* CRC  (src/tools/crc.rs:15-27; table :41-298 is the standard MSB-first      */
/* CRC-32 table for polynomial 0x04C11DB7, generated here instead of listed)  */
/* ------------------------------------------------------------------------- */
static uint32_t g_crc_table[256];
static pthread_once_t g_crc_once = PTHREAD_ONCE_INIT;
static void crc_table_init(void) {
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i << 24;
        for (int k = 0; k < 8; k++) c = (c & 0x80000000u) ? (c << 1) ^ 0x04C11DB7u : (c << 1);
        g_crc_table[i] = c;
    }
}
uint32_t ref_do_crc(uint32_t existing_crc, const uint8_t *data, size_t n) {
    pthread_once(&g_crc_once, crc_table_init);
    uint32_t crc = ~existing_crc;                       /* crc.rs:17 */
    for (size_t i = 0; i < n; i++)                      /* crc.rs:18-20 */
        crc = (crc << 8) ^ g_crc_table[(crc >> 24) ^ data[i]];
    return ~crc;                                        /* crc.rs:21 */
}
uint32_t ref_do_stream_crc(uint32_t s, uint32_t b) {    /* crc.rs:25-27 */
    return ((s << 1) | (s >> 31)) ^ b;
}

/* ------------------------------------------------------------------------- */
/* RLE1 block iterator  (src/tools/rle1.rs:33-264)                            */
/* The buffer/refill geometry is kept because EOF behaviour depends on it.    */
/* ------------------------------------------------------------------------- */
#define MAX_RUN 260                                     /* rle1.rs:29 */
struct ref_rle1_iter {
    const uint8_t *src; size_t src_len, src_pos;        /* stands in for `source: R` */
    size_t block_size;
    uint8_t *buf; size_t len, cap;                      /* `buffer: Vec<u8>` */
    size_t cursor;                                      /* buffer_cursor */
    int data_gone;
    uint32_t block_crc;
    uint8_t *out; size_t out_len, out_cap;
    size_t consumed;                                    /* bytes fed to do_crc for this block */
    int panicked;
};

ref_rle1_iter *ref_rle1_new(const uint8_t *src, size_t n, size_t block_size) {   /* rle1.rs:49-58 */
    ref_rle1_iter *it = (ref_rle1_iter *)calloc(1, sizeof *it);
    it->src = src; it->src_len = n; it->block_size = block_size;
    it->cap = 2 * block_size + 2 * MAX_RUN + 16;
    it->buf = (uint8_t *)malloc(it->cap);
    it->out_cap = block_size + MAX_RUN + 16;
    it->out = (uint8_t *)malloc(it->out_cap);
    return it;
}
void ref_rle1_free(ref_rle1_iter *it) { if (it) { free(it->buf); free(it->out); free(it); } }

static void it_drain(ref_rle1_iter *it, size_t k) {     /* Vec::drain(..k) */
    if (k > it->len) { it->panicked = 1; return; }
    memmove(it->buf, it->buf + k, it->len - k);
    it->len -= k;
}
static int it_refill(ref_rle1_iter *it) {               /* rle1.rs:63-85 */
    if (it->data_gone || it->len - it->cursor < MAX_RUN) {
        it_drain(it, it->cursor);
        it->cursor = 0;
        size_t received = it->src_len - it->src_pos;
        if (received > it->block_size) received = it->block_size;
        if (it->len + received > it->cap) {
            it->cap = it->len + received + 16;
            it->buf = (uint8_t *)realloc(it->buf, it->cap);
        }
        memcpy(it->buf + it->len, it->src + it->src_pos, received);
        it->len += received; it->src_pos += received;
        if (received < it->block_size) { it->data_gone = 1; return 0; }
    }
    return 1;
}
/* checked buffer read: the reference panics on out-of-range indexing */
static uint8_t it_at(ref_rle1_iter *it, size_t i) {
    if (i >= it->len) { it->panicked = 1; return 0; }
    return it->buf[i];
}
/* out.extend_from_slice(&buffer[a..b]) and/or do_crc over the slice */
static void it_emit(ref_rle1_iter *it, size_t a, size_t b, int crc_too, size_t crc_b) {
    if (a > b || b > it->len) { it->panicked = 1; return; }
    if (it->out_len + (b - a) + 1 > it->out_cap) {
        it->out_cap = it->out_len + (b - a) + 64;
        it->out = (uint8_t *)realloc(it->out, it->out_cap);
    }
    memcpy(it->out + it->out_len, it->buf + a, b - a);
    it->out_len += b - a;
    if (crc_too) {
        if (a > crc_b || crc_b > it->len) { it->panicked = 1; return; }
        it->block_crc = ref_do_crc(it->block_crc, it->buf + a, crc_b - a);
        it->consumed += crc_b - a;
    }
}
static uint8_t it_count_dups(ref_rle1_iter *it) {       /* rle1.rs:226-241 */
    size_t c = it->cursor;
    uint8_t v = it_at(it, c);
    if (v != it_at(it, c + 4)) return 0;                /* rle1.rs:228 (may panic) */
    if (it->panicked) return 0;
    size_t k = 0;
    while (k < 251 && c + 4 + k < it->len && it->buf[c + 4 + k] == v) k++;
    /* position(|x| x != compare) over skip(c+4).take(251); unwrap_or(min(251, len-(c+4))) */
    return (uint8_t)k;
}
static void it_get_block(ref_rle1_iter *it, int *last) { /* rle1.rs:89-223 */
    it->out_len = 0;
    size_t start = 0;
    size_t remaining = it->len;
    while (it->out_len + (it->cursor - start) < it->block_size) {   /* rle1.rs:110 */
        if (it->panicked) return;
        if (remaining == 0) {                            /* rle1.rs:115-124 */
            it->len = 0; it->cursor = 0;
            *last = it->data_gone;                        /* data_gone && buffer.is_empty() */
            return;
        } else if (remaining <= 3) {                     /* rle1.rs:126-139 */
            it->cursor = it->len;
            it_emit(it, start, it->cursor, 1, it->cursor);
            it_drain(it, it->cursor);
            it->cursor = 0;
            *last = it->data_gone && it->len == 0;
            return;
        } else {
            if (remaining < MAX_RUN && !it->data_gone) { /* rle1.rs:143-155 */
                if (it->cursor == 0) { it->panicked = 1; return; }   /* usize underflow */
                it->cursor -= 1;
                it_emit(it, start, it->cursor, 1, it->cursor);
                it_drain(it, it->cursor);
                it->cursor = 0;
                it_refill(it);
                remaining = it->len;
                start = 0;
            }
            if (it->cursor == 0) { it->cursor += 1; remaining -= 1; }   /* rle1.rs:157-160 */
            size_t c = it->cursor;
            if (it_at(it, c) != it_at(it, c + 2)) { it->cursor += 2; remaining -= 2; continue; }   /* :164-168 */
            if (it_at(it, c) != it_at(it, c + 1)) { it->cursor += 2; remaining -= 2; continue; }   /* :169-173 */
            if (it_at(it, c) != it_at(it, c - 1)) {      /* rle1.rs:175-181, && short-circuits */
                if (it_at(it, c) != it_at(it, c + 3)) { it->cursor += 2; remaining -= 2; continue; }
            }
            if (it->panicked) return;
            if (it_at(it, c) == it_at(it, c - 1)) it->cursor -= 1;      /* rle1.rs:183-185 */
            {
                uint8_t dups = it_count_dups(it);        /* rle1.rs:189 */
                if (it->panicked) return;
                it->cursor += 4 + dups;                  /* rle1.rs:191 */
                /* crc over start..cursor, output start..cursor-dups then the count byte (:193-201) */
                it_emit(it, start, it->cursor - dups, 1, it->cursor);
                if (it->panicked) return;
                it->out[it->out_len++] = dups;
                start = it->cursor;                      /* rle1.rs:203 */
                it->cursor += 1;                         /* rle1.rs:205 */
                remaining = it->len - start;             /* rle1.rs:207 */
            }
        }
    }
    /* block filled: rle1.rs:212-222 */
    it_emit(it, start, it->cursor, 1, it->cursor);
    it_drain(it, it->cursor);
    it->cursor = 0;
    *last = it->data_gone && it->len == 0;
}
int ref_rle1_next(ref_rle1_iter *it, uint32_t *crc, const uint8_t **block, size_t *block_len,
                  int *last, size_t *consumed) {         /* rle1.rs:250-263 */
    if (it->data_gone && it->len == 0) return 0;
    it_refill(it);
    it->block_crc = 0;
    it->consumed = 0;
    int l = 0;
    it_get_block(it, &l);
    if (it->panicked) return REF_ERR_PANIC;
    *crc = it->block_crc; *block = it->out; *block_len = it->out_len; *last = l;
    if (consumed) *consumed = it->consumed;
    return 1;
}

size_t ref_rle1_decode_reference(const uint8_t *r, size_t n, uint8_t *out, size_t cap) {  /* rle1.rs:267-316 */
    size_t start = 0, cursor = 1, o = 0;
    if (n < 4) { /* `rle1.len() - 4` underflows in the reference; treat as literal copy */
        if (n > cap) return (size_t)-1;
        memcpy(out, r, n); return n;
    }
    while (cursor < n - 4) {
        if (r[cursor] != r[cursor + 2]) { cursor += 2; continue; }
        if (r[cursor] != r[cursor + 1]) { cursor += 2; continue; }
        if (r[cursor] != r[cursor - 1] && r[cursor] != r[cursor + 3]) { cursor += 2; continue; }
        if (r[cursor] == r[cursor - 1]) cursor -= 1;
        size_t lit = cursor + 4 - start, rep = r[cursor + 4];
        if (o + lit + rep > cap) return (size_t)-1;
        memcpy(out + o, r + start, lit); o += lit;
        memset(out + o, r[cursor], rep); o += rep;
        cursor += 5; start = cursor; cursor += 1;
    }
    if (o + (n - start) > cap) return (size_t)-1;
    memcpy(out + o, r + start, n - start); o += n - start;
    return o;
}
size_t ref_rle1_decode_standard(const uint8_t *r, size_t n, uint8_t *out, size_t cap) {
    size_t o = 0, i = 0;
    int run = 0; int prev = -1;
    while (i < n) {
        uint8_t c = r[i++];
        if (run == 4) {            /* c is a repeat count */
            if (o + c > cap) return (size_t)-1;
            memset(out + o, prev, c); o += c; run = 0; prev = -1;
            continue;
        }
        if ((int)c == prev) run++; else { run = 1; prev = c; }
        if (o + 1 > cap) return (size_t)-1;
        out[o++] = c;
    }
    return o;
}

/* ------------------------------------------------------------------------- */
/* BWT, native path  (src/bwt_algorithms/bwt_sort.rs:27-86)                   */
/* ------------------------------------------------------------------------- */
typedef struct { const uint8_t *x; uint32_t n; } rot_ctx;
/* block_compare, bwt_sort.rs:61-86: lexicographic compare of the two full cyclic rotations */
static int rot_cmp(const void *pa, const void *pb, void *pc) {
    const rot_ctx *c = (const rot_ctx *)pc;
    uint32_t a = *(const uint32_t *)pa, b = *(const uint32_t *)pb, n = c->n;
    if (a == b) return 0;
    const uint8_t *x = c->x;
    uint32_t left = n;
    while (left) {
        uint32_t seg = n - (a > b ? a : b);
        if (seg > left) seg = left;
        int r = memcmp(x + a, x + b, seg);
        if (r) return r;
        a += seg; if (a == n) a = 0;
        b += seg; if (b == n) b = 0;
        left -= seg;
    }
    return 0;
}
/* smallest p | n with x p-periodic as a cyclic string (n if primitive); KMP failure function */
static uint32_t cyclic_period(const uint8_t *x, uint32_t n) {
    if (n == 0) return 0;
    uint32_t *f = (uint32_t *)malloc((size_t)(n + 1) * 4);
    f[0] = 0; f[1] = 0;
    uint32_t k = 0;
    for (uint32_t i = 1; i < n; i++) {
        while (k && x[i] != x[k]) k = f[k];
        if (x[i] == x[k]) k++;
        f[i + 1] = k;
    }
    uint32_t q = n - f[n];
    free(f);
    return (n % q == 0) ? q : n;
}
static void bwt_from_index(const uint8_t *x, uint32_t n, const uint32_t *index, uint32_t *key, uint8_t *bwt) {
    /* bwt_sort.rs:45-56, with the SURVEY D.2 rule for fully periodic blocks:
       key = first row of the class of rotations equal to rotation 0 */
    uint32_t p = cyclic_period(x, n);
    uint32_t k = 0;
    for (uint32_t i = 0; i < n; i++) {
        if (index[i] == 0) k = i;
        bwt[i] = index[i] == 0 ? x[n - 1] : x[index[i] - 1];
    }
    if (p != n) while (k > 0 && index[k - 1] % p == 0) k--;
    *key = k;
}
static int bwt_native(const uint8_t *x, uint32_t n, uint32_t *key, uint8_t *bwt) {
    uint32_t *index = (uint32_t *)malloc((size_t)n * 4 + 4);
    for (uint32_t i = 0; i < n; i++) index[i] = i;     /* bwt_sort.rs:36 */
    rot_ctx c = { x, n };
    qsort_r(index, n, 4, rot_cmp, &c);                  /* bwt_sort.rs:39-43 (order of equal rotations is immaterial) */
    bwt_from_index(x, n, index, key, bwt);
    free(index);
    return REF_OK;
}

/* O(n log n)-ish cyclic prefix doubling: same output as bwt_native, usable on
 * inputs where the reference's comparison sort would take hours. */
static int bwt_doubling(const uint8_t *x, uint32_t n, uint32_t *key, uint8_t *bwt) {
    if (n == 0) { *key = 0; return REF_OK; }
    uint32_t *sa = (uint32_t *)malloc((size_t)n * 4);
    uint32_t *rk = (uint32_t *)malloc((size_t)n * 4);   /* rank = head index of the group */
    uint32_t *tmp = (uint32_t *)malloc((size_t)n * 4);
    uint32_t *keyv = (uint32_t *)malloc((size_t)n * 4);
    uint32_t *cnt = (uint32_t *)calloc(65537, 4);
    /* counting sort on the first two bytes */
    for (uint32_t i = 0; i < n; i++) cnt[(((uint32_t)x[i] << 8) | x[(i + 1) % n]) + 1]++;
    for (uint32_t b = 0; b < 65536; b++) cnt[b + 1] += cnt[b];
    for (uint32_t i = 0; i < n; i++) { uint32_t b = ((uint32_t)x[i] << 8) | x[(i + 1) % n]; rk[i] = cnt[b]; }
    {
        uint32_t *pos = (uint32_t *)malloc(65536 * 4);
        memcpy(pos, cnt, 65536 * 4);
        for (uint32_t i = 0; i < n; i++) { uint32_t b = ((uint32_t)x[i] << 8) | x[(i + 1) % n]; sa[pos[b]++] = i; }
        free(pos);
    }
    for (uint32_t h = 2; h < n; h *= 2) {
        int any = 0;
        uint32_t i = 0;
        /* gather keys first so that every group in this round sees rank_h */
        for (uint32_t j = 0; j < n; j++) { uint32_t s = sa[j] + h; keyv[j] = rk[s >= n ? s % n : s]; }
        while (i < n) {
            uint32_t head = rk[sa[i]], e = i + 1;
            while (e < n && rk[sa[e]] == head) e++;
            if (e - i > 1) {
                any = 1;
                /* sort sa[i..e) by keyv: shell-free simple approach = radix on 2x16 bits via tmp when large, insertion when small */
                uint32_t len = e - i;
                if (len <= 16) {
                    for (uint32_t a = i + 1; a < e; a++) {
                        uint32_t kv = keyv[a], sv = sa[a]; uint32_t b = a;
                        while (b > i && keyv[b - 1] > kv) { keyv[b] = keyv[b - 1]; sa[b] = sa[b - 1]; b--; }
                        keyv[b] = kv; sa[b] = sv;
                    }
                } else {
                    /* LSD radix, 3 passes of 8 bits when ranks < 2^24 (n <= 900k), else 4 */
                    int passes = n <= (1u << 16) ? 2 : (n <= (1u << 24) ? 3 : 4);
                    uint32_t *ks = keyv + i, *ss = sa + i, *kt = tmp, *st = (uint32_t *)malloc((size_t)len * 4);
                    uint32_t *kt_base = kt, *st_base = st;
                    for (int p = 0; p < passes; p++) {
                        uint32_t c[257]; memset(c, 0, sizeof c);
                        int sh = 8 * p;
                        for (uint32_t a = 0; a < len; a++) c[((ks[a] >> sh) & 255) + 1]++;
                        for (int b = 0; b < 256; b++) c[b + 1] += c[b];
                        for (uint32_t a = 0; a < len; a++) { uint32_t d = (ks[a] >> sh) & 255; kt[c[d]] = ks[a]; st[c[d]++] = ss[a]; }
                        uint32_t *t1 = ks; ks = kt; kt = t1; t1 = ss; ss = st; st = t1;
                    }
                    if (ks != keyv + i) { memcpy(keyv + i, ks, (size_t)len * 4); memcpy(sa + i, ss, (size_t)len * 4); }
                    (void)kt_base; free(st_base);
                }
            }
            i = e;
        }
        if (!any) break;
        /* re-rank: new head wherever old group changes or key changes */
        i = 0;
        while (i < n) {
            uint32_t head = rk[sa[i]], e = i + 1;
            while (e < n && rk[sa[e]] == head) e++;
            if (e - i > 1) {
                uint32_t cur = i;
                tmp[i] = i;
                for (uint32_t a = i + 1; a < e; a++) { if (keyv[a] != keyv[a - 1]) cur = a; tmp[a] = cur; }
            } else tmp[i] = i;
            i = e;
        }
        for (uint32_t j = 0; j < n; j++) rk[sa[j]] = tmp[j];
    }
    /* key = rank of rotation 0 = head of its class (first row of class) */
    for (uint32_t j = 0; j < n; j++) bwt[j] = sa[j] == 0 ? x[n - 1] : x[sa[j] - 1];
    *key = rk[0];
    free(sa); free(rk); free(tmp); free(keyv); free(cnt);
    return REF_OK;
}

/* ------------------------------------------------------------------------- */
/* SA-IS fallback, bug-for-bug  (src/bwt_algorithms/sais_fallback.rs)         */
/* ------------------------------------------------------------------------- */
#define NONE 0xFFFFFFFFu
typedef struct { uint8_t *s, *lms; uint32_t last, lms_count, s_count; } lms_t;

static void lms_init(lms_t *L, const uint32_t *data, uint32_t n) {   /* sais_fallback.rs:59-131 */
    L->last = n;
    L->s = (uint8_t *)calloc((size_t)n + 1, 1);
    L->lms = (uint8_t *)calloc((size_t)n + 1, 1);
    L->s[n] = 1; L->lms[n] = 1;                          /* :87-88 sentinel is S and LMS */
    int current = 0;                                    /* L */
    uint32_t prev = data[n - 1];
    for (uint32_t k = n - 1; k-- > 0;) {                /* idx = n-2 .. 0  (:96) */
        uint32_t el = data[k];
        if (el < prev) { L->s[k] = 1; current = 1; }    /* :99-104 */
        else if (el == prev) { if (current) L->s[k] = 1; }   /* :107-113 */
        else { if (current) { L->lms[k + 1] = 1; current = 0; } }   /* :115-122 */
        prev = el;
    }
    uint32_t lc = 0, sc = 0;
    for (uint32_t i = 0; i <= n; i++) { lc += L->lms[i]; sc += L->s[i]; }
    L->lms_count = lc; L->s_count = sc;
}
static void lms_free(lms_t *L) { free(L->s); free(L->lms); }

static int is_unequal_lms(const lms_t *L, const uint32_t *data, uint32_t a, uint32_t b) {   /* :158-192 */
    if (a == L->last || b == L->last) return 1;
    uint32_t i = (a > b ? b : a) + 1;
    uint32_t diff = a > b ? a - b : b - a;
    while (i != L->last - diff) {
        uint32_t bb = i + diff;
        if (L->lms[i] && L->lms[bb]) return 0;           /* :177-179 (characters at the LMS positions never compared) */
        if (L->lms[i] || L->lms[bb]) return 1;
        if (data[i] != data[bb]) return 1;
        i++;
    }
    return 1;
}
static void bucket_sizes(const uint32_t *data, uint32_t n, uint32_t size, uint32_t *sz) {   /* :285-314 */
    memset(sz, 0, (size_t)size * 4);
    for (uint32_t i = 0; i < n; i++) sz[data[i]]++;
}
static void bucket_heads(const uint32_t *sz, uint32_t size, uint32_t *h) {   /* :317-330 */
    uint32_t idx = 1; for (uint32_t i = 0; i < size; i++) { h[i] = idx; idx += sz[i]; }
}
static void bucket_tails(const uint32_t *sz, uint32_t size, uint32_t *t) {   /* :332-345 */
    uint32_t idx = 1; for (uint32_t i = 0; i < size; i++) { idx += sz[i]; t[i] = idx - 1; }
}
void ref_bucket_sizes_heads_tails(const uint32_t *data, uint32_t n, uint32_t size,
                                  uint32_t *sizes, uint32_t *heads, uint32_t *tails) {
    bucket_sizes(data, n, size, sizes); bucket_heads(sizes, size, heads); bucket_tails(sizes, size, tails);
}
static int induced_sort_l(const uint32_t *data, uint32_t *bk, const uint32_t *sz, uint32_t asz, const lms_t *L) {   /* :400-430 */
    uint32_t *heads = (uint32_t *)malloc((size_t)asz * 4);
    bucket_heads(sz, asz, heads);
    for (uint32_t idx = 0; idx < L->last; idx++) {
        if (bk[idx] != NONE) {
            uint32_t prev = bk[idx] == 0 ? L->last : bk[idx] - 1;
            if (!L->s[prev]) {
                uint32_t h = heads[data[prev]];
                if (h > L->last) { free(heads); return REF_ERR_PANIC; }
                bk[h] = prev; heads[data[prev]]++;
            }
        }
    }
    free(heads); return REF_OK;
}
static int induced_sort_s(const uint32_t *data, uint32_t *bk, const uint32_t *sz, uint32_t asz, const lms_t *L) {   /* :433-467 */
    uint32_t *tails = (uint32_t *)malloc((size_t)asz * 4);
    bucket_tails(sz, asz, tails);
    for (uint32_t idx = L->last; idx > 0; idx--) {
        if (bk[idx] == NONE) { free(tails); return REF_ERR_PANIC; }   /* unwrap() on None :452 */
        if (bk[idx] != 0) {
            uint32_t prev = bk[idx] - 1;
            if (L->s[prev]) {
                if (prev >= L->last) { free(tails); return REF_ERR_PANIC; }
                uint32_t t = tails[data[prev]];
                if (t > L->last) { free(tails); return REF_ERR_PANIC; }
                bk[t] = prev; tails[data[prev]]--;
            }
        }
    }
    free(tails); return REF_OK;
}
/* sa_is, sais_fallback.rs:469-578.  Returns malloc'd SA of n entries (NULL on "panic"). */
static uint32_t *sa_is(const uint32_t *data, uint32_t n, uint32_t asz, int *err) {
    if (n == 0) return (uint32_t *)calloc(1, 4);
    lms_t L; lms_init(&L, data, n);
    uint32_t *sz = (uint32_t *)malloc((size_t)asz * 4);
    uint32_t *tails = (uint32_t *)malloc((size_t)asz * 4);
    uint32_t *bk = (uint32_t *)malloc(((size_t)n + 1) * 4);
    uint32_t *names = NULL, *offs = NULL, *summary = NULL, *offsets = NULL, *ssv = NULL, *result = NULL;
    bucket_sizes(data, n, asz, sz);
    /* initial_buckets_sort :375-397 */
    bucket_tails(sz, asz, tails);
    for (uint32_t i = 0; i <= n; i++) bk[i] = NONE;
    bk[0] = L.last;
    for (uint32_t idx = L.last; idx-- > 0;) {
        if (L.lms[idx]) { bk[tails[data[idx]]] = idx; tails[data[idx]]--; }
    }
    bk[0] = n;
    if ((*err = induced_sort_l(data, bk, sz, asz, &L)) != REF_OK) goto done;      /* STEP 4 */
    if ((*err = induced_sort_s(data, bk, sz, asz, &L)) != REF_OK) goto done;      /* STEP 5 */
    /* make_summary :623-664 */
    names = (uint32_t *)malloc(((size_t)n + 1) * 4);
    offs = (uint32_t *)malloc(((size_t)n + 1) * 4);
    for (uint32_t i = 0; i <= n; i++) { names[i] = NONE; offs[i] = NONE; }
    uint32_t current_name = 0;
    names[bk[0]] = 0; offs[bk[0]] = L.last;
    uint32_t prev_lms = bk[0];
    for (uint32_t j = 1; j <= n; j++) {
        uint32_t p = bk[j];
        if (p == NONE) { *err = REF_ERR_PANIC; goto done; }
        if (L.lms[p]) {
            if (is_unequal_lms(&L, data, prev_lms, p)) { prev_lms = p; current_name++; }
            names[p] = current_name; offs[p] = p;
        }
    }
    uint32_t summary_size = current_name + 1;
    uint32_t sl = 0;
    summary = (uint32_t *)malloc(((size_t)L.lms_count + 1) * 4);
    offsets = (uint32_t *)malloc(((size_t)L.lms_count + 1) * 4);
    for (uint32_t i = 0; i <= n; i++) if (names[i] != NONE) { summary[sl] = names[i]; offsets[sl] = offs[i]; sl++; }
    /* make_summary_suffix_vec :668-688 */
    ssv = (uint32_t *)malloc(((size_t)sl + 1) * 4);
    if (summary_size != L.lms_count) {
        uint32_t *rec = sa_is(summary, sl, summary_size, err);
        if (*err != REF_OK) { free(rec); goto done; }
        ssv[0] = sl;
        memcpy(ssv + 1, rec, (size_t)sl * 4);
        free(rec);
    } else {
        for (uint32_t i = 0; i <= sl; i++) ssv[i] = summary_size;
        for (uint32_t i = 0; i < sl; i++) ssv[summary[i] + 1] = i;
    }
    /* STEP 7 :531-546 */
    for (uint32_t i = 0; i <= n; i++) bk[i] = NONE;
    bucket_tails(sz, asz, tails);
    for (uint32_t k = sl + 1; k-- > 2;) {               /* iter().skip(2).rev() */
        uint32_t di = offsets[ssv[k]];
        if (di >= n) { *err = REF_ERR_PANIC; goto done; }
        uint32_t b = data[di];
        bk[tails[b]] = di; tails[b]--;
    }
    bk[0] = n;
    if ((*err = induced_sort_l(data, bk, sz, asz, &L)) != REF_OK) goto done;      /* STEP 9 */
    if ((*err = induced_sort_s(data, bk, sz, asz, &L)) != REF_OK) goto done;      /* STEP 10 */
    result = (uint32_t *)malloc((size_t)n * 4);
    for (uint32_t i = 0; i < n; i++) {
        if (bk[i + 1] == NONE) { *err = REF_ERR_PANIC; free(result); result = NULL; goto done; }
        result[i] = bk[i + 1];
    }
done:
    lms_free(&L); free(sz); free(tails); free(bk); free(names); free(offs); free(summary); free(offsets); free(ssv);
    return result;
}
uint32_t ref_duval(const uint8_t *in, uint32_t n) {      /* sais_fallback.rs:781-804 */
    uint32_t final_start = 0, i = 0;
    while (i < n) {
        uint32_t j = i + 1, k = i;
        while (j < n && in[k] <= in[j]) { if (in[k] < in[j]) k = i; else k++; j++; }
        while (i <= k) { final_start = i; i += j - k; }
    }
    return final_start;
}
static int bwt_sais(const uint8_t *x, uint32_t n, uint32_t *key, uint8_t *bwt) {   /* sais_entry :582-620 */
    uint32_t off = ref_duval(x, n);                      /* rotate_duval :808-816 */
    uint32_t *d = (uint32_t *)malloc((size_t)n * 4);
    for (uint32_t i = 0; i < n; i++) d[i] = x[(i + off) % n];
    int err = REF_OK;
    uint32_t *index = sa_is(d, n, 256, &err);
    if (err != REF_OK || !index) { free(d); free(index); return REF_ERR_PANIC; }
    uint32_t k = 0, zero_pos = n - off;                 /* :601 */
    for (uint32_t i = 0; i < n; i++) {
        if (index[i] == zero_pos) k = i;
        bwt[i] = (uint8_t)(index[i] == 0 ? d[n - 1] : d[index[i] - 1]);
    }
    *key = k;
    free(d); free(index);
    return REF_OK;
}
void ref_lms_types(const uint8_t *data, uint32_t n, uint8_t *is_s, uint8_t *is_lms) {
    uint32_t *d = (uint32_t *)malloc((size_t)n * 4);
    for (uint32_t i = 0; i < n; i++) d[i] = data[i];
    lms_t L; lms_init(&L, d, n);
    memcpy(is_s, L.s, (size_t)n + 1); memcpy(is_lms, L.lms, (size_t)n + 1);
    lms_free(&L); free(d);
}
uint32_t ref_lms_count(const uint8_t *data, uint32_t n) {   /* lms_complexity numerator, :821-829 */
    uint32_t m = n < 5000 ? n : 5000;
    if (m == 0) return 0;
    uint32_t *d = (uint32_t *)malloc((size_t)m * 4);
    for (uint32_t i = 0; i < m; i++) d[i] = data[i];
    lms_t L; lms_init(&L, d, m);
    uint32_t c = L.lms_count;
    lms_free(&L); free(d);
    return c;
}
int ref_bwt_encode(const uint8_t *block, uint32_t n, int mode, uint32_t *key, uint8_t *bwt, int *path_used) {
    int path = 0;
    /* bwt_sort.rs:29: len > 5000 && lms_count/5000.0 < 0.3  <=>  lms_count <= 1499 */
    if (mode == REF_BWT_EXACT && n > 5000 && ref_lms_count(block, n) <= 1499) path = 1;
    if (path_used) *path_used = path;
    if (n == 0) { *key = 0; return REF_OK; }
    if (path == 1) return bwt_sais(block, n, key, bwt);
    if (mode == REF_BWT_SPEC_FAST) return bwt_doubling(block, n, key, bwt);
    return bwt_native(block, n, key, bwt);
}
int ref_bwt_decode(uint32_t key, const uint8_t *bwt, uint32_t n, uint8_t *out) {   /* bwt_sort.rs:91-130 */
    if (n == 0) return REF_OK;
    if (key >= n) return REF_ERR_ARG;
    uint32_t freq[256] = {0}, cum[256];
    for (uint32_t i = 0; i < n; i++) freq[bwt[i]]++;
    cum[0] = 0; for (int i = 0; i < 255; i++) cum[i + 1] = cum[i] + freq[i];
    uint32_t *t = (uint32_t *)calloc(n, 4);
    for (uint32_t i = 0; i < n; i++) { uint8_t s = bwt[i]; t[i] |= (uint32_t)s << 24; t[cum[s]] |= i; cum[s]++; }
    uint32_t el = t[key]; key = el & 0xFFFFFF; out[n - 1] = (uint8_t)(el >> 24);
    for (uint32_t i = 1; i < n; i++) { el = t[key]; key = el & 0xFFFFFF; out[i - 1] = (uint8_t)(el >> 24); }
    free(t);
    return REF_OK;
}

/* ------------------------------------------------------------------------- */
/* MTF + RLE2  (src/tools/rle2_mtf.rs:23-177, :293-322)                        */
/* ------------------------------------------------------------------------- */
int ref_rle2_mtf_encode(const uint8_t *block, uint32_t n, uint16_t *rle2, uint32_t *m_out,
                        uint32_t freqs[256], uint16_t symmap[17], int *nmap) {
    uint8_t used[256] = {0};
    for (uint32_t i = 0; i < n; i++) used[block[i]] = 1;           /* :26-29 */
    uint8_t mtf[256] = {0}; int cnt = 0;
    for (int s = 0; s < 256; s++) if (used[s]) mtf[cnt++] = (uint8_t)s;   /* :30-39 */
    uint16_t eob = (uint16_t)(cnt + 1);                             /* :42 */
    /* encode_sym_map_from_bool_map :293-322 */
    uint16_t maps[17] = {0};
    for (int idx = 0; idx < 256; idx++) if (used[idx]) {
        maps[0] |= 0x8000 >> (idx >> 4);
        maps[1 + (idx >> 4)] |= 0x8000 >> (idx & 15);
    }
    int k = 0; for (int i = 0; i < 17; i++) if (maps[i] > 0) symmap[k++] = maps[i];
    *nmap = k;
    memset(freqs, 0, 256 * 4);
    uint32_t out = 0; size_t zeros = 0;
#define FLUSH_ZEROS() do { \
        if (zeros == 1) { rle2[out++] = 0; freqs[0]++; } \
        else if (zeros == 2) { rle2[out++] = 1; freqs[1]++; } \
        else if (zeros > 2) { size_t nn = zeros - 1; for (;;) { rle2[out++] = (uint16_t)(nn & 1); freqs[nn & 1]++; if (nn < 2) break; nn = (nn - 2) >> 1; } } \
        zeros = 0; } while (0)
    for (uint32_t i = 0; i < n; i++) {                              /* :61-131 */
        uint8_t byte = block[i];
        int idx = 0; while (mtf[idx] != byte) idx++;
        if (idx == 0) { zeros++; continue; }
        FLUSH_ZEROS();
        freqs[idx]++;                                               /* :104 (position, not emitted symbol) */
        rle2[out++] = (uint16_t)(idx + 1);
        memmove(mtf + 1, mtf, (size_t)idx);
        mtf[0] = byte;
    }
    FLUSH_ZEROS();                                                  /* :134-164 */
#undef FLUSH_ZEROS
    rle2[out++] = eob;                                              /* :166-167 */
    *m_out = out;
    return REF_OK;
}

/* ------------------------------------------------------------------------- */
/* BitPacker  (src/bitstream/bitpacker.rs:17-112)                             */
/* ------------------------------------------------------------------------- */
void ref_bp_init(ref_bitpacker *bp, uint8_t *buf, size_t cap) {
    bp->out = buf; bp->len = 0; bp->cap = cap; bp->queue = 0; bp->q_bits = 0; bp->padding = 0; bp->overflow = 0;
}
static void bp_write_stream(ref_bitpacker *bp) {        /* :45-51 */
    while (bp->q_bits > 7) {
        uint8_t byte = (uint8_t)(bp->queue >> (bp->q_bits - 8));
        if (bp->len < bp->cap) bp->out[bp->len] = byte; else bp->overflow = 1;
        bp->len++;
        bp->q_bits -= 8;
    }
}
void ref_bp_out24(ref_bitpacker *bp, uint32_t data) {   /* :64-70 */
    uint32_t depth = data >> 24;
    bp->queue <<= depth;
    if (depth) bp->queue |= (uint64_t)(data & (0xffffffffu >> (32 - depth)));
    bp->q_bits += depth;
    bp_write_stream(bp);
}
void ref_bp_out32(ref_bitpacker *bp, uint32_t data) {   /* :73-78 */
    bp->queue <<= 32; bp->queue |= data; bp->q_bits += 32; bp_write_stream(bp);
}
void ref_bp_out16(ref_bitpacker *bp, uint16_t data) {   /* :81-86 */
    bp->queue <<= 16; bp->queue |= data; bp->q_bits += 16; bp_write_stream(bp);
}
void ref_bp_flush(ref_bitpacker *bp) {                  /* :98-106 */
    if (bp->q_bits > 0) {
        bp->padding = (uint8_t)(8 - bp->q_bits % 8);
        bp->queue <<= bp->padding;
        bp->q_bits += bp->padding;
        bp_write_stream(bp);
    }
}

/* ------------------------------------------------------------------------- */
/* Huffman  (src/huffman_coding/huffman.rs, huffman_code_from_weights.rs)     */
/* ------------------------------------------------------------------------- */
void ref_init_tables(const uint32_t freqs[256], int table_count, uint16_t eob, uint32_t tables[6][258]) {   /* huffman.rs:472-532 */
    for (int t = 0; t < 6; t++) for (int i = 0; i < 258; i++) tables[t][i] = 15;
    uint32_t sum = 0; for (int i = 0; i < 256; i++) sum += freqs[i];
    uint32_t limit = sum / (uint32_t)table_count;                   /* :478 */
    int ti = table_count - 1; uint32_t portion = 0;
    int lim = (int)eob + 1; if (lim > 256) lim = 256;               /* freqs has 256 entries, take(eob+1) :511 */
    for (int i = 0; i < lim; i++) {
        uint32_t f = freqs[i];
        if (portion + f > limit && (ti == 2 || ti == 4)) {          /* :513-521 */
            ti = ti > 0 ? ti - 1 : 0;
            tables[ti][i] = 0;
            portion = f;
            if (portion > limit) { tables[ti][i] = 0; ti = ti > 0 ? ti - 1 : 0; portion = 0; }
        } else {                                                    /* :522-529 */
            portion += f;
            tables[ti][i] = 0;
            if (portion > limit) { ti = ti > 0 ? ti - 1 : 0; portion = 0; }
        }
    }
}

typedef struct { uint32_t weight, syms; int16_t left, right; uint8_t depth; } hnode;
/* Node ordering huffman.rs:55-73: descending (weight, syms); smallest at the end is popped. */
static int key_less(const hnode *a, const hnode *b) {   /* key(a) < key(b) */
    return a->weight < b->weight || (a->weight == b->weight && a->syms < b->syms);
}
static uint32_t add_weights(uint32_t a, uint32_t b) {   /* huffman_code_from_weights.rs:105-109 */
    uint32_t da = a & 0xff, db = b & 0xff;
    return ((a & 0xffffff00u) + (b & 0xffffff00u)) | (1 + (da > db ? da : db));
}
static int node_desc_cmp(const void *pa, const void *pb, void *ctx) {
    const hnode *nodes = (const hnode *)ctx;
    const hnode *a = nodes + *(const int16_t *)pa, *b = nodes + *(const int16_t *)pb;
    if (key_less(b, a)) return -1;
    if (key_less(a, b)) return 1;
    return 0;
}
static void improve_code_len(uint32_t *codes, const uint32_t *sym_weight, uint16_t eob,
                             uint32_t *tie_events, uint32_t *retries, uint32_t *tie_unpinned) {   /* huffman_code_from_weights.rs:17-84 */
    int nsym = (int)eob + 1;
    uint32_t weight[258];
    for (int i = 0; i < nsym; i++) weight[i] = sym_weight[i] == 0 ? 256 : sym_weight[i] << 8;   /* :31 */
    hnode nodes[520]; int16_t order[260];
    for (;;) {
        int nn = nsym, cnt = nsym;
        for (int i = 0; i < nsym; i++) {
            nodes[i].weight = weight[i]; nodes[i].syms = (uint32_t)i; nodes[i].depth = 0;
            nodes[i].left = nodes[i].right = -1; order[i] = (int16_t)i;
        }
        /* tree.sort_unstable() on leaves: keys unique (syms unique), any correct sort gives the same order */
        qsort_r(order, (size_t)cnt, sizeof(int16_t), node_desc_cmp, nodes);
        while (cnt > 1) {                                           /* :46-60 */
            int16_t r = order[--cnt], l = order[--cnt];             /* right = smallest, left = next */
            hnode *p = &nodes[nn];
            p->weight = add_weights(nodes[l].weight, nodes[r].weight);
            p->depth = (uint8_t)((nodes[l].depth > nodes[r].depth ? nodes[l].depth : nodes[r].depth) + 1);
            p->syms = nodes[l].syms + nodes[r].syms;
            p->left = l; p->right = r;
            /* the reference re-sorts the whole vector (sort_unstable, :49); the vector is sorted
               except for the pushed parent, so this is an insertion.  Tie rule:
               the parent goes AFTER existing equal keys, i.e. nearer the popped end. */
            int pos = cnt;
            while (pos > 0 && key_less(&nodes[order[pos - 1]], p)) { order[pos] = order[pos - 1]; pos--; }
            if (pos > 0 && tie_events) {
                const hnode *q = &nodes[order[pos - 1]];
                if (q->weight == p->weight && q->syms == p->syms) {
                    (*tie_events)++;
                    /* the reference sorts a vector of cnt + 1 nodes here (:49).  rustc 1.65's sort_unstable is an insertion
                       sort up to 20 elements and repairs a nearly sorted input from 50 on (both keep the parent behind its
                       equal); 21..49 elements go through an unstable partition: the order of the two equal nodes -- and with
                       it the code lengths -- is not pinned by anything in this repository */
                    if (tie_unpinned && cnt + 1 >= 21 && cnt + 1 <= 49) (*tie_unpinned)++;
                }
            }
            order[pos] = (int16_t)nn;
            cnt++; nn++;
        }
        hnode *root = &nodes[order[0]];
        if (root->depth <= 17) {                                    /* :65-72 */
            /* return_leaves :88-101: leaf depth below the root */
            uint8_t d[520];
            d[order[0]] = 0;
            for (int i = nn - 1; i >= nsym; i--) { d[nodes[i].left] = d[i] + 1; d[nodes[i].right] = d[i] + 1; }
            if (nsym == 1) d[0] = 0;
            for (int i = 0; i < nsym; i++) codes[i] = d[i];
            return;
        }
        if (retries) (*retries)++;
        for (int i = 0; i < nsym; i++) { uint32_t j = weight[i] >> 8; j = 1 + j / 2; weight[i] = j << 8; }   /* :76-80 */
    }
}

void ref_improve_code_len(uint32_t *codes, const uint32_t *sym_weight, uint16_t eob,
                          uint32_t *tie_events, uint32_t *retries) {
    improve_code_len(codes, sym_weight, eob, tie_events, retries, NULL);
}

int ref_huf_encode(ref_bitpacker *bp, const uint16_t *rle2, uint32_t m, const uint32_t freq[256],
                   uint16_t eob, const uint16_t *symmap, int nmap, ref_huf_info *info) {   /* huffman.rs:79-468 */
    int T = m < 200 ? 2 : m < 600 ? 3 : m < 1200 ? 4 : m < 2400 ? 5 : 6;   /* :87-93 */
    static __thread uint32_t tables[6][258];
    ref_init_tables(freq, T, eob, tables);                          /* :96 */
    uint32_t G = m / 50 + (m % 50 != 0);                            /* :99 */
    uint8_t *sel = (uint8_t *)malloc(G ? G : 1);
    uint32_t ties = 0, retries = 0, unpinned = 0;
    for (int iter = 0; iter < 4; iter++) {                          /* :114 */
        static __thread uint32_t rfreq[6][258];
        memset(rfreq, 0, sizeof rfreq);
        for (uint32_t g = 0; g < G; g++) {                          /* :137 chunks(50) */
            uint32_t a = g * 50, b = a + 50 > m ? m : a + 50;
            uint32_t cost[6] = {0, 0, 0, 0, 0, 0};
            for (uint32_t i = a; i < b; i++) for (int t = 0; t < T; t++) cost[t] += tables[t][rle2[i]];   /* :143-147 */
            int bt = 0; for (int t = 1; t < T; t++) if (cost[t] < cost[bt]) bt = t;   /* first minimum :150-153 */
            for (uint32_t i = a; i < b; i++) rfreq[bt][rle2[i]]++;  /* :165-167 */
            if (iter == 3) sel[g] = (uint8_t)bt;                    /* :171-173 */
        }
        for (int t = 0; t < T; t++) improve_code_len(tables[t], rfreq[t], eob, &ties, &retries, &unpinned);   /* :197-199 */
    }
    for (int i = 0; i < nmap; i++) ref_bp_out16(bp, symmap[i]);     /* :209-212 */
    ref_bp_out24(bp, (3u << 24) | (uint32_t)T);                     /* :216 */
    ref_bp_out24(bp, (15u << 24) | G);                              /* :224 */
    /* selector MTF + unary :237-292 */
    {
        int idx6[6] = {0, 1, 2, 3, 4, 5};
        for (uint32_t g = 0; g < G; g++) {
            int p = 0; while (idx6[p] != sel[g]) p++;
            int v = idx6[p];
            for (int k = p; k > 0; k--) idx6[k] = idx6[k - 1];
            idx6[0] = v;
            ref_bp_out24(bp, ((uint32_t)(p + 1) << 24) | ((1u << (p + 1)) - 2));   /* p ones then a zero */
        }
    }
    /* per table: canonical codes, then origin + deltas  :311-447 */
    static __thread uint32_t codes[6][258];
    int nsym = (int)eob + 1;
    for (int t = 0; t < T; t++) {
        /* sort (len, sym) ascending :330; assign codes :365-374 */
        uint32_t minlen = 99, maxlen = 0;
        for (int s = 0; s < nsym; s++) { if (tables[t][s] < minlen) minlen = tables[t][s]; if (tables[t][s] > maxlen) maxlen = tables[t][s]; }
        uint32_t code = 0, curlen = minlen;
        for (uint32_t len = minlen; len <= maxlen; len++) {
            for (int s = 0; s < nsym; s++) if (tables[t][s] == len) {
                if (len != curlen) { code <<= (len - curlen); curlen = len; }
                codes[t][s] = (len << 24) | code;
                code++;
            }
        }
        uint32_t origin = tables[t][0];                             /* :391 */
        ref_bp_out24(bp, (5u << 24) | origin);                      /* :398 */
        for (int s = 0; s < nsym; s++) {                            /* :401-438 */
            int delta = (int)tables[t][s] - (int)origin;
            origin = tables[t][s];
            while (delta > 0) { ref_bp_out24(bp, 0x02000002u); delta--; }
            while (delta < 0) { ref_bp_out24(bp, 0x02000003u); delta++; }
            ref_bp_out24(bp, 0x01000000u);
        }
    }
    for (uint32_t g = 0; g < G; g++) {                              /* :452-466 */
        uint32_t a = g * 50, b = a + 50 > m ? m : a + 50;
        for (uint32_t i = a; i < b; i++) ref_bp_out24(bp, codes[sel[g]][rle2[i]]);
    }
    if (info) {
        info->table_count = T; info->selector_count = G; info->tie_events = ties; info->retries = retries; info->tie_unpinned = unpinned;
        for (int t = 0; t < 6; t++) for (int s = 0; s < 258; s++) info->lengths[t][s] = t < T && s < nsym ? (uint8_t)tables[t][s] : 0;
        if (info->selectors) memcpy(info->selectors, sel, G);
    }
    free(sel);
    return REF_OK;
}

/* ------------------------------------------------------------------------- */
/* compress_block  (src/compression/compress_block.rs:24-67)                  */
/* ------------------------------------------------------------------------- */
int ref_compress_block(const uint8_t *block, uint32_t n, uint32_t crc, int bwt_mode,
                       uint8_t *out, size_t cap, size_t *out_len, uint8_t *padding, ref_block_info *info) {
    if (n == 0) return REF_ERR_PANIC;                   /* rle2[rle2.len()-1] with no symbol map -> invalid */
    ref_bitpacker bp; ref_bp_init(&bp, out, cap);
    ref_bp_out24(&bp, 0x18314159u);                     /* :34 */
    ref_bp_out24(&bp, 0x18265359u);                     /* :35 */
    ref_bp_out32(&bp, crc);                             /* :36 */
    ref_bp_out24(&bp, 0x01000000u);                     /* :41 */
    uint8_t *bwt = (uint8_t *)malloc(n);
    uint16_t *rle2 = (uint16_t *)malloc(((size_t)n + 1) * 2);
    uint32_t key = 0, m = 0, freq[256]; uint16_t symmap[17]; int nmap = 0, path = 0;
    int rc = ref_bwt_encode(block, n, bwt_mode, &key, bwt, &path);   /* :44 */
    if (rc != REF_OK) { free(bwt); free(rle2); return rc; }
    ref_bp_out24(&bp, 0x18000000u | key);               /* :48 */
    ref_rle2_mtf_encode(bwt, n, rle2, &m, freq, symmap, &nmap);   /* :50 */
    uint16_t eob = rle2[m - 1];                         /* :53 */
    ref_huf_info hi; memset(&hi, 0, sizeof hi);
    ref_huf_encode(&bp, rle2, m, freq, eob, symmap, nmap, &hi);   /* :56 */
    ref_bp_flush(&bp);                                  /* :65 */
    free(bwt); free(rle2);
    if (info) { info->key = key; info->path_used = path; info->m = m; info->huf = hi; info->huf.selectors = NULL; }
    *out_len = bp.len; *padding = bp.padding;
    return bp.overflow ? REF_ERR_CAP : REF_OK;
}

/* ------------------------------------------------------------------------- */
/* compress + BitWriter  (compress.rs:40-136, bitwriter.rs:42-173)            */
/* ------------------------------------------------------------------------- */
typedef struct { uint8_t *out; size_t len, cap; uint64_t queue; uint32_t q_bits; int overflow; } bitwriter;
static void bw_out8(bitwriter *w, uint8_t d) {          /* bitwriter.rs:135-153 (queue flushing is an implementation detail) */
    w->queue = (w->queue << 8) | d; w->q_bits += 8;
    while (w->q_bits >= 8 + 7) {  /* keep < 8 spare bits pending so that un-padding can shift them out */
        uint8_t byte = (uint8_t)(w->queue >> (w->q_bits - 8));
        if (w->len < w->cap) w->out[w->len] = byte; else w->overflow = 1;
        w->len++; w->q_bits -= 8;
    }
}
static void bw_flush(bitwriter *w) {                    /* :158-172 */
    while (w->q_bits > 7) {
        uint8_t byte = (uint8_t)(w->queue >> (w->q_bits - 8));
        if (w->len < w->cap) w->out[w->len] = byte; else w->overflow = 1;
        w->len++; w->q_bits -= 8;
    }
    if (w->q_bits > 0) {
        uint8_t byte = (uint8_t)((w->queue & (0xffu >> (8 - w->q_bits))) << (8 - w->q_bits));
        if (w->len < w->cap) w->out[w->len] = byte; else w->overflow = 1;
        w->len++; w->q_bits = 0;
    }
}

typedef struct {
    uint8_t *data; size_t len; uint32_t crc; int last;
    uint8_t *packed; size_t packed_len; uint8_t padding; int rc; ref_block_info info; int divergent;
} blk_t;
typedef struct { blk_t *blks; uint32_t nb; volatile uint32_t next; int mode; int want_div; pthread_mutex_t mu; } pool_t;
static void compress_one(blk_t *b, int mode, int want_div) {
    size_t cap = b->len + b->len / 2 + 4096;
    b->packed = (uint8_t *)malloc(cap);
    b->rc = ref_compress_block(b->data, (uint32_t)b->len, b->crc, mode, b->packed, cap, &b->packed_len, &b->padding, &b->info);
    b->divergent = 0;
    if (want_div && b->rc == REF_OK && b->info.path_used == 1) {
        uint8_t *b1 = (uint8_t *)malloc(b->len), *b2 = (uint8_t *)malloc(b->len); uint32_t k1, k2; int p;
        ref_bwt_encode(b->data, (uint32_t)b->len, REF_BWT_EXACT, &k1, b1, &p);
        ref_bwt_encode(b->data, (uint32_t)b->len, REF_BWT_SPEC_FAST, &k2, b2, &p);
        /* divergent iff the block would not decode to the input: BWT bytes differ, or key outside rotation-0's class */
        if (memcmp(b1, b2, b->len) != 0) b->divergent = 1;
        else {
            uint8_t *d = (uint8_t *)malloc(b->len);
            ref_bwt_decode(k1, b1, (uint32_t)b->len, d);
            if (memcmp(d, b->data, b->len) != 0) b->divergent = 1;
            free(d);
        }
        free(b1); free(b2);
    }
}
static void *pool_worker(void *arg) {
    pool_t *p = (pool_t *)arg;
    for (;;) {
        pthread_mutex_lock(&p->mu);
        uint32_t i = p->next++;
        pthread_mutex_unlock(&p->mu);
        if (i >= p->nb) break;
        compress_one(&p->blks[i], p->mode, p->want_div);
    }
    return NULL;
}
int ref_compress_stream(const uint8_t *in, size_t n, int level, int bwt_mode, int threads,
                        uint8_t *out, size_t cap, size_t *out_len, ref_stream_stats *stats) {
    if (level < 1 || level > 9) return REF_ERR_ARG;
    size_t block_size = (size_t)level * 100000 - 19;    /* compress.rs:55 */
    ref_rle1_iter *it = ref_rle1_new(in, n, block_size);
    blk_t *blks = NULL; uint32_t nb = 0, cb = 0; int rc = REF_OK;
    for (;;) {                                          /* the iterator is sequential (par_bridge mutex) */
        uint32_t crc; const uint8_t *bp; size_t bl; int last;
        int r = ref_rle1_next(it, &crc, &bp, &bl, &last, NULL);
        if (r == 0) break;
        if (r < 0) { rc = r; break; }
        if (nb == cb) { cb = cb ? cb * 2 : 16; blks = (blk_t *)realloc(blks, cb * sizeof(blk_t)); }
        memset(&blks[nb], 0, sizeof(blk_t));
        blks[nb].data = (uint8_t *)malloc(bl ? bl : 1); memcpy(blks[nb].data, bp, bl);
        blks[nb].len = bl; blks[nb].crc = crc; blks[nb].last = last;
        nb++;
    }
    ref_rle1_free(it);
    if (rc == REF_OK) {
        pool_t p; p.blks = blks; p.nb = nb; p.next = 0; p.mode = bwt_mode; p.want_div = stats != NULL && bwt_mode == REF_BWT_EXACT;
        pthread_mutex_init(&p.mu, NULL);
        if (threads <= 1) pool_worker(&p);
        else {
            pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
            for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, pool_worker, &p);
            for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
            free(th);
        }
        pthread_mutex_destroy(&p.mu);
    }
    /* writer thread: BitWriter::add_block in order, bitwriter.rs:77-132 */
    bitwriter w = { out, 0, cap, 0, 0, 0 };
    uint32_t stream_crc = 0;
    if (stats) memset(stats, 0, sizeof *stats);
    for (uint32_t i = 0; i < nb && rc == REF_OK; i++) {
        blk_t *b = &blks[i];
        if (b->rc != REF_OK) { rc = b->rc; break; }
        if (stream_crc == 0) {                          /* :84-86 (re-emits the header whenever the running crc is 0) */
            bw_out8(&w, 'B'); bw_out8(&w, 'Z'); bw_out8(&w, 'h'); bw_out8(&w, (uint8_t)(level + 0x30));   /* :67-72 */
        }
        uint32_t bcrc = ((uint32_t)b->packed[6] << 24) | ((uint32_t)b->packed[7] << 16) | ((uint32_t)b->packed[8] << 8) | b->packed[9];   /* :89 */
        stream_crc = ref_do_stream_crc(stream_crc, bcrc);   /* :91 */
        for (size_t k = 0; k < b->packed_len; k++) bw_out8(&w, b->packed[k]);   /* :94 */
        if (b->padding > 0) { w.queue >>= b->padding; w.q_bits -= b->padding; }   /* :97-100 */
        if (b->last) {                                  /* :103-114 */
            static const uint8_t magic[6] = {0x17, 0x72, 0x45, 0x38, 0x50, 0x90};
            for (int k = 0; k < 6; k++) bw_out8(&w, magic[k]);
            bw_out8(&w, (uint8_t)(stream_crc >> 24)); bw_out8(&w, (uint8_t)(stream_crc >> 16));
            bw_out8(&w, (uint8_t)(stream_crc >> 8)); bw_out8(&w, (uint8_t)stream_crc);
            bw_flush(&w);
        }
        if (stats) {
            stats->n_blocks++;
            if (b->info.path_used) stats->n_sais++; else stats->n_native++;
            stats->n_sais_divergent += (uint32_t)b->divergent;
            stats->tie_events += b->info.huf.tie_events; stats->retries += b->info.huf.retries; stats->tie_unpinned += b->info.huf.tie_unpinned;
        }
    }
    if (stats) stats->combined_crc = stream_crc;
    for (uint32_t i = 0; i < nb; i++) { free(blks[i].data); free(blks[i].packed); }
    free(blks);
    if (rc != REF_OK) return rc;
    *out_len = w.len;
    return w.overflow ? REF_ERR_CAP : REF_OK;
}

/* ------------------------------------------------------------------------- */
/* Standard bzip2 single-stream decoder (cross-check of libbz2; the structure  */
/* follows decompress.rs:38-404 but without the reference decoder's defects,   */
/* SURVEY D.6)                                                                */
/* ------------------------------------------------------------------------- */
typedef struct { const uint8_t *p; size_t n; size_t bitpos; int eof; } bitrd;
static uint32_t rd_bits(bitrd *r, int k) {
    uint32_t v = 0;
    for (int i = 0; i < k; i++) {
        size_t byte = r->bitpos >> 3;
        if (byte >= r->n) { r->eof = 1; return 0; }
        v = (v << 1) | ((r->p[byte] >> (7 - (r->bitpos & 7))) & 1);
        r->bitpos++;
    }
    return v;
}
/* reference_semantics != 0: the reference's own decoder behaviour -- rle1_decode with its tail defect (rle1.rs:267-316,
 * SURVEY D.6) and CRC mismatches that are only logged (decompress.rs:376-386, :394-402): they are counted, not fatal. */
static int decompress_impl(const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *out_len, int reference_semantics,
                           uint32_t *n_blocks, uint32_t *n_crc_mismatch) {
    bitrd r = { in, n, 0, 0 };
    if (rd_bits(&r, 8) != 'B' || rd_bits(&r, 8) != 'Z' || rd_bits(&r, 8) != 'h') return REF_ERR_FORMAT;
    int level = (int)rd_bits(&r, 8) - '0';
    if (level < 1 || level > 9) return REF_ERR_FORMAT;
    size_t maxblk = (size_t)level * 100000;
    uint8_t *tt = (uint8_t *)malloc(maxblk + 16), *blk = (uint8_t *)malloc(maxblk + 16);
    size_t o = 0; uint32_t combined = 0; int rc = REF_OK;
    for (;;) {
        uint32_t m1 = rd_bits(&r, 24), m2 = rd_bits(&r, 24);
        if (r.eof) { rc = REF_ERR_FORMAT; break; }
        if (m1 == 0x177245 && m2 == 0x385090) {
            uint32_t sc = rd_bits(&r, 32);
            if (r.eof || sc != combined) rc = REF_ERR_FORMAT;
            break;
        }
        if (m1 != 0x314159 || m2 != 0x265359) { rc = REF_ERR_FORMAT; break; }
        uint32_t bcrc = rd_bits(&r, 32);
        if (rd_bits(&r, 1)) { rc = REF_ERR_FORMAT; break; }   /* randomised blocks unsupported */
        uint32_t key = rd_bits(&r, 24);
        uint32_t l1 = rd_bits(&r, 16); uint8_t seq[256]; int nused = 0;
        for (int i = 0; i < 16; i++) if (l1 & (0x8000u >> i)) {
            uint32_t l2 = rd_bits(&r, 16);
            for (int j = 0; j < 16; j++) if (l2 & (0x8000u >> j)) seq[nused++] = (uint8_t)(i * 16 + j);
        }
        if (nused == 0) { rc = REF_ERR_FORMAT; break; }
        int alpha = nused + 2;
        int T = (int)rd_bits(&r, 3); uint32_t G = rd_bits(&r, 15);
        if (T < 2 || T > 6 || G < 1) { rc = REF_ERR_FORMAT; break; }
        uint8_t *sel = (uint8_t *)malloc(G);
        { uint8_t l6[6] = {0, 1, 2, 3, 4, 5};
          for (uint32_t g = 0; g < G && rc == REF_OK; g++) {
              int j = 0; while (rd_bits(&r, 1)) { j++; if (j >= T) { rc = REF_ERR_FORMAT; break; } }
              if (rc != REF_OK) break;
              uint8_t v = l6[j]; for (int k = j; k > 0; k--) l6[k] = l6[k - 1]; l6[0] = v; sel[g] = v;
          } }
        uint8_t len[6][258];
        for (int t = 0; t < T && rc == REF_OK; t++) {
            int c = (int)rd_bits(&r, 5);
            for (int s = 0; s < alpha; s++) {
                for (;;) {
                    if (c < 1 || c > 20) { rc = REF_ERR_FORMAT; break; }
                    if (!rd_bits(&r, 1)) break;
                    c += rd_bits(&r, 1) ? -1 : 1;
                }
                if (rc != REF_OK) break;
                len[t][s] = (uint8_t)c;
            }
        }
        if (rc != REF_OK || r.eof) { free(sel); rc = REF_ERR_FORMAT; break; }
        int32_t limit[6][22], base[6][22]; uint16_t perm[6][258]; int minl[6];
        for (int t = 0; t < T; t++) {
            int mn = 32, mx = 0; for (int s = 0; s < alpha; s++) { if (len[t][s] < mn) mn = len[t][s]; if (len[t][s] > mx) mx = len[t][s]; }
            minl[t] = mn; int pp = 0;
            for (int l = mn; l <= mx; l++) for (int s = 0; s < alpha; s++) if (len[t][s] == l) perm[t][pp++] = (uint16_t)s;
            int32_t cnt[22] = {0}; for (int s = 0; s < alpha; s++) cnt[len[t][s]]++;
            int32_t code = 0, idx = 0;
            for (int l = 1; l <= 20; l++) { base[t][l] = idx - code; code += cnt[l]; idx += cnt[l]; limit[t][l] = code - 1; code <<= 1; }
            for (int l = 1; l <= 20; l++) if (l < mn || l > mx) limit[t][l] = (l > mx) ? 0x7fffffff : -1;
        }
        /* decode symbols, inverse RLE2/MTF */
        uint32_t nblk = 0, cftab[257]; memset(cftab, 0, sizeof cftab);
        uint32_t runlen = 0, runbit = 1, g = 0, gpos = 50; int t = 0, done = 0;
        while (!done && rc == REF_OK) {
            if (gpos == 50) { if (g >= G) { rc = REF_ERR_FORMAT; break; } t = sel[g++]; gpos = 0; }
            gpos++;
            int l = minl[t]; int32_t code = (int32_t)rd_bits(&r, l);
            while (l <= 20 && code > limit[t][l]) { l++; code = (code << 1) | (int32_t)rd_bits(&r, 1); }
            if (l > 20 || r.eof) { rc = REF_ERR_FORMAT; break; }
            int32_t pi = code + base[t][l];
            if (pi < 0 || pi >= alpha) { rc = REF_ERR_FORMAT; break; }
            uint16_t s = perm[t][pi];
            if (s <= 1) { runlen += runbit << s; runbit <<= 1; continue; }
            if (runlen) {
                if (nblk + runlen > maxblk) { rc = REF_ERR_FORMAT; break; }
                memset(tt + nblk, seq[0], runlen); nblk += runlen; runlen = 0;
            }
            runbit = 1;
            if (s == alpha - 1) { done = 1; break; }
            uint8_t v = seq[s - 1]; memmove(seq + 1, seq, s - 1); seq[0] = v;
            if (nblk + 1 > maxblk) { rc = REF_ERR_FORMAT; break; }
            tt[nblk++] = v;
        }
        free(sel);
        if (rc != REF_OK) break;
        if (key >= nblk) { rc = REF_ERR_FORMAT; break; }
        ref_bwt_decode(key, tt, nblk, blk);
        /* inverse RLE1 + crc */
        size_t before = o;
        size_t got = reference_semantics ? ref_rle1_decode_reference(blk, nblk, out + o, cap - o)
                                         : ref_rle1_decode_standard(blk, nblk, out + o, cap - o);
        if (got == (size_t)-1) { rc = REF_ERR_CAP; break; }
        o += got;
        if (n_blocks) (*n_blocks)++;
        if (ref_do_crc(0, out + before, got) != bcrc) {
            if (!reference_semantics) { rc = REF_ERR_FORMAT; break; }
            if (n_crc_mismatch) (*n_crc_mismatch)++;
        }
        combined = ref_do_stream_crc(combined, bcrc);
    }
    free(tt); free(blk);
    *out_len = o;
    return rc;
}
int ref_decompress_stream(const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *out_len) {
    return decompress_impl(in, n, out, cap, out_len, 0, NULL, NULL);
}
int ref_decompress_stream_reference(const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *out_len,
                                    uint32_t *n_blocks, uint32_t *n_crc_mismatch) {
    if (n_blocks) *n_blocks = 0;
    if (n_crc_mismatch) *n_crc_mismatch = 0;
    return decompress_impl(in, n, out, cap, out_len, 1, n_blocks, n_crc_mismatch);
}

/* ------------------------------------------------------------------------- */
/* Test hooks for the reference's own bit-stream unit vectors                  */
/* (bitwriter.rs:179-210, bitreader.rs:175-240): the stream writer and the    */
/* decoder's bit reader above, driven call by call.                           */
/* ------------------------------------------------------------------------- */
/* BitWriter::out8 for every byte of `bytes`, then flush (bitwriter.rs:135-172); returns the output length */
size_t ref_bw_out8_flush(const uint8_t *bytes, size_t n, uint8_t *out, size_t cap) {
    bitwriter w = { out, 0, cap, 0, 0, 0 };
    for (size_t i = 0; i < n; i++) bw_out8(&w, bytes[i]);
    bw_flush(&w);
    return w.overflow ? (size_t)-1 : w.len;
}
/* BitReader::bint(k) / bit / byte (bitreader.rs:60-150) as a sequence of reads: widths[i] bits each, MSB first;
 * values[i] receives the value, *bitpos_out the position after the last read ("[byte.bit]" of BitReader::loc);
 * returns the number of reads that were served before the input ran out */
size_t ref_br_read_sequence(const uint8_t *in, size_t n, const int *widths, size_t nreads, uint32_t *values,
                            size_t *bitpos_out) {
    bitrd r = { in, n, 0, 0 };
    size_t done = 0;
    for (size_t i = 0; i < nreads; i++) {
        uint32_t v = rd_bits(&r, widths[i]);
        if (r.eof) break;
        values[i] = v; done++;
    }
    if (bitpos_out) *bitpos_out = r.bitpos;
    return done;
}

