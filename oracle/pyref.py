"""ctypes binding of the CPU oracle (oracle/libbz2ref.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libbz2ref.so")

SPEC, SPEC_FAST, EXACT = 0, 1, 2
OK, ERR_PANIC, ERR_CAP, ERR_ARG, ERR_FORMAT = 0, -1, -2, -3, -4


def build(force=False):
    src = os.path.join(_HERE, "bz2ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


class HufInfo(C.Structure):
    _fields_ = [("table_count", C.c_int), ("selector_count", C.c_uint32),
                ("lengths", (C.c_uint8 * 258) * 6), ("selectors", C.POINTER(C.c_uint8)),
                ("tie_events", C.c_uint32), ("retries", C.c_uint32), ("tie_unpinned", C.c_uint32)]


class BlockInfo(C.Structure):
    _fields_ = [("key", C.c_uint32), ("path_used", C.c_int), ("m", C.c_uint32), ("huf", HufInfo)]


class StreamStats(C.Structure):
    _fields_ = [("n_blocks", C.c_uint32), ("n_native", C.c_uint32), ("n_sais", C.c_uint32),
                ("n_sais_divergent", C.c_uint32), ("tie_events", C.c_uint32), ("retries", C.c_uint32),
                ("combined_crc", C.c_uint32), ("tie_unpinned", C.c_uint32)]


class BitPacker(C.Structure):
    _fields_ = [("out", C.c_void_p), ("len", C.c_size_t), ("cap", C.c_size_t), ("queue", C.c_uint64),
                ("q_bits", C.c_uint32), ("padding", C.c_uint8), ("overflow", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.ref_do_crc.restype = C.c_uint32
        L.ref_do_crc.argtypes = [C.c_uint32, C.c_char_p, C.c_size_t]
        L.ref_do_stream_crc.restype = C.c_uint32
        L.ref_do_stream_crc.argtypes = [C.c_uint32, C.c_uint32]
        L.ref_rle1_new.restype = C.c_void_p
        L.ref_rle1_new.argtypes = [C.c_char_p, C.c_size_t, C.c_size_t]
        L.ref_rle1_free.argtypes = [C.c_void_p]
        L.ref_rle1_next.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_void_p),
                                    C.POINTER(C.c_size_t), C.POINTER(C.c_int), C.POINTER(C.c_size_t)]
        L.ref_rle1_decode_reference.restype = C.c_size_t
        L.ref_rle1_decode_reference.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.ref_rle1_decode_standard.restype = C.c_size_t
        L.ref_rle1_decode_standard.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.ref_bwt_encode.argtypes = [C.c_char_p, C.c_uint32, C.c_int, C.POINTER(C.c_uint32), C.c_void_p,
                                     C.POINTER(C.c_int)]
        L.ref_lms_count.restype = C.c_uint32
        L.ref_lms_count.argtypes = [C.c_char_p, C.c_uint32]
        L.ref_duval.restype = C.c_uint32
        L.ref_duval.argtypes = [C.c_char_p, C.c_uint32]
        L.ref_lms_types.argtypes = [C.c_char_p, C.c_uint32, C.c_void_p, C.c_void_p]
        L.ref_bucket_sizes_heads_tails.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                                                   C.c_void_p]
        L.ref_bwt_decode.argtypes = [C.c_uint32, C.c_char_p, C.c_uint32, C.c_void_p]
        L.ref_rle2_mtf_encode.argtypes = [C.c_char_p, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32), C.c_void_p,
                                          C.c_void_p, C.POINTER(C.c_int)]
        L.ref_bp_init.argtypes = [C.POINTER(BitPacker), C.c_void_p, C.c_size_t]
        L.ref_bp_out24.argtypes = [C.POINTER(BitPacker), C.c_uint32]
        L.ref_bp_out32.argtypes = [C.POINTER(BitPacker), C.c_uint32]
        L.ref_bp_out16.argtypes = [C.POINTER(BitPacker), C.c_uint16]
        L.ref_bp_flush.argtypes = [C.POINTER(BitPacker)]
        L.ref_huf_encode.argtypes = [C.POINTER(BitPacker), C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint16,
                                     C.c_void_p, C.c_int, C.POINTER(HufInfo)]
        L.ref_improve_code_len.argtypes = [C.c_void_p, C.c_void_p, C.c_uint16, C.POINTER(C.c_uint32),
                                           C.POINTER(C.c_uint32)]
        L.ref_init_tables.argtypes = [C.c_void_p, C.c_int, C.c_uint16, C.c_void_p]
        L.ref_compress_block.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_size_t,
                                         C.POINTER(C.c_size_t), C.POINTER(C.c_uint8), C.POINTER(BlockInfo)]
        L.ref_compress_stream.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                          C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(StreamStats)]
        L.ref_decompress_stream.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                            C.POINTER(C.c_size_t)]
        _lib = L
    return _lib


class RefPanic(Exception):
    """The reference would panic (or emit an invalid stream) on this input."""


def _check(rc, what):
    if rc == ERR_PANIC:
        raise RefPanic(what)
    if rc != OK:
        raise RuntimeError("%s failed: %d" % (what, rc))


def crc(data, existing=0):
    return lib().ref_do_crc(existing, bytes(data), len(data))


def stream_crc(s, b):
    return lib().ref_do_stream_crc(s, b)


def rle1_blocks(data, level=9):
    """Yields (crc, rle1_block_bytes, last, consumed_input_bytes) like RLE1Block (rle1.rs:245-264)."""
    data = bytes(data)
    L = lib()
    it = L.ref_rle1_new(data, len(data), level * 100000 - 19)
    try:
        while True:
            c, p, n, last, cons = C.c_uint32(), C.c_void_p(), C.c_size_t(), C.c_int(), C.c_size_t()
            r = L.ref_rle1_next(it, C.byref(c), C.byref(p), C.byref(n), C.byref(last), C.byref(cons))
            if r == 0:
                return
            _check(r if r < 0 else OK, "rle1_next")
            yield c.value, C.string_at(p.value, n.value) if n.value else b"", bool(last.value), cons.value
    finally:
        L.ref_rle1_free(it)


def rle1_decode_standard(b):
    buf = C.create_string_buffer(len(b) * 52 + 16)
    n = lib().ref_rle1_decode_standard(bytes(b), len(b), buf, len(buf))
    return buf.raw[:n]


def rle1_decode_reference(b):
    buf = C.create_string_buffer(len(b) * 52 + 16)
    n = lib().ref_rle1_decode_reference(bytes(b), len(b), buf, len(buf))
    return buf.raw[:n]


def bwt_encode(block, mode=SPEC):
    block = bytes(block)
    key, path = C.c_uint32(), C.c_int()
    out = C.create_string_buffer(max(len(block), 1))
    _check(lib().ref_bwt_encode(block, len(block), mode, C.byref(key), out, C.byref(path)), "bwt_encode")
    return key.value, out.raw[:len(block)], path.value


def bwt_decode(key, bwt):
    out = C.create_string_buffer(max(len(bwt), 1))
    _check(lib().ref_bwt_decode(key, bytes(bwt), len(bwt), out), "bwt_decode")
    return out.raw[:len(bwt)]


def lms_count(data):
    return lib().ref_lms_count(bytes(data), len(data))


def duval(data):
    return lib().ref_duval(bytes(data), len(data))


def lms_types(data):
    n = len(data)
    s, l = C.create_string_buffer(n + 1), C.create_string_buffer(n + 1)
    lib().ref_lms_types(bytes(data), n, s, l)
    return list(s.raw[:n + 1]), list(l.raw[:n + 1])


def buckets(data, size):
    n = len(data)
    d = (C.c_uint32 * n)(*data)
    a, b, c = (C.c_uint32 * size)(), (C.c_uint32 * size)(), (C.c_uint32 * size)()
    lib().ref_bucket_sizes_heads_tails(d, n, size, a, b, c)
    return list(a), list(b), list(c)


def rle2_mtf_encode(bwt):
    """-> (symbols list[u16] incl. EOB, freq[256], symmap list[u16])  (rle2_mtf.rs:23)."""
    import numpy as np
    bwt = bytes(bwt)
    n = len(bwt)
    sym = np.zeros(n + 1, dtype=np.uint16)
    freq = np.zeros(256, dtype=np.uint32)
    smap = np.zeros(17, dtype=np.uint16)
    m, nmap = C.c_uint32(), C.c_int()
    _check(lib().ref_rle2_mtf_encode(bwt, n, sym.ctypes.data, C.byref(m), freq.ctypes.data, smap.ctypes.data,
                                     C.byref(nmap)), "rle2_mtf_encode")
    return sym[:m.value].copy(), freq, smap[:nmap.value].copy()


def huf_encode(sym, freq, symmap):
    """Runs huf_encode (huffman.rs:79) on a fresh BitPacker. -> (bytes, n_bits, info dict)."""
    import numpy as np
    sym = np.ascontiguousarray(sym, dtype=np.uint16)
    freq = np.ascontiguousarray(freq, dtype=np.uint32)
    symmap = np.ascontiguousarray(symmap, dtype=np.uint16)
    m = len(sym)
    eob = int(sym[-1])
    cap = m * 3 + 8192
    buf = C.create_string_buffer(cap)
    bp = BitPacker()
    L = lib()
    L.ref_bp_init(C.byref(bp), buf, cap)
    info = HufInfo()
    G = (m + 49) // 50
    selbuf = (C.c_uint8 * max(G, 1))()
    info.selectors = C.cast(selbuf, C.POINTER(C.c_uint8))
    _check(L.ref_huf_encode(C.byref(bp), sym.ctypes.data, m, freq.ctypes.data, eob, symmap.ctypes.data,
                            len(symmap), C.byref(info)), "huf_encode")
    nbits = bp.len * 8 + bp.q_bits
    L.ref_bp_flush(C.byref(bp))
    lengths = np.array([[info.lengths[t][s] for s in range(258)] for t in range(6)], dtype=np.uint8)
    return buf.raw[:bp.len], nbits, dict(table_count=info.table_count, selectors=bytes(selbuf[:G]),
                                         lengths=lengths, tie_events=info.tie_events, retries=info.retries)


def improve_code_len(weights, eob):
    import numpy as np
    codes = np.zeros(258, dtype=np.uint32)
    w = np.zeros(258, dtype=np.uint32)
    w[:len(weights)] = weights
    ties, retries = C.c_uint32(0), C.c_uint32(0)
    lib().ref_improve_code_len(codes.ctypes.data, w.ctypes.data, eob, C.byref(ties), C.byref(retries))
    return codes[:eob + 1].copy(), ties.value, retries.value


def init_tables(freq, table_count, eob):
    import numpy as np
    f = np.ascontiguousarray(freq, dtype=np.uint32)
    t = np.zeros((6, 258), dtype=np.uint32)
    lib().ref_init_tables(f.ctypes.data, table_count, eob, t.ctypes.data)
    return t


def compress_block(block, block_crc, mode=SPEC):
    """compress_block (compress_block.rs:24) -> (packed bytes, padding, info)."""
    block = bytes(block)
    cap = len(block) * 3 // 2 + 4096
    buf = C.create_string_buffer(cap)
    n, pad, info = C.c_size_t(), C.c_uint8(), BlockInfo()
    _check(lib().ref_compress_block(block, len(block), block_crc, mode, buf, cap, C.byref(n), C.byref(pad),
                                    C.byref(info)), "compress_block")
    return buf.raw[:n.value], pad.value, dict(key=info.key, path=info.path_used, m=info.m,
                                              ties=info.huf.tie_events, tables=info.huf.table_count)


def compress_stream(data, level=9, mode=SPEC, threads=1, want_stats=False):
    data = bytes(data)
    cap = len(data) * 3 // 2 + 65536
    buf = C.create_string_buffer(cap)
    n = C.c_size_t()
    st = StreamStats()
    _check(lib().ref_compress_stream(data, len(data), level, mode, threads, buf, cap, C.byref(n),
                                     C.byref(st) if want_stats else None), "compress_stream")
    out = buf.raw[:n.value]
    if want_stats:
        return out, {k: getattr(st, k) for k, _ in StreamStats._fields_}
    return out


def decompress_stream(data, cap=None):
    data = bytes(data)
    cap = cap or max(len(data) * 60, 1 << 20)
    buf = C.create_string_buffer(cap)
    n = C.c_size_t()
    rc = lib().ref_decompress_stream(data, len(data), buf, cap, C.byref(n))
    if rc != OK:
        raise RuntimeError("ref_decompress_stream failed: %d" % rc)
    return buf.raw[:n.value]


def decompress_stream_reference(data, cap=None):
    """The reference's decoder semantics (decompress.rs:38-404 + rle1.rs:267-316) -> (bytes, blocks, blocks whose CRC the
    reference would only have logged as wrong)."""
    data = bytes(data)
    cap = cap or max(len(data) * 60, 1 << 20)
    buf = C.create_string_buffer(cap)
    n = C.c_size_t()
    nb, bad = C.c_uint32(), C.c_uint32()
    L = lib()
    L.ref_decompress_stream_reference.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t),
                                                  C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    rc = L.ref_decompress_stream_reference(data, len(data), buf, cap, C.byref(n), C.byref(nb), C.byref(bad))
    if rc != OK:
        raise RuntimeError("ref_decompress_stream_reference failed: %d" % rc)
    return buf.raw[:n.value], nb.value, bad.value


class Packer:
    """BitPacker (bitpacker.rs:17-112) for the reference's own unit vectors."""

    def __init__(self, cap=1024):
        self._buf = C.create_string_buffer(cap)
        self._bp = BitPacker()
        lib().ref_bp_init(C.byref(self._bp), self._buf, cap)

    def out24(self, v): lib().ref_bp_out24(C.byref(self._bp), v)
    def out32(self, v): lib().ref_bp_out32(C.byref(self._bp), v)
    def out16(self, v): lib().ref_bp_out16(C.byref(self._bp), v)
    def flush(self): lib().ref_bp_flush(C.byref(self._bp))
    @property
    def output(self): return self._buf.raw[:self._bp.len]
    @property
    def padding(self): return self._bp.padding
    def loc(self):
        b = self._bp.len * 8 + self._bp.q_bits
        return "[%d.%d]" % (b // 8, b % 8)


def bw_out8_flush(data):
    """BitWriter::out8 for every byte, then flush (bitwriter.rs:135-172): the reference's bitwriter unit vectors."""
    data = bytes(data)
    buf = C.create_string_buffer(len(data) + 8)
    L = lib()
    L.ref_bw_out8_flush.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t]
    L.ref_bw_out8_flush.restype = C.c_size_t
    n = L.ref_bw_out8_flush(data, len(data), buf, len(data) + 8)
    return buf.raw[:n]


def br_read(data, widths):
    """The decoder's bit reader driven like BitReader::bit / bint / byte (bitreader.rs:52-150): one read per width,
    MSB first -> (values served before the input ran out, "[byte.bit]" position as BitReader::loc prints it)."""
    data = bytes(data)
    w = (C.c_int * len(widths))(*widths)
    v = (C.c_uint32 * len(widths))()
    pos = C.c_size_t()
    L = lib()
    L.ref_br_read_sequence.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_int), C.c_size_t, C.POINTER(C.c_uint32),
                                       C.POINTER(C.c_size_t)]
    L.ref_br_read_sequence.restype = C.c_size_t
    done = L.ref_br_read_sequence(data, len(data), w, len(widths), v, C.byref(pos))
    return [int(v[i]) for i in range(done)], "[%d.%d]" % (pos.value // 8, pos.value % 8)
