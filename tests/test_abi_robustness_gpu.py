"""T3: robustness at the C ABI -- tiny inputs, exact-4 tails, too-small output buffers, bad pointers.
Every failure is an error code (and the context keeps working); nothing crashes or unwinds across the boundary.
The reference panics or writes invalid streams on several of these; libbz2 is the authority for what the
bytes must decode to."""
import bz2
import ctypes as C

import numpy as np
import pytest
import torch

import bzip2_rust_b200 as bz
from bzip2_rust_b200 import corpus

pytestmark = pytest.mark.gpu


def test_tiny_inputs_and_exact4_tails(engine, ref):
    for data in (b"", b"a", b"ab", b"abc", b"aaaa", b"xaaaa", b"xyaaaa", b"xyyy", b"aaaaxyyy", b"\x00", b"\xff" * 4,
                 b"abcd" * 3 + b"zzzz", b"q" * 259, b"q" * 260, b"q" * 255 + b"r" * 4):
        for level in (1, 9):
            s = engine.compress(data, level)
            assert bz2.decompress(s) == data, (data[:16], level)
            assert engine.decompress(s) == data
    # where the reference is well defined (no exact-4 tail, not empty) the bytes are the oracle's
    for data in (b"a", b"ab", b"abc", b"abcd", b"aaaaa", b"hello world\n"):
        assert engine.compress(data, 9) == ref.compress_stream(data, 9, ref.SPEC)


def test_output_too_small_is_E_CAP(engine):
    L = bz.load_library()
    data = corpus.text(3_000_000, 3)
    n = C.c_size_t()
    for cap in (0, 8, 64, 1000, 100_000):
        out = np.zeros(max(cap, 1), dtype=np.uint8)
        rc = L.bz2b200_compress_stream(engine._h, data.ctypes.data, data.size, 9, out.ctypes.data, cap, C.byref(n))
        assert rc == bz.E_CAP, cap
    stream = np.frombuffer(engine.compress(data, 9), dtype=np.uint8)
    for cap in (0, 1, 4096, data.size - 1):
        out = np.zeros(max(cap, 1), dtype=np.uint8)
        rc = L.bz2b200_decompress_stream(engine._h, stream.ctypes.data, stream.size, out.ctypes.data, cap, C.byref(n))
        assert rc == bz.E_CAP, cap
        assert n.value >= min(data.size, 1)                         # the size that was needed (so far) is reported
    # compress_blocks: one of two output buffers too small
    blk = [bytes(corpus.text(50_000, 4)), bytes(corpus.random_bytes(50_000, 5))]
    arrs = [np.frombuffer(b, dtype=np.uint8) for b in blk]
    outs = [np.zeros(100_000, dtype=np.uint8), np.zeros(16, dtype=np.uint8)]
    lens = np.array([a.size for a in arrs], dtype=np.uint32)
    crcs = np.array([engine.crc32(b) for b in blk], dtype=np.uint32)
    caps = np.array([o.size for o in outs], dtype=np.uint64)
    bits = np.zeros(2, dtype=np.uint64)
    inp = (C.c_void_p * 2)(*[a.ctypes.data for a in arrs])
    outp = (C.c_void_p * 2)(*[o.ctypes.data for o in outs])
    rc = L.bz2b200_compress_blocks(engine._h, 2, inp, lens.ctypes.data, crcs.ctypes.data, outp, caps.ctypes.data, bits.ctypes.data)
    assert rc == bz.E_CAP
    assert engine.decompress(engine.compress(data, 9), max_out=data.size + 8) == data.tobytes()    # still healthy


def test_bad_arguments(engine):
    L = bz.load_library()
    n = C.c_size_t()
    out = np.zeros(4096, dtype=np.uint8)
    data = corpus.text(10_000, 6)
    assert L.bz2b200_compress_stream(None, data.ctypes.data, data.size, 9, out.ctypes.data, out.size, C.byref(n)) == bz.E_ARG
    assert L.bz2b200_compress_stream(engine._h, None, 10, 9, out.ctypes.data, out.size, C.byref(n)) == bz.E_ARG
    assert L.bz2b200_compress_stream(engine._h, data.ctypes.data, data.size, 0, out.ctypes.data, out.size, C.byref(n)) == bz.E_ARG
    assert L.bz2b200_compress_stream(engine._h, data.ctypes.data, data.size, 10, out.ctypes.data, out.size, C.byref(n)) == bz.E_ARG
    assert L.bz2b200_compress_stream(engine._h, data.ctypes.data, data.size, 9, None, out.size, C.byref(n)) == bz.E_ARG
    assert L.bz2b200_decompress_stream(engine._h, data.ctypes.data, data.size, out.ctypes.data, out.size, C.byref(n)) == bz.E_FORMAT
    assert L.bz2b200_decompress_stream(engine._h, data.ctypes.data, 3, out.ctypes.data, out.size, C.byref(n)) == bz.E_ARG
    # block seam: empty and oversized blocks
    big = np.zeros(bz.MAX_BLOCK + 1, dtype=np.uint8)
    for arr in (np.zeros(0, dtype=np.uint8), big):
        lens = np.array([arr.size], dtype=np.uint32)
        crc = np.zeros(1, dtype=np.uint32)
        caps = np.array([out.size], dtype=np.uint64)
        bits = np.zeros(1, dtype=np.uint64)
        inp = (C.c_void_p * 1)(arr.ctypes.data if arr.size else out.ctypes.data)
        outp = (C.c_void_p * 1)(out.ctypes.data)
        assert L.bz2b200_compress_blocks(engine._h, 1, inp, lens.ctypes.data, crc.ctypes.data, outp, caps.ctypes.data,
                                         bits.ctypes.data) == bz.E_ARG
    # device entry points: output pointers must be 4-byte aligned (the bit merge works on 32-bit words)
    d_in = torch.from_numpy(data).cuda()
    d_out = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    assert L.bz2b200_compress_stream_dev(engine._h, d_in.data_ptr(), data.size, 9, d_out.data_ptr() + 1, 1 << 19, C.byref(n)) == bz.E_ARG
    assert L.bz2b200_shift_bits_dev(engine._h, d_out.data_ptr() + 2, 100, 3, d_out.data_ptr() + 4096) == bz.E_ARG
    assert L.bz2b200_compress_stream_dev(engine._h, d_in.data_ptr(), data.size, 9, d_out.data_ptr(), 1 << 20, C.byref(n)) == bz.OK
    assert bz2.decompress(d_out[:n.value].cpu().numpy().tobytes()) == data.tobytes()


def test_contexts_come_and_go(engine):
    """Creating and destroying contexts (and multi contexts) leaves no state behind: a recycled address starts clean."""
    data = corpus.text(2_000_000, 7)
    want = engine.compress(data, 9)
    for _ in range(3):
        e = bz.Engine(0)
        assert e.compress(data, 9) == want
        e.close()
        m = bz.MultiEngine([0, 0])
        assert m.compress(data, 9) == want
        m.close()
