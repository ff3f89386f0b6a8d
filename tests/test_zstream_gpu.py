"""bz2b200_zstream_*: `compress` (compress.rs:40-136) as a pipeline -- input in pieces, output through a sink.  Whatever
the piece sizes and the window size, the sink receives the bytes of the whole-buffer call (hence the oracle's)."""
import bz2
import io

import numpy as np
import pytest

import bzip2_rust_b200 as bz
from bzip2_rust_b200 import corpus

pytestmark = pytest.mark.gpu


def _stream(engine, data, level, pieces):
    out = io.BytesIO()
    with bz.ZStream(engine, out, level) as z:
        pos = 0
        for p in pieces:
            z.write(data[pos:pos + p])
            pos += p
        if pos < len(data):
            z.write(data[pos:])
    assert z.total_in == len(data) and z.total_out == len(out.getvalue())
    return out.getvalue()


def test_pieces_and_windows_give_the_whole_buffer_stream(engine, monkeypatch):
    rng = np.random.default_rng(8)
    data = np.concatenate([corpus.text(9_000_000, 61), corpus.repetitive(9_000_000, 62), corpus.random_bytes(3_000_000, 63),
                           corpus.text(2_500_000, 64)]).tobytes()
    for level in (9, 2):
        want = engine.compress(data, level)
        for win_mb in (64, 4, 1):                               # 1 MiB windows: every level-9 block needs several windows
            monkeypatch.setenv("BZ2B200_ZSTREAM_WINDOW_MB", str(win_mb))
            sizes = [int(x) for x in rng.integers(1, 3_000_000, 40)]
            got = _stream(engine, data, level, sizes)
            assert got == want, (level, win_mb)
    assert bz2.decompress(want) == data


def test_tiny_empty_and_single_call(engine, monkeypatch):
    monkeypatch.setenv("BZ2B200_ZSTREAM_WINDOW_MB", "2")
    for data in (b"", b"a", b"hello world\n", b"aaaa", b"q" * 300_000, corpus.text(5_000_000, 65).tobytes()):
        want = engine.compress(data, 9)
        assert _stream(engine, data, 9, []) == want                                    # one write
        assert _stream(engine, data, 9, [1] * min(len(data), 50)) == want              # byte by byte at the start
    assert bz2.decompress(_stream(engine, b"", 5, [])) == b""


def test_sink_errors_and_bad_arguments(engine):
    class Full(io.RawIOBase):
        def write(self, b):
            raise OSError("disk full")
    with pytest.raises(OSError):
        with bz.ZStream(engine, Full(), 9) as z:
            z.write(corpus.text(100_000, 66).tobytes())
    L = bz.load_library()
    import ctypes as C
    h = C.c_void_p()
    assert L.bz2b200_zstream_open(engine._h, 0, bz.SINK(lambda u, d, n: 0), None, C.byref(h)) == bz.E_ARG
    assert L.bz2b200_zstream_write(None, None, 0) == bz.E_ARG
    # the context is still healthy
    d = corpus.text(1_000_000, 67).tobytes()
    assert bz2.decompress(engine.compress(d, 9)) == d
