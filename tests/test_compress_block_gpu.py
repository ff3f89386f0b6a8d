"""T2: batched compress_block seam (bz2b200_compress_blocks) vs the oracle (compress_block.rs:24-67). Byte exact."""
import bz2

import numpy as np
import pytest

from bzip2_rust_b200 import corpus
from inputs import small_cases

pytestmark = pytest.mark.gpu


def _stream_from_blocks(ref, level, packed):
    """BitWriter::add_block restated in Python for the test (bitwriter.rs:77-132)."""
    bits = []
    crc = 0
    out = bytearray(b"BZh" + bytes([48 + level]))
    acc = 0
    nacc = 0
    for data, pad, bcrc in packed:
        crc = ref.stream_crc(crc, bcrc)
        nb = len(data) * 8 - pad
        v = int.from_bytes(data, "big") >> pad
        acc = (acc << nb) | v
        nacc += nb
    footer = (0x177245385090 << 32) | crc
    acc = (acc << 80) | footer
    nacc += 80
    padb = (8 - nacc % 8) % 8
    acc <<= padb
    out += acc.to_bytes((nacc + padb) // 8, "big")
    return bytes(out)


def test_small_blocks_one_batch(engine, ref):
    blocks, crcs = [], []
    for name, data in small_cases():
        blocks.append(data)
        crcs.append(ref.crc(data))
    got = engine.compress_blocks(blocks, crcs)
    for (name, data), c, (gb, gpad) in zip(small_cases(), crcs, got):
        rb, rpad, _ = ref.compress_block(data, c, ref.SPEC_FAST)
        assert gpad == rpad, name
        assert gb == rb, name


def test_mix1m_level9_stream_identical(engine, ref):
    """BASELINE config 1: 1 MB mixed text/binary at -9 (one 900 kB block + tail)."""
    data = corpus.mix1m(1).tobytes()
    blocks = list(ref.rle1_blocks(data, 9))
    got = engine.compress_blocks([b for _, b, _, _ in blocks], [c for c, _, _, _ in blocks])
    stream = _stream_from_blocks(ref, 9, [(gb, pad, c) for (gb, pad), (c, _, _, _) in zip(got, blocks)])
    want = ref.compress_stream(data, 9, ref.EXACT)
    assert stream == want
    assert bz2.decompress(stream) == data


def test_levels_1_to_9(engine, ref):
    data = corpus.mixed(2_000_000, 5).tobytes()
    for level in range(1, 10):
        blocks = list(ref.rle1_blocks(data, level))
        got = engine.compress_blocks([b for _, b, _, _ in blocks], [c for c, _, _, _ in blocks])
        stream = _stream_from_blocks(ref, level, [(gb, pad, c) for (gb, pad), (c, _, _, _) in zip(got, blocks)])
        assert stream == ref.compress_stream(data, level, ref.SPEC_FAST), level
    assert bz2.decompress(stream) == data


def test_batch_size_does_not_change_bytes(engine, ref):
    data = corpus.text(1_500_000, 8).tobytes()
    blocks = list(ref.rle1_blocks(data, 3))
    bl = [b for _, b, _, _ in blocks]
    cr = [c for c, _, _, _ in blocks]
    all_at_once = engine.compress_blocks(bl, cr)
    one_by_one = [engine.compress_block(b, c) for b, c in zip(bl, cr)]
    assert all_at_once == one_by_one
