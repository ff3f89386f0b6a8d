"""Round trip: the CUDA decoder (bz2b200_decompress_stream, bz2b200_bwt_decode) against libbz2 and the oracle
(decompress.rs:38-404, bwt_sort.rs:91-130).  libbz2 is the decode authority."""
import bz2

import numpy as np
import pytest

import bzip2_rust_b200 as bz
from bzip2_rust_b200 import corpus
from inputs import small_cases

pytestmark = pytest.mark.gpu


def test_bwt_decode_inverts_encode(engine, ref):
    for name, data in small_cases():
        key, bwt, _ = ref.bwt_encode(data, ref.SPEC_FAST)
        assert engine.bwt_decode(key, bwt) == data, name
        assert ref.bwt_decode(key, bwt) == data, name
    blk = corpus.mix1m(1)[:899_986].tobytes()
    key, bwt = engine.bwt_encode(blk)
    assert engine.bwt_decode(key, bwt) == blk
    per = (b"aaaa\xfb" * 50_000)
    key, bwt = engine.bwt_encode(per)                 # fully periodic block: many LF cycles
    assert engine.bwt_decode(key, bwt) == per
    same = b"z" * 70_000
    key, bwt = engine.bwt_encode(same)
    assert engine.bwt_decode(key, bwt) == same


def test_decompress_own_and_libbz2_streams(engine, ref):
    rng = np.random.default_rng(3)
    inputs = [b"a", b"hello world\n", corpus.mix1m(1).tobytes(), corpus.repetitive(2_000_000, 5).tobytes(),
              corpus.text(1_200_000, 6).tobytes(), bytes(rng.integers(0, 256, 300_000, dtype=np.uint8)),
              b"\x00" * 3_000_000]
    for i, data in enumerate(inputs):
        for level in (1, 9):
            ours = engine.compress(data, level)
            assert engine.decompress(ours) == data, (i, level)
            theirs = bz2.compress(data, level)            # a different encoder's stream (other tables / splits)
            assert engine.decompress(theirs) == data, (i, level)
            assert ref.decompress_stream(ours, cap=len(data) + 1024) == data


def test_empty_stream(engine):
    assert engine.decompress(bz2.compress(b"")) == b""


def test_corruption_is_detected(engine):
    data = corpus.text(300_000, 9).tobytes()
    s = bytearray(engine.compress(data, 1))
    s[len(s) // 2] ^= 0x10
    with pytest.raises(bz.Bz2B200Error):
        engine.decompress(bytes(s))
    with pytest.raises(bz.Bz2B200Error):
        engine.decompress(b"BZh9" + b"\x00" * 20)
    with pytest.raises(bz.Bz2B200Error):
        engine.decompress(b"not a bzip2 stream at all")


def test_roundtrip_all_levels_mixed(engine):
    data = corpus.mixed(3_000_000, 5).tobytes()
    for level in range(1, 10):
        assert engine.decompress(engine.compress(data, level)) == data


def test_concatenated_streams(engine):
    """Several .bz2 streams one after the other (libbz2 / `bzip2 -d` accept that; decompress.rs does not): own streams,
    libbz2's, mixed levels, an empty stream in the middle."""
    a = corpus.text(700_000, 21).tobytes()
    b = corpus.repetitive(400_000, 22).tobytes()
    c = b"tail"
    cat = engine.compress(a, 9) + bz2.compress(b, 3) + bz2.compress(b"") + engine.compress(c, 1)
    assert bz2.decompress(cat) == a + b + c
    assert engine.decompress(cat) == a + b + c
    cat2 = engine.compress(b"x", 1) + engine.compress(a, 5)         # a later stream announces the larger block size
    assert engine.decompress(cat2) == b"x" + a
    with pytest.raises(bz.Bz2B200Error) as e:
        engine.decompress(engine.compress(a, 9) + b"garbage after the stream")
    assert e.value.rc == bz.E_FORMAT


def test_more_blocks_than_one_decoder_batch(engine):
    data = corpus.mixed(24_000_000, 31)                             # level 1: about 240 blocks, two batches of 128
    s = engine.compress(data, 1)
    assert engine.decompress(s, max_out=data.size + 16) == data.tobytes()


def _rand_flips(n):
    """Byte positions a legacy "randomised" block XORs with 1 (BZ2_rNums of the .bz2 format, taken from libbz2 itself)."""
    import ctypes
    t = list((ctypes.c_int32 * 512).in_dll(ctypes.CDLL("libbz2.so.1.0"), "BZ2_rNums"))
    pos, out, k = 0, [], 0
    while pos < n:
        p = pos + t[k % 512] - 2
        if p < n:
            out.append(p)
        pos += t[k % 512]
        k += 1
    return out


def test_randomised_block(engine, ref):
    """bzip2 <= 0.9.0 could mark a block "randomised"; no current encoder does (compress_block.rs:41 writes 0), so the
    vector is built here: flip the RLE1 bytes at the format's pseudo-random positions, compress that block with the
    oracle, set the randomised bit.  libbz2 and the GPU decoder must both undo it."""
    data = corpus.text(300_000, 33).tobytes() + b"q" * 1000 + corpus.random_bytes(5000, 34).tobytes()
    (crc, rle1, last, consumed), = list(ref.rle1_blocks(data, 9))
    r = bytearray(rle1)
    for p in _rand_flips(len(r)):
        r[p] ^= 1
    packed, pad, _ = ref.compress_block(bytes(r), crc, ref.SPEC_FAST)
    pb = bytearray(packed)
    assert (pb[10] >> 7) == 0
    pb[10] |= 0x80                                                  # bit 80: after the 48-bit magic and the 32-bit CRC
    stream = bz.merge_streams(9, [(bytes(pb), len(pb) * 8 - pad, [crc])])
    assert bz2.decompress(stream) == data                          # libbz2 agrees that this IS the stream of `data`
    assert engine.decompress(stream) == data


def test_walk_leaves_its_jump_tables(engine, monkeypatch):
    """A chance block magic inside a block's data cuts the range the jump tables of k_dec_jumps cover (decode.cu); the
    walk must then finish code by code with the same result.  BZ2B200_DEC_RANGE_LIMIT cuts every range artificially:
    inside the first window, in the middle of a block, and at zero bits (no tables at all)."""
    data = corpus.text(1_500_000, 33).tobytes() + bytes(np.random.default_rng(5).integers(0, 256, 200_000, dtype=np.uint8))
    streams = [engine.compress(data, 9), bz2.compress(data, 3)]
    for limit in ("1", "700", "100000", "1500001"):
        monkeypatch.setenv("BZ2B200_DEC_RANGE_LIMIT", limit)
        for s in streams:
            assert engine.decompress(s) == data, limit
    monkeypatch.delenv("BZ2B200_DEC_RANGE_LIMIT")
    assert engine.decompress(streams[0]) == data
