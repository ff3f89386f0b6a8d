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
