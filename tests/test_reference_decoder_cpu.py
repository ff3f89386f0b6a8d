"""north_star: streams "must round-trip through both the reference decoder and the new one".  The reference's decoder
(decompress.rs:38-404) is restated in the oracle with its own semantics: rle1_decode as written (rle1.rs:267-316) and CRC
mismatches that are only logged (decompress.rs:376-386) -- here they are COUNTED.  SURVEY D.6: that rle1_decode does not
expand a run group that sits in a block's last five bytes with the wrong cursor parity, so some valid streams come back
wrong from the reference itself; libbz2 is the decode authority.  This test runs the restated reference decoder over the
parity corpora (streams of the oracle encoder, i.e. the bytes the GPU engine is tested to produce) and lists the
exceptions: none on the corpora, the documented one on the SURVEY example."""
import pytest

from bzip2_rust_b200 import corpus

CASES = [
    ("text2m", lambda: corpus.text(2_000_000, 2).tobytes()),
    ("mix1m", lambda: corpus.mix1m(1).tobytes()),
    ("rep3m", lambda: corpus.repetitive(3_000_000, 3).tobytes()),
    ("mixed3m", lambda: corpus.mixed(3_000_000, 5).tobytes()),
    ("walk1m", lambda: corpus.random_walk(1_000_000, 8).tobytes()),
    ("markov2m", lambda: corpus.markov(2_000_000, 6).tobytes()),
]


@pytest.mark.parametrize("name,make", CASES)
def test_parity_corpora_round_trip_through_the_reference_decoder(ref, name, make):
    data = make()
    for level in (1, 9):
        stream = ref.compress_stream(data, level, ref.SPEC_FAST, threads=4)
        out, blocks, logged = ref.decompress_stream_reference(stream, cap=len(data) * 2 + 1024)
        assert blocks >= 1
        assert logged == 0, "%s level %d: the reference decoder gets %d of %d blocks wrong" % (name, level, logged, blocks)
        assert out == data
        assert ref.decompress_stream(stream, cap=len(data) + 1024) == data       # the standard decoder agrees


def test_known_exception_run_at_the_block_tail(ref):
    """`xyaaaaaa` -> RLE1 `xyaaaa\\x02`; the group lies in the last five bytes with even parity, the
    reference's rle1_decode copies it literally: 7 bytes come back instead of 8 and the block CRC does not match
    (which the reference only logs)."""
    data = b"xyaaaaaa"
    stream = ref.compress_stream(data, 9, ref.SPEC)
    out, blocks, logged = ref.decompress_stream_reference(stream)
    assert (blocks, logged) == (1, 1) and out == b"xyaaaa\x02"
    assert ref.decompress_stream(stream) == data                                  # libbz2 semantics: correct
    import bz2
    assert bz2.decompress(stream) == data


def test_short_and_run_heavy_inputs_listed(ref):
    """Fuzz-sized view of D.6: single-block inputs made of short runs; the reference decoder fails on a minority, always
    at the tail, and never when the standard decoder (== libbz2) fails."""
    import numpy as np
    rng = np.random.default_rng(11)
    bad = 0
    total = 200
    for k in range(total):
        parts = []
        for _ in range(int(rng.integers(1, 6))):
            parts.append(bytes([int(rng.integers(97, 100))]) * int(rng.integers(1, 9)))
        data = b"".join(parts)
        if len(data) >= 4 and data[-4:] == data[-1:] * 4 and (len(data) == 4 or data[-5] != data[-1]):
            continue                                                              # exact-4 tail: reference encoder defect D.4
        try:
            stream = ref.compress_stream(data, 9, ref.SPEC)
        except ref.RefPanic:
            continue                                                              # the reference ENCODER panics here (D.4)
        try:
            std = ref.decompress_stream(stream)
        except RuntimeError:
            continue                                                              # ... or emits an invalid stream (D.4)
        assert std == data
        out, _, logged = ref.decompress_stream_reference(stream)
        if out != data:
            bad += 1
            assert logged == 1
            assert out[:max(0, len(out) - 6)] == data[:max(0, len(out) - 6)]     # the damage is confined to the tail
    assert 0 < bad < total // 2
