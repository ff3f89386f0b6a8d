"""Full-size parity (BASELINE.json configs 2 and 3): the sizes bench.py times, checked end to end.

text100m at -9 is compared byte for byte with the oracle's stream for the WHOLE input (the oracle runs its blocks
on all host cores), and round-trips through libbz2 and through the GPU decoder.  rep256m (long runs, short periods:
up to 15 doubling rounds in the BWT) is checked through the size-independent properties -- libbz2 and own round
trip, stored block CRCs -- plus byte identity on a leading sample, because the oracle needs minutes on it."""
import bz2
import os

import numpy as np
import pytest

from bzip2_rust_b200 import corpus

pytestmark = pytest.mark.gpu


def test_text100m_whole_stream_identical_and_round_trips(engine, ref):
    data = corpus.text(100_000_000, 2)
    raw = data.tobytes()
    got = engine.compress(data, 9)
    want = ref.compress_stream(raw, 9, ref.SPEC_FAST, threads=os.cpu_count() or 1)
    assert len(got) == len(want)
    assert got == want, "stream differs from the oracle (first difference at byte %d)" % next(
        i for i, (a, b) in enumerate(zip(got, want)) if a != b)
    assert bz2.decompress(got) == raw
    assert engine.decompress(got, max_out=len(raw) + 1024) == raw


def test_rep256m_round_trips_and_sample_identical(engine, ref):
    data = corpus.repetitive(256_000_000, 3)
    raw = data.tobytes()
    got = engine.compress(data, 9)
    assert bz2.decompress(got) == raw
    assert engine.decompress(got, max_out=len(raw) + 1024) == raw
    k = 4_000_000
    assert engine.compress(data[:k], 9) == ref.compress_stream(raw[:k], 9, ref.SPEC_FAST, threads=os.cpu_count() or 1)


def test_levels_agree_on_1gb_of_mixed_data_checksum(engine):
    """Config 5 shape at a size the suite can afford (256 MB): every level round-trips through the GPU decoder and
    the decoded bytes hash to the same value (a checksum of checksums over levels)."""
    import zlib
    data = corpus.mixed(256_000_000, 5)
    want = zlib.adler32(data.tobytes())
    for level in (1, 5, 9):
        stream = engine.compress(data, level)
        out = engine.decompress(stream, max_out=data.size + 1024)
        assert zlib.adler32(out) == want, level
