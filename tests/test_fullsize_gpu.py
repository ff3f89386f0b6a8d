"""Full-size parity (BASELINE.json configs 2-5): the sizes bench.py and tools/run_configs.py time, checked end to end.

Every stream is compared byte for byte with the CPU oracle's stream for the WHOLE input (the oracle runs its blocks on
all host cores) and round-trips through the GPU decoder (and libbz2 where that is affordable):
  config 2  text100m            100 MB text, level 9
  config 3  rep256m             256 MB of long runs / short periods (up to 15 doubling rounds in the BWT), level 9
  config 4  corpus8g            8 GB, level 9, one GPU and the multi-rank path (ranks share the GPU here; distinct GPUs
                                run under bench.py / tools/run_corpus8g.py): same bytes for every rank count
  config 5  sweep1g             1 GB random+text, all nine levels
The oracle needs about 25 ms per MB and core: on a box with fewer than 24 cores corpus8g is compared on its leading
4 GB (blocks are independent, so the stream prefix up to the last block of the prefix is the oracle's prefix stream)."""
import bz2
import os
import zlib

import numpy as np
import pytest

import bzip2_rust_b200 as bz
from bzip2_rust_b200 import corpus

pytestmark = pytest.mark.gpu

CORES = os.cpu_count() or 1
WORKERS = corpus.default_workers()


def _first_diff(a, b):
    n = min(len(a), len(b))
    x = np.frombuffer(a, dtype=np.uint8, count=n) != np.frombuffer(b, dtype=np.uint8, count=n)
    i = int(np.argmax(x)) if x.any() else n
    return i


def test_text100m_whole_stream_identical_and_round_trips(engine, ref):
    data = corpus.text(100_000_000, 2)
    raw = data.tobytes()
    got = engine.compress(data, 9)
    want = ref.compress_stream(raw, 9, ref.SPEC_FAST, threads=CORES)
    assert len(got) == len(want)
    assert got == want, "stream differs from the oracle (first difference at byte %d)" % _first_diff(got, want)
    assert bz2.decompress(got) == raw
    assert engine.decompress(got, max_out=len(raw) + 1024) == raw


def test_markov64m_whole_stream_identical(engine, ref):
    """Higher-order text (word transitions from a sparse table): deeper contexts than text100m, more doubling rounds."""
    data = corpus.markov(64_000_000, 6, workers=WORKERS)
    got = engine.compress(data, 9)
    want = ref.compress_stream(data.tobytes(), 9, ref.SPEC_FAST, threads=CORES)
    assert got == want, "first difference at byte %d" % _first_diff(got, want)
    assert engine.decompress(got, max_out=data.size + 1024) == data.tobytes()


def test_rep256m_whole_stream_identical_and_round_trips(engine, ref):
    data = corpus.rep_segments(256_000_000, 3, workers=WORKERS)
    raw = data.tobytes()
    got = engine.compress(data, 9)
    want = ref.compress_stream(raw, 9, ref.SPEC_FAST, threads=CORES)
    assert got == want, "first difference at byte %d" % _first_diff(got, want)
    assert bz2.decompress(got) == raw
    assert engine.decompress(got, max_out=len(raw) + 1024) == raw


def test_sweep1g_every_level_identical_and_round_trips(engine, ref):
    """Config 5 at its stated size: 1 GB of random+text, levels 1..9, every stream == the oracle's; the decoded bytes
    of every level hash to the input's checksum (a checksum of checksums over the levels)."""
    data = corpus.mixed(1_000_000_000, 5, workers=WORKERS)
    raw = data.tobytes()
    want_sum = zlib.adler32(raw)
    for level in range(1, 10):
        got = engine.compress(data, level)
        want = ref.compress_stream(raw, level, ref.SPEC_FAST, threads=CORES)
        assert got == want, "level %d: first difference at byte %d" % (level, _first_diff(got, want))
        out = engine.decompress(got, max_out=data.size + 1024)
        assert zlib.adler32(out) == want_sum, level


def test_corpus8g_identical_for_one_and_several_ranks(engine, ref):
    """Config 4: 8 GB at level 9.  One GPU context, and the library's multi-rank scheduler (30-32 windows dealt round
    robin to 2 and 3 ranks that share the GPU, or to 2 and 8 distinct GPUs), give the same bytes; those bytes are the
    oracle's."""
    n = 8_000_000_000
    data = corpus.corpus(n, 4, workers=WORKERS)
    got = engine.compress(data, 9)
    engine.trim()                                               # ranks that share this GPU need the room (20 GB per rank)
    for world in ((2, 8) if _ngpu() >= 8 else (2, 3)):
        m = bz.MultiEngine([r % max(1, _ngpu()) for r in range(world)])
        try:
            assert m.compress(data, 9) == got, "world %d differs from the single-context stream" % world
        finally:
            m.close()
    p = n if CORES >= 24 else 4_000_000_000
    want = ref.compress_stream(data[:p].tobytes(), 9, ref.SPEC_FAST, threads=CORES)
    k = len(want) if p == n else len(want) - 2_000_000          # the prefix stream's last block ends at ITS end of input
    assert got[:k] == want[:k], "first difference at byte %d" % _first_diff(got[:k], want[:k])
    if p == n:
        assert len(got) == len(want)
    out = engine.decompress(got, max_out=n + 1024)              # > 4 GiB of output: the decoder works window by window
    assert len(out) == n and zlib.adler32(out) == zlib.adler32(data.tobytes())


def _ngpu():
    import torch
    return torch.cuda.device_count()
