"""T1: CUDA MTF+RLE2 (bz2b200_mtf_rle2) vs the oracle's rle2_mtf_encode restatement (rle2_mtf.rs:23-177)."""
import numpy as np
import pytest

from bzip2_rust_b200 import corpus
from inputs import small_cases

pytestmark = pytest.mark.gpu


def _check(engine, ref, bwt, name=""):
    sym, freq, smap = engine.rle2_mtf_encode(bwt)
    rs, rf, rm = ref.rle2_mtf_encode(bwt)
    assert len(sym) == len(rs), "%s: m differs %d vs %d" % (name, len(sym), len(rs))
    assert np.array_equal(sym, rs), "%s: symbols differ at %s" % (name, np.nonzero(sym != rs)[0][:5])
    assert np.array_equal(freq, rf), name
    assert list(smap) == list(rm), name


def test_small_cases(engine, ref):
    for name, data in small_cases():
        _, bwt, _ = ref.bwt_encode(data, ref.SPEC_FAST)
        _check(engine, ref, bwt, name)
        _check(engine, ref, data, name + "/raw")      # any byte string is a valid input of this stage


def test_chunk_boundaries(engine, ref):
    # zero runs and first occurrences straddling the 1024-byte chunk edges
    rng = np.random.default_rng(3)
    for n in (1023, 1024, 1025, 2047, 2048, 2049, 4096, 5000):
        _check(engine, ref, bytes(n), "zeros%d" % n)
        _check(engine, ref, b"\x07" * n, "sevens%d" % n)
        a = np.zeros(n, dtype=np.uint8)
        a[rng.integers(0, n, 5)] = rng.integers(1, 255, 5)
        _check(engine, ref, a.tobytes(), "sparse%d" % n)
        a = np.repeat(rng.integers(0, 256, n // 100 + 1, dtype=np.uint8), 100)[:n]
        _check(engine, ref, a.tobytes(), "runs100_%d" % n)


def test_all_256_symbols(engine, ref):
    rng = np.random.default_rng(4)
    a = rng.integers(0, 256, 100_000, dtype=np.uint8)
    _check(engine, ref, a.tobytes(), "rand256")       # eob = 257: freq[256] truncation case of SURVEY B.3
    a = np.concatenate([np.arange(256, dtype=np.uint8)[::-1], np.zeros(5000, dtype=np.uint8), np.arange(256, dtype=np.uint8)])
    _check(engine, ref, a.tobytes(), "desc_then_zero")


def test_full_size_block(engine, ref):
    blk = corpus.text(899_981, 41).tobytes()
    _, bwt, _ = ref.bwt_encode(blk, ref.SPEC_FAST)
    _check(engine, ref, bwt, "text900k")
    blk = corpus.repetitive(899_000, 42).tobytes()
    _, bwt, _ = ref.bwt_encode(blk, ref.SPEC_FAST)
    _check(engine, ref, bwt, "rep900k")
