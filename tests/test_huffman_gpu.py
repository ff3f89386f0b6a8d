"""T1: CUDA Huffman stage (bz2b200_huffman) vs the oracle's huf_encode restatement (huffman.rs:79-468)."""
import numpy as np
import pytest

from bzip2_rust_b200 import corpus
from inputs import small_cases

pytestmark = pytest.mark.gpu


def _check(engine, ref, data, name=""):
    _, bwt, _ = ref.bwt_encode(data, ref.SPEC_FAST)
    sym, freq, smap = ref.rle2_mtf_encode(bwt)
    rb, rbits, rinfo = ref.huf_encode(sym, freq, smap)
    gb, gbits, ginfo = engine.huf_encode(sym, freq, smap)
    assert ginfo["table_count"] == rinfo["table_count"], name
    assert ginfo["selectors"] == rinfo["selectors"], "%s: selectors differ" % name
    T = rinfo["table_count"]
    nsym = int(sym[-1]) + 1
    assert np.array_equal(ginfo["lengths"][:T, :nsym], rinfo["lengths"][:T, :nsym]), "%s: code lengths differ" % name
    assert gbits == rbits, "%s: bit length %d vs %d" % (name, gbits, rbits)
    assert gb == rb, "%s: packed bits differ" % name
    return rinfo


def test_small_cases(engine, ref):
    for name, data in small_cases():
        _check(engine, ref, data, name)


def test_table_count_thresholds(engine, ref):
    # m just below / above 200, 600, 1200, 2400 (huffman.rs:87-93)
    rng = np.random.default_rng(9)
    for n in (190, 198, 199, 200, 210, 590, 598, 599, 610, 1190, 1199, 1200, 1210, 2390, 2399, 2400, 2410):
        data = bytes(rng.integers(0, 256, n, dtype=np.uint8))
        _check(engine, ref, data, "rand%d" % n)


def test_group_tail_sizes(engine, ref):
    rng = np.random.default_rng(10)
    for n in (49, 50, 51, 99, 100, 101, 12799, 12800, 12801, 25600):
        data = bytes(rng.integers(0, 7, n, dtype=np.uint8))
        _check(engine, ref, data, "seven%d" % n)


def _fib_symbols(nsym=30, seed=11):
    """MTF/RLE2-stage output with Fibonacci symbol frequencies (drives the tree past depth 17)."""
    parts = []
    a, b = 1, 1
    for s in range(2, nsym):
        parts.append(np.full(a, s, dtype=np.uint16))
        a, b = b, a + b
    arr = np.concatenate(parts)
    np.random.default_rng(seed).shuffle(arr)
    sym = np.concatenate([arr, np.array([nsym], dtype=np.uint16)])
    freq = np.zeros(256, dtype=np.uint32)
    for v, c in zip(*np.unique(arr, return_counts=True)):
        freq[v - 1] += c                      # rle2_mtf.rs:104 indexing
    maps = [0] * 17
    for i in range(nsym - 1):
        maps[0] |= 0x8000 >> (i >> 4)
        maps[1 + (i >> 4)] |= 0x8000 >> (i & 15)
    return sym, freq, np.array([m for m in maps if m], dtype=np.uint16)


def test_depth_limit_retry_and_tie(engine, ref):
    # trees deeper than 17 -> weight halving (huffman_code_from_weights.rs:73-81); this input also
    # produces one (weight, syms) tie between distinct nodes
    sym, freq, smap = _fib_symbols()
    rb, rbits, rinfo = ref.huf_encode(sym, freq, smap)
    assert rinfo["retries"] > 0
    gb, gbits, ginfo = engine.huf_encode(sym, freq, smap)
    T = rinfo["table_count"]
    assert ginfo["selectors"] == rinfo["selectors"]
    assert np.array_equal(ginfo["lengths"][:T, :31], rinfo["lengths"][:T, :31])
    assert (gbits, gb) == (rbits, rb)


def test_full_size_blocks(engine, ref):
    total_ties = 0
    for name, data in (("text", corpus.text(899_981, 51)), ("rand", corpus.random_bytes(899_981, 52)),
                       ("rep", corpus.repetitive(899_981, 53)), ("walk", corpus.random_walk(500_000, 54))):
        info = _check(engine, ref, data.tobytes(), name)
        total_ties += info["tie_events"]
    print("huffman (weight,syms) tie events seen:", total_ties)
