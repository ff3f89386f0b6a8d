"""T1: CUDA BWT (bz2b200_bwt_encode*) vs the oracle's bwt_encode restatement (bwt_sort.rs:27-58). Bit exact."""
import numpy as np
import pytest

from bzip2_rust_b200 import corpus
from inputs import small_cases

pytestmark = pytest.mark.gpu


def _check(engine, ref, blocks, mode=None):
    mode = ref.SPEC_FAST if mode is None else mode
    got = engine.bwt_encode_batch(blocks)
    for i, (blk, (key, bwt)) in enumerate(zip(blocks, got)):
        k, b, _ = ref.bwt_encode(blk, mode)
        assert bwt == b, "bwt bytes differ for block %d (n=%d)" % (i, len(blk))
        assert key == k, "origin pointer differs for block %d (n=%d): %d vs %d" % (i, len(blk), key, k)


def test_small_cases_one_by_one(engine, ref):
    for name, data in small_cases():
        key, bwt = engine.bwt_encode(data)
        k, b, _ = ref.bwt_encode(data, ref.SPEC)
        assert (key, bwt) == (k, b), name


def test_small_cases_as_one_ragged_batch(engine, ref):
    _check(engine, ref, [d for _, d in small_cases()])


def test_every_length_1_to_70(engine, ref):
    rng = np.random.default_rng(5)
    blocks = [bytes(rng.integers(97, 100, n, dtype=np.uint8)) for n in range(1, 71)]
    _check(engine, ref, blocks, ref.SPEC)


def test_full_size_blocks(engine, ref):
    n = 899_981
    blocks = [corpus.text(n, 21).tobytes(), corpus.random_bytes(n, 22).tobytes(),
              corpus.random_walk(n + 5, 23).tobytes()]
    _check(engine, ref, blocks)
    st = engine.bwt_stats()
    assert st["rounds"] >= 1


def test_pathological_blocks(engine, ref):
    n = 300_000
    rle1_runs = (b"aaaa\xfb" * (n // 5 + 1))[:n]            # post-RLE1 image of one long run (period 5)
    blocks = [rle1_runs, rle1_runs[:n - 3], b"ab" * (n // 2), (b"abc" * (n // 3 + 1))[:n - 1],
              corpus.repetitive(n, 31).tobytes(), bytes(n)]
    _check(engine, ref, blocks)


def test_real_rle1_blocks_of_repetitive_input(engine, ref):
    data = corpus.repetitive(2_500_000, 3).tobytes()
    blocks = [b for _, b, _, _ in ref.rle1_blocks(data, 9)]
    _check(engine, ref, blocks)


def test_inverse_property_at_full_size(engine, ref):
    blk = corpus.mix1m(1)[:899_986].tobytes()
    key, bwt = engine.bwt_encode(blk)
    assert ref.bwt_decode(key, bwt) == blk


def test_reference_path_selector_count(engine, ref):
    """k_ref_path counts the blocks the reference would route to its SA-IS fallback (bwt_sort.rs:29): longer than 5000
    bytes and at most 1499 LMS positions (sentinel included) in the first 5000."""
    blocks = [corpus.text(300_000, 5).tobytes(), corpus.random_walk(300_000, 6).tobytes(),
              corpus.random_bytes(200_000, 7).tobytes(), bytes(range(256)) * 40, b"ab" * 4000,
              corpus.text(4000, 8).tobytes(), corpus.repetitive(100_000, 9).tobytes()]
    want = sum(1 for b in blocks if len(b) > 5000 and ref.lms_count(b) <= 1499)
    assert 0 < want < len(blocks)
    engine.bwt_encode_batch(blocks)
    assert engine.bwt_stats()["ref_sais_blocks"] == want
