"""bz2b200_compress_stream_multi: `compress` (compress.rs:40-136) over several ranks of ONE process -- one host thread +
one context per rank, the block chain handed over through host memory, every rank's bit string copied to its final
position.  On a single-GPU box the ranks share the GPU (device ids repeat; nothing on the device waits for another
rank).  The merged stream must be the bytes one single-GPU call produces, for every world size, and the oracle's."""
import bz2
import os

import numpy as np
import pytest
import torch

import bzip2_rust_b200 as bz
from bzip2_rust_b200 import corpus

pytestmark = pytest.mark.gpu


def _devices(world):
    n = torch.cuda.device_count()
    return [r % n for r in range(world)]


@pytest.fixture(scope="module")
def data80():
    # text / repetitive / random thirds: long runs make some blocks span more input than the 4 MB look-ahead
    return np.concatenate([corpus.text(30_000_000, 41), corpus.repetitive(30_000_000, 42), corpus.random_bytes(8_000_000, 43),
                           corpus.text(12_000_000, 44)])


@pytest.mark.parametrize("world", [2, 3, 8])
def test_worlds_produce_the_single_gpu_stream(engine, data80, world):
    want = engine.compress(data80, 9)
    m = bz.MultiEngine(_devices(world))
    try:
        got = m.compress(data80, 9)
        assert got == want, "world %d differs from the single-GPU stream" % world
        st = m.stats()
        assert st["h2d_bytes"] >= data80.size and st["d2h_bytes"] >= len(want) - 32
        # a second call on the same context (buffers reused, hand-off state reset), other level
        assert m.compress(data80, 3) == engine.compress(data80, 3)
    finally:
        m.close()
    assert bz2.decompress(got) == data80.tobytes()


def test_several_windows_per_rank(engine, data80, monkeypatch):
    """A rank works through its windows one after the other while the next window's upload and the previous window's
    download are in flight (two buffers each): 3 and 5 windows per rank must give the same bytes, and so must two windows
    of unequal size (the first one is the smaller: BZ2B200_MULTI_FIRST_PCT)."""
    want = engine.compress(data80, 9)
    for per, world in ((3, 2), (5, 3), (2, 2), (2, 4)):
        monkeypatch.setenv("BZ2B200_MULTI_WINDOWS", str(per))
        m = bz.MultiEngine(_devices(world))
        try:
            assert m.compress(data80, 9) == want, (per, world)
        finally:
            m.close()


def test_multi_equals_oracle_on_two_text_segments(engine, ref):
    data = np.concatenate([corpus.text(20_000_000, 2), corpus.text(20_000_000, 3)])
    m = bz.MultiEngine(_devices(2))
    try:
        got = m.compress(data, 9)
    finally:
        m.close()
    want = ref.compress_stream(data.tobytes(), 9, ref.SPEC_FAST, threads=os.cpu_count() or 1)
    assert got == want


def test_small_inputs_and_errors(engine):
    m = bz.MultiEngine(_devices(4))
    try:
        for n in (0, 1, 5, 100_000, 3_000_000):                 # below the multi threshold: the single-GPU path, same bytes
            d = corpus.text(n, 7) if n else np.zeros(0, dtype=np.uint8)
            assert m.compress(d, 9) == engine.compress(d, 9)
        d = corpus.text(40_000_000, 8)
        out = np.empty(1000, dtype=np.uint8)                    # far too small: an error code, not a crash
        with pytest.raises(bz.Bz2B200Error) as e:
            m.compress_into(d.ctypes.data, d.size, 9, out.ctypes.data, out.size)
        assert e.value.rc == bz.E_CAP
        assert m.compress(d, 9) == engine.compress(d, 9)        # and the context still works afterwards
    finally:
        m.close()


def test_all_visible_gpus(engine):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one GPU visible: the multi-rank path on distinct GPUs runs under gpurun --gpus N / bench.py --gpus N")
    data = corpus.text(n * 30_000_000, 2)
    m = bz.MultiEngine(list(range(n)))
    try:
        assert m.compress(data, 9) == engine.compress(data, 9)
    finally:
        m.close()
