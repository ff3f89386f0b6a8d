"""SURVEY 8e on one GPU: the ranks of a sharded compression are emulated one after the other (the B200 profiling
guide forbids kernels that wait on one another on a single GPU, and none do here: the hand-off is a host number).
The merged stream must be the bytes a single whole-stream call produces."""
import bz2

import numpy as np
import pytest
import torch

import bzip2_rust_b200 as bz
from bzip2_rust_b200 import corpus

pytestmark = pytest.mark.gpu


def _sharded(engine, data, level, world, window_slack):
    L = bz.load_library()
    total = data.size
    d_in = torch.from_numpy(data).cuda()
    per = total // world
    cap = int(L.bz2b200_compress_bound(total))
    parts = []
    start = 0
    for r in range(world):
        lo, hi = r * per, (total if r == world - 1 else (r + 1) * per)
        # rank r only has [lo, hi + slack) resident; grow the window when the plan asks for it
        slack = window_slack
        while True:
            win_len = min(total - lo, hi - lo + slack)
            d_win = d_in[lo:lo + win_len].clone()
            torch.cuda.synchronize()        # the context works on its own non-blocking stream: torch's copy must have landed
            if r % 2 == 0:
                engine.shard_scan(d_win.data_ptr(), lo, win_len, total, level)     # optional early phase
            try:
                nxt, nb = engine.shard_plan(d_win.data_ptr(), lo, win_len, total, level, start, hi)
                break
            except bz.Bz2B200Error as e:
                assert e.rc == bz.E_CAP
                assert lo + win_len < total
                slack *= 4
        d_out = torch.zeros(cap, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        bits, crcs = engine.shard_compress(nb, d_out.data_ptr(), cap)
        parts.append((d_out, bits, crcs))
        assert nxt >= min(hi, total)
        start = nxt
    assert start == total
    # ordered merge with pre-shifted parts (what bench.py does across ranks)
    pos = 32
    final = torch.zeros(cap + 64, dtype=torch.uint8, device="cuda")
    combined = 0
    host_parts = []
    for d_out, bits, crcs in parts:
        phase = pos % 8
        d_shift = torch.zeros(cap + 64, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        engine.shift_bits(d_out.data_ptr(), bits, phase, d_shift.data_ptr())
        nby = (bits + phase + 7) // 8
        final[pos // 8:pos // 8 + nby].bitwise_or_(d_shift[:nby])
        host_parts.append((d_out[:(bits + 7) // 8].cpu().numpy().tobytes(), bits, [int(c) for c in crcs]))
        for c in crcs:
            combined = (((combined << 1) | (combined >> 31)) & 0xFFFFFFFF) ^ int(c)
        pos += bits
    total_bits = pos + 80
    nfinal = (total_bits + 7) // 8
    buf = final[:nfinal].cpu().numpy()
    buf[0:4] = np.frombuffer(b"BZh" + bytes([48 + level]), dtype=np.uint8)
    foot = ((0x177245385090 << 32) | combined) << ((8 - total_bits % 8) % 8)
    buf[nfinal - 11:nfinal] |= np.frombuffer(foot.to_bytes(11, "big"), dtype=np.uint8)
    return buf.tobytes(), bz.merge_streams(level, host_parts)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_stream_equals_single_stream(engine, world):
    data = corpus.text(6_000_000, 21)
    whole = engine.compress(data, 3)
    dev_merge, host_merge = _sharded(engine, data, 3, world, 512 << 10)
    assert host_merge == whole
    assert dev_merge == whole
    assert bz2.decompress(dev_merge) == data.tobytes()


def test_sharded_run_heavy_input_with_tiny_windows(engine):
    # blocks that span far more input than a shard: windows must be grown, some ranks get zero blocks
    data = np.concatenate([corpus.repetitive(1_500_000, 5), np.zeros(9_000_000, dtype=np.uint8),
                           corpus.text(800_000, 6), np.full(3_000_000, 7, dtype=np.uint8), corpus.text(300_000, 7)])
    whole = engine.compress(data, 1)
    dev_merge, host_merge = _sharded(engine, data, 1, 6, 64 << 10)
    assert host_merge == whole
    assert dev_merge == whole
