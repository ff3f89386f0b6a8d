"""N > 1 host logic on CPU (gloo, world_size 2; no GPU): each rank produces the packed blocks of its share, the pieces
travel over torch.distributed, and rank 0 runs the product's ordered merge (bz2b200_merge_streams: bit-granular
concatenation + combined CRC + header/footer, bitwriter.rs:77-132).  The block compressor here is the ORACLE, used
as the stand-in data source; the thing under test is the sharding arithmetic, the hand-off and the host merge."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import bzip2_rust_b200 as bz
        from bzip2_rust_b200 import corpus
        from oracle import pyref
        level = 1
        data = corpus.mix1m(7, 700_000).tobytes()
        blocks = list(pyref.rle1_blocks(data, level))            # (crc, rle1 bytes, last, consumed)
        nb = len(blocks)
        # the chain of block starts is handed from rank to rank as one number (here: the block index)
        if rank == 0:
            first = 0
        else:
            t = torch.zeros(1, dtype=torch.int64)
            dist.recv(t, rank - 1)
            first = int(t.item())
        last = nb if rank == world - 1 else (rank + 1) * nb // world
        if rank < world - 1:
            dist.send(torch.tensor([last], dtype=torch.int64), rank + 1)
        # this rank's bit string: its blocks concatenated at bit granularity
        acc, nbits, crcs = 0, 0, []
        for crc, blk, _, _ in blocks[first:last]:
            packed, pad, _ = pyref.compress_block(blk, crc, pyref.SPEC_FAST)
            bits = len(packed) * 8 - pad
            acc = (acc << bits) | (int.from_bytes(packed, "big") >> pad)
            nbits += bits
            crcs.append(crc)
        padb = (8 - nbits % 8) % 8
        part = (acc << padb).to_bytes((nbits + padb) // 8, "big")
        meta = torch.tensor([nbits, len(crcs), len(part)], dtype=torch.int64)
        metas = [torch.zeros(3, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(metas, meta)
        if rank == 0:
            parts = [(part, nbits, crcs)]
            for r in range(1, world):
                nb_r, nc_r, len_r = [int(x) for x in metas[r]]
                buf = torch.empty(len_r, dtype=torch.uint8)
                cr = torch.empty(nc_r, dtype=torch.int64)
                dist.recv(buf, r)
                dist.recv(cr, r)
                parts.append((buf.numpy().tobytes(), nb_r, [int(x) for x in cr]))
            merged = bz.merge_streams(level, parts)
            want = pyref.compress_stream(data, level, pyref.SPEC_FAST)
            import bz2
            q.put((merged == want, bz2.decompress(merged) == data, nb))
        else:
            dist.send(torch.frombuffer(bytearray(part), dtype=torch.uint8), 0)
            dist.send(torch.tensor(crcs, dtype=torch.int64), 0)
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_shard_handoff_and_merge_gloo():
    from bzip2_rust_b200 import build
    build.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 300)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    same, roundtrip, nb = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert nb >= 4
    assert same and roundtrip
