"""BASELINE config 4: the 8 GB corpus at -9, STRONG scaling over 1/2/4/8 GPUs of one box, through the library's
multi-GPU entry point (bz2b200_compress_stream_multi: host buffers in, host buffers out, copies inside the timed call).

    python tools/run_corpus8g.py [--gb 8] [--gpus 1,2,4,8] [--reps 2]

One JSON object per GPU count: MB/s, speed-up and efficiency against 1 GPU, and `identical` = the stream has the same
SHA-256 as the 1-GPU stream (which tests/test_fullsize_gpu.py compares with the CPU oracle).
"""
import argparse
import hashlib
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bzip2_rust_b200 as bz                 # noqa: E402
from bzip2_rust_b200 import corpus           # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gb", type=float, default=8.0)
    ap.add_argument("--gpus", default="1,2,4,8")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--level", type=int, default=9)
    a = ap.parse_args()
    n = int(a.gb * 1e9)
    t0 = time.perf_counter()
    data = corpus.corpus(n, 4, workers=corpus.default_workers())
    gen_s = time.perf_counter() - t0
    L = bz.load_library()
    h_in = torch.from_numpy(data).pin_memory()
    del data
    cap = int(L.bz2b200_compress_bound(n))
    h_out = torch.empty(cap, dtype=torch.uint8).pin_memory()
    avail = torch.cuda.device_count()
    base = None
    sha1 = None
    for g in [int(x) for x in a.gpus.split(",")]:
        if g > avail:
            continue
        m = bz.MultiEngine(list(range(g)))
        try:
            ln = m.compress_into(h_in.data_ptr(), n, a.level, h_out.data_ptr(), cap)       # warm-up: allocations, page faults
            best = 1e30
            for _ in range(a.reps):
                t0 = time.perf_counter()
                ln = m.compress_into(h_in.data_ptr(), n, a.level, h_out.data_ptr(), cap)
                best = min(best, time.perf_counter() - t0)
            st = m.stats()
        finally:
            m.close()
        sha = hashlib.sha256(h_out[:ln].numpy().tobytes()).hexdigest()
        if sha1 is None:
            sha1 = sha
        rate = n / 1e6 / best
        if base is None:
            base = rate
        print(json.dumps({"config": "corpus8g", "level": a.level, "input_bytes": n, "gpus": g, "scaling": "strong",
                          "MBps_e2e": round(rate, 1), "ms": round(best * 1e3, 2), "speedup_vs_1": round(rate / base, 3),
                          "efficiency": round(rate / base / g, 3), "compressed_bytes": int(ln), "sha256": sha[:16],
                          "identical_to_1gpu": sha == sha1, "h2d_bytes": st["h2d_bytes"], "d2h_bytes": st["d2h_bytes"],
                          "api": "bz2b200_compress_stream_multi", "corpus_generation_s": round(gen_s, 1),
                          "host_cores": os.cpu_count()}), flush=True)


if __name__ == "__main__":
    main()
