"""Runs the BASELINE.json configs on one B200 and prints one JSON object per config.

    python tools/run_configs.py [text100m] [markov100m] [rep256m] [sweep] [big] [decode]

Throughput = input bytes / wall time of bz2b200_compress_stream_dev (input + output resident in HBM), best of 3
after one warm-up.  Parity: byte identity against the CPU oracle on a leading sample (the oracle needs seconds
per block) and a full libbz2 round trip.
"""
import bz2
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bzip2_rust_b200 as bz                 # noqa: E402
from bzip2_rust_b200 import corpus           # noqa: E402
from oracle import pyref                     # noqa: E402


def timed_compress(eng, d_in, n, level, d_out, cap, reps=3):
    eng.compress_dev(d_in.data_ptr(), n, level, d_out.data_ptr(), cap)
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ln = eng.compress_dev(d_in.data_ptr(), n, level, d_out.data_ptr(), cap)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best, ln


def run(name, data, level, eng, L, oracle_sample=4_000_000, verify=True, ref_paths=False):
    n = data.size
    d_in = torch.from_numpy(data).cuda()
    cap = int(L.bz2b200_compress_bound(n))
    d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    eng.set_timing(1)
    dt, ln = timed_compress(eng, d_in, n, level, d_out, cap)
    st = eng.timing()
    bw = eng.bwt_stats()
    res = {"config": name, "level": level, "input_bytes": n, "compressed_bytes": ln, "MBps": n / 1e6 / dt,
           "ms": dt * 1e3, "stage_ms": st, "bwt_rounds_last_batch": int(bw["rounds"]),
           "bwt_list_sum_per_n_last_batch": round(bw["list_sum"] / max(1, min(n, 272 << 20)), 3),
           "bwt_big_path_share": round(bw["big_sum"] / max(1, bw["list_sum"]), 4)}
    if ref_paths:
        # blocks native / sais-valid / sais-divergent (the oracle's bug-for-bug EXACT mode on the flagged blocks)
        import bench
        res["ref_path"] = bench.ref_path_accounting(eng, data[:min(n, 64_000_000)], level, limit=24)
    if verify:
        stream = d_out[:ln].cpu().numpy().tobytes()
        raw = data.tobytes()
        res["libbz2_roundtrip"] = bz2.decompress(stream) == raw
        # byte identity against the oracle on a leading sample compressed as its own stream
        k = min(n, oracle_sample)
        ours = eng.compress(data[:k], level)
        want, stats = pyref.compress_stream(raw[:k], level, pyref.EXACT, threads=os.cpu_count() or 1, want_stats=True)
        spec = pyref.compress_stream(raw[:k], level, pyref.SPEC_FAST, threads=os.cpu_count() or 1)
        res["oracle_sample_bytes"] = k
        res["identical_to_ref_spec"] = ours == spec
        res["identical_to_ref_exact"] = ours == want
        res["ref_exact_paths"] = {"native": stats["n_native"], "sais": stats["n_sais"],
                                  "sais_divergent": stats["n_sais_divergent"]}
    del d_in, d_out
    print(json.dumps(res), flush=True)
    return res


def main():
    which = set(sys.argv[1:]) or {"text100m", "markov100m", "rep256m", "sweep", "decode"}
    eng = bz.Engine(0)
    L = bz.load_library()
    wk = corpus.default_workers()
    if "text100m" in which:
        run("text100m", corpus.text(100_000_000, 2), 9, eng, L, ref_paths=True)
    if "markov100m" in which:
        run("markov100m", corpus.markov(100_000_000, 6, workers=wk), 9, eng, L, ref_paths=True)
    if "rep256m" in which:
        run("rep256m", corpus.rep_segments(256_000_000, 3, workers=wk), 9, eng, L, ref_paths=True)
    if "sweep" in which:
        data = corpus.mixed(1_000_000_000 if "big" in which else 256_000_000, 5, workers=wk)
        for level in range(1, 10):
            run("sweep_mixed", data, level, eng, L, verify=(level in (1, 9)), ref_paths=(level == 9))
    if "decode" in which:
        import ctypes as C
        data = corpus.text(100_000_000, 2)
        stream = np.frombuffer(eng.compress(data, 9), dtype=np.uint8)
        # pinned host buffers, straight through the C ABI (the ctypes convenience wrapper allocates and copies)
        h_in = torch.from_numpy(stream.copy()).pin_memory()
        h_out = torch.empty(data.size + 1024, dtype=torch.uint8).pin_memory()
        n_out = C.c_size_t()

        def dec():
            rc = L.bz2b200_decompress_stream(eng._h, h_in.data_ptr(), stream.size, h_out.data_ptr(), h_out.numel(), C.byref(n_out))
            assert rc == 0, rc
        dec()
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            dec()
            best = min(best, time.perf_counter() - t0)
        eng.set_timing(2)
        eng.reset_kernel_stats()
        dec()
        ks = {k: round(v[0], 3) for k, v in sorted(eng.kernel_stats().items(), key=lambda kv: -kv[1][0])}
        ok = n_out.value == data.size and bool((h_out[:data.size].numpy() == data).all())
        print(json.dumps({"config": "decode_text100m", "MBps_output": data.size / 1e6 / best, "ms": best * 1e3,
                          "ok": ok, "kernel_ms": ks, "kernel_ms_sum": round(sum(ks.values()), 3),
                          "note": "pinned host buffers (H2D of .bz2 and D2H of the text included)"}),
              flush=True)


if __name__ == "__main__":
    main()
