"""Per-kernel CUDA-event times of one HBM-resident whole-stream compression (dev loop helper).

    python tools/kernel_times.py [MB=100] [corpus=text|mixed|rep] [level=9] [check=0|1]

Prints one JSON line: stage times, per-kernel ms / launches, output size and a checksum of the output so that two
kernel variants (selected through BZ2B200_* environment knobs) can be compared byte for byte across runs."""
import json
import os
import sys
import time
import zlib

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bzip2_rust_b200 as bz
from bzip2_rust_b200 import corpus

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 100
kind = sys.argv[2] if len(sys.argv) > 2 else "text"
level = int(sys.argv[3]) if len(sys.argv) > 3 else 9
check = int(sys.argv[4]) if len(sys.argv) > 4 else 0
n = mb * 1_000_000
data = {"text": lambda: corpus.text(n, 2), "mixed": lambda: corpus.mixed(n, 5),
        "rep": lambda: corpus.repetitive(n, 3)}[kind]()
d_in = torch.from_numpy(data).cuda()
eng = bz.Engine(0)
cap = int(bz.load_library().bz2b200_compress_bound(data.size))
d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
for i in range(2):
    eng.compress_dev(d_in.data_ptr(), data.size, level, d_out.data_ptr(), cap)
torch.cuda.synchronize()
t0 = time.perf_counter()
z = eng.compress_dev(d_in.data_ptr(), data.size, level, d_out.data_ptr(), cap)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) * 1e3
eng.set_timing(2)
eng.reset_kernel_stats()
z = eng.compress_dev(d_in.data_ptr(), data.size, level, d_out.data_ptr(), cap)
torch.cuda.synchronize()
ks = eng.kernel_stats()
stage = eng.timing()
out = d_out[:z].cpu().numpy()
rec = {"corpus": kind, "mb": mb, "level": level, "wall_ms_untimed_mode": round(wall, 3), "z": int(z),
       "adler": zlib.adler32(out.tobytes()),
       "stage_ms": {k: round(v, 3) for k, v in stage.items()} if isinstance(stage, dict) else stage,
       "bwt_stats": eng.bwt_stats(),
       "env": {k: v for k, v in os.environ.items() if k.startswith("BZ2B200_")},
       "kernels": sorted(([k, round(v[0], 3), v[1]] for k, v in ks.items()), key=lambda r: -r[1])}
if check:
    import bz2
    rec["libbz2_roundtrip"] = bz2.decompress(out.tobytes()) == data.tobytes()
print(json.dumps(rec))
