"""One warm-up + one profiled whole-stream compression of text100m (HBM resident).  Run under ncu:
   ncu --metrics gpu__time_duration.sum --clock-control none -s <skip> -c <n> --csv --log-file X python tools/profile_step.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bzip2_rust_b200 as bz
from bzip2_rust_b200 import corpus

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 100
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 1
data = corpus.text(mb * 1_000_000, 2)
d_in = torch.from_numpy(data).cuda()
eng = bz.Engine(0)
cap = int(bz.load_library().bz2b200_compress_bound(data.size))
d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
for i in range(warm + 1):
    l0 = eng.launches
    n = eng.compress_dev(d_in.data_ptr(), data.size, 9, d_out.data_ptr(), cap)
    print("step", i, "bytes", n, "launches", eng.launches - l0)
