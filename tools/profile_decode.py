"""One decompression of the text100m stream (pinned host buffers).  Run under ncu, e.g.
   ncu --set full --clock-control none -k regex:'k_dec_jumps|k_dec_bounds|k_dec_expand' -c 3 -o X python tools/profile_decode.py
(the jump-table producer is launched before the walkers, so a tool that serialises kernels still terminates)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bzip2_rust_b200 as bz
from bzip2_rust_b200 import corpus

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 100
data = corpus.text(mb * 1_000_000, 2)
eng = bz.Engine(0)
L = bz.load_library()
stream = np.frombuffer(eng.compress(data, 9), dtype=np.uint8)
h_in = torch.from_numpy(stream.copy()).pin_memory()
h_out = torch.empty(data.size + 1024, dtype=torch.uint8).pin_memory()
n_out = C.c_size_t()
rc = L.bz2b200_decompress_stream(eng._h, h_in.data_ptr(), stream.size, h_out.data_ptr(), h_out.numel(), C.byref(n_out))
print("rc", rc, "bytes", n_out.value, "ok", n_out.value == data.size and bool((h_out[:data.size].numpy() == data).all()))
