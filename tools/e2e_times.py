"""End-to-end (pinned host buffers) timing of bz2b200_compress_stream for a few window counts (dev loop helper)."""
import ctypes as C, json, os, sys, time, zlib
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bzip2_rust_b200 as bz
from bzip2_rust_b200 import corpus
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 100
data = corpus.text(mb * 1_000_000, 2)
L = bz.load_library()
eng = bz.Engine(0)
h_in = torch.from_numpy(data).pin_memory()
cap = int(L.bz2b200_compress_bound(data.size))
h_out = torch.empty(cap, dtype=torch.uint8).pin_memory()
out_len = C.c_size_t()
def step():
    rc = L.bz2b200_compress_stream(eng._h, h_in.data_ptr(), data.size, 9, h_out.data_ptr(), cap, C.byref(out_len))
    assert rc == 0, rc
for _ in range(3): step()
ts = []
for _ in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter(); step(); ts.append((time.perf_counter() - t0) * 1e3)
print(json.dumps({"env": os.environ.get("BZ2B200_E2E_WINDOWS"), "ms": [round(t, 2) for t in ts], "best_MBps": round(data.size / 1e3 / min(ts), 1),
                  "adler": zlib.adler32(h_out[:out_len.value].numpy().tobytes()), "len": out_len.value}))
