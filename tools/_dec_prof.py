import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bzip2_rust_b200 as bz
from bzip2_rust_b200 import corpus
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 100
data = corpus.text(mb * 1_000_000, 2)
eng = bz.Engine(0); L = bz.load_library()
stream = np.frombuffer(eng.compress(data, 9), dtype=np.uint8)
h_in = torch.from_numpy(stream.copy()).pin_memory()
h_out = torch.empty(data.size + 1024, dtype=torch.uint8).pin_memory()
n_out = C.c_size_t()
for _ in range(2):
    rc = L.bz2b200_decompress_stream(eng._h, h_in.data_ptr(), stream.size, h_out.data_ptr(), h_out.numel(), C.byref(n_out))
    assert rc == 0
print("ok", n_out.value)
