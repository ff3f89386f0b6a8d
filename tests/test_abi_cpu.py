"""CPU-side checks of the drop-in boundary: the library builds, loads and exports every symbol include/bz2b200.h
declares; host-only entry points behave; there is no silent CPU fallback."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import bzip2_rust_b200 as bz

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "bz2b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bz2b200_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from bzip2_rust_b200 import build
    build.build()
    return bz.load_library()


def test_every_declared_symbol_is_exported(lib):
    declared = _declared_symbols()
    assert len(declared) >= 25
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert sorted(bz.EXPORTS) == declared          # the Python binding covers the whole header


def test_version_and_bound(lib):
    assert b"sm_100a" in lib.bz2b200_version()
    assert lib.bz2b200_compress_bound(0) >= 14
    assert lib.bz2b200_compress_bound(10**8) > 10**8


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    assert lib.bz2b200_create(-1, C.byref(h)) == bz.E_CUDA
    with pytest.raises(bz.Bz2B200Error):
        bz.Engine()


def test_merge_streams_is_the_bitwriter(lib, ref):
    """bz2b200_merge_streams (host code) == BitWriter::add_block over the oracle's packed blocks (bitwriter.rs:77-132)."""
    from bzip2_rust_b200 import corpus
    data = corpus.mix1m(2, 450_000).tobytes()
    level = 1
    want = ref.compress_stream(data, level, ref.SPEC_FAST)
    packed = []
    for crc, blk, last, cons in ref.rle1_blocks(data, level):
        pb, pad, _ = ref.compress_block(blk, crc, ref.SPEC_FAST)
        packed.append((pb, len(pb) * 8 - pad, [crc]))
    assert bz.merge_streams(level, packed) == want
    # grouping blocks into larger parts first (what a rank does) gives the same stream
    def cat(parts):
        acc, n, crcs = 0, 0, []
        for pb, nb, c in parts:
            acc = (acc << nb) | (int.from_bytes(pb, "big") >> (len(pb) * 8 - nb))
            n += nb
            crcs += c
        pad = (8 - n % 8) % 8
        return ((acc << pad).to_bytes((n + pad) // 8, "big"), n, crcs)
    k = len(packed) // 2
    assert bz.merge_streams(level, [cat(packed[:k]), cat(packed[k:])]) == want


def test_bad_arguments_are_errors_not_crashes(lib):
    assert lib.bz2b200_merge_streams(0, 0, None, None, None, None, None, 0, None) == bz.E_ARG
    out = np.zeros(8, dtype=np.uint8)
    ln = C.c_size_t()
    assert lib.bz2b200_merge_streams(9, 0, None, None, None, None, out.ctypes.data, 8, C.byref(ln)) == bz.E_CAP


def test_header_is_plain_c_and_every_symbol_links(lib, tmp_path):
    """The boundary is a C ABI: the header must compile as C99 (no C++-isms), and a C caller that references every
    declared entry point must link against the shared library (nothing is called except bz2b200_version)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    syms = _declared_symbols()
    # bz2b200_sink is a callback TYPE, not a function of the library
    syms = [s for s in syms if s != "bz2b200_sink"]
    src = tmp_path / "abi.c"
    src.write_text('#include "bz2b200.h"\n#include <stdio.h>\nint main(void) {\n    void *p[] = {%s};\n'
                   '    printf("%%s %%d\\n", bz2b200_version(), (int)(sizeof p / sizeof p[0]));\n    return 0;\n}\n'
                   % ", ".join("(void *)%s" % s for s in syms))
    exe = tmp_path / "abi"
    so = os.path.join(ROOT, "bzip2_rust_b200", "libbz2b200.so")
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                        so, "-Wl,-rpath," + os.path.dirname(so)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    out = subprocess.run([str(exe)], stdout=subprocess.PIPE, text=True).stdout
    assert "sm_100a" in out and out.split()[-1] == str(len(syms))


def test_rust_binding_declares_the_header(lib):
    """integration/rust/bz2b200-sys/src/lib.rs (unbuilt here: no rustc) must declare exactly the functions of the header."""
    text = open(os.path.join(ROOT, "integration", "rust", "bz2b200-sys", "src", "lib.rs")).read()
    rust = sorted(set(re.findall(r"pub fn (bz2b200_[a-z0-9_]+)\s*\(", text)))
    declared = [s for s in _declared_symbols() if s != "bz2b200_sink"]
    assert rust == declared
    assert "pub type bz2b200_sink" in text
