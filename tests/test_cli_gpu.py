"""tools/bz2b200_cli.py: the reference's flag surface (src/tools/cli.rs:124-288) over libbz2b200.  -1..-9 / --fast / --best
pick the block size, -z / -d / -t the mode, -c writes to stdout, -f overwrites, -k is accepted (input files are ALWAYS
kept, as in the reference: cli.rs:314), several files and stdin/stdout work (advertised by the reference's help text,
cli.rs:331-333, not implemented there)."""
import bz2
import os
import subprocess
import sys

import pytest

from bzip2_rust_b200 import corpus

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = [sys.executable, os.path.join(ROOT, "tools", "bz2b200_cli.py")]


def _run(args, cwd, stdin=None):
    return subprocess.run(CLI + args, cwd=cwd, input=stdin, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)


def test_cli_round_trips(tmp_path, ref):
    d = str(tmp_path)
    a = corpus.text(1_500_000, 51).tobytes()
    b = corpus.mixed(1_200_000, 52).tobytes()
    open(os.path.join(d, "a.txt"), "wb").write(a)
    open(os.path.join(d, "b.bin"), "wb").write(b)
    # default: compress at level 9, FILE.bz2 next to the input (compress.rs:59-60), inputs kept
    r = _run(["a.txt", "b.bin"], d)
    assert r.returncode == 0, r.stderr
    assert os.path.exists(os.path.join(d, "a.txt")) and os.path.exists(os.path.join(d, "b.bin"))
    za = open(os.path.join(d, "a.txt.bz2"), "rb").read()
    assert za[:4] == b"BZh9" and bz2.decompress(za) == a
    assert za == ref.compress_stream(a, 9, ref.SPEC_FAST, threads=4)           # the file the reference would write
    assert bz2.decompress(open(os.path.join(d, "b.bin.bz2"), "rb").read()) == b
    # an existing output is not overwritten without -f
    r = _run(["-1", "a.txt"], d)
    assert r.returncode == 1 and open(os.path.join(d, "a.txt.bz2"), "rb").read() == za
    r = _run(["-1", "-f", "-k", "a.txt"], d)
    assert r.returncode == 0
    z1 = open(os.path.join(d, "a.txt.bz2"), "rb").read()
    assert z1[:4] == b"BZh1" and bz2.decompress(z1) == a
    # every level through stdout
    for lv in range(1, 10):
        r = _run(["-%d" % lv, "-c", "a.txt"], d)
        assert r.returncode == 0 and r.stdout[:4] == b"BZh%d" % lv and bz2.decompress(r.stdout) == a
    assert _run(["--fast", "-c", "a.txt"], d).stdout[:4] == b"BZh1"
    assert _run(["--best", "-z", "-c", "a.txt"], d).stdout[:4] == b"BZh9"
    # -t: integrity test, no output file; a damaged file fails
    os.remove(os.path.join(d, "a.txt"))
    r = _run(["-t", "a.txt.bz2", "b.bin.bz2"], d)
    assert r.returncode == 0 and not os.path.exists(os.path.join(d, "a.txt"))
    bad = bytearray(z1)
    bad[len(bad) // 2] ^= 1
    open(os.path.join(d, "bad.bz2"), "wb").write(bytes(bad))
    assert _run(["-t", "-q", "bad.bz2"], d).returncode == 2
    # -d: strips .bz2, keeps the archive
    r = _run(["-d", "a.txt.bz2"], d)
    assert r.returncode == 0 and open(os.path.join(d, "a.txt"), "rb").read() == a and os.path.exists(os.path.join(d, "a.txt.bz2"))
    # stdin -> stdout, both directions; concatenated archives decode to the concatenation
    r = _run(["-5"], d, stdin=b)
    assert r.returncode == 0 and bz2.decompress(r.stdout) == b
    r2 = _run(["-d"], d, stdin=r.stdout + z1)
    assert r2.returncode == 0 and r2.stdout == b + a
    # version / license
    assert b"bz2b200" in _run(["-V"], d).stdout and b"bz2b200" in _run(["--license"], d).stdout
