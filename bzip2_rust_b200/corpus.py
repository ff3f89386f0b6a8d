"""Seeded synthetic inputs for the BASELINE.json configs.  numpy only; deterministic.

    mix1m(seed=1)        1 000 000 B: alternating 64 KiB spans of text / random / random-walk binary
    text(n, seed=2)      text-like: Zipf-weighted words from a 4096-word vocabulary, lines of 40-100
    repetitive(n, seed=3) long runs, short-period repeats, `aaaa\\xfb`-style post-RLE1 pathologies
    mixed(n, seed=5)     50/50 random bytes and text in 1 MiB spans (block-size sweep)
    corpus(n, seed=4)    text model with a different seed per 32 MiB segment plus ~10% binary segments (config 4)
    markov(n, seed=6)    order-2 text: word TRANSITIONS are drawn from a sparse table, so contexts repeat far deeper
                         than in the i.i.d. word model (more prefix-doubling rounds, larger groups)
The big corpora are made of independently seeded segments; `workers` > 1 generates them in parallel processes (same
bytes as the serial call).
"""
import os

import numpy as np

_LETTERS = np.frombuffer(b"etaoinshrdlcumwfgypbvkjxqz", dtype=np.uint8)
_LETTER_P = np.array([12.7, 9.1, 8.2, 7.5, 7.0, 6.7, 6.3, 6.1, 6.0, 4.3, 4.0, 2.8, 2.8, 2.4, 2.4, 2.2, 2.0, 2.0,
                      1.9, 1.5, 1.0, 0.8, 0.15, 0.15, 0.10, 0.07])
_LETTER_P = _LETTER_P / _LETTER_P.sum()


def _vocab(rng, nwords=4096, maxlen=12):
    lens = np.clip(rng.poisson(4.2, nwords) + 1, 1, maxlen)
    mat = np.full((nwords, maxlen + 1), ord(" "), dtype=np.uint8)
    for i in range(nwords):
        mat[i, :lens[i]] = rng.choice(_LETTERS, size=lens[i], p=_LETTER_P)
    return mat, lens


def text(n, seed=2):
    """Text-like bytes: words, spaces, newlines every 40-100 chars, ~2% digits/punctuation."""
    rng = np.random.default_rng(seed)
    mat, lens = _vocab(rng)
    nw = mat.shape[0]
    p = 1.0 / np.arange(1, nw + 1) ** 1.05
    p /= p.sum()
    out = np.empty(n + 64, dtype=np.uint8)
    pos = 0
    while pos < n:
        k = min(1 << 20, (n - pos) // 5 + 16)
        ids = rng.choice(nw, size=k, p=p)
        wl = lens[ids] + 1                      # word + separator
        off = np.cumsum(wl) - wl
        total = int(off[-1] + wl[-1])
        rep = np.repeat(np.arange(k), wl)
        within = np.arange(total) - off[rep]
        chunk = mat[ids[rep], within]
        # separators: mostly space, sometimes punctuation+space is emulated by replacing the last letter
        seps = off + wl - 1
        r = rng.random(k)
        chunk[seps[r < 0.012]] = ord(",")
        chunk[seps[(r >= 0.012) & (r < 0.02)]] = ord(".")
        dig = rng.random(total) < 0.008
        chunk[dig] = rng.integers(ord("0"), ord("9") + 1, size=int(dig.sum()), dtype=np.uint8)
        # newlines: every 40..100 characters turn the next separator into '\n'
        nl = np.cumsum(rng.integers(40, 101, size=total // 40 + 2))
        nl = nl[nl < total]
        idx = np.searchsorted(seps, nl)
        idx = idx[idx < k]
        chunk[seps[idx]] = ord("\n")
        take = min(total, n + 64 - pos)
        out[pos:pos + take] = chunk[:take]
        pos += take
    return out[:n].copy()


def random_bytes(n, seed=7):
    return np.random.default_rng(seed).integers(0, 256, size=n, dtype=np.uint8)


def random_walk(n, seed=8):
    rng = np.random.default_rng(seed)
    steps = rng.integers(-1, 2, size=n, dtype=np.int64)
    return (np.cumsum(steps) + 128).astype(np.uint8)


def mix1m(seed=1, n=1_000_000):
    span = 64 * 1024
    parts = []
    k = 0
    while sum(p.size for p in parts) < n:
        kind = k % 3
        if kind == 0:
            parts.append(text(span, seed * 1000 + k))
        elif kind == 1:
            parts.append(random_bytes(span, seed * 1000 + k))
        else:
            parts.append(random_walk(span, seed * 1000 + k))
        k += 1
    return np.concatenate(parts)[:n].copy()


def repetitive(n, seed=3):
    """Long runs (1..10000), short-period repeats (period 2..64) and aaaa\\xfb-like material."""
    rng = np.random.default_rng(seed)
    parts = []
    total = 0
    while total < n:
        kind = rng.integers(0, 4)
        if kind == 0:      # long run of one byte
            a = np.full(int(rng.integers(1, 10001)), rng.integers(0, 256), dtype=np.uint8)
        elif kind == 1:    # short-period repeat, thousands of repetitions
            per = int(rng.integers(2, 65))
            pat = rng.integers(0, 256, size=per, dtype=np.uint8)
            a = np.tile(pat, int(rng.integers(200, 4000)))
        elif kind == 2:    # very long run: becomes aaaa\xfb aaaa\xfb ... after RLE1
            a = np.full(int(rng.integers(20000, 200000)), rng.integers(0, 256), dtype=np.uint8)
        else:              # a little text so that blocks are not purely periodic
            a = text(int(rng.integers(100, 3000)), int(rng.integers(1 << 30)))
        parts.append(a)
        total += a.size
    out = np.concatenate(parts)[:n].copy()
    # must not end in a run whose final RLE1 group is exactly 4: end with distinct bytes
    if n >= 4:
        out[-4:] = np.array([1, 2, 3, 5], dtype=np.uint8)
    return out


def _segment(job):
    kind, m, seed = job
    if kind == "text":
        return text(m, seed)
    if kind == "walk":
        return random_walk(m, seed)
    if kind == "random":
        return random_bytes(m, seed)
    if kind == "rep":
        return repetitive(m, seed)
    if kind == "markov":
        return _markov_segment(m, seed)
    raise ValueError(kind)


def _assemble(jobs, n, workers):
    """Concatenates the segments of `jobs` (kind, length, seed) into one array of n bytes.  workers > 1: the segments are
    generated by that many child interpreters writing into one shared file (plain subprocesses: neither a CUDA context
    nor the parent's __main__ is inherited)."""
    workers = min(workers or 1, len(jobs))
    if workers <= 1:
        out = np.empty(n, dtype=np.uint8)
        pos = 0
        for j in jobs:
            a = _segment(j)[:n - pos]
            out[pos:pos + a.size] = a
            pos += a.size
        return out
    import json
    import subprocess
    import sys
    import tempfile
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else None
    fd, path = tempfile.mkstemp(prefix="bz2b200_corpus_", dir=shm)
    try:
        os.ftruncate(fd, n)
        os.close(fd)
        offs, pos = [], 0
        for j in jobs:
            offs.append(pos)
            pos += j[1]
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        procs = []
        for w in range(workers):
            mine = [(jobs[i][0], int(jobs[i][1]), int(jobs[i][2]), int(offs[i])) for i in range(w, len(jobs), workers)]
            p = subprocess.Popen([sys.executable, "-c",
                                  "import sys; sys.path.insert(0, %r); from bzip2_rust_b200 import corpus; corpus._worker()" % root],
                                 stdin=subprocess.PIPE)
            p.stdin.write(json.dumps({"path": path, "n": n, "jobs": mine}).encode())
            p.stdin.close()
            procs.append(p)
        for p in procs:
            if p.wait() != 0:
                raise RuntimeError("corpus worker failed")
        return np.fromfile(path, dtype=np.uint8, count=n)
    finally:
        try:
            os.unlink(path)
        except OSError:
            pass


def _worker():
    import json
    import sys
    spec = json.loads(sys.stdin.read())
    mm = np.memmap(spec["path"], dtype=np.uint8, mode="r+", shape=(spec["n"],))
    for kind, m, seed, off in spec["jobs"]:
        a = _segment((kind, m, seed))[:spec["n"] - off]
        mm[off:off + a.size] = a
    mm.flush()


def default_workers():
    return max(1, min(32, (os.cpu_count() or 1)))


def mixed(n, seed=5, workers=1):
    span = 1 << 20
    jobs = []
    total = 0
    k = 0
    while total < n:
        jobs.append(("random" if k % 2 == 0 else "text", span, seed * 100 + k))
        total += span
        k += 1
    return _assemble(jobs, n, workers)


def corpus(n, seed=4, workers=1):
    """text model with a different seed per 32 MiB segment plus ~10% binary segments (config 4: 8 GB)."""
    seg = 32 << 20
    jobs = []
    total = 0
    k = 0
    while total < n:
        m = min(seg, n - total)
        jobs.append(("walk" if k % 10 == 9 else "text", m, seed * 100 + k))
        total += m
        k += 1
    return _assemble(jobs, n, workers)


def rep_segments(n, seed=3, workers=1, seg=16 << 20):
    """`repetitive` material in independently seeded 16 MiB segments (config 3 at 256 MB; the serial generator needs
    minutes at that size)."""
    jobs = []
    total = 0
    k = 0
    while total < n:
        m = min(seg, n - total)
        jobs.append(("rep", m, seed * 1000 + k))
        total += m
        k += 1
    out = _assemble(jobs, n, workers)
    if n >= 4:
        out[-4:] = np.array([1, 2, 3, 5], dtype=np.uint8)       # no exact-4 run at the very end
    return out


def _markov_segment(n, seed):
    """Words follow each other through a sparse transition table (8 successors per word, Zipf weighted): the same word
    PAIRS and TRIPLES recur, as in natural text."""
    rng = np.random.default_rng(seed)
    mat, lens = _vocab(rng, nwords=2048)
    nw = mat.shape[0]
    succ = rng.integers(0, nw, size=(nw, 8))
    pw = 1.0 / np.arange(1, 9) ** 1.2
    pw /= pw.sum()
    k = n // 4 + 16
    choice = rng.choice(8, size=k, p=pw)
    jump = rng.random(k) < 0.03                                 # now and then an unrelated word
    rnd = rng.integers(0, nw, size=k)
    ids = np.empty(k, dtype=np.int64)
    cur = int(rng.integers(0, nw))
    for i in range(k):                                          # a chain: inherently serial, ~1 us per word
        cur = int(rnd[i]) if jump[i] else int(succ[cur, choice[i]])
        ids[i] = cur
    wl = lens[ids] + 1
    off = np.cumsum(wl) - wl
    total = int(off[-1] + wl[-1])
    rep = np.repeat(np.arange(k), wl)
    within = np.arange(total) - off[rep]
    chunk = mat[ids[rep], within]
    seps = off + wl - 1
    r = rng.random(k)
    chunk[seps[r < 0.03]] = ord("\n")
    chunk[seps[(r >= 0.03) & (r < 0.05)]] = ord(",")
    out = chunk[:n]
    if out.size < n:
        out = np.concatenate([out, text(n - out.size, seed + 1)])
    return out.copy()


def markov(n, seed=6, workers=1, seg=8 << 20):
    jobs = []
    total = 0
    k = 0
    while total < n:
        m = min(seg, n - total)
        jobs.append(("markov", m, seed * 1000 + k))
        total += m
        k += 1
    return _assemble(jobs, n, workers)
