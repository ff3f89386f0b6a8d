"""Seeded synthetic inputs for the BASELINE.json configs.  numpy only; deterministic.

    mix1m(seed=1)        1 000 000 B: alternating 64 KiB spans of text / random / random-walk binary
    text(n, seed=2)      text-like: Zipf-weighted words from a 4096-word vocabulary, lines of 40-100
    repetitive(n, seed=3) long runs, short-period repeats, `aaaa\\xfb`-style post-RLE1 pathologies
    mixed(n, seed=5)     50/50 random bytes and text in 1 MiB spans (block-size sweep)
"""
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


def mixed(n, seed=5):
    span = 1 << 20
    parts = []
    k = 0
    total = 0
    while total < n:
        a = random_bytes(span, seed * 100 + k) if k % 2 == 0 else text(span, seed * 100 + k)
        parts.append(a)
        total += a.size
        k += 1
    return np.concatenate(parts)[:n].copy()


def corpus(n, seed=4):
    """text model with a different seed per 256 MB segment plus ~10% binary segments (config 4)."""
    seg = 32 << 20
    parts = []
    total = 0
    k = 0
    while total < n:
        m = min(seg, n - total)
        if k % 10 == 9:
            a = random_walk(m, seed * 100 + k)
        else:
            a = text(m, seed * 100 + k)
        parts.append(a)
        total += m
        k += 1
    return np.concatenate(parts)
