"""Seeded inputs shared by the parity tests."""
import numpy as np

from bzip2_rust_b200 import corpus


def small_cases():
    """(name, bytes) edge cases: tiny, single symbol, periodic (p | n and p !| n), runs, two symbols ..."""
    rng = np.random.default_rng(1234)
    cases = [
        ("one", b"a"), ("two", b"ab"), ("abc", b"abc"), ("abcd", b"abcd"), ("aaaaa", b"aaaaa"),
        ("hello", b"hello world\n"), ("acababab", b"acababab"), ("banana", b"banana"),
        ("mississippi", b"mississippi"), ("same8", b"a" * 8), ("same9", b"z" * 9), ("same1000", b"q" * 1000),
        ("ab_x500", b"ab" * 500), ("abc_x333", b"abc" * 333), ("abc_x333_d", b"abc" * 333 + b"d"),
        ("period7x64", b"abcdefg" * 64), ("period5_rle", b"aaaa\xfb" * 300), ("period5_rle_tail", b"aaaa\xfb" * 300 + b"aa"),
        ("sorted", bytes(range(256)) * 3), ("revsorted", bytes(range(255, -1, -1)) * 3),
        ("two_sym", bytes(rng.integers(0, 2, 5000, dtype=np.uint8) + 65)),
        ("four_sym", bytes(rng.integers(0, 4, 7001, dtype=np.uint8) + 65)),
        ("rand_4097", bytes(rng.integers(0, 256, 4097, dtype=np.uint8))),
        ("rand_4096", bytes(rng.integers(0, 256, 4096, dtype=np.uint8))),
        ("rand_12289", bytes(rng.integers(0, 256, 12289, dtype=np.uint8))),
        ("text_30k", corpus.text(30000, 11).tobytes()),
        ("walk_20k", corpus.random_walk(20000, 12).tobytes()),
        ("rep_50k", corpus.repetitive(50000, 13).tobytes()),
        ("long_repeat", (corpus.text(3000, 14).tobytes() * 9)[:25000]),
        ("fib", _fib(20000)),
    ]
    return cases


def _fib(n):
    a, b = b"a", b"ab"
    while len(b) < n:
        a, b = b, b + a
    return b[:n]
