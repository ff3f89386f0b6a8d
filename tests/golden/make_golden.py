"""Generates tests/golden/streams.json with the CPU oracle (the reference is Rust and cannot run here, so
these vectors pin the ORACLE's behaviour at commit time; App. E entries come from SURVEY.md and three of
them equal libbz2's own encoder output).  Run from the repo root:  python tests/golden/make_golden.py"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import pyref as R          # noqa: E402
from bzip2_rust_b200 import corpus     # noqa: E402
from inputs import small_cases         # noqa: E402

APP_E = {
    "a": "425a683931415926535919939b6b00000001002000200021184682ee48a70a120332736d60",
    "ab": "425a6839314159265359e993fdcd000000010030002000210082b177245385090e993fdcd0",
    "abc": "425a6839314159265359648cbb73000000010038002000219819846177245385090648cbb730",
    "abcd": "425a68393141592653593d4c334b00000001003c00200021860c30dc45dc914e14240f530cd2c0",
    "aaaaa": "425a683931415926535944a4303d00000241002000200020002100820b17724538509044a4303d",
    "hello world\n": "425a68393141592653594eece8360000025180001040000644908020002200f4420182c0d4b618fc5dc914e142413bb3a0d8",
}


def main():
    out = {"app_e": APP_E, "small": [], "generated": []}
    for name, data in small_cases():
        if len(data) > 8000:
            continue
        try:
            s = R.compress_stream(data, 9, R.SPEC)
        except R.RefPanic:
            continue
        out["small"].append({"name": name, "input_hex": data.hex(), "level": 9, "stream_hex": s.hex()})
    gens = [("mix1m", 1, 1_000_000, 9), ("mix1m", 1, 1_000_000, 1), ("text", 2, 1_500_000, 5),
            ("repetitive", 3, 1_200_000, 9), ("mixed", 5, 1_100_000, 3)]
    for gen, seed, n, level in gens:
        data = getattr(corpus, gen)(n, seed).tobytes() if gen != "mix1m" else corpus.mix1m(seed, n).tobytes()
        s = R.compress_stream(data, level, R.SPEC_FAST, threads=8)
        blocks = [(c, hashlib.sha256(b).hexdigest(), cons) for c, b, _, cons in R.rle1_blocks(data, level)]
        out["generated"].append({"gen": gen, "seed": seed, "n": n, "level": level, "stream_len": len(s),
                                 "stream_sha256": hashlib.sha256(s).hexdigest(),
                                 "input_sha256": hashlib.sha256(data).hexdigest(),
                                 "blocks": [{"crc": c, "rle1_sha256": h, "consumed": cons} for c, h, cons in blocks]})
    with open(os.path.join(ROOT, "tests", "golden", "streams.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(out["small"]), "small +", len(out["generated"]), "generated vectors")


if __name__ == "__main__":
    main()
