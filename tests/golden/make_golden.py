#!/usr/bin/env python3
"""Generates tests/golden/vectors.json from the REFERENCE's own liblz4 (oracle/_ref/libreflz4.so, compiled from
/root/reference/internal/pkg/clz4/lz4.c by oracle/Makefile), called with the argument patterns of clz4.go.

Run in the build container (needs oracle/_ref):   python tests/golden/make_golden.py
The vectors pin: level-1 block bytes for assorted inputs and capacities, dictionary-path bytes, decode return
codes (including malformed inputs), xxh32 values, and whole frames (via oracle/frame_oracle.py on top of the
same library).  Inputs are regenerated from seeds by tests/datagen.py, so the file stays small.
"""
import base64, hashlib, json, os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import Ref, Port
from oracle import frame_oracle as F
from tests.datagen import make

ref, port = Ref(), Port()
b64 = lambda b: base64.b64encode(b).decode()
sha = lambda b: hashlib.sha256(b).hexdigest()
out = {"source": "liblz4 1.10.0 vendored in prequel-dev/plz4 (internal/pkg/clz4/lz4.c), compiled -O3", "blocks": [], "dict_blocks": [], "decode": [], "xxh32": [], "frames": []}

for kind in ["log", "words", "runs", "ab", "zeros", "random"]:
    for n in [0, 1, 5, 12, 13, 64, 300, 4096, 65536, 70000]:
        s = make(kind, n)
        for cap in (None, n):
            c = ref.compress(s, cap)
            e = {"kind": kind, "n": n, "cap": cap, "len": None if c is None else len(c), "sha256": None if c is None else sha(c)}
            if c is not None and len(c) <= 96:
                e["bytes"] = b64(c)
            out["blocks"].append(e)

for dn in [3, 8, 1000, 65536, 70000]:
    d = make("words", dn, seed=3)
    rd = ref.dict_create(d)
    for n in [0, 13, 4096, 4097, 20000]:
        s = (d[-40:] + make("words", n, seed=3))[:n] if dn >= 64 else make("words", n, seed=3)
        c = rd.compress(s)
        out["dict_blocks"].append({"dict_n": dn, "n": n, "tail40": dn >= 64, "len": len(c), "sha256": sha(c)})

rng = random.Random(2024)
for i in range(200):
    n = rng.choice([0, 5, 13, 20, 64, 300, 1000])
    c = bytearray(ref.compress(make(rng.choice(["words", "ab", "runs", "random"]), n, seed=i)))
    m = rng.randrange(5)
    if m == 0 and c: c[rng.randrange(len(c))] = rng.getrandbits(8)
    elif m == 1 and c: c = c[: rng.randrange(len(c) + 1)]
    elif m == 2: c += rng.randbytes(rng.randint(1, 12))
    elif m == 3: c = bytearray(rng.randbytes(rng.randint(1, 40)))
    cap = n + rng.choice([0, 1, 12, 33, 100])
    z = port.lib.orc_dbg_zero_offset_hits()
    port.decompress(bytes(c), cap)
    if port.lib.orc_dbg_zero_offset_hits() != z:
        continue                                    # offset 0: liblz4's output is not defined
    r, data = ref.decompress(bytes(c), cap)
    out["decode"].append({"stream": b64(bytes(c)), "cap": cap, "ret": r, "sha256": None if data is None else sha(data)})

for n in [0, 1, 15, 16, 17, 100, 65536]:
    s = make("random", n)
    out["xxh32"].append({"n": n, "value": port.xxh32(s)})        # == xxhash.xxh32(seed=0), tests/test_oracle_vs_ref.py

for bi, bx, cx, n in [(4, False, False, 5), (4, True, True, 70000), (5, True, False, 300000), (7, False, True, 100)]:
    s = make("log", n)
    f = F.write_frame(s, F.Opts(block_idx=bi, block_checksum=bx, content_checksum=cx), ref)
    out["frames"].append({"block_idx": bi, "bx": bx, "cx": cx, "n": n, "len": len(f), "sha256": sha(f), "head": b64(f[:32])})

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "vectors.json"), "w") as fp:
    json.dump(out, fp, indent=0)
print({k: len(v) for k, v in out.items() if isinstance(v, list)})
