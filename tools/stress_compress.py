#!/usr/bin/env python3
"""Seed sweep: GPU-compress many inputs, decode with the oracle, report mismatches and worst size ratio."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import plz4_b200 as P
from oracle import oracle as O
from tests.datagen import KINDS, make
P.init(0)
ref = O.best()
seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 20
worst = {}
bad = 0
for seed in range(seeds):
    for bsz, sizes in [(65536, [0, 1, 13, 100, 1000, 4096, 30000, 65535, 65536]), (262144, [70000, 262144]), (4 << 20, [1 << 20, 4 << 20])]:
        srcs = [make(k, n, seed=seed) for k in KINDS for n in sizes]
        buf = b"".join(srcs); lens = [len(s) for s in srcs]; off = np.cumsum([0] + lens)[:-1]
        packed, poff = P.compress_batch(buf, off, lens, P.compress_block_bound(bsz), raw_blocks=True)
        for i, s in enumerate(srcs):
            c = packed[int(poff[i]):int(poff[i + 1])].tobytes()
            r, data = ref.decompress(c, len(s))
            if r != len(s) or data != s:
                bad += 1
                print("MISMATCH seed", seed, "kind", KINDS[i // len(sizes)], "n", len(s), "ret", r)
                open(f"/tmp/bad_{seed}_{i}.bin", "wb").write(s)
                continue
            rc = ref.compress(s)
            k = (KINDS[i // len(sizes)], len(s))
            ratio = len(c) / max(len(rc), 1)
            if len(c) > len(rc) + 8 and ratio > worst.get(k, (0,))[0]:
                worst[k] = (ratio, seed, len(c), len(rc))
print("bad", bad)
for k, v in sorted(worst.items(), key=lambda kv: -kv[1][0])[:12]:
    print(k, "ratio %.4f seed %d gpu %d ref %d" % v)
