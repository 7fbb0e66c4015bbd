import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np
import plz4_b200 as P
from oracle import oracle as O
from tests.datagen import make, logtext
P.init(0)
ref = O.best()
for kind, n, seed in [("log", 4 << 20, 4), ("log", 1 << 20, 0), ("log", 262144, 0), ("log", 200000, 1), ("words", 1 << 20, 0)]:
    s = make(kind, n, seed=seed)
    packed, poff = P.compress_batch(s, [0], [n], P.compress_block_bound(n), raw_blocks=True)
    c = packed[:int(poff[1])].tobytes()
    r, data = ref.decompress(c, n)
    if r == n and data == s:
        print(kind, n, seed, "ok", len(c)); continue
    a = np.frombuffer(data, np.uint8) if r == n else None
    b = np.frombuffer(s, np.uint8)
    if a is None:
        print(kind, n, seed, "ret", r); continue
    bad = np.nonzero(a != b)[0]
    print(kind, n, seed, "first bad", bad[:5], "count", len(bad), "last", bad[-3:], "frag of first", bad[0] // 65536, bad[0] % 65536)
