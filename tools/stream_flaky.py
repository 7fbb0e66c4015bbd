#!/usr/bin/env python3
"""Writer or reader?  One frame written N times (each walked on the host), each read back M times."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import plz4_b200 as P
from plz4_b200 import _lib
L = _lib.lib(); P.init(0)
from tools.stream_probe_lib import c_compress, c_decompress, vp
n = (int(sys.argv[1]) if len(sys.argv) > 1 else 512) << 20
data = np.empty(n, dtype=np.uint8); L.plz4cu_gen_logtext_host(0x504C5A34, 0, vp(data), n)
other = np.empty(n, dtype=np.uint8)
fbuf = np.empty(n + (1 << 20), dtype=np.uint8)
def walk(buf, flen, bsz):
    pos, nb = 7, 0
    while True:
        if pos + 4 > flen: return f"ran off the end at {pos} after {nb} blocks"
        w = int.from_bytes(buf[pos:pos + 4].tobytes(), "little")
        if w == 0: break
        sz = w & 0x7FFFFFFF
        if sz > bsz: return f"size word {sz:#x} at {pos} (block {nb})"
        pos += 8 + sz; nb += 1
    return f"ok {nb} blocks, end {pos + 4} of {flen}"
for bidx, bsz in ((4, 65536), (7, 4 << 20)):
    for rep in range(4):
        flen = c_compress(data, fbuf, block_size_idx=bidx, block_checksum=True, content_checksum=False)
        print("bsz", bsz, "write", rep, "flen", flen, walk(fbuf, flen, bsz), flush=True)
        for r in range(3):
            try:
                m = c_decompress(fbuf, flen, other)
                ok = m == n and bool((other == data).all())
                print("   read", r, "->", m, "match" if ok else "MISMATCH", flush=True)
            except AssertionError as e:
                print("   read", r, "error", e, flush=True)
