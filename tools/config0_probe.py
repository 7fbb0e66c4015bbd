#!/usr/bin/env python3
"""BASELINE configs[0] read side, every pass printed: 256 MiB log text, 4 MiB blocks, block checksums, with and without the
content checksum, NewReader.WriteTo over in-memory endpoints (PLZ4CU_STREAM_PROF=1 adds the stage times)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import plz4_b200 as P
from plz4_b200 import _lib
L = _lib.lib(); P.init(0)
from tools.stream_probe_lib import c_compress, c_decompress, vp
n = 256 << 20
data = np.empty(n, dtype=np.uint8); L.plz4cu_gen_logtext_host(0x504C5A34, 0, vp(data), n)
fbuf = np.empty(n + (1 << 20), dtype=np.uint8); obuf = np.empty(n, dtype=np.uint8)
for cx in (True, False):
    flen = c_compress(data, fbuf, block_size_idx=7, block_checksum=True, content_checksum=cx, parallel=-1)
    ts = []
    for _ in range(8):
        t0 = time.perf_counter(); c_decompress(fbuf, flen, obuf, parallel=-1); ts.append(time.perf_counter() - t0)
    print("content checksum", cx, "read ms:", " ".join("%.1f" % (t * 1e3) for t in ts), flush=True)
