#!/usr/bin/env python3
"""Reader stage times (PLZ4CU_STREAM_PROF=1): one 2 GiB frame of 64 KiB blocks read back through NewReader.WriteTo."""
import os, sys, time
os.environ["PLZ4CU_STREAM_PROF"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import plz4_b200 as P
from plz4_b200 import _lib
L = _lib.lib(); P.init(0)
from tools.stream_probe_lib import c_compress, c_decompress, vp
n = (int(sys.argv[1]) if len(sys.argv) > 1 else 2048) << 20
bidx = int(sys.argv[2]) if len(sys.argv) > 2 else 4
data = np.empty(n, dtype=np.uint8); L.plz4cu_gen_logtext_host(0x504C5A34, 0, vp(data), n)
other = np.empty(n, dtype=np.uint8); fbuf = np.empty(n + (1 << 20), dtype=np.uint8)
flen = c_compress(data, fbuf, block_size_idx=bidx, block_checksum=True, content_checksum=False)
pend = int(os.environ.get("READ_PROF_PENDING_MIB", "0")) << 20
for rep in range(3):
    t0 = time.perf_counter(); c_decompress(fbuf, flen, other, pending_size=pend); t = time.perf_counter() - t0
    print(f"read {n / t / 1e9:.2f} GB/s ({t * 1e3:.0f} ms)", flush=True)
