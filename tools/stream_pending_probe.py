#!/usr/bin/env python3
"""Stream throughput vs batch size (WithPendingSize) on a 2 GiB log-text stream, 64 KiB blocks, no content checksum."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import plz4_b200 as P
from plz4_b200 import _lib
from tools.stream_probe_lib import c_compress, c_decompress, best, vp
L = _lib.lib(); P.init(0)
n = (int(sys.argv[1]) if len(sys.argv) > 1 else 2048) << 20
bidx = int(sys.argv[2]) if len(sys.argv) > 2 else 4
data = np.empty(n, dtype=np.uint8); L.plz4cu_gen_logtext_host(0x504C5A34, 0, vp(data), n)
fbuf = np.empty(n + (1 << 20), dtype=np.uint8); other = np.empty(n, dtype=np.uint8)
for pend in (16, 32, 64, 128, 256, 512):
    o = dict(block_size_idx=bidx, block_checksum=True, content_checksum=False, pending_size=pend << 20)
    flen = c_compress(data, fbuf, **o)
    tw = best(lambda: c_compress(data, fbuf, **o), 2)
    tr = best(lambda: c_decompress(fbuf, flen, other, pending_size=pend << 20), 2)
    print("pending %4d MiB: write %.2f GB/s  read %.2f GB/s" % (pend, n / tw / 1e9, n / tr / 1e9), flush=True)
