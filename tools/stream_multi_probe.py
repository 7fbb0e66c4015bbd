#!/usr/bin/env python3
"""BASELINE configs[4] in one process: NewWriter / NewReader over mixed-entropy data, 256 KiB blocks, pageable buffers,
in-memory C endpoints, one stream spread over 1..N GPUs (opts.n_devices)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import plz4_b200 as P
from plz4_b200 import _lib
L = _lib.lib(); P.init(0)
from tools.stream_probe_lib import best, c_compress, c_decompress, vp
gib = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
n = int(gib * (1 << 30))
ndev = P.device_count()
P.init_devices(list(range(ndev)))
# mixed entropy: log text, random, zero pages, repeated records, 1 MiB each in rotation
data = np.empty(n, dtype=np.uint8)
L.plz4cu_gen_logtext_host(0x504C5A34, 0, vp(data), n)
rng = np.random.default_rng(3)
seg = 1 << 20
for i in range(n // seg):
    k = i % 10
    if k in (2, 8):
        data[i * seg:(i + 1) * seg] = rng.integers(0, 256, seg, dtype=np.uint8)
    elif k == 3:
        data[i * seg:(i + 1) * seg] = 0
    elif k in (4, 6):
        rec = rng.integers(0, 256, 1025, dtype=np.uint8)
        data[i * seg:(i + 1) * seg] = np.resize(rec, seg)
other = np.empty(n, dtype=np.uint8)
fbuf = np.empty(n + (16 << 20), dtype=np.uint8)
gb = lambda t: "%.2f GB/s (%.0f ms)" % (n / t / 1e9, t * 1e3)
for nd in sorted({1, 2, 4, ndev} & set(range(1, ndev + 1))):
    o = dict(block_size_idx=5, block_checksum=True, content_checksum=False, n_devices=nd)
    flen = c_compress(data, fbuf, **o)
    tw = best(lambda: c_compress(data, fbuf, **o), 2)
    tr = best(lambda: c_decompress(fbuf, flen, other, n_devices=nd), 2)
    assert (other == data).all()
    print(f"configs[4] {gib} GiB mixed, 256 KiB blocks, n_devices={nd}: write {gb(tw)}  read {gb(tr)}  ratio {flen / n:.3f}", flush=True)
