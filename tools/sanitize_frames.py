#!/usr/bin/env python3
"""compute-sanitizer workload for the paths tools/sanitize_smoke.py does not reach: device-resident frames (parallel
frame walk forced onto small frames with PLZ4CU_WALK_MIN_BLOCKS=4), fragment-parallel encode of large blocks + stitch,
dictionary blocks, record packing.  Buffers are exactly sized device allocations."""
import ctypes as C, os, sys
os.environ.setdefault("PLZ4CU_WALK_MIN_BLOCKS", "4")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import plz4_b200 as P
from plz4_b200 import _lib
from tests.datagen import make
L = _lib.lib(); P.init(0)

def exact(b):
    """device tensor whose allocation ends right behind the data (rounded to the documented 16-byte granule)"""
    n = len(b)
    t = torch.empty((n + 15) // 16 * 16, dtype=torch.uint8, device="cuda")
    if n:
        t[:n] = torch.frombuffer(bytearray(b), dtype=torch.uint8).cuda()
    return t[:n]

torch.cuda.memory._set_allocator_settings("") if False else None
for kind in ["log", "random", "zeros", "words"]:
    for n in [0, 1, 70000, 600000, 3 << 20]:
        data = make(kind, n, seed=n & 7)
        d = exact(data)
        for bidx, bx in ((4, True), (4, False), (5, True), (7, True)):
            frame = P.compress_frame_device(d, block_size_idx=bidx, block_checksum=bx)
            tight = exact(bytes(frame.cpu().numpy()))
            out, info = P.decompress_frame_device(tight)
            assert bytes(out.cpu().numpy()) == data and info.frame_len == tight.numel(), (kind, n, bidx, bx)
dct = P.Dict(make("log", 65536, seed=1))
msgs = make("log", 64 * 4096, seed=2)
packed, poff = P.compress_batch(msgs, [i * 4096 for i in range(64)], [4096] * 64, P.compress_block_bound(4096), raw_blocks=True, dict=dct)
sizes = np.diff(poff).astype(np.uint32)
out, res = P.decompress_batch(packed, poff[:-1], 4096, raw_len=sizes, dict=dct)
assert (res == 4096).all() and out.reshape(-1).tobytes() == msgs
print("sanitize frames workload ok")
