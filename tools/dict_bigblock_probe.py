#!/usr/bin/env python3
"""Device-resident compress of large blocks with a 64 KiB dictionary (the span encoder with the dictionary as the fragment before
the block) beside the same blocks without one; PLZ4CU_SPANS=0 gives round 1's path (one warp per block with a dictionary)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import plz4_b200 as P
from plz4_b200 import _lib
from plz4_b200._lib import check
from tests.datagen import make
L = _lib.lib(); P.init(0)
p = lambda t: C.c_void_p(t.data_ptr())
total = 256 << 20
src = torch.empty(total, dtype=torch.uint8, device="cuda")
check(L.plz4cu_gen_logtext_device(None, 0x504C5A34, 0, p(src), total))
d = src[:65536].cpu().numpy().tobytes()
gd = P.Dict(d)
for bsz in (262144, 1 << 20, 4 << 20):
    nblk = total // bsz; stride = (int(L.plz4cu_compress_bound(bsz)) + 8 + 15) // 16 * 16
    recs = torch.empty(nblk * stride, dtype=torch.uint8, device="cuda")
    off = torch.arange(nblk, dtype=torch.int64, device="cuda") * bsz; ln = torch.full((nblk,), bsz, dtype=torch.int32, device="cuda")
    rl = torch.zeros(nblk, dtype=torch.int32, device="cuda")
    row = []
    for dct in (None, gd):
        h = C.c_void_p(dct.handle) if dct is not None else None
        def enc(): check(L.plz4cu_compress_batch_device(None, p(src), p(off), p(ln), nblk, bsz, 1, 0, h, p(recs), stride, p(rl)))
        enc(); torch.cuda.synchronize(); best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); enc(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1) / 1e3)
        row.append((total / best / 1e9, float(rl.sum().item()) / total))
    print(f"bsz {bsz:8d}  blocks {nblk:4d}  no dictionary {row[0][0]:7.2f} GB/s ratio {row[0][1]:.4f}   64 KiB dictionary {row[1][0]:7.2f} GB/s ratio {row[1][1]:.4f}", flush=True)
