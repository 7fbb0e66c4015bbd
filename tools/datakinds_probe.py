#!/usr/bin/env python3
"""Device-resident compress / decompress throughput per kind of data (tests/datagen.py kinds tiled to 512 MiB of
64 KiB blocks): where are the weak spots off the benchmark workload?"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from plz4_b200 import _lib
from plz4_b200._lib import check
from tests.datagen import make
L = _lib.lib(); check(L.plz4cu_init(0))
p = lambda t: C.c_void_p(t.data_ptr())
BSZ = 65536; nblk = 8192; total = BSZ * nblk
for kind in ("log", "words", "runs", "zeros", "record1025", "ab", "random"):
    pat = np.frombuffer(make(kind, 4 << 20, seed=3), dtype=np.uint8)
    src = torch.from_numpy(np.tile(pat, total // pat.size)).cuda()
    stride = BSZ + 16
    recs = torch.empty(nblk * stride, dtype=torch.uint8, device="cuda"); rl = torch.zeros(nblk, dtype=torch.int32, device="cuda")
    off = torch.arange(nblk, dtype=torch.int64, device="cuda") * BSZ; ln = torch.full((nblk,), BSZ, dtype=torch.int32, device="cuda")
    out = torch.empty(total, dtype=torch.uint8, device="cuda"); ol = torch.zeros(nblk, dtype=torch.int32, device="cuda")
    roff = torch.arange(nblk, dtype=torch.int64, device="cuda") * stride
    def comp(): check(L.plz4cu_compress_batch_device(None, p(src), p(off), p(ln), nblk, BSZ, 1, 0, None, p(recs), stride, p(rl)))
    def dec(): check(L.plz4cu_decompress_batch_device(None, p(recs), p(roff), None, nblk, BSZ, 1, 0, None, p(out), BSZ, p(ol)))
    ts = []
    for f in (comp, dec):
        f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / 1e3)
    assert torch.equal(out, src)
    print(f"{kind:12s} compress {total/ts[0]/1e9:7.2f} GB/s  decompress {total/ts[1]/1e9:7.2f} GB/s  ratio {float(rl.sum())/total:.4f}")
