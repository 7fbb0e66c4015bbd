#!/usr/bin/env python3
"""Compress the same device buffer many times: every run must give the same record lengths and bytes."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from plz4_b200 import _lib
from plz4_b200._lib import check
L = _lib.lib(); check(L.plz4cu_init(0))
for BSZ, gib in ((65536, 2.0), (262144, 1.0), (4 << 20, 1.0)):
    dev = torch.device("cuda", 0)
    n = int(gib * (1 << 30)) // BSZ * BSZ; nblk = n // BSZ; stride = BSZ + 16
    src = torch.empty(n, dtype=torch.uint8, device=dev)
    recs = [torch.zeros(nblk * stride, dtype=torch.uint8, device=dev) for _ in range(2)]
    rlen = [torch.zeros(nblk, dtype=torch.int32, device=dev) for _ in range(2)]
    off = torch.arange(nblk, dtype=torch.int64, device=dev) * BSZ
    ln = torch.full((nblk,), BSZ, dtype=torch.int32, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())
    check(L.plz4cu_gen_logtext_device(None, 0x504C5A34, 0, p(src), n))
    bad = 0
    for it in range(12):
        k = it & 1
        recs[k].zero_()
        check(L.plz4cu_compress_batch_device(None, p(src), p(off), p(ln), nblk, BSZ, 1, 0, None, p(recs[k]), stride, p(rlen[k])))
        torch.cuda.synchronize()
        if it and not (torch.equal(rlen[0], rlen[1]) and torch.equal(recs[0], recs[1])):
            d = (rlen[0] != rlen[1]).nonzero().flatten()
            bad += 1
            print("  run", it, "differs: blocks with other lengths", d[:6].tolist(), "count", int(d.numel()))
    print("bsz", BSZ, "blocks", nblk, "runs 12", "DIFFERENCES %d" % bad if bad else "identical")
