#!/usr/bin/env python3
"""Device-resident kernel throughput vs block size (how parallelism-starved are large blocks?)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from plz4_b200 import _lib
from plz4_b200._lib import check
L = _lib.lib(); check(L.plz4cu_init(0))
dev = torch.device("cuda", 0)
total = 1 << 30
p = lambda t: C.c_void_p(t.data_ptr())
src = torch.empty(total, dtype=torch.uint8, device=dev)
check(L.plz4cu_gen_logtext_device(None, 0x504C5A34, 0, p(src), total))
for bsz in (65536, 262144, 1 << 20, 4 << 20):
    nblk = total // bsz; stride = bsz + 16
    recs = torch.empty(nblk * stride, dtype=torch.uint8, device=dev); out = torch.empty(total, dtype=torch.uint8, device=dev)
    off = torch.arange(nblk, dtype=torch.int64, device=dev) * bsz; ln = torch.full((nblk,), bsz, dtype=torch.int32, device=dev)
    roff = torch.arange(nblk, dtype=torch.int64, device=dev) * stride
    rl = torch.zeros(nblk, dtype=torch.int32, device=dev); ol = torch.zeros(nblk, dtype=torch.int32, device=dev)
    def comp(): check(L.plz4cu_compress_batch_device(None, p(src), p(off), p(ln), nblk, bsz, 1, 0, None, p(recs), stride, p(rl)))
    def dec(): check(L.plz4cu_decompress_batch_device(None, p(recs), p(roff), None, nblk, bsz, 1, 0, None, p(out), bsz, p(ol)))
    ts = []
    for f in (comp, dec):
        f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / 1e3)
    assert torch.equal(out, src) and bool((ol == bsz).all())
    print(f"bsz {bsz:8d} blocks {nblk:6d}  compress {total/ts[0]/1e9:7.2f} GB/s  decompress {total/ts[1]/1e9:7.2f} GB/s  ratio {float(rl.sum())/total:.4f}")
