#!/usr/bin/env python3
"""Compress throughput on data with LONG literal runs between matches (random text interleaved with repeats of
earlier data): the case where one lane of flush_queue would copy hundreds of literal bytes on its own."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from plz4_b200 import _lib
from plz4_b200._lib import check
L = _lib.lib(); check(L.plz4cu_init(0))
p = lambda t: C.c_void_p(t.data_ptr())
rng = np.random.default_rng(5)
BSZ = 65536; nblk = 8192; total = BSZ * nblk
for run, rep in ((24, 24), (200, 200), (1000, 300), (5000, 100)):
    # one 1 MiB pattern: `run` random bytes, then `rep` bytes copied from 3000 bytes back, ...
    pat = np.empty(1 << 20, dtype=np.uint8); pos = 0
    while pos < pat.size:
        k = min(run, pat.size - pos); pat[pos:pos + k] = rng.integers(32, 127, k, dtype=np.uint8); pos += k
        if pos >= pat.size: break
        k = min(rep, pat.size - pos); s = max(0, pos - 3000)
        pat[pos:pos + k] = pat[s:s + k]; pos += k
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
    print(f"literal runs {run:5d} / matches {rep:4d}: compress {total/ts[0]/1e9:7.2f} GB/s  decompress {total/ts[1]/1e9:7.2f} GB/s  ratio {float(rl.sum())/total:.4f}")
