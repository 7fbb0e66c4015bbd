#!/usr/bin/env python3
"""Device-resident decode throughput of few-large-block launches (the team kernel's case): total bytes x block size,
with and without block-checksum verification.  Run with PLZ4CU_TEAM=0 for the one-warp-per-block kernel."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from plz4_b200 import _lib
from plz4_b200._lib import check
L = _lib.lib(); check(L.plz4cu_init(0))
dev = torch.device("cuda", 0)
p = lambda t: C.c_void_p(t.data_ptr())
print("PLZ4CU_TEAM =", os.environ.get("PLZ4CU_TEAM", "(default)"), " PLZ4CU_TEAM_DBG =", os.environ.get("PLZ4CU_TEAM_DBG", "-"),
      " PLZ4CU_TEAM_COPY =", os.environ.get("PLZ4CU_TEAM_COPY", "-"))
QUICK = bool(os.environ.get("TEAM_PROBE_QUICK"))
for total in ((256 << 20,) if QUICK else (256 << 20, 1 << 30)):
    src = torch.empty(total, dtype=torch.uint8, device=dev)
    check(L.plz4cu_gen_logtext_device(None, 0x504C5A34, 0, p(src), total))
    for bsz in ((4 << 20,) if QUICK else (262144, 1 << 20, 4 << 20)):
        nblk = total // bsz; stride = bsz + 16
        if nblk >= 1024:
            continue
        recs = torch.empty(nblk * stride, dtype=torch.uint8, device=dev); out = torch.zeros(total, dtype=torch.uint8, device=dev)
        off = torch.arange(nblk, dtype=torch.int64, device=dev) * bsz; ln = torch.full((nblk,), bsz, dtype=torch.int32, device=dev)
        roff = torch.arange(nblk, dtype=torch.int64, device=dev) * stride
        rl = torch.zeros(nblk, dtype=torch.int32, device=dev); ol = torch.zeros(nblk, dtype=torch.int32, device=dev)
        check(L.plz4cu_compress_batch_device(None, p(src), p(off), p(ln), nblk, bsz, 1, 0, None, p(recs), stride, p(rl)))
        row = []
        for verify in (0, 1):
            def dec(): check(L.plz4cu_decompress_batch_device(None, p(recs), p(roff), None, nblk, bsz, verify, 0, None, p(out), bsz, p(ol)))
            dec(); torch.cuda.synchronize()
            best = 1e9
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); dec(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1) / 1e3)
            if not os.environ.get("PLZ4CU_TEAM_DBG"):
                assert bool((ol == bsz).all()), ol[:8]
                assert torch.equal(out, src)
            row.append(total / best / 1e9)
        print(f"total {total >> 20:5d} MiB  bsz {bsz:8d}  blocks {nblk:4d}  decode {row[0]:7.2f} GB/s  with checksum {row[1]:7.2f} GB/s", flush=True)
    del src
