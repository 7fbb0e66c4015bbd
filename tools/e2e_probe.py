#!/usr/bin/env python3
"""Where does the host-buffer path spend its time?  PCIe ceilings vs plz4cu_*_batch_host."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from plz4_b200 import _lib
from plz4_b200._lib import check
L = _lib.lib(); check(L.plz4cu_init(0))
BSZ = 65536; gib = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
n = int(gib * (1 << 30)) // BSZ * BSZ; nblk = n // BSZ
dev = torch.device("cuda", 0)
d = torch.empty(n, dtype=torch.uint8, device=dev)
check(L.plz4cu_gen_logtext_device(None, 0x504C5A34, 0, C.c_void_p(d.data_ptr()), n)); torch.cuda.synchronize()
h = torch.empty(n, dtype=torch.uint8).pin_memory(); h.copy_(d)
def t(f, reps=3):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
print("H2D pinned GB/s %.1f" % (n / t(lambda: d.copy_(h, non_blocking=True)) / 1e9))
print("D2H pinned GB/s %.1f" % (n / t(lambda: h.copy_(d, non_blocking=True)) / 1e9))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
d2 = torch.empty_like(d); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    s1.synchronize(); s2.synchronize()
print("H2D+D2H concurrent, each GB/s %.1f" % (n / t(both) / 1e9))
packed = torch.empty(nblk * (BSZ + 8), dtype=torch.uint8).pin_memory()
out = torch.empty(n, dtype=torch.uint8).pin_memory()
off = np.arange(nblk, dtype=np.uint64) * BSZ; ln = np.full(nblk, BSZ, dtype=np.uint32)
poff = np.zeros(nblk + 1, dtype=np.uint64); res = np.zeros(nblk, dtype=np.int32)
vp = lambda a: C.c_void_p(a.ctypes.data); hp = lambda x: C.c_void_p(x.data_ptr())
comp = lambda: check(L.plz4cu_compress_batch_host(hp(h), vp(off), vp(ln), nblk, BSZ, 1, 0, None, hp(packed), packed.numel(), vp(poff)))
tc = t(comp); c = int(poff[nblk])
dec = lambda: check(L.plz4cu_decompress_batch_host(hp(packed), c, vp(poff), None, nblk, BSZ, 1, 0, None, hp(out), BSZ, vp(res)))
td = t(dec)
assert (res == BSZ).all() and torch.equal(out, h)
print("compress_batch_host   %.1f ms  %.1f GB/s (ratio %.3f)" % (tc * 1e3, n / tc / 1e9, c / n))
print("decompress_batch_host %.1f ms  %.1f GB/s" % (td * 1e3, n / td / 1e9))
print("combined 2U/(tc+td)   %.1f GB/s" % (2 * n / (tc + td) / 1e9))
