#!/usr/bin/env python3
"""Where does NewWriter / NewReader time go?  256 MiB log text, per block size: batch calls with pinned vs pageable
buffers, plain host memcpy, host xxh32, and the stream objects over in-memory C endpoints."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import plz4_b200 as P
from plz4_b200 import _lib, stream as S
from plz4_b200._lib import check
L = _lib.lib(); P.init(0)
vp = lambda a: C.c_void_p(a.ctypes.data); hp = lambda x: C.c_void_p(x.data_ptr()); fn = lambda f: C.cast(f, C.c_void_p)
n = int(sys.argv[1]) << 20 if len(sys.argv) > 1 else 256 << 20

from tools.stream_probe_lib import best, c_compress, c_decompress

data = np.empty(n, dtype=np.uint8); L.plz4cu_gen_logtext_host(0x504C5A34, 0, vp(data), n)
other = np.empty(n, dtype=np.uint8)
gb = lambda t: "%.2f GB/s (%.1f ms)" % (n / t / 1e9, t * 1e3)
print("host memcpy pageable->pageable  ", gb(best(lambda: np.copyto(other, data))))
pin = torch.empty(n, dtype=torch.uint8).pin_memory(); pin_np = pin.numpy()
print("host memcpy pageable->pinned    ", gb(best(lambda: np.copyto(pin_np, data))))
L.plz4cu_xxh32_host.restype = C.c_uint32
print("host xxh32 (1 core)             ", gb(best(lambda: L.plz4cu_xxh32_host(vp(data), n))))
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
def h2d(src):
    dev.copy_(src, non_blocking=True); torch.cuda.synchronize()
print("H2D from pinned                 ", gb(best(lambda: h2d(pin))))
pg = torch.from_numpy(data)
print("H2D from pageable               ", gb(best(lambda: h2d(pg))))

for bidx, bsz in ((4, 64 << 10), (5, 256 << 10), (7, 4 << 20)):
    nblk = n // bsz
    off = np.arange(nblk, dtype=np.uint64) * bsz; ln = np.full(nblk, bsz, dtype=np.uint32)
    poff = np.zeros(nblk + 1, dtype=np.uint64); res = np.zeros(nblk, dtype=np.int32)
    packed = torch.empty(nblk * (bsz + 8), dtype=torch.uint8).pin_memory()
    out = torch.empty(n, dtype=torch.uint8).pin_memory()
    np.copyto(pin_np, data)
    print("--- block size %d KiB (%d blocks)" % (bsz >> 10, nblk))
    tc = best(lambda: check(L.plz4cu_compress_batch_host(hp(pin), vp(off), vp(ln), nblk, bsz, 1, 0, None, hp(packed), packed.numel(), vp(poff))))
    c = int(poff[nblk])
    td = best(lambda: check(L.plz4cu_decompress_batch_host(hp(packed), c, vp(poff), None, nblk, bsz, 1, 0, None, hp(out), bsz, vp(res))))
    assert (res == bsz).all()
    print("batch_host pinned:    compress", gb(tc), " decompress", gb(td))
    pk_pg = np.empty(nblk * (bsz + 8), dtype=np.uint8)
    tc = best(lambda: check(L.plz4cu_compress_batch_host(vp(data), vp(off), vp(ln), nblk, bsz, 1, 0, None, vp(pk_pg), pk_pg.size, vp(poff))))
    td = best(lambda: check(L.plz4cu_decompress_batch_host(vp(pk_pg), c, vp(poff), None, nblk, bsz, 1, 0, None, vp(other), bsz, vp(res))))
    print("batch_host pageable:  compress", gb(tc), " decompress", gb(td))
    fbuf = np.empty(n + (1 << 20), dtype=np.uint8)
    for cx in (False, True):
        o = dict(block_size_idx=bidx, block_checksum=True, content_checksum=cx)
        flen = c_compress(data, fbuf, **o)
        tw = best(lambda: c_compress(data, fbuf, **o), 2)
        tw1 = best(lambda: c_compress(data, fbuf, chunk=1 << 20, **o), 2)
        flen = c_compress(data, fbuf, **o)            # the frame that is read back (batch boundaries may change the bytes)
        tr = best(lambda: c_decompress(fbuf, flen, other), 2)
        assert other.tobytes() == data.tobytes()
        print("stream cx=%d: write(one call)" % cx, gb(tw), " write(1 MiB calls)", gb(tw1), " read", gb(tr))
