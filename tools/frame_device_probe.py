#!/usr/bin/env python3
"""Device-resident frame decode: time of the parallel frame walk (plz4cu_frame_index_device) against the decode it
feeds, on the bench workload held as ONE LZ4 frame in device memory.  usage: tools/frame_device_probe.py [GiB]"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import plz4_b200 as P
from plz4_b200 import _lib
from plz4_b200._lib import check
L = _lib.lib(); P.init(0)
BSZ = 65536; gib = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
n = int(gib * (1 << 30)) // BSZ * BSZ; nblk = n // BSZ
dp = lambda t: C.c_void_p(t.data_ptr())
src = torch.empty(n, dtype=torch.uint8, device="cuda")
check(L.plz4cu_gen_logtext_device(None, 0x504C5A34, 0, dp(src), n))
off = torch.arange(nblk, dtype=torch.int64, device="cuda") * BSZ
ln = torch.full((nblk,), BSZ, dtype=torch.int32, device="cuda")
stride = BSZ + 16
recs = torch.empty(nblk * stride, dtype=torch.uint8, device="cuda"); rl = torch.empty(nblk, dtype=torch.int32, device="cuda")
check(L.plz4cu_compress_batch_device(None, dp(src), dp(off), dp(ln), nblk, BSZ, 1, 0, None, dp(recs), stride, dp(rl)))
total = int(rl.to(torch.int64).sum())
hdr = bytes([0x04, 0x22, 0x4d, 0x18, 0x70, 0x40])                      # version 1, independent, block checksums; 64 KiB
hdr += bytes([(L.plz4cu_xxh32_host(hdr[4:], 2) >> 8) & 0xFF])
frame = torch.empty(len(hdr) + total + 4, dtype=torch.uint8, device="cuda")
frame[: len(hdr)] = torch.tensor(list(hdr), dtype=torch.uint8, device="cuda")
poff = torch.empty(nblk + 1, dtype=torch.int64, device="cuda")
body = frame[len(hdr): len(hdr) + total]
check(L.plz4cu_pack_records_device(None, dp(recs), stride, dp(rl), nblk, dp(body), dp(poff)))
frame[len(hdr) + total:] = 0                                             # EndMark
torch.cuda.synchronize(); del recs
print("frame: %.3f GB for %.3f GB of data, %d blocks" % (frame.numel() / 1e9, n / 1e9, nblk))

rec_off = torch.empty(nblk, dtype=torch.int64, device="cuda"); out_len = torch.empty(nblk, dtype=torch.int32, device="cuda")
dst = torch.empty(n + 16, dtype=torch.uint8, device="cuda")
nb, end = C.c_uint64(), C.c_uint64()
def index():
    check(L.plz4cu_frame_index_device(None, C.c_void_p(frame.data_ptr() + len(hdr)), total + 4, BSZ, 1, dp(rec_off), nblk, C.byref(nb), C.byref(end)))
def best(f, reps=5):
    f(); t = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize(); t.append(time.perf_counter() - t0)
    return min(t)
ti = best(index)
assert nb.value == nblk and torch.equal(rec_off, poff[:nblk])
info = _lib.FrameInfo()
def whole():
    rc = L.plz4cu_decompress_frame_device(None, dp(frame), frame.numel(), None, dp(dst), n, dp(rec_off), dp(out_len), nblk, C.byref(info))
    assert rc == 0, rc
tw = best(whole)
assert info.out_bytes == n and torch.equal(dst[:n], src)
def decode_only():
    check(L.plz4cu_decompress_batch_device(None, C.c_void_p(frame.data_ptr() + len(hdr)), dp(rec_off), None, nblk, BSZ, 1, 0, None, dp(dst), BSZ, dp(out_len)))
td = best(decode_only)
print("frame walk on device  %.2f ms  (%.0f GB/s of frame bytes)%s" % (ti * 1e3, frame.numel() / ti / 1e9,
      "  [PLZ4CU_SERIAL_WALK: one thread chasing size words]" if os.environ.get("PLZ4CU_SERIAL_WALK") else ""))
print("decode kernel alone   %.2f ms  (%.1f GB/s)" % (td * 1e3, n / td / 1e9))
print("whole frame, device to device  %.2f ms  (%.1f GB/s of decoded bytes)" % (tw * 1e3, n / tw / 1e9))
