"""Compress-kernel probe: device-resident log text, 64 KiB blocks; time, size and round trip of launch_compress.
PLZ4CU_CTA_MIN=0 selects the one-warp-per-block kernel, the default the CTA-per-block kernel."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from plz4_b200 import _lib
from plz4_b200._lib import check

BSZ = 64 << 10
gib = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
L = _lib.lib()
check(L.plz4cu_init(0), "init")
dev = torch.device("cuda", 0)
nbytes = int(gib * (1 << 30)) // BSZ * BSZ
nblk = nbytes // BSZ
stride = BSZ + 16
st = torch.cuda.current_stream()
sh = C.c_void_p(st.cuda_stream)
u8 = lambda n: torch.empty(n, dtype=torch.uint8, device=dev)
src, recs, out = u8(nbytes), u8(nblk * stride), u8(nbytes)
src_off = torch.arange(nblk, dtype=torch.int64, device=dev) * BSZ
src_len = torch.full((nblk,), BSZ, dtype=torch.int32, device=dev)
rec_off = torch.arange(nblk, dtype=torch.int64, device=dev) * stride
rec_len = torch.zeros(nblk, dtype=torch.int32, device=dev)
out_len = torch.zeros(nblk, dtype=torch.int32, device=dev)
p = lambda t: C.c_void_p(t.data_ptr())
check(L.plz4cu_gen_logtext_device(sh, 0x504C5A34, 0, p(src), nbytes), "gen")
torch.cuda.synchronize()


def compress():
    check(L.plz4cu_compress_batch_device(sh, p(src), p(src_off), p(src_len), nblk, BSZ, 1, 0, None, p(recs), stride, p(rec_len)), "c")


def decompress():
    check(L.plz4cu_decompress_batch_device(sh, p(recs), p(rec_off), None, nblk, BSZ, 1, 0, None, p(out), BSZ, p(out_len)), "d")


compress(); decompress(); torch.cuda.synchronize()
ok = bool((out_len == BSZ).all()) and torch.equal(out, src)
csize = int(rec_len.to(torch.int64).sum())
best = 1e9
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st); compress(); e1.record(st); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(f"cta_min={os.environ.get('PLZ4CU_CTA_MIN', 'default')} blocks={nblk} roundtrip={'ok' if ok else 'FAIL'} "
      f"ratio={csize / nbytes:.5f} compress {best:.2f} ms = {nbytes / best / 1e6:.1f} GB/s")
if not ok:
    bad = (out_len != BSZ).nonzero().flatten()[:8].tolist()
    print("bad blocks", bad, out_len[bad].tolist() if bad else "")
    sys.exit(1)
if hasattr(L, "plz4cu_debug_cta_prof") or True:
    try:
        buf = (C.c_ulonglong * 16)()
        if L.plz4cu_debug_cta_prof(buf, 1) == 0:
            names = ["pre-token", "wait token", "hold token", "verify", "parse", "coop", "wait entry", "sizes", "emit", "wait out"]
            tot = sum(buf[:10]) or 1
            ntile = nblk * 64 * (reps + 1)
            print("stage cycles per tile (lane 0 of the tile's worker): " + ", ".join(f"{n} {buf[i] / ntile:.0f}" for i, n in enumerate(names)) + f"  | total {tot / ntile:.0f}")
    except AttributeError:
        pass
