import ctypes as C, os, sys
sys.path.insert(0, "/root/repo")
import torch
from plz4_b200 import _lib
from plz4_b200._lib import check
L = _lib.lib(); check(L.plz4cu_init(0))
dev = torch.device("cuda", 0); total = 1 << 29; bsz = 4 << 20
p = lambda t: C.c_void_p(t.data_ptr())
src = torch.empty(total, dtype=torch.uint8, device=dev)
check(L.plz4cu_gen_logtext_device(None, 0x504C5A34, 0, p(src), total))
nblk = total // bsz; stride = bsz + 16
recs = torch.empty(nblk * stride, dtype=torch.uint8, device=dev); out = torch.empty(total, dtype=torch.uint8, device=dev)
off = torch.arange(nblk, dtype=torch.int64, device=dev) * bsz; ln = torch.full((nblk,), bsz, dtype=torch.int32, device=dev)
roff = torch.arange(nblk, dtype=torch.int64, device=dev) * stride
rl = torch.zeros(nblk, dtype=torch.int32, device=dev); ol = torch.zeros(nblk, dtype=torch.int32, device=dev)
check(L.plz4cu_compress_batch_device(None, p(src), p(off), p(ln), nblk, bsz, 1, 0, None, p(recs), stride, p(rl)))
for _ in range(2): check(L.plz4cu_decompress_batch_device(None, p(recs), p(roff), None, nblk, bsz, 1, 0, None, p(out), bsz, p(ol)))
torch.cuda.synchronize(); assert torch.equal(out, src)
