#!/usr/bin/env python3
"""Small, allocation-tight workloads for compute-sanitizer: buffers are exactly sized device allocations
(cudaMalloc via plz4cu_device_alloc), so any out-of-bounds access of the kernels is reported."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from plz4_b200 import _lib
from plz4_b200._lib import check
from tests.datagen import make
L = _lib.lib(); check(L.plz4cu_init(0))
PAD = int(os.environ.get('PAD16', '0'))   # extra 16-byte granules behind every input buffer
def dev(nbytes): 
    p = L.plz4cu_device_alloc(nbytes); assert p; return p
import torch
def h2d(dptr, arr):
    t = torch.from_numpy(arr); 
    torch.cuda.current_stream().synchronize()
    C.cdll.LoadLibrary("libcudart.so.12") if False else None
    # use torch's runtime for the copy
    import torch.cuda
    tmp = t.cuda()
    C.memmove  # no-op
    cudamemcpy(dptr, tmp.data_ptr(), arr.nbytes)
from torch.utils.cpp_extension import CUDA_HOME  # noqa
rt = C.CDLL(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so.12"))
def cudamemcpy(dst, src, n, kind=4):
    r = rt.cudaMemcpy(C.c_void_p(dst), C.c_void_p(src), C.c_size_t(n), kind); assert r == 0, r
FAST = os.environ.get("SAN_FAST", "0")
KINDS_ = ["log", "words", "runs", "random", "zeros"] if FAST == "0" else (["log", "runs"] if FAST == "1" else ["log"])
SIZES_ = [0, 1, 13, 63, 64, 100, 4096, 65535, 65536, 100000, 300001, 2600000] if FAST == "0" else ([0, 13, 100, 4096, 65536, 100000, 300001, 1200000] if FAST == "1" else [4096, 65536, 300001])
MIS_ = [0, 1, 7] if FAST == "0" else ([0, 7] if FAST == "1" else [0])
for kind in KINDS_:
    for n in SIZES_:
        for misalign in MIS_:
            data = np.frombuffer(make(kind, n), dtype=np.uint8)
            d_src = dev((n + misalign + 15) // 16 * 16 + 16 * PAD); 
            if n: cudamemcpy(d_src + misalign, data.ctypes.data, n, 1)
            cap = int(L.plz4cu_compress_bound(n)); stride = (max(cap, n) + 8 + 15) // 16 * 16
            d_rec = dev(stride); d_off = dev(8); d_len = dev(4); d_rl = dev(4)
            off = np.array([misalign], dtype=np.uint64); ln = np.array([n], dtype=np.uint32)
            cudamemcpy(d_off, off.ctypes.data, 8, 1); cudamemcpy(d_len, ln.ctypes.data, 4, 1)
            check(L.plz4cu_compress_batch_device(None, d_src, d_off, d_len, 1, min(cap, 65536 if n <= 65536 else cap), 1, 0, None, d_rec, stride, d_rl))
            rl = np.zeros(1, dtype=np.uint32); cudamemcpy(rl.ctypes.data, d_rl, 4, 2)
            # decode from an exactly sized copy of the record into an exactly sized output
            rec = np.zeros(int(rl[0]), dtype=np.uint8); cudamemcpy(rec.ctypes.data, d_rec, int(rl[0]), 2)
            d_in = dev((int(rl[0]) + misalign + 15) // 16 * 16 + 16 * PAD); cudamemcpy(d_in + misalign, rec.ctypes.data, int(rl[0]), 1)
            d_out = dev(cap + 16); d_ol = dev(4)
            check(L.plz4cu_decompress_batch_device(None, d_in, d_off, None, 1, cap, 1, 0, None, d_out, cap + 16, d_ol))
            ol = np.zeros(1, dtype=np.int32); cudamemcpy(ol.ctypes.data, d_ol, 4, 2)
            out = np.zeros(max(n, 1), dtype=np.uint8); cudamemcpy(out.ctypes.data, d_out, max(n, 1), 2) if n else None
            assert ol[0] == n and out[:n].tobytes() == data.tobytes(), (kind, n, misalign, ol[0])
            for p in (d_src, d_rec, d_off, d_len, d_rl, d_in, d_out, d_ol): L.plz4cu_device_free(p)
# large blocks with a dictionary: the span encoder with the dictionary as the fragment before the block
for dn in ([65536, 777] if FAST != "2" else [777]):
    dct = make("log", dn, seed=3)
    L.plz4cu_dict_create.restype = C.c_void_p; L.plz4cu_dict_create.argtypes = [C.c_char_p, C.c_size_t]
    gd = L.plz4cu_dict_create(dct, len(dct)); assert gd
    for n in (70000, 300001):
        data = np.frombuffer(dct[-500:] + make("log", n - 500, seed=4), dtype=np.uint8)
        d_src = dev((n + 15) // 16 * 16); cudamemcpy(d_src, data.ctypes.data, n, 1)
        cap = int(L.plz4cu_compress_bound(n)); stride = (cap + 8 + 15) // 16 * 16
        d_rec = dev(stride); d_off = dev(8); d_len = dev(4); d_rl = dev(4)
        off = np.array([0], dtype=np.uint64); ln = np.array([n], dtype=np.uint32)
        cudamemcpy(d_off, off.ctypes.data, 8, 1); cudamemcpy(d_len, ln.ctypes.data, 4, 1)
        check(L.plz4cu_compress_batch_device(None, d_src, d_off, d_len, 1, cap, 1, 0, C.c_void_p(gd), d_rec, stride, d_rl))
        rl = np.zeros(1, dtype=np.uint32); cudamemcpy(rl.ctypes.data, d_rl, 4, 2)
        d_out = dev(cap + 16); d_ol = dev(4)
        check(L.plz4cu_decompress_batch_device(None, d_rec, d_off, None, 1, cap, 1, 0, C.c_void_p(gd), d_out, cap + 16, d_ol))
        ol = np.zeros(1, dtype=np.int32); cudamemcpy(ol.ctypes.data, d_ol, 4, 2)
        out = np.zeros(n, dtype=np.uint8); cudamemcpy(out.ctypes.data, d_out, n, 2)
        assert ol[0] == n and out.tobytes() == data.tobytes(), (dn, n, ol[0])
        for p in (d_src, d_rec, d_off, d_len, d_rl, d_out, d_ol): L.plz4cu_device_free(p)
print("sanitize workload ok")
