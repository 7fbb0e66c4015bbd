#!/usr/bin/env python3
"""Latency of the per-block shims (one block per call): plz4cu_compress_fast / plz4cu_decompress_safe
on 4 KiB, 64 KiB and 4 MiB blocks of log text, beside the reference's liblz4 on one core."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import plz4_b200 as P
from plz4_b200 import _lib
from oracle import oracle
L = _lib.lib(); P.init(0)
ref = oracle.Ref()
vp = lambda a: C.c_void_p(a.ctypes.data)
for n in (4096, 65536, 4 << 20):
    src = np.empty(n, dtype=np.uint8); L.plz4cu_gen_logtext_host(0x504C5A34, 0, vp(src), n)
    cap = int(L.plz4cu_compress_bound(n)); dst = np.empty(cap, dtype=np.uint8); out = np.empty(n, dtype=np.uint8)
    reps = 300 if n <= 65536 else 30
    c = L.plz4cu_compress_fast(vp(src), n, vp(dst), cap); assert c > 0
    assert L.plz4cu_decompress_safe(vp(dst), c, vp(out), n) == n and (out == src).all()
    t0 = time.perf_counter()
    for _ in range(reps): L.plz4cu_compress_fast(vp(src), n, vp(dst), cap)
    tc = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for _ in range(reps): L.plz4cu_decompress_safe(vp(dst), c, vp(out), n)
    td = (time.perf_counter() - t0) / reps
    raw = src.tobytes()
    t0 = time.perf_counter()
    for _ in range(reps): rc = ref.compress(raw)
    rtc = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for _ in range(reps): ref.decompress(rc, n)
    rtd = (time.perf_counter() - t0) / reps
    print(f"block {n:8d} B: compress {tc * 1e6:8.0f} us  decompress {td * 1e6:8.0f} us   liblz4 on one core (through ctypes): {rtc * 1e6:8.0f} / {rtd * 1e6:8.0f} us", flush=True)
