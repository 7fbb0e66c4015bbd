"""CPU-side checks of the drop-in boundary: the library loads, exports what include/plz4cu.h declares,
and refuses to work without a GPU instead of falling back."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from plz4_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_header_symbol_is_exported_and_bound():
    syms = _lib.header_symbols()
    assert len(syms) >= 20
    L = _lib.lib()
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/plz4cu.h but not exported by libplz4cu.so"
    assert set(syms) == set(_lib.SIGNATURES), "ctypes table out of sync with the header"


def test_header_cites_reference_lines():
    text = open(os.path.join(ROOT, "include", "plz4cu.h")).read()
    for needle in ["clz4.go:31-45", "clz4.go:47-60", "blk/blk.go:69-109", "blk/frame.go:114-127", "compress/compress.go:83-85"]:
        assert needle in text


def test_per_block_codes_agree_between_header_and_mirror():
    """The negative out_len codes are part of the boundary: the ctypes mirror must carry the header's values."""
    text = open(os.path.join(ROOT, "include", "plz4cu.h")).read()
    codes = {m.group(1): -int(m.group(2), 16) for m in re.finditer(r"#define (PLZ4CU_E_\w+)\s+\(\(int32_t\)-0x([0-9A-Fa-f]+)\)", text)}
    assert codes == {"PLZ4CU_E_BLOCKHASH": _lib.E_BLOCKHASH, "PLZ4CU_E_OVERFLOW": _lib.E_OVERFLOW, "PLZ4CU_E_STALL": _lib.E_STALL}


def test_compress_bound_matches_reference(port):
    L = _lib.lib()
    # block_test.go:338-353 (monotonic) + lz4.h:215 values quoted in SURVEY.md a14
    prev = 0
    for n in [0, 1, 15, 255, 256, 4096, 65536, 4 << 20, 0x7E000000]:
        b = L.plz4cu_compress_bound(n)
        assert b == port.compress_bound(n)
        assert b >= prev
        prev = b
    assert L.plz4cu_compress_bound(65536) == 65809
    assert L.plz4cu_compress_bound(4 << 20) == 4210768
    assert L.plz4cu_compress_bound(0) == 16
    assert L.plz4cu_compress_bound(0x7E000001) == 0


def test_host_logtext_is_deterministic_and_segmented():
    L = _lib.lib()
    a = np.empty(3 * 65536 + 123, dtype=np.uint8)
    b = np.empty(65536, dtype=np.uint8)
    L.plz4cu_gen_logtext_host(7, 10, C.c_void_p(a.ctypes.data), a.size)
    L.plz4cu_gen_logtext_host(7, 12, C.c_void_p(b.ctypes.data), b.size)
    assert a[2 * 65536: 3 * 65536].tobytes() == b.tobytes()       # segment 12 is the same bytes either way
    assert re.match(rb"^\d{10}\.\d{6} [A-Z]+ \[[a-z]+\] pid=\d+ tid=\d+ ", a[:80].tobytes())


def test_no_gpu_means_loud_failure_not_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = _lib.lib()
    assert L.plz4cu_device_count() == _lib.ERR_NODEVICE
    assert L.plz4cu_init(0) < 0
    src = np.frombuffer(b"hello hello hello hello", dtype=np.uint8)
    dst = np.zeros(64, dtype=np.uint8)
    r = L.plz4cu_compress_fast(C.c_void_p(src.ctypes.data), src.size, C.c_void_p(dst.ctypes.data), 64)
    assert r == _lib.INT32_MIN
    assert b"CUDA" in L.plz4cu_last_error() or b"device" in L.plz4cu_last_error()
    with pytest.raises(_lib.Plz4cuError):
        import plz4_b200
        plz4_b200.compress_block(b"hello")


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under plz4_b200/ may import, link or dlopen it."""
    pkg = os.path.join(ROOT, "plz4_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text, f"{f} mentions the oracle"
                assert "liborc" not in text and "libreflz4" not in text
