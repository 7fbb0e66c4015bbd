"""ctypes front-end for the CPU oracle (TEST INFRASTRUCTURE — never imported by plz4_b200/).

Two libraries are exposed with the same Python surface:

* ``Port``  — oracle/liborc.so, our C restatement (oracle/lz4_port.c).
* ``Ref``   — oracle/_ref/libreflz4.so, the reference's own vendored liblz4 1.10.0
              (internal/pkg/clz4/lz4.c) compiled by oracle/Makefile; called with exactly the
              argument patterns of internal/pkg/clz4/clz4.go.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_u8p = C.POINTER(C.c_uint8)


def build(quiet: bool = True) -> None:
    """Compile liborc.so / cpu_driver.so (and _ref/ when the reference checkout is present)."""
    subprocess.run(["make", "-C", _HERE], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _buf(b):
    """bytes/bytearray/memoryview/numpy -> (ctypes pointer, length, keepalive)."""
    import numpy as np
    if isinstance(b, np.ndarray):
        a = np.ascontiguousarray(b).view(np.uint8).reshape(-1)
        return a.ctypes.data_as(C.c_void_p), a.size, a
    if isinstance(b, (bytes, bytearray, memoryview)):
        a = np.frombuffer(b, dtype=np.uint8)
        return a.ctypes.data_as(C.c_void_p), a.size, a
    raise TypeError(type(b))


class Port:
    """oracle/liborc.so"""

    def __init__(self):
        path = os.path.join(_HERE, "liborc.so")
        if not os.path.exists(path):
            build()
        self.lib = L = C.CDLL(path)
        L.orc_compress_bound.restype = C.c_int
        L.orc_compress_bound.argtypes = [C.c_int]
        L.orc_xxh32.restype = C.c_uint32
        L.orc_xxh32.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_compress_fast.restype = C.c_int
        L.orc_compress_fast.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orc_dict_create.restype = C.c_void_p
        L.orc_dict_create.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_dict_destroy.argtypes = [C.c_void_p]
        L.orc_compress_dict.restype = C.c_int
        L.orc_compress_dict.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orc_decompress_safe.restype = C.c_int
        L.orc_decompress_safe.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orc_decompress_dict.restype = C.c_int
        L.orc_decompress_dict.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orc_block_record.restype = C.c_int
        L.orc_block_record.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]

    kind = "port"

    def compress_bound(self, n: int) -> int:
        return self.lib.orc_compress_bound(n)

    def xxh32(self, data) -> int:
        p, n, _k = _buf(data)
        return self.lib.orc_xxh32(p, n)

    def compress(self, src, cap: int | None = None) -> bytes | None:
        """LZ4_compress_fast(acc=1); None when it does not fit `cap` (ret 0)."""
        p, n, _k = _buf(src)
        if cap is None:
            cap = self.compress_bound(n)
        dst = C.create_string_buffer(max(cap, 1))
        r = self.lib.orc_compress_fast(p, n, dst, cap)
        return None if r == 0 else dst.raw[:r]

    def decompress(self, src, cap: int):
        """LZ4_decompress_safe -> (ret, bytes|None)."""
        p, n, _k = _buf(src)
        dst = C.create_string_buffer(max(cap, 1))
        r = self.lib.orc_decompress_safe(p, n, dst, cap)
        return r, (dst.raw[:r] if r >= 0 else None)

    def dict_create(self, d) -> "PortDict":
        return PortDict(self, d)

    def block_record(self, src, bsz: int, checksum: bool, dict_: "PortDict | None" = None) -> bytes:
        p, n, _k = _buf(src)
        rec = C.create_string_buffer(bsz + 8)
        r = self.lib.orc_block_record(dict_.h if dict_ else None, p, n, bsz, int(checksum), rec)
        return rec.raw[:r]


class PortDict:
    def __init__(self, port: Port, d):
        self.port = port
        p, n, _k = _buf(d)
        self.data = bytes(_k[-65536:].tobytes()) if n else b""
        self.h = port.lib.orc_dict_create(p, n)

    def compress(self, src, cap: int | None = None) -> bytes | None:
        p, n, _k = _buf(src)
        if cap is None:
            cap = self.port.compress_bound(n)
        dst = C.create_string_buffer(max(cap, 1))
        r = self.port.lib.orc_compress_dict(self.h, p, n, dst, cap)
        return None if r == 0 else dst.raw[:r]

    def decompress(self, src, cap: int):
        p, n, _k = _buf(src)
        dp, dn, _dk = _buf(self.data)
        dst = C.create_string_buffer(max(cap, 1))
        r = self.port.lib.orc_decompress_dict(p, n, dst, cap, dp, dn)
        return r, (dst.raw[:r] if r >= 0 else None)

    def __del__(self):
        try:
            self.port.lib.orc_dict_destroy(self.h)
        except Exception:
            pass


class Ref:
    """oracle/_ref/libreflz4.so — the reference's own liblz4, driven like clz4.go does."""

    kind = "reference"

    def __init__(self):
        path = os.path.join(_HERE, "_ref", "libreflz4.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = L = C.CDLL(path)
        L.LZ4_compressBound.restype = C.c_int
        L.LZ4_compressBound.argtypes = [C.c_int]
        L.LZ4_compress_fast.restype = C.c_int
        L.LZ4_compress_fast.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.LZ4_decompress_safe.restype = C.c_int
        L.LZ4_decompress_safe.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.LZ4_decompress_safe_usingDict.restype = C.c_int
        L.LZ4_decompress_safe_usingDict.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.LZ4_compress_HC.restype = C.c_int
        L.LZ4_compress_HC.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.LZ4_resetStream_fast.argtypes = [C.c_void_p]
        L.LZ4_loadDictSlow.restype = C.c_int
        L.LZ4_loadDictSlow.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.LZ4_attach_dictionary.argtypes = [C.c_void_p, C.c_void_p]
        L.LZ4_compress_fast_continue.restype = C.c_int
        L.LZ4_compress_fast_continue.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]

    @staticmethod
    def available() -> bool:
        return os.path.exists(os.path.join(_HERE, "_ref", "libreflz4.so"))

    def compress_bound(self, n: int) -> int:
        return self.lib.LZ4_compressBound(n)

    def compress(self, src, cap: int | None = None) -> bytes | None:
        p, n, _k = _buf(src)
        if cap is None:
            cap = self.compress_bound(n)
        dst = C.create_string_buffer(max(cap, 1) + 8)
        r = self.lib.LZ4_compress_fast(p if n else None, dst, n, cap, 1)
        return None if r == 0 else dst.raw[:r]

    def compress_hc(self, src, level: int, cap: int | None = None) -> bytes | None:
        p, n, _k = _buf(src)
        if cap is None:
            cap = self.compress_bound(n)
        dst = C.create_string_buffer(max(cap, 1) + 8)
        r = self.lib.LZ4_compress_HC(p if n else None, dst, n, cap, level)
        return None if r == 0 else dst.raw[:r]

    def decompress(self, src, cap: int):
        p, n, _k = _buf(src)
        dst = C.create_string_buffer(max(cap, 1) + 8)
        r = self.lib.LZ4_decompress_safe(p if n else None, dst if cap else None, n, cap)
        return r, (dst.raw[:r] if r >= 0 else None)

    def dict_create(self, d) -> "RefDict":
        return RefDict(self, d)

    def xxh32(self, data) -> int:
        """The reference's xxh32 is Go; the pinned port (== python-xxhash, tests/test_oracle_vs_ref.py) stands in."""
        if not hasattr(self, "_port"):
            self._port = Port()
        return self._port.xxh32(data)

    def block_record(self, src, bsz: int, checksum: bool, dict_: "RefDict | None" = None) -> bytes:
        """blk.CompressToBlk (blk/blk.go:69-109) on top of the reference's liblz4."""
        src = bytes(src)
        c = dict_.compress(src, bsz) if dict_ else self.compress(src, bsz)
        if c is None:
            c, word = src, len(src) | 0x80000000
        else:
            word = len(c)
        rec = word.to_bytes(4, "little") + c
        if checksum:
            rec += self.xxh32(c).to_bytes(4, "little")
        return rec


class RefDict:
    """clz4.go:96-120 NewDictCtx + :151-179 StreamIndieCtx (fresh working ctx per call)."""

    def __init__(self, ref: Ref, d):
        self.ref = ref
        p, n, k = _buf(d)
        self.data = bytes(k[-65536:].tobytes()) if n else b""    # compress/dict.go:43-56
        self._dbuf = C.create_string_buffer(self.data, max(len(self.data), 1))
        self._strm = C.create_string_buffer(16416 + 16)
        self._sp = (C.addressof(self._strm) + 15) & ~15
        ref.lib.LZ4_resetStream_fast(self._sp)
        ref.lib.LZ4_loadDictSlow(self._sp, self._dbuf if self.data else None, len(self.data))
        self._work = C.create_string_buffer(16416 + 16)
        self._wp = (C.addressof(self._work) + 15) & ~15

    def compress(self, src, cap: int | None = None, reuse_ctx: bool = True) -> bytes | None:
        p, n, _k = _buf(src)
        if cap is None:
            cap = self.ref.compress_bound(n)
        if not reuse_ctx:
            C.memset(self._wp, 0, 16416)
        dst = C.create_string_buffer(max(cap, 1) + 8)
        L = self.ref.lib
        L.LZ4_resetStream_fast(self._wp)
        L.LZ4_attach_dictionary(self._wp, self._sp)
        r = L.LZ4_compress_fast_continue(self._wp, p if n else None, dst, n, cap, 1)
        return None if r == 0 else dst.raw[:r]

    def decompress(self, src, cap: int):
        p, n, _k = _buf(src)
        dst = C.create_string_buffer(max(cap, 1) + 8)
        r = self.ref.lib.LZ4_decompress_safe_usingDict(
            p if n else None, dst, n, cap, self._dbuf if self.data else None, len(self.data))
        return r, (dst.raw[:r] if r >= 0 else None)


def best():
    """The strongest codec available: the compiled reference if present, else the port."""
    return Ref() if Ref.available() else Port()
