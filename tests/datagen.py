"""Seeded test inputs shared by the CPU and GPU suites."""
import ctypes as C
import random
import zlib

import numpy as np


def logtext(nbytes: int, seed: int = 0x504C5A34, first_seg: int = 0) -> bytes:
    """The benchmark's synthetic log text (plz4_b200/csrc/logtext.h), generated on the host."""
    from plz4_b200 import _lib
    buf = np.empty(max(nbytes, 1), dtype=np.uint8)
    _lib.lib().plz4cu_gen_logtext_host(seed, first_seg, C.c_void_p(buf.ctypes.data), nbytes)
    return buf[:nbytes].tobytes()


def make(kind: str, n: int, seed: int = 1) -> bytes:
    rng = random.Random((zlib.crc32(kind.encode()) & 0xFFFF) * 1000003 + seed * 7919 + n)
    if kind == "random":
        return rng.randbytes(n)
    if kind == "zeros":
        return bytes(n)
    if kind == "ab":
        return bytes(rng.choice(b"ab") for _ in range(n))
    if kind == "words":
        words = [rng.randbytes(rng.randint(1, 12)) for _ in range(40)]
        out = bytearray()
        while len(out) < n:
            out += rng.choice(words)
        return bytes(out[:n])
    if kind == "log":
        return logtext(n, seed=0x504C5A34 + seed)
    if kind == "record1025":     # the repeated-record pattern of rd_test.go:1561-1573
        rec = rng.randbytes(1025)
        return (rec * (n // 1025 + 1))[:n]
    if kind == "runs":           # short periods: overlapping matches with offsets 1..40
        out = bytearray()
        while len(out) < n:
            p = rng.randbytes(rng.randint(1, 40))
            out += p * rng.randint(1, 30)
        return bytes(out[:n])
    raise ValueError(kind)


KINDS = ["random", "zeros", "ab", "words", "log", "record1025", "runs"]
