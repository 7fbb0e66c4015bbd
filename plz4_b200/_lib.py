"""ctypes binding of libplz4cu.so (include/plz4cu.h).

This module is plumbing: it declares the C ABI to ctypes and nothing else.  There is no CPU
codec behind it — if the shared library is missing it is built, and if it cannot be built or
loaded the import fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import re

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(_HERE, "..", "include", "plz4cu.h")

ERR_CUDA, ERR_ARG, ERR_NOMEM, ERR_NODEVICE = -1, -2, -3, -4
E_BLOCKHASH = -0x7F000001
E_OVERFLOW = -0x7F000002
E_STALL = -0x7F000003
STORED_BIT = 0x80000000
INT32_MIN = -(1 << 31)

_vp, _u32, _u64, _i32, _int, _sz = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int32, C.c_int, C.c_size_t

# name -> (restype, argtypes); must list every PLZ4CU_API symbol of the header (tests check this)
SIGNATURES = {
    "plz4cu_device_count": (_int, []),
    "plz4cu_init": (_int, [_int]),
    "plz4cu_init_devices": (_int, [_int, C.POINTER(C.c_int)]),
    "plz4cu_registered_devices": (_int, []),
    "plz4cu_last_error": (C.c_char_p, []),
    "plz4cu_version": (C.c_char_p, []),
    "plz4cu_launch_count": (_u64, []),
    "plz4cu_compress_bound": (_sz, [_sz]),
    "plz4cu_host_alloc": (_vp, [_sz]),
    "plz4cu_host_free": (None, [_vp]),
    "plz4cu_host_trim": (None, []),
    "plz4cu_host_outstanding": (C.c_int64, []),
    "plz4cu_device_alloc": (_vp, [_sz]),
    "plz4cu_device_free": (None, [_vp]),
    "plz4cu_dict_create": (_vp, [_vp, _sz]),
    "plz4cu_dict_destroy": (None, [_vp]),
    "plz4cu_compress_batch_device": (_int, [_vp, _vp, _vp, _vp, _u32, _u32, _int, _int, _vp, _vp, _u32, _vp]),
    "plz4cu_decompress_batch_device": (_int, [_vp, _vp, _vp, _vp, _u32, _u32, _int, _int, _vp, _vp, _u64, _vp]),
    "plz4cu_pack_records_device": (_int, [_vp, _vp, _u32, _vp, _u32, _vp, _vp]),
    "plz4cu_frame_index_device": (_int, [_vp, _vp, _u64, _u32, _int, _vp, _u32, _vp, _vp]),
    "plz4cu_decompress_frame_device": (_int, [_vp, _vp, _u64, _vp, _vp, _u64, _vp, _vp, _u32, _vp]),
    "plz4cu_compress_frame_device": (_int, [_vp, _vp, _u64, _vp, _vp, _vp, _u64, _vp]),
    "plz4cu_gen_logtext_device": (_int, [_vp, _u32, _u64, _vp, _u64]),
    "plz4cu_gen_logtext_host": (_int, [_u32, _u64, _vp, _u64]),
    "plz4cu_compress_batch_host": (_int, [_vp, _vp, _vp, _u32, _u32, _int, _int, _vp, _vp, _u64, _vp]),
    "plz4cu_decompress_batch_host": (_int, [_vp, _u64, _vp, _vp, _u32, _u32, _int, _int, _vp, _vp, _u64, _vp]),
    "plz4cu_compress_fast": (_int, [_vp, _int, _vp, _int]),
    "plz4cu_compress_fast_dict": (_int, [_vp, _vp, _int, _vp, _int]),
    "plz4cu_decompress_safe": (_int, [_vp, _int, _vp, _int]),
    "plz4cu_decompress_safe_dict": (_int, [_vp, _vp, _int, _vp, _int]),
    "plz4cu_xxh32_batch_device": (_int, [_vp, _vp, _vp, _vp, _u32, _vp]),
    # frame streams
    "plz4cu_opts_default": (None, [_vp]),
    "plz4cu_err_corrupted": (_int, [_int]),
    "plz4cu_strerror": (C.c_char_p, [_int]),
    "plz4cu_writer_new": (_vp, [_vp, _vp, _vp]),
    "plz4cu_writer_write": (C.c_int64, [_vp, _vp, _sz]),
    "plz4cu_writer_read_from": (C.c_int64, [_vp, _vp, _vp]),
    "plz4cu_writer_flush": (_int, [_vp]),
    "plz4cu_writer_close": (_int, [_vp]),
    "plz4cu_writer_free": (None, [_vp]),
    "plz4cu_reader_new": (_vp, [_vp, _vp, _vp, _vp]),
    "plz4cu_reader_read": (C.c_int64, [_vp, _vp, _sz]),
    "plz4cu_reader_write_to": (C.c_int64, [_vp, _vp, _vp]),
    "plz4cu_reader_close": (_int, [_vp]),
    "plz4cu_reader_free": (None, [_vp]),
    "plz4cu_write_skip_frame_header": (_int, [_vp, _vp, C.c_uint8, _u32]),
    "plz4cu_xxh32_host": (_u32, [_vp, _sz]),
    "plz4cu_frame_header": (_int, [_vp, _vp]),
    "plz4cu_membuf_new": (_vp, [_vp, _sz, _sz]),
    "plz4cu_membuf_free": (None, [_vp]),
    "plz4cu_membuf_len": (_sz, [_vp]),
    "plz4cu_membuf_read": (C.c_int64, [_vp, _vp, _sz]),
    "plz4cu_membuf_write": (C.c_int64, [_vp, _vp, _sz]),
    "plz4cu_membuf_seek": (_int, [_vp, C.c_int64]),
}

WRITE_FN = C.CFUNCTYPE(C.c_int64, _vp, _vp, _sz)
READ_FN = C.CFUNCTYPE(C.c_int64, _vp, _vp, _sz)
SEEK_FN = C.CFUNCTYPE(_int, _vp, C.c_int64)
TASK_FN = C.CFUNCTYPE(None, _vp)
SUBMIT_FN = C.CFUNCTYPE(_int, _vp, TASK_FN, _vp)
PROGRESS_FN = C.CFUNCTYPE(None, _vp, C.c_int64, C.c_int64)
SKIP_FN = C.CFUNCTYPE(_int, _vp, C.c_uint8, _vp, _u32)
DICT_FN = C.CFUNCTYPE(_int, _vp, _u32, C.POINTER(_vp), C.POINTER(_sz))


class Opts(C.Structure):
    """plz4cu_opts_t (include/plz4cu.h)."""
    _fields_ = [
        ("level", C.c_int32), ("n_parallel", C.c_int32), ("pending_size", C.c_int32), ("block_size_idx", C.c_int32),
        ("block_checksum", C.c_int32), ("content_checksum", C.c_int32), ("block_linked", C.c_int32),
        ("has_content_size", C.c_int32), ("content_size", C.c_uint64),
        ("has_dict_id", C.c_int32), ("dict_id", C.c_uint32),
        ("dict", _vp), ("dict_len", _sz),
        ("read_offset", C.c_int64), ("content_size_check", C.c_int32), ("n_devices", C.c_int32),
        ("progress", PROGRESS_FN), ("progress_ctx", _vp),
        ("skip_cb", SKIP_FN), ("skip_ctx", _vp),
        ("dict_cb", DICT_FN), ("dict_ctx", _vp),
        ("submit", SUBMIT_FN), ("submit_ctx", _vp),
    ]


class FrameInfo(C.Structure):
    """plz4cu_frame_info_t (include/plz4cu.h)."""
    _fields_ = [
        ("block_size", C.c_uint32), ("header_len", C.c_uint32),
        ("block_checksum", C.c_int32), ("content_checksum", C.c_int32), ("has_content_size", C.c_int32), ("has_dict_id", C.c_int32),
        ("dict_id", C.c_uint32), ("content_hash", C.c_uint32),
        ("content_size", C.c_uint64), ("nblk", C.c_uint64), ("frame_len", C.c_uint64), ("out_bytes", C.c_uint64),
        ("contiguous", C.c_int32), ("reserved0", C.c_int32),
    ]


def header_symbols() -> list[str]:
    """Every function the public header declares."""
    with open(HEADER) as f:
        return sorted(set(re.findall(r"PLZ4CU_API[^;(]*?\b(plz4cu_\w+)\s*\(", f.read())))


_lib = None


def lib() -> C.CDLL:
    """Load (building if stale) libplz4cu.so and declare every signature."""
    global _lib
    if _lib is None:
        path = _build.build()
        L = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)            # AttributeError here == header/library mismatch: loud
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class Plz4cuError(RuntimeError):
    """Infrastructure failure inside the engine (CUDA error, bad argument, no device)."""


def check(rc: int, what: str = "") -> int:
    if rc < 0:
        msg = lib().plz4cu_last_error().decode(errors="replace")
        raise Plz4cuError(f"{what or 'plz4cu'} failed ({rc}): {msg}")
    return rc
