"""Host-side mirror of plz4's block-level API on top of the C ABI (include/plz4cu.h).

Names, argument meaning and error behaviour follow the reference:
  compress_block / decompress_block / compress_block_bound   plz4_block.go:78-172
  Lz4Error hierarchy + lz4_corrupted()                         plz4_err.go:11-45, zerr/zerr.go:11-41
  compress_batch / decompress_batch                            the batched form of blk.CompressToBlk
                                                               (blk/blk.go:69-109) and BlkT.Decompress
                                                               (blk/blk.go:50-61) + frame.go:79-127 checks
Everything computes on the GPU through libplz4cu.so; numpy is only used to hold host buffers.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Sequence

import numpy as np

from . import _lib
from ._lib import E_BLOCKHASH, E_OVERFLOW, E_STALL, INT32_MIN, STORED_BIT, Plz4cuError, check

BLOCK_IDX_64KB, BLOCK_IDX_256KB, BLOCK_IDX_1MB, BLOCK_IDX_4MB = 4, 5, 6, 7      # descriptor/index.go:5-14
BLOCK_SIZES = {4: 64 << 10, 5: 256 << 10, 6: 1 << 20, 7: 4 << 20}
MAX_TRIES, INIT_MULTIPLE = 3, 4                                                 # plz4_block.go:8-11


# ---------------------------------------------------------------- errors (zerr/zerr.go)

class Lz4Error(Exception):
    """Base of the reference's sentinel errors; `kinds` plays the role of errors.Is."""
    kinds: tuple[str, ...] = ()

    def __init__(self, msg: str, kinds: Iterable[str] = ()):
        super().__init__(msg)
        self.kinds = tuple(kinds)

    def is_(self, kind: str) -> bool:
        return kind in self.kinds


ERR_CORRUPTED = "lz4 corrupted"
ERR_DECOMPRESS = "lz4 fail decompress"
ERR_COMPRESS = "lz4 fail compress"
ERR_BLOCK_HASH = "lz4 block hash mismatch"
ERR_BLOCK_SIZE_OVERFLOW = "lz4 block size overflow"


def lz4_corrupted(err: BaseException) -> bool:
    """plz4.Lz4Corrupted (plz4_err.go:43-45)."""
    return isinstance(err, Lz4Error) and err.is_(ERR_CORRUPTED)


def block_error(code: int) -> Lz4Error:
    """Map a per-block negative out_len onto the reference's error joins."""
    if code == E_BLOCKHASH:      # blk/frame.go:122-124
        return Lz4Error(f"{ERR_CORRUPTED}: {ERR_BLOCK_HASH}", (ERR_CORRUPTED, ERR_BLOCK_HASH))
    if code == E_OVERFLOW:       # blk/frame.go:79-81
        return Lz4Error(f"{ERR_CORRUPTED}: {ERR_BLOCK_SIZE_OVERFLOW}", (ERR_CORRUPTED, ERR_BLOCK_SIZE_OVERFLOW))
    if code == E_STALL:          # an engine fault, not a verdict on the data
        raise Plz4cuError("a decode team stalled (PLZ4CU_E_STALL)")
    # compress/decompress.go:33-36 + clz4.go:55-57
    return Lz4Error(f"{ERR_CORRUPTED}\n{ERR_DECOMPRESS}\nlz4 fail decompress: code {code}", (ERR_CORRUPTED, ERR_DECOMPRESS))


# ---------------------------------------------------------------- helpers

def _np(b) -> np.ndarray:
    if isinstance(b, np.ndarray):
        return np.ascontiguousarray(b).view(np.uint8).reshape(-1)
    return np.frombuffer(b, dtype=np.uint8) if len(b) else np.zeros(0, dtype=np.uint8)


def _ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data) if a.size else C.c_void_p(0)


class Dict:
    """compress.DictT + clz4.DictCtx: last 64 KiB, resident on the device (compress/dict.go:5-56)."""

    def __init__(self, data):
        a = _np(data)
        self._keep = a
        self.handle = _lib.lib().plz4cu_dict_create(_ptr(a), a.size)
        if not self.handle:
            raise Plz4cuError(_lib.lib().plz4cu_last_error().decode())
        self.data = bytes(a[-65536:].tobytes())

    def close(self):
        if self.handle:
            _lib.lib().plz4cu_dict_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def init(device: int = 0) -> None:
    check(_lib.lib().plz4cu_init(device), "plz4cu_init")


def init_devices(devices: Sequence[int]) -> None:
    """Register the GPUs one stream may spread its batches over (NewWriter / NewReader with n_devices > 1 or -1)."""
    arr = (C.c_int * len(devices))(*[int(d) for d in devices])
    check(_lib.lib().plz4cu_init_devices(len(devices), arr), "plz4cu_init_devices")


def device_count() -> int:
    return int(_lib.lib().plz4cu_device_count())


def compress_block_bound(n: int) -> int:
    """plz4.CompressBlockBound (plz4_block.go:78-80)."""
    return int(_lib.lib().plz4cu_compress_bound(n))


# ---------------------------------------------------------------- batched host API

def compress_batch(src, offsets: Sequence[int], lengths: Sequence[int], dst_cap: int, *,
                   block_checksum: bool = False, raw_blocks: bool = False, dict: Dict | None = None):
    """Compress independent blocks of one host buffer.  Returns (packed uint8 array, offsets[nblk+1])."""
    L = _lib.lib()
    a = _np(src)
    off = np.ascontiguousarray(offsets, dtype=np.uint64)
    ln = np.ascontiguousarray(lengths, dtype=np.uint32)
    nblk = int(off.size)
    slot = max(int(dst_cap), int(ln.max()) if nblk else 0) + 8
    packed = np.empty(max(1, nblk * slot), dtype=np.uint8)
    poff = np.zeros(nblk + 1, dtype=np.uint64)
    check(L.plz4cu_compress_batch_host(_ptr(a), _ptr(off), _ptr(ln), nblk, dst_cap, int(block_checksum),
                                       int(raw_blocks), dict.handle if dict else None,
                                       _ptr(packed), packed.size, _ptr(poff)), "compress_batch_host")
    return packed[: int(poff[nblk])], poff


def decompress_batch(recs, rec_off: Sequence[int], dst_cap: int, *, verify_checksum: bool = False,
                     raw_len: Sequence[int] | None = None, dict: Dict | None = None):
    """Decode independent block records (or raw blocks when raw_len is given) of one host buffer.
    Returns (out uint8 [nblk, dst_cap], out_len int32 [nblk])."""
    L = _lib.lib()
    a = _np(recs)
    off = np.ascontiguousarray(rec_off, dtype=np.uint64)
    nblk = int(off.size)
    rl = np.ascontiguousarray(raw_len, dtype=np.uint32) if raw_len is not None else None
    out = np.zeros((max(nblk, 1), max(int(dst_cap), 1)), dtype=np.uint8)
    res = np.zeros(max(nblk, 1), dtype=np.int32)
    check(L.plz4cu_decompress_batch_host(_ptr(a), a.size, _ptr(off), _ptr(rl) if rl is not None else None, nblk,
                                         dst_cap, int(verify_checksum), int(rl is not None),
                                         dict.handle if dict else None, _ptr(out), out.shape[1], _ptr(res)),
          "decompress_batch_host")
    return out[:nblk], res[:nblk]


# ---------------------------------------------------------------- device-resident frames

def decompress_frame_device(frame, *, dict: Dict | None = None, stream=None):
    """Decode one LZ4 frame held in a CUDA uint8 tensor without moving it to the host: header, device-side block
    walk (blk/frame.go:54-112 done in parallel), batched decode with block-checksum verification.
    Returns (out, info): `out` is a CUDA uint8 tensor of info.out_bytes decoded bytes, `info` the FrameInfo
    (block size, block count, bytes the frame occupied, content checksum value: reported, not verified).
    Raises stream.StreamError with the reference's error taxonomy on a malformed frame."""
    import torch
    from .stream import StreamError
    L = _lib.lib()
    assert frame.is_cuda and frame.dtype == torch.uint8 and frame.is_contiguous()
    info = _lib.FrameInfo()
    st = None if stream is None else C.c_void_p(stream)
    fp = C.c_void_p(frame.data_ptr())
    rc = L.plz4cu_decompress_frame_device(st, fp, frame.numel(), None, None, 0, None, None, 0, C.byref(info))   # sizing call
    if rc < 0 and rc != _lib.ERR_ARG:
        if rc > -100:
            check(rc, "decompress_frame_device")
        raise StreamError(rc)
    nblk, bsz = int(info.nblk), int(info.block_size)
    if nblk == 0:
        return torch.empty(0, dtype=torch.uint8, device=frame.device), info
    dst = torch.empty(nblk * bsz + 16, dtype=torch.uint8, device=frame.device)
    rec_off = torch.empty(nblk, dtype=torch.int64, device=frame.device)
    out_len = torch.empty(nblk, dtype=torch.int32, device=frame.device)
    rc = L.plz4cu_decompress_frame_device(st, fp, frame.numel(), dict.handle if dict else None, C.c_void_p(dst.data_ptr()),
                                          nblk * bsz, C.c_void_p(rec_off.data_ptr()), C.c_void_p(out_len.data_ptr()), nblk,
                                          C.byref(info))
    if rc < 0:
        if rc > -100:
            check(rc, "decompress_frame_device")
        raise StreamError(rc)
    if info.contiguous:
        return dst[: int(info.out_bytes)], info
    # blocks shorter than the block size in mid-stream (Flush): gather the pieces
    lens = out_len.to(torch.int64)
    parts = [dst[b * bsz: b * bsz + int(n)] for b, n in enumerate(lens.tolist())]
    return torch.cat(parts), info


def compress_frame_device(src, *, dict: Dict | None = None, stream=None, **options):
    """n bytes of a CUDA uint8 tensor -> one complete LZ4 frame in a CUDA uint8 tensor, byte for byte what this library's
    NewWriter writes for the same bytes and options (block_size_idx, block_checksum, content_size, dict_id; no content
    checksum).  Any LZ4 frame reader decodes it; the bytes are not those plz4 / liblz4 would write (the parse differs)."""
    import torch
    from .stream import StreamError, _opts
    L = _lib.lib()
    assert src.is_cuda and src.dtype == torch.uint8 and src.is_contiguous()
    options.setdefault("content_checksum", False)
    o, keep = _opts(**options)
    bsz = 1 << (8 + 2 * int(o.block_size_idx))
    n = src.numel()
    cap = 19 + ((n + bsz - 1) // bsz) * (bsz + 8) + 4
    frame = torch.empty(cap, dtype=torch.uint8, device=src.device)
    flen = C.c_uint64()
    rc = L.plz4cu_compress_frame_device(None if stream is None else C.c_void_p(stream), C.c_void_p(src.data_ptr()), n, C.byref(o),
                                        dict.handle if dict else None, C.c_void_p(frame.data_ptr()), cap, C.byref(flen))
    if rc < 0:
        if rc > -100:
            check(rc, "compress_frame_device")
        raise StreamError(rc)
    return frame[: flen.value]


# ---------------------------------------------------------------- raw block API (plz4_block.go)

def compress_block(src, *, dst_cap: int | None = None, dict: Dict | None = None) -> bytes:
    """plz4.CompressBlock (plz4_block.go:96-119).  `dst_cap` plays WithBlockDst's len(dst)."""
    L = _lib.lib()
    a = _np(src)
    cap = compress_block_bound(a.size) if dst_cap is None else int(dst_cap)
    dst = np.empty(max(cap, 1), dtype=np.uint8)
    if dict is not None:
        r = L.plz4cu_compress_fast_dict(dict.handle, _ptr(a), a.size, _ptr(dst), cap)
    else:
        r = L.plz4cu_compress_fast(_ptr(a), a.size, _ptr(dst), cap)
    if r == INT32_MIN:
        raise Plz4cuError(L.plz4cu_last_error().decode())
    if r == 0:      # compress/indie.go:69-71
        raise Lz4Error(f"{ERR_COMPRESS}\nlz4 fail compress; insufficient destination buffer", (ERR_COMPRESS,))
    return dst[:r].tobytes()


def _decompress_once(src: np.ndarray, cap: int, dict: Dict | None) -> tuple[int, np.ndarray]:
    L = _lib.lib()
    dst = np.empty(max(cap, 1), dtype=np.uint8)
    if dict is not None and len(dict.data):
        r = L.plz4cu_decompress_safe_dict(dict.handle, _ptr(src), src.size, _ptr(dst), cap)
    else:
        r = L.plz4cu_decompress_safe(_ptr(src), src.size, _ptr(dst), cap)
    if r == INT32_MIN:
        raise Plz4cuError(L.plz4cu_last_error().decode())
    return r, dst


def decompress_block(src, *, dst_cap: int | None = None, dict: Dict | None = None) -> bytes:
    """plz4.DecompressBlock (plz4_block.go:125-172) incl. the 4x/8x/16x grow-and-retry."""
    a = _np(src)
    if dst_cap is not None:
        r, dst = _decompress_once(a, int(dst_cap), dict)
        if r < 0:
            raise block_error(r)
        return dst[:r].tobytes()
    n_try, size = 1, a.size * INIT_MULTIPLE
    while True:
        r, dst = _decompress_once(a, size, dict)
        if r >= 0:
            return dst[:r].tobytes()
        if n_try < MAX_TRIES:
            n_try += 1
            size *= 2
        else:
            raise block_error(r)
