"""NewWriter / NewReader — the Python face of the frame streams in include/plz4cu.h.

Same options, defaults and error behaviour as plz4 (plz4_writer.go:40-53, plz4_reader.go:28-33,
plz4_opts.go:70-255); the work happens in libplz4cu.so (host_stream.cu + the GPU engine).
"""
from __future__ import annotations

import ctypes as C
import io
from typing import Callable

from . import _lib
from .api import Lz4Error

Z_NAMES = {
    -101: "ErrClosed", -102: "ErrHeaderHash", -103: "ErrBlockHash", -104: "ErrContentHash", -105: "ErrHeaderRead",
    -106: "ErrHeaderWrite", -107: "ErrMagic", -108: "ErrVersion", -109: "ErrBlockSizeRead", -110: "ErrBlockRead",
    -111: "ErrBlockSizeOverflow", -112: "ErrDecompress", -113: "ErrReserveBitSet", -114: "ErrBlockDescriptor",
    -115: "ErrContentHashRead", -116: "ErrContentSize", -117: "ErrReadOffset", -118: "ErrReadOffsetLinked",
    -119: "ErrSkip", -120: "ErrNibble", -121: "ErrUnsupported", -122: "ErrWrite", -123: "ErrEngine",
}


class StreamError(Lz4Error):
    def __init__(self, code: int):
        L = _lib.lib()
        msg = L.plz4cu_strerror(code).decode()
        if code == -123:
            msg += ": " + L.plz4cu_last_error().decode(errors="replace")
        kinds = [Z_NAMES.get(code, str(code))]
        if L.plz4cu_err_corrupted(code):
            kinds.append("lz4 corrupted")
        super().__init__(msg, kinds)
        self.code = code
        self.name = Z_NAMES.get(code, str(code))


def _opts(*, level=1, parallel=1, pending_size=0, block_size_idx=7, block_checksum=False, content_checksum=True,
          block_linked=False, content_size=None, dict_id=None, dictionary=None, read_offset=0,
          content_size_check=True, progress=None, skip_callback=None, dict_callback=None, n_devices=0, worker_pool=None):
    """Build a plz4cu_opts_t from With* style keyword options; returns (struct, keepalive list)."""
    o = _lib.Opts()
    _lib.lib().plz4cu_opts_default(C.byref(o))
    keep = []
    o.level, o.n_parallel, o.pending_size = level, parallel, pending_size
    o.block_size_idx = block_size_idx
    o.block_checksum, o.content_checksum, o.block_linked = int(block_checksum), int(content_checksum), int(block_linked)
    if content_size is not None:
        o.has_content_size, o.content_size = 1, content_size
    if dict_id is not None:
        o.has_dict_id, o.dict_id = 1, dict_id
    if dictionary:
        buf = C.create_string_buffer(bytes(dictionary), len(dictionary))
        keep.append(buf)
        o.dict, o.dict_len = C.cast(buf, C.c_void_p), len(dictionary)
    o.read_offset = read_offset
    o.content_size_check = int(content_size_check)
    o.n_devices = int(n_devices)                # WithParallel across GPUs: devices registered by init_devices() (-1: all)
    if worker_pool is not None:                 # WithWorkerPool (plz4_opts.go:107): an object with submit(callable)
        def _submit(_ctx, task, arg):
            try:
                worker_pool.submit(lambda: task(arg))
                return 0
            except Exception:
                return -1
        cb = _lib.SUBMIT_FN(_submit)
        keep.append(cb)
        o.submit = cb
    if progress:
        cb = _lib.PROGRESS_FN(lambda _ctx, s, d: progress(s, d))
        keep.append(cb)
        o.progress = cb
    if skip_callback:
        def _skip(_ctx, nibble, payload, sz):
            try:
                skip_callback(nibble, C.string_at(payload, sz) if sz else b"")
                return 0
            except Exception:
                return -1
        cb = _lib.SKIP_FN(_skip)
        keep.append(cb)
        o.skip_cb = cb
    if dict_callback:
        def _dict(_ctx, did, pp, plen):
            try:
                d = dict_callback(did)
            except Exception:
                return -1
            if d:
                buf = C.create_string_buffer(bytes(d), len(d))
                keep.append(buf)
                pp[0] = C.cast(buf, C.c_void_p).value
                plen[0] = len(d)
            return 0
        cb = _lib.DICT_FN(_dict)
        keep.append(cb)
        o.dict_cb = cb
    return o, keep


class Writer:
    """plz4.Writer: write() / read_from() / flush() / close()."""

    def __init__(self, dst, **options):
        self._L = _lib.lib()
        self._dst = dst
        self._o, self._keep = _opts(**options)

        def _wr(_ctx, data, n):
            try:
                r = dst.write(C.string_at(data, n))
                return n if r is None else r
            except Exception:
                return -1
        self._wr = _lib.WRITE_FN(_wr)
        self._h = self._L.plz4cu_writer_new(self._wr, None, C.byref(self._o))

    def write(self, data) -> int:
        b = bytes(data)
        r = self._L.plz4cu_writer_write(self._h, b, len(b))
        if r < 0:
            raise StreamError(int(r))
        return int(r)

    def read_from(self, src) -> int:
        def _rd(_ctx, buf, n):
            try:
                d = src.read(n)
                C.memmove(buf, d, len(d))
                return len(d)
            except Exception:
                return -1
        cb = _lib.READ_FN(_rd)
        r = self._L.plz4cu_writer_read_from(self._h, cb, None)
        if r < 0:
            raise StreamError(int(r))
        return int(r)

    def flush(self) -> None:
        r = self._L.plz4cu_writer_flush(self._h)
        if r < 0:
            raise StreamError(r)

    def close(self) -> None:
        r = self._L.plz4cu_writer_close(self._h)
        if r < 0:
            raise StreamError(r)

    def __del__(self):
        try:
            if self._h:
                self._L.plz4cu_writer_free(self._h)
                self._h = None
        except Exception:
            pass


class Reader:
    """plz4.Reader: read() / write_to() / close()."""

    def __init__(self, src, **options):
        self._L = _lib.lib()
        self._src = src
        self._o, self._keep = _opts(**options)

        def _rd(_ctx, buf, n):
            try:
                d = src.read(n)
                if d:
                    C.memmove(buf, d, len(d))
                return len(d)
            except Exception:
                return -1
        self._rd = _lib.READ_FN(_rd)
        seek = None
        if hasattr(src, "seek") and getattr(src, "seekable", lambda: False)():
            def _seek(_ctx, delta):
                try:
                    src.seek(delta, io.SEEK_CUR)
                    return 0
                except Exception:
                    return -1
            seek = _lib.SEEK_FN(_seek)
        self._seek = seek
        self._h = self._L.plz4cu_reader_new(self._rd, seek if seek else C.cast(None, _lib.SEEK_FN), None, C.byref(self._o))

    def read(self, n: int) -> bytes:
        """Up to n bytes; b'' at end of stream (io.EOF)."""
        buf = C.create_string_buffer(max(n, 1))
        r = self._L.plz4cu_reader_read(self._h, buf, n)
        if r < 0:
            raise StreamError(int(r))
        return buf.raw[:r]

    def read_all(self) -> bytes:
        out = bytearray()
        while True:
            d = self.read(1 << 20)
            if not d:
                return bytes(out)
            out += d

    def write_to(self, dst) -> int:
        def _wr(_ctx, data, n):
            try:
                r = dst.write(C.string_at(data, n))
                return n if r is None else r
            except Exception:
                return -1
        cb = _lib.WRITE_FN(_wr)
        r = self._L.plz4cu_reader_write_to(self._h, cb, None)
        if r < 0:
            raise StreamError(int(r))
        return int(r)

    def close(self) -> None:
        r = self._L.plz4cu_reader_close(self._h)
        if r < 0:
            raise StreamError(r)

    def __del__(self):
        try:
            if self._h:
                self._L.plz4cu_reader_free(self._h)
                self._h = None
        except Exception:
            pass


def NewWriter(dst, **options) -> Writer:
    return Writer(dst, **options)


def NewReader(src, **options) -> Reader:
    return Reader(src, **options)


def write_skip_frame_header(dst, nibble: int, sz: int) -> int:
    """plz4.WriteSkipFrameHeader."""
    if nibble > 0xF:
        raise StreamError(-120)
    cb = _lib.WRITE_FN(lambda _c, data, n: (dst.write(C.string_at(data, n)) or n))
    r = _lib.lib().plz4cu_write_skip_frame_header(cb, None, nibble, sz)
    if r < 0:
        raise StreamError(r)
    return r
