"""frame_oracle.py — CPU ORACLE for the LZ4 *frame* layer (test infrastructure, NOT product code).

A small pure-Python restatement of the Go host code that surrounds the block engine, used to
produce "reference-produced frames" and to check frames the product writes.  Block bytes come
from the pinned C oracle (oracle/lz4_port.c == the reference's liblz4, see tests/test_oracle_vs_ref.py).

Restates (file:line relative to /root/reference):
  write_header        internal/pkg/header/write.go:23-73, descriptor/flags.go:3-42, descriptor/block.go:9-29
  read_header         internal/pkg/header/read.go:26-119, header/skip.go:38-76
  write_frame         internal/pkg/sync/writer.go:62-122,136-190,265-290 (block slicing, trailer, content hash)
                      internal/pkg/blk/blk.go:69-109 (record), trailer/trailer.go:10-19
  read_frame(s)       internal/pkg/rdr/rdr.go:242-296 (header mode, ReadOffset), blk/frame.go:54-139,
                      rdr/rdr.go:91-101 (content size check), async/reader.go:236-241 (content hash check)
Parity status: PINNED by the reference's own golden frames (tests/test_golden.py: G1-G6).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

MAGIC = bytes([0x04, 0x22, 0x4D, 0x18])
SKIP_MAGIC = 0x184D2A50
BLOCK_SIZES = {4: 64 << 10, 5: 256 << 10, 6: 1 << 20, 7: 4 << 20}


class FrameError(Exception):
    def __init__(self, kind: str, corrupted: bool = False):
        super().__init__(("lz4 corrupted: " if corrupted else "") + kind)
        self.kind = kind
        self.corrupted = corrupted


@dataclass
class Opts:
    block_idx: int = 7                 # plz4_opts.go:238-255 defaults
    block_checksum: bool = False
    content_checksum: bool = True
    content_size: int | None = None
    dict_id: int | None = None
    linked: bool = False
    dictionary: bytes | None = None


@dataclass
class Header:
    size: int = 0
    flags: int = 0
    bd: int = 0
    content_size: int | None = None
    dict_id: int | None = None

    @property
    def block_checksum(self): return bool(self.flags & 0x10)
    @property
    def content_checksum(self): return bool(self.flags & 0x04)
    @property
    def independent(self): return bool(self.flags & 0x20)
    @property
    def block_size(self): return BLOCK_SIZES[(self.bd >> 4) & 7]


def write_header(o: Opts, xxh32) -> bytes:
    flags = 1 << 6
    if not o.linked:
        flags |= 0x20
    if o.block_checksum:
        flags |= 0x10
    if o.content_checksum:
        flags |= 0x04
    body = b""
    if o.content_size is not None:
        flags |= 0x08
        body += struct.pack("<Q", o.content_size)
    if o.dict_id is not None:
        flags |= 0x01
        body += struct.pack("<I", o.dict_id)
    desc = bytes([flags, (o.block_idx & 7) << 4]) + body
    return MAGIC + desc + bytes([(xxh32(desc) >> 8) & 0xFF])


def read_header(buf: bytes, pos: int, xxh32):
    """-> ("frame", Header, new_pos) | ("skip", nibble, payload, new_pos) | ("eof",)"""
    if pos == len(buf):
        return ("eof",)
    if len(buf) - pos < 7:
        raise FrameError("lz4 fail read header")
    if buf[pos:pos + 4] != MAGIC:
        m = struct.unpack_from("<I", buf, pos)[0]
        if m >> 4 != SKIP_MAGIC >> 4:
            raise FrameError("lz4 bad magic", True)
        if len(buf) - pos < 8:
            raise FrameError("lz4 fail read header")
        sz = struct.unpack_from("<I", buf, pos + 4)[0]
        if len(buf) - pos - 8 < sz:
            raise FrameError("lz4 fail skip")
        return ("skip", m & 0xF, buf[pos + 8: pos + 8 + sz], pos + 8 + sz)
    h = Header(flags=buf[pos + 4], bd=buf[pos + 5])
    if (h.flags >> 6) & 3 != 1:
        raise FrameError("lz4 unsupported version")
    if h.flags & 0x02:
        raise FrameError("lz4 reserved bit set", True)
    if ((h.bd >> 4) & 7) < 4 or (h.bd & 0x80) or (h.bd & 0x0F):
        raise FrameError("lz4 invalid BD byte", True)
    n = 7
    if h.flags & 0x08:
        if len(buf) - pos < n + 8:
            raise FrameError("lz4 fail read header")
        h.content_size = struct.unpack_from("<Q", buf, pos + 6)[0]
        n += 8
    if h.flags & 0x01:
        if len(buf) - pos < n + 4:
            raise FrameError("lz4 fail read header")
        h.dict_id = struct.unpack_from("<I", buf, pos + n - 1)[0]
        n += 4
    if ((xxh32(buf[pos + 4: pos + n - 1]) >> 8) & 0xFF) != buf[pos + n - 1]:
        raise FrameError("lz4 header hash mismatch", True)
    h.size = n
    return ("frame", h, pos + n)


def write_frame(data: bytes, o: Opts, port, progress=None) -> bytes:
    """Whole-buffer equivalent of NewWriter(...).Write(data); Close()."""
    bsz = BLOCK_SIZES[o.block_idx]
    dict_ = port.dict_create(o.dictionary) if o.dictionary is not None else None
    out = bytearray(write_header(o, port.xxh32))
    src_mark = 0
    for i in range(0, len(data), bsz):
        blk = data[i:i + bsz]
        if progress:
            progress(src_mark, len(out))
        out += port.block_record(blk, bsz, o.block_checksum, dict_)
        src_mark += len(blk)
    if progress:
        progress(src_mark, len(out))
    out += b"\x00\x00\x00\x00"
    if o.content_checksum:
        out += struct.pack("<I", port.xxh32(data))
    return bytes(out)


def read_frames(buf: bytes, port, dictionary: bytes | None = None, read_offset: int = 0,
                content_size_check: bool = True, progress=None, skip_cb=None) -> bytes:
    """Whole-buffer equivalent of NewReader(...).WriteTo(): concatenated + skippable frames."""
    out = bytearray()
    pos = 0
    first = True
    dict_ = port.dict_create(dictionary) if dictionary else None
    while True:
        r = read_header(buf, pos, port.xxh32)
        if r[0] == "eof":
            return bytes(out)
        if r[0] == "skip":
            if skip_cb:
                skip_cb(r[1], r[2])
            pos = r[3]
            continue
        h, pos = r[1], r[2]
        hdr_start = pos - h.size
        check_content_hash = h.content_checksum
        check_size = content_size_check and h.content_size is not None
        if first and read_offset:
            # rdr/rdr.go:261-285
            if not h.independent:
                raise FrameError("lz4 read offset unsupported in block linked mode")
            if read_offset < h.size:
                raise FrameError("lz4 bad read offset")
            pos = hdr_start + read_offset
            check_content_hash = False
            check_size = False
        first = False
        bsz = h.block_size
        frame_out = bytearray()
        while True:
            if len(buf) - pos < 4:
                raise FrameError("lz4 fail read block size")
            word = struct.unpack_from("<I", buf, pos)[0]
            if progress and word != 0:
                progress(pos - hdr_start, len(frame_out))
            pos += 4
            if word == 0:
                if h.content_checksum:
                    if len(buf) - pos < 4:
                        raise FrameError("lz4 fail read content hash")
                    want = struct.unpack_from("<I", buf, pos)[0]
                    pos += 4
                    if check_content_hash and want != port.xxh32(bytes(frame_out)):
                        raise FrameError("lz4 content hash mismatch", True)
                break
            n = word & 0x7FFFFFFF
            if n > bsz:
                raise FrameError("lz4 block size overflow", True)
            need = n + (4 if h.block_checksum else 0)
            if len(buf) - pos < need:
                raise FrameError("lz4 fail read block")
            payload = buf[pos:pos + n]
            if h.block_checksum and struct.unpack_from("<I", buf, pos + n)[0] != port.xxh32(payload):
                raise FrameError("lz4 block hash mismatch", True)
            pos += need
            if word & 0x80000000:
                frame_out += payload
            else:
                rc, dec = (dict_.decompress(payload, bsz) if dict_ else port.decompress(payload, bsz))
                if rc < 0:
                    raise FrameError("lz4 fail decompress", True)
                frame_out += dec
        if progress:
            progress(pos - hdr_start - (8 if h.content_checksum else 4), len(frame_out))
        if check_size and h.content_size != len(frame_out):
            raise FrameError("lz4 content size mismatch", True)
        out += frame_out
