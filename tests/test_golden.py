"""The reference's own golden vectors against the oracle."""
import bz2
import hashlib
import os

import pytest

from oracle import frame_oracle as F

HELLO_FRAME = bytes.fromhex("04224d18607073060000005068656c6c6f00000000")               # plz4_test.go:12,74  (G1)
THE_WORKS_54 = bytes([                                                                    # rd_test.go:527-538   (G2)
    0x04, 0x22, 0x4d, 0x18,
    0x7d, 0x70, 0x09, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x0b, 0x00, 0x00, 0x00, 0x0c,
    0x06, 0x00, 0x00, 0x00, 0x50, 0x74, 0x65, 0x73, 0x74, 0x79, 0xcf, 0x22, 0x82, 0x16,
    0x05, 0x00, 0x00, 0x00, 0x40, 0x63, 0x6f, 0x64, 0x65, 0xc5, 0x63, 0x71, 0xe5,
    0x00, 0x00, 0x00, 0x00, 0x4a, 0x73, 0x1c, 0xae])
ONE_FRAME = bytes.fromhex("04224d186440a706000080746573 74790a000000005dc73f2a".replace(" ", ""))   # rd_test.go:714 (G3)
ONE_FRAME_NOHASH = bytes.fromhex("04224d18604082060000807465 7374790a00000000".replace(" ", ""))      # rd_test.go:715
THE_WORKS_HDR = bytes([0x04, 0x22, 0x4d, 0x18, 0x7d, 0x40, 0x07, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0x0f,   # header/read_test.go:15 (G4)
                       0x07, 0x00, 0x00, 0x80, 0x74, 0x65, 0x73, 0x74, 0x79, 0x0a, 0x0a, 0xb3, 0x89, 0x63, 0xba,
                       0x00, 0x00, 0x00, 0x00, 0xb3, 0x89, 0x63, 0xba])

HEADERS = {   # header/write_test.go:25-83 (G5): opts -> bytes after the magic
    "bsz_4M": (dict(block_idx=7), "607073"),
    "bsz_1M": (dict(block_idx=6), "606051"),
    "bsz_256KB": (dict(block_idx=5), "6050fb"),
    "bsz_64KB": (dict(block_idx=4), "604082"),
    "linked": (dict(block_idx=7, linked=True), "4070df"),
    "block_checksum": (dict(block_idx=7, block_checksum=True), "707072"),
    "content_checksum": (dict(block_idx=7, content_checksum=True), "6470b9"),
    "content_size": (dict(block_idx=7, content_size=11), "68700b0000000000000038"),
    "dict_id": (dict(block_idx=7, dict_id=6789), "6170851a0000af"),
    "dict_id+content_size": (dict(block_idx=7, dict_id=6789, content_size=11), "69700b00000000000000851a0000e2"),
}


def test_g1_hello_frame_decode_and_byte_exact_encode(port):
    assert F.read_frames(HELLO_FRAME, port) == b"hello"
    assert F.write_frame(b"hello", F.Opts(content_checksum=False), port) == HELLO_FRAME
    assert port.compress(b"hello") == bytes.fromhex("5068656c6c6f")      # 6 > 5 bytes yet kept compressed


def test_g2_annotated_frame_block_and_content_checksums(port):
    assert F.read_frames(THE_WORKS_54, port) == b"testycode"
    assert port.xxh32(bytes.fromhex("507465737479")) == int.from_bytes(bytes.fromhex("cf228216"), "little")
    assert port.xxh32(bytes.fromhex("40636f6465")) == int.from_bytes(bytes.fromhex("c56371e5"), "little")
    assert port.xxh32(b"testycode") == int.from_bytes(bytes.fromhex("4a731cae"), "little")
    kind, h, pos = F.read_header(THE_WORKS_54, 0, port.xxh32)
    assert (h.size, h.content_size, h.dict_id, h.block_size) == (19, 9, 11, 4 << 20)
    # every strict prefix is a short read, never "corrupted" (rd_test.go:636-641)
    for cut in range(1, len(THE_WORKS_54)):
        with pytest.raises(F.FrameError) as e:
            F.read_frames(THE_WORKS_54[:cut], port)
        assert not e.value.corrupted, cut


def test_g3_stored_block_frames_and_content_crc(port):
    assert F.read_frames(ONE_FRAME, port) == b"testy\n"
    assert F.read_frames(ONE_FRAME_NOHASH, port) == b"testy\n"
    assert hashlib.sha256(b"testy\n").hexdigest() == "4e64edc52754ee847f3f043382f70d8cc4f83e38113d3555bdee20442d0d5f50"   # rd_test.go:717
    bad = bytearray(ONE_FRAME); bad[-1] ^= 1
    with pytest.raises(F.FrameError) as e:
        F.read_frames(bytes(bad), port)
    assert e.value.corrupted and e.value.kind == "lz4 content hash mismatch"
    # the golden frame stores "testy\n" raw (another writer made it); liblz4 keeps it compressed (7 B):
    # different bytes, same content
    assert F.read_frames(F.write_frame(b"testy\n", F.Opts(block_idx=4), port), port) == b"testy\n"


def test_g4_the_works_header(port):
    kind, h, pos = F.read_header(THE_WORKS_HDR, 0, port.xxh32)
    assert kind == "frame" and pos == 19 and h.content_size == 7 and h.dict_id == 0
    assert h.block_checksum and h.content_checksum and h.independent and h.block_size == 65536
    assert F.read_frames(THE_WORKS_HDR, port) == b"testy\n\n"


@pytest.mark.parametrize("name", sorted(HEADERS))
def test_g5_header_bytes(port, name):
    kw, hexbytes = HEADERS[name]
    kw.setdefault("content_checksum", False)
    hdr = F.write_header(F.Opts(**kw), port.xxh32)
    assert hdr == F.MAGIC + bytes.fromhex(hexbytes)
    kind, h, pos = F.read_header(hdr + b"\0\0\0\0", 0, port.xxh32)
    assert pos == len(hdr)


def test_g5_content_size_headers(port):
    # header/read_test.go:172-209
    ok = {0: "68400000000000000000 05", 1: "68400100000000000000 2c", 0xFFFFFFFF: "6840ffffffff00000000 5e",
          2**64 - 2: "6840feffffffffffffff 86", 2**64 - 1: "6840ffffffffffffffff a7"}
    for sz, hx in ok.items():
        raw = F.MAGIC + bytes.fromhex(hx.replace(" ", ""))
        assert F.read_header(raw, 0, port.xxh32)[1].content_size == sz
        assert F.write_header(F.Opts(block_idx=4, content_checksum=False, content_size=sz), port.xxh32) == raw
    with pytest.raises(F.FrameError) as e:
        F.read_header(F.MAGIC + bytes.fromhex("6840828000000000000004"), 0, port.xxh32)
    assert e.value.kind == "lz4 header hash mismatch"


def test_g6_skip_frames(port):
    skip_empty = bytes.fromhex("502a4d1800000000")                     # rd_test.go:203-205
    skip_one = bytes.fromhex("5f2a4d1801000000f7")
    skip_bad = bytes.fromhex("602a4d1801000000f7")
    sz_zero = bytes.fromhex("04224d18684000000000000000000500000000")
    seen = []
    out = F.read_frames(skip_empty + skip_one + sz_zero + HELLO_FRAME, port, skip_cb=lambda nib, p: seen.append((nib, p)))
    assert out == b"hello" and seen == [(0, b""), (15, b"\xf7")]
    with pytest.raises(F.FrameError) as e:
        F.read_frames(skip_bad, port)
    assert e.value.corrupted and e.value.kind == "lz4 bad magic"


def test_g7_malformed_raw_blocks(port, codec):
    # block_test.go:315-319: {} , "not-a-valid-lz4-block", 64 x 0xFF -> liblz4 returns -1, -10, -51
    for blk, code in [(b"", -1), (b"not-a-valid-lz4-block", -10), (b"\xff" * 64, -51)]:
        for c in (port, codec):
            assert c.decompress(blk, len(blk) * 4)[0] == code


def test_g8_empty_input_is_one_zero_byte(port, codec):
    for c in (port, codec):                      # block_test.go:22-52
        assert c.compress(b"") == b"\x00"
        assert c.decompress(b"\x00", 0) == (0, b"")
        assert c.decompress(b"\x00", 16) == (0, b"")


def test_g10_dict_sample_if_present(port):
    p = "/root/reference/internal/test/samples/dict.bin.bz2"     # only on the build box; not needed on the GPU box
    if not os.path.exists(p):
        pytest.skip("reference checkout not present")
    d = bz2.decompress(open(p, "rb").read())
    assert len(d) == 65536
    assert hashlib.sha256(d).hexdigest().startswith("fb0f084f")
    dc = port.dict_create(d)
    msg = d[1000:3000] + b"tail"
    c = dc.compress(msg)
    assert len(c) < 64 and dc.decompress(c, len(msg)) == (len(msg), msg)


def test_frame_oracle_roundtrip_matrix(port):
    from tests.datagen import make
    for bi in (4, 5):
        for bx in (False, True):
            for cx in (False, True):
                for n in (0, 1, 65535, 65536, 65537, 300000):
                    data = make("log", n)
                    marks = []
                    f = F.write_frame(data, F.Opts(block_idx=bi, block_checksum=bx, content_checksum=cx), port,
                                      progress=lambda s, d: marks.append((s, d)))
                    assert F.read_frames(f, port) == data
                    assert F.read_frames(f + f, port) == data + data          # concatenated frames
                    for s, d in marks[1:-1]:                                  # WithReadOffset random access
                        assert F.read_frames(f, port, read_offset=d) == data[s:]
