"""Frame streams (NewWriter / NewReader) on the GPU engine, mirroring the reference's own stream tests
(internal/test/wr_test.go, rd_test.go, plz4_test.go) with the frame oracle as the independent checker."""
import io
import random

import pytest

from oracle import frame_oracle as F
from tests.datagen import make
from tests.test_golden import HELLO_FRAME, ONE_FRAME, ONE_FRAME_NOHASH, THE_WORKS_54, THE_WORKS_HDR

pytestmark = pytest.mark.gpu


def compress(gpu, data, chunk=None, **opts):
    dst = io.BytesIO()
    w = gpu.NewWriter(dst, **opts)
    if chunk is None:
        w.write(data)
    else:
        for i in range(0, len(data), chunk):
            w.write(data[i:i + chunk])
    w.close()
    return dst.getvalue()


def decompress(gpu, frame, **opts):
    r = gpu.NewReader(io.BytesIO(frame), **opts)
    try:
        return r.read_all()
    finally:
        r.close()


def test_example_hello_frame_byte_exact(gpu):
    # plz4_test.go:41-75 ExampleNewWriter: WithParallel(1), WithContentChecksum(false), Write, Flush, Close
    dst = io.BytesIO()
    w = gpu.NewWriter(dst, parallel=1, content_checksum=False)
    w.write(b"hello")
    w.flush()
    w.close()
    with pytest.raises(gpu.StreamError) as e:            # "double close is ok": it reports ErrClosed (async/writer.go:136-138)
        w.close()
    assert e.value.name == "ErrClosed"
    assert dst.getvalue() == HELLO_FRAME
    # plz4_test.go:9-38 ExampleNewReader
    out = io.BytesIO()
    r = gpu.NewReader(io.BytesIO(HELLO_FRAME), parallel=0)
    assert r.write_to(out) == 5 and out.getvalue() == b"hello"
    r.close()


def test_reference_golden_frames_decode(gpu):
    assert decompress(gpu, THE_WORKS_54) == b"testycode"          # rd_test.go:527-538
    assert decompress(gpu, ONE_FRAME) == b"testy\n"               # rd_test.go:714
    assert decompress(gpu, ONE_FRAME_NOHASH) == b"testy\n"
    assert decompress(gpu, THE_WORKS_HDR) == b"testy\n\n"         # header/read_test.go:15


def test_short_read_matrix_is_never_corrupted(gpu):
    """rd_test.go:521-706: clip the annotated 54-byte frame everywhere; a short read is not corruption."""
    expect = {19: "ErrBlockSizeRead", 21: "ErrBlockSizeRead", 23: "ErrBlockRead", 30: "ErrBlockRead", 33: "ErrBlockSizeRead",
              40: "ErrBlockRead", 46: "ErrBlockSizeRead", 50: "ErrContentHashRead", 52: "ErrContentHashRead", 3: "ErrHeaderRead", 12: "ErrHeaderRead"}
    for cut in range(1, len(THE_WORKS_54)):
        with pytest.raises(gpu.StreamError) as e:
            decompress(gpu, THE_WORKS_54[:cut])
        assert not gpu.lz4_corrupted(e.value), cut
        if cut in expect:
            assert e.value.name == expect[cut], (cut, e.value.name)
    # data decoded before the cut is still delivered (deferred error, rdr/rdr.go:66-75)
    r = gpu.NewReader(io.BytesIO(THE_WORKS_54[:40]))
    assert r.read(100) == b"testy"
    with pytest.raises(gpu.StreamError):
        r.read(100)
    r.close()


def test_corruption_taxonomy(gpu):
    def err(frame, **kw):
        with pytest.raises(gpu.StreamError) as e:
            decompress(gpu, bytes(frame), **kw)
        return e.value
    f = bytearray(THE_WORKS_54); f[24] ^= 1
    e = err(f); assert e.name == "ErrBlockHash" and gpu.lz4_corrupted(e)          # rd_test.go:926-954
    f = bytearray(THE_WORKS_54); f[-1] ^= 1
    e = err(f); assert e.name == "ErrContentHash" and gpu.lz4_corrupted(e)        # rd_test.go:710-810
    assert decompress(gpu, bytes(f), content_checksum=False) == b"testycode"      # check disabled by option
    f = bytearray(THE_WORKS_54); f[18] ^= 1
    e = err(f); assert e.name == "ErrHeaderHash" and gpu.lz4_corrupted(e)
    f = bytearray(THE_WORKS_54); f[0] = 0x05
    e = err(f); assert e.name == "ErrMagic" and gpu.lz4_corrupted(e)
    f = bytearray(HELLO_FRAME); f[4] |= 0x02
    e = err(f); assert e.name == "ErrReserveBitSet" and gpu.lz4_corrupted(e)      # rd_test.go:26-128
    f = bytearray(HELLO_FRAME); f[4] = (f[4] & 0x3F) | 0x80
    e = err(f); assert e.name == "ErrVersion" and not gpu.lz4_corrupted(e)
    f = bytearray(HELLO_FRAME); f[5] = 0x30
    e = err(f); assert e.name == "ErrBlockDescriptor" and gpu.lz4_corrupted(e)
    f = bytearray(ONE_FRAME_NOHASH); f[7:11] = (65537).to_bytes(4, "little")      # rd_test.go:896-923
    e = err(f); assert e.name == "ErrBlockSizeOverflow" and gpu.lz4_corrupted(e)
    f = bytearray(HELLO_FRAME); f[11] = 0xF0                                      # garbage LZ4 block
    e = err(f); assert e.name == "ErrDecompress" and gpu.lz4_corrupted(e)


def test_content_size_check(gpu):
    # rd_test.go:132-195
    sz_one = bytes.fromhex("04224d18684001000000000000002c0100008000") + b"\0\0\0\0"
    sz_zero_with_one = bytes.fromhex("04224d1868400000000000000000050100008000") + b"\0\0\0\0"
    assert decompress(gpu, sz_one) == b"\0"
    with pytest.raises(gpu.StreamError) as e:
        decompress(gpu, sz_zero_with_one)
    assert e.value.name == "ErrContentSize" and gpu.lz4_corrupted(e.value)
    assert decompress(gpu, sz_zero_with_one, content_size_check=False) == b"\0"
    data = make("log", 100000)
    f = compress(gpu, data, content_size=len(data), block_size_idx=4)
    assert decompress(gpu, f) == data
    assert F.read_header(f, 0, lambda b: __import__("xxhash").xxh32(b).intdigest())[1].content_size == len(data)


@pytest.mark.parametrize("bidx", [4, 5, 7])
@pytest.mark.parametrize("parallel", [0, 1, -1])
def test_option_matrix_roundtrip_and_interop(gpu, port, bidx, parallel):
    """wr_test.go:50-200: every option combination round-trips, and interoperates with the reference format."""
    bsz = F.BLOCK_SIZES[bidx]
    for bx in (False, True):
        for cx in (False, True):
            for n in (0, 1, bsz - 1, bsz, bsz + 1, 3 * bsz + 12345):
                data = make("log", n, seed=bidx)
                f = compress(gpu, data, block_size_idx=bidx, block_checksum=bx, content_checksum=cx, parallel=parallel)
                assert F.read_frames(f, port) == data, (bx, cx, n)              # the reference format reader accepts it
                assert decompress(gpu, f, parallel=parallel) == data
                ref_frame = F.write_frame(data, F.Opts(block_idx=bidx, block_checksum=bx, content_checksum=cx), port)
                assert decompress(gpu, ref_frame, parallel=parallel) == data    # and we accept reference-produced frames
                assert len(f) <= len(ref_frame) * 1.03 + 16


def test_progress_offsets_and_read_offset(gpu, port):
    """wr_test.go:202-232,1198-1235 + rd_test.go:1077-1176: marks from the writer are valid WithReadOffset starts."""
    bsz = 65536
    data = make("log", 10 * bsz + 777)
    marks = []
    f = compress(gpu, data, block_size_idx=4, block_checksum=True, progress=lambda s, d: marks.append((s, d)))
    assert [m[0] for m in marks] == [min(i * bsz, len(data)) for i in range(11)] + [len(data)]
    ref_marks = []
    F.write_frame(data, F.Opts(block_idx=4, block_checksum=True), port, progress=lambda s, d: ref_marks.append(s))
    assert [m[0] for m in marks] == ref_marks
    for s, d in marks[:-1]:
        r = gpu.NewReader(io.BytesIO(f), read_offset=d)
        assert r.read(bsz) == data[s:s + bsz]                       # validatePos reads one block
        r.close()
        assert decompress(gpu, f, read_offset=d) == data[s:]         # and to the end (content checksum is skipped)
    # reader-side progress agrees with the writer's (src/dst swapped roles)
    rmarks = []
    assert decompress(gpu, f, progress=lambda s, d: rmarks.append((d, s))) == data
    assert rmarks[: len(marks) - 1] == marks[:-1]
    # no seek available: offset is honoured by read-and-discard (rd_test.go:1255-1273)
    class NoSeek(io.RawIOBase):
        def __init__(self, b): self.b = io.BytesIO(b)
        def read(self, n=-1): return self.b.read(n)
        def readable(self): return True
    r = gpu.NewReader(NoSeek(f), read_offset=marks[3][1])
    assert r.read(1024) == data[3 * bsz: 3 * bsz + 1024]
    r.close()
    with pytest.raises(gpu.StreamError) as e:                        # rd_test.go:1277-1297
        decompress(gpu, f, read_offset=3)
    assert e.value.name == "ErrReadOffset"


def test_flush_makes_blocks_and_write_chunking(gpu, port):
    """wr_test.go:238-346: flushing creates a block each time; odd write sizes do not matter."""
    data = make("words", 5000)
    dst = io.BytesIO()
    w = gpu.NewWriter(dst, block_size_idx=4, content_checksum=True)
    for i in range(0, len(data), 100):
        w.write(data[i:i + 100])
        w.flush()
        w.flush()                                        # nothing pending: no empty block
    w.close()
    nblocks = []
    F.read_frames(dst.getvalue(), port, progress=lambda s, d: nblocks.append(d))
    assert len(nblocks) == 50 + 1
    assert decompress(gpu, dst.getvalue()) == data
    big = make("log", 1_000_000)
    for chunk in (1, 7, 65536, 65537, 300000):
        if chunk == 1:
            f = compress(gpu, big[:3000], chunk=1, block_size_idx=4)
            assert decompress(gpu, f) == big[:3000]
            continue
        f = compress(gpu, big, chunk=chunk, block_size_idx=4, block_checksum=True)
        assert f == compress(gpu, big, block_size_idx=4, block_checksum=True)      # output is a pure function of the data
        assert F.read_frames(f, port) == big
    dst = io.BytesIO()
    w = gpu.NewWriter(dst, block_size_idx=4, block_checksum=True)
    w.write(big[:1000])
    assert w.read_from(io.BytesIO(big[1000:])) == len(big) - 1000                # wr_test.go:662-723 interleave
    w.close()
    assert decompress(gpu, dst.getvalue()) == big


def test_concatenated_and_skip_frames(gpu):
    # rd_test.go:198-373, wr_test.go:727-847
    a, b = make("log", 70000), make("words", 1000)
    fa, fb = compress(gpu, a, block_size_idx=4), compress(gpu, b, block_size_idx=5, block_checksum=True)
    skip = io.BytesIO()
    assert gpu.write_skip_frame_header(skip, 0xF, 3) == 8
    skip.write(b"abc")
    empty_skip = bytes.fromhex("502a4d1800000000")
    seen = []
    out = decompress(gpu, empty_skip + fa + skip.getvalue() + fb + empty_skip, skip_callback=lambda nib, p: seen.append((nib, p)))
    assert out == a + b and seen == [(0, b""), (15, b"abc"), (0, b"")]
    assert decompress(gpu, fa + skip.getvalue() + fb) == a + b            # no callback: payload discarded
    with pytest.raises(gpu.StreamError) as e:
        gpu.write_skip_frame_header(io.BytesIO(), 16, 0)
    assert e.value.name == "ErrNibble"
    # a skip frame that claims 4 GiB and delivers ten bytes: ErrSkip, without the reader allocating what the header says
    # (header/skip.go:38-76 streams the payload), with and without a callback; a 200 KiB payload arrives whole
    liar = bytes.fromhex("532a4d18f0ffffff") + b"0123456789"
    for cb in (None, lambda nib, p: None):
        with pytest.raises(gpu.StreamError) as e:
            decompress(gpu, fa + liar, skip_callback=cb)
        assert e.value.name == "ErrSkip"
    big = bytes(range(256)) * 800
    hdr = io.BytesIO()
    gpu.write_skip_frame_header(hdr, 2, len(big))
    seen = []
    assert decompress(gpu, hdr.getvalue() + big + fb, skip_callback=lambda nib, p: seen.append((nib, p))) == b
    assert seen == [(2, big)]
    assert decompress(gpu, hdr.getvalue() + big + fb) == b


def test_dictionary_frames(gpu, port):
    # wr_test.go:416-625, rd_test.go:376-442,1373-1487
    from tests.datagen import logtext
    corpus = logtext(400000, seed=123)
    d, data = corpus[:65536], corpus[100000:300000]
    f = compress(gpu, data, block_size_idx=4, dictionary=d, dict_id=6789, block_checksum=True)
    plain = compress(gpu, data, block_size_idx=4, block_checksum=True)
    assert len(f) < len(plain)                                                 # the dictionary is applied
    assert F.read_frames(f, port, dictionary=d) == data                        # reference-format reader + dict
    assert decompress(gpu, f, dictionary=d) == data
    ids = []
    assert decompress(gpu, f, dict_callback=lambda i: (ids.append(i), d)[1]) == data and ids == [6789]
    ref = F.write_frame(data, F.Opts(block_idx=4, dictionary=d, dict_id=6789, content_checksum=True), port)
    assert decompress(gpu, ref, dictionary=d) == data
    with pytest.raises(gpu.StreamError):                                       # wrong / missing dictionary
        decompress(gpu, f, dictionary=corpus[70000:135536])


def test_unsupported_and_state_errors(gpu):
    w = gpu.NewWriter(io.BytesIO(), level=3)
    with pytest.raises(gpu.StreamError) as e:
        w.write(b"x")
    assert e.value.name == "ErrUnsupported"
    w.close()
    w = gpu.NewWriter(io.BytesIO(), block_linked=True)
    with pytest.raises(gpu.StreamError) as e:
        w.write(b"x")
    assert e.value.name == "ErrUnsupported"
    # wr_test.go:1087-1101: API calls after Close report ErrClosed
    w = gpu.NewWriter(io.BytesIO())
    w.write(b"abc")
    w.close()
    with pytest.raises(gpu.StreamError) as e:
        w.write(b"more")
    assert e.value.name == "ErrClosed"
    with pytest.raises(gpu.StreamError):
        w.flush()
    # failing io.Writer (wr_test.go:852-1031): error surfaces once, Close is then clean
    class Boom:
        def __init__(self): self.n = 0
        def write(self, b):
            self.n += 1
            if self.n >= 2:
                raise IOError("disk full")
            return len(b)
    w = gpu.NewWriter(Boom(), block_size_idx=4, parallel=0)
    with pytest.raises(gpu.StreamError) as e:
        w.write(make("log", 200000))
        w.flush()
    assert e.value.name == "ErrWrite"
    w.close()
    r = gpu.NewReader(io.BytesIO(HELLO_FRAME))
    assert r.read(10) == b"hello" and r.read(10) == b""          # io.EOF, never (0, nil) mid-stream
    r.close()
    with pytest.raises(gpu.StreamError) as e:
        r.read(1)
    assert e.value.name == "ErrClosed"


def test_random_read_sizes_large_stream(gpu, port):
    # rd_test.go:813-893 + a multi-batch stream (pending_size forces several engine calls)
    data = make("log", 9 * (1 << 20) + 31, seed=4) + make("random", 200000) + bytes(300000)
    f = compress(gpu, data, chunk=1 << 20, block_size_idx=4, block_checksum=True, pending_size=2 << 20)
    assert F.read_frames(f, port) == data
    rng = random.Random(5)
    r = gpu.NewReader(io.BytesIO(f), pending_size=1 << 20)
    out = bytearray()
    while True:
        chunk = r.read(rng.choice([1, 2, 100, 4096, 65536, 70001, 1 << 20]))
        if not chunk:
            break
        out += chunk
    r.close()
    assert bytes(out) == data


def test_reader_leaves_the_source_where_the_frame_ends(gpu):
    """The reader fills its record area in large reads when the source can seek and gives back what it read past the
    frame, so bytes that follow an LZ4 frame in the same source are still there for the caller (the reference reads
    block by block and never runs ahead, blk/frame.go:54-112)."""
    data = make("log", 3_000_000, seed=31)
    for opts in (dict(block_size_idx=4, block_checksum=True), dict(block_size_idx=5, content_checksum=True)):
        frame = compress(gpu, data, **opts)
        trailer = b"TRAILING BYTES THAT ARE NOT LZ4" * 1000
        src = io.BytesIO(frame + trailer)
        r = gpu.NewReader(src)
        got = bytearray()
        while len(got) < len(data):
            chunk = r.read(len(data) - len(got))
            assert chunk
            got += chunk
        assert bytes(got) == data
        # io.Reader semantics: the frame's end is only seen by the read that finds the EndMark; ask for it
        with pytest.raises(gpu.StreamError):
            r.read(1)                                       # what follows is not a frame header: ErrMagic
        r.close()
    # closing inside a body also puts the source back
    frame = compress(gpu, data, block_size_idx=4)
    src = io.BytesIO(frame)
    r = gpu.NewReader(src)
    first = r.read(100_000)
    assert first == data[:100_000]
    r.close()
    assert src.tell() <= len(frame)
