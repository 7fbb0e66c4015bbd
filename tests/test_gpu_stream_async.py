"""The batching flavour of NewWriter / NewReader (n_parallel != 0): pinned double-buffered staging, the engine and
content-hash threads, read-ahead batches.  Mirrors the reference's async stream tests (internal/test/wr_test.go
run with WithParallel(-1), rd_test.go:813-893, wr_test.go:852-1031) at sizes that cross several batches."""
import ctypes as C
import io
import random
import threading

import numpy as np
import pytest

from oracle import frame_oracle as F
from tests.test_gpu_stream import compress, decompress

pytestmark = pytest.mark.gpu

MiB = 1 << 20


def logtext(n, seed=7):
    from plz4_b200 import _lib
    a = np.empty(n, dtype=np.uint8)
    _lib.lib().plz4cu_gen_logtext_host(seed, 0, C.c_void_p(a.ctypes.data), n)
    return a.tobytes()


@pytest.fixture(scope="module")
def big():
    return logtext(40 * MiB + 12345)


def test_multi_batch_pipeline_roundtrip(gpu, port, big):
    """12 MiB batches (> the pageable staging limit) => pinned slabs rotate while the engine thread works."""
    marks = []
    f = compress(gpu, big, chunk=MiB + 3, block_size_idx=4, block_checksum=True, content_checksum=True,
                 pending_size=12 * MiB, progress=lambda s, d: marks.append((s, d)))
    nblk = (len(big) + 65535) // 65536
    assert [m[0] for m in marks] == [min(i * 65536, len(big)) for i in range(nblk)] + [len(big)]
    assert all(b[1] > a[1] for a, b in zip(marks, marks[1:]))
    assert marks[-1][1] == len(f) - 8                                   # EndMark + content checksum follow
    assert F.read_frames(f, port) == big                                 # independent decoder, checks both checksums
    # the frame is a pure function of the data: synchronous flavour, one write call, other batch sizes
    assert f == compress(gpu, big, block_size_idx=4, block_checksum=True, content_checksum=True, parallel=0)
    assert f == compress(gpu, big, block_size_idx=4, block_checksum=True, content_checksum=True)
    # read-ahead reader, odd read sizes, several batch sizes (slow start: 1, 4, 12 MiB ...)
    rng = random.Random(11)
    for pend in (0, 12 * MiB, 3 * MiB):
        r = gpu.NewReader(io.BytesIO(f), pending_size=pend)
        out = bytearray()
        while True:
            chunk = r.read(rng.choice([1, 4096, 65536, 70001, MiB, 5 * MiB]))
            if not chunk:
                break
            out += chunk
        r.close()
        assert bytes(out) == big
    # reader-side progress is in block order across batch boundaries
    rmarks = []
    assert decompress(gpu, f, pending_size=12 * MiB, progress=lambda s, d: rmarks.append((d, s))) == big
    assert rmarks[:nblk] == marks[:nblk]


def test_flush_is_a_barrier_mid_stream(gpu, port, big):
    dst = io.BytesIO()
    w = gpu.NewWriter(dst, block_size_idx=5, block_checksum=True, pending_size=16 * MiB)
    w.write(big[:20 * MiB + 5])
    w.flush()
    # everything written so far is decodable once the frame is terminated by hand (EndMark, no content checksum ...)
    part = dst.getvalue()
    assert len(part) > 7
    w.write(big[20 * MiB + 5:])
    w.close()
    f = dst.getvalue()
    assert f.startswith(part)
    assert F.read_frames(f, port) == big
    assert decompress(gpu, f) == big


def test_async_sink_failure_surfaces_once(gpu, big):
    """wr_test.go:852-1031 with WithParallel(-1): the error shows up on a later call, exactly once; Close is clean."""
    class Boom:
        def __init__(self, ok): self.ok, self.n = ok, 0
        def write(self, b):
            self.n += 1
            if self.n > self.ok:
                raise IOError("disk full")
            return len(b)
    for ok in (0, 1, 2):
        w = gpu.NewWriter(Boom(ok), block_size_idx=4, pending_size=12 * MiB)
        with pytest.raises(gpu.StreamError) as e:
            for i in range(0, len(big), MiB):
                w.write(big[i:i + MiB])
            w.flush()
        assert e.value.name in ("ErrWrite", "ErrHeaderWrite")
        w.close()                                                        # already reported


def test_corrupt_block_behind_the_read_ahead(gpu, big):
    f = bytearray(compress(gpu, big, block_size_idx=4, block_checksum=True, pending_size=12 * MiB))
    marks = []
    compress(gpu, big, block_size_idx=4, block_checksum=True, pending_size=12 * MiB, progress=lambda s, d: marks.append((s, d)))
    s_bad, d_bad = marks[400]                                            # a block in a later batch
    f[d_bad + 10] ^= 0x40
    r = gpu.NewReader(io.BytesIO(bytes(f)), pending_size=12 * MiB)
    out = bytearray()
    with pytest.raises(gpu.StreamError) as e:
        while True:
            chunk = r.read(MiB)
            if not chunk:
                break
            out += chunk
    assert e.value.name == "ErrBlockHash" and gpu.lz4_corrupted(e.value)
    assert bytes(out) == big[:s_bad]                                     # everything before the bad block was delivered
    r.close()


def test_pinned_source_is_compressed_in_place(gpu, port, big):
    from plz4_b200 import _lib, stream as S
    L = _lib.lib()
    n = 32 * MiB
    L.plz4cu_host_alloc.restype = C.c_void_p
    src = L.plz4cu_host_alloc(n)
    assert src
    C.memmove(src, big[:n], n)
    dst = np.empty(n + MiB, dtype=np.uint8)
    o, keep = S._opts(block_size_idx=4, block_checksum=True, content_checksum=True, pending_size=12 * MiB)
    sink = L.plz4cu_membuf_new(C.c_void_p(dst.ctypes.data), 0, dst.size)
    fn = lambda f: C.cast(f, C.c_void_p)
    w = L.plz4cu_writer_new(fn(L.plz4cu_membuf_write), sink, C.byref(o))
    assert L.plz4cu_writer_write(w, C.c_void_p(src), n) == n
    assert L.plz4cu_writer_close(w) == 0
    flen = L.plz4cu_membuf_len(sink)
    L.plz4cu_writer_free(w); L.plz4cu_membuf_free(sink)
    L.plz4cu_host_free(C.c_void_p(src))
    frame = dst[:flen].tobytes()
    assert F.read_frames(frame, port) == big[:n]
    assert frame == compress(gpu, big[:n], block_size_idx=4, block_checksum=True, content_checksum=True)


def test_concurrent_streams(gpu, big):
    """Several writers and readers at once: each leases its own engine pipe; results are independent."""
    want = compress(gpu, big[:24 * MiB], block_size_idx=4, block_checksum=True, pending_size=12 * MiB)
    errs = []

    def job(k):
        try:
            for _ in range(2):
                f = compress(gpu, big[:24 * MiB], chunk=(k + 1) * MiB, block_size_idx=4, block_checksum=True, pending_size=12 * MiB)
                assert f == want
                assert decompress(gpu, f, pending_size=(k + 2) * 4 * MiB) == big[:24 * MiB]
        except BaseException as e:           # noqa: BLE001 - reported by the main thread
            errs.append(repr(e))

    ts = [threading.Thread(target=job, args=(k,)) for k in range(6)]
    for t in ts: t.start()
    for t in ts: t.join()
    assert not errs, errs
    from plz4_b200 import _lib
    assert _lib.lib().plz4cu_host_outstanding() >= 0


class FailAfter(io.RawIOBase):
    """failReader (rd_test.go:1493-1505): hands out the bytes, then raises at the n-th read call."""
    def __init__(self, data, ok_calls):
        self.b, self.ok, self.n = io.BytesIO(data), ok_calls, 0
    def readable(self): return True
    def read(self, n=-1):
        self.n += 1
        if self.n > self.ok:
            raise IOError("cable pulled")
        return self.b.read(n)


def test_failing_source_is_an_io_error_not_corruption(gpu, big):
    """rd_test.go:959-1075, wr_test.go:852-1031: injected read failures at the n-th call, both flavours."""
    f = compress(gpu, big[:20 * MiB], block_size_idx=4, block_checksum=True, pending_size=12 * MiB)
    for parallel in (0, -1):
        for ok in (0, 1, 3, 40, 200):
            r = gpu.NewReader(FailAfter(f, ok), parallel=parallel, pending_size=12 * MiB)
            out = bytearray()
            with pytest.raises(gpu.StreamError) as e:
                while True:
                    chunk = r.read(MiB)
                    if not chunk:
                        break
                    out += chunk
            assert not gpu.lz4_corrupted(e.value), (parallel, ok, e.value.name)
            assert bytes(out) == big[:len(out)] and len(out) % 65536 == 0       # whole blocks before the failure
            with pytest.raises(gpu.StreamError):                                   # the error is sticky
                r.read(1)
            r.close()
    # Writer.ReadFrom with a failing source: the error surfaces, Close afterwards is clean (already reported)
    w = gpu.NewWriter(io.BytesIO(), block_size_idx=4, pending_size=12 * MiB)
    with pytest.raises(gpu.StreamError) as e:
        w.read_from(FailAfter(big, 5))
    assert e.value.name == "ErrBlockRead"
    w.close()


def test_bad_seeker_and_early_close(gpu, big):
    from plz4_b200 import _lib
    L = _lib.lib()
    marks = []
    f = compress(gpu, big, block_size_idx=4, block_checksum=True, pending_size=12 * MiB, progress=lambda s, d: marks.append((s, d)))

    class BadSeeker(io.BytesIO):                                                  # rd_test.go:1629-1640
        def seek(self, *a):
            raise IOError("no seeking today")
    with pytest.raises(gpu.StreamError) as e:
        decompress(gpu, BadSeeker(f).getvalue() and f, read_offset=3)             # misaligned offset
    assert e.value.name == "ErrReadOffset"
    r = gpu.NewReader(BadSeeker(f), read_offset=marks[5][1])
    with pytest.raises(gpu.StreamError) as e:
        r.read(10)
    assert e.value.name == "ErrReadOffset"
    r.close()

    # slow consumer gives up early (rd_test.go:1180-1250): Close with a batch still being decoded ahead must neither
    # hang nor touch the source afterwards, and every pinned slab goes back
    import gc
    del r, e
    gc.collect()
    before = L.plz4cu_host_outstanding()

    def give_up_early():
        src = io.BytesIO(f)
        r = gpu.NewReader(src, pending_size=12 * MiB)
        assert r.read(3 * MiB) == big[:3 * MiB]
        r.close()
        pos = src.tell()
        with pytest.raises(gpu.StreamError) as e:
            r.read(1)
        assert e.value.name == "ErrClosed"
        assert src.tell() == pos
        # a writer dropped without Close finishes its queued work and frees its staging
        w = gpu.NewWriter(io.BytesIO(), block_size_idx=4, pending_size=12 * MiB)
        w.write(big[:30 * MiB])

    give_up_early()
    gc.collect()
    assert L.plz4cu_host_outstanding() == before


def test_flush_storm_after_the_stream_went_threaded(gpu, port, big):
    """wr_test.go:238-346 on a stream that already runs its helper threads: every Flush is a barrier and makes a block."""
    dst = io.BytesIO()
    w = gpu.NewWriter(dst, block_size_idx=4, block_checksum=True, content_checksum=True, pending_size=12 * MiB)
    w.write(big[:30 * MiB])                                              # slabs rotate, engine / sink / hash threads run
    pos = 30 * MiB
    sizes = []
    for i in range(40):
        n = 1 + (i * 7919) % 3000
        w.write(big[pos:pos + n]); pos += n
        w.flush()
        sizes.append(len(dst.getvalue()))
        w.flush()                                                        # nothing pending: no empty block
        assert len(dst.getvalue()) == sizes[-1]
    w.write(big[pos:])
    w.close()
    assert all(b > a for a, b in zip(sizes, sizes[1:]))
    f = dst.getvalue()
    assert F.read_frames(f, port) == big
    assert decompress(gpu, f, pending_size=12 * MiB) == big
