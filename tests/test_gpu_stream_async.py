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
