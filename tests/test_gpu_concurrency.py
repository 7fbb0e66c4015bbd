"""Threading and leak discipline of the boundary (SURVEY.md §8b threading row, §3.5 item 8)."""
import io
import threading

import numpy as np
import pytest

from tests.datagen import make

pytestmark = pytest.mark.gpu


def test_shared_engine_is_safe_from_many_threads(gpu, codec):
    """The reference shares ONE Decompressor between all reader goroutines (async/reader.go:71-74,202) and gives every
    writer goroutine its own Compressor: the engine's entry points must tolerate concurrent callers."""
    blocks = [make(k, n, seed=t) for t in range(6) for k, n in (("log", 65536), ("words", 30000), ("runs", 4096))]
    comp = [codec.compress(b) for b in blocks]
    errors = []

    def worker(tid):
        try:
            for it in range(8):
                i = (tid * 5 + it) % len(blocks)
                c = gpu.compress_block(blocks[i])
                assert codec.decompress(c, len(blocks[i])) == (len(blocks[i]), blocks[i])
                assert gpu.decompress_block(comp[i], dst_cap=len(blocks[i])) == blocks[i]
                off = np.cumsum([0] + [len(x) for x in comp[:6]])[:-1]
                out, res = gpu.decompress_batch(b"".join(comp[:6]), off, 65536, raw_len=[len(x) for x in comp[:6]])
                for j in range(6):
                    assert res[j] == len(blocks[j]) and out[j, : res[j]].tobytes() == blocks[j]
        except Exception as e:                                       # noqa: BLE001
            errors.append((tid, repr(e)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_no_pinned_slab_leaks_on_any_path(gpu):
    """testBorrowed (wr_test.go:29-33): every borrowed block goes back on every path, errors included."""
    L = gpu._lib.lib()
    base = L.plz4cu_host_outstanding()
    big = make("log", 24 << 20)

    def cycle(fail=False):
        class Sink:
            def __init__(self): self.n = 0
            def write(self, b):
                self.n += 1
                if fail and self.n > 1:
                    raise IOError("boom")
                return len(b)
        w = gpu.NewWriter(Sink(), block_size_idx=4, block_checksum=True)
        try:
            w.write(big)
            w.close()
        except gpu.StreamError:
            try:
                w.close()
            except gpu.StreamError:
                pass
        del w

    cycle()
    cycle(fail=True)
    dst = io.BytesIO()
    w = gpu.NewWriter(dst, block_size_idx=4)
    w.write(big)
    w.close()
    del w
    for cut in (None, len(dst.getvalue()) // 2):
        r = gpu.NewReader(io.BytesIO(dst.getvalue()[:cut]))
        try:
            r.read_all()
        except gpu.StreamError:
            pass
        r.close()
        del r
    import gc
    gc.collect()
    assert L.plz4cu_host_outstanding() == base
    L.plz4cu_host_trim()
    assert L.plz4cu_host_outstanding() == base
