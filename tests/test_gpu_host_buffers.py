"""plz4cu_*_batch_host on page-locked and on ordinary host memory: same bytes, same lengths, same return codes.  Ordinary
memory is staged through the engine's pinned slabs by several threads (engine.cu: stage_in / stage_out); page-locked memory
moves by DMA in place; PLZ4CU_STAGE=0 hands ordinary memory to cudaMemcpyAsync as round 1 did."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from tests.datagen import make

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
vp = lambda a: C.c_void_p(a.ctypes.data)


def _pinned(L, n):
    L.plz4cu_host_alloc.restype = C.c_void_p
    p = L.plz4cu_host_alloc(max(n, 1))
    assert p
    return p, np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(max(n, 1),))


@pytest.mark.parametrize("bsz,nblk", [(65536, 700), (4096, 5000), (1 << 20, 40)])
def test_batch_calls_agree_on_pinned_and_ordinary_memory(gpu, bsz, nblk):
    from plz4_b200 import _lib
    from plz4_b200._lib import check
    L = _lib.lib()
    kinds = ["log", "words", "runs", "random", "zeros"]
    blocks = [make(kinds[i % 5], bsz if i % 7 else max(1, bsz // 3), seed=i) for i in range(nblk)]
    lens = np.array([len(b) for b in blocks], dtype=np.uint32)
    off = np.concatenate([[0], np.cumsum(lens[:-1], dtype=np.uint64)]).astype(np.uint64)
    raw = b"".join(blocks)
    n = len(raw)
    cap = int(L.plz4cu_compress_bound(bsz))
    pcap = nblk * (cap + 8)
    results = []
    for pinned in (False, True):
        if pinned:
            ps, src = _pinned(L, n); pp, packed = _pinned(L, pcap); po, out = _pinned(L, nblk * (bsz + 16))
        else:
            src = np.empty(n, dtype=np.uint8); packed = np.empty(pcap, dtype=np.uint8); out = np.empty(nblk * (bsz + 16), dtype=np.uint8)
        src[:n] = np.frombuffer(raw, dtype=np.uint8)
        poff = np.zeros(nblk + 1, dtype=np.uint64)
        check(L.plz4cu_compress_batch_host(vp(src), vp(off), vp(lens), nblk, bsz, 1, 0, None, vp(packed), pcap, vp(poff)))
        total = int(poff[nblk])
        res = np.zeros(nblk, dtype=np.int32)
        # an output stride that is not the engine's own (bsz + 16): the rows are spread over the caller's buffer one by one
        check(L.plz4cu_decompress_batch_host(vp(packed), total, vp(poff), None, nblk, bsz, 1, 0, None, vp(out), bsz + 16, vp(res)))
        assert (res == lens.astype(np.int32)).all()
        rows = out[: nblk * (bsz + 16)].reshape(nblk, bsz + 16)
        for i in (0, 1, nblk // 2, nblk - 1):
            assert rows[i, : lens[i]].tobytes() == blocks[i], (pinned, i)
        back = b"".join(rows[i, : lens[i]].tobytes() for i in range(nblk))
        assert back == raw
        results.append((packed[:total].tobytes(), poff.copy()))
        if pinned:
            for p in (ps, pp, po):
                L.plz4cu_host_free(C.c_void_p(p))
    assert results[0][0] == results[1][0] and (results[0][1] == results[1][1]).all()


def test_ordinary_memory_without_the_staging():
    """PLZ4CU_STAGE=0: the same test with pageable buffers handed straight to cudaMemcpyAsync."""
    env = dict(os.environ, PLZ4CU_STAGE="0")
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "-k", "pinned_and_ordinary",
                        "tests/test_gpu_host_buffers.py"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
