"""One stream over several GPUs in one process (BASELINE configs[4]): batches are cut into runs of
blocks, one run per device, the dictionary is replicated, and the frame is byte for byte the one-GPU writer's.  Also the
WithWorkerPool hook (plz4_opts.go:107): the stream's stages run on the caller's pool."""
import io
import threading

import pytest

from tests.datagen import make

pytestmark = pytest.mark.gpu


def _mixed(n_mib):
    parts = []
    for i in range(n_mib):
        kind = ["log", "log", "random", "zeros", "record1025", "log", "words", "log", "random", "log"][i % 10]
        parts.append(make(kind, 1 << 20, seed=i))
    return b"".join(parts)


def _write(gpu, data, **o):
    dst = io.BytesIO()
    w = gpu.NewWriter(dst, **o)
    for i in range(0, len(data), 5_000_011):
        w.write(data[i:i + 5_000_011])
    w.close()
    return dst.getvalue()


def _read(gpu, frame, **o):
    r = gpu.NewReader(io.BytesIO(frame), **o)
    out = r.read_all()
    r.close()
    return out


def test_one_stream_over_two_gpus_is_byte_identical(gpu):
    if gpu.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2; log in profiles/r02_multi_gpu.txt)")
    gpu.init_devices([0, 1])
    data = _mixed(96)
    d = make("log", 65536, seed=99)
    for opts in (dict(block_size_idx=5, block_checksum=True, content_checksum=True),
                 dict(block_size_idx=4, block_checksum=True, content_checksum=False, dictionary=d, dict_id=7),
                 dict(block_size_idx=7, block_checksum=False, content_checksum=True)):
        marks1, marks2 = [], []
        one = _write(gpu, data, n_devices=1, progress=lambda s, t: marks1.append((s, t)), **opts)
        two = _write(gpu, data, n_devices=2, progress=lambda s, t: marks2.append((s, t)), **opts)
        assert one == two and marks1 == marks2
        ro = {"dictionary": d} if "dictionary" in opts else {}
        assert _read(gpu, two, n_devices=2, **ro) == data
        assert _read(gpu, two, n_devices=-1, **ro) == data
        assert _read(gpu, one, n_devices=1, **ro) == data


def test_registered_devices_and_bad_lists(gpu):
    L = gpu._lib.lib()
    import ctypes as C
    n = gpu.device_count()
    gpu.init_devices(list(range(n)))
    assert L.plz4cu_registered_devices() == n
    bad = (C.c_int * 2)(0, 0)
    assert L.plz4cu_init_devices(2, bad) < 0                       # listed twice
    bad = (C.c_int * 1)(n)
    assert L.plz4cu_init_devices(1, bad) < 0                       # out of range
    # a stream asking for more devices than there are uses what is registered
    data = make("log", 3 << 20, seed=3)
    f = _write(gpu, data, n_devices=-1, block_size_idx=4)
    assert _read(gpu, f, n_devices=8) == data


class _Pool:
    """opts.WorkerPool (internal/pkg/opts/opts.go:43-45): Submit(task) runs it on a worker of the caller's choosing."""

    def __init__(self):
        self.tasks = 0
        self.threads = []

    def submit(self, fn):
        self.tasks += 1
        t = threading.Thread(target=fn, name=f"pool-{self.tasks}", daemon=True)
        self.threads.append(t)
        t.start()


def test_worker_pool_runs_the_stream_stages(gpu):
    data = _mixed(40)
    pool = _Pool()
    f = _write(gpu, data, worker_pool=pool, block_size_idx=4, block_checksum=True, content_checksum=True)
    assert pool.tasks >= 2                                          # engine + sink (+ content hash) were handed to the pool
    assert f == _write(gpu, data, block_size_idx=4, block_checksum=True, content_checksum=True)
    rpool = _Pool()
    assert _read(gpu, f, worker_pool=rpool) == data
    assert rpool.tasks >= 1
    for t in pool.threads + rpool.threads:
        t.join(timeout=10)
        assert not t.is_alive()                                     # the loops leave when the stream is freed
