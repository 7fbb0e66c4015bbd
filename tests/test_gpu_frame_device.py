"""Device-resident frames: the frame walk of blk/frame.go:54-112 done on the GPU (parallel candidate chains for frames
with many blocks, plain walk for few), checked against the host reader and the frame oracle on the same frames."""
import ctypes as C
import io

import numpy as np
import pytest
import torch

from oracle import frame_oracle as F
from tests.datagen import make
from tests.test_gpu_stream import compress, decompress
from tests.test_golden import HELLO_FRAME, THE_WORKS_54

pytestmark = pytest.mark.gpu
MiB = 1 << 20


def logtext(n, seed=7):
    from plz4_b200 import _lib
    a = np.empty(n, dtype=np.uint8)
    _lib.lib().plz4cu_gen_logtext_host(seed, 0, C.c_void_p(a.ctypes.data), n)
    return a.tobytes()


def dev(b):
    return torch.frombuffer(bytearray(b), dtype=torch.uint8).cuda()


def host_offsets(frame, port):
    """Record offsets (body-relative) by the plain host walk of the frame oracle's progress marks."""
    marks = []
    F.read_frames(frame, port, progress=lambda s, d: marks.append(d))
    return marks


@pytest.fixture(scope="module")
def mixed():
    # log text (compressible), random bytes (stored blocks full of accidental size words), zeros (tiny records)
    rng = np.random.default_rng(21)
    return logtext(96 * MiB) + rng.integers(0, 256, 40 * MiB, dtype=np.uint8).tobytes() + bytes(8 * MiB) + logtext(5 * MiB + 321, seed=9)


def test_golden_frames_from_device(gpu):
    out, info = gpu.decompress_frame_device(dev(HELLO_FRAME))
    assert bytes(out.cpu().numpy()) == b"hello" and info.nblk == 1 and info.frame_len == len(HELLO_FRAME) and info.block_size == 4 * MiB
    out, info = gpu.decompress_frame_device(dev(THE_WORKS_54 + b"trailing garbage"))
    assert bytes(out.cpu().numpy()) == b"testycode" and info.nblk == 2 and info.frame_len == len(THE_WORKS_54)
    assert info.block_checksum and info.content_checksum and info.content_hash == int.from_bytes(THE_WORKS_54[-4:], "little")
    assert info.has_dict_id and info.has_content_size and info.content_size == 9


@pytest.mark.parametrize("bidx,bx", [(4, True), (4, False), (5, True)])
def test_parallel_walk_matches_the_serial_walk(gpu, port, mixed, bidx, bx):
    """> 1024 blocks => the candidate-chain path; offsets must equal the host walk's, bytes the original."""
    from plz4_b200 import _lib
    L = _lib.lib()
    frame = compress(gpu, mixed, block_size_idx=bidx, block_checksum=bx, content_checksum=False)
    bsz = 1 << (8 + 2 * bidx)
    nblk = (len(mixed) + bsz - 1) // bsz
    if bidx == 4:
        assert (len(frame) - 7) // bsz > 1024                            # really takes the parallel path
    d = dev(frame)
    out, info = gpu.decompress_frame_device(d)
    assert info.nblk == nblk and info.frame_len == len(frame) and info.out_bytes == len(mixed)
    assert bytes(out.cpu().numpy()) == mixed
    # the index itself, against offsets found by walking the frame on the host
    want = []
    p = 7
    while int.from_bytes(frame[p:p + 4], "little") != 0:
        want.append(p - 7)
        p += 4 + (int.from_bytes(frame[p:p + 4], "little") & 0x7FFFFFFF) + (4 if bx else 0)
    rec_off = torch.zeros(nblk, dtype=torch.int64, device="cuda")
    n, end = C.c_uint64(), C.c_uint64()
    rc = L.plz4cu_frame_index_device(None, C.c_void_p(d.data_ptr() + 7), len(frame) - 7, bsz, int(bx), C.c_void_p(rec_off.data_ptr()),
                                     nblk, C.byref(n), C.byref(end))
    assert rc == 0 and n.value == nblk and end.value == p - 7 + 4
    assert rec_off.cpu().tolist() == want


def test_broken_frames_report_like_the_host_reader(gpu, mixed):
    frame = bytearray(compress(gpu, mixed[:90 * MiB], block_size_idx=4, block_checksum=True, content_checksum=True))
    marks = []
    compress(gpu, mixed[:90 * MiB], block_size_idx=4, block_checksum=True, content_checksum=True, progress=lambda s, d: marks.append(d))

    def both(fr):
        with pytest.raises(gpu.StreamError) as e1:
            decompress(gpu, bytes(fr))
        with pytest.raises(gpu.StreamError) as e2:
            gpu.decompress_frame_device(dev(bytes(fr)))
        return e1.value.name, e2.value.name

    f = bytearray(frame); f[marks[700] + 2] = 0x7F                        # size word far above the block size
    assert both(f) == ("ErrBlockSizeOverflow", "ErrBlockSizeOverflow")
    f = bytearray(frame); f[marks[900] + 9] ^= 0x10                       # payload bit flip: block checksum
    assert both(f) == ("ErrBlockHash", "ErrBlockHash")
    f = bytearray(frame); f[marks[300]] ^= 0x01                           # size off by one: the chain breaks
    n1, n2 = both(f)
    assert n2 in ("ErrBlockHash", "ErrBlockSizeOverflow", "ErrBlockRead", "ErrBlockSizeRead", "ErrDecompress") and n1 == n2
    assert both(frame[: marks[1200] + 100]) == ("ErrBlockRead", "ErrBlockRead")       # cut inside a record
    assert both(frame[: marks[1200]]) == ("ErrBlockSizeRead", "ErrBlockSizeRead")     # cut at a boundary
    f = bytearray(frame); f[5] = 0x30
    assert both(f) == ("ErrBlockDescriptor", "ErrBlockDescriptor")
    f = bytearray(frame); f[6] ^= 1
    assert both(f) == ("ErrHeaderHash", "ErrHeaderHash")


def test_flushed_short_blocks_and_big_blocks(gpu, port):
    # mid-stream short blocks (Flush) make the output non-contiguous in its slots; 4 MiB blocks use the plain walk
    data = make("log", 300_000, seed=3)
    dst = io.BytesIO()
    w = gpu.NewWriter(dst, block_size_idx=4, block_checksum=True)
    for i in range(0, len(data), 50_000):
        w.write(data[i:i + 50_000]); w.flush()
    w.close()
    out, info = gpu.decompress_frame_device(dev(dst.getvalue()))
    assert bytes(out.cpu().numpy()) == data and not info.contiguous and info.nblk == 6
    big = logtext(20 * MiB + 5)
    frame = F.write_frame(big, F.Opts(block_idx=7, block_checksum=True, content_checksum=True), port)   # reference-side writer
    out, info = gpu.decompress_frame_device(dev(frame))
    assert bytes(out.cpu().numpy()) == big and info.nblk == 6 and info.frame_len == len(frame)


def test_compress_frame_device_equals_the_host_writer(gpu, port, mixed):
    """The device-side writer emits byte for byte the frame NewWriter produces (same blocks, same order, same header)."""
    data = mixed[90 * MiB: 130 * MiB + 777]
    d = dev(data)
    for opts in (dict(block_size_idx=4, block_checksum=True), dict(block_size_idx=5), dict(block_size_idx=7, block_checksum=True, content_size=len(data), dict_id=77)):
        frame = gpu.compress_frame_device(d, **opts)
        fb = bytes(frame.cpu().numpy())
        assert fb == compress(gpu, data, content_checksum=False, **opts)
        assert F.read_frames(fb, port) == data                            # the independent decoder accepts it
        out, info = gpu.decompress_frame_device(frame)                    # and device -> device closes the loop
        assert torch.equal(out, d) and info.frame_len == len(fb)
    # empty input: header + EndMark (wr_test.go: zero-length streams)
    e = gpu.compress_frame_device(torch.empty(0, dtype=torch.uint8, device="cuda"), block_size_idx=4)
    assert bytes(e.cpu().numpy()) == compress(gpu, b"", block_size_idx=4, content_checksum=False)
    with pytest.raises(gpu.StreamError) as ex:                            # serial content checksum: host-side job
        gpu.compress_frame_device(d, content_checksum=True)
    assert ex.value.name == "ErrUnsupported"
    dct = gpu.Dict(mixed[:65536])
    frame = gpu.compress_frame_device(d[: 3 * MiB], dict=dct, block_size_idx=4, block_checksum=True, dict_id=5)
    out, info = gpu.decompress_frame_device(frame, dict=dct)
    assert torch.equal(out, d[: 3 * MiB]) and info.dict_id == 5
    assert F.read_frames(bytes(frame.cpu().numpy()), port, dictionary=mixed[:65536]) == data[: 3 * MiB]


def test_frame_index_edges(gpu, mixed):
    from plz4_b200 import _lib
    L = _lib.lib()
    frame = compress(gpu, mixed[:2 * MiB + 5], block_size_idx=4, block_checksum=False, content_checksum=False)
    d = dev(frame)
    nblk = 33
    rec_off = torch.zeros(64, dtype=torch.int64, device="cuda")
    n, end = C.c_uint64(), C.c_uint64()
    call = lambda ln, cap: L.plz4cu_frame_index_device(None, C.c_void_p(d.data_ptr() + 7), ln, 65536, 0, C.c_void_p(rec_off.data_ptr()),
                                                       cap, C.byref(n), C.byref(end))
    assert call(len(frame) - 7, 64) == 0 and n.value == nblk and end.value == len(frame) - 7
    assert call(len(frame) - 7, 3) == _lib.ERR_ARG and n.value == nblk          # too little room: the count still comes back
    assert call(3, 64) == -109 and n.value == 0                                  # ErrBlockSizeRead: not even a size word
    assert call(len(frame) - 7 - 4, 64) == -109 and n.value == nblk              # EndMark cut off
    assert call(len(frame) - 7 - 6, 64) == -110 and n.value == nblk - 1          # last record cut short: ErrBlockRead
    # whole-frame call on truncated headers
    for cut in (0, 3, 6):
        with pytest.raises(gpu.StreamError) as e:
            gpu.decompress_frame_device(dev(frame[:cut] + b"\\0" * (1 if cut == 0 else 0)))
        assert e.value.name in ("ErrHeaderRead", "ErrMagic")
