"""The other BASELINE.json configs as parity cases (scaled to test size), plus size-independent
properties at a larger scale (round trip, checksum of checksums) on device-resident data."""
import ctypes as C
import io
import random

import numpy as np
import pytest

from oracle import frame_oracle as F
from tests.datagen import logtext, make

pytestmark = pytest.mark.gpu


def test_config1_cli_equivalent_4mib_blocks_both_checksums(gpu, port):
    """configs[0]: compress+decompress of log text, independent 4 MiB blocks, level 1, block+content checksums."""
    data = logtext(20 * (1 << 20) + 4321, seed=11)
    dst = io.BytesIO()
    w = gpu.NewWriter(dst, block_size_idx=7, block_checksum=True, content_checksum=True, parallel=-1)
    assert w.read_from(io.BytesIO(data)) == len(data)
    w.close()
    frame = dst.getvalue()
    assert F.read_frames(frame, port) == data                        # reference-format reader, all checksums verified
    out = io.BytesIO()
    r = gpu.NewReader(io.BytesIO(frame), parallel=-1)
    assert r.write_to(out) == len(data) and out.getvalue() == data
    r.close()
    ref = F.write_frame(data, F.Opts(block_idx=7, block_checksum=True, content_checksum=True), port)
    assert len(frame) <= 1.03 * len(ref)


def test_config3_random_access_into_reference_frames(gpu, port):
    """configs[2]: decompress-only of reference-produced frames (4 MiB blocks, block checksums) from
    WithReadOffset starts taken from the writer's progress marks."""
    bsz = 4 << 20
    data = logtext(6 * bsz + 12345, seed=12)
    marks = []
    frame = F.write_frame(data, F.Opts(block_idx=7, block_checksum=True, content_checksum=False), port,
                          progress=lambda s, d: marks.append((s, d)))
    rng = random.Random(3)
    for s, d in rng.sample(marks[:-1], 4):
        r = gpu.NewReader(io.BytesIO(frame), read_offset=d)
        assert r.read_all() == data[s:]
        r.close()


def test_config4_small_messages_with_dictionary_batch(gpu, port, codec):
    """configs[3]: batched 4 KiB payloads with a 64 KiB dictionary through the raw block API (one engine call)."""
    corpus = logtext(8 << 20, seed=13)
    d = corpus[:65536]
    nmsg = 8192
    rng = np.random.default_rng(1)
    starts = rng.integers(65536, len(corpus) - 4096, size=nmsg).astype(np.uint64)
    gd, cd = gpu.Dict(d), codec.dict_create(d)
    packed, poff = gpu.compress_batch(corpus, starts, [4096] * nmsg, gpu.compress_block_bound(4096), raw_blocks=True, dict=gd)
    sizes = np.diff(poff).astype(np.int64)
    assert (sizes > 0).all()
    sample = rng.choice(nmsg, size=200, replace=False)
    ref_total = 0
    for i in sample:
        m = corpus[int(starts[i]): int(starts[i]) + 4096]
        c = packed[int(poff[i]): int(poff[i + 1])].tobytes()
        assert cd.decompress(c, 4096) == (4096, m)
        ref_total += len(cd.compress(m))
    assert sizes[sample].sum() <= 1.03 * ref_total
    out, res = gpu.decompress_batch(packed, poff[:-1], 4096, raw_len=sizes.astype(np.uint32), dict=gd)
    assert (res == 4096).all()
    want = np.frombuffer(corpus, dtype=np.uint8)
    for i in sample:
        assert out[i].tobytes() == want[int(starts[i]): int(starts[i]) + 4096].tobytes()


def test_config5_mixed_entropy_stream_256k_blocks(gpu, port):
    """configs[4] in miniature: logtext / random (-> stored blocks) / zero pages / repeated 1025-byte record,
    256 KiB blocks, block + content checksums, streamed through NewWriter / NewReader."""
    seg = 1 << 20
    parts = []
    for i in range(24):
        kind = ["log", "log", "random", "zeros", "record1025", "log", "record1025", "log", "random", "log"][i % 10]
        parts.append(make(kind, seg, seed=i))
    data = b"".join(parts)
    dst = io.BytesIO()
    w = gpu.NewWriter(dst, block_size_idx=5, block_checksum=True, content_checksum=True, pending_size=8 << 20)
    for i in range(0, len(data), 3_000_001):
        w.write(data[i:i + 3_000_001])
    w.close()
    frame = dst.getvalue()
    stored = []
    pos = 7
    while True:
        word = int.from_bytes(frame[pos:pos + 4], "little")
        if word == 0:
            break
        stored.append(bool(word & 0x80000000))
        pos += 4 + (word & 0x7FFFFFFF) + 4
    assert any(stored) and not all(stored)                       # random segments are stored raw, the rest compressed
    assert F.read_frames(frame, port) == data
    r = gpu.NewReader(io.BytesIO(frame), pending_size=4 << 20)
    assert r.read_all() == data
    r.close()


def test_config2_device_resident_properties_at_scale(gpu, port):
    """configs[1] at 1 GiB, device-resident: exact round trip, every block decoded, and the xxh32 trailers of a
    sample of records equal the oracle's xxh32 of the same payloads (a checksum of checksums)."""
    torch = pytest.importorskip("torch")
    L = gpu._lib.lib()
    bsz, nblk = 65536, 16384
    n = bsz * nblk
    dev = torch.device("cuda", 0)
    stride = bsz + 16
    u8 = lambda k: torch.empty(k, dtype=torch.uint8, device=dev)
    src, recs, out = u8(n), u8(nblk * stride), u8(n)
    off = torch.arange(nblk, dtype=torch.int64, device=dev) * bsz
    ln = torch.full((nblk,), bsz, dtype=torch.int32, device=dev)
    roff = torch.arange(nblk, dtype=torch.int64, device=dev) * stride
    rlen = torch.zeros(nblk, dtype=torch.int32, device=dev)
    olen = torch.zeros(nblk, dtype=torch.int32, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())
    gpu._lib.check(L.plz4cu_gen_logtext_device(None, 0x504C5A34, 0, p(src), n))
    gpu._lib.check(L.plz4cu_compress_batch_device(None, p(src), p(off), p(ln), nblk, bsz, 1, 0, None, p(recs), stride, p(rlen)))
    gpu._lib.check(L.plz4cu_decompress_batch_device(None, p(recs), p(roff), None, nblk, bsz, 1, 0, None, p(out), bsz, p(olen)))
    torch.cuda.synchronize()
    assert bool((olen == bsz).all()) and torch.equal(out, src)
    assert src[:200000].cpu().numpy().tobytes() == logtext(200000)           # device generator == host generator
    recs_h = recs.view(nblk, stride)[::997].cpu().numpy()
    lens_h = rlen[::997].cpu().numpy()
    for row, rl in zip(recs_h, lens_h):
        c = int.from_bytes(row[:4].tobytes(), "little") & 0x7FFFFFFF
        assert rl == c + 8
        assert int.from_bytes(row[4 + c: 8 + c].tobytes(), "little") == port.xxh32(row[4: 4 + c].tobytes())
    ratio = float(rlen.to(torch.int64).sum()) / n
    assert 0.36 < ratio < 0.40


# ---------------------------------------------------------------- the configs at their stated scale
# Size-independent properties (round trip, checksum of checksums, the reference decodes a sample) at the sizes
# BASELINE.json quotes: configs[0] 256 MiB, configs[2] 64 starts into a 1 GiB reference frame, configs[3] 1 Mi payloads.

def test_config1_at_scale_256mib_4mib_blocks_both_checksums(gpu, port):
    data = logtext(256 << 20, seed=21)
    dst = io.BytesIO()
    w = gpu.NewWriter(dst, block_size_idx=7, block_checksum=True, content_checksum=True, content_size=len(data))
    assert w.read_from(io.BytesIO(data)) == len(data)
    w.close()
    frame = dst.getvalue()
    assert F.read_frames(frame, port) == data                        # reference-format reader: every block + content checksum
    out = io.BytesIO()
    r = gpu.NewReader(io.BytesIO(frame))
    assert r.write_to(out) == len(data)
    r.close()
    assert out.getbuffer().nbytes == len(data) and out.getvalue() == data
    # per block within tolerance of liblz4 on the same block (64 blocks of 4 MiB)
    pos, b = 7 + 8, 0
    while True:
        word = int.from_bytes(frame[pos:pos + 4], "little")
        if word == 0:
            break
        n = word & 0x7FFFFFFF
        ref_rec = port.block_record(data[b * (4 << 20):(b + 1) * (4 << 20)], 4 << 20, True)
        assert 8 + n <= 1.03 * len(ref_rec), (b, n, len(ref_rec))
        pos += 8 + n
        b += 1
    assert b == 64


def test_config3_at_scale_64_starts_into_a_1gib_reference_frame(gpu, port):
    bsz = 4 << 20
    data = logtext(1 << 30, seed=22)
    marks = []
    frame = F.write_frame(data, F.Opts(block_idx=7, block_checksum=True, content_checksum=False), port,
                          progress=lambda s, d: marks.append((s, d)))
    assert len(marks) >= 256
    rng = random.Random(5)
    want = 32 << 20                                                  # bytes read from every start
    for s, d in rng.sample(marks[:-1], 64):
        r = gpu.NewReader(io.BytesIO(frame), read_offset=d)
        got = bytearray()
        while len(got) < want:
            chunk = r.read(want - len(got))
            if not chunk:
                break
            got += chunk
        r.close()
        assert bytes(got) == data[s:s + want], (s, d)


def test_config4_at_scale_one_million_payloads_with_dictionary(gpu, port, codec):
    torch = pytest.importorskip("torch")
    L = gpu._lib.lib()
    corpus = logtext(64 << 20, seed=23)
    d = corpus[:65536]
    nmsg, msz = 1 << 20, 4096
    rng = np.random.default_rng(7)
    starts = rng.integers(65536, len(corpus) - msz, size=nmsg).astype(np.int64)
    gd, cd = gpu.Dict(d), codec.dict_create(d)
    dev = torch.device("cuda", 0)
    src = torch.frombuffer(bytearray(corpus), dtype=torch.uint8).to(dev)
    off = torch.from_numpy(starts).to(dev)
    ln = torch.full((nmsg,), msz, dtype=torch.int32, device=dev)
    cap = gpu.compress_block_bound(msz)
    stride = (cap + 15) // 16 * 16
    recs = torch.empty(nmsg * stride, dtype=torch.uint8, device=dev)
    rlen = torch.zeros(nmsg, dtype=torch.int32, device=dev)
    out = torch.empty(nmsg * msz, dtype=torch.uint8, device=dev)
    olen = torch.zeros(nmsg, dtype=torch.int32, device=dev)
    roff = torch.arange(nmsg, dtype=torch.int64, device=dev) * stride
    ooff = torch.arange(nmsg, dtype=torch.int64, device=dev) * msz
    h_src = torch.zeros(nmsg, dtype=torch.int32, device=dev)
    h_out = torch.zeros(nmsg, dtype=torch.int32, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())
    dh = gd.handle if hasattr(gd, "handle") else gd._h
    gpu._lib.check(L.plz4cu_compress_batch_device(None, p(src), p(off), p(ln), nmsg, cap, 0, 1, dh, p(recs), stride, p(rlen)))
    gpu._lib.check(L.plz4cu_decompress_batch_device(None, p(recs), p(roff), p(rlen), nmsg, msz, 0, 1, dh, p(out), msz, p(olen)))
    gpu._lib.check(L.plz4cu_xxh32_batch_device(None, p(src), p(off), p(ln), nmsg, p(h_src)))
    gpu._lib.check(L.plz4cu_xxh32_batch_device(None, p(out), p(ooff), p(ln), nmsg, p(h_out)))
    torch.cuda.synchronize()
    assert bool((rlen > 0).all()) and bool((olen == msz).all())
    assert torch.equal(h_src, h_out)                                 # checksum of checksums: every payload came back
    sample = rng.choice(nmsg, size=300, replace=False)
    rl = rlen.cpu().numpy()
    tot_gpu = tot_ref = 0
    for i in sample:
        m = corpus[int(starts[i]): int(starts[i]) + msz]
        c = recs[int(i) * stride: int(i) * stride + int(rl[i])].cpu().numpy().tobytes()
        assert cd.decompress(c, msz) == (msz, m)                     # the reference decodes it with the same dictionary
        tot_gpu += len(c)
        tot_ref += len(cd.compress(m))
    assert tot_gpu <= 1.03 * tot_ref
