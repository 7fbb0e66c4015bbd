"""GPU encode parity: every block the CUDA encoder emits is decoded byte-exactly by the reference
decoder, framed exactly like blk.CompressToBlk, and within 3 % of liblz4's size."""
import numpy as np
import pytest

from tests.datagen import KINDS, make

pytestmark = pytest.mark.gpu

SIZES = [0, 1, 5, 12, 13, 14, 31, 32, 33, 64, 100, 1000, 4096, 65535, 65536]
TOLERANCE = 1.03    # BASELINE.json north_star: compressed size within 3 % of liblz4 at the same level


def test_golden_hello_and_empty(gpu):
    assert gpu.compress_block(b"hello") == bytes.fromhex("5068656c6c6f")       # plz4_test.go:74 (G1)
    assert gpu.compress_block(b"") == b"\x00"                                   # block_test.go:22-52 (G8)
    assert gpu.decompress_block(bytes.fromhex("5068656c6c6f")) == b"hello"      # plz4_test.go:12
    assert gpu.decompress_block(b"\x00") == b""


def test_raw_blocks_decode_with_reference(gpu, codec):
    srcs = [make(k, n) for k in KINDS for n in SIZES]
    buf, off = b"".join(srcs), np.cumsum([0] + [len(s) for s in srcs])[:-1]
    lens = [len(s) for s in srcs]
    packed, poff = gpu.compress_batch(buf, off, lens, gpu.compress_block_bound(65536), raw_blocks=True)
    for i, s in enumerate(srcs):
        c = packed[int(poff[i]): int(poff[i + 1])].tobytes()
        assert len(c) > 0
        r, data = codec.decompress(c, len(s))          # exact capacity: enforces the end-of-block rules
        assert r == len(s) and data == s, (i, r, len(s))
        ref_c = codec.compress(s)
        assert len(c) <= max(len(ref_c) * TOLERANCE, len(ref_c) + 8), (i, len(c), len(ref_c))


@pytest.mark.parametrize("bsz", [65536, 262144, 4 << 20])
@pytest.mark.parametrize("checksum", [False, True])
def test_frame_records_layout(gpu, port, codec, bsz, checksum):
    """Record bytes follow blk/blk.go:87-106: size word, stored bit, payload, xxh32 over the payload."""
    kinds = ["log", "random", "zeros", "words", "runs", "record1025"]
    blocks = [make(k, bsz) for k in kinds] + [make("log", 1000), make("random", 300), b"hello"]
    buf, off = b"".join(blocks), np.cumsum([0] + [len(b) for b in blocks])[:-1]
    packed, poff = gpu.compress_batch(buf, off, [len(b) for b in blocks], bsz, block_checksum=checksum)
    tot_gpu = tot_ref = 0
    for i, b in enumerate(blocks):
        rec = packed[int(poff[i]): int(poff[i + 1])].tobytes()
        word = int.from_bytes(rec[:4], "little")
        n = word & 0x7FFFFFFF
        assert len(rec) == 4 + n + (4 if checksum else 0)
        payload = rec[4: 4 + n]
        if checksum:
            assert int.from_bytes(rec[4 + n:], "little") == port.xxh32(payload)
        ref_rec = port.block_record(b, bsz, checksum)
        ref_stored = bool(int.from_bytes(ref_rec[:4], "little") & 0x80000000)
        if word & 0x80000000:
            assert payload == b
            assert ref_stored or len(ref_rec) >= len(b) * 0.97
        else:
            assert n <= bsz
            r, data = codec.decompress(payload, bsz)
            assert r == len(b) and data == b
        # per block, every kind, every block size (large blocks: liblz4 hashes five bytes there, lz4.c:785-795,1391-1400)
        assert len(rec) <= max(len(ref_rec) * TOLERANCE, len(ref_rec) + 8), (i, len(rec), len(ref_rec))
        tot_gpu += len(rec)
        tot_ref += len(ref_rec)
    assert tot_gpu <= tot_ref * TOLERANCE


def test_logtext_ratio_vs_liblz4(gpu, codec):
    """The benchmark workload itself: 64 x 64 KiB blocks of log text, total size within tolerance."""
    from tests.datagen import logtext
    bsz, nblk = 65536, 64
    data = logtext(bsz * nblk)
    off = np.arange(nblk, dtype=np.uint64) * bsz
    packed, poff = gpu.compress_batch(data, off, [bsz] * nblk, bsz, block_checksum=True)
    ref_total = sum(len(codec.compress(data[i * bsz:(i + 1) * bsz], bsz)) + 8 for i in range(nblk))
    ratio = int(poff[nblk]) / ref_total
    print(f"gpu/liblz4 size ratio on logtext: {ratio:.4f}")
    assert ratio <= TOLERANCE
    out, res = gpu.decompress_batch(packed, poff[:-1], bsz, verify_checksum=True)
    assert (res == bsz).all()
    assert out.reshape(-1).tobytes() == data


def test_dictionary_compress_small_messages(gpu, port, codec):
    """BASELINE configs[3] in miniature: 4 KiB payloads sharing a 64 KiB dictionary (raw block API)."""
    from tests.datagen import logtext
    corpus = logtext(1 << 20, seed=77)
    d = corpus[:65536]
    rng = np.random.default_rng(5)
    starts = rng.integers(65536, len(corpus) - 4096, size=256)
    msgs = [corpus[s: s + 4096] for s in starts] + [b"", b"x", d[-100:], d[1000:1200] * 3, make("random", 4096)]
    gd, cd = gpu.Dict(d), codec.dict_create(d)
    buf, lens = b"".join(msgs), [len(m) for m in msgs]
    off = np.cumsum([0] + lens)[:-1]
    packed, poff = gpu.compress_batch(buf, off, lens, gpu.compress_block_bound(4096), raw_blocks=True, dict=gd)
    tot_gpu = tot_ref = tot_nodict = 0
    for i, m in enumerate(msgs):
        c = packed[int(poff[i]): int(poff[i + 1])].tobytes()
        r, data = cd.decompress(c, len(m))               # the reference decoder with the same dictionary
        assert r == len(m) and data == m, i
        tot_gpu += len(c)
        tot_ref += len(cd.compress(m))
        tot_nodict += len(codec.compress(m))
    print(f"dict: gpu {tot_gpu} liblz4+dict {tot_ref} liblz4 no dict {tot_nodict}")
    assert tot_gpu <= tot_ref * TOLERANCE
    assert tot_gpu < tot_nodict * 0.92                   # the dictionary is really used (wr_test.go:471-625)
    # and back through the GPU decoder with the dictionary
    out, res = gpu.decompress_batch(packed, poff[:-1], 4096, raw_len=np.diff(poff).astype(np.uint32), dict=gd)
    for i, m in enumerate(msgs):
        assert res[i] == len(m) and out[i, : len(m)].tobytes() == m


@pytest.mark.parametrize("dn", [0, 3, 4, 100, 5000, 65535, 65536, 70000])
def test_dictionary_sizes_and_block_api(gpu, codec, dn):
    d = make("words", dn, seed=9)
    gd, cd = gpu.Dict(d), codec.dict_create(d)
    for n in [0, 1, 13, 300, 4096, 65536, 100000]:
        for kind in ["words", "log"]:
            m = make(kind, n, seed=9)
            if dn >= 64 and n >= 64:
                m = (d[-50:] + m)[:n]                    # begins with the dictionary's tail
            c = gpu.compress_block(m, dict=gd)
            r, data = cd.decompress(c, len(m))
            assert r == len(m) and data == m, (dn, n, kind)
            assert gpu.decompress_block(c, dst_cap=len(m), dict=gd) == m


def test_large_blocks_with_dictionary(gpu, codec):
    """Blocks above 64 KiB with a dictionary go through the span encoder with the dictionary as the fragment before
    the block (round 1 gave them one warp each).  The reference decodes them with the same dictionary, every block is
    within the tolerance of liblz4 + dictionary, the dictionary is really used, and a block's bytes do not depend on
    what else the batch holds (small blocks with a dictionary take another kernel in the same call)."""
    from tests.datagen import logtext
    corpus = logtext(6 << 20, seed=41)
    for dn in (65536, 20000, 777):
        d = corpus[100000: 100000 + dn]
        gd, cd = gpu.Dict(d), codec.dict_create(d)
        blocks = [d[-min(dn, 3000):] * 2 + corpus[300000: 300000 + n] for n in (66000, 200000, 1 << 20)]
        rep = d[-min(dn, 40000):]                        # (a copy exactly 64 KiB back is out of an offset's reach)
        blocks += [corpus[2 << 20: (2 << 20) + (3 << 20)], make("words", 300000, seed=4), rep * (200000 // len(rep) + 1)]
        small = [corpus[50000:54096], b"", d[-200:] * 4]
        mixed = blocks + small
        cap = max(len(b) for b in mixed)
        lens = [len(b) for b in mixed]
        off = np.cumsum([0] + lens)[:-1]
        packed, poff = gpu.compress_batch(b"".join(mixed), off, lens, gpu.compress_block_bound(cap), raw_blocks=True, dict=gd)
        got = [packed[int(poff[i]): int(poff[i + 1])].tobytes() for i in range(len(mixed))]
        for i, m in enumerate(mixed):
            r, data = cd.decompress(got[i], len(m))
            assert r == len(m) and data == m, (dn, i)
            ref = len(cd.compress(m))
            # (a block that is nearly one match compresses to a few hundred bytes: there a fragment boundary's 4-5 bytes and
            # a first match that starts some dozen bytes later are percents of nothing — a thousandth of the input is allowed)
            assert len(got[i]) <= max(ref * TOLERANCE, ref + len(m) // 1000 + 16), (dn, i, len(got[i]), ref)
        # the dictionary is used: the block that repeats it costs next to nothing, and far less than without it
        assert len(got[len(blocks) - 1]) < 0.02 * len(blocks[-1]) + 64
        if dn >= 20000:
            assert len(got[0]) < len(codec.compress(blocks[0]))
        # alone, a large block gives the same bytes as in the batch (small blocks with a dictionary keep round 1's encoder,
        # whose table size follows the capacity the call states)
        for i in (0, 2):
            assert gpu.compress_block(mixed[i], dict=gd) == got[i], (dn, i)
        # and back through the GPU decoder
        out, res = gpu.decompress_batch(packed, poff[:-1], cap, raw_len=np.diff(poff).astype(np.uint32), dict=gd)
        for i, m in enumerate(mixed):
            assert res[i] == len(m) and out[i, : len(m)].tobytes() == m, (dn, i)


def test_compress_is_deterministic(gpu):
    """The same bytes compress to the same bytes whatever the timing between a block's tiles (stream writers rely on it:
    progress marks of one run address the frame of another, tests/test_gpu_frame_device.py)."""
    from tests.datagen import logtext
    bsz = 65536
    blocks = [logtext(bsz, seed=1000 + i) for i in range(48)] + [make(k, bsz, seed=3) for k in ["words", "runs", "ab", "zeros", "record1025", "random"]] * 4
    buf, off = b"".join(blocks), np.arange(len(blocks), dtype=np.uint64) * bsz
    first = None
    for _ in range(16):
        packed, poff = gpu.compress_batch(buf, off, [bsz] * len(blocks), bsz, block_checksum=True)
        got = (packed[: int(poff[-1])].tobytes(), poff.tobytes())
        if first is None:
            first = got
        assert got == first


@pytest.mark.parametrize("n", [100000, 262144, 1 << 20, (4 << 20) + 12345])
@pytest.mark.parametrize("kind", ["log", "words", "ab", "runs", "zeros", "record1025", "random"])
def test_large_blocks_per_block_size_and_reference_decode(gpu, codec, kind, n):
    """Blocks above 64 KiB (spans of fragments with the table and a 64 KiB window kept, compress_cta.cu): the reference decodes
    every one and each stays within tolerance of liblz4 on the same block."""
    s = make(kind, n, seed=5)
    c = gpu.compress_block(s)
    r, data = codec.decompress(c, n)
    assert r == n and data == s
    ref_c = codec.compress(s)
    assert len(c) <= max(len(ref_c) * TOLERANCE, len(ref_c) + 8), (len(c), len(ref_c))


def test_block_longer_than_the_room_offered(gpu, codec):
    """plz4_block.go:100-109 WithBlockDst: the caller may offer less room than the block is long; a compressible block still
    goes in whole (fragments are sized from the data, not from the room), an incompressible one is refused."""
    s = make("log", 1 << 20, seed=8)
    c = gpu.compress_block(s, dst_cap=600 << 10)
    assert 0 < len(c) <= 600 << 10
    back, data = codec.decompress(c, len(s))
    assert back == len(s) and data == s
    with pytest.raises(gpu.Lz4Error):
        gpu.compress_block(make("random", 1 << 20, seed=8), dst_cap=600 << 10)
    # just above the 66-fragment mark of the one-warp-per-fragment path
    big = make("log", (4 << 20) + (200 << 10), seed=9)
    c = gpu.compress_block(big)
    back, data = codec.decompress(c, len(big))
    assert back == len(big) and data == big


def test_a_block_compresses_the_same_in_any_batch(gpu):
    """Kernels are chosen per block (up to 64 KiB: one CTA; above: spans of sixteen fragments), never by what else the launch
    holds: the bytes of a block do not depend on its batch — what lets several writers, batch sizes or GPUs produce one frame."""
    sizes = [4 << 20, 30_000, 200_000, 65536, 10, 1 << 20, 65537, 0, 3_000_000]
    blocks = [make("log", n, seed=40 + i) for i, n in enumerate(sizes)]
    cap = gpu.compress_block_bound(max(sizes))
    buf, off = b"".join(blocks), np.cumsum([0] + sizes)[:-1]
    packed, poff = gpu.compress_batch(buf, off, sizes, cap, raw_blocks=True)
    together = [packed[int(poff[i]): int(poff[i + 1])].tobytes() for i in range(len(sizes))]
    for i, b in enumerate(blocks):
        if not b:
            continue
        alone, p1 = gpu.compress_batch(b, [0], [len(b)], cap, raw_blocks=True)
        assert alone[: int(p1[1])].tobytes() == together[i], (i, sizes[i])
    # and in a batch of its own kind
    same, p2 = gpu.compress_batch(blocks[3] * 3, [0, 65536, 131072], [65536] * 3, cap, raw_blocks=True)
    assert same[: int(p2[1])].tobytes() == together[3]
