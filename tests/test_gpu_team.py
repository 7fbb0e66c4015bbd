"""The team decoder (one CTA per block: parser warp, copy warps, checksum warp — decompress.cu) takes launches of few,
large blocks.  Same bar as the one-warp decoder: bytes and return codes bit-exact against the oracle."""
import os
import random
import subprocess
import sys

import numpy as np
import pytest

from tests.datagen import make

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
E_BLOCKHASH = -0x7F000001


def _records(port, blocks, bsz, checksum):
    recs, offs, pos = bytearray(), [], 0
    for b in blocks:
        r = port.block_record(b, bsz, checksum)
        offs.append(pos)
        recs += r
        pos += len(r)
    return bytes(recs), offs


@pytest.mark.parametrize("bsz", [70000, 262144, 4 << 20])
@pytest.mark.parametrize("checksum", [False, True])
def test_large_block_records_decode_exact(gpu, port, bsz, checksum):
    kinds = ["log", "words", "runs", "record1025", "zeros", "ab", "random"]
    blocks = [make(k, bsz, seed=3) for k in kinds] + [make("log", bsz - 1), make("log", 65537), make("words", 13), b"", b"hello"]
    recs, offs = _records(port, blocks, bsz, checksum)
    out, res = gpu.decompress_batch(recs, offs, bsz, verify_checksum=checksum)
    for i, b in enumerate(blocks):
        assert res[i] == len(b), (i, res[i], len(b))
        assert out[i, : len(b)].tobytes() == b, i


def test_large_block_checksum_and_overflow_codes(gpu, port):
    bsz = 1 << 20
    blocks = [make("log", bsz), make("random", bsz), make("words", 300000)]
    recs, offs = _records(port, blocks, bsz, True)
    bad = bytearray(recs)
    bad[offs[0] + 4 + 1000] ^= 1            # compressed payload of block 0
    bad[offs[1] + 4 + 77] ^= 0x80           # stored payload of block 1
    out, res = gpu.decompress_batch(bytes(bad), offs, bsz, verify_checksum=True)
    assert res[0] == E_BLOCKHASH and res[1] == E_BLOCKHASH and res[2] == 300000
    assert out[2, :300000].tobytes() == blocks[2]
    # without verification the flipped stored block is simply delivered
    out, res = gpu.decompress_batch(bytes(bad), offs, bsz, verify_checksum=False)
    assert res[1] == bsz and res[2] == 300000
    big = bytearray(recs)
    big[offs[0]: offs[0] + 4] = (bsz + 1).to_bytes(4, "little")
    out, res = gpu.decompress_batch(bytes(big), offs[:1], bsz, verify_checksum=True)
    assert res[0] == -0x7F000002


def test_large_block_corruption_same_codes(gpu, port, codec):
    """Intact, truncated and mutated streams of a few hundred KiB at exact and tight capacities: the parser warp is the
    one-warp decoder's own parse, so every code must be liblz4's."""
    rng = random.Random(4242)
    hits = port.lib.orc_dbg_zero_offset_hits
    hits.restype = __import__("ctypes").c_uint64
    cases = []
    for it in range(160):
        n = rng.choice([66000, 100000, 200000, 300000])
        kind = rng.choice(["log", "words", "runs", "record1025", "zeros"])
        c = bytearray(codec.compress(make(kind, n, seed=it)))
        m = rng.randrange(6)
        if m == 0:
            for _ in range(rng.randint(1, 3)):
                c[rng.randrange(len(c))] = rng.getrandbits(8)
        elif m == 1:
            c = c[: rng.randrange(1, len(c) + 1)]
        elif m == 2:
            c[rng.randrange(len(c))] = rng.choice([0xFF, 0xF0, 0xF4, 0x0F, 0x00])
        elif m == 3:
            i = rng.randrange(len(c)); c[i:i] = rng.randbytes(rng.randint(1, 3))
        cases.append((bytes(c), n + rng.choice([0, 0, 5, 11, 12, 13, 31, 32, 33, 100]) - rng.choice([0, 0, 0, 1, 7])))
    for group_cap in sorted({c[1] for c in cases}):
        grp = [c[0] for c in cases if c[1] == group_cap]
        buf, off = b"".join(grp), np.cumsum([0] + [len(c) for c in grp])[:-1]
        out, res = gpu.decompress_batch(buf, off, group_cap, raw_len=[len(c) for c in grp])
        for i, c in enumerate(grp):
            z = hits()
            want, data = port.decompress(c, group_cap)
            assert res[i] == want, (len(c), group_cap, res[i], want)
            if want >= 0:
                assert out[i, :want].tobytes() == data
            elif hits() == z:
                rw, _ = codec.decompress(c, group_cap)
                assert rw == want


def test_large_block_dictionary_decode(gpu, port):
    d = make("log", 65536, seed=9)
    pd = port.dict_create(d)
    srcs = [make("log", n, seed=9) for n in (70000, 262144, 1 << 20)] + [d[1000:5000] * 40]
    comp = [pd.compress(s) for s in srcs]
    gd = gpu.Dict(d)
    cap = 1 << 20
    buf, off = b"".join(comp), np.cumsum([0] + [len(c) for c in comp])[:-1]
    out, res = gpu.decompress_batch(buf, off, cap, raw_len=[len(c) for c in comp], dict=gd)
    for i, s in enumerate(srcs):
        assert res[i] == len(s), (i, res[i])
        assert out[i, : len(s)].tobytes() == s


def test_long_literal_runs_jump_over_superwindows(gpu, codec):
    """Random islands inside compressible text become literal runs of up to 70 KB: decode_one carries the parser over
    several 4 KiB superwindows at once and the table warps have to skip ahead with it."""
    rng = random.Random(7)
    blocks = []
    for it in range(10):
        parts, total = [], 0
        while total < 500000:
            if rng.random() < 0.5:
                p = rng.randbytes(rng.choice([100, 4000, 4200, 9000, 20000, 70000]))
            else:
                p = make("log", rng.choice([3000, 50000, 150000]), seed=it)
            parts.append(p)
            total += len(p)
        blocks.append(b"".join(parts))
    comp = [codec.compress(b) for b in blocks]
    cap = max(len(b) for b in blocks)
    buf, off = b"".join(comp), np.cumsum([0] + [len(c) for c in comp])[:-1]
    out, res = gpu.decompress_batch(buf, off, cap, raw_len=[len(c) for c in comp])
    for i, b in enumerate(blocks):
        assert res[i] == len(b), (i, res[i], len(b))
        assert out[i, : len(b)].tobytes() == b, i
    # the same streams cut short / with too little room must give liblz4's codes
    for i, c in enumerate(comp[:4]):
        for cut, room in ((len(c) // 2, cap), (len(c), len(blocks[i]) - 1), (len(c) - 1, cap)):
            want, _ = codec.decompress(c[:cut], room)
            _, r = gpu.decompress_batch(c[:cut], [0], room, raw_len=[cut])
            assert r[0] == want, (i, cut, room, r[0], want)


@pytest.mark.parametrize("pair", ["0", "1"])
def test_whole_decode_suite_through_the_team_kernel(pair):
    """PLZ4CU_TEAM=2 sends every launch below 1024 blocks through the team kernel, whatever the capacity: the one-warp
    decoder's own parity suite (capacity edges, 4500 corrupted streams, dictionaries) must pass unchanged, and so must
    the tests above, with one team per SM (64 KiB output window) and with two (32 KiB)."""
    env = dict(os.environ, PLZ4CU_TEAM="2", PLZ4CU_TEAM_PAIR=pair)
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "-k", "not whole_decode_suite",
                        "tests/test_gpu_decompress.py", "tests/test_gpu_team.py"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
