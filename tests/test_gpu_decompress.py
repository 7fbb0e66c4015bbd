"""GPU decode parity: the CUDA decoder against the oracle on the same bytes (bit-exact, incl. codes)."""
import random

import os
import subprocess
import sys

import numpy as np
import pytest

from tests.datagen import KINDS, make

pytestmark = pytest.mark.gpu

SIZES = [0, 1, 5, 12, 13, 14, 31, 32, 33, 64, 100, 1000, 4096, 65535, 65536]


def _records(port, blocks, bsz, checksum):
    recs, offs, pos = bytearray(), [], 0
    for b in blocks:
        r = port.block_record(b, bsz, checksum)
        offs.append(pos)
        recs += r
        pos += len(r)
    return bytes(recs), offs


def test_raw_blocks_roundtrip_exact(gpu, codec):
    srcs = [make(k, n) for k in KINDS for n in SIZES]
    comp = [codec.compress(s) for s in srcs]
    buf, off = b"".join(comp), np.cumsum([0] + [len(c) for c in comp])[:-1]
    out, res = gpu.decompress_batch(buf, off, 65536, raw_len=[len(c) for c in comp])
    for i, s in enumerate(srcs):
        assert res[i] == len(s), (i, res[i], len(s))
        assert out[i, : len(s)].tobytes() == s


@pytest.mark.parametrize("slack", [0, 1, 5, 11, 12, 13, 31, 32, 33, 64])
def test_capacity_edges_match_oracle(gpu, codec, slack):
    """liblz4's accept/reject depends on dst capacity near the end of a block; codes must agree."""
    for kind in ["words", "ab", "log", "runs"]:
        for n in [13, 20, 64, 300, 5000]:
            s = make(kind, n)
            c = codec.compress(s)
            for cap in {n + slack, max(n - slack, 0)}:
                want, data = codec.decompress(c, cap)
                out, res = gpu.decompress_batch(c, [0], cap, raw_len=[len(c)])
                assert res[0] == want, (kind, n, cap, res[0], want)
                if want >= 0:
                    assert out[0, :want].tobytes() == data


def test_corrupted_streams_same_codes(gpu, port, codec):
    rng = random.Random(99)
    cases = []
    for it in range(3000):
        n = rng.choice([0, 1, 5, 13, 20, 64, 100, 300, 1000, 5000])
        c = bytearray(codec.compress(make(rng.choice(["words", "ab", "random", "zeros", "runs"]), n, seed=it)))
        m = rng.randrange(5)
        if m == 0 and c:
            for _ in range(rng.randint(1, 3)):
                c[rng.randrange(len(c))] = rng.getrandbits(8)
        elif m == 1 and c:
            c = c[: rng.randrange(len(c) + 1)]
        elif m == 2:
            c += rng.randbytes(rng.randint(1, 20))
        elif m == 3 and c:
            c[rng.randrange(len(c))] = rng.choice([0xFF, 0xF0, 0x0F, 0x00])
        else:
            c = bytearray(rng.randbytes(rng.randint(1, 64)))
        if len(c) == 0:
            continue
        cases.append((bytes(c), n + rng.choice([0, 1, 12, 40, 100])))
    cases += [(b"not-a-valid-lz4-block", 84), (b"\xff" * 64, 256)]          # block_test.go:315-319 (G7)
    cap = max(c[1] for c in cases)
    hits = port.lib.orc_dbg_zero_offset_hits
    hits.restype = __import__("ctypes").c_uint64
    for group_cap in sorted({c[1] for c in cases}):
        grp = [c[0] for c in cases if c[1] == group_cap]
        buf, off = b"".join(grp), np.cumsum([0] + [len(c) for c in grp])[:-1]
        out, res = gpu.decompress_batch(buf, off, group_cap, raw_len=[len(c) for c in grp])
        for i, c in enumerate(grp):
            z = hits()
            want, data = port.decompress(c, group_cap)        # port == liblz4 (tests/test_oracle_vs_ref.py)
            assert res[i] == want, (c.hex(), group_cap, res[i], want)
            if want >= 0:
                assert out[i, :want].tobytes() == data
            elif hits() == z:
                rw, _ = codec.decompress(c, group_cap)
                assert rw == want


def _litruns(rng, n):
    """random literal runs of 10..90 bytes separated by copies of earlier data: tokens with literal nibble 15"""
    out = bytearray()
    while len(out) < n:
        out += rng.randbytes(rng.randint(10, 90))
        if len(out) > 40:
            k = rng.randint(4, 40)
            s = rng.randrange(0, len(out) - k)
            out += out[s:s + k]
    return bytes(out[:n])


def test_long_literal_runs_same_codes(gpu, port, codec):
    """Sequences with a literal-length extension byte go through the batched path: exact lengths, bytes and error
    codes against liblz4 on intact, truncated and mutated streams, at exact and tight capacities."""
    rng = random.Random(1234)
    cases = []
    for it in range(1500):
        n = rng.choice([60, 200, 1000, 5000, 20000, 70000])
        data = _litruns(rng, n)
        c = bytearray(codec.compress(data))
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
        hits = port.lib.orc_dbg_zero_offset_hits
        hits.restype = __import__("ctypes").c_uint64
        for i, c in enumerate(grp):
            z = hits()
            want, data = port.decompress(c, group_cap)
            assert res[i] == want, (c.hex()[:200], group_cap, res[i], want)
            if want >= 0:
                assert out[i, :want].tobytes() == data
            elif hits() == z:
                rw, _ = codec.decompress(c, group_cap)
                assert rw == want


def test_frame_records_checksum_stored_overflow(gpu, port):
    bsz = 65536
    blocks = [make("log", bsz), make("random", bsz), make("zeros", bsz), make("words", 777), make("random", 100), b"hello"]
    recs, offs = _records(port, blocks, bsz, True)
    out, res = gpu.decompress_batch(recs, offs, bsz, verify_checksum=True)
    for i, b in enumerate(blocks):
        assert res[i] == len(b)
        assert out[i, : len(b)].tobytes() == b
    # flip one payload byte in every record -> block hash mismatch (blk/frame.go:114-127)
    bad = bytearray(recs)
    for o in offs:
        bad[o + 5] ^= 0x40
    out, res = gpu.decompress_batch(bytes(bad), offs, bsz, verify_checksum=True)
    assert all(r == gpu._lib.E_BLOCKHASH for r in res)
    # size word larger than the block size -> overflow (blk/frame.go:79-81)
    big = bytearray(recs)
    big[offs[0]: offs[0] + 4] = (bsz + 1).to_bytes(4, "little")
    out, res = gpu.decompress_batch(bytes(big), offs[:1], bsz, verify_checksum=True)
    assert res[0] == gpu._lib.E_OVERFLOW


def test_dictionary_decode(gpu, port, codec):
    d = make("words", 70000, seed=5)
    pd = port.dict_create(d)
    gd = gpu.Dict(d)
    srcs = [make("words", n, seed=5) for n in [0, 5, 100, 4096, 20000, 65536]] + [d[-3000:] + make("words", 1000, seed=5)]
    comp = [pd.compress(s) for s in srcs]
    buf, off = b"".join(comp), np.cumsum([0] + [len(c) for c in comp])[:-1]
    out, res = gpu.decompress_batch(buf, off, 65536, raw_len=[len(c) for c in comp], dict=gd)
    for i, s in enumerate(srcs):
        assert res[i] == len(s)
        assert out[i, : len(s)].tobytes() == s
    # wrong dictionary must not silently give the right bytes (block_test.go:223-307)
    g2 = gpu.Dict(make("words", 70000, seed=6))
    out2, res2 = gpu.decompress_batch(comp[-1], [0], 65536, raw_len=[len(comp[-1])], dict=g2)
    assert res2[0] < 0 or out2[0, : res2[0]].tobytes() != srcs[-1]


@pytest.mark.parametrize("duo", ["0", "1"])
def test_whole_decode_suite_through_either_decoder(duo):
    """Launches that cannot fill the SMs with one warp per block give a block two warps (one parses and lists batches, one
    produces the bytes: lz4_decompress_duo_kernel), the others one (lz4_decompress_kernel); PLZ4CU_DEC_DUO=0 / 1 forces
    either for every launch.  Bytes and return codes must be the same: the whole parity suite, the at-scale configs and the
    streams run through each."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PLZ4CU_DEC_DUO=duo)
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "-k", "not whole_decode_suite",
                        "tests/test_gpu_decompress.py", "tests/test_gpu_configs.py", "tests/test_gpu_stream.py"],
                       cwd=root, env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
