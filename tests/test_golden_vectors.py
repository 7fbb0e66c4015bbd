"""Committed golden vectors (tests/golden/vectors.json, generated from the reference's own liblz4 by
tests/golden/make_golden.py): the oracle port must reproduce them on CPU, the CUDA engine on the GPU."""
import base64
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import frame_oracle as F
from tests.datagen import make

V = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vectors.json")))
sha = lambda b: hashlib.sha256(b).hexdigest()
unb = lambda s: base64.b64decode(s)


def test_port_reproduces_reference_block_bytes(port):
    for e in V["blocks"]:
        c = port.compress(make(e["kind"], e["n"]), e["cap"])
        assert (None if c is None else len(c)) == e["len"], e
        if c is not None:
            assert sha(c) == e["sha256"], e
            if "bytes" in e:
                assert c == unb(e["bytes"])


def test_port_reproduces_reference_dictionary_bytes(port):
    for e in V["dict_blocks"]:
        d = make("words", e["dict_n"], seed=3)
        s = (d[-40:] + make("words", e["n"], seed=3))[: e["n"]] if e["tail40"] else make("words", e["n"], seed=3)
        c = port.dict_create(d).compress(s)
        assert len(c) == e["len"] and sha(c) == e["sha256"], e


def test_port_reproduces_reference_decode_codes(port):
    for e in V["decode"]:
        r, data = port.decompress(unb(e["stream"]), e["cap"])
        assert r == e["ret"], e
        assert (None if data is None else sha(data)) == e["sha256"]


def test_port_xxh32_and_frames(port):
    for e in V["xxh32"]:
        assert port.xxh32(make("random", e["n"])) == e["value"]
    for e in V["frames"]:
        f = F.write_frame(make("log", e["n"]), F.Opts(block_idx=e["block_idx"], block_checksum=e["bx"], content_checksum=e["cx"]), port)
        assert len(f) == e["len"] and sha(f) == e["sha256"] and f[:32] == unb(e["head"])


@pytest.mark.gpu
def test_gpu_decoder_reproduces_reference_decode_codes(gpu):
    """Same streams, same capacities, same return codes and bytes as the reference decoder produced."""
    by_cap = {}
    for e in V["decode"]:
        by_cap.setdefault(e["cap"], []).append(e)
    for cap, es in by_cap.items():
        streams = [unb(e["stream"]) for e in es]
        keep = [i for i, s in enumerate(streams) if len(s)]
        buf = b"".join(streams[i] for i in keep)
        off = np.cumsum([0] + [len(streams[i]) for i in keep])[:-1]
        out, res = gpu.decompress_batch(buf, off, cap, raw_len=[len(streams[i]) for i in keep])
        for j, i in enumerate(keep):
            assert res[j] == es[i]["ret"], (es[i], res[j])
            if es[i]["ret"] >= 0:
                assert sha(out[j, : res[j]].tobytes()) == es[i]["sha256"]


@pytest.mark.gpu
def test_gpu_encoder_output_is_decoded_by_reference_and_close_in_size(gpu, codec):
    tot_gpu = tot_ref = 0
    for e in V["blocks"]:
        if e["cap"] is not None or e["len"] is None:
            continue
        s = make(e["kind"], e["n"])
        c = gpu.compress_block(s)
        assert codec.decompress(c, len(s)) == (len(s), s)
        tot_gpu += len(c)
        tot_ref += e["len"]
    assert tot_gpu <= 1.03 * tot_ref
    for e in V["frames"]:
        import io
        dst = io.BytesIO()
        w = gpu.NewWriter(dst, block_size_idx=e["block_idx"], block_checksum=e["bx"], content_checksum=e["cx"])
        w.write(make("log", e["n"]))
        w.close()
        f = dst.getvalue()
        assert f[:7] == unb(e["head"])[:7]                       # byte-identical frame header
        assert F.read_frames(f, codec) == make("log", e["n"])
