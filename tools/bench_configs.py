#!/usr/bin/env python3
"""Secondary measurements: the BASELINE.json configs that are NOT the headline bench line, scaled to
minutes, GPU engine next to the reference's liblz4 on all host cores.  Prints one JSON object.
These are reported context (profiles/r02_configs.json), not bench.py lines."""
import ctypes as C, io, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import plz4_b200 as P
from plz4_b200 import _lib
from plz4_b200._lib import check
from oracle import oracle as O
from tests.datagen import make

L = _lib.lib(); P.init(0)
ref, port = O.Ref(), O.Port()
drv = C.CDLL(os.path.join(os.path.dirname(O.__file__), "cpu_driver.so"))
ncpu = os.cpu_count()
vp = lambda a: C.c_void_p(a.ctypes.data)
fn = lambda f: C.cast(f, C.c_void_p)
res = {"host_cores": ncpu}

from plz4_b200 import stream as S

def c_compress(src_np, dst_np, **opts):
    """NewWriter over in-memory C endpoints (no Python in the data path); returns the frame length."""
    o, keep = S._opts(**opts)
    sink = L.plz4cu_membuf_new(vp(dst_np), 0, dst_np.size)
    w = L.plz4cu_writer_new(fn(L.plz4cu_membuf_write), sink, C.byref(o))
    r = L.plz4cu_writer_write(w, vp(src_np), src_np.size); assert r == src_np.size, r
    assert L.plz4cu_writer_close(w) == 0
    n = L.plz4cu_membuf_len(sink)
    L.plz4cu_writer_free(w); L.plz4cu_membuf_free(sink)
    return n

def c_decompress(frame_np, flen, dst_np, **opts):
    o, keep = S._opts(**opts)
    srcb = L.plz4cu_membuf_new(vp(frame_np), flen, flen)
    sink = L.plz4cu_membuf_new(vp(dst_np), 0, dst_np.size)
    r = L.plz4cu_reader_new(fn(L.plz4cu_membuf_read), fn(L.plz4cu_membuf_seek), srcb, C.byref(o))
    n = L.plz4cu_reader_write_to(r, fn(L.plz4cu_membuf_write), sink); assert n >= 0, n
    L.plz4cu_reader_close(r); L.plz4cu_reader_free(r); L.plz4cu_membuf_free(srcb); L.plz4cu_membuf_free(sink)
    return n

def logtext(n, seed=0x504C5A34):
    a = np.empty(n, dtype=np.uint8); L.plz4cu_gen_logtext_host(seed, 0, vp(a), n); return a

def best(f, reps=3):
    f(); t = []
    for _ in range(reps):
        t0 = time.perf_counter(); f(); t.append(time.perf_counter() - t0)
    return min(t)

def cpu_blocks(data, bsz, checksum=1):
    n = data.size; nblk = (n + bsz - 1) // bsz
    recs = np.empty(nblk * (bsz + 8), dtype=np.uint8); rl = np.zeros(nblk, dtype=np.uint32)
    out = np.empty(nblk * bsz, dtype=np.uint8); ol = np.zeros(nblk, dtype=np.uint32)
    roff = np.arange(nblk, dtype=np.uint64) * (bsz + 8)
    drv.drv_compress_blocks.argtypes = [C.c_void_p] * 3 + [C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    drv.drv_decompress_blocks.argtypes = [C.c_void_p] * 5 + [C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    tc = best(lambda: drv.drv_compress_blocks(fn(ref.lib.LZ4_compress_fast), fn(port.lib.orc_xxh32), vp(data), n, bsz, checksum, vp(recs), vp(rl), ncpu))
    td = best(lambda: drv.drv_decompress_blocks(fn(ref.lib.LZ4_decompress_safe), fn(port.lib.orc_xxh32), vp(recs), vp(roff), vp(rl), nblk, bsz, checksum, vp(out), vp(ol), ncpu))
    return n / tc / 1e9, n / td / 1e9, int(rl.sum()) / n

# ---- configs[0]: 256 MiB log text, 4 MiB blocks, block + content checksums, host-resident streams
data = logtext(256 << 20)
raw = data.tobytes()
fbuf = np.empty(data.size + (1 << 20), dtype=np.uint8); obuf = np.empty(data.size, dtype=np.uint8)
o0 = dict(block_size_idx=7, block_checksum=True, content_checksum=True, parallel=-1)
flen = c_compress(data, fbuf, **o0)
assert c_decompress(fbuf, flen, obuf, parallel=-1) == data.size and obuf.tobytes() == raw
tw = best(lambda: c_compress(data, fbuf, **o0), 4)
tr = best(lambda: c_decompress(fbuf, flen, obuf, parallel=-1), 4)
o0n = dict(o0, content_checksum=False)
flen_n = c_compress(data, fbuf, **o0n)
tw_n = best(lambda: c_compress(data, fbuf, **o0n), 4)
tr_n = best(lambda: c_decompress(fbuf, flen_n, obuf, parallel=-1), 4)
frame = bytes(flen)
cc, cd, cr = cpu_blocks(data, 4 << 20)
res["config0_256MiB_4MiB_blocks_bx_cx_streams"] = {
    "gpu_write_gbs": round(len(raw) / tw / 1e9, 2), "gpu_read_gbs": round(len(raw) / tr / 1e9, 2), "gpu_ratio": round(len(frame) / len(raw), 4),
    "gpu_write_gbs_no_content_checksum": round(len(raw) / tw_n / 1e9, 2), "gpu_read_gbs_no_content_checksum": round(len(raw) / tr_n / 1e9, 2),
    "cpu_compress_gbs": round(cc, 2), "cpu_decompress_gbs": round(cd, 2), "cpu_ratio": round(cr, 4),
    "note": "NewWriter/NewReader over in-memory C endpoints, pageable caller buffers; CPU columns are block-level (no stream layer, no content checksum); "
            "decode of 4 MiB blocks is one CTA per block (64 blocks here): DESIGN.md 4.1b"}

# ---- configs[2]: decode reference-produced frames (4 MiB blocks, bx) from 64 random WithReadOffset starts into a 1 GiB frame,
# 64 MiB read from every start; beside it the reference's liblz4 decoding the same blocks on all host cores
from oracle import frame_oracle as F
big1 = logtext(1 << 30, seed=0x1234)
sub = big1.tobytes()
marks = []
rframe = F.write_frame(sub, F.Opts(block_idx=7, block_checksum=True, content_checksum=False), port, progress=lambda s, d: marks.append((s, d)))
rng = np.random.default_rng(3)
picks = [marks[i] for i in rng.integers(0, len(marks) - 17, size=64)]
rf_np = np.frombuffer(rframe, dtype=np.uint8)
want = 64 << 20
ra_out = np.empty(want + (4 << 20), dtype=np.uint8)
def c_read_some(frame_np, dst_np, nbytes, **opts):
    o, keep = S._opts(**opts)
    srcb = L.plz4cu_membuf_new(vp(frame_np), frame_np.size, frame_np.size)
    r = L.plz4cu_reader_new(fn(L.plz4cu_membuf_read), fn(L.plz4cu_membuf_seek), srcb, C.byref(o))
    got = 0
    while got < nbytes:
        k = L.plz4cu_reader_read(r, C.c_void_p(dst_np.ctypes.data + got), nbytes - got)
        assert k >= 0, k
        if k == 0:
            break
        got += k
    L.plz4cu_reader_close(r); L.plz4cu_reader_free(r); L.plz4cu_membuf_free(srcb)
    return got
L.plz4cu_reader_read.restype = C.c_int64
L.plz4cu_reader_read.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
def ra():
    tot = 0
    for s_, d_ in picks:
        tot += c_read_some(rf_np, ra_out, want, read_offset=d_)
    return tot
tot = ra(); t3 = best(ra, 2)
assert ra_out[:want].tobytes() == sub[picks[-1][0]: picks[-1][0] + want]
_, cd2, _ = cpu_blocks(big1[: 256 << 20], 4 << 20)
res["config2_random_access_reference_frames_4MiB"] = {"starts": len(picks), "frame_bytes": len(rframe), "read_per_start": want, "decoded_bytes": tot,
                                                      "gpu_gbs": round(tot / t3 / 1e9, 2), "cpu_decompress_gbs_same_blocks_all_cores": round(cd2, 2),
                                                      "note": "NewReader(WithReadOffset) over a reference-written 1 GiB frame, pageable buffers; the CPU column decodes 4 MiB blocks of the same text on all host cores (block level, no stream layer)"}

# ---- configs[3]: 4 KiB payloads + 64 KiB dictionary, one batch call (device work + PCIe), vs liblz4 amortised dict ctx
corpus = logtext(64 << 20, seed=13)
nmsg = 1 << 18
starts = rng.integers(65536, corpus.size - 4096, size=nmsg).astype(np.uint64)
d = corpus[:65536].tobytes()
gd = P.Dict(d)
lens = np.full(nmsg, 4096, dtype=np.uint32)
hsrc = torch.from_numpy(corpus).pin_memory()
cap = P.compress_block_bound(4096)
packed = torch.empty(nmsg * (cap + 8), dtype=torch.uint8).pin_memory(); poff = np.zeros(nmsg + 1, dtype=np.uint64)
hp = lambda t: C.c_void_p(t.data_ptr())
comp = lambda: check(L.plz4cu_compress_batch_host(hp(hsrc), vp(starts), vp(lens), nmsg, cap, 0, 1, gd.handle, hp(packed), packed.numel(), vp(poff)))
t4c = best(comp)
sizes = np.diff(poff).astype(np.uint32)
out = torch.empty(nmsg * 4096, dtype=torch.uint8).pin_memory(); rs = np.zeros(nmsg, dtype=np.int32)
dec = lambda: check(L.plz4cu_decompress_batch_host(hp(packed), int(poff[nmsg]), vp(poff), vp(sizes), nmsg, 4096, 0, 1, gd.handle, hp(out), 4096, vp(rs)))
t4d = best(dec)
assert (rs == 4096).all()
rd_ = ref.dict_create(d)
dst = np.empty(nmsg * cap, dtype=np.uint8); ol = np.zeros(nmsg, dtype=np.uint32)
drv.drv_compress_dict_msgs.argtypes = [C.c_void_p] * 6 + [C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_int]
t4cpu = best(lambda: drv.drv_compress_dict_msgs(fn(ref.lib.LZ4_resetStream_fast), fn(ref.lib.LZ4_attach_dictionary), fn(ref.lib.LZ4_compress_fast_continue),
                                                C.c_void_p(rd_._sp), vp(corpus), vp(starts), 4096, nmsg, vp(dst), cap, vp(ol), ncpu))
res["config3_4KiB_payloads_64KiB_dict"] = {
    "messages": nmsg, "gpu_compress_gbs": round(nmsg * 4096 / t4c / 1e9, 2), "gpu_compress_Mmsg_s": round(nmsg / t4c / 1e6, 2),
    "gpu_decompress_gbs": round(nmsg * 4096 / t4d / 1e9, 2), "gpu_ratio": round(int(sizes.sum()) / (nmsg * 4096), 4),
    "cpu_compress_gbs_amortised_dict_ctx": round(nmsg * 4096 / t4cpu / 1e9, 2), "cpu_ratio": round(int(ol.sum()) / (nmsg * 4096), 4),
    "note": "host buffers, PCIe inside the timing; CPU figure keeps one dict ctx for all payloads (the reference rebuilds it per call)"}

# ---- configs[4] in miniature: mixed-entropy stream, 256 KiB blocks, block + content checksum, NewWriter / NewReader
seg = 1 << 20
kinds = ["log", "log", "random", "zeros", "record1025", "log", "record1025", "log", "random", "log"]
mixed = b"".join(make(kinds[i % 10], seg, seed=i) for i in range(256)) * 4          # 1 GiB: 256 distinct MiB, four times
mixed_np = np.frombuffer(mixed, dtype=np.uint8)
f5buf = np.empty(mixed_np.size + (1 << 20), dtype=np.uint8); o5buf = np.empty(mixed_np.size, dtype=np.uint8)
o5 = dict(block_size_idx=5, block_checksum=True, content_checksum=True)
f5len = c_compress(mixed_np, f5buf, **o5)
assert c_decompress(f5buf, f5len, o5buf) == mixed_np.size and o5buf.tobytes() == mixed
t5w = best(lambda: c_compress(mixed_np, f5buf, **o5), 2); t5r = best(lambda: c_decompress(f5buf, f5len, o5buf), 2)
o5n = dict(o5, content_checksum=False)
f5len_n = c_compress(mixed_np, f5buf, **o5n)
t5w_n = best(lambda: c_compress(mixed_np, f5buf, **o5n), 2); t5r_n = best(lambda: c_decompress(f5buf, f5len_n, o5buf), 2)
f5 = bytes(f5len)
cc5, cd5, cr5 = cpu_blocks(np.frombuffer(mixed, dtype=np.uint8), 256 << 10)
res["config4_mixed_1GiB_256KiB_blocks_bx_cx_streams"] = {
    "gpu_write_gbs": round(len(mixed) / t5w / 1e9, 2), "gpu_read_gbs": round(len(mixed) / t5r / 1e9, 2), "gpu_ratio": round(len(f5) / len(mixed), 4),
    "gpu_write_gbs_no_content_checksum": round(len(mixed) / t5w_n / 1e9, 2), "gpu_read_gbs_no_content_checksum": round(len(mixed) / t5r_n / 1e9, 2),
    "cpu_compress_gbs": round(cc5, 2), "cpu_decompress_gbs": round(cd5, 2), "cpu_ratio": round(cr5, 4)}

# ---- what one host thread can do on this box: the ceilings of any stream layer that copies caller bytes once
# (staging in, result out) and hashes the content serially
big = mixed_np[: 256 << 20]; tmp = np.empty_like(big)
L.plz4cu_xxh32_host.restype = C.c_uint32
res["host_single_thread_bounds"] = {
    "memcpy_gbs": round(big.size / best(lambda: np.copyto(tmp, big)) / 1e9, 2),
    "xxh32_gbs": round(big.size / best(lambda: L.plz4cu_xxh32_host(vp(big), big.size)) / 1e9, 2),
    "note": "content checksum (xxh32 of the uncompressed stream) is serial by definition: streams with it cannot exceed xxh32_gbs"}
print(json.dumps(res, indent=1))
