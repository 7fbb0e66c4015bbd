"""Helpers shared by the stream probes: NewWriter / NewReader over in-memory C endpoints (no Python in the data path)."""
import ctypes as C, time
from plz4_b200 import _lib, stream as S
L = _lib.lib()
vp = lambda a: C.c_void_p(a.ctypes.data); hp = lambda x: C.c_void_p(x.data_ptr()); fn = lambda f: C.cast(f, C.c_void_p)

def best(f, reps=3):
    f(); t = []
    for _ in range(reps):
        t0 = time.perf_counter(); f(); t.append(time.perf_counter() - t0)
    return min(t)

def c_compress(src_np, dst_np, chunk=0, **opts):
    o, keep = S._opts(**opts)
    sink = L.plz4cu_membuf_new(vp(dst_np), 0, dst_np.size)
    w = L.plz4cu_writer_new(fn(L.plz4cu_membuf_write), sink, C.byref(o))
    if chunk:
        base = src_np.ctypes.data
        for o_ in range(0, src_np.size, chunk):
            k = min(chunk, src_np.size - o_)
            r = L.plz4cu_writer_write(w, C.c_void_p(base + o_), k); assert r == k, r
    else:
        r = L.plz4cu_writer_write(w, vp(src_np), src_np.size); assert r == src_np.size, r
    assert L.plz4cu_writer_close(w) == 0
    m = L.plz4cu_membuf_len(sink)
    L.plz4cu_writer_free(w); L.plz4cu_membuf_free(sink)
    return m

def c_decompress(frame_np, flen, dst_np, **opts):
    o, keep = S._opts(**opts)
    srcb = L.plz4cu_membuf_new(vp(frame_np), flen, flen)
    sink = L.plz4cu_membuf_new(vp(dst_np), 0, dst_np.size)
    r = L.plz4cu_reader_new(fn(L.plz4cu_membuf_read), fn(L.plz4cu_membuf_seek), srcb, C.byref(o))
    m = L.plz4cu_reader_write_to(r, fn(L.plz4cu_membuf_write), sink); assert m >= 0, m
    L.plz4cu_reader_close(r); L.plz4cu_reader_free(r); L.plz4cu_membuf_free(srcb); L.plz4cu_membuf_free(sink)
    return m

