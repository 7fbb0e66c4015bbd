// host_stream.cu — host side of the path: LZ4 frame Writer / Reader on top of the batched GPU engine.
//
// A C++ mirror of the Go host code that surrounds plz4's block engine (this image has no Go toolchain;
// INTEGRATION.md shows how the same batching slots into the Go writer/reader).  Reference being mirrored:
//   header   internal/pkg/header/write.go:23-73, read.go:26-119, skip.go:18-76, descriptor/*.go, trailer/trailer.go:10-19
//   writer   internal/pkg/sync/writer.go:53-290 and internal/pkg/async/writer.go:81-191,284-381 (ordering, marks,
//            progress callback, Flush barrier, sticky error + "reported" flag)
//   reader   internal/pkg/rdr/rdr.go:39-366 (header/body modes, deferred errors, ReadOffset, frame concatenation,
//            content-size check), internal/pkg/blk/frame.go:54-139 (size word walk), async/reader.go:223-271
//   xxh32    internal/pkg/xxh32/xxh32zero.go:22-235 (streaming form, for the serial content checksum)
// What changes: instead of one goroutine per block, whole batches of blocks go through
// plz4cu_compress_batch_host / plz4cu_decompress_batch_host; block checksums are made / verified on the GPU.
#include "../../include/plz4cu.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <future>
#include <string>
#include <vector>

namespace {

// ---------------------------------------------------------------- streaming xxh32, seed 0

constexpr uint32_t P1 = 2654435761u, P2 = 2246822519u, P3 = 3266489917u, P4 = 668265263u, P5 = 374761393u;
inline uint32_t rol(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }

struct XXH32 {
    uint32_t v[4];
    uint64_t total = 0;
    uint8_t buf[16];
    int nbuf = 0;
    XXH32() { reset(); }
    void reset() { v[0] = P1 + P2; v[1] = P2; v[2] = 0; v[3] = 0u - P1; total = 0; nbuf = 0; }
    void stripe(const uint8_t* p)
    {
        for (int i = 0; i < 4; i++) v[i] = rol(v[i] + rd32(p + 4 * i) * P2, 13) * P1;
    }
    void update(const uint8_t* p, size_t n)
    {
        total += n;
        if (nbuf) {
            size_t take = std::min<size_t>(16 - nbuf, n);
            memcpy(buf + nbuf, p, take);
            nbuf += (int)take; p += take; n -= take;
            if (nbuf < 16) return;
            stripe(buf);
            nbuf = 0;
        }
        for (; n >= 16; p += 16, n -= 16) stripe(p);
        if (n) { memcpy(buf, p, n); nbuf = (int)n; }
    }
    uint32_t digest() const
    {
        uint32_t h = total >= 16 ? rol(v[0], 1) + rol(v[1], 7) + rol(v[2], 12) + rol(v[3], 18) : P5;
        h += (uint32_t)total;
        int i = 0;
        for (; i + 4 <= nbuf; i += 4) h = rol(h + rd32(buf + i) * P3, 17) * P4;
        for (; i < nbuf; i++) h = rol(h + buf[i] * P5, 11) * P1;
        h ^= h >> 15; h *= P2; h ^= h >> 13; h *= P3; h ^= h >> 16;
        return h;
    }
};

uint32_t xxh32_once(const void* p, size_t n)
{
    XXH32 x;
    x.update(static_cast<const uint8_t*>(p), n);
    return x.digest();
}

// ---------------------------------------------------------------- frame constants

const uint8_t kMagic[4] = {0x04, 0x22, 0x4d, 0x18};
constexpr uint32_t kSkipMagic = 0x184D2A50u;
constexpr size_t kAutoBatchBytes = 256u << 20;       // blocks gathered per engine call when batching (n_parallel != 0)

int block_size_of(int idx)
{
    switch (idx) {                                    // descriptor/index.go:26-38
    case 4: return 64 << 10;
    case 5: return 256 << 10;
    case 6: return 1 << 20;
    case 7: return 4 << 20;
    }
    return 0;
}

inline void put32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }
inline uint32_t get32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

// Grow-only pinned staging (blk.BorrowBlk's job in the GPU build): no zero-fill, DMA-able, counted by
// plz4cu_host_outstanding() so the leak discipline of the reference's tests (wr_test.go:29-33) still has a gauge.
struct PinnedBuf {
    uint8_t* p = nullptr;
    size_t cap = 0;
    bool pinned = false;
    ~PinnedBuf() { release(); }
    void release()
    {
        if (p) { if (pinned) plz4cu_host_free(p); else free(p); }
        p = nullptr; cap = 0;
    }
    bool reserve(size_t n)
    {
        if (n <= cap) return true;
        release();
        // small streams are not worth pinning; large ones borrow a pooled pinned slab
        pinned = n >= (4u << 20);
        p = static_cast<uint8_t*>(pinned ? plz4cu_host_alloc(n) : malloc(std::max<size_t>(n, 64)));
        if (!p) return false;
        cap = n;
        return true;
    }
};

struct Opts {
    plz4cu_opts_t o;
    std::vector<uint8_t> dict;                        // owned copy (the caller's buffer need not outlive the call)
    explicit Opts(const plz4cu_opts_t* in)
    {
        if (in) o = *in; else plz4cu_opts_default(&o);
        if (o.block_size_idx < 4 || o.block_size_idx > 7) o.block_size_idx = 7;   // plz4_opts.go:160-168
        if (o.level < 1) o.level = 1;
        if (o.level > 12) o.level = 12;
        if (o.dict && o.dict_len) dict.assign(static_cast<const uint8_t*>(o.dict), static_cast<const uint8_t*>(o.dict) + o.dict_len);
        o.dict = nullptr;
    }
    size_t batch_bytes(int bsz) const
    {
        if (o.n_parallel == 0) return (size_t)bsz;                                // synchronous flavour: one block per call
        size_t want = o.pending_size > 0 ? (size_t)o.pending_size : kAutoBatchBytes;
        want = std::max<size_t>(want, (size_t)bsz);
        return want / bsz * bsz;
    }
};

// header/write.go:23-73
std::vector<uint8_t> make_header(const plz4cu_opts_t& o)
{
    std::vector<uint8_t> h(kMagic, kMagic + 4);
    uint8_t flags = 1u << 6;
    if (!o.block_linked) flags |= 1u << 5;
    if (o.block_checksum) flags |= 1u << 4;
    if (o.content_checksum) flags |= 1u << 2;
    if (o.has_content_size) flags |= 1u << 3;
    if (o.has_dict_id) flags |= 1u << 0;
    h.push_back(flags);
    h.push_back((uint8_t)((o.block_size_idx & 7) << 4));
    if (o.has_content_size) for (int i = 0; i < 8; i++) h.push_back((uint8_t)(o.content_size >> (8 * i)));
    if (o.has_dict_id) for (int i = 0; i < 4; i++) h.push_back((uint8_t)(o.dict_id >> (8 * i)));
    h.push_back((uint8_t)((xxh32_once(h.data() + 4, h.size() - 4) >> 8) & 0xFF));
    return h;
}

}  // namespace

// ================================================================ Writer

struct plz4cu_writer {
    plz4cu_write_fn wr;
    void* wr_ctx;
    Opts opt;
    int bsz;
    size_t batch;
    bool header_written = false, closed = false, reported = false;
    int state = 0;                                    // sticky error (first error wins, async/writer.go:553-555)
    int64_t src_mark = 0, dst_mark = 0;
    XXH32 hasher;
    plz4cu_dict_t* dict = nullptr;
    // staging: pageable while small, one pinned slab once a stream proves to be large
    std::vector<uint8_t> small;
    uint8_t* slab = nullptr;
    size_t fill = 0;
    PinnedBuf packed;
    std::vector<uint64_t> offs, poff;
    std::vector<uint32_t> lens;

    plz4cu_writer(plz4cu_write_fn w, void* c, const plz4cu_opts_t* o) : wr(w), wr_ctx(c), opt(o)
    {
        bsz = block_size_of(opt.o.block_size_idx);
        batch = opt.batch_bytes(bsz);
        if (opt.o.level != 1 || opt.o.block_linked) state = PLZ4CU_Z_UNSUPPORTED;
        if (!opt.dict.empty() && state == 0) {
            dict = plz4cu_dict_create(opt.dict.data(), opt.dict.size());
            if (!dict) state = PLZ4CU_Z_ENGINE;
        }
    }
    ~plz4cu_writer()
    {
        if (slab) plz4cu_host_free(slab);
        if (dict) plz4cu_dict_destroy(dict);
    }
    int report() { if (state) reported = true; return state; }
    void set_error(int e) { if (!state) state = e; }

    uint8_t* stage_ptr() { return slab ? slab : small.data(); }
    bool reserve(size_t want)
    {
        if (slab) return true;
        if (want <= (8u << 20) || batch <= (8u << 20)) {
            if (small.size() < want) small.resize(std::max(want, small.size() * 2));
            return true;
        }
        slab = static_cast<uint8_t*>(plz4cu_host_alloc(batch + 16));
        if (!slab) return false;
        memcpy(slab, small.data(), fill);
        small.clear(); small.shrink_to_fit();
        return true;
    }

    int write_all(const uint8_t* p, size_t n, int err_code)
    {
        if (n == 0) return 0;
        int64_t r = wr(wr_ctx, p, n);
        if (r < 0 || (size_t)r != n) return err_code;
        return 0;
    }
    int ensure_header()
    {
        if (header_written) return 0;
        std::vector<uint8_t> h = make_header(opt.o);
        if (int e = write_all(h.data(), h.size(), PLZ4CU_Z_HEADER_WRITE)) return e;
        dst_mark = (int64_t)h.size();
        header_written = true;
        return 0;
    }

    // Compress data[0..n) as consecutive bsz-sized blocks (the last may be short) and write them in order.
    int emit(const uint8_t* data, size_t n)
    {
        if (n == 0) return 0;
        if (int e = ensure_header()) return e;
        const uint32_t nblk = (uint32_t)((n + bsz - 1) / bsz);
        offs.resize(nblk); lens.resize(nblk); poff.resize(nblk + 1);
        for (uint32_t i = 0; i < nblk; i++) { offs[i] = (uint64_t)i * bsz; lens[i] = (uint32_t)std::min<size_t>(bsz, n - offs[i]); }
        const size_t packed_cap = (size_t)nblk * (bsz + 8);
        if (!packed.reserve(packed_cap)) return PLZ4CU_Z_ENGINE;
        // the serial content checksum runs on a host core while the GPU works (async/hash.go)
        std::future<void> hf;
        if (opt.o.content_checksum) hf = std::async(std::launch::async, [&] { hasher.update(data, n); });
        int rc = plz4cu_compress_batch_host(data, offs.data(), lens.data(), nblk, (uint32_t)bsz, opt.o.block_checksum, 0, dict,
                                            packed.p, packed_cap, poff.data());
        if (hf.valid()) hf.get();
        if (rc < 0) return PLZ4CU_Z_ENGINE;
        if (!opt.o.progress) {
            // nobody watches block boundaries: one write for the whole batch
            if (int e = write_all(packed.p, (size_t)poff[nblk], PLZ4CU_Z_WRITE)) return e;
            src_mark += (int64_t)n; dst_mark += (int64_t)poff[nblk];
            return 0;
        }
        for (uint32_t i = 0; i < nblk; i++) {
            const size_t len = (size_t)(poff[i + 1] - poff[i]);
            int e = write_all(packed.p + poff[i], len, PLZ4CU_Z_WRITE);
            opt.o.progress(opt.o.progress_ctx, src_mark, dst_mark);        // async/writer.go:327-331
            src_mark += lens[i]; dst_mark += (int64_t)len;
            if (e) return e;
        }
        return 0;
    }

    int64_t write(const uint8_t* src, size_t n)
    {
        if (state) return report();
        size_t done = 0;
        while (done < n && !state) {
            if (fill == 0 && n - done >= batch) {
                // large caller buffer: compress whole batches in place, no staging copy (sync/writer.go:99-109)
                size_t take = (n - done) / batch * batch;
                take = std::min(take, batch);
                if (int e = emit(src + done, take)) set_error(e);
                done += take;
                continue;
            }
            size_t take = std::min(n - done, batch - fill);
            if (!reserve(fill + take)) { set_error(PLZ4CU_Z_ENGINE); break; }
            memcpy(stage_ptr() + fill, src + done, take);
            fill += take; done += take;
            if (fill == batch) {
                if (int e = emit(stage_ptr(), fill)) set_error(e);
                fill = 0;
            }
        }
        if (state) return report();
        return (int64_t)done;
    }
    int flush_pending()
    {
        if (fill == 0) return 0;
        int e = emit(stage_ptr(), fill);
        fill = 0;
        return e;
    }
    int flush()
    {
        if (state) return report();
        if (int e = flush_pending()) set_error(e);
        return report();
    }
    int close()
    {
        if (closed) return report();
        if (!state) {
            int e = flush_pending();
            if (!e) e = ensure_header();
            if (!e) {
                if (opt.o.progress) opt.o.progress(opt.o.progress_ctx, src_mark, dst_mark);   // async/writer.go:368
                uint8_t t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                size_t tn = 4;
                if (opt.o.content_checksum) { put32(t + 4, hasher.digest()); tn = 8; }       // trailer/trailer.go:10-19
                e = write_all(t, tn, PLZ4CU_Z_WRITE);
            }
            if (e) set_error(e);
        }
        closed = true;
        int ret;
        if (reported) ret = 0;                         // already surfaced once: Close succeeds (async/writer.go:175-190)
        else if (!state) { state = PLZ4CU_Z_CLOSED; reported = true; ret = 0; }
        else ret = report();
        return ret;
    }
};

// ================================================================ Reader

struct plz4cu_reader {
    plz4cu_read_fn rd;
    plz4cu_seek_fn seek;
    void* ctx;
    Opts opt;
    bool closed = false;
    int state = 0;                                    // sticky error; 1 = clean end of stream
    int64_t src_pos = 0, dst_pos = 0;
    int64_t read_offset;
    bool skip_content_size;
    plz4cu_dict_t* dict = nullptr;
    std::vector<uint8_t> cur_dict;

    // current frame
    bool in_body = false;
    int bsz = 0;
    bool blk_check = false, has_content_hash = false, verify_content_hash = false;
    bool has_content_size = false;
    uint64_t hdr_content_size = 0, content_acc = 0;
    XXH32 hasher;

    // current batch of decoded blocks
    PinnedBuf recs, out;
    size_t recs_len = 0;
    std::vector<uint64_t> rec_off;
    std::vector<uint32_t> rec_read;                   // input bytes each block consumed (size word + body + hash)
    std::vector<int32_t> out_len;
    uint32_t nblk = 0, cur = 0;
    size_t cur_off = 0, cur_len = 0;
    bool have_block = false;
    int tail_event = 0;                               // after the batch: 0 nothing, 2 EndMark, <0 error
    uint32_t endmark_read = 0;                        // bytes the EndMark (+ content hash) consumed
    uint32_t content_hash_read = 0;
    uint32_t tail_read = 0;                           // bytes consumed by a failed trailing read (for src_pos)

    plz4cu_reader(plz4cu_read_fn r, plz4cu_seek_fn s, void* c, const plz4cu_opts_t* o) : rd(r), seek(s), ctx(c), opt(o)
    {
        read_offset = opt.o.read_offset;
        skip_content_size = !opt.o.content_size_check;
        cur_dict = opt.dict;
    }
    ~plz4cu_reader() { if (dict) plz4cu_dict_destroy(dict); }

    // io.ReadFull: 0 = ok, 1 = clean EOF before any byte, -1 = short / error
    int read_full(uint8_t* p, size_t n, size_t* got)
    {
        size_t g = 0;
        while (g < n) {
            int64_t r = rd(ctx, p + g, n - g);
            if (r <= 0) break;
            g += (size_t)r;
        }
        *got = g;
        if (g == n) return 0;
        return g == 0 ? 1 : -1;
    }

    // header/read.go:26-119 + rdr/rdr.go:242-296.  Returns 0 ok, 1 clean EOF, <0 error.
    int read_header()
    {
        for (;;) {
            uint8_t h[19];
            size_t got = 0;
            int r = read_full(h, 7, &got);
            src_pos += (int64_t)got;
            if (r == 1) return 1;
            if (r < 0) return PLZ4CU_Z_HEADER_READ;
            if (memcmp(h, kMagic, 4) != 0) {
                // header/skip.go:38-76
                uint32_t m = get32(h);
                if ((m >> 4) != (kSkipMagic >> 4)) return PLZ4CU_Z_MAGIC;
                r = read_full(h + 7, 1, &got);
                src_pos += (int64_t)got;
                if (r != 0) return PLZ4CU_Z_HEADER_READ;
                uint32_t sz = get32(h + 4);
                std::vector<uint8_t> payload(sz);
                r = read_full(payload.data(), sz, &got);
                src_pos += (int64_t)got;
                if (r != 0 && sz) return PLZ4CU_Z_SKIP;
                if (opt.o.skip_cb && opt.o.skip_cb(opt.o.skip_ctx, (uint8_t)(m & 0xF), payload.data(), sz) != 0) return PLZ4CU_Z_SKIP;
                continue;                               // a skipped frame puts the reader back in header mode
            }
            const uint8_t flags = h[4], bd = h[5];
            if (((flags >> 6) & 3) != 1) return PLZ4CU_Z_VERSION;
            if (flags & 0x02) return PLZ4CU_Z_RESERVE_BIT;
            if (((bd >> 4) & 7) < 4 || (bd & 0x80) || (bd & 0x0F)) return PLZ4CU_Z_BLOCK_DESCRIPTOR;
            size_t n = 7;
            uint64_t csz = 0;
            uint32_t did = 0;
            if (flags & 0x08) {
                r = read_full(h + 7, 8, &got);
                src_pos += (int64_t)got;
                if (r != 0) return PLZ4CU_Z_HEADER_READ;
                for (int i = 0; i < 8; i++) csz |= (uint64_t)h[6 + i] << (8 * i);
                n = 15;
            }
            if (flags & 0x01) {
                r = read_full(h + n, 4, &got);
                src_pos += (int64_t)got;
                if (r != 0) return PLZ4CU_Z_HEADER_READ;
                did = get32(h + n - 1);
                n += 4;
            }
            if (((xxh32_once(h + 4, n - 5) >> 8) & 0xFF) != h[n - 1]) return PLZ4CU_Z_HEADER_HASH;

            if ((flags & 0x01) && opt.o.dict_cb) {       // rdr/rdr.go:254-259
                const void* dp = nullptr; size_t dl = 0;
                if (opt.o.dict_cb(opt.o.dict_ctx, did, &dp, &dl) != 0) return PLZ4CU_Z_HEADER_READ;
                if (dp) { cur_dict.assign(static_cast<const uint8_t*>(dp), static_cast<const uint8_t*>(dp) + dl); if (dict) { plz4cu_dict_destroy(dict); dict = nullptr; } }
            }
            const bool independent = (flags & 0x20) != 0;
            bool check_hash = (flags & 0x04) != 0 && opt.o.content_checksum;
            if (read_offset != 0 && read_offset != (int64_t)n) {   // rdr/rdr.go:261-285
                if (read_offset < (int64_t)n) return PLZ4CU_Z_READ_OFFSET;
                if (!independent) return PLZ4CU_Z_READ_OFFSET_LINKED;
                int64_t skip = read_offset - (int64_t)n;
                if (seek) {
                    if (seek(ctx, skip) != 0) return PLZ4CU_Z_READ_OFFSET;
                } else {
                    std::vector<uint8_t> junk(1 << 16);
                    int64_t left = skip;
                    while (left > 0) {
                        size_t g = 0;
                        int rr = read_full(junk.data(), (size_t)std::min<int64_t>(left, (int64_t)junk.size()), &g);
                        left -= (int64_t)g;
                        if (rr != 0) return PLZ4CU_Z_READ_OFFSET;
                    }
                }
                src_pos += skip;
                check_hash = false;                     // the checksum covers bytes we skipped
                skip_content_size = true;
            }
            read_offset = 0;                            // applies to the first frame only
            if (!independent) return PLZ4CU_Z_UNSUPPORTED;   // linked frames are decoded on CPU cores by the reference
            if (!cur_dict.empty() && !dict) {
                dict = plz4cu_dict_create(cur_dict.data(), cur_dict.size());
                if (!dict) return PLZ4CU_Z_ENGINE;
            }
            bsz = block_size_of((bd >> 4) & 7);
            blk_check = (flags & 0x10) != 0;
            has_content_hash = (flags & 0x04) != 0;
            verify_content_hash = check_hash;
            has_content_size = (flags & 0x08) != 0;
            hdr_content_size = csz;
            content_acc = 0;
            hasher.reset();
            in_body = true;
            return 0;
        }
    }

    // blk/frame.go:54-112 for up to a batch of blocks, then one engine call.
    void fill_batch()
    {
        const size_t batch_blocks = std::max<size_t>(1, opt.batch_bytes(bsz) / (size_t)bsz);
        recs_len = 0; rec_off.clear(); rec_read.clear();
        nblk = 0; cur = 0; tail_event = 0; tail_read = 0;
        if (!recs.reserve(batch_blocks * ((size_t)bsz + 8))) { tail_event = PLZ4CU_Z_ENGINE; return; }
        while (nblk < batch_blocks) {
            uint8_t w[4];
            size_t got = 0;
            int r = read_full(w, 4, &got);
            if (r != 0) { tail_event = PLZ4CU_Z_BLOCK_SIZE_READ; tail_read = (uint32_t)got; break; }
            uint32_t word = get32(w);
            if (word == 0) {                            // EndMark (+ content checksum)
                endmark_read = 4;
                tail_event = 2;
                if (has_content_hash) {
                    uint8_t c[4];
                    r = read_full(c, 4, &got);
                    endmark_read += (uint32_t)got;
                    if (r != 0) { tail_event = PLZ4CU_Z_CONTENT_HASH_READ; tail_read = endmark_read; break; }
                    content_hash_read = get32(c);
                }
                break;
            }
            uint32_t n = word & 0x7FFFFFFFu;
            if (n > (uint32_t)bsz) { tail_event = PLZ4CU_Z_BLOCK_SIZE_OVERFLOW; tail_read = 4; break; }
            const size_t body = (size_t)n + (blk_check ? 4 : 0);
            const size_t at = recs_len;
            memcpy(recs.p + at, w, 4);
            r = read_full(recs.p + at + 4, body, &got);
            if (r != 0) { tail_event = PLZ4CU_Z_BLOCK_READ; tail_read = 4 + (uint32_t)got; break; }
            recs_len = at + 4 + body;
            rec_off.push_back(at);
            rec_read.push_back((uint32_t)(4 + body));
            nblk++;
        }
        if (nblk) {
            out_len.resize(nblk);
            int rc = out.reserve((size_t)nblk * bsz) ? 0 : -1;
            if (rc == 0) rc = plz4cu_decompress_batch_host(recs.p, recs_len, rec_off.data(), nullptr, nblk, (uint32_t)bsz, blk_check, 0,
                                                           dict, out.p, (uint64_t)bsz, out_len.data());
            if (rc < 0) { nblk = 0; tail_event = PLZ4CU_Z_ENGINE; }
        }
    }

    // rdr/rdr.go:207-227 nextBlock: 0 = a block is current, 2 = EndMark, <0 error
    int next_block()
    {
        have_block = false;
        if (cur >= nblk && tail_event == 0) fill_batch();
        if (opt.o.progress) opt.o.progress(opt.o.progress_ctx, src_pos, dst_pos);
        if (cur < nblk) {
            const int32_t r = out_len[cur];
            src_pos += rec_read[cur];
            if (r < 0) {
                nblk = 0;
                if (r == PLZ4CU_E_BLOCKHASH) return PLZ4CU_Z_BLOCK_HASH;
                if (r == PLZ4CU_E_OVERFLOW) return PLZ4CU_Z_BLOCK_SIZE_OVERFLOW;
                return PLZ4CU_Z_DECOMPRESS;
            }
            cur_off = 0; cur_len = (size_t)r;
            dst_pos += r; content_acc += (uint64_t)r;
            if (verify_content_hash) hasher.update(out.p + (size_t)cur * bsz, (size_t)r);
            have_block = true;
            cur++;
            return 0;
        }
        const int ev = tail_event;
        tail_event = 0;
        if (ev == 2) {
            src_pos += endmark_read;
            if (verify_content_hash && hasher.digest() != content_hash_read) return PLZ4CU_Z_CONTENT_HASH;
            return 2;
        }
        src_pos += tail_read;
        return ev;
    }
    const uint8_t* block_ptr() const { return out.p + (size_t)(cur - 1) * bsz; }

    // rdr/rdr.go:91-101
    int handle_end_mark()
    {
        int e = 0;
        if (has_content_size && !skip_content_size && hdr_content_size != content_acc) e = PLZ4CU_Z_CONTENT_SIZE;
        in_body = false;
        have_block = false;
        return e;
    }

    int64_t read(uint8_t* dst, size_t n)
    {
        if (state) return state == 1 ? 0 : state;
        size_t produced = 0;
        for (;;) {
            if (!in_body) {
                int r = read_header();
                if (r == 1) { state = 1; return (int64_t)produced; }          // io.EOF
                if (r < 0) { state = r; return produced ? (int64_t)produced : r; }
            }
            int err = 0;
            for (;;) {
                if (have_block && cur_off < cur_len) {
                    size_t k = std::min(n - produced, cur_len - cur_off);
                    memcpy(dst + produced, block_ptr() + cur_off, k);
                    cur_off += k; produced += k;
                    if (produced == n) return (int64_t)produced;
                }
                err = next_block();
                if (err) break;
            }
            if (err == 2) {
                int e = handle_end_mark();
                if (e) { state = e; return produced ? (int64_t)produced : e; }
                if (produced == 0 && n > 0) continue;     // never return (0, nil) at a frame boundary (rdr/rdr.go:61-64)
                return (int64_t)produced;
            }
            // defer the error when some data was produced (rdr/rdr.go:66-75)
            state = err;
            return produced ? (int64_t)produced : err;
        }
    }

    int64_t write_to(plz4cu_write_fn w, void* wctx)
    {
        int64_t sum = 0;
        while (state == 0) {
            if (!in_body) {
                int r = read_header();
                if (r == 1) break;                        // io.EOF on a header boundary ends WriteTo cleanly
                if (r < 0) { state = r; break; }
            }
            int err = 0;
            for (;;) {
                if (have_block && cur_off < cur_len) {
                    int64_t k = w(wctx, block_ptr() + cur_off, cur_len - cur_off);
                    if (k > 0) { cur_off += (size_t)k; sum += k; }
                    if (k < 0 || cur_off < cur_len) { err = PLZ4CU_Z_WRITE; break; }
                }
                err = next_block();
                if (err) break;
            }
            if (err == 2) { int e = handle_end_mark(); if (e) state = e; }
            else state = err;
        }
        return (state && state != 1) ? (int64_t)state : sum;
    }
    int close()
    {
        if (closed) return state;                       // rdr/rdr.go:109-112: a second Close reports the state (ErrClosed)
        closed = true;
        if (state == 0 || state == 1) state = PLZ4CU_Z_CLOSED;
        return 0;
    }
};

// ================================================================ C ABI

extern "C" {

void plz4cu_opts_default(plz4cu_opts_t* o)
{
    memset(o, 0, sizeof *o);
    o->level = 1;
    o->n_parallel = 1;
    o->block_size_idx = 7;
    o->content_checksum = 1;
    o->content_size_check = 1;
}

int plz4cu_err_corrupted(int code)
{
    switch (code) {
    case PLZ4CU_Z_HEADER_HASH: case PLZ4CU_Z_BLOCK_HASH: case PLZ4CU_Z_CONTENT_HASH: case PLZ4CU_Z_MAGIC:
    case PLZ4CU_Z_BLOCK_SIZE_OVERFLOW: case PLZ4CU_Z_DECOMPRESS: case PLZ4CU_Z_RESERVE_BIT:
    case PLZ4CU_Z_BLOCK_DESCRIPTOR: case PLZ4CU_Z_CONTENT_SIZE:
        return 1;
    }
    return 0;
}

const char* plz4cu_strerror(int code)
{
    switch (code) {                                   // zerr/zerr.go:11-36
    case 0: return "ok";
    case PLZ4CU_Z_CLOSED: return "lz4 closed";
    case PLZ4CU_Z_HEADER_HASH: return "lz4 corrupted: lz4 header hash mismatch";
    case PLZ4CU_Z_BLOCK_HASH: return "lz4 corrupted: lz4 block hash mismatch";
    case PLZ4CU_Z_CONTENT_HASH: return "lz4 corrupted: lz4 content hash mismatch";
    case PLZ4CU_Z_HEADER_READ: return "lz4 fail read header";
    case PLZ4CU_Z_HEADER_WRITE: return "lz4 fail write header";
    case PLZ4CU_Z_MAGIC: return "lz4 corrupted: lz4 bad magic";
    case PLZ4CU_Z_VERSION: return "lz4 unsupported version";
    case PLZ4CU_Z_BLOCK_SIZE_READ: return "lz4 fail read block size";
    case PLZ4CU_Z_BLOCK_READ: return "lz4 fail read block";
    case PLZ4CU_Z_BLOCK_SIZE_OVERFLOW: return "lz4 corrupted: lz4 block size overflow";
    case PLZ4CU_Z_DECOMPRESS: return "lz4 corrupted: lz4 fail decompress";
    case PLZ4CU_Z_RESERVE_BIT: return "lz4 corrupted: lz4 reserved bit set";
    case PLZ4CU_Z_BLOCK_DESCRIPTOR: return "lz4 corrupted: lz4 invalid BD byte";
    case PLZ4CU_Z_CONTENT_HASH_READ: return "lz4 fail read content hash";
    case PLZ4CU_Z_CONTENT_SIZE: return "lz4 corrupted: lz4 content size mismatch";
    case PLZ4CU_Z_READ_OFFSET: return "lz4 bad read offset";
    case PLZ4CU_Z_READ_OFFSET_LINKED: return "lz4 read offset unsupported in block linked mode";
    case PLZ4CU_Z_SKIP: return "lz4 fail skip";
    case PLZ4CU_Z_NIBBLE: return "lz4 bad nibble";
    case PLZ4CU_Z_UNSUPPORTED: return "lz4 unsupported feature";
    case PLZ4CU_Z_WRITE: return "write callback failed";
    case PLZ4CU_Z_ENGINE: return "plz4cu engine failure";
    }
    return "unknown";
}

plz4cu_writer_t* plz4cu_writer_new(plz4cu_write_fn wr, void* wr_ctx, const plz4cu_opts_t* opts)
{
    if (!wr) return nullptr;
    return new plz4cu_writer(wr, wr_ctx, opts);
}
int64_t plz4cu_writer_write(plz4cu_writer_t* w, const void* src, size_t n) { return w->write(static_cast<const uint8_t*>(src), n); }
int64_t plz4cu_writer_read_from(plz4cu_writer_t* w, plz4cu_read_fn rd, void* rd_ctx)
{
    if (w->state) return w->report();
    std::vector<uint8_t> buf(1 << 20);
    int64_t total = 0;
    for (;;) {
        int64_t r = rd(rd_ctx, buf.data(), buf.size());
        if (r < 0) { w->set_error(PLZ4CU_Z_BLOCK_READ); return w->report(); }
        if (r == 0) break;
        int64_t k = w->write(buf.data(), (size_t)r);
        if (k < 0) return k;
        total += k;
    }
    return total;
}
int plz4cu_writer_flush(plz4cu_writer_t* w) { return w->flush(); }
int plz4cu_writer_close(plz4cu_writer_t* w) { return w->close(); }
void plz4cu_writer_free(plz4cu_writer_t* w) { delete w; }

plz4cu_reader_t* plz4cu_reader_new(plz4cu_read_fn rd, plz4cu_seek_fn seek, void* rd_ctx, const plz4cu_opts_t* opts)
{
    if (!rd) return nullptr;
    return new plz4cu_reader(rd, seek, rd_ctx, opts);
}
int64_t plz4cu_reader_read(plz4cu_reader_t* r, void* dst, size_t n) { return r->read(static_cast<uint8_t*>(dst), n); }
int64_t plz4cu_reader_write_to(plz4cu_reader_t* r, plz4cu_write_fn wr, void* wr_ctx) { return r->write_to(wr, wr_ctx); }
int plz4cu_reader_close(plz4cu_reader_t* r) { return r->close(); }
void plz4cu_reader_free(plz4cu_reader_t* r) { delete r; }

int plz4cu_write_skip_frame_header(plz4cu_write_fn wr, void* wr_ctx, uint8_t nibble, uint32_t sz)
{
    if (nibble > 0xF) return PLZ4CU_Z_NIBBLE;
    uint8_t p[8];
    put32(p, kSkipMagic | nibble);
    put32(p + 4, sz);
    int64_t r = wr(wr_ctx, p, 8);
    return (r == 8) ? 8 : PLZ4CU_Z_WRITE;
}

uint32_t plz4cu_xxh32_host(const void* p, size_t n) { return xxh32_once(p, n); }

// ---- in-memory endpoints (bytes.Reader / bytes.Buffer for C callers)
struct plz4cu_membuf { uint8_t* data; size_t len, cap, pos; };
plz4cu_membuf_t* plz4cu_membuf_new(void* data, size_t len, size_t cap)
{
    plz4cu_membuf* m = new plz4cu_membuf{static_cast<uint8_t*>(data), len, cap, 0};
    return m;
}
void plz4cu_membuf_free(plz4cu_membuf_t* m) { delete m; }
size_t plz4cu_membuf_len(const plz4cu_membuf_t* m) { return m->len; }
int64_t plz4cu_membuf_read(void* ctx, void* buf, size_t n)
{
    plz4cu_membuf* m = static_cast<plz4cu_membuf*>(ctx);
    size_t k = std::min(n, m->len - m->pos);
    memcpy(buf, m->data + m->pos, k);
    m->pos += k;
    return (int64_t)k;
}
int64_t plz4cu_membuf_write(void* ctx, const void* data, size_t n)
{
    plz4cu_membuf* m = static_cast<plz4cu_membuf*>(ctx);
    if (m->len + n > m->cap) return -1;
    memcpy(m->data + m->len, data, n);
    m->len += n;
    return (int64_t)n;
}
int plz4cu_membuf_seek(void* ctx, int64_t delta)
{
    plz4cu_membuf* m = static_cast<plz4cu_membuf*>(ctx);
    if (delta < 0 || m->pos + (size_t)delta > m->len) return -1;
    m->pos += (size_t)delta;
    return 0;
}

}  // extern "C"
