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

#include <cuda_runtime.h>
#if defined(__x86_64__) && defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <future>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <new>
#include <vector>

std::vector<int> plz4cu_internal_devices();            // engine.cu: what plz4cu_init_devices registered

namespace {

// ---------------------------------------------------------------- in-order background work
//
// The reference overlaps its stages with goroutines and channels (async/writer.go:232-381, async/reader.go:128-271,
// async/hash.go:14-111).  Here a stream owns at most two helper threads, each running its jobs strictly in
// submission order: one drives the GPU engine and the sink / source callbacks, one runs the serial content
// checksum.  The thread starts with the first job, so small or synchronous streams never create one.
class SerialExec {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::deque<std::function<void()>> q;
    uint64_t submitted = 0, completed = 0;
    bool stop = false, started = false, finished = false;
    plz4cu_submit_fn spawn = nullptr;                 // WithWorkerPool: who runs this stage's loop (opts/opts.go:43-45)
    void* spawn_ctx = nullptr;
    static void entry(void* self) { static_cast<SerialExec*>(self)->run(); }
    void run()
    {
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv_work.wait(lk, [&] { return stop || !q.empty(); });
            if (q.empty()) break;
            std::function<void()> f = std::move(q.front());
            q.pop_front();
            lk.unlock();
            f();
            lk.lock();
            completed++;
            cv_done.notify_all();
        }
        finished = true;
        cv_done.notify_all();
    }
public:
    ~SerialExec()
    {
        {
            std::unique_lock<std::mutex> lk(mu);
            cv_done.wait(lk, [&] { return completed >= submitted; });
            stop = true;
        }
        cv_work.notify_all();
        if (th.joinable()) th.join();
        else if (started) {                               // the loop runs on a worker of the caller's pool: wait for it to leave
            std::unique_lock<std::mutex> lk(mu);
            cv_done.wait(lk, [&] { return finished; });
        }
    }
    void use_pool(plz4cu_submit_fn f, void* ctx) { spawn = f; spawn_ctx = ctx; }
    uint64_t submit(std::function<void()> f)               // returns the ticket to wait() on
    {
        std::lock_guard<std::mutex> lk(mu);
        if (!started) {
            started = true;
            // the reference hands its loops to opts.WorkerPool.Submit (async/writer.go:439-467); its stub is `go task()`
            if (!spawn || spawn(spawn_ctx, &SerialExec::entry, this) != 0) th = std::thread([this] { run(); });
        }
        q.push_back(std::move(f));
        cv_work.notify_one();
        return ++submitted;
    }
    void wait(uint64_t ticket)
    {
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return completed >= ticket; });
    }
    void drain()
    {
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return completed >= submitted; });
    }
};

// Bulk copies between caller memory and staging move each byte once and do not re-read it soon, so from 64 KiB up
// they use non-temporal stores (no read-for-ownership of the destination lines, no cache pollution), walking four
// pages side by side with the next four prefetched, which keeps several DRAM pages open.  glibc's memcpy only
// switches to such a mode for copies of tens of MiB; blocks and Write() calls are usually smaller.  Measured on
// the B200 box's host: 12-14 GB/s per thread against 6-9 GB/s for memcpy on 64 KiB .. 16 MiB pieces.
void bulk_copy(void* dst, const void* src, size_t n)
{
#if defined(__x86_64__) && defined(__SSE2__)
    if (n >= (64u << 10)) {
        uint8_t* d = static_cast<uint8_t*>(dst);
        const uint8_t* s = static_cast<const uint8_t*>(src);
        const size_t head = (64 - (reinterpret_cast<uintptr_t>(d) & 63)) & 63;
        memcpy(d, s, head);
        d += head; s += head; n -= head;
        auto line = [](uint8_t* dp, const uint8_t* sp) {
            const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(sp));
            const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(sp + 16));
            const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(sp + 32));
            const __m128i e = _mm_loadu_si128(reinterpret_cast<const __m128i*>(sp + 48));
            _mm_stream_si128(reinterpret_cast<__m128i*>(dp), a);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dp + 16), b);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dp + 32), c);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dp + 48), e);
        };
        constexpr size_t kPage = 4096, kGroup = 4 * kPage;
        for (; n >= 2 * kGroup; d += kGroup, s += kGroup, n -= kGroup) {       // 2x: the prefetch stays inside src
            for (size_t j = 0; j < kPage; j += 64) {
                for (size_t k = 0; k < kGroup; k += kPage) {
                    _mm_prefetch(reinterpret_cast<const char*>(s + k + j + kGroup), _MM_HINT_T0);
                    line(d + k + j, s + k + j);
                }
            }
        }
        for (; n >= 64; d += 64, s += 64, n -= 64) line(d, s);
        _mm_sfence();
        memcpy(d, s, n);
        return;
    }
#endif
    memcpy(dst, src, n);
}

// One thread moves 12-14 GB/s; the PCIe link behind the staging moves four times that.  Copies of 1 MiB and more are cut
// into pieces for a small pool of helper threads (the caller takes the first piece), so a stream fed from pageable memory
// is no longer bound by one core's memcpy (async/writer.go:81-107 has the same shape: the caller only slices, workers do
// the rest).  PLZ4CU_COPY_THREADS sets the number of helpers (default 3, 0 = none).
class CopyPool {
    std::vector<std::thread> th;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<std::function<void()>> q;
    bool stop = false;
public:
    explicit CopyPool(int n)
    {
        for (int i = 0; i < n; i++) th.emplace_back([this] {
            std::unique_lock<std::mutex> lk(mu);
            for (;;) {
                cv.wait(lk, [&] { return stop || !q.empty(); });
                if (q.empty()) return;
                std::function<void()> f = std::move(q.front());
                q.pop_front();
                lk.unlock();
                f();
                lk.lock();
            }
        });
    }
    ~CopyPool()
    {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv.notify_all();
        for (auto& t : th) t.join();
    }
    size_t size() const { return th.size(); }
    void submit(std::function<void()> f)
    {
        { std::lock_guard<std::mutex> lk(mu); q.push_back(std::move(f)); }
        cv.notify_one();
    }
};

CopyPool& copy_pool()
{
    static CopyPool* pool = [] {
        int n = 3;
        if (const char* e = getenv("PLZ4CU_COPY_THREADS")) n = std::max(0, std::min(32, atoi(e)));
        return new CopyPool(n);                             // lives as long as the process: streams may end at exit
    }();
    return *pool;
}

void bulk_copy_mt(void* dst, const void* src, size_t n, int site = 0)
{
    static const int sites = getenv("PLZ4CU_COPY_SITES") ? atoi(getenv("PLZ4CU_COPY_SITES")) : 15;
    if (!(sites & site)) { bulk_copy(dst, src, n); return; }
    constexpr size_t kMin = 1u << 20, kPiece = 256u << 10;
    CopyPool& pool = copy_pool();
    if (n < kMin || pool.size() == 0) { bulk_copy(dst, src, n); return; }
    const size_t parts = std::min(pool.size() + 1, n / kPiece);
    const size_t per = ((n / parts) + 4095) & ~size_t(4095);
    std::mutex mu;
    std::condition_variable cv;
    size_t left = 0;
    uint8_t* d = static_cast<uint8_t*>(dst);
    const uint8_t* s = static_cast<const uint8_t*>(src);
    for (size_t off = per; off < n; off += per) {
        const size_t len = std::min(per, n - off);
        { std::lock_guard<std::mutex> lk(mu); left++; }
        pool.submit([&, off, len] {
            bulk_copy(d + off, s + off, len);
            std::lock_guard<std::mutex> lk(mu);
            if (--left == 0) cv.notify_all();
        });
    }
    bulk_copy(d, s, std::min(per, n));
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return left == 0; });
    static const bool check = getenv("PLZ4CU_COPY_CHECK") != nullptr;
    if (check && memcmp(dst, src, n) != 0) { fprintf(stderr, "bulk_copy_mt: MISMATCH n=%zu site=%d\n", n, site); abort(); }
    if (check) fprintf(stderr, "bulk_copy_mt ok n=%zu site=%d parts=%zu\n", n, site, parts);
}

// The devices a stream spreads its batches over (opts.n_devices): whole runs of independent blocks per
// device, each through its own engine pipeline, no exchange between devices.
std::vector<int> stream_devices(int n_devices, int fallback)
{
    std::vector<int> reg = plz4cu_internal_devices();
    if (n_devices == 0 || n_devices == 1 || reg.size() < 2) return {fallback};
    if (n_devices > 0 && (size_t)n_devices < reg.size()) reg.resize((size_t)n_devices);
    return reg;
}

// run fn(part, first_block, block_count) for `parts` contiguous runs of nblk blocks, all at once, first error wins
template <typename F>
int for_each_part(uint32_t nblk, size_t parts, F fn)
{
    parts = std::min<size_t>(parts, std::max<uint32_t>(nblk, 1));
    if (parts <= 1) return fn(0, 0u, nblk);
    std::vector<std::future<int>> fut;
    uint32_t b0 = 0;
    std::vector<std::pair<uint32_t, uint32_t>> runs;
    for (size_t i = 0; i < parts; i++) {
        const uint32_t b1 = (uint32_t)(((uint64_t)nblk * (i + 1)) / parts);
        runs.emplace_back(b0, b1 - b0);
        b0 = b1;
    }
    for (size_t i = 1; i < parts; i++) fut.push_back(std::async(std::launch::async, [&, i] { return fn(i, runs[i].first, runs[i].second); }));
    int rc = fn(0, runs[0].first, runs[0].second);
    for (auto& f : fut) { const int r = f.get(); if (rc >= 0 && r < 0) rc = r; }
    return rc;
}

// stage timing of a reader for tuning (PLZ4CU_STREAM_PROF=1): microseconds per stage, printed when the reader is freed
struct StageClock {
    std::atomic<int64_t> us[6];
    StageClock() { for (auto& u : us) u = 0; }
    static bool on() { static const bool v = getenv("PLZ4CU_STREAM_PROF") != nullptr; return v; }
    static int64_t now() { return std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
};

int current_device()
{
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess) d = 0;
    return d;
}

// true when p is page-locked memory the GPU can DMA from directly (plz4cu_host_alloc, cudaHostAlloc, cudaHostRegister)
bool is_pinned(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

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
// Bytes of blocks gathered per engine call when batching (n_parallel != 0) and no pending size is given: enough blocks
// to occupy the GPU (one warp or one 64 KiB fragment per block), small enough that staging, engine and sink overlap.
constexpr size_t kAutoBatchMin = 64u << 20, kAutoBatchMax = 128u << 20;
constexpr size_t kAutoBatchBlocks = 256;
constexpr size_t kSmallStage = 8u << 20;             // streams shorter than this never touch pinned memory

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
    bool reserve(size_t n) { return grow(n, 0); }
    // capacity >= n, keeping the first `keep` bytes
    bool grow(size_t n, size_t keep)
    {
        if (n <= cap) return true;
        // small streams are not worth pinning; large ones borrow a pooled pinned slab
        const bool pin = n >= (4u << 20);
        uint8_t* q = static_cast<uint8_t*>(pin ? plz4cu_host_alloc(n) : malloc(std::max<size_t>(n, 64)));
        if (!q) return false;
        if (keep) memcpy(q, p, keep);
        release();
        p = q; cap = n; pinned = pin;
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
    // `decode`: a decode batch costs at least one block's serial decode time, so large blocks want more bytes per
    // batch; the encoder splits large blocks into 64 KiB fragments and is content with the minimum
    size_t batch_bytes(int bsz, bool decode) const
    {
        if (o.n_parallel == 0) return (size_t)bsz;                                // synchronous flavour: one block per call
        size_t want = kAutoBatchMin;
        if (o.pending_size > 0) want = (size_t)o.pending_size;
        else if (decode) want = std::min(kAutoBatchMax, std::max(kAutoBatchMin, kAutoBatchBlocks * (size_t)bsz));
        else if (bsz >= (1 << 20)) want = 2 * kAutoBatchMin;   // fragments + stitch cost a few ms per call whatever its size
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

// the several-thread copy for the engine's own staging of pageable caller buffers (engine.cu)
void plz4cu_internal_copy(void* dst, const void* src, size_t n) { bulk_copy_mt(dst, src, n, 1); }

// ================================================================ Writer

struct plz4cu_writer {
    plz4cu_write_fn wr;
    void* wr_ctx;
    Opts opt;
    int bsz;
    size_t batch;
    const bool async;                                 // n_parallel != 0: stage, engine + sink, content hash overlap
    bool threaded = false;                            // ... from the moment the stream proves large (first slab)
    const int device;
    bool header_written = false, closed = false, reported = false;
    std::atomic<int> state{0};                        // sticky error (first error wins, async/writer.go:553-555)
    int64_t src_mark = 0, dst_mark = 0;               // sink stage only
    XXH32 hasher;                                     // hash thread only (caller thread when synchronous)
    std::vector<int> devs;                            // devices the batches are spread over (one unless opts.n_devices says so)
    std::vector<plz4cu_dict_t*> dicts;                // the dictionary on each of them (nullptr: none)
    // staging: pageable while the stream is small, then two pinned slabs of one batch each — the caller fills one
    // while the engine thread works on the other (the reference's blocks-in-flight window, opts/opts.go:62-95)
    std::vector<uint8_t> small;
    struct Slab { uint8_t* p = nullptr; uint64_t engine_ticket = 0, hash_ticket = 0; };
    Slab slabs[2];
    int cur_slab = -1;                                // -1: still in `small`
    size_t fill = 0;
    struct Packed { PinnedBuf buf; uint64_t sink_ticket = 0; };
    Packed packed[2];                                 // engine thread fills, sink thread drains
    int pk_cur = 0;
    std::vector<uint64_t> offs, poff;                 // engine thread only
    std::vector<uint32_t> lens;
    // stages, each strictly in order: engine (H2D, kernels, D2H) -> sink (caller's io.Writer, marks, progress);
    // the content checksum of the same bytes runs beside them
    SerialExec engine_q, sink_q, hash_q;

    plz4cu_writer(plz4cu_write_fn w, void* c, const plz4cu_opts_t* o)
        : wr(w), wr_ctx(c), opt(o), async(opt.o.n_parallel != 0), device(current_device())
    {
        bsz = block_size_of(opt.o.block_size_idx);
        batch = opt.batch_bytes(bsz, false);
        if (opt.o.level != 1 || opt.o.block_linked) state = PLZ4CU_Z_UNSUPPORTED;
        devs = stream_devices(opt.o.n_devices, device);
        dicts.assign(devs.size(), nullptr);
        if (!opt.dict.empty() && state == 0) {
            for (size_t i = 0; i < devs.size() && state == 0; i++) {         // the dictionary lives on every device the stream uses
                cudaSetDevice(devs[i]);
                dicts[i] = plz4cu_dict_create(opt.dict.data(), opt.dict.size());
                if (!dicts[i]) state = PLZ4CU_Z_ENGINE;
            }
            cudaSetDevice(device);
        }
        if (opt.o.submit) { engine_q.use_pool(opt.o.submit, opt.o.submit_ctx); sink_q.use_pool(opt.o.submit, opt.o.submit_ctx); hash_q.use_pool(opt.o.submit, opt.o.submit_ctx); }
    }
    ~plz4cu_writer()
    {
        engine_q.drain();
        sink_q.drain();
        hash_q.drain();
        for (Slab& s : slabs) if (s.p) plz4cu_host_free(s.p);
        for (plz4cu_dict_t* d : dicts) if (d) plz4cu_dict_destroy(d);
    }
    int report() { int s = state; if (s) reported = true; return s; }
    void set_error(int e) { int expect = 0; state.compare_exchange_strong(expect, e); }

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

    // Compress data[0..n) as consecutive bsz-sized blocks (the last may be short), then hand the packed records to
    // the sink stage.  Runs on the engine thread when async; the two packed buffers alternate so that the sink can
    // still be writing batch k while batch k+1 is being compressed.
    int emit(const uint8_t* data, size_t n)
    {
        if (n == 0) return 0;
        const uint32_t nblk = (uint32_t)((n + bsz - 1) / bsz);
        offs.resize(nblk); lens.resize(nblk); poff.resize(nblk + 1);
        for (uint32_t i = 0; i < nblk; i++) { offs[i] = (uint64_t)i * bsz; lens[i] = (uint32_t)std::min<size_t>(bsz, n - offs[i]); }
        const size_t packed_cap = (size_t)nblk * (bsz + 8);
        Packed& pk = packed[pk_cur];
        pk_cur ^= 1;
        if (threaded) sink_q.wait(pk.sink_ticket);
        if (!pk.buf.reserve(packed_cap)) return PLZ4CU_Z_ENGINE;
        if (devs.size() > 1 && nblk >= 2 * devs.size()) {
            // several devices: contiguous runs of blocks, one per device, each through its own pipeline and into its own
            // stretch of the packed buffer; the sink then writes the runs in order (the bytes are those of one device)
            struct Run { uint32_t b0, nb; std::vector<uint64_t> off, poff; std::vector<uint32_t> len; };
            auto runs = std::make_shared<std::vector<Run>>(devs.size());
            const int rc = for_each_part(nblk, devs.size(), [&](size_t i, uint32_t b0, uint32_t nb) -> int {
                Run& r = (*runs)[i];
                r.b0 = b0; r.nb = nb;
                if (nb == 0) return 0;
                cudaSetDevice(devs[i]);
                r.off.resize(nb); r.len.resize(nb); r.poff.resize(nb + 1);
                for (uint32_t j = 0; j < nb; j++) { r.off[j] = (uint64_t)j * bsz; r.len[j] = lens[b0 + j]; }
                return plz4cu_compress_batch_host(data + (size_t)b0 * bsz, r.off.data(), r.len.data(), nb, (uint32_t)bsz, opt.o.block_checksum, 0,
                                                  dicts[i], pk.buf.p + (size_t)b0 * (bsz + 8), (size_t)nb * (bsz + 8), r.poff.data());
            });
            cudaSetDevice(device);
            if (rc < 0) return PLZ4CU_Z_ENGINE;
            const uint8_t* base = pk.buf.p;
            auto deliver_runs = [this, base, runs, n]() -> int {
                for (const Run& r : *runs) {
                    if (r.nb == 0) continue;
                    const size_t bytes = std::min<size_t>((size_t)r.nb * bsz, n - (size_t)r.b0 * bsz);
                    std::vector<uint64_t> jp = r.poff;
                    if (!opt.o.progress) jp.assign({0, r.poff[r.nb]});
                    if (int e = deliver(base + (size_t)r.b0 * (bsz + 8), bytes, r.len, jp)) return e;
                }
                return 0;
            };
            if (!threaded) return deliver_runs();
            pk.sink_ticket = sink_q.submit([this, deliver_runs] {
                if (state) return;
                if (int e = deliver_runs()) set_error(e);
            });
            return 0;
        }
        int rc = plz4cu_compress_batch_host(data, offs.data(), lens.data(), nblk, (uint32_t)bsz, opt.o.block_checksum, 0, dicts[0],
                                            pk.buf.p, packed_cap, poff.data());
        if (rc < 0) { if (getenv("PLZ4CU_DEBUG")) fprintf(stderr, "writer: engine call failed (%d): %s\n", rc, plz4cu_last_error()); return PLZ4CU_Z_ENGINE; }
        if (!threaded) return deliver(pk.buf.p, n, lens, poff);
        // the sink job owns copies of the per-block tables only when somebody watches block boundaries
        std::vector<uint32_t> jl;
        std::vector<uint64_t> jp;
        if (opt.o.progress) { jl = lens; jp = poff; } else jp.assign({0, poff[nblk]});
        const uint8_t* base = pk.buf.p;
        pk.sink_ticket = sink_q.submit([this, base, n, jl = std::move(jl), jp = std::move(jp)] {
            if (state) return;
            if (int e = deliver(base, n, jl, jp)) set_error(e);
        });
        return 0;
    }
    // Write one batch of packed records in order (sink thread when async): header first, marks, progress callbacks.
    int deliver(const uint8_t* base, size_t n, const std::vector<uint32_t>& blens, const std::vector<uint64_t>& bpoff)
    {
        if (int e = ensure_header()) return e;
        if (!opt.o.progress) {
            // nobody watches block boundaries: one write for the whole batch
            const uint64_t total = bpoff.back();
            if (int e = write_all(base, (size_t)total, PLZ4CU_Z_WRITE)) return e;
            src_mark += (int64_t)n; dst_mark += (int64_t)total;
            return 0;
        }
        for (size_t i = 0; i + 1 < bpoff.size(); i++) {
            const size_t len = (size_t)(bpoff[i + 1] - bpoff[i]);
            int e = write_all(base + bpoff[i], len, PLZ4CU_Z_WRITE);
            opt.o.progress(opt.o.progress_ctx, src_mark, dst_mark);        // async/writer.go:327-331
            src_mark += blens[i]; dst_mark += (int64_t)len;
            if (e) return e;
        }
        return 0;
    }

    // Hand data[0..n) to the engine (and the content hasher).  With a slab the call returns at once and the slab
    // carries the tickets; caller-owned or pageable staging memory is waited for before returning.
    void dispatch(const uint8_t* data, size_t n, Slab* slab)
    {
        if (n == 0) return;
        if (!threaded) {
            // synchronous flavour (sync/writer.go:53-290), and the first bytes of any stream: everything on the
            // caller's thread; only a large batch is worth a helper for the content checksum
            std::future<void> hf;
            if (opt.o.content_checksum) {
                if (n >= (1u << 20)) hf = std::async(std::launch::async, [=] { hasher.update(data, n); });
                else hasher.update(data, n);
            }
            int e = emit(data, n);
            if (hf.valid()) hf.get();
            if (e) set_error(e);
            return;
        }
        uint64_t ht = 0;
        if (opt.o.content_checksum) ht = hash_q.submit([=] { hasher.update(data, n); });      // async/hash.go
        const uint64_t et = engine_q.submit([=] {
            if (state) return;                         // after the first error nothing more is compressed or written
            cudaSetDevice(device);
            if (int e = emit(data, n)) set_error(e);
        });
        if (slab) { slab->engine_ticket = et; slab->hash_ticket = ht; return; }
        engine_q.wait(et);
        if (ht) hash_q.wait(ht);
    }
    void wait_slab(Slab& s)
    {
        engine_q.wait(s.engine_ticket);
        if (s.hash_ticket) hash_q.wait(s.hash_ticket);
    }

    uint8_t* stage_ptr() { return cur_slab < 0 ? small.data() : slabs[cur_slab].p; }
    // make room for `want` staged bytes; moves from the pageable vector to the first pinned slab when a stream grows
    bool reserve(size_t want)
    {
        if (cur_slab >= 0) return true;
        if (!async || want <= kSmallStage || batch <= kSmallStage) {
            if (small.size() < want) small.resize(std::max(want, small.size() * 2));
            return true;
        }
        slabs[0].p = static_cast<uint8_t*>(plz4cu_host_alloc(batch));
        if (!slabs[0].p) return false;
        threaded = true;                               // the stream is large: from here on the stages overlap
        memcpy(slabs[0].p, small.data(), fill);
        small.clear(); small.shrink_to_fit();
        cur_slab = 0;
        return true;
    }
    // send the staged bytes off and get an empty staging area
    bool submit_stage()
    {
        if (fill == 0) return true;
        if (cur_slab < 0) { dispatch(small.data(), fill, nullptr); fill = 0; return true; }
        dispatch(slabs[cur_slab].p, fill, &slabs[cur_slab]);
        fill = 0;
        cur_slab ^= 1;
        Slab& s = slabs[cur_slab];
        if (s.p) { wait_slab(s); return true; }
        s.p = static_cast<uint8_t*>(plz4cu_host_alloc(batch));
        return s.p != nullptr;
    }
    // room left in the staging area right now (for callers that produce straight into it)
    uint8_t* stage_space(size_t want, size_t* avail)
    {
        const size_t take = std::min(want, batch - fill);
        if (!reserve(fill + take)) return nullptr;
        *avail = take;
        return stage_ptr() + fill;
    }
    void stage_commit(size_t n)
    {
        fill += n;
        if (fill == batch && !submit_stage()) set_error(PLZ4CU_Z_ENGINE);
    }

    int64_t write(const uint8_t* src, size_t n)
    {
        if (state) return report();
        size_t done = 0;
        while (done < n && !state) {
            if (fill == 0 && n - done >= batch && (!async || is_pinned(src + done))) {
                // whole batches straight from the caller's buffer, no staging copy (sync/writer.go:99-109); when
                // batching this is only a win for page-locked memory — pageable bytes are staged so that the copy
                // overlaps the GPU instead of crawling over PCIe through the driver's bounce buffer
                if (async) threaded = true;
                dispatch(src + done, batch, nullptr);
                done += batch;
                continue;
            }
            size_t take = 0;
            uint8_t* dst = stage_space(n - done, &take);
            if (!dst) { set_error(PLZ4CU_Z_ENGINE); break; }
            bulk_copy_mt(dst, src + done, take, 1);
            done += take;
            stage_commit(take);
        }
        if (state) return report();
        return (int64_t)done;
    }
    int64_t read_from(plz4cu_read_fn rd, void* rd_ctx)
    {
        if (state) return report();
        int64_t total = 0;
        while (!state) {
            size_t room = 0;
            uint8_t* dst = stage_space(1 << 20, &room);       // the source fills the staging area directly
            if (!dst) { set_error(PLZ4CU_Z_ENGINE); break; }
            int64_t r = rd(rd_ctx, dst, room);
            if (r < 0) { set_error(PLZ4CU_Z_BLOCK_READ); break; }
            if (r == 0) break;
            stage_commit((size_t)r);
            total += r;
        }
        if (state) return report();
        return total;
    }
    // Flush barrier (async/writer.go:284-314): everything written so far is compressed and handed to the sink
    void barrier()
    {
        if (!state && !submit_stage()) set_error(PLZ4CU_Z_ENGINE);
        engine_q.drain();
        sink_q.drain();
        hash_q.drain();
    }
    int flush()
    {
        if (state) return report();
        barrier();
        return report();
    }
    int close()
    {
        if (closed) return report();
        barrier();
        if (!state) {
            int e = ensure_header();
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
    std::vector<int> devs;                            // devices the batches are spread over
    std::vector<plz4cu_dict_t*> dicts;                // the frame's dictionary on each of them (empty: none yet)
    std::vector<uint8_t> cur_dict;

    // current frame
    bool in_body = false;
    int bsz = 0;
    bool blk_check = false, has_content_hash = false, verify_content_hash = false;
    bool has_content_size = false;
    uint64_t hdr_content_size = 0, content_acc = 0;
    XXH32 hasher;

    // Decoded blocks arrive in batches.  When batching (n_parallel != 0) a source thread reads the records of the next
    // batches while earlier ones are being decoded and the caller drains the current one (async/reader.go:128-221's
    // read-ahead); every batch slot decodes on its own engine thread, so several batches can be on the GPU at once —
    // a batch of large blocks keeps only a few warps busy, and only concurrency between batches fills the device.
    // The content checksum of the decoded bytes runs on its own thread (async/hash.go).  Batches grow from small to
    // full size so that a short stream neither waits for nor pins a full-size staging area.
    struct Batch {
        PinnedBuf recs, out;
        size_t recs_len = 0;
        std::vector<uint64_t> rec_off;
        std::vector<uint32_t> rec_read;               // input bytes each block consumed (size word + body + hash)
        std::vector<int32_t> out_len;
        uint32_t nblk = 0;
        int tail_event = 0;                           // after the batch: 0 nothing, 2 EndMark, <0 error
        uint32_t endmark_read = 0;                    // bytes the EndMark (+ content hash) consumed
        uint32_t content_hash_read = 0;
        uint32_t tail_read = 0;                       // bytes consumed by a failed trailing read (for src_pos)
        uint32_t digest = 0;                          // content checksum up to the end of this batch
        uint64_t ticket = 0, hash_ticket = 0;         // decode job on engine_q[slot], hash job on hash_q
    };
    static constexpr int kSlots = 8;
    Batch bt[kSlots];
    // ring bookkeeping: batch number k lives in slot k % kSlots.  `produced` batches have been read (decode submitted),
    // `released` have been handed back by the caller; the source loop runs while produced - released <= depth.
    std::mutex ring_mu;
    std::condition_variable ring_cv;
    uint64_t produced = 0, released = 0;
    bool source_done = true, stop_source = false;
    uint64_t cb_index = 0;                            // batch the caller is draining (valid while have_batch)
    bool have_batch = false;
    size_t next_batch_bytes = 0;
    const bool async;
    const int device;
    uint32_t cur = 0, run_first = 0;                  // next block to serve; first block of the piece being served
    size_t cur_off = 0, cur_len = 0;
    bool have_block = false;
    SerialExec source_q, engine_q[kSlots], hash_q;
    Batch& cbatch() { return bt[cb_index % kSlots]; }

    plz4cu_reader(plz4cu_read_fn r, plz4cu_seek_fn s, void* c, const plz4cu_opts_t* o)
        : rd(r), seek(s), ctx(c), opt(o), async(opt.o.n_parallel != 0), device(current_device())
    {
        read_offset = opt.o.read_offset;
        skip_content_size = !opt.o.content_size_check;
        cur_dict = opt.dict;
        devs = stream_devices(opt.o.n_devices, device);
        if (opt.o.submit) {
            source_q.use_pool(opt.o.submit, opt.o.submit_ctx); hash_q.use_pool(opt.o.submit, opt.o.submit_ctx);
            for (SerialExec& e : engine_q) e.use_pool(opt.o.submit, opt.o.submit_ctx);
        }
    }
    void drop_dicts()
    {
        for (plz4cu_dict_t* d : dicts) if (d) plz4cu_dict_destroy(d);
        dicts.clear();
    }
    ~plz4cu_reader()
    {
        quiesce();
        drop_dicts();
        if (StageClock::on())
            fprintf(stderr, "reader stages (ms): source read %.1f  decode (sum over engine threads) %.1f  caller waits for source %.1f + for decode %.1f  sink writes %.1f\n",
                    clk.us[0] / 1e3, clk.us[1] / 1e3, clk.us[2] / 1e3, clk.us[4] / 1e3, clk.us[3] / 1e3);
        if (StageClock::on()) fprintf(stderr, "  of the source read: %.1f ms inside %lld read callbacks, %.1f ms in %lld growths of the record area\n", clk.us[5] / 1e3, (long long)rd_calls.load(), grow_us.load() / 1e3, (long long)grows.load());
    }
    void quiesce()
    {
        { std::lock_guard<std::mutex> lk(ring_mu); stop_source = true; }
        ring_cv.notify_all();
        source_q.drain();
        for (SerialExec& q : engine_q) q.drain();
        hash_q.drain();
        stop_source = false;
    }

    // io.ReadFull: 0 = ok, 1 = clean EOF before any byte, -1 = short read or I/O error (an error is never an EOF)
    int read_full(uint8_t* p, size_t n, size_t* got)
    {
        size_t g = 0;
        bool failed = false;
        while (g < n) {
            int64_t r = rd(ctx, p + g, n - g);
            if (r < 0) failed = true;
            if (r <= 0) break;
            g += (size_t)r;
        }
        *got = g;
        if (g == n) return 0;
        return (g == 0 && !failed) ? 1 : -1;
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
                // The size comes from the stream: never allocate on its say-so.  Without a callback the payload is read
                // and dropped 64 KiB at a time (header/skip.go: io.CopyN(io.Discard)); with one the buffer grows only as
                // bytes really arrive, and an allocation failure is this frame's error, not the process's end.
                uint32_t sz = get32(h + 4);
                std::vector<uint8_t> payload;
                try {
                    const size_t kPiece = 64u << 10;
                    if (!opt.o.skip_cb) payload.resize((size_t)std::min<uint64_t>(sz, kPiece));
                    for (uint64_t done = 0; done < sz;) {
                        const size_t want = (size_t)std::min<uint64_t>(sz - done, kPiece);
                        uint8_t* at = payload.data();
                        if (opt.o.skip_cb) { payload.resize((size_t)done + want); at = payload.data() + done; }
                        r = read_full(at, want, &got);
                        src_pos += (int64_t)got;
                        if (r != 0) return PLZ4CU_Z_SKIP;
                        done += want;
                    }
                } catch (const std::bad_alloc&) {
                    return PLZ4CU_Z_SKIP;
                }
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
                if (dp) { cur_dict.assign(static_cast<const uint8_t*>(dp), static_cast<const uint8_t*>(dp) + dl); drop_dicts(); }
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
            if (!cur_dict.empty() && dicts.empty()) {
                for (int d : devs) {                                    // one copy per device the batches go to
                    cudaSetDevice(d);
                    dicts.push_back(plz4cu_dict_create(cur_dict.data(), cur_dict.size()));
                    if (!dicts.back()) { cudaSetDevice(device); return PLZ4CU_Z_ENGINE; }
                }
                cudaSetDevice(device);
            }
            bsz = block_size_of((bd >> 4) & 7);
            blk_check = (flags & 0x10) != 0;
            has_content_hash = (flags & 0x04) != 0;
            verify_content_hash = check_hash;
            has_content_size = (flags & 0x08) != 0;
            hdr_content_size = csz;
            content_acc = 0;
            quiesce();                                  // no read-ahead or hashing crosses a frame boundary
            hasher.reset();
            have_batch = false;
            for (Batch& b : bt) { b.nblk = 0; b.tail_event = 0; }
            cur = 0;
            in_body = true;
            return 0;
        }
    }

    // blk/frame.go:54-112 for up to `want_bytes` of blocks: the records of one batch, in order (source thread).
    // Bytes read past the last record of a batch while the body goes on (bulk reads, below); they open the next batch.
    std::vector<uint8_t> carry;
    StageClock clk;                                   // 0 source read, 1 decode, 2 wait for a batch, 3 sink write
    std::atomic<int64_t> rd_calls{0}, grow_us{0}, grows{0};
    double ratio_seen = 0.0;                          // largest record bytes / block bytes of a batch so far (source thread only)

    void read_records(Batch& b, size_t want_bytes)
    {
        const size_t batch_blocks = std::max<size_t>(1, want_bytes / (size_t)bsz);
        b.recs_len = 0; b.rec_off.clear(); b.rec_read.clear();
        b.nblk = 0; b.tail_event = 0; b.tail_read = 0; b.hash_ticket = 0; b.ticket = 0;
        uint32_t nblk = 0;
        // The reference reads a size word, then a body, block after block (blk/frame.go:54-112) — two small reads per
        // block, which one thread cannot do faster than a few GB/s.  When the source can seek, the record area is filled
        // in pieces of up to 8 MiB instead (one read callback each, copied by several threads when it comes from memory)
        // and walked in place; what was read beyond the batch is kept for the next one, and what was read beyond the
        // frame is given back with one seek, so the source stands exactly where the reference would leave it.
        const bool bulk = async && seek != nullptr;
        const size_t cap_max = batch_blocks * ((size_t)bsz + 8) + 64 + (bulk ? (9u << 20) : 0);
        size_t fill = 0;                                // bytes of the stream sitting in b.recs
        bool eof = false, failed = false;
        // the record area grows with what actually arrives (a short stream never pins a full batch); once a batch has shown
        // how the stream compresses, later areas are sized for that in one step instead of by doubling
        const size_t expect = ratio_seen > 0.0 ? (size_t)((double)batch_blocks * (double)bsz * ratio_seen * 1.125) + (bulk ? (9u << 20) : 0) : 0;
        auto room_for = [&](size_t upto) -> bool {
            if (upto <= b.recs.cap) return true;
            const int64_t tg = StageClock::on() ? StageClock::now() : 0;
            const bool ok = b.recs.grow(std::min(cap_max, std::max(std::max(upto, expect), 2 * b.recs.cap)), fill);
            if (StageClock::on()) { grow_us += StageClock::now() - tg; grows++; }
            return ok;
        };
        if (bulk && !carry.empty()) {
            if (!room_for(carry.size())) { b.tail_event = PLZ4CU_Z_ENGINE; b.nblk = 0; return; }
            memcpy(b.recs.p, carry.data(), carry.size());
            fill = carry.size();
            carry.clear();
        }
        // make stream bytes [0, upto) present in b.recs; false = the stream ended (or failed) first
        auto need = [&](size_t upto) -> bool {
            while (fill < upto && !eof) {
                if (!bulk) {
                    size_t got = 0;
                    const int r = read_full(b.recs.p + fill, upto - fill, &got);
                    fill += got;
                    if (r != 0) { eof = true; failed = r < 0; }
                    break;
                }
                size_t piece = std::min<size_t>(8u << 20, std::max<size_t>(64u << 10, b.recs.cap - fill));
                if (ratio_seen > 0.0) {
                    // near the end of the batch, read about what its remaining blocks should take: what is read beyond the
                    // batch has to be carried over to the next one by copy
                    const size_t ahead = fill - std::min(fill, b.recs_len);
                    const size_t rest = (size_t)((double)(batch_blocks - nblk) * (double)bsz * ratio_seen * 1.03) + (64u << 10);
                    piece = std::min(piece, std::max<size_t>(rest > ahead ? rest - ahead : 0, 64u << 10));
                }
                piece = std::max(piece, upto - fill);
                if (!room_for(fill + piece)) { eof = true; failed = true; break; }
                const int64_t tr0 = StageClock::on() ? StageClock::now() : 0;
                const int64_t r = rd(ctx, b.recs.p + fill, std::min(piece, b.recs.cap - fill));
                if (StageClock::on()) { clk.us[5] += StageClock::now() - tr0; rd_calls++; }
                if (r <= 0) { eof = true; failed = r < 0; break; }
                fill += (size_t)r;
            }
            return fill >= upto;
        };
        while (nblk < batch_blocks) {
            const size_t at = b.recs_len;
            if (!room_for(at + 4)) { b.tail_event = PLZ4CU_Z_ENGINE; break; }
            if (!need(at + 4)) { b.tail_event = PLZ4CU_Z_BLOCK_SIZE_READ; b.tail_read = (uint32_t)(fill - at); b.recs_len = fill; break; }
            const uint32_t word = get32(b.recs.p + at);
            if (word == 0) {                            // EndMark (+ content checksum)
                b.endmark_read = 4;
                b.tail_event = 2;
                b.recs_len = at + 4;
                if (has_content_hash) {
                    if (!room_for(at + 8) || !need(at + 8)) {
                        b.endmark_read += (uint32_t)(fill - (at + 4));
                        b.tail_event = PLZ4CU_Z_CONTENT_HASH_READ; b.tail_read = b.endmark_read;
                        b.recs_len = fill;
                        break;
                    }
                    b.endmark_read += 4;
                    b.content_hash_read = get32(b.recs.p + at + 4);
                    b.recs_len = at + 8;
                }
                break;
            }
            const uint32_t n = word & 0x7FFFFFFFu;
            if (n > (uint32_t)bsz) { b.tail_event = PLZ4CU_Z_BLOCK_SIZE_OVERFLOW; b.tail_read = 4; b.recs_len = at + 4; break; }
            const size_t body = (size_t)n + (blk_check ? 4 : 0);
            // the record area grows with what actually arrives (a short stream never pins a full batch)
            if (!room_for(at + 4 + body)) { b.tail_event = PLZ4CU_Z_ENGINE; break; }
            if (!need(at + 4 + body)) { b.tail_event = PLZ4CU_Z_BLOCK_READ; b.tail_read = (uint32_t)(fill - at); b.recs_len = fill; break; }
            b.recs_len = at + 4 + body;
            b.rec_off.push_back(at);
            b.rec_read.push_back((uint32_t)(4 + body));
            nblk++;
        }
        b.nblk = nblk;
        if (nblk > 0) ratio_seen = std::max(ratio_seen, (double)b.recs_len / ((double)nblk * (double)bsz));
        if (bulk && fill > b.recs_len) {
            const size_t extra = fill - b.recs_len;
            if (b.tail_event == 0) carry.assign(b.recs.p + b.recs_len, b.recs.p + fill);        // the body goes on
            else if (seek(ctx, -(int64_t)extra) != 0) { b.tail_event = PLZ4CU_Z_ENGINE; }        // past the frame: give it back
        }
        if (b.tail_event != 0) {
            // the records end where the last whole one ends (what follows is the tail event's own bytes)
            b.recs_len = b.rec_off.empty() ? 0 : (size_t)b.rec_off.back() + b.rec_read.back();
        }
        (void)failed;
    }
    // one engine call for the batch (the slot's engine thread when async)
    void decode_records(Batch& b)
    {
        if (b.nblk == 0) return;
        b.out_len.resize(b.nblk);
        int rc = b.out.reserve((size_t)b.nblk * bsz) ? 0 : -1;
        // several devices: contiguous runs of the batch's blocks, one per device; every block has its own output slot, so
        // the runs need no joining
        const size_t parts = (devs.size() > 1 && b.nblk >= 2 * devs.size()) ? devs.size() : 1;
        if (rc == 0) rc = for_each_part(b.nblk, parts, [&](size_t i, uint32_t b0, uint32_t nb) -> int {
            if (nb == 0) return 0;
            cudaSetDevice(devs[parts > 1 ? i : 0]);
            plz4cu_dict_t* d = dicts.empty() ? nullptr : dicts[parts > 1 ? i : 0];
            return plz4cu_decompress_batch_host(b.recs.p, b.recs_len, b.rec_off.data() + b0, nullptr, nb, (uint32_t)bsz, blk_check, 0,
                                                d, b.out.p + (size_t)b0 * bsz, (uint64_t)bsz, b.out_len.data() + b0);
        });
        if (rc < 0) { if (getenv("PLZ4CU_DEBUG")) fprintf(stderr, "reader: engine call failed (%d): %s\n", rc, plz4cu_last_error()); b.nblk = 0; b.tail_event = PLZ4CU_Z_ENGINE; }
    }
    // the serial checksum of the decoded bytes, in stream order, stops at the first block that failed; submitted by the
    // caller's thread when it starts on the batch, so the jobs line up in batch order whatever order decodes finish in
    void hash_batch(Batch& b)
    {
        if (!verify_content_hash) return;
        auto job = [this, &b] {
            for (uint32_t i = 0; i < b.nblk && b.out_len[i] >= 0; i++) hasher.update(b.out.p + (size_t)i * bsz, (size_t)b.out_len[i]);
            b.digest = hasher.digest();
        };
        // small batches are hashed in place (in order: the hash thread, if it ever started, is idle first)
        if (async && (size_t)b.nblk * bsz >= (1u << 20)) b.hash_ticket = hash_q.submit(job);
        else { hash_q.drain(); job(); }
    }
    size_t next_fill_bytes()
    {
        const size_t limit = opt.batch_bytes(bsz, true);
        if (!async) return limit;
        // first batch: 4 MiB of small blocks, four large ones — the caller (and the serial content checksum behind it) gets its
        // first bytes after one block's decode time; the batches that follow are four times as large each and decode
        // beside one another, so the stream's rate does not depend on the small start
        if (next_batch_bytes == 0) next_batch_bytes = bsz >= (1 << 20) ? 4 * (size_t)bsz : std::max<size_t>(64 * (size_t)bsz, 1u << 20);
        const size_t want = std::min(next_batch_bytes, limit);
        next_batch_bytes = std::min(limit, next_batch_bytes * 4);
        return want;
    }
    // batches ahead of the caller: four for small blocks (one being read, two decoding, one ready; 3 / 4 / 5 / 6 measured
    // with tools/read_sweep.sh: 19-24 / 26-27 / 23-27 / 20-28 GB/s), more for large ones
    int read_ahead_depth() const
    {
        static const int forced = getenv("PLZ4CU_READ_AHEAD") ? atoi(getenv("PLZ4CU_READ_AHEAD")) : 0;   // measurements
        if (forced > 0) return std::min(kSlots - 1, forced);
        return bsz >= (1 << 20) ? kSlots - 1 : std::min(kSlots - 1, 4);
    }
    // source thread: read batch after batch in order, hand each to its slot's engine thread, stay at most
    // read_ahead_depth() batches ahead of the caller; ends with the batch that carries the body's tail event
    void source_loop()
    {
        cudaSetDevice(device);
        for (;;) {
            uint64_t k;
            {
                std::unique_lock<std::mutex> lk(ring_mu);
                ring_cv.wait(lk, [&] { return stop_source || produced - released <= (uint64_t)read_ahead_depth(); });
                if (stop_source) { source_done = true; ring_cv.notify_all(); return; }
                k = produced;
            }
            Batch& b = bt[k % kSlots];
            if (b.hash_ticket) hash_q.wait(b.hash_ticket);      // the previous tenant of these buffers may still be hashed
            const int64_t t0 = StageClock::on() ? StageClock::now() : 0;
            read_records(b, next_fill_bytes());
            if (StageClock::on()) clk.us[0] += StageClock::now() - t0;
            const bool last = b.tail_event != 0;
            if (b.nblk) b.ticket = engine_q[k % kSlots].submit([this, &b] {
                cudaSetDevice(device);
                const int64_t t1 = StageClock::on() ? StageClock::now() : 0;
                decode_records(b);
                if (StageClock::on()) clk.us[1] += StageClock::now() - t1;
            });
            {
                std::lock_guard<std::mutex> lk(ring_mu);
                produced = k + 1;
                if (last) source_done = true;
            }
            ring_cv.notify_all();
            if (last) return;
        }
    }
    // make the next batch current.  With small blocks the first batch of a body is read and decoded on the caller's
    // thread (it is waited for anyway), so a stream that fits one batch never starts a thread; if the body goes on, the
    // source loop takes over.  With large blocks the loop starts at once: one block alone takes tens of milliseconds
    // to decode, and the batches behind the first should be decoding during that time.
    void advance_batch()
    {
        uint64_t next = have_batch ? cb_index + 1 : 0;
        if (!async) {
            // synchronous flavour: one slot, refilled in place
            read_records(bt[0], next_fill_bytes());
            decode_records(bt[0]);
            next = 0;
        } else if (!have_batch && bsz < (1 << 20)) {
            Batch& b = bt[0];
            if (b.hash_ticket) hash_q.wait(b.hash_ticket);
            read_records(b, next_fill_bytes());
            decode_records(b);
            produced = 1; released = 0;
            source_done = b.tail_event != 0;
            if (!source_done) source_q.submit([this] { source_loop(); });
        } else {
            if (!have_batch) {
                produced = 0; released = 0;
                source_done = false;
                source_q.submit([this] { source_loop(); });
            }
            bool there;
            {
                std::unique_lock<std::mutex> lk(ring_mu);
                released = next;                                // everything before `next` may be overwritten
                ring_cv.notify_all();
                const int64_t t2 = StageClock::on() ? StageClock::now() : 0;
                ring_cv.wait(lk, [&] { return produced > next || source_done; });
                there = produced > next;
                if (StageClock::on()) clk.us[2] += StageClock::now() - t2;
            }
            Batch& b = bt[next % kSlots];
            const int64_t t3 = StageClock::on() ? StageClock::now() : 0;
            if (there) engine_q[next % kSlots].wait(b.ticket);
            else { b.nblk = 0; b.tail_event = PLZ4CU_Z_ENGINE; }         // the source loop stopped short: cannot happen while reading
            if (StageClock::on()) clk.us[4] += StageClock::now() - t3;
        }
        cb_index = next;
        have_batch = true;
        cur = 0;
        hash_batch(cbatch());
    }

    // rdr/rdr.go:207-227 nextBlock: 0 = a block is current, 2 = EndMark, <0 error
    int next_block()
    {
        have_block = false;
        if (!have_batch || (cur >= cbatch().nblk && cbatch().tail_event == 0)) advance_batch();
        Batch& b = cbatch();
        if (opt.o.progress) opt.o.progress(opt.o.progress_ctx, src_pos, dst_pos);
        if (cur < b.nblk) {
            const int32_t r = b.out_len[cur];
            src_pos += b.rec_read[cur];
            if (r < 0) {
                cur = b.nblk;                           // nothing after a bad block is served (the error is sticky)
                if (r == PLZ4CU_E_BLOCKHASH) return PLZ4CU_Z_BLOCK_HASH;
                if (r == PLZ4CU_E_OVERFLOW) return PLZ4CU_Z_BLOCK_SIZE_OVERFLOW;
                if (r == PLZ4CU_E_STALL) return PLZ4CU_Z_ENGINE;    // an engine fault, not a verdict on the data
                return PLZ4CU_Z_DECOMPRESS;
            }
            cur_off = 0; cur_len = (size_t)r;
            dst_pos += r; content_acc += (uint64_t)r;
            have_block = true;
            run_first = cur++;
            // full blocks sit back to back in the output area: when nobody watches block boundaries, serve the whole
            // run of good blocks as one piece (one large copy / one sink call instead of one per block)
            if (!opt.o.progress) {
                for (int32_t last = r; last == bsz && cur < b.nblk && b.out_len[cur] >= 0; cur++) {
                    last = b.out_len[cur];
                    src_pos += b.rec_read[cur];
                    cur_len += (size_t)last; dst_pos += last; content_acc += (uint64_t)last;
                }
            }
            return 0;
        }
        const int ev = b.tail_event;
        b.tail_event = 0;
        if (ev == 2) {
            src_pos += b.endmark_read;
            if (verify_content_hash) {
                hash_q.wait(b.hash_ticket);
                if (b.digest != b.content_hash_read) return PLZ4CU_Z_CONTENT_HASH;
            }
            return 2;
        }
        src_pos += b.tail_read;
        return ev;
    }
    const uint8_t* block_ptr() { return cbatch().out.p + (size_t)run_first * bsz; }

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
                    bulk_copy_mt(dst + produced, block_ptr() + cur_off, k, 2);
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
                    const int64_t t4 = StageClock::on() ? StageClock::now() : 0;
                    int64_t k = w(wctx, block_ptr() + cur_off, cur_len - cur_off);
                    if (StageClock::on()) clk.us[3] += StageClock::now() - t4;
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
        quiesce();                                      // the source callback is not used after Close
        if (!carry.empty() && seek) { seek(ctx, -(int64_t)carry.size()); carry.clear(); }   // closed inside a body: give back what was read ahead
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
int64_t plz4cu_writer_read_from(plz4cu_writer_t* w, plz4cu_read_fn rd, void* rd_ctx) { return w->read_from(rd, rd_ctx); }
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

int plz4cu_frame_header(const plz4cu_opts_t* opts, uint8_t out[19])
{
    Opts o(opts);                                    // same normalisation as NewWriter (block size default, level clamp)
    const std::vector<uint8_t> h = make_header(o.o);
    memcpy(out, h.data(), h.size());
    return (int)h.size();
}

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
    bulk_copy_mt(buf, m->data + m->pos, k, 4);
    m->pos += k;
    return (int64_t)k;
}
int64_t plz4cu_membuf_write(void* ctx, const void* data, size_t n)
{
    plz4cu_membuf* m = static_cast<plz4cu_membuf*>(ctx);
    if (m->len + n > m->cap) return -1;
    bulk_copy_mt(m->data + m->len, data, n, 8);
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
