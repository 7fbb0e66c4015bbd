// engine.cu — the C ABI declared in include/plz4cu.h.
//
// Device-resident entry points are thin launches on the caller's stream.  Host-resident entry
// points lease a pipeline context of their device (grow-only device scratch, kLanes lanes of
// stream + buffers) so that H2D of chunk k+1, kernels of chunk k and D2H of chunk k-1 overlap.
// There is NO CPU codec in this library: without a usable GPU every entry point fails loudly.
#include "../../include/plz4cu.h"
#include "kernels.h"
#include "logtext.h"

#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>
#include <algorithm>
#include <unordered_map>

using namespace plz4;

static_assert(PLZ4CU_E_BLOCKHASH == PLZ4CU_E_BLOCKHASH_, "header/kernels mismatch");
static_assert(PLZ4CU_E_OVERFLOW == PLZ4CU_E_OVERFLOW_, "header/kernels mismatch");
static_assert(PLZ4CU_E_STALL == PLZ4CU_E_STALL_, "header/kernels mismatch");

void plz4cu_internal_copy(void* dst, const void* src, size_t n);    // host_stream.cu: copy by several threads

namespace {

std::mutex g_slab_mu;
std::unordered_map<void*, size_t> g_slab_size;                 // every live slab -> its size class
std::unordered_map<size_t, std::vector<void*>> g_slab_free;    // cached, not borrowed
size_t g_slab_cached = 0;
const size_t kSlabCacheMax = 4ull << 30;

thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};
std::atomic<int64_t> g_host_outstanding{0};

int fail(int code, const char* what, cudaError_t e = cudaSuccess)
{
    char buf[512];
    if (e != cudaSuccess) snprintf(buf, sizeof buf, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    else snprintf(buf, sizeof buf, "%s", what);
    g_err = buf;
    return code;
}
#define CU(call)                                                       \
    do {                                                               \
        cudaError_t e__ = (call);                                      \
        if (e__ != cudaSuccess) return fail(PLZ4CU_ERR_CUDA, #call, e__); \
    } while (0)

// host pipeline tunables (see "host-resident batches" below)
constexpr int kMaxLanes = 8;
int kLanes = 4;                                  // lanes in use                   (PLZ4CU_LANES)
bool g_stage_pageable = true;                    // pageable caller buffers go through pinned slabs, several threads (PLZ4CU_STAGE=0: straight cudaMemcpyAsync)
bool g_spin_wait = false;                        // busy-wait on the GPU instead of sleeping (PLZ4CU_SPIN, measurements)
uint32_t kChunkBlocks = 256;                     // blocks per chunk we aim for    (PLZ4CU_CHUNK_BLOCKS)
uint64_t kChunkMinBytes = 16ull << 20;           // ... but small payloads are gathered up to this many bytes (PLZ4CU_CHUNK_MIB)
const uint64_t kChunkBytes = 512ull << 20;       // upper bound on a chunk's input span

void read_tuning_env()
{
    static std::once_flag once;
    std::call_once(once, [] {
        if (const char* e = getenv("PLZ4CU_LANES")) { int v = atoi(e); if (v >= 1 && v <= kMaxLanes) kLanes = v; }
        if (const char* e = getenv("PLZ4CU_CHUNK_BLOCKS")) { int v = atoi(e); if (v >= 1) kChunkBlocks = (uint32_t)v; }
        if (const char* e = getenv("PLZ4CU_SPIN")) g_spin_wait = atoi(e) != 0;
        if (const char* e = getenv("PLZ4CU_STAGE")) g_stage_pageable = atoi(e) != 0;
        if (const char* e = getenv("PLZ4CU_CHUNK_MIB")) { int v = atoi(e); if (v >= 1) kChunkMinBytes = (uint64_t)v << 20; }
    });
}

int ensure_configured()
{
    read_tuning_env();
    // per-device function attributes must be set on every device we run on
    static std::mutex mu;
    static std::vector<int> done;
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return fail(PLZ4CU_ERR_NODEVICE, "no CUDA device (the engine has no CPU fallback)", e);
    std::lock_guard<std::mutex> lk(mu);
    if (std::find(done.begin(), done.end(), dev) != done.end()) return 0;
    e = configure_compress();
    if (e != cudaSuccess) return fail(PLZ4CU_ERR_CUDA, "configure_compress", e);
    e = configure_compress_cta();
    if (e != cudaSuccess) return fail(PLZ4CU_ERR_CUDA, "configure_compress_cta", e);
    e = configure_decompress();
    if (e != cudaSuccess) return fail(PLZ4CU_ERR_CUDA, "configure_decompress", e);
    done.push_back(dev);
    return 0;
}

// grow-only device buffer
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = std::max(n, (size_t)1 << 20);
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    template <typename T> T* as() { return reinterpret_cast<T*>(p); }
};
struct HostBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaHostAlloc(&p, n, cudaHostAllocPortable);
        if (e == cudaSuccess) cap = n;
        return e;
    }
    template <typename T> T* as() { return reinterpret_cast<T*>(p); }
};

// one pipeline lane: a stream plus its private scratch
// Is this host pointer page-locked (cudaHostAlloc / cudaHostRegister)?  Pageable buffers are staged through the lanes' pinned
// slabs by several host threads: a pageable cudaMemcpyAsync goes through the driver's own bounce buffer at 9-11 GB/s.
static bool host_pinned(const void* p)
{
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged || at.type == cudaMemoryTypeDevice;
}

struct Lane {
    cudaStream_t st = nullptr;
    cudaEvent_t done = nullptr;          // sizes of a compressed chunk are on the host
    cudaEvent_t fin = nullptr;           // everything issued on the lane for its chunk has completed
    bool fin_recorded = false;
    DevBuf in, out, packed, off, res, poff;
    HostBuf h_meta, h_res;      // pinned staging for small metadata (up / down)
    HostBuf h_in, h_out;        // pinned staging for a caller's pageable buffers (one chunk up, one chunk down)
    uint64_t out_pos = 0, out_bytes = 0;   // where the chunk's bytes go in the caller's buffer once they are down
};

struct Pipe {
    Lane lane[kMaxLanes];
    bool ready = false;
    int init()
    {
        if (ready) return 0;
        // The host waits on events that block instead of spinning: a waiting call costs no core, which is what lets
        // eight ranks (or many streams) share one host without starving each other's copy and hash threads.
        const unsigned flags = cudaEventDisableTiming | (g_spin_wait ? 0u : (unsigned)cudaEventBlockingSync);
        for (auto& l : lane) {
            CU(cudaStreamCreateWithFlags(&l.st, cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&l.done, flags));
            CU(cudaEventCreateWithFlags(&l.fin, flags));
        }
        ready = true;
        return 0;
    }
};

// Host-resident calls from different threads (many writers / readers, the reference's one-goroutine-per-stream use)
// each lease a whole pipe, so their copies and kernels overlap on the device; up to kMaxPipes per device, further
// callers queue for a free one.
constexpr int kMaxPipes = 8;
struct PipePool {
    std::mutex mu;
    std::condition_variable cv;
    std::vector<Pipe*> idle;
    int created = 0;
};
std::mutex g_pools_mu;
std::vector<PipePool*> g_pools;

class PipeLease {
    PipePool* pool = nullptr;
    Pipe* pipe = nullptr;
public:
    PipeLease()
    {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return;
        {
            std::lock_guard<std::mutex> lk(g_pools_mu);
            if ((int)g_pools.size() <= dev) g_pools.resize(dev + 1, nullptr);
            if (!g_pools[dev]) g_pools[dev] = new PipePool();
            pool = g_pools[dev];
        }
        std::unique_lock<std::mutex> lk(pool->mu);
        if (pool->idle.empty() && pool->created < kMaxPipes) { pool->created++; pipe = new Pipe(); return; }
        pool->cv.wait(lk, [&] { return !pool->idle.empty(); });
        pipe = pool->idle.back();
        pool->idle.pop_back();
    }
    ~PipeLease()
    {
        if (!pipe) return;
        { std::lock_guard<std::mutex> lk(pool->mu); pool->idle.push_back(pipe); }
        pool->cv.notify_one();
    }
    PipeLease(const PipeLease&) = delete;
    PipeLease& operator=(const PipeLease&) = delete;
    Pipe* get() const { return pipe; }
};

inline uint32_t round_up16(uint32_t v) { return (v + 15u) & ~15u; }

}  // namespace

struct plz4cu_dict {
    uint8_t* d_bytes = nullptr;     // device copy of the last <= 64 KiB (+ zeroed slack)
    uint16_t* d_tables = nullptr;   // encoder tables for 11, 12 and 13 hash bits, back to back
    uint32_t size = 0;
    std::vector<uint8_t> h_bytes;
    const uint16_t* table(int bits) const
    {
        return d_tables + (bits == 11 ? 0 : bits == 12 ? (1 << 11) : (1 << 11) + (1 << 12));
    }
};

// devices one stream may spread over (plz4cu_init_devices); read by host_stream.cu
static std::mutex g_devs_mu;
static std::vector<int> g_devs;
std::vector<int> plz4cu_internal_devices()
{
    std::lock_guard<std::mutex> lk(g_devs_mu);
    return g_devs;
}

extern "C" {

int plz4cu_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) return fail(PLZ4CU_ERR_NODEVICE, "no CUDA device (the engine has no CPU fallback)", e);
    return n;
}

int plz4cu_init(int device)
{
    int n = plz4cu_device_count();
    if (n < 0) return n;
    if (device < 0 || device >= n) return fail(PLZ4CU_ERR_ARG, "plz4cu_init: device out of range");
    CU(cudaSetDevice(device));
    CU(cudaFree(0));
    return ensure_configured();
}

int plz4cu_init_devices(int ndev, const int* devs)
{
    int n = plz4cu_device_count();
    if (n < 0) return n;
    if (ndev <= 0 || !devs) return fail(PLZ4CU_ERR_ARG, "plz4cu_init_devices: no devices given");
    std::vector<int> v;
    for (int i = 0; i < ndev; i++) {
        if (devs[i] < 0 || devs[i] >= n) return fail(PLZ4CU_ERR_ARG, "plz4cu_init_devices: device out of range");
        if (std::find(v.begin(), v.end(), devs[i]) != v.end()) return fail(PLZ4CU_ERR_ARG, "plz4cu_init_devices: device listed twice");
        v.push_back(devs[i]);
    }
    for (int d : v) {
        CU(cudaSetDevice(d));
        CU(cudaFree(0));
        if (int r = ensure_configured()) return r;
    }
    CU(cudaSetDevice(v[0]));
    std::lock_guard<std::mutex> lk(g_devs_mu);
    g_devs = v;
    return 0;
}
int plz4cu_registered_devices(void)
{
    std::lock_guard<std::mutex> lk(g_devs_mu);
    return (int)g_devs.size();
}

const char* plz4cu_last_error(void) { return g_err.c_str(); }
const char* plz4cu_version(void) { return "plz4cu 0.1 sm_100a"; }
uint64_t plz4cu_launch_count(void) { return g_launches.load(); }

size_t plz4cu_compress_bound(size_t n)
{
    if (n > 0x7E000000u) return 0;
    return n + n / 255 + 16;
}

// Pinned slabs are pooled by power-of-two size class: pinning is slow (~0.3 ms per MiB), and the reference hands its
// blocks out of sync.Pools for the same reason (blk/pool.go:22-27).  Freed slabs stay cached for the next borrower.
void* plz4cu_host_alloc(size_t n)
{
    size_t cls = 1u << 16;
    while (cls < n) cls <<= 1;
    {
        std::lock_guard<std::mutex> lk(g_slab_mu);
        auto it = g_slab_free.find(cls);
        if (it != g_slab_free.end() && !it->second.empty()) {
            void* p = it->second.back();
            it->second.pop_back();
            g_slab_cached -= cls;
            g_host_outstanding++;
            return p;
        }
    }
    void* p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, cls, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        plz4cu_host_trim();                      // give cached slabs back and retry once
        e = cudaHostAlloc(&p, cls, cudaHostAllocPortable);
    }
    if (e != cudaSuccess) { fail(PLZ4CU_ERR_NOMEM, "cudaHostAlloc", e); return nullptr; }
    {
        std::lock_guard<std::mutex> lk(g_slab_mu);
        g_slab_size[p] = cls;
    }
    g_host_outstanding++;
    return p;
}
void plz4cu_host_free(void* p)
{
    if (!p) return;
    g_host_outstanding--;
    std::lock_guard<std::mutex> lk(g_slab_mu);
    auto it = g_slab_size.find(p);
    if (it == g_slab_size.end()) { cudaFreeHost(p); return; }
    const size_t cls = it->second;
    if (g_slab_cached + cls > kSlabCacheMax) {
        g_slab_size.erase(it);
        cudaFreeHost(p);
        return;
    }
    g_slab_free[cls].push_back(p);
    g_slab_cached += cls;
}
void plz4cu_host_trim(void)
{
    std::lock_guard<std::mutex> lk(g_slab_mu);
    for (auto& kv : g_slab_free) {
        for (void* p : kv.second) { g_slab_size.erase(p); cudaFreeHost(p); }
        kv.second.clear();
    }
    g_slab_cached = 0;
}
int64_t plz4cu_host_outstanding(void) { return g_host_outstanding.load(); }

void* plz4cu_device_alloc(size_t n)
{
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, n ? n : 1);
    if (e != cudaSuccess) { fail(PLZ4CU_ERR_NOMEM, "cudaMalloc", e); return nullptr; }
    return p;
}
void plz4cu_device_free(void* p) { if (p) cudaFree(p); }

plz4cu_dict_t* plz4cu_dict_create(const void* d, size_t n)
{
    const uint8_t* b = static_cast<const uint8_t*>(d);
    if (n > PLZ4CU_DICT_MAX) { b += n - PLZ4CU_DICT_MAX; n = PLZ4CU_DICT_MAX; }   // compress/dict.go:43-56
    plz4cu_dict* dc = new plz4cu_dict();
    dc->size = (uint32_t)n;
    dc->h_bytes.assign(b, b + n);
    if (n) {
        // 32 bytes of zeroed slack: the encoder fetches aligned 16-byte chunks around a candidate
        cudaError_t e = cudaMalloc((void**)&dc->d_bytes, n + 48);
        if (e == cudaSuccess) e = cudaMemset(dc->d_bytes, 0, n + 48);
        if (e == cudaSuccess) e = cudaMemcpy(dc->d_bytes, b, n, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMalloc((void**)&dc->d_tables, sizeof(uint16_t) * ((1 << 11) + (1 << 12) + (1 << 13)));
        for (int bits = 11; bits <= 13 && e == cudaSuccess; bits++) {
            e = launch_dict_build(dc->d_bytes, dc->size, bits, const_cast<uint16_t*>(dc->table(bits)), nullptr);
            g_launches++;
        }
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            fail(PLZ4CU_ERR_CUDA, "plz4cu_dict_create", e);
            if (dc->d_bytes) cudaFree(dc->d_bytes);
            if (dc->d_tables) cudaFree(dc->d_tables);
            delete dc;
            return nullptr;
        }
    }
    return dc;
}
void plz4cu_dict_destroy(plz4cu_dict_t* dc)
{
    if (!dc) return;
    if (dc->d_bytes) cudaFree(dc->d_bytes);
    if (dc->d_tables) cudaFree(dc->d_tables);
    delete dc;
}

int plz4cu_compress_batch_device(plz4cu_stream_t stream, const void* src_base, const uint64_t* src_off,
                                 const uint32_t* src_len, uint32_t nblk, uint32_t dst_cap, int block_checksum,
                                 int raw_blocks, const plz4cu_dict_t* dict, void* rec_base, uint32_t rec_stride,
                                 uint32_t* rec_len)
{
    if (int r = ensure_configured()) return r;
    if (nblk == 0) return 0;
    if (!src_base || !src_off || !src_len || !rec_base || !rec_len) return fail(PLZ4CU_ERR_ARG, "compress_batch_device: null pointer");
    if (rec_stride & 15u) return fail(PLZ4CU_ERR_ARG, "compress_batch_device: rec_stride must be a multiple of 16");
    if ((reinterpret_cast<uintptr_t>(rec_base) & 15u) != 0) return fail(PLZ4CU_ERR_ARG, "compress_batch_device: rec_base must be 16-byte aligned");
    if ((uint64_t)rec_stride < (uint64_t)dst_cap + (raw_blocks ? 0 : PLZ4CU_REC_OVERHEAD)) return fail(PLZ4CU_ERR_ARG, "compress_batch_device: rec_stride too small for dst_cap");
    EncodeArgs a{};
    a.src_base = static_cast<const uint8_t*>(src_base);
    a.src_off = src_off; a.src_len = src_len; a.nblk = nblk; a.dst_cap = dst_cap;
    a.block_checksum = block_checksum; a.raw_blocks = raw_blocks;
    a.rec_base = static_cast<uint8_t*>(rec_base); a.rec_stride = rec_stride; a.rec_len = rec_len;
    // frame path: a block is at most the frame's block size (dst_cap); raw path: only the slot bounds it
    a.max_src_len = raw_blocks ? rec_stride : dst_cap;
    if (dict && dict->size) { a.dict = dict->d_bytes; a.dict_size = dict->size; a.dict_table = dict->table(compress_hash_bits(dst_cap)); }
    CU(launch_compress(a, static_cast<cudaStream_t>(stream)));
    g_launches++;
    return 0;
}

int plz4cu_decompress_batch_device(plz4cu_stream_t stream, const void* rec_base, const uint64_t* rec_off,
                                   const uint32_t* raw_len, uint32_t nblk, uint32_t dst_cap, int verify_checksum,
                                   int raw_blocks, const plz4cu_dict_t* dict, void* dst_base, uint64_t dst_stride,
                                   int32_t* out_len)
{
    if (int r = ensure_configured()) return r;
    if (nblk == 0) return 0;
    if (!rec_base || !rec_off || !dst_base || !out_len) return fail(PLZ4CU_ERR_ARG, "decompress_batch_device: null pointer");
    if (raw_blocks && !raw_len) return fail(PLZ4CU_ERR_ARG, "decompress_batch_device: raw_len required for raw blocks");
    if (dst_stride < dst_cap) return fail(PLZ4CU_ERR_ARG, "decompress_batch_device: dst_stride < dst_cap");
    DecodeArgs a{};
    a.rec_base = static_cast<const uint8_t*>(rec_base);
    a.rec_off = rec_off; a.raw_len = raw_len; a.nblk = nblk; a.dst_cap = dst_cap;
    a.verify_checksum = verify_checksum; a.raw_blocks = raw_blocks;
    a.dict = (dict && dict->size) ? dict->d_bytes : nullptr;
    a.dict_size = dict ? dict->size : 0;
    a.dst_base = static_cast<uint8_t*>(dst_base); a.dst_stride = dst_stride; a.out_len = out_len;
    CU(launch_decompress(a, static_cast<cudaStream_t>(stream)));
    g_launches++;
    return 0;
}

int plz4cu_pack_records_device(plz4cu_stream_t stream, const void* rec_base, uint32_t rec_stride,
                               const uint32_t* rec_len, uint32_t nblk, void* packed, uint64_t* packed_off)
{
    if (int r = ensure_configured()) return r;
    if (!packed_off) return fail(PLZ4CU_ERR_ARG, "pack_records_device: null pointer");
    CU(launch_pack(static_cast<const uint8_t*>(rec_base), rec_stride, rec_len, nblk,
                   static_cast<uint8_t*>(packed), packed_off, static_cast<cudaStream_t>(stream)));
    g_launches += nblk ? 2 : 1;
    return 0;
}

int plz4cu_xxh32_batch_device(plz4cu_stream_t stream, const void* base, const uint64_t* off, const uint32_t* len,
                              uint32_t nblk, uint32_t* out)
{
    if (int r = ensure_configured()) return r;
    CU(launch_xxh32(static_cast<const uint8_t*>(base), off, len, nblk, out, static_cast<cudaStream_t>(stream)));
    if (nblk) g_launches++;
    return 0;
}

int plz4cu_gen_logtext_device(plz4cu_stream_t stream, uint32_t seed, uint64_t first_seg, void* dst, uint64_t n)
{
    if (int r = ensure_configured()) return r;
    CU(launch_gen_logtext(seed, first_seg, static_cast<uint8_t*>(dst), n, static_cast<cudaStream_t>(stream)));
    if (n) g_launches++;
    return 0;
}

int plz4cu_gen_logtext_host(uint32_t seed, uint64_t first_seg, void* dst, uint64_t n)
{
    uint8_t* out = static_cast<uint8_t*>(dst);
    for (uint64_t s = 0, pos = 0; pos < n; s++, pos += LOGTEXT_SEG) {
        uint32_t len = (n - pos < LOGTEXT_SEG) ? (uint32_t)(n - pos) : LOGTEXT_SEG;
        lt_fill_segment(seed, first_seg + s, out + pos, len);
    }
    return 0;
}

// ---------------------------------------------------------------- device-resident frames

int plz4cu_frame_index_device(plz4cu_stream_t stream, const void* body, uint64_t len, uint32_t block_size, int block_checksum,
                              uint64_t* rec_off, uint32_t cap, uint64_t* nblk, uint64_t* end_off)
{
    if (int r = ensure_configured()) return r;
    if (!body || !rec_off || !nblk || !end_off || block_size == 0) return fail(PLZ4CU_ERR_ARG, "frame_index_device: null pointer");
    FrameIndexResult res{};
    uint64_t launches = 0;
    cudaError_t e = launch_frame_index(static_cast<const uint8_t*>(body), len, block_size, block_checksum, rec_off, cap, &res,
                                       &launches, static_cast<cudaStream_t>(stream));
    g_launches += launches;
    if (e != cudaSuccess) return fail(PLZ4CU_ERR_CUDA, "frame_index_device", e);
    *nblk = res.nblk;
    *end_off = res.stop_off + (res.why == 1 ? 4 : 0);
    switch (res.why) {
    case 1: break;
    case 2: return PLZ4CU_Z_BLOCK_SIZE_OVERFLOW;
    case 3: return PLZ4CU_Z_BLOCK_READ;
    default: return PLZ4CU_Z_BLOCK_SIZE_READ;
    }
    if (res.nblk > cap) return fail(PLZ4CU_ERR_ARG, "frame_index_device: more blocks than rec_off can hold");
    return 0;
}

int plz4cu_decompress_frame_device(plz4cu_stream_t stream, const void* frame, uint64_t frame_len, const plz4cu_dict_t* dict,
                                   void* dst, uint64_t dst_cap, uint64_t* rec_off, int32_t* out_len, uint32_t cap,
                                   plz4cu_frame_info_t* info)
{
    if (int r = ensure_configured()) return r;
    if (!frame || !info) return fail(PLZ4CU_ERR_ARG, "decompress_frame_device: null pointer");
    memset(info, 0, sizeof *info);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const uint8_t* f = static_cast<const uint8_t*>(frame);

    // header/read.go:26-119 on the first <= 19 bytes
    uint8_t h[19] = {0};
    const size_t hn = (size_t)std::min<uint64_t>(frame_len, sizeof h);
    CU(cudaMemcpyAsync(h, f, hn, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (hn < 7) return PLZ4CU_Z_HEADER_READ;
    static const uint8_t magic[4] = {0x04, 0x22, 0x4d, 0x18};
    if (memcmp(h, magic, 4) != 0) return PLZ4CU_Z_MAGIC;
    const uint8_t flags = h[4], bd = h[5];
    if (((flags >> 6) & 3) != 1) return PLZ4CU_Z_VERSION;
    if (flags & 0x02) return PLZ4CU_Z_RESERVE_BIT;
    if (((bd >> 4) & 7) < 4 || (bd & 0x80) || (bd & 0x0F)) return PLZ4CU_Z_BLOCK_DESCRIPTOR;
    size_t n = 7;
    if (flags & 0x08) n += 8;
    if (flags & 0x01) n += 4;
    if (hn < n) return PLZ4CU_Z_HEADER_READ;
    if (flags & 0x08) for (int i = 0; i < 8; i++) info->content_size |= (uint64_t)h[6 + i] << (8 * i);
    if (flags & 0x01) memcpy(&info->dict_id, h + n - 5, 4);
    if (((plz4cu_xxh32_host(h + 4, n - 5) >> 8) & 0xFF) != h[n - 1]) return PLZ4CU_Z_HEADER_HASH;
    info->block_size = 1u << (8 + 2 * ((bd >> 4) & 7));                 // 4..7 -> 64 KiB .. 4 MiB
    info->header_len = (uint32_t)n;
    info->block_checksum = (flags >> 4) & 1;
    info->content_checksum = (flags >> 2) & 1;
    info->has_content_size = (flags >> 3) & 1;
    info->has_dict_id = flags & 1;
    if (!(flags & 0x20)) return PLZ4CU_Z_UNSUPPORTED;                   // linked blocks decode in order on CPU cores

    uint64_t nblk = 0, end_off = 0;
    int rc;
    if (!rec_off || !out_len || cap == 0) {
        // sizing call: count the blocks only
        uint64_t* one = nullptr;
        CU(cudaMalloc((void**)&one, sizeof(uint64_t)));
        rc = plz4cu_frame_index_device(stream, f + n, frame_len - n, info->block_size, info->block_checksum, one, 0, &nblk, &end_off);
        cudaFree(one);
        cap = 0;
    } else {
        rc = plz4cu_frame_index_device(stream, f + n, frame_len - n, info->block_size, info->block_checksum, rec_off, cap,
                                       &nblk, &end_off);
    }
    info->nblk = nblk;
    if (rc == PLZ4CU_ERR_CUDA) return rc;
    // a broken walk is reported AFTER the blocks before the break were decoded and checked: the first error in
    // stream order wins, as with the host reader (a wrong size word usually shows up as that block's hash mismatch)
    const int walk_rc = (rc == PLZ4CU_ERR_ARG) ? 0 : rc;
    if (walk_rc == 0) {
        info->frame_len = n + end_off + (info->content_checksum ? 4 : 0);
        if (info->frame_len > frame_len) return PLZ4CU_Z_CONTENT_HASH_READ;
        if (info->content_checksum) {
            uint8_t c[4];
            CU(cudaMemcpyAsync(c, f + n + end_off, 4, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            memcpy(&info->content_hash, c, 4);
        }
    }
    if (nblk == 0) { info->contiguous = 1; return walk_rc; }
    if (nblk > cap) return fail(PLZ4CU_ERR_ARG, "decompress_frame_device: scratch holds fewer entries than the frame has blocks");
    if (!dst || nblk * (uint64_t)info->block_size > dst_cap) return fail(PLZ4CU_ERR_ARG, "decompress_frame_device: dst too small");
    rc = plz4cu_decompress_batch_device(stream, f + n, rec_off, nullptr, (uint32_t)nblk, info->block_size, info->block_checksum, 0,
                                        dict, dst, info->block_size, out_len);
    if (rc < 0) return rc;
    std::vector<int32_t> lens(nblk);
    CU(cudaMemcpyAsync(lens.data(), out_len, nblk * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    info->contiguous = 1;
    for (uint64_t b = 0; b < nblk; b++) {
        const int32_t r = lens[b];
        if (r < 0) {
            if (r == PLZ4CU_E_BLOCKHASH) return PLZ4CU_Z_BLOCK_HASH;
            if (r == PLZ4CU_E_OVERFLOW) return PLZ4CU_Z_BLOCK_SIZE_OVERFLOW;
            if (r == PLZ4CU_E_STALL) return fail(PLZ4CU_Z_ENGINE, "decompress_frame_device: a decode team stalled");
            return PLZ4CU_Z_DECOMPRESS;
        }
        info->out_bytes += (uint64_t)r;
        if (b + 1 < nblk && (uint32_t)r != info->block_size) info->contiguous = 0;
    }
    return walk_rc;
}

int plz4cu_compress_frame_device(plz4cu_stream_t stream, const void* src, uint64_t n, const plz4cu_opts_t* opts,
                                 const plz4cu_dict_t* dict, void* frame, uint64_t frame_cap, uint64_t* frame_len)
{
    if (int r = ensure_configured()) return r;
    if (!frame || !frame_len || (!src && n)) return fail(PLZ4CU_ERR_ARG, "compress_frame_device: null pointer");
    plz4cu_opts_t o;
    if (opts) o = *opts; else plz4cu_opts_default(&o);
    if (o.level > 1 || o.block_linked || o.content_checksum) return PLZ4CU_Z_UNSUPPORTED;
    uint8_t hdr[19];
    const uint64_t hn = (uint64_t)plz4cu_frame_header(&o, hdr);
    const uint32_t bsz = 1u << (8 + 2 * ((hdr[5] >> 4) & 7));
    if (n / bsz >= 0xFFFFFFF0ull) return fail(PLZ4CU_ERR_ARG, "compress_frame_device: too many blocks for one call");
    const uint32_t nblk = (uint32_t)((n + bsz - 1) / bsz);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint8_t* out = static_cast<uint8_t*>(frame);
    if (frame_cap < hn + 4) return fail(PLZ4CU_ERR_ARG, "compress_frame_device: frame buffer too small");
    CU(cudaMemcpyAsync(out, hdr, hn, cudaMemcpyHostToDevice, st));
    uint64_t total = 0;
    if (nblk) {
        // scratch: block table, record slots, record lengths, packed offsets
        const uint32_t stride = round_up16(bsz + PLZ4CU_REC_OVERHEAD);
        const uint64_t tbl_bytes = ((uint64_t)nblk * 12 + 15) & ~15ull, len_bytes = ((uint64_t)nblk * 4 + 15) & ~15ull;
        const uint64_t bytes = tbl_bytes + (uint64_t)nblk * stride + len_bytes + (uint64_t)(nblk + 1) * 8;
        uint8_t* scratch = nullptr;
        CU(cudaMallocAsync((void**)&scratch, bytes, st));
        uint64_t* d_off = reinterpret_cast<uint64_t*>(scratch);
        uint32_t* d_len = reinterpret_cast<uint32_t*>(d_off + nblk);
        uint8_t* d_rec = scratch + tbl_bytes;
        uint32_t* d_rl = reinterpret_cast<uint32_t*>(d_rec + (uint64_t)nblk * stride);
        uint64_t* d_poff = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(d_rl) + len_bytes);
        std::vector<uint64_t> tbl((tbl_bytes + 7) / 8);
        uint32_t* h_len = reinterpret_cast<uint32_t*>(tbl.data() + nblk);
        for (uint32_t b = 0; b < nblk; b++) { tbl[b] = (uint64_t)b * bsz; h_len[b] = (uint32_t)std::min<uint64_t>(bsz, n - tbl[b]); }
        int rc = 0;
        cudaError_t e = cudaMemcpyAsync(scratch, tbl.data(), (uint64_t)nblk * 12, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) {
            rc = plz4cu_compress_batch_device(stream, src, d_off, d_len, nblk, bsz, o.block_checksum, 0, dict, d_rec, stride, d_rl);
            if (rc == 0) e = launch_scan_u32(d_rl, nblk, d_poff, st);
        }
        if (e == cudaSuccess && rc == 0) e = cudaMemcpyAsync(&total, d_poff + nblk, 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && rc == 0) e = cudaStreamSynchronize(st);
        if (e == cudaSuccess && rc == 0) {
            if (hn + total + 4 > frame_cap) rc = fail(PLZ4CU_ERR_ARG, "compress_frame_device: frame buffer too small");
            else e = launch_pack(d_rec, stride, d_rl, nblk, out + hn, d_poff, st);
        }
        g_launches += 3;
        cudaFreeAsync(scratch, st);
        if (e != cudaSuccess) return fail(PLZ4CU_ERR_CUDA, "compress_frame_device", e);
        if (rc < 0) return rc;
    }
    CU(cudaMemsetAsync(out + hn + total, 0, 4, st));                  // EndMark (trailer/trailer.go:10-19)
    CU(cudaStreamSynchronize(st));
    *frame_len = hn + total + 4;
    return 0;
}

// ---------------------------------------------------------------- host-resident batches
//
// Blocks are cut into chunks and software-pipelined over kLanes lanes (stream + private scratch
// each), so that the H2D of chunk k, the kernels of chunk k-1 and the D2H of chunk k-2 overlap.
// A chunk is 16 MiB and 256 blocks at least: the encoder gives every block (or span of a large block) a CTA that
// is done in a fraction of a millisecond, so a launch of that size fills the GPU for long enough, chunks of
// different lanes run side by side, and a 64 MiB batch of a stream already pipelines over the lanes (it was one
// chunk when a warp worked on a block for milliseconds and a chunk had to bring thousands of blocks).
// All small metadata crosses PCIe through pinned staging, so no call in the loop blocks the host
// except the explicit waits.

namespace {
struct Chunk { uint32_t b0, b1; uint64_t lo, hi; };     // blocks [b0,b1), input byte span [lo,hi)
}

int plz4cu_compress_batch_host(const void* src, const uint64_t* src_off, const uint32_t* src_len, uint32_t nblk,
                               uint32_t dst_cap, int block_checksum, int raw_blocks, const plz4cu_dict_t* dict,
                               void* packed, uint64_t packed_cap, uint64_t* packed_off)
{
    if (int r = ensure_configured()) return r;
    if (!packed_off) return fail(PLZ4CU_ERR_ARG, "compress_batch_host: null pointer");
    packed_off[0] = 0;
    if (nblk == 0) return 0;
    if (!src || !src_off || !src_len || !packed) return fail(PLZ4CU_ERR_ARG, "compress_batch_host: null pointer");
    PipeLease lease;
    Pipe* pp = lease.get();
    if (!pp) return fail(PLZ4CU_ERR_NODEVICE, "no CUDA device");
    if (int r = pp->init()) return r;

    const uint8_t* hsrc = static_cast<const uint8_t*>(src);
    uint8_t* hout = static_cast<uint8_t*>(packed);
    uint32_t max_len = 0;
    uint64_t total_len = 0;
    uint32_t min_len = 0xFFFFFFFFu;
    for (uint32_t b = 0; b < nblk; b++) { max_len = std::max(max_len, src_len[b]); min_len = std::min(min_len, src_len[b]); total_len += src_len[b]; }
    // a chunk is an eighth of the call, between kChunkMinBytes (16 MiB: a stream's 64 MiB batch still pipelines over the
    // lanes) and 64 MiB (a whole buffer is not cut finer than its copies need: every chunk costs two host waits)
    const uint64_t chunk_min = std::min<uint64_t>(std::max<uint64_t>(total_len / 8, kChunkMinBytes), std::max<uint64_t>(kChunkMinBytes, 64ull << 20));
    const uint32_t slot_payload = std::max(dst_cap, max_len);
    const uint32_t stride = round_up16(slot_payload + (raw_blocks ? 0 : PLZ4CU_REC_OVERHEAD));

    // blocks a chunk should hold at least: 256 of up to 64 KiB; large blocks are spans of 1 MiB to the encoder, so 32 MiB of
    // them fill the lanes as well, and a stream's batch of 4 MiB blocks (32 of them) pipelines instead of going up, through
    // the kernels and down in one piece
    const uint32_t chunk_blocks = std::max<uint32_t>(1, std::min<uint64_t>(kChunkBlocks, (32ull << 20) / std::max<uint32_t>(max_len, 1)));
    std::vector<Chunk> chunks;
    for (uint32_t b = 0; b < nblk;) {
        Chunk c{b, b, src_off[b], src_off[b] + src_len[b]};
        while (c.b1 < nblk) {
            uint64_t lo = std::min(c.lo, src_off[c.b1]), hi = std::max(c.hi, src_off[c.b1] + src_len[c.b1]);
            if (c.b1 > c.b0 && (hi - lo > kChunkBytes || (c.b1 - c.b0 >= chunk_blocks && c.hi - c.lo >= chunk_min))) break;
            c.lo = lo; c.hi = hi; c.b1++;
        }
        chunks.push_back(c);
        b = c.b1;
    }
    const int nchunks = (int)chunks.size();
    uint64_t out_pos = 0;
    const bool stage_in = g_stage_pageable && !host_pinned(hsrc), stage_out = g_stage_pageable && !host_pinned(hout);
    // a call of a block or two (the per-block shims): the caller waits for exactly this work, so the host spins on the stream
    // instead of sleeping on an event (a blocking wait wakes up 50-100 us late, twice per call)
    const bool tiny = nchunks == 1 && total_len <= (1u << 20);

    // stage 1: inputs up, kernels, packed offsets down
    auto issue = [&](int k) -> int {
        Lane& L = pp->lane[k % kLanes];
        const Chunk& c = chunks[k];
        const uint32_t cnt = c.b1 - c.b0;
        const uint64_t span = c.hi - c.lo;
        CU(L.in.reserve(span + 16));
        CU(L.out.reserve((uint64_t)cnt * stride));
        CU(L.packed.reserve((uint64_t)cnt * stride));
        CU(L.off.reserve((uint64_t)cnt * 12));
        CU(L.res.reserve((uint64_t)cnt * 4));
        CU(L.poff.reserve((uint64_t)(cnt + 1) * 8));
        CU(L.h_meta.reserve((uint64_t)cnt * 12));
        CU(L.h_res.reserve((uint64_t)(cnt + 1) * 8));
        uint64_t* h_rel = L.h_meta.as<uint64_t>();
        uint32_t* h_len = reinterpret_cast<uint32_t*>(h_rel + cnt);
        for (uint32_t i = 0; i < cnt; i++) { h_rel[i] = src_off[c.b0 + i] - c.lo; h_len[i] = src_len[c.b0 + i]; }
        const uint8_t* up = hsrc + c.lo;
        if (stage_in) {
            CU(L.h_in.reserve(span));
            plz4cu_internal_copy(L.h_in.p, up, span);
            up = L.h_in.as<uint8_t>();
        }
        CU(cudaMemcpyAsync(L.in.p, up, span, cudaMemcpyHostToDevice, L.st));
        CU(cudaMemcpyAsync(L.off.p, h_rel, (uint64_t)cnt * 12, cudaMemcpyHostToDevice, L.st));
        EncodeArgs a{};
        a.src_base = L.in.as<uint8_t>(); a.src_off = L.off.as<uint64_t>();
        a.src_len = reinterpret_cast<const uint32_t*>(L.off.as<uint64_t>() + cnt);
        a.nblk = cnt; a.dst_cap = dst_cap; a.block_checksum = block_checksum; a.raw_blocks = raw_blocks;
        a.rec_base = L.out.as<uint8_t>(); a.rec_stride = stride; a.rec_len = L.res.as<uint32_t>();
        a.max_src_len = max_len; a.min_src_len = min_len;
        if (dict && dict->size) { a.dict = dict->d_bytes; a.dict_size = dict->size; a.dict_table = dict->table(compress_hash_bits(dst_cap)); }
        CU(launch_compress(a, L.st));
        CU(launch_pack(L.out.as<uint8_t>(), stride, L.res.as<uint32_t>(), cnt, L.packed.as<uint8_t>(), L.poff.as<uint64_t>(), L.st));
        g_launches += 3;
        CU(cudaMemcpyAsync(L.h_res.p, L.poff.p, (uint64_t)(cnt + 1) * 8, cudaMemcpyDeviceToHost, L.st));
        CU(cudaEventRecord(L.done, L.st));
        L.fin_recorded = false;
        return 0;
    };
    // stage 2: once the sizes are known on the host, bring the packed records down
    auto post = [&](int k) -> int {
        Lane& L = pp->lane[k % kLanes];
        const Chunk& c = chunks[k];
        const uint32_t cnt = c.b1 - c.b0;
        if (tiny) CU(cudaStreamSynchronize(L.st)); else CU(cudaEventSynchronize(L.done));
        const uint64_t* hoff = L.h_res.as<uint64_t>();
        const uint64_t total = hoff[cnt];
        if (out_pos + total > packed_cap) return fail(PLZ4CU_ERR_ARG, "compress_batch_host: packed buffer too small");
        uint8_t* down = hout + out_pos;
        L.out_pos = out_pos; L.out_bytes = stage_out ? total : 0;
        if (stage_out) { CU(L.h_out.reserve(total)); down = L.h_out.as<uint8_t>(); }
        CU(cudaMemcpyAsync(down, L.packed.p, total, cudaMemcpyDeviceToHost, L.st));
        CU(cudaEventRecord(L.fin, L.st));
        L.fin_recorded = true;
        for (uint32_t i = 0; i <= cnt; i++) packed_off[c.b0 + i] = out_pos + hoff[i];
        out_pos += total;
        return 0;
    };
    auto finish = [&](int k) -> int {
        Lane& L = pp->lane[k % kLanes];
        if (L.fin_recorded && !tiny) CU(cudaEventSynchronize(L.fin));
        else CU(cudaStreamSynchronize(L.st));             // a tiny call, or the error path: the chunk never got as far as post()
        if (L.fin_recorded && L.out_bytes) { plz4cu_internal_copy(hout + L.out_pos, L.h_out.p, L.out_bytes); L.out_bytes = 0; }
        return 0;
    };
    // chunk j is issued at step j, its packed bytes are requested at step j + kLanes - 1 (so kLanes - 1 younger
    // chunks keep the GPU busy while the host waits for j's sizes), and its lane is recycled at step j + kLanes
    int rc = 0, posted = 0;
    for (int k = 0; k < nchunks && !rc; k++) {
        if (k >= kLanes) rc = finish(k - kLanes);
        if (!rc) rc = issue(k);
        if (!rc && k >= kLanes - 1) { rc = post(k - (kLanes - 1)); posted = k - (kLanes - 1) + 1; }
    }
    for (; posted < nchunks && !rc; posted++) rc = post(posted);
    for (int k = std::max(0, nchunks - kLanes); k < nchunks; k++) { int r2 = finish(k); if (!rc) rc = r2; }
    return rc;
}

int plz4cu_decompress_batch_host(const void* recs, uint64_t recs_bytes, const uint64_t* rec_off, const uint32_t* raw_len,
                                 uint32_t nblk, uint32_t dst_cap, int verify_checksum, int raw_blocks,
                                 const plz4cu_dict_t* dict, void* dst, uint64_t dst_stride, int32_t* out_len)
{
    if (int r = ensure_configured()) return r;
    if (nblk == 0) return 0;
    if (!recs || !rec_off || !dst || !out_len) return fail(PLZ4CU_ERR_ARG, "decompress_batch_host: null pointer");
    if (raw_blocks && !raw_len) return fail(PLZ4CU_ERR_ARG, "decompress_batch_host: raw_len required for raw blocks");
    if (dst_stride < dst_cap) return fail(PLZ4CU_ERR_ARG, "decompress_batch_host: dst_stride < dst_cap");
    PipeLease lease;
    Pipe* pp = lease.get();
    if (!pp) return fail(PLZ4CU_ERR_NODEVICE, "no CUDA device");
    if (int r = pp->init()) return r;

    const uint8_t* hrec = static_cast<const uint8_t*>(recs);
    uint8_t* hdst = static_cast<uint8_t*>(dst);

    // extent of record b in the input: needs the size word on the frame path (host-side read)
    auto rec_extent = [&](uint32_t b, uint64_t* lo, uint64_t* hi) -> int {
        uint64_t o = rec_off[b];
        if (raw_blocks) { *lo = o; *hi = o + raw_len[b]; }
        else {
            if (o + 4 > recs_bytes) return fail(PLZ4CU_ERR_ARG, "decompress_batch_host: record offset outside input");
            uint32_t w; memcpy(&w, hrec + o, 4);
            const uint64_t sz = (uint64_t)(w & 0x7FFFFFFFu);
            *lo = o;
            // an oversized size word is reported per block by the kernel from the word alone
            *hi = (sz > dst_cap) ? o + 4 : o + 4 + sz + (verify_checksum ? 4 : 0);
        }
        if (*hi > recs_bytes) return fail(PLZ4CU_ERR_ARG, "decompress_batch_host: record runs past the input (short read)");
        return 0;
    };

    std::vector<Chunk> chunks;
    const uint64_t total_out = (uint64_t)nblk * dst_cap;
    const uint64_t chunk_min = std::min<uint64_t>(std::max<uint64_t>(total_out / 8, kChunkMinBytes), std::max<uint64_t>(kChunkMinBytes, 64ull << 20));
    const uint32_t max_blk_per_chunk = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(kChunkBlocks, chunk_min / std::max<uint32_t>(dst_cap, 1)),
                                                                    std::max<uint64_t>(1, kChunkBytes / std::max<uint32_t>(dst_cap, 1)));
    for (uint32_t b = 0; b < nblk;) {
        uint64_t lo, hi;
        if (int r = rec_extent(b, &lo, &hi)) return r;
        Chunk c{b, b + 1, lo, hi};
        while (c.b1 < nblk && c.b1 - c.b0 < max_blk_per_chunk) {
            uint64_t l2, h2;
            if (int r = rec_extent(c.b1, &l2, &h2)) return r;
            uint64_t nlo = std::min(c.lo, l2), nhi = std::max(c.hi, h2);
            if (nhi - nlo > kChunkBytes) break;
            c.lo = nlo; c.hi = nhi; c.b1++;
        }
        chunks.push_back(c);
        b = c.b1;
    }
    const int nchunks = (int)chunks.size();
    const uint64_t dstride = ((uint64_t)dst_cap + 15u) & ~15ull;
    const bool stage_in = g_stage_pageable && !host_pinned(hrec), stage_out = g_stage_pageable && !host_pinned(hdst);
    const bool tiny = nchunks == 1 && total_out <= (1u << 20);      // (see compress_batch_host)

    auto issue = [&](int k) -> int {
        Lane& L = pp->lane[k % kLanes];
        const Chunk& c = chunks[k];
        const uint32_t cnt = c.b1 - c.b0;
        const uint64_t span = c.hi - c.lo;
        CU(L.in.reserve(span + 16));
        CU(L.out.reserve((uint64_t)cnt * dstride + 16));
        CU(L.off.reserve((uint64_t)cnt * 12));
        CU(L.res.reserve((uint64_t)cnt * 4));
        CU(L.h_meta.reserve((uint64_t)cnt * 12));
        CU(L.h_res.reserve((uint64_t)cnt * 4));
        uint64_t* h_rel = L.h_meta.as<uint64_t>();
        uint32_t* h_len = reinterpret_cast<uint32_t*>(h_rel + cnt);
        for (uint32_t i = 0; i < cnt; i++) { h_rel[i] = rec_off[c.b0 + i] - c.lo; h_len[i] = raw_blocks ? raw_len[c.b0 + i] : 0u; }
        const uint8_t* up = hrec + c.lo;
        if (stage_in) {
            CU(L.h_in.reserve(span));
            plz4cu_internal_copy(L.h_in.p, up, span);
            up = L.h_in.as<uint8_t>();
        }
        CU(cudaMemcpyAsync(L.in.p, up, span, cudaMemcpyHostToDevice, L.st));
        CU(cudaMemcpyAsync(L.off.p, h_rel, (uint64_t)cnt * 12, cudaMemcpyHostToDevice, L.st));
        DecodeArgs a{};
        a.rec_base = L.in.as<uint8_t>(); a.rec_off = L.off.as<uint64_t>();
        a.raw_len = reinterpret_cast<const uint32_t*>(L.off.as<uint64_t>() + cnt);
        a.nblk = cnt; a.dst_cap = dst_cap; a.verify_checksum = verify_checksum; a.raw_blocks = raw_blocks;
        a.dict = (dict && dict->size) ? dict->d_bytes : nullptr; a.dict_size = dict ? dict->size : 0;
        a.dst_base = L.out.as<uint8_t>(); a.dst_stride = dstride; a.out_len = L.res.as<int32_t>();
        CU(launch_decompress(a, L.st));
        g_launches++;
        CU(cudaMemcpyAsync(L.h_res.p, L.res.p, (uint64_t)cnt * 4, cudaMemcpyDeviceToHost, L.st));
        if (stage_out) {
            // down into the lane's pinned slab in one piece; finish() spreads the slots over the caller's buffer
            CU(L.h_out.reserve((uint64_t)cnt * dstride));
            CU(cudaMemcpyAsync(L.h_out.p, L.out.p, (uint64_t)cnt * dstride, cudaMemcpyDeviceToHost, L.st));
        } else if (dstride == dst_stride) {
            CU(cudaMemcpyAsync(hdst + (uint64_t)c.b0 * dst_stride, L.out.p, (uint64_t)cnt * dstride, cudaMemcpyDeviceToHost, L.st));
        } else {
            CU(cudaMemcpy2DAsync(hdst + (uint64_t)c.b0 * dst_stride, dst_stride, L.out.p, dstride, dst_cap, cnt,
                                 cudaMemcpyDeviceToHost, L.st));
        }
        CU(cudaEventRecord(L.fin, L.st));
        return 0;
    };
    auto finish = [&](int k) -> int {
        Lane& L = pp->lane[k % kLanes];
        const Chunk& c = chunks[k];
        if (tiny) CU(cudaStreamSynchronize(L.st)); else CU(cudaEventSynchronize(L.fin));
        memcpy(out_len + c.b0, L.h_res.p, (size_t)(c.b1 - c.b0) * 4);
        if (stage_out) {
            const uint32_t cnt = c.b1 - c.b0;
            if (dstride == dst_stride) plz4cu_internal_copy(hdst + (uint64_t)c.b0 * dst_stride, L.h_out.p, (size_t)cnt * dstride);
            else for (uint32_t i = 0; i < cnt; i++) memcpy(hdst + (uint64_t)(c.b0 + i) * dst_stride, L.h_out.as<uint8_t>() + (uint64_t)i * dstride, dst_cap);
        }
        return 0;
    };
    int rc = 0;
    for (int k = 0; k < nchunks && !rc; k++) {
        if (k >= kLanes) rc = finish(k - kLanes);
        if (!rc) rc = issue(k);
    }
    for (int k = std::max(0, nchunks - kLanes); k < nchunks; k++) { int r2 = finish(k); if (!rc) rc = r2; }
    return rc;
}

// ---------------------------------------------------------------- per-block shims

static int one_block_compress(const plz4cu_dict_t* dict, const void* src, int n, void* dst, int cap)
{
    if (n < 0 || (uint32_t)n > 0x7E000000u) return 0;
    if (cap <= 0) return 0;
    uint64_t off = 0, poff[2] = {0, 0};
    uint32_t len = (uint32_t)n;
    // LZ4_compress_fast never writes more than bound(n) bytes, so a larger cap changes nothing
    uint32_t eff_cap = (uint32_t)std::min<uint64_t>((uint64_t)cap, plz4cu_compress_bound((size_t)n));
    std::vector<uint8_t> tmp((size_t)eff_cap + 32);
    uint8_t dummy = 0;
    int r = plz4cu_compress_batch_host(n ? src : &dummy, &off, &len, 1, eff_cap, 0, 1, dict, tmp.data(), tmp.size(), poff);
    if (r < 0) return INT32_MIN;
    uint64_t c = poff[1] - poff[0];
    if (c == 0 || c > (uint64_t)cap) return 0;
    memcpy(dst, tmp.data(), c);
    return (int)c;
}

int plz4cu_compress_fast(const void* src, int n, void* dst, int cap) { return one_block_compress(nullptr, src, n, dst, cap); }
int plz4cu_compress_fast_dict(const plz4cu_dict_t* dict, const void* src, int n, void* dst, int cap)
{
    return one_block_compress(dict, src, n, dst, cap);
}

static int one_block_decompress(const plz4cu_dict_t* dict, const void* src, int n, void* dst, int cap)
{
    if (src == nullptr || cap < 0) return -1;     // lz4.c:2033
    if (n < 0) return -1;
    if (n == 0) {
        // lz4.c:2064-2069: both special cases end in -1 for an empty input
        return -1;
    }
    uint64_t off = 0;
    uint32_t len = (uint32_t)n;
    int32_t res = 0;
    uint8_t dummy[16];
    int r = plz4cu_decompress_batch_host(src, (uint64_t)n, &off, &len, 1, (uint32_t)cap, 0, 1, dict,
                                         cap ? dst : dummy, (uint64_t)(cap ? cap : 16), &res);
    if (r < 0) return INT32_MIN;
    return res;
}
int plz4cu_decompress_safe(const void* src, int n, void* dst, int cap) { return one_block_decompress(nullptr, src, n, dst, cap); }
int plz4cu_decompress_safe_dict(const plz4cu_dict_t* dict, const void* src, int n, void* dst, int cap)
{
    return one_block_decompress(dict, src, n, dst, cap);
}

}  // extern "C"
