// frame_index.cu — block boundaries of a DEVICE-resident LZ4 frame body, found without a serial walk.
//
// The reference walks a frame one size word at a time (blk/frame.go:54-112): each word locates the next block, so
// on the host it is a serial pointer chase through the stream.  Done that way on a GPU it costs one dependent HBM
// load per block (~1 us): 131072 blocks = more than the whole decode.  Here the walk is turned inside out:
//   1. every byte position is tested as a size word ("candidate": word == 0, or size <= block size and the record
//      fits the body).  In compressed data about bsz / 2^31 of all positions pass by accident;
//   2. candidates are compacted in position order (flags + count per tile, scan, fill from the flags);
//   3. every candidate computes where ITS successor would start and looks that position up among the candidates
//      (binary search): the true boundaries form one chain from position 0 to the EndMark, accidental candidates
//      form stray chains that dead-end or merge into the true one;
//   4. pointer doubling builds jump tables nxt^(2^k);
//   5. one thread counts the hops from candidate 0 to the EndMark with the tables (binary lifting), then thread r
//      walks r hops the same way and writes the offset of block r.
// Work O(L + C log C) for L body bytes and C candidates, depth O(log C); the one streaming pass over the body
// dominates.  Frames with few blocks (large block sizes) use the plain walk: it is short, and accidental candidates
// would outnumber real ones by far.  The plain walk also names the error when the chain is broken.
#include "common.cuh"
#include "kernels.h"

#include <cstdlib>

namespace plz4 {

namespace {

constexpr int kThreads = 256;
constexpr int kPerThread = 32;
constexpr int kTile = kThreads * kPerThread;          // positions per CTA
constexpr uint32_t kTerm = 0xFFFFFFFEu;               // successor of an EndMark
constexpr uint32_t kDead = 0xFFFFFFFFu;               // successor position is not a candidate

struct Probe {
    const uint8_t* body;
    uint64_t len;
    uint32_t bsz;
    uint32_t trailer;                                 // 4 when records carry a block checksum
};

// 32 consecutive positions per thread: flags of the plausible size words.  The cheap test (size field <= block size,
// which includes the EndMark's zero word) runs on every position; the 64-bit "record and a following size word fit
// the body" test only on the few that pass.
__device__ __forceinline__ uint32_t probe32(const Probe& a, uint64_t p0)
{
    if (p0 + 4 > a.len) return 0;
    const uintptr_t addr = reinterpret_cast<uintptr_t>(a.body) + p0;
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(addr & ~uintptr_t(3));
    const uint32_t off = (uint32_t)(addr & 3);
    // aligned words covering bytes p0 .. p0+34 (never past the word holding the body's last byte)
    const uint64_t last_word = (reinterpret_cast<uintptr_t>(a.body) + a.len - 1 - (addr & ~uintptr_t(3))) >> 2;
    uint32_t w[10];
#pragma unroll
    for (int i = 0; i < 10; i++) w[i] = ((uint64_t)i <= last_word) ? wp[i] : 0u;
    uint32_t flags = 0;
#pragma unroll
    for (int j = 0; j < kPerThread; j++) {
        const uint32_t b = off + j;
        const uint32_t v = __funnelshift_r(w[b >> 2], w[(b >> 2) + 1], (b & 3) * 8);
        if ((v & 0x7FFFFFFFu) <= a.bsz) flags |= 1u << j;
    }
    const uint64_t room = a.len - 3 - p0;                               // positions from p0 on that hold a whole word
    if (room < 32) flags &= (1u << room) - 1u;
    for (uint32_t rest = flags; rest; rest &= rest - 1) {
        const int j = __ffs(rest) - 1;
        const uint32_t b = off + j;
        const uint32_t v = __funnelshift_r(w[b >> 2], w[(b >> 2) + 1], (b & 3) * 8);
        if (v != 0 && p0 + j + 4 + (v & 0x7FFFFFFFu) + a.trailer + 4 > a.len) flags &= ~(1u << j);
    }
    return flags;
}

// pass 1: flags of every position (one word per thread, kept for pass 2) and the number of candidates per tile
__global__ void __launch_bounds__(kThreads) count_candidates_kernel(Probe a, uint32_t* __restrict__ flag_words,
                                                                    uint32_t* __restrict__ tile_count)
{
    const uint64_t t = (uint64_t)blockIdx.x * kThreads + threadIdx.x;
    const uint32_t flags = probe32(a, t * kPerThread);
    flag_words[t] = flags;
    __shared__ int warp_sum[kThreads / 32];
    int s = __popc(flags);
#pragma unroll
    for (int d = 16; d; d >>= 1) s += __shfl_down_sync(FULL_MASK, s, d);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int sum = 0;
#pragma unroll
        for (int i = 0; i < kThreads / 32; i++) sum += warp_sum[i];
        tile_count[blockIdx.x] = (uint32_t)sum;
    }
}

// pass 2: candidate positions in increasing order
__global__ void __launch_bounds__(kThreads) fill_candidates_kernel(const uint32_t* __restrict__ flag_words,
                                                                   const uint64_t* __restrict__ tile_off,
                                                                   uint64_t* __restrict__ cand_pos)
{
    const uint64_t t = (uint64_t)blockIdx.x * kThreads + threadIdx.x;
    uint32_t flags = flag_words[t];
    const int n = __popc(flags);
    // exclusive rank of this thread's first candidate inside the tile
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int up = __shfl_up_sync(FULL_MASK, incl, d);
        if (lane >= d) incl += up;
    }
    __shared__ int warp_sum[kThreads / 32];
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    int before = incl - n;
    for (int i = 0; i < warp; i++) before += warp_sum[i];
    uint64_t at = tile_off[blockIdx.x] + (uint64_t)before;
    while (flags) {
        const int j = __ffs(flags) - 1;
        flags &= flags - 1;
        cand_pos[at++] = t * kPerThread + j;
    }
}

// successor of every candidate, as an index into cand_pos
__global__ void __launch_bounds__(256) link_kernel(Probe a, const uint64_t* __restrict__ cand_pos, uint32_t ncand,
                                                   uint32_t* __restrict__ nxt)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncand) return;
    const uint64_t p = cand_pos[i];
    const uint32_t w = load_le32(a.body + p);
    if (w == 0) { nxt[i] = kTerm; return; }
    const uint64_t target = p + 4 + (w & 0x7FFFFFFFu) + a.trailer;
    uint32_t lo = i + 1, hi = ncand;                                   // first candidate at or after target
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (cand_pos[mid] < target) lo = mid + 1; else hi = mid;
    }
    nxt[i] = (lo < ncand && cand_pos[lo] == target) ? lo : kDead;
}

__global__ void __launch_bounds__(256) double_kernel(const uint32_t* __restrict__ cur, uint32_t* __restrict__ next_level, uint32_t ncand)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncand) return;
    const uint32_t t = cur[i];
    next_level[i] = (t >= kTerm) ? t : cur[t];
}

// result[0] = data blocks on the chain from candidate 0, result[1] = 1 if the chain ends in an EndMark,
// result[2..3] = body offset of the chain's last node
__global__ void count_hops_kernel(const uint64_t* __restrict__ cand_pos, const uint32_t* __restrict__ jump, uint32_t ncand,
                                  int levels, uint64_t* __restrict__ result)
{
    if (threadIdx.x || blockIdx.x) return;
    if (ncand == 0 || cand_pos[0] != 0) { result[0] = 0; result[1] = 0; result[2] = 0; return; }
    uint32_t node = 0;
    uint64_t hops = 0;
    for (int k = levels - 1; k >= 0; k--) {
        const uint32_t t = jump[(uint64_t)k * ncand + node];
        if (t < kTerm) { node = t; hops += 1ull << k; }
    }
    result[0] = hops;
    result[1] = (jump[node] == kTerm) ? 1 : 0;
    result[2] = cand_pos[node];
}

__global__ void __launch_bounds__(256) emit_offsets_kernel(const uint64_t* __restrict__ cand_pos, const uint32_t* __restrict__ jump,
                                                           uint32_t ncand, int levels, const uint64_t* __restrict__ result,
                                                           uint64_t* __restrict__ rec_off, uint32_t cap)
{
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= result[0] || r >= cap) return;
    uint32_t node = 0;
    for (int k = 0; k < levels; k++)
        if ((r >> k) & 1) node = jump[(uint64_t)k * ncand + node];
    rec_off[r] = cand_pos[node];
}

// The plain walk (blk/frame.go:54-112): result[0] = blocks found, result[1] = 1 EndMark reached, 2 size overflow,
// 3 record runs past the body, 4 body ends where a size word should be; result[2] = offset where the walk stopped.
__global__ void walk_kernel(Probe a, uint64_t* __restrict__ rec_off, uint32_t cap, uint64_t* __restrict__ result)
{
    if (threadIdx.x || blockIdx.x) return;
    uint64_t p = 0, n = 0;
    uint32_t why;
    for (;;) {
        if (p + 4 > a.len) { why = 4; break; }
        const uint32_t w = load_le32(a.body + p);
        if (w == 0) { why = 1; break; }
        const uint32_t size = w & 0x7FFFFFFFu;
        if (size > a.bsz) { why = 2; break; }
        if (p + 4 + size + a.trailer > a.len) { why = 3; break; }
        if (n < cap) rec_off[n] = p;
        n++;
        p += 4 + (uint64_t)size + a.trailer;
    }
    result[0] = n; result[1] = why; result[2] = p;
}

}  // namespace

cudaError_t launch_frame_index(const uint8_t* body, uint64_t len, uint32_t bsz, int blk_check, uint64_t* rec_off, uint32_t cap,
                               FrameIndexResult* out, uint64_t* launches, cudaStream_t stream)
{
    Probe a{body, len, bsz, blk_check ? 4u : 0u};
    uint64_t* d_result = nullptr;
    uint64_t h_result[3] = {0, 0, 0};
    cudaError_t e = cudaMallocAsync((void**)&d_result, 3 * sizeof(uint64_t), stream);
    if (e != cudaSuccess) return e;
    auto walk = [&]() -> cudaError_t {
        walk_kernel<<<1, 32, 0, stream>>>(a, rec_off, cap, d_result);
        ++*launches;
        cudaError_t e2 = cudaMemcpyAsync(h_result, d_result, sizeof h_result, cudaMemcpyDeviceToHost, stream);
        if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(stream);
        out->nblk = h_result[0]; out->why = (uint32_t)h_result[1]; out->stop_off = h_result[2];
        return e2;
    };
    // few blocks: the serial walk is short, and accidental candidates would swamp the real ones
    // measurement / test knobs: force the one-thread walk, or let small frames take the parallel path
    static const bool force_walk = getenv("PLZ4CU_SERIAL_WALK") != nullptr;
    static const uint64_t min_blocks = getenv("PLZ4CU_WALK_MIN_BLOCKS") ? strtoull(getenv("PLZ4CU_WALK_MIN_BLOCKS"), nullptr, 10) : 1024;
    if (len / bsz <= min_blocks || force_walk) {
        e = walk();
        cudaFreeAsync(d_result, stream);
        return e;
    }

    const uint64_t ntiles = (len + kTile - 1) / kTile;
    uint32_t* tile_count = nullptr;
    uint32_t* flag_words = nullptr;
    uint64_t* tile_off = nullptr;
    uint64_t* cand_pos = nullptr;
    uint32_t* jump = nullptr;
    auto cleanup = [&] {
        if (tile_count) cudaFreeAsync(tile_count, stream);
        if (flag_words) cudaFreeAsync(flag_words, stream);
        if (tile_off) cudaFreeAsync(tile_off, stream);
        if (cand_pos) cudaFreeAsync(cand_pos, stream);
        if (jump) cudaFreeAsync(jump, stream);
        cudaFreeAsync(d_result, stream);
    };
    e = cudaMallocAsync((void**)&tile_count, ntiles * sizeof(uint32_t), stream);
    if (e == cudaSuccess) e = cudaMallocAsync((void**)&tile_off, (ntiles + 1) * sizeof(uint64_t), stream);
    if (e == cudaSuccess) e = cudaMallocAsync((void**)&flag_words, ntiles * kThreads * sizeof(uint32_t), stream);
    if (e != cudaSuccess) { cleanup(); return e; }
    count_candidates_kernel<<<(unsigned)ntiles, kThreads, 0, stream>>>(a, flag_words, tile_count);
    e = launch_scan_u32(tile_count, (uint32_t)ntiles, tile_off, stream);
    *launches += 2;
    uint64_t ncand64 = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&ncand64, tile_off + ntiles, sizeof ncand64, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) { cleanup(); return e; }
    if (ncand64 == 0 || ncand64 >= kTerm) {            // nothing plausible at all (or absurdly many): let the walk explain
        e = walk();
        cleanup();
        return e;
    }
    const uint32_t ncand = (uint32_t)ncand64;
    int levels = 1;
    while ((1ull << levels) < (uint64_t)ncand) levels++;
    e = cudaMallocAsync((void**)&cand_pos, (uint64_t)ncand * sizeof(uint64_t), stream);
    if (e == cudaSuccess) e = cudaMallocAsync((void**)&jump, (uint64_t)levels * ncand * sizeof(uint32_t), stream);
    if (e != cudaSuccess) { cleanup(); return e; }
    fill_candidates_kernel<<<(unsigned)ntiles, kThreads, 0, stream>>>(flag_words, tile_off, cand_pos);
    const unsigned g = (ncand + 255) / 256;
    link_kernel<<<g, 256, 0, stream>>>(a, cand_pos, ncand, jump);
    for (int k = 1; k < levels; k++)
        double_kernel<<<g, 256, 0, stream>>>(jump + (uint64_t)(k - 1) * ncand, jump + (uint64_t)k * ncand, ncand);
    count_hops_kernel<<<1, 32, 0, stream>>>(cand_pos, jump, ncand, levels, d_result);
    emit_offsets_kernel<<<g, 256, 0, stream>>>(cand_pos, jump, ncand, levels, d_result, rec_off, cap);
    *launches += 3 + levels;
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_result, d_result, sizeof h_result, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e == cudaSuccess) {
        if (h_result[1] == 1) { out->nblk = h_result[0]; out->why = 1; out->stop_off = h_result[2]; }
        else e = walk();                               // broken chain: the plain walk says where and why
    }
    cleanup();
    return e;
}

}  // namespace plz4
