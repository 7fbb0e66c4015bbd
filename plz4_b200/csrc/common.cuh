// common.cuh — device helpers shared by the sm_100a kernels (warp-level byte plumbing + xxh32).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define FULL_MASK 0xffffffffu

namespace plz4 {

constexpr int MINMATCH = 4;
constexpr int LASTLITERALS = 5;
constexpr int MFLIMIT = 12;
constexpr uint32_t MAX_DISTANCE = 65535;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// ---------------------------------------------------------------- unaligned access

// 4 bytes at an arbitrary byte address, read as two aligned words + funnel shift.  The aligned
// words may start up to 3 bytes before p / end up to 3 bytes after p+4; callers guarantee those
// bytes lie inside the same allocation (true for any interior pointer of a CUDA allocation).
__device__ __forceinline__ uint32_t load_u32_unaligned(const uint8_t* p)
{
    uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
    uint32_t sh = (uint32_t)(a & 3) * 8;
    uint32_t lo = w[0];
    if (sh == 0) return lo;
    return __funnelshift_r(lo, w[1], sh);
}

__device__ __forceinline__ uint32_t load_u16le(const uint8_t* p)
{
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8);
}

// ---------------------------------------------------------------- warp copies

// Forward, non-overlapping copy of n bytes by one warp.  Handles any alignment; moves 16 bytes per
// lane per step once dst is 16-byte aligned and src happens to share the alignment, 4 bytes per lane
// when only word alignment can be shared (src re-aligned with a funnel shift), bytes otherwise.
__device__ __forceinline__ void warp_copy(uint8_t* dst, const uint8_t* src, uint32_t n, int lane)
{
    if (n < 64) {
        for (uint32_t k = lane; k < n; k += 32) dst[k] = src[k];
        return;
    }
    // head: bring dst to 16-byte alignment
    uint32_t head = (uint32_t)((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15);
    if ((uint32_t)lane < head) dst[lane] = src[lane];
    dst += head; src += head; n -= head;
    uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15);
    uint32_t nvec = n >> 4;
    if (mis == 0) {
        const uint4* s = reinterpret_cast<const uint4*>(src);
        uint4* d = reinterpret_cast<uint4*>(dst);
        uint32_t k = lane;
        for (; k + 96 < nvec; k += 128) {      // 4 loads in flight per lane
            uint4 a = s[k], b = s[k + 32], c = s[k + 64], e = s[k + 96];
            d[k] = a; d[k + 32] = b; d[k + 64] = c; d[k + 96] = e;
        }
        for (; k < nvec; k += 32) d[k] = s[k];
    } else if ((mis & 3) == 0) {
        const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
        uint4* d = reinterpret_cast<uint4*>(dst);
        for (uint32_t k = lane; k < nvec; k += 32) {
            uint4 v;
            v.x = s[4 * k]; v.y = s[4 * k + 1]; v.z = s[4 * k + 2]; v.w = s[4 * k + 3];
            d[k] = v;
        }
    } else {
        // src is byte-misaligned relative to dst: aligned word loads + funnel shift, 16-byte stores
        const uint32_t* s = reinterpret_cast<const uint32_t*>(src - (mis & 3));
        uint32_t sh = (mis & 3) * 8;
        uint4* d = reinterpret_cast<uint4*>(dst);
        for (uint32_t k = lane; k < nvec; k += 32) {
            uint32_t w0 = s[4 * k], w1 = s[4 * k + 1], w2 = s[4 * k + 2], w3 = s[4 * k + 3], w4 = s[4 * k + 4];
            uint4 v;
            v.x = __funnelshift_r(w0, w1, sh); v.y = __funnelshift_r(w1, w2, sh);
            v.z = __funnelshift_r(w2, w3, sh); v.w = __funnelshift_r(w3, w4, sh);
            d[k] = v;
        }
    }
    uint32_t done = nvec << 4;
    for (uint32_t k = done + lane; k < n; k += 32) dst[k] = src[k];
}

// ---------------------------------------------------------------- xxh32 (seed 0)

constexpr uint32_t XP1 = 2654435761u, XP2 = 2246822519u, XP3 = 3266489917u, XP4 = 668265263u,
                   XP5 = 374761393u;

__device__ __forceinline__ uint32_t rol32(uint32_t x, int r) { return __funnelshift_l(x, x, r); }

// The four accumulators live in lanes j = lane & 3 (every group of four lanes carries a copy).
__device__ __forceinline__ uint32_t xxh32_init(int lane)
{
    const int j = lane & 3;
    return (j == 0) ? (XP1 + XP2) : (j == 1) ? XP2 : (j == 2) ? 0u : (0u - XP1);
}

// Merge of the accumulators, the 0..15 trailing bytes of p[0..n) and the avalanche (xxh32zero.go:259-277).
__device__ __forceinline__ uint32_t xxh32_finish(uint32_t acc, const uint8_t* p, uint32_t n)
{
    uint32_t h;
    if (n >= 16) {
        uint32_t v0 = __shfl_sync(FULL_MASK, acc, 0), v1 = __shfl_sync(FULL_MASK, acc, 1);
        uint32_t v2 = __shfl_sync(FULL_MASK, acc, 2), v3 = __shfl_sync(FULL_MASK, acc, 3);
        h = rol32(v0, 1) + rol32(v1, 7) + rol32(v2, 12) + rol32(v3, 18);
    } else {
        h = XP5;
    }
    h += n;
    uint32_t i = (n >> 4) << 4;
    for (; i + 4 <= n; i += 4) {
        uint32_t v = (uint32_t)p[i] | ((uint32_t)p[i + 1] << 8) | ((uint32_t)p[i + 2] << 16) | ((uint32_t)p[i + 3] << 24);
        h = rol32(h + v * XP3, 17) * XP4;
    }
    for (; i < n; i++) h = rol32(h + (uint32_t)p[i] * XP5, 11) * XP1;
    h ^= h >> 15; h *= XP2; h ^= h >> 13; h *= XP3; h ^= h >> 16;
    return h;
}

// `nstripes` whole stripes from word-aligned memory that answers quickly (a shared-memory tile): the chain only.
__device__ __forceinline__ uint32_t xxh32_consume_words(uint32_t acc, const uint32_t* w, uint32_t nstripes, int lane)
{
    const uint32_t nwords = nstripes << 2;
    for (uint32_t base = 0; base < nstripes; base += 32) {
        uint32_t y[4];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const uint32_t idx = base * 4 + r * 32 + lane;
            y[r] = (idx < nwords ? w[idx] : 0u) * XP2;
        }
        const uint32_t left = nstripes - base;
        if (left >= 32) {
#pragma unroll
            for (int r = 0; r < 4; r++) {
#pragma unroll
                for (int t = 0; t < 8; t++) acc = rol32(acc + __shfl_sync(FULL_MASK, y[r], 4 * t + (lane & 3)), 13) * XP1;
            }
        } else {
#pragma unroll
            for (int r = 0; r < 4; r++) {
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    const uint32_t x = __shfl_sync(FULL_MASK, y[r], 4 * t + (lane & 3));
                    if ((uint32_t)(r * 8 + t) < left) acc = rol32(acc + x, 13) * XP1;
                }
            }
        }
    }
    return acc;
}

// xxh32.ChecksumZero (xxh32/xxh32zero.go:238-280) of p[0..n) computed by one warp; every lane
// returns the digest.  The four accumulator chains are inherently serial along the stripes, so
// lanes 0..3 carry them; the other lanes only help to fetch: a chunk of 32 stripes (512 B) is
// loaded coalesced (4 words per lane) and handed to the chain lanes with shuffles.
static __device__ __noinline__ uint32_t warp_xxh32(const uint8_t* p, uint32_t n, int lane)
{
    uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
    const uint32_t sh = (uint32_t)(a & 3) * 8;
    const uint32_t nstripes = n >> 4;
    const uint32_t nwords = nstripes << 2;
    uint32_t acc = xxh32_init(lane);
    // one chunk = 32 stripes = 4 words per lane, already multiplied by XP2 (that product is off the serial chain)
    auto load_chunk = [&](uint32_t base, uint32_t (&y)[4]) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const uint32_t idx = base * 4 + r * 32 + lane;
            uint32_t v = 0;
            if (idx < nwords) {
                v = wp[idx];
                if (sh) v = __funnelshift_r(v, wp[idx + 1], sh);
            }
            y[r] = v * XP2;
        }
    };
    uint32_t cur[4] = {0, 0, 0, 0};
    if (nstripes) load_chunk(0, cur);
    for (uint32_t base = 0; base < nstripes; base += 32) {
        // the next chunk's loads are in flight while this one goes down the chain, and the lines eight chunks
        // further on are asked into L1: one chunk of chain work is far shorter than a trip to HBM
        uint32_t nxt[4] = {0, 0, 0, 0};
        if (base + 32 < nstripes) load_chunk(base + 32, nxt);
        {
            const uint32_t ahead = (base + 32 * 8) * 4 + (uint32_t)lane * 4;
            if (ahead < nwords) asm volatile("prefetch.global.L1 [%0];" ::"l"(wp + ahead));
        }
        const uint32_t left = nstripes - base;      // stripes in this chunk (warp-uniform)
        if (left >= 32) {
            // full chunk: nothing but add, rotate, multiply on the chain; the shuffles do not depend on it
#pragma unroll
            for (int r = 0; r < 4; r++) {
#pragma unroll
                for (int t = 0; t < 8; t++) acc = rol32(acc + __shfl_sync(FULL_MASK, cur[r], 4 * t + (lane & 3)), 13) * XP1;
            }
        } else {
#pragma unroll
            for (int r = 0; r < 4; r++) {
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    const uint32_t x = __shfl_sync(FULL_MASK, cur[r], 4 * t + (lane & 3));
                    if ((uint32_t)(r * 8 + t) < left) acc = rol32(acc + x, 13) * XP1;
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 4; r++) cur[r] = nxt[r];
    }
    return xxh32_finish(acc, p, n);
}

// ---------------------------------------------------------------- xxh32 in pieces (a payload hashed while it is produced / consumed)

// `nchunks` chunks of 512 bytes (32 stripes) starting at word pointer wp (+ byte shift sh), read around L1
// kBypassL1: the bytes were just written by other warps of this CTA (read around L1); otherwise plain loads, which also
// leave the lines in L1 for a consumer that parses the same bytes next
template <bool kBypassL1 = true>
__device__ __forceinline__ uint32_t xxh32_consume_global(uint32_t acc, const uint32_t* wp, uint32_t sh, int nchunks, int lane)
{
    auto load = [&](int c, uint32_t (&y)[4]) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int idx = c * 128 + r * 32 + lane;
            uint32_t v = kBypassL1 ? __ldcg(wp + idx) : wp[idx];
            if (sh) v = __funnelshift_r(v, kBypassL1 ? __ldcg(wp + idx + 1) : wp[idx + 1], sh);
            y[r] = v * XP2;
        }
    };
    uint32_t cur[4] = {0, 0, 0, 0};
    if (nchunks > 0) load(0, cur);
    for (int c = 0; c < nchunks; c++) {
        uint32_t nxt[4] = {0, 0, 0, 0};
        if (c + 1 < nchunks) load(c + 1, nxt);              // in flight while this chunk goes down the chain
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int t = 0; t < 8; t++) acc = rol32(acc + __shfl_sync(FULL_MASK, cur[r], 4 * t + (lane & 3)), 13) * XP1;
        }
#pragma unroll
        for (int r = 0; r < 4; r++) cur[r] = nxt[r];
    }
    return acc;
}

// the rest of a payload of n bytes of which `done` (a multiple of 16) are already in acc: whole stripes, then the tail
__device__ __forceinline__ uint32_t xxh32_finish_global(uint32_t acc, const uint8_t* p, uint32_t done, uint32_t n, int lane)
{
    const uint32_t stripes = (n >> 4) - (done >> 4);
    for (uint32_t s0 = 0; s0 < stripes; s0 += 8) {
        // 8 stripes per round: lane l holds word l of the round
        const uint32_t idx = done + s0 * 16 + (uint32_t)lane * 4;
        uint32_t v = 0;
        if (idx + 4 <= (n & ~15u)) {
            const uint8_t* q = p + idx;
            v = (uint32_t)__ldcg(q) | ((uint32_t)__ldcg(q + 1) << 8) | ((uint32_t)__ldcg(q + 2) << 16) | ((uint32_t)__ldcg(q + 3) << 24);
        }
        v *= XP2;
        const uint32_t left = min(8u, stripes - s0);
#pragma unroll
        for (int t = 0; t < 8; t++) {
            const uint32_t x = __shfl_sync(FULL_MASK, v, 4 * t + (lane & 3));
            if ((uint32_t)t < left) acc = rol32(acc + x, 13) * XP1;
        }
    }
    uint32_t h;
    if (n >= 16) {
        const uint32_t v0 = __shfl_sync(FULL_MASK, acc, 0), v1 = __shfl_sync(FULL_MASK, acc, 1);
        const uint32_t v2 = __shfl_sync(FULL_MASK, acc, 2), v3 = __shfl_sync(FULL_MASK, acc, 3);
        h = rol32(v0, 1) + rol32(v1, 7) + rol32(v2, 12) + rol32(v3, 18);
    } else {
        h = XP5;
    }
    h += n;
    uint32_t i = n & ~15u;
    for (; i + 4 <= n; i += 4) {
        const uint32_t v = (uint32_t)__ldcg(p + i) | ((uint32_t)__ldcg(p + i + 1) << 8) | ((uint32_t)__ldcg(p + i + 2) << 16) |
                           ((uint32_t)__ldcg(p + i + 3) << 24);
        h = rol32(h + v * XP3, 17) * XP4;
    }
    for (; i < n; i++) h = rol32(h + (uint32_t)__ldcg(p + i) * XP5, 11) * XP1;
    h ^= h >> 15; h *= XP2; h ^= h >> 13; h *= XP3; h ^= h >> 16;
    return h;
}

__device__ __forceinline__ void store_le32(uint8_t* p, uint32_t v)
{
    p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
}
__device__ __forceinline__ uint32_t load_le32(const uint8_t* p)
{
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

}  // namespace plz4
