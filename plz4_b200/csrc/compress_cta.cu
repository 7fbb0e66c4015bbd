// compress_cta.cu — LZ4 level-1 block encode, one CTA per block of up to 64 KiB (sm_100a).
//
// Replaces, for a whole batch of independent blocks, what plz4 does per block on a goroutine:
//   async/writer.go:232-282 compressLoop -> blk/blk.go:69-109 CompressToBlk
//     -> compress/indie.go:66-74 -> clz4.go:31-45 -> lz4.c:930-1338 LZ4_compress_generic_validated
//   plus the record framing (size word / stored fallback / xxh32 trailer) of blk/blk.go:87-106.
//
// The block is brought into shared memory ONCE by a TMA bulk copy (cp.async.bulk + mbarrier) and every later
// access — hashing, candidate verification, match extension, literal copies — is a shared-memory access.
// The CTA's warps are workers on tiles of 1024 positions (9 workers, tiles round-robin) plus one hasher; what
// must happen in order passes from tile to tile through mbarriers (parked waits), everything else runs in parallel:
//   (1) index: for EVERY position the most recent earlier position with the same hash (liblz4's hash4,
//       lz4.c:777-783, 4096 slots).  The 4-byte values of the tile are read ahead into registers; when the token
//       arrives the warp issues one shared-memory atomicMax per 32 positions (positions only grow, so max is
//       "most recent"; the value returned is the table entry from before, or a lower lane of the same group) and
//       hands the token on.
//   (2) verify: every candidate is compared with its position (4 equal bytes) -> one 32-bit map per lane.
//   (3) parse: every LANE parses its own 32 positions serially — greedy with a one-step lazy check, forward
//       extension in 4-byte steps, backward extension, at most 8 matches — with no knowledge of its neighbours;
//       matches may run past the lane's end; matches of 64 bytes and more are completed by the whole warp.
//   (4) resolve, in tile order: a prefix maximum over the lanes' match ends tells each lane where the parse of the
//       lanes before it stops, what it must drop or trim, and where its first literal run begins; sizes are
//       prefix-summed.  Entry state (covered-up-to, literal anchor, output offset) passes from tile to tile.
//   (5) emit: every lane writes its own sequences into a staging tile, which goes out with 16-byte stores.
//   * hasher (1 warp): block checksum (xxh32.ChecksumZero, xxh32/xxh32zero.go:238-280) over the payload, chunk
//     by chunk as tiles complete; then last literals (lz4.c:1302-1329), stored fallback, size word, trailer.
// The parse is not liblz4's (bytes differ, the reference decodes them; size within the tolerance pinned by
// tests/test_gpu_compress.py): all positions enter the table, and the lazy step more than pays for the lanes'
// independent starts (tools/parse_model.c is the CPU model the design was sized with).
#include "common.cuh"
#include "kernels.h"

#include <cstdlib>

namespace plz4 {

// Stage timing for tuning (nvcc -DPLZ4CU_CTA_PROF): cycles summed over all tiles of all blocks by lane 0 of every worker.
#ifdef PLZ4CU_CTA_PROF
__device__ unsigned long long g_cta_prof[16];
#define PROF_DECL long long prof_t = clock64()
#define PROF(i) do { const long long now__ = clock64(); if (lane == 0) atomicAdd(&g_cta_prof[i], (unsigned long long)(now__ - prof_t)); prof_t = now__; } while (0)
#else
#define PROF_DECL
#define PROF(i)
#endif

namespace {

constexpr int kCtaThreads = 384;
constexpr int kWorkers = 11;                    // warps 0..8; warp 9 hashes
constexpr int kTile = 1024;                     // positions per tile: 32 lanes x 32 positions
constexpr int kMaxTiles = 64;
constexpr int kPvStride = 34;                   // u16 per group in prev[]: 17 words, so lanes reading their own group hit 32 banks
constexpr int kStage = 2 * 32 * kPvStride;      // staging bytes per worker (prev[] of a tile lives in the same bytes)
constexpr int kLaneCap = 64;                    // a lane extends a match this far by itself
constexpr int kLazyBelow = 16;                  // the lazy check is made for matches shorter than this
constexpr int kLongLit = 48;                    // literal runs from this length on are copied by the whole warp
constexpr int kWinPad = 32;
#ifndef PLZ4CU_WAIT_NAP
#define PLZ4CU_WAIT_NAP 64
#endif
constexpr unsigned kWaitNap = PLZ4CU_WAIT_NAP;     // ns between two polls of an mbarrier
constexpr int kHashChunk = 512;                 // bytes the hasher consumes per step (32 stripes)

constexpr int kSpanWorkers = 22;                // large blocks: one CTA per SM, every warp a worker
constexpr int kSpanThreads = kSpanWorkers * 32;

template <int kWin, int kNW>
struct __align__(16) CtaSmemT {
    uint8_t win[kWin + kWinPad];                // the block (large blocks: the fragment before the current one, then the current one)
    uint8_t stage[kNW][kStage];                 // per worker: prev[] of its tile (u16 x 1024) while parsing, staging tile while emitting
    uint16_t table[4096];                       // hash -> a recent position (its low 16 bits), 0xFFFF = none
    uint32_t recs[kNW][8][32];                  // [record][lane]: offset<<16 | min(len,2047)<<5 | start-b0
    unsigned long long bar_load;
    unsigned long long bar_token[kMaxTiles + 1];      // [t]: the table holds every position before tile t
    unsigned long long bar_entry[kMaxTiles + 1];      // [t]: entry state (covered-up-to, anchor) of tile t is published
    unsigned long long bar_out[kMaxTiles + 1];        // [t]: output offset of tile t is published
    unsigned long long bar_done[kMaxTiles];           // [t]: tile t's bytes are in global memory
    int st_x[kMaxTiles + 1], st_anchor[kMaxTiles + 1], st_out[kMaxTiles + 1];
    int fail;
    uint32_t hash_acc[32];                            // the hasher's running state, parked for the finish
    int hash_done;
};
using CtaSmem = CtaSmemT<65536, kWorkers>;
using SpanSmem = CtaSmemT<131072, kSpanWorkers>;

// What a worker needs to know about the stretch of input it works on.  Positions are absolute inside the block; w32 / win
// are biased so that w32[p >> 2] / win[p] is position p whichever part of the block the shared-memory window holds.
struct TileEnv {
    const uint32_t* w32;
    const uint8_t* win;
    const uint8_t* gsrc;   // the block in global memory (large blocks: literal runs that reach behind the window)
    uint8_t* payload;      // where the sequences go
    int pos0;              // position of tile 0
    int ntiles;
    int hash_end;          // positions below it are hashed
    int mf_end;            // a match may start at p < mf_end        (lz4.c:963: last match starts <= n-12)
    int match_end;         // and must end at or before match_end    (last 5 bytes are literals)
    int lo_valid;          // lowest position a candidate may have
    int cap;
    uint32_t parity;       // phase of the per-tile barriers (large blocks reuse them fragment after fragment)
    bool index_only;       // large blocks: the fragment before a span only warms the table
};

// ---------------------------------------------------------------- mbarrier / TMA plumbing

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
// Parked wait for the phase of the given parity; a stall of seconds is a bug and ends the kernel instead of the box.
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity)
{
    const uint32_t a = smem_addr(bar);
    const long long t0 = clock64();
    for (uint32_t spins = 1;; spins++) {
        uint32_t ok;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(a), "r"(parity), "r"(20000u)
            : "memory");
        if (ok) return;
        __nanosleep(kWaitNap);                                // a waiting warp leaves its issue slots to the workers
        if ((spins & 255u) == 0 && clock64() - t0 > 4000000000ll) __trap();
    }
}
__device__ __forceinline__ void tma_load_bulk(void* dst_smem, const void* src_gmem, uint32_t bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

// ---------------------------------------------------------------- small helpers

// 4 bytes at byte position p of the window (two aligned words + funnel shift; reads up to p+7)
__device__ __forceinline__ uint32_t ld4(const uint32_t* __restrict__ w32, int p)
{
    const int i = p >> 2;
    return __funnelshift_r(w32[i], w32[i + 1], (uint32_t)(p & 3) * 8u);
}

__device__ __forceinline__ int ext_len(int rest) { return rest / 255 + 1; }     // bytes of a 255-run for a saturated nibble

__device__ __forceinline__ int seq_size(int lit, int len)
{
    int s = 3 + lit;
    if (lit >= 15) s += ext_len(lit - 15);
    if (len >= 19) s += ext_len(len - 19);
    return s;
}

__device__ __forceinline__ int warp_incl_max(int v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int u = __shfl_up_sync(FULL_MASK, v, d);
        if (lane >= d) v = max(v, u);
    }
    return v;
}
__device__ __forceinline__ int warp_incl_sum(int v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int u = __shfl_up_sync(FULL_MASK, v, d);
        if (lane >= d) v += u;
    }
    return v;
}

// equal bytes of win[a..] and win[b..], at most `limit`, by the whole warp: 128 bytes per round
__device__ __noinline__ int warp_count_equal(const uint32_t* __restrict__ w32, int a, int b, int limit, int lane)
{
    for (int total = 0; total < limit; total += 128) {
        const int k = total + 4 * lane;
        uint32_t x = 0;
        if (k < limit) {
            x = ld4(w32, a + k) ^ ld4(w32, b + k);
            const int rem = limit - k;
            if (rem < 4) x &= (1u << (8 * rem)) - 1u;
        }
        const uint32_t bal = __ballot_sync(FULL_MASK, x != 0);
        if (bal) {
            const int l = __ffs(bal) - 1;
            const uint32_t xl = __shfl_sync(FULL_MASK, x, l);
            return total + 4 * l + ((__ffs(xl) - 1) >> 3);
        }
    }
    return limit;
}

// n bytes from shared memory (any alignment) to global memory (any alignment) by `nthr` threads: 16-byte stores on the
// destination's alignment, bytes at the two ends
__device__ __forceinline__ void copy_s2g(uint8_t* dst, const uint8_t* src, int n, int tid, int nthr)
{
    if (n <= 0) return;
    const int head = min(n, (int)((16u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u));
    for (int k = tid; k < head; k += nthr) dst[k] = src[k];
    const int nvec = (n - head) >> 4;
    const uint8_t* s = src + head;
    uint4* d = reinterpret_cast<uint4*>(dst + head);
    const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(s) & 3u) * 8u;
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(s) & ~uintptr_t(3));
    for (int k = tid; k < nvec; k += nthr) {
        const uint32_t w0 = sw[4 * k], w1 = sw[4 * k + 1], w2 = sw[4 * k + 2], w3 = sw[4 * k + 3];
        uint4 v;
        if (sh == 0) {
            v = make_uint4(w0, w1, w2, w3);
        } else {
            const uint32_t w4 = sw[4 * k + 4];
            v.x = __funnelshift_r(w0, w1, sh); v.y = __funnelshift_r(w1, w2, sh);
            v.z = __funnelshift_r(w2, w3, sh); v.w = __funnelshift_r(w3, w4, sh);
        }
        d[k] = v;
    }
    for (int k = head + (nvec << 4) + tid; k < n; k += nthr) dst[k] = src[k];
}

// ---------------------------------------------------------------- workers

// equal bytes of a position and its candidate, from their first byte on: 16 bytes per round — five words of each side
// in flight at once — at most kLaneCap + 16 bytes (then `un`: the whole warp completes it); fewer than 4 = no match
__device__ __forceinline__ void lane_extend(const uint32_t* __restrict__ w32, int p, int c, int lim, int& ml, bool& un)
{
    const uint32_t* __restrict__ wa = w32 + (p >> 2);
    const uint32_t* __restrict__ wb = w32 + (c >> 2);
    const uint32_t sa = (uint32_t)(p & 3) * 8u, sb = (uint32_t)(c & 3) * 8u;
    uint32_t a0 = wa[0], b0 = wb[0];
    int done = 0;
    un = false;
    for (;;) {
        const uint32_t a1 = wa[1], a2 = wa[2], a3 = wa[3], a4 = wa[4];
        const uint32_t b1 = wb[1], b2 = wb[2], b3 = wb[3], b4 = wb[4];
        const uint32_t x0 = __funnelshift_r(a0, a1, sa) ^ __funnelshift_r(b0, b1, sb);
        const uint32_t x1 = __funnelshift_r(a1, a2, sa) ^ __funnelshift_r(b1, b2, sb);
        const uint32_t x2 = __funnelshift_r(a2, a3, sa) ^ __funnelshift_r(b2, b3, sb);
        const uint32_t x3 = __funnelshift_r(a3, a4, sa) ^ __funnelshift_r(b3, b4, sb);
        if (x0 | x1 | x2 | x3) {
            const uint32_t xs = x0 ? x0 : (x1 ? x1 : (x2 ? x2 : x3));
            const int skip = x0 ? 0 : (x1 ? 4 : (x2 ? 8 : 12));
            done += skip + ((__ffs(xs) - 1) >> 3);
            break;
        }
        done += 16;
        if (done >= lim) break;
        if (done >= kLaneCap) { un = true; break; }
        a0 = a4; b0 = b4;
        wa += 4; wb += 4;
    }
    ml = min(done, lim);
}

__device__ __forceinline__ void put_ext_bytes(uint8_t* o, int rest)
{
    for (; rest >= 255; rest -= 255) *o++ = 255;
    *o = (uint8_t)rest;
}

// hash of the 4 (blocks up to 64 KiB: liblz4's hash4, lz4.c:777-783) or 5 (larger blocks: its hash5, lz4.c:785-795) bytes at a
// position, as the byte offset of the table slot; w0, w1 are the aligned words holding them, lsh the position's byte shift
template <bool kLarge>
__device__ __forceinline__ uint32_t slot_offset(uint32_t w0, uint32_t w1, uint32_t lsh)
{
    const uint32_t v = __funnelshift_r(w0, w1, lsh);
    if (!kLarge) return ((v * 2654435761u) >> 19) & 0x1FFEu;
    // bits [28, 40) of (five bytes) * 889523592379 mod 2^40
    const uint32_t b4 = (w1 >> lsh) & 0xFFu;
    const uint32_t lo = v * 0x1BBCDCBBu, hi = __umulhi(v, 0x1BBCDCBBu);
    const uint32_t top = (hi + v * 0xCFu + b4 * 0x1BBCDCBBu) & 0xFFu;
    return (((top << 4) | (lo >> 28)) << 1) & 0x1FFEu;
}

// candidate position from its 16 stored bits: the one at most 65535 below p (large blocks; small ones store it whole)
__device__ __forceinline__ int cand_abs(int p, uint32_t c16)
{
    int c = (p & ~0xFFFF) | (int)c16;
    if (c >= p) c -= 65536;
    return c;
}

template <bool kLarge, int kNW, typename SM>
__device__ __forceinline__ void run_worker(SM& S, const TileEnv& C, int pw, int lane, int& my_x)
{
    const uint32_t* __restrict__ w32 = C.w32;
    const uint8_t* __restrict__ win = C.win;
    uint8_t* payload = C.payload;
    const int cap = C.cap;
    const int hash_end = C.hash_end;
    const int ntiles = C.ntiles;
    const uint32_t par = C.parity;
    uint8_t* stage = S.stage[pw];
    uint16_t* pv = reinterpret_cast<uint16_t*>(S.stage[pw]);
    uint32_t(*recs)[32] = S.recs[pw];

    for (int t = pw; t < ntiles; t += kNW) {
        const int tile_base = C.pos0 + t * kTile;
        const int b0 = tile_base + 32 * lane;
        PROF_DECL;

        // ---- (1) index.  Position tile_base + 32 g + lane: the word index is the same for four lanes and the byte
        // shift is the lane's own, so the values come in with immediate offsets.  Every position is looked up, the
        // EVEN positions are entered (half the table traffic, and a table that churns half as fast gives longer
        // matches: size -0.4 % on log text against entering all of them).
        //  (1a) candidates inside the group (what matters is short-period data, where any lower member of the run
        //       serves; done in (2), off the token's path).  Even lanes write their position into a private scratch table
        //       with ONE store: when lanes of a warp store to the same shared-memory address, sm_100 keeps the value of
        //       the lowest lane (tools/micro/sts_winner.cu; architecturally unspecified), so a slot ends up holding the
        //       lowest position of the group that hashes to it.  Every lane reads its slot back and keeps what it finds
        //       if that lies in this group, below it, and really has the same hash (one shuffle).  Were the winner any
        //       other lane the candidates would still be valid, only fewer: tests/test_gpu_compress.py pins the size on
        //       short-period data.
        //  (1b) with the token: per group one read of the table slot (the entry from earlier groups) and one write by
        //       the even lanes — 64 shared-memory instructions that depend on nothing but the slots, back to back.
        // Shared-memory atomics would do the exchange in one instruction but retire about one lane every two cycles
        // per SM (measured: 165 cycles per 32-lane atomicMax with two CTAs resident), and match.any costs ~390.
        const uint32_t* __restrict__ wt = w32 + (tile_base >> 2) + (lane >> 2);
        const uint32_t lsh = (uint32_t)(lane & 3) * 8u;
        const int room = hash_end - tile_base - lane;             // group g is hashed iff 32 g < room
        const int q = (lane & 1) ? -1 : 0;                        // >= 0: an entering lane (even positions)
        {
            uint8_t* tb = reinterpret_cast<uint8_t*>(S.table);
            uint32_t hs2[16];                                     // byte offsets of the slots, two groups per register
#pragma unroll
            for (int g = 0; g < 32; g++) {
                const uint32_t h = slot_offset<kLarge>(wt[8 * g], wt[8 * g + 1], lsh);
                if (g & 1) hs2[g >> 1] |= h << 16; else hs2[g >> 1] = h;
            }
            PROF(0);
            mbar_wait(&S.bar_token[t], par);
            PROF(1);
            uint32_t old[32];
#pragma unroll
            for (int g = 0; g < 32; g++) {
                volatile uint16_t* slot = reinterpret_cast<volatile uint16_t*>(tb + ((hs2[g >> 1] >> (16 * (g & 1))) & 0xFFFFu));
                old[g] = *slot;
                if (q >= 0 && 32 * g < room) *slot = (uint16_t)(tile_base + 32 * g + lane);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.bar_token[t + 1]);
            PROF(2);
#pragma unroll
            for (int g = 0; g < 32; g++) pv[g * kPvStride + lane] = (uint16_t)old[g];
        }
        if (kLarge && C.index_only) {                                    // table warmed; keep the other chains' phases in step
            if (lane == 0) { mbar_arrive(&S.bar_entry[t + 1]); mbar_arrive(&S.bar_out[t + 1]); mbar_arrive(&S.bar_done[t]); }
            continue;
        }
        const int xhint = my_x;          // a lower bound of this tile's entry state that does not depend on timing: this worker's previous tile

        // ---- (2) candidate of every position — inside the group if there is one (1a), else the table's — and a first
        // check: lane g ends up with the map of positions b0 .. b0+31 whose candidate agrees in the bytes its first
        // aligned word holds (1 to 4 of them); the parse checks the rest.  Eight groups per round, every stage of a
        // round issued for all eight before its results are used.
        uint32_t bits = 0;
        if (xhint < tile_base + kTile) {
            uint8_t* sb = reinterpret_cast<uint8_t*>(S.recs[pw]); // 512 x u16 of scratch (the records come later)
#pragma unroll 1
            for (int gb = 0; gb < 32; gb += 8) {
                const uint32_t* __restrict__ wg = wt + 8 * gb;
                uint16_t* pg = pv + gb * kPvStride + lane;
                uint32_t own[8], h[8], w[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    own[j] = __funnelshift_r(wg[8 * j], wg[8 * j + 1], lsh);
                    h[j] = slot_offset<kLarge>(wg[8 * j], wg[8 * j + 1], lsh);
                    volatile uint16_t* sc = reinterpret_cast<volatile uint16_t*>(sb + (h[j] & 0x3FEu));
                    const bool enter = q >= 0 && 32 * (gb + j) < room;
                    const uint16_t p16 = (uint16_t)(tile_base + 32 * (gb + j) + lane);
                    if (enter) *sc = p16;            // lanes sharing a slot: the lowest one's value stays (see (1a))
                    __syncwarp();
                    w[j] = *sc;
                }
                uint32_t c[8];
#pragma unroll
                for (int j = 0; j < 8; j++) c[j] = pg[j * kPvStride];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    // 16-bit arithmetic: a group never straddles a multiple of 65536
                    const uint32_t g16 = (uint32_t)(tile_base + 32 * (gb + j)) & 0xFFFFu;
                    const uint32_t hw = __shfl_sync(FULL_MASK, h[j], (int)(w[j] & 31u));
                    // (the scratch is cleared when the CTA starts, so what is read here was written by this worker while it
                    // worked on this block: the choice of candidate never depends on what ran on the SM before)
                    if (w[j] >= g16 && w[j] < g16 + (uint32_t)lane && hw == h[j]) {
                        c[j] = w[j];
                        pg[j * kPvStride] = (uint16_t)w[j];
                    }
                }
                uint32_t cw[8];
                int ca[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int p = tile_base + 32 * (gb + j) + lane;
                    if (kLarge) {
                        ca[j] = cand_abs(p, c[j]);
                        // none: an empty slot, a position from before the span, or one exactly 65536 back (same low bits: an
                        // offset LZ4 cannot express)
                        if (c[j] == 0xFFFFu || ca[j] < C.lo_valid || ca[j] == p - 65536) ca[j] = p;
                    } else {
                        ca[j] = (int)c[j] < p ? (int)c[j] : p;
                    }
                    cw[j] = w32[ca[j] >> 2];
                }
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int p = tile_base + 32 * (gb + j) + lane;
                    const uint32_t sh = (uint32_t)(ca[j] & 3) * 8u;
                    const bool okb = ca[j] < p && p < C.mf_end && (((cw[j] >> sh) ^ own[j]) & (0xFFFFFFFFu >> sh)) == 0;
                    const uint32_t word = __ballot_sync(FULL_MASK, okb);
                    if (lane == gb + j) bits = word;
                }
            }
        }
        __syncwarp();

        PROF(3);
        // ---- (2) every lane parses its own 32 positions
        int cnt = 0, last_q = 0, last_ml = 0, last_off = 0;
        bool unfin = false;
        {
            int la = max(b0, xhint);
            bits = (la - b0 >= 32) ? 0u : (bits & (0xFFFFFFFFu << (la - b0)));
            while (bits) {
                const int r = __ffs(bits) - 1;
                int p = b0 + r;
                int c = kLarge ? cand_abs(p, pv[lane * kPvStride + r]) : (int)pv[lane * kPvStride + r];
                int ml;
                bool un;
                lane_extend(w32, p, c, C.match_end - p, ml, un);
                if (ml < MINMATCH) { bits &= bits - 1; continue; }               // the partial check of (2) let it through
                if (ml < kLazyBelow && r < 31 && ((bits >> (r + 1)) & 1u)) {     // one-step lazy: is the next position better?
                    // not if it continues the same source (one byte shorter by construction), and only if the byte
                    // that would make it longer is there
                    const int c2 = kLarge ? cand_abs(p + 1, pv[lane * kPvStride + r + 1]) : (int)pv[lane * kPvStride + r + 1];
                    if (c2 != c + 1 && win[p + 1 + ml] == win[c2 + ml]) {
                        int ml2;
                        bool un2;
                        lane_extend(w32, p + 1, c2, C.match_end - p - 1, ml2, un2);
                        if (ml2 > ml) { p++; c = c2; ml = ml2; un = un2; }
                    }
                }
                int q = p, cc = c;
                while (q > la && cc > C.lo_valid && win[q - 1] == win[cc - 1]) { q--; cc--; ml++; }
                recs[cnt][lane] = ((uint32_t)(q - cc) << 16) | ((uint32_t)min(ml, 2047) << 5) | (uint32_t)(q - b0);
                cnt++;
                last_q = q; last_ml = ml; last_off = q - cc; unfin = un;
                la = q + ml;
                const int rel = la - b0;
                bits = (un || rel >= 32) ? 0u : (bits & (0xFFFFFFFFu << rel));
            }
        }
        __syncwarp();                                                    // prev[] is dead from here on: its bytes become the staging tile

        PROF(4);
        // ---- long matches: completed by the whole warp, lowest lane first; one that starts inside a completed
        // match is dropped (the match before it covers its start; only its tail beyond could have been used)
        {
            int covered = 0;
            for (uint32_t todo = __ballot_sync(FULL_MASK, unfin); todo; todo &= todo - 1) {
                const int l = __ffs(todo) - 1;
                const int q = __shfl_sync(FULL_MASK, last_q, l), ml = __shfl_sync(FULL_MASK, last_ml, l);
                const int off = __shfl_sync(FULL_MASK, last_off, l);
                if (q < covered) {
                    if (lane == l) {
                        cnt--;
                        if (cnt > 0) {
                            const uint32_t rc = recs[cnt - 1][lane];
                            last_q = b0 + (int)(rc & 31u); last_ml = (int)((rc >> 5) & 2047u); last_off = (int)(rc >> 16);
                        }
                    }
                    continue;
                }
                const int more = warp_count_equal(w32, q + ml, q + ml - off, C.match_end - q - ml, lane);
                if (lane == l) last_ml = ml + more;
                covered = q + ml + more;
            }
        }

        // ---- (3) resolve against the entry state of the tile
        const int E = cnt ? last_q + last_ml : 0;
        const int pm = warp_incl_max(E, lane);
        int Xl = __shfl_up_sync(FULL_MASK, pm, 1);                       // where the lanes before this one stop, as far as this tile knows
        if (lane == 0) Xl = 0;
        const int pm_all = __shfl_sync(FULL_MASK, pm, 31);
        PROF(5);
        mbar_wait(&S.bar_entry[t], par);
        PROF(6);
        // Two chains run through the tiles.  The first carries (covered-up-to, literal anchor) and is a handful of
        // instructions per tile: the next tile's lanes need it before they can size their sequences.  The second
        // carries the output offset, which takes the sizes of all lanes — work that tiles do side by side.
        const int x_in = S.st_x[t], anchor_in = S.st_anchor[t];
        int X = max(Xl, x_in);                                           // matches of this lane may start here
        bool emits = false;
        if (cnt) {
            const int st = max(last_q, X);
            emits = (E - st >= MINMATCH) && st < C.mf_end;
        }
        const int am = warp_incl_max(emits ? E : 0, lane);
        const int x_out = max(x_in, pm_all);
        my_x = x_out;
        const int anchor_out = max(anchor_in, __shfl_sync(FULL_MASK, am, 31));
        if (lane == 0) {
            S.st_x[t + 1] = x_out; S.st_anchor[t + 1] = anchor_out;
            mbar_arrive(&S.bar_entry[t + 1]);
        }
        int A = __shfl_up_sync(FULL_MASK, am, 1);
        if (lane == 0) A = 0;
        A = max(A, anchor_in);                                           // literal run of this lane's first sequence starts here
        int size = 0;
        if (emits) {
            int lit_from = A;
            for (int i = 0; i < cnt; i++) {
                const uint32_t rc = recs[i][lane];
                const int q = b0 + (int)(rc & 31u);
                const int ml = (i == cnt - 1) ? last_ml : (int)((rc >> 5) & 2047u);
                const int st = max(q, X), len = q + ml - st;
                if (len < MINMATCH || st >= C.mf_end) continue;
                size += seq_size(st - lit_from, len);
                lit_from = st + len;
            }
        }
        const int incl = warp_incl_sum(size, lane);
        const int total = __shfl_sync(FULL_MASK, incl, 31);
        PROF(7);
        mbar_wait(&S.bar_out[t], par);
        const int out_in = S.st_out[t];
        const int out_out = out_in + total;
        const bool failed = (*reinterpret_cast<volatile int*>(&S.fail) != 0) || out_out > cap;
        if (lane == 0) {
            S.st_out[t + 1] = out_out;
            if (failed) *reinterpret_cast<volatile int*>(&S.fail) = 1;
            mbar_arrive(&S.bar_out[t + 1]);
        }

        PROF(9);
        // ---- (4) emit: into the staging tile when the tile's bytes fit there (then out with 16-byte stores),
        // straight to global memory otherwise (long literal runs: poorly compressible data)
        if (!failed && total > 0) {
            uint8_t* gdst = payload + out_in;
            const int g0 = (int)(reinterpret_cast<uintptr_t>(gdst) & 15u);
            const bool staged = g0 + total <= kStage;
            int long_from = 0, long_n = 0, long_at = 0;                  // a literal run left to the whole warp: source, length, offset in the tile's bytes
            auto emit_lane = [&](uint8_t* base) {
                uint8_t* o = base + (incl - size);
                int lit_from = A;
                for (int i = 0; i < cnt; i++) {
                    const uint32_t rc = recs[i][lane];
                    const int q = b0 + (int)(rc & 31u);
                    const int ml = (i == cnt - 1) ? last_ml : (int)((rc >> 5) & 2047u);
                    const int st = max(q, X), len = q + ml - st;
                    if (len < MINMATCH || st >= C.mf_end) continue;
                    const int lit = st - lit_from, mc = len - MINMATCH;
                    const uint32_t off = rc >> 16;
                    *o++ = (uint8_t)(((lit < 15 ? lit : 15) << 4) | (mc < 15 ? mc : 15));
                    if (lit >= 15) { put_ext_bytes(o, lit - 15); o += ext_len(lit - 15); }
                    if (lit >= kLongLit) { long_from = lit_from; long_n = lit; long_at = (int)(o - base); }
                    else for (int j = 0; j < lit; j++) o[j] = win[lit_from + j];
                    o += lit;
                    o[0] = (uint8_t)off; o[1] = (uint8_t)(off >> 8);
                    o += 2;
                    if (mc >= 15) { put_ext_bytes(o, mc - 15); o += ext_len(mc - 15); }
                    lit_from = st + len;
                }
            };
            if (emits) {
                if (staged) emit_lane(S.stage[pw] + g0); else emit_lane(gdst);
            }
            // long literal runs: only a lane's first sequence can have one
            for (uint32_t todo = __ballot_sync(FULL_MASK, long_n > 0); todo; todo &= todo - 1) {
                const int l = __ffs(todo) - 1;
                const int from = __shfl_sync(FULL_MASK, long_from, l), cnt_l = __shfl_sync(FULL_MASK, long_n, l);
                const int at = __shfl_sync(FULL_MASK, long_at, l);
                // a long run may begin behind the window when blocks are large: those come from global memory
                if (staged) { for (int k = lane; k < cnt_l; k += 32) S.stage[pw][g0 + at + k] = kLarge ? C.gsrc[from + k] : win[from + k]; }
                else if (kLarge) warp_copy(gdst + at, C.gsrc + from, (uint32_t)cnt_l, lane);
                else copy_s2g(gdst + at, win + from, cnt_l, lane, 32);
            }
            __syncwarp();
            if (staged) {
                uint8_t* gb = gdst - g0;                                  // 16-byte aligned
                const int end = g0 + total;
                for (int c16 = lane * 16; c16 < end; c16 += 512) {
                    const int lo = max(c16, g0), hi = min(c16 + 16, end);
                    if (hi - lo == 16) *reinterpret_cast<uint4*>(gb + c16) = *reinterpret_cast<const uint4*>(stage + c16);
                    else for (int j = lo; j < hi; j++) gb[j] = stage[j];
                }
                __syncwarp();
            }
        }
        if (lane == 0) mbar_arrive(&S.bar_done[t]);
        PROF(8);
    }
}

// ---------------------------------------------------------------- hasher / finisher

}  // namespace

__global__ void __launch_bounds__(kCtaThreads, 2)
lz4_compress_cta_kernel(EncodeArgs a)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    CtaSmem& S = *reinterpret_cast<CtaSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t b = blockIdx.x;
    const uint8_t* src = a.src_base + a.src_off[b];
    const int n_in = (int)a.src_len[b];
    uint8_t* rec = a.rec_base + (uint64_t)b * a.rec_stride;
    uint8_t* payload = a.raw_blocks ? rec : rec + 4;
    const int cap = (int)a.dst_cap;
    if (a.split_by_size && n_in > 65536) return;              // the span kernel's block
    // a block this kernel cannot hold is stored / refused (launch_compress never sends one: see kernels.h)
    const bool oversize = n_in > 65536;
    const int n = oversize ? 0 : n_in;
    const int ntiles = n >= MFLIMIT + 1 ? (n - MFLIMIT + 1 + kTile - 1) / kTile : 0;
    const bool aligned = (reinterpret_cast<uintptr_t>(src) & 15u) == 0;
    const uint32_t bulk = aligned ? ((uint32_t)n & ~15u) : 0u;

    if (tid == 0) {
        mbar_init(&S.bar_load, 1);
        for (int i = 0; i <= ntiles; i++) mbar_init(&S.bar_token[i], 1);
        for (int i = 0; i <= ntiles; i++) { mbar_init(&S.bar_entry[i], 1); mbar_init(&S.bar_out[i], 1); }
        for (int i = 0; i < ntiles; i++) mbar_init(&S.bar_done[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        S.st_x[0] = 0; S.st_anchor[0] = 0; S.st_out[0] = 0;
        S.fail = 0;
        if (bulk) {
            mbar_arrive_expect_tx(&S.bar_load, bulk);
            tma_load_bulk(S.win, src, bulk, &S.bar_load);                 // the block, once, into shared memory
        }
    }
    {
        uint4* t4 = reinterpret_cast<uint4*>(S.table);
        const uint4 fill = make_uint4(~0u, ~0u, ~0u, ~0u);       // -1 everywhere
        for (int i = tid; i < (int)(sizeof(S.table) / 16); i += kCtaThreads) t4[i] = fill;
        uint4* r4 = reinterpret_cast<uint4*>(S.recs);            // the workers' scratch tables start empty (see run_worker (2))
        for (int i = tid; i < (int)(sizeof(S.recs) / 16); i += kCtaThreads) r4[i] = fill;
        // what the bulk copy does not bring: the last n & 15 bytes (or everything, from an unaligned source), zero padding
        if (aligned) {
            for (int k = (int)bulk + tid; k < n; k += kCtaThreads) S.win[k] = src[k];
        } else {
            const int head = min(n, (int)((16u - (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u)) & 15u));
            for (int k = tid; k < head; k += kCtaThreads) S.win[k] = src[k];
            const uint4* s16 = reinterpret_cast<const uint4*>(src + head);
            const int nvec = (n - head) >> 4;
            for (int k = tid; k < nvec; k += kCtaThreads) {
                const uint4 v = s16[k];
                uint8_t* d = S.win + head + 16 * k;                      // shared side is misaligned by `head`
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 16; j++) d[j] = (uint8_t)(w[j >> 2] >> (8 * (j & 3)));
            }
            for (int k = head + (nvec << 4) + tid; k < n; k += kCtaThreads) S.win[k] = src[k];
        }
        for (int k = n + tid; k < n + kWinPad; k += kCtaThreads) S.win[k] = 0;
    }
    __syncthreads();
    if (tid == 0) {
        if (!bulk) mbar_arrive(&S.bar_load);
        mbar_arrive(&S.bar_entry[0]);
        mbar_arrive(&S.bar_out[0]);
        mbar_arrive(&S.bar_token[0]);
    }
    mbar_wait(&S.bar_load, 0);

    if (warp < kWorkers) {
        TileEnv env;
        env.w32 = reinterpret_cast<const uint32_t*>(S.win); env.win = S.win; env.gsrc = src; env.payload = payload;
        env.pos0 = 0; env.ntiles = ntiles;
        env.hash_end = n - 3; env.mf_end = n - MFLIMIT + 1; env.match_end = n - LASTLITERALS;
        env.lo_valid = 0; env.cap = cap; env.parity = 0; env.index_only = false;
        int my_x = 0;                                         // covered-up-to at the exit of this worker's previous tile
        run_worker<false, kWorkers>(S, env, warp, lane, my_x);
    } else if (a.block_checksum && !a.raw_blocks) {
        // block checksum as the payload appears: whole 512-byte chunks behind the last completed tile
        const uintptr_t pa = reinterpret_cast<uintptr_t>(payload);
        const uint32_t* wp = reinterpret_cast<const uint32_t*>(pa & ~uintptr_t(3));
        const uint32_t sh = (uint32_t)(pa & 3u) * 8u;
        uint32_t acc = xxh32_init(lane);
        int hashed = 0;
        for (int t = 0; t < ntiles; t++) {
            mbar_wait(&S.bar_done[t], 0);
            if (*reinterpret_cast<volatile int*>(&S.fail)) break;
            const int avail = S.st_out[t + 1];
            const int nch = (avail - hashed) / kHashChunk;
            if (nch > 0) {
                acc = xxh32_consume_global(acc, wp + hashed / 4, sh, nch, lane);
                hashed += nch * kHashChunk;
            }
        }
        // park the running state for the finish below (same warp)
        S.hash_acc[lane] = acc;
        if (lane == 0) S.hash_done = hashed;
    }
    __syncthreads();

    // ---- last literals (lz4.c:1302-1329), stored fallback, framing (blk/blk.go:78-106)
    const int anchor = S.st_anchor[ntiles], out_end = S.st_out[ntiles];
    const int run = n - anchor;
    const int run_ext = run >= 15 ? ext_len(run - 15) : 0;
    int c = out_end + 1 + run_ext + run;
    const bool fits = !oversize && S.fail == 0 && c <= cap;
    if (fits) {
        if (tid == 0) {
            payload[out_end] = (uint8_t)((run < 15 ? run : 15) << 4);
            if (run >= 15) put_ext_bytes(payload + out_end + 1, run - 15);
        }
        copy_s2g(payload + out_end + 1 + run_ext, S.win + anchor, run, tid, kCtaThreads);
    } else {
        c = 0;
    }
    if (a.raw_blocks) {
        if (tid == 0) a.rec_len[b] = (uint32_t)c;          // 0 = does not fit (clz4.go:40-42)
        return;
    }
    uint32_t word = (uint32_t)c;
    if (c == 0) {                                           // blk/blk.go:78-92: store raw
        if (oversize) {
            for (int k = tid; k < n_in; k += kCtaThreads) payload[k] = src[k];
        } else {
            copy_s2g(payload, S.win, n, tid, kCtaThreads);
        }
        c = n_in;
        word = (uint32_t)n_in | 0x80000000u;
    }
    __syncthreads();                                        // the payload is complete (block-wide visibility)
    if (warp != kWorkers) return;
    if (lane == 0) store_le32(rec, word);
    uint32_t total = 4u + (uint32_t)c;
    if (a.block_checksum) {
        uint32_t acc = xxh32_init(lane);
        uint32_t hashed = 0;
        uint32_t x;
        if (fits) {
            acc = S.hash_acc[lane]; hashed = (uint32_t)S.hash_done;
            const uintptr_t pa = reinterpret_cast<uintptr_t>(payload);
            const int nch = (int)(((uint32_t)c - hashed) / kHashChunk);
            acc = xxh32_consume_global(acc, reinterpret_cast<const uint32_t*>(pa & ~uintptr_t(3)) + hashed / 4, (uint32_t)(pa & 3u) * 8u, nch, lane);
            hashed += (uint32_t)nch * kHashChunk;
            x = xxh32_finish_global(acc, payload, hashed, (uint32_t)c, lane);
        } else if (!oversize) {
            // a stored block is the window itself: hash it where it lies
            acc = xxh32_consume_words(acc, reinterpret_cast<const uint32_t*>(S.win), (uint32_t)n >> 4, lane);
            x = xxh32_finish(acc, S.win, (uint32_t)n);
        } else {
            const uintptr_t pa = reinterpret_cast<uintptr_t>(payload);
            const int nch = c / kHashChunk;
            acc = xxh32_consume_global(acc, reinterpret_cast<const uint32_t*>(pa & ~uintptr_t(3)), (uint32_t)(pa & 3u) * 8u, nch, lane);
            x = xxh32_finish_global(acc, payload, (uint32_t)nch * kHashChunk, (uint32_t)c, lane);
        }
        if (lane == 0) store_le32(payload + c, x);
        total += 4;
    }
    if (lane == 0) a.rec_len[b] = total;
}


// ---------------------------------------------------------------- blocks larger than 64 KiB
//
// One CTA per SPAN: a run of consecutive 64 KiB fragments of one block, encoded one after the other by the same
// workers with the table kept.  The window holds the fragment before the current one next to it (moved down when
// the next one is loaded), so a candidate up to 65535 bytes back is always in shared memory and the parse behaves as
// liblz4's does on a large block — five-byte hash included (lz4.c:785-795,1391-1400).  A match stops at the end of its
// fragment (the next one is not loaded yet) and goes on as a new sequence with no literals: four or five bytes per
// 64 KiB on data that is one long match.  A block is cut into several spans only when the launch has fewer blocks than
// the GPU has SMs; a span that does not begin its block first runs the fragment before it through the index alone, and
// lz4_stitch_kernel (compress.cu) joins the spans' streams: the literals left at the end of one become part of the
// first sequence of the next.
__global__ void __launch_bounds__(kSpanThreads, 1)
lz4_compress_span_kernel(EncodeArgs a, uint8_t* tmp, uint32_t slot_stride, uint32_t spans_per_block, uint32_t span_bytes,
                         int32_t* span_len, uint32_t* span_tail)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    SpanSmem& S = *reinterpret_cast<SpanSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t w = blockIdx.x, b = w / spans_per_block, sp = w % spans_per_block;
    if (b >= a.nblk) return;
    const uint8_t* src = a.src_base + a.src_off[b];
    const int n_blk = (int)a.src_len[b];
    if (n_blk <= 65536) return;                                   // the CTA kernel's block
    // With a dictionary the positions are those of a virtual block that starts with the dictionary's 64 KiB slot (the
    // dictionary at its end, lz4.c:1556-1604: it lies just below the block): the fragment before a block's first span is
    // then the dictionary, indexed like any fragment before a span, and offsets into it come out as plain distances.
    const int shift = a.dict_size ? 65536 : 0;
    const int n_v = n_blk + shift;
    const int span_start = (int)(sp * span_bytes) + shift;
    if (span_start >= n_v && sp != 0) {
        if (tid == 0) span_len[w] = -2;                           // this span does not exist
        return;
    }
    const int span_end = min(n_v, span_start + (int)span_bytes);
    uint8_t* out = tmp + (uint64_t)w * slot_stride;

    if (tid == 0) {
        mbar_init(&S.bar_load, 1);
        for (int i = 0; i <= kMaxTiles; i++) { mbar_init(&S.bar_token[i], 1); mbar_init(&S.bar_entry[i], 1); mbar_init(&S.bar_out[i], 1); }
        for (int i = 0; i < kMaxTiles; i++) mbar_init(&S.bar_done[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        S.fail = 0;
    }
    {
        uint4* t4 = reinterpret_cast<uint4*>(S.table);
        const uint4 fill = make_uint4(~0u, ~0u, ~0u, ~0u);
        for (int i = tid; i < (int)(sizeof(S.table) / 16); i += kSpanThreads) t4[i] = fill;
        uint4* r4 = reinterpret_cast<uint4*>(S.recs);
        for (int i = tid; i < (int)(sizeof(S.recs) / 16); i += kSpanThreads) r4[i] = fill;
    }
    uint8_t* cur = S.win + 65536;                                 // the current fragment; the one before it lies below
    int my_x = span_start;
    int carry_x = span_start, carry_anchor = span_start, carry_out = 0;
    uint32_t it = 0;
    const bool has_before = sp != 0 || shift != 0;                // a fragment (or the dictionary) lies before the span
    for (int fs = span_start - (has_before ? 65536 : 0); fs < span_end; fs += 65536, it++) {
        const bool warm = fs < span_start;
        const int fe = warm ? fs + 65536 : min(fs + 65536, span_end);      // end of the fragment's bytes
        __syncthreads();                                          // every worker is done with the fragment before
        if (it > 0) {
            uint4* d = reinterpret_cast<uint4*>(S.win);
            const uint4* s4 = reinterpret_cast<const uint4*>(cur);
            for (int i = tid; i < 4096; i += kSpanThreads) d[i] = s4[i];
        }
        __syncthreads();
        const bool dict_frag = fs < shift;                        // the dictionary's slot: zeros, then the dictionary
        const uint8_t* fsrc = src + (fs - shift);
        const int fn = fe - fs;
        const bool aligned = !dict_frag && (reinterpret_cast<uintptr_t>(fsrc) & 15u) == 0;
        const uint32_t bulk = aligned ? ((uint32_t)fn & ~15u) : 0u;
        if (tid == 0 && bulk) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the bulk copy overwrites bytes just read by ordinary loads
            mbar_arrive_expect_tx(&S.bar_load, bulk);
            tma_load_bulk(cur, fsrc, bulk, &S.bar_load);
        }
        if (dict_frag) {
            const int gap = 65536 - (int)a.dict_size;
            for (int k = tid; k < gap; k += kSpanThreads) cur[k] = 0;
            for (int k = tid; k < (int)a.dict_size; k += kSpanThreads) cur[gap + k] = a.dict[k];
        } else {
            for (int k = (int)bulk + tid; k < fn; k += kSpanThreads) cur[k] = fsrc[k];
        }
        for (int k = fn + tid; k < fn + kWinPad && k < 65536 + kWinPad; k += kSpanThreads) cur[k] = 0;
        TileEnv env;
        env.w32 = reinterpret_cast<const uint32_t*>(S.win) + ((65536 - fs) >> 2);   // fs is a multiple of 65536
        env.win = S.win + (65536 - fs);
        env.gsrc = src - shift; env.payload = out; env.pos0 = fs;                   // (read at block positions only)
        env.lo_valid = sp != 0 ? span_start - 65536 : shift - (int)a.dict_size;
        env.hash_end = min(n_v - 4, fe - 4);                      // five bytes at p, all of them loaded
        env.mf_end = warm ? fs : min(n_v - MFLIMIT + 1, fe - 3);
        env.match_end = min(n_v - LASTLITERALS, fe);
        env.ntiles = warm ? kMaxTiles : max(0, (env.mf_end - fs + kTile - 1) / kTile);
        env.cap = 0x7FFFFFF0; env.parity = it & 1u; env.index_only = warm;
        if (tid == 0) {
            S.st_x[0] = carry_x; S.st_anchor[0] = carry_anchor; S.st_out[0] = carry_out;
        }
        __syncthreads();
        if (tid == 0) {
            if (!bulk) mbar_arrive(&S.bar_load);
            mbar_arrive(&S.bar_token[0]); mbar_arrive(&S.bar_entry[0]); mbar_arrive(&S.bar_out[0]);
        }
        mbar_wait(&S.bar_load, it & 1u);
        run_worker<true, kSpanWorkers>(S, env, warp, lane, my_x);
        __syncthreads();
        if (!warm && env.ntiles > 0) {
            carry_x = S.st_x[env.ntiles]; carry_anchor = S.st_anchor[env.ntiles]; carry_out = S.st_out[env.ntiles];
        }
        // barriers of tiles this fragment did not have stay one phase behind: only the last fragment of a block is short
    }
    if (tid == 0) {
        span_len[w] = carry_out;                                  // bytes of sequences; the final literal run is the stitcher's
        span_tail[w] = (uint32_t)(span_end - carry_anchor);
    }
}

cudaError_t configure_compress_cta()
{
    cudaError_t e = cudaFuncSetAttribute(lz4_compress_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CtaSmem));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(lz4_compress_span_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SpanSmem));
}

cudaError_t launch_compress_spans(const EncodeArgs& a, uint8_t* tmp, uint32_t slot_stride, uint32_t spans_per_block, uint32_t span_bytes,
                                  int32_t* span_len, uint32_t* span_tail, cudaStream_t stream)
{
    lz4_compress_span_kernel<<<a.nblk * spans_per_block, kSpanThreads, sizeof(SpanSmem), stream>>>(a, tmp, slot_stride, spans_per_block,
                                                                                                   span_bytes, span_len, span_tail);
    return cudaGetLastError();
}

cudaError_t launch_compress_cta(const EncodeArgs& a, cudaStream_t stream)
{
    lz4_compress_cta_kernel<<<a.nblk, kCtaThreads, sizeof(CtaSmem), stream>>>(a);
    return cudaGetLastError();
}

}  // namespace plz4
