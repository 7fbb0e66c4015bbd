// compress.cu — LZ4 level-1 block encode, one warp per block (sm_100a).
//
// Replaces, for a whole batch of independent blocks, what plz4 does per block on a goroutine:
//   async/writer.go:232-282 compressLoop -> blk/blk.go:69-109 CompressToBlk
//     -> compress/indie.go:66-74 -> clz4.go:31-45 -> lz4.c:930-1338 LZ4_compress_generic_validated
//   plus the record framing (size word / stored fallback / xxh32 trailer) of blk/blk.go:87-106.
//
// It is NOT liblz4's serial parse.  Per 32-position group the warp
//   (a) hashes all 32 positions at once (the bytes of the next two groups are already in registers /
//       in flight), looks every one up in a per-warp shared-memory table and resolves same-group
//       duplicates with match.any, so each position sees its most recent earlier occurrence (what
//       liblz4's mutating table gives it one position at a time); then EVERY lane measures its own
//       candidate in parallel — two aligned 16-byte loads around the candidate, own bytes from the
//       neighbours' registers by shuffle — giving a match length of up to 15, a backward byte and a
//       one-step-lazy hint, then
//   (b) walks the measured candidates with no memory access on the common path (only matches longer
//       than 15 go back to memory) and appends every chosen sequence to a shared-memory queue;
//       every 32 sequences the queue is written out one sequence per lane (sizes prefix-summed,
//       each lane stores its own token, length bytes, literals and offset).
// The table holds the low 16 bits of positions: LZ4's window is 64 KiB, so that is enough for any
// block size (a stale entry aliases into the window and is caught by the byte comparison).
// The output is a valid LZ4 block obeying the end-of-block rules liblz4's decoder enforces
// (last match starts <= n-12 and ends <= n-5, lz4.c:245-246,963-964); bytes differ from liblz4's,
// size stays within the tolerance pinned by tests/test_gpu_compress.py.
#include "common.cuh"
#include "kernels.h"

#include <algorithm>
#include <cstdlib>

namespace plz4 {

constexpr int kEncodeWarps = 4;                 // blocks per CTA
constexpr int kLongLiterals = 64;               // literal runs from this length on are copied by the whole warp
constexpr int kQueueLen = 48;                   // sequences queued per warp: 32 to flush + the <= 8 a group can add

// Table geometry.  The 12-bit table uses 7/8 of its 4096 slots: 3584 x u16 = 7 KiB per warp, which lets a seventh
// CTA (28 warps instead of 24) fit an SM's 227 KiB of shared memory next to the sequence queues.
__host__ __device__ constexpr int table_slots(int bits) { return bits == 12 ? 3584 : (1 << bits); }
__host__ __device__ constexpr int table_bytes(int bits) { return 2 * table_slots(bits); }
__host__ __device__ __forceinline__ uint32_t table_slot(uint32_t v, int bits)
{
    const uint32_t h = (v * 2654435761u) >> (32 - bits);       // liblz4's hash4 (lz4.c:777-783)
    return bits == 12 ? (h * 7u) >> 3 : h;
}
// liblz4's hash5 (lz4.c:785-795): blocks of 64 KiB + 11 bytes and more are hashed on FIVE bytes (lz4.c:1391-1400 picks
// byU32 there, and LZ4_hashPosition hashes 5 bytes for every table type but byU16 on 64-bit builds).  The top `bits` bits of
// ((sequence << 24) * 889523592379) are bits [40-bits, 40) of sequence * prime mod 2^40; b4 is the fifth byte.
__host__ __device__ __forceinline__ uint32_t table_slot5(uint32_t v, uint32_t b4, int bits)
{
    const uint64_t lo = (uint64_t)v * 0x1BBCDCBBull;
    const uint32_t top = ((uint32_t)(lo >> 32) + v * 0xCFu + b4 * 0x1BBCDCBBu) & 0xFFu;
    const uint32_t h = ((top << (bits - 8)) | ((uint32_t)lo >> (40 - bits))) & ((1u << bits) - 1u);
    return bits == 12 ? (h * 7u) >> 3 : h;
}
__constant__ int g_back_dev = 1;                // tuning knobs (see configure_compress)
__constant__ int g_jump_dev = 64;
__constant__ int g_lazy_dev = 1;                // 0 = greedy; k>0 = take p+1 if its match is longer by >= k

// x << n with the hardware's semantics (a shift count >= 32 gives 0), which C leaves undefined
__device__ __forceinline__ uint32_t shl_sat(uint32_t x, int n)
{
    uint32_t r;
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(n));
    return r;
}

// bytes needed for the 255-run extension of a length whose nibble saturated
__device__ __forceinline__ int ext_bytes(int rest) { return rest / 255 + 1; }

__device__ __forceinline__ void put_ext(uint8_t* o, int rest, int lane)
{
    int nb = ext_bytes(rest);
    for (int k = lane; k < nb; k += 32) o[k] = (k == nb - 1) ? (uint8_t)(rest - 255 * (nb - 1)) : (uint8_t)255;
}

// Long-match tail: number of equal bytes of a[0..] and b[0..], at most `limit`.  64 bytes per round, all four loads in
// flight before the first ballot.
__device__ __noinline__ int count_equal(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int limit, int lane)
{
    int total = 0;
    for (;;) {
        const int k0 = total + lane, k1 = k0 + 32;
        uint32_t a0 = 0, b0 = 1, a1 = 0, b1 = 1;
        if (k0 < limit) { a0 = a[k0]; b0 = b[k0]; }
        if (k1 < limit) { a1 = a[k1]; b1 = b[k1]; }
        const uint32_t n0 = __ballot_sync(FULL_MASK, a0 != b0);
        if (n0) return total + (__ffs(n0) - 1);
        const uint32_t n1 = __ballot_sync(FULL_MASK, a1 != b1);
        if (n1) return total + 32 + (__ffs(n1) - 1);
        total += 64;
    }
}

// Sequences chosen by the walk wait in a per-warp shared-memory queue as (literal start, literal count, match
// length, offset) and are written 32 at a time, one sequence per lane: sizes are prefix-summed, each lane stores
// its own token, length bytes, literals and offset.  Returns the bytes written, or -1 if they do not fit `room`.
__device__ __noinline__ int flush_queue(const uint4* queue, int cnt, const uint8_t* __restrict__ src, uint8_t* out, int room, int lane)
{
    const bool mine = lane < cnt;
    const uint4 e = mine ? queue[lane] : make_uint4(0, 0, MINMATCH, 0);
    const int lit = (int)e.y, ml = (int)e.z - MINMATCH;
    const int le = lit >= 15 ? ext_bytes(lit - 15) : 0;
    const int me = ml >= 15 ? ext_bytes(ml - 15) : 0;
    const int size = mine ? 1 + le + lit + 2 + me : 0;
    int incl = size;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int up = __shfl_up_sync(FULL_MASK, incl, d);
        if (lane >= d) incl += up;
    }
    const int total = __shfl_sync(FULL_MASK, incl, 31);
    if (total > room) return -1;
    if (mine) {
        uint8_t* o = out + (incl - size);
        *o++ = (uint8_t)(((lit < 15 ? lit : 15) << 4) | (ml < 15 ? ml : 15));
        if (le) { int r = lit - 15; for (; r >= 255; r -= 255) *o++ = 255; *o++ = (uint8_t)r; }
        const uint8_t* lp = src + (int)e.x;
        if (lit < kLongLiterals) for (int j = 0; j < lit; j++) o[j] = lp[j];
        o += lit;
        o[0] = (uint8_t)e.w; o[1] = (uint8_t)(e.w >> 8);
        o += 2;
        if (me) { int r = ml - 15; for (; r >= 255; r -= 255) *o++ = 255; *o++ = (uint8_t)r; }
    }
    // long literal runs (poorly compressible stretches) are copied by the whole warp, one run after the other,
    // instead of byte by byte by the lane that owns the sequence
    for (uint32_t todo = __ballot_sync(FULL_MASK, mine && lit >= kLongLiterals); todo; todo &= todo - 1) {
        const int l = __ffs(todo) - 1;
        const int from = (int)__shfl_sync(FULL_MASK, e.x, l), n = __shfl_sync(FULL_MASK, lit, l);
        const int at = __shfl_sync(FULL_MASK, incl - size + 1 + le, l);
        warp_copy(out + at, src + from, (uint32_t)n, lane);
    }
    return total;
}

// kDict: the block may reference a read-only dictionary (a4: compress/indie.go:14-35, clz4.go:160-179).
// Positions then live in a virtual space where the dictionary ends at 65536 and the block starts there;
// the table starts as a copy of the dictionary's table instead of empty.
//
// kFrag: the input is one 64 KiB fragment of a larger block (see lz4_compress_frag_kernel).  `prefix` bytes of the
// same block precede src[0] in memory and may be matched (the table is pre-warmed with the last kPrewarm of
// them); the final literal run is NOT emitted — its length goes to *tail_out and the stitcher merges it into the
// next fragment's first sequence.
template <int kHashBits, bool kDict, bool kFrag>
__device__ __forceinline__ int encode_block(const uint8_t* __restrict__ src, int n, uint8_t* dst,
                                            int cap, uint16_t* table, int lane,
                                            const uint8_t* __restrict__ dict, int dsz, const uint16_t* __restrict__ dict_table,
                                            int prefix, uint32_t* tail_out, uint4* queue)
{
    constexpr uint32_t kEmpty = 0xFFFFu;
    constexpr int kProbe = 15;                       // bytes of every candidate examined in parallel
    const int kJumpAt = g_jump_dev;
    const bool kLazy = g_lazy_dev != 0;
    const uint32_t kLazyGain = (uint32_t)g_lazy_dev - 1u;
    const bool kBack = g_back_dev != 0;
    {
        uint4 fill = make_uint4(~0u, ~0u, ~0u, ~0u);
        uint4* t4 = reinterpret_cast<uint4*>(table);
        const uint4* d4 = reinterpret_cast<const uint4*>(dict_table);
        constexpr int kVecs = table_bytes(kHashBits) / 16;
        for (int i = lane; i < kVecs; i += 32) t4[i] = kDict ? d4[i] : fill;
    }
    __syncwarp();
    constexpr int kVirt = kDict ? 65536 : 0;          // virtual position of block byte 0

    int op = 0, anchor = 0, qn = 0;
    if (n >= MFLIMIT + 1) {
        const int mf_end = n - MFLIMIT + 1;          // a match may start at p < mf_end
        const int match_end = n - LASTLITERALS;      // and must end at or before match_end
        const int ld_end = n - 3;                    // 4 bytes can be read at q < ld_end
        // 32-bit addressing relative to aligned bases: byte q of src is byte (d16+q) of src16
        const uintptr_t sa = reinterpret_cast<uintptr_t>(src);
        const uint4* __restrict__ src16 = reinterpret_cast<const uint4*>(sa & ~uintptr_t(15));
        const uint32_t* __restrict__ src4 = reinterpret_cast<const uint32_t*>(sa & ~uintptr_t(15));
        const uint32_t d16 = (uint32_t)sa & 15u;
        const uint32_t end16 = (d16 + (uint32_t)n + 15u) >> 4;       // 16-byte chunks holding valid bytes
        const uint32_t last4 = (d16 + (uint32_t)n - 1u) >> 2;        // last word holding a valid byte
        auto own4 = [&](int q) -> uint32_t {                          // 4 bytes at position q (-prefix <= q < ld_end)
            const int a = (int)d16 + q;
            const uint32_t lo = src4[a >> 2], hi = src4[min((a >> 2) + 1, (int)last4)];
            return __funnelshift_r(lo, hi, (uint32_t)(a & 3) * 8u);
        };
        if (kFrag && prefix > 0) {
            // pre-warm: the last kPrewarm bytes of the previous fragment enter the table (no candidates wanted yet)
            constexpr int kPrewarm = 16384;
            for (int q0 = -min(prefix, kPrewarm); q0 < 0; q0 += 32) {
                const int q = q0 + lane;
                const uint32_t hv = table_slot5(own4(q), own4(q + 1) >> 24, kHashBits);     // kFrag: five-byte hash
                const uint32_t same = __match_any_sync(FULL_MASK, hv);
                if ((same >> lane) == 1u) table[hv] = (uint16_t)q;
                __syncwarp();                        // a later group may overwrite the same slot: keep the order
            }
        }
        const int lo_cand = kFrag ? 1 - min(prefix, 65535) : 1;      // lowest usable candidate position
        const uintptr_t da = reinterpret_cast<uintptr_t>(dict);
        const uint4* __restrict__ dict16 = reinterpret_cast<const uint4*>(da & ~uintptr_t(15));
        const uint32_t dd16 = (uint32_t)da & 15u;
        int base = 0;
        uint32_t v_prv = 0;                          // previous group's bytes (valid whenever literals carry over)
        uint32_t v_cur = (lane < ld_end) ? own4(lane) : 0u;
        uint32_t v_nxt = (lane + 32 < ld_end) ? own4(lane + 32) : 0u;
        while (base < mf_end) {
            const int p = base + lane;
            const bool valid = p < mf_end;
            const uint32_t v = v_cur;
            // bytes two groups ahead go in flight now; they are consumed at the bottom of the loop
            const uint32_t v_far = (p + 64 < ld_end) ? own4(p + 64) : 0u;

            // ---- (a1) hash, table lookup, same-group duplicates, table update
            uint32_t h = 0x80000000u | (uint32_t)lane;
            int cand = -0x40000000;                               // "none": fails the distance test below
            uint32_t b4 = 0;                                      // fifth byte of the position (large blocks hash five)
            if (kFrag) {
                const uint32_t up1 = __shfl_down_sync(FULL_MASK, v_cur, 1), nx0 = __shfl_sync(FULL_MASK, v_nxt, 0);
                b4 = (lane == 31 ? nx0 : up1) >> 24;
            }
            if (valid) {
                h = kFrag ? table_slot5(v, b4, kHashBits) : table_slot(v, kHashBits);
                const uint32_t c = table[h];
                const int vp = p + kVirt;
                int q = (int)(((uint32_t)vp & 0xFFFF0000u) | c);
                if (q >= vp) q -= 65536;
                if (c != kEmpty) cand = q - kVirt;               // < 0: inside the dictionary (kDict only)
            }
            const uint32_t same = __match_any_sync(FULL_MASK, h);
            const uint32_t lower = same & ((1u << lane) - 1u);
            if (lower) cand = base + 31 - __clz(lower);
            if (kFrag) {
                // large blocks enter the even positions only: a table that churns half as fast keeps longer matches
                const uint32_t writers = same & ((base & 1) ? 0xAAAAAAAAu : 0x55555555u);
                if (valid && !(p & 1) && (writers >> lane) == 1u) table[h] = (uint16_t)p;
            } else if (valid && (same >> lane) == 1u) table[h] = (uint16_t)p;   // most recent occurrence wins
            __syncwarp();

            // ---- (a2) every lane measures its own candidate: 1 byte backwards, 15 bytes forwards.
            // positions already covered by the previous match were inserted above but need no candidate;
            // candidate 0 is skipped so that the byte before the candidate always exists.
            // dictionary candidates (cand < 0) must leave a byte before them inside the dictionary
            const bool in_dict = kDict && cand < 0;
            const int didx = cand + dsz;                              // index inside the dictionary when in_dict
            const bool want = valid && p >= anchor && (uint32_t)(p - cand) <= MAX_DISTANCE &&
                              (in_dict ? didx >= 1 : cand >= lo_cand);
            uint4 r0 = make_uint4(0, 0, 0, 0), r1 = make_uint4(0, 0, 0, 0);
            uint32_t t = 0;
            if (want) {
                // fetch starts at the byte before the candidate
                const int ca = in_dict ? ((int)dd16 + didx - 1) : ((int)d16 + cand - 1);   // may be negative in a fragment
                const uint4* __restrict__ cb16 = in_dict ? dict16 : src16;
                const int ci = ca >> 4;
                t = (uint32_t)(ca & 15);
                r0 = cb16[ci];
                if (in_dict || ci + 1 < (int)end16) r1 = cb16[ci + 1];    // the dictionary buffer carries 32 B of zeroed slack
            }
            // own bytes p-1 .. p+14 as four words, from the neighbours' registers
            uint32_t o0, o1, o2, o3;
            {
                const uint32_t up = __shfl_up_sync(FULL_MASK, v_cur, 1);
                o0 = lane ? up : (v_cur << 8);                        // lane 0: no backward byte
                const int s1 = lane + 3, s2 = lane + 7, s3 = lane + 11;
                const uint32_t a1 = __shfl_sync(FULL_MASK, v_cur, s1), b1 = __shfl_sync(FULL_MASK, v_nxt, s1);
                const uint32_t a2 = __shfl_sync(FULL_MASK, v_cur, s2), b2 = __shfl_sync(FULL_MASK, v_nxt, s2);
                const uint32_t a3 = __shfl_sync(FULL_MASK, v_cur, s3), b3 = __shfl_sync(FULL_MASK, v_nxt, s3);
                o1 = (s1 < 32) ? a1 : b1; o2 = (s2 < 32) ? a2 : b2; o3 = (s3 < 32) ? a3 : b3;
            }
            int eqlen;
            bool backeq;
            {
                // barrel-select the 5 words that hold fetched bytes t .. t+19
                const bool w1 = (t & 4u) != 0, w2 = (t & 8u) != 0;
                const uint32_t T0 = w1 ? r0.y : r0.x, T1 = w1 ? r0.z : r0.y, T2 = w1 ? r0.w : r0.z, T3 = w1 ? r1.x : r0.w,
                               T4 = w1 ? r1.y : r1.x, T5 = w1 ? r1.z : r1.y, T6 = w1 ? r1.w : r1.z;
                const uint32_t S0 = w2 ? T2 : T0, S1 = w2 ? T3 : T1, S2 = w2 ? T4 : T2, S3 = w2 ? T5 : T3, S4 = w2 ? T6 : T4;
                const uint32_t bs = (t & 3u) * 8u;
                const uint32_t x0 = __funnelshift_r(S0, S1, bs) ^ o0;
                const uint32_t x1 = __funnelshift_r(S1, S2, bs) ^ o1;
                const uint32_t x2 = __funnelshift_r(S2, S3, bs) ^ o2;
                const uint32_t x3 = __funnelshift_r(S3, S4, bs) ^ o3;
                backeq = (x0 & 0xFFu) == 0;
                // first differing forward byte: pick the first non-zero word, then ctz>>3 (ctz(0)==32 -> 4)
                const uint32_t y0 = x0 >> 8;
                const bool z0 = y0 == 0, z1 = x1 == 0, z2 = x2 == 0;
                const uint32_t xs = z0 ? (z1 ? (z2 ? x3 : x2) : x1) : y0;
                const int skip = z0 ? (z1 ? (z2 ? 11 : 7) : 3) : 0;
                eqlen = skip + (__clz(__brev(xs)) >> 3);
                int room = match_end - p;
                if (in_dict && dsz - didx < room) room = dsz - didx;   // a dictionary match stops at the dictionary's end
                if (eqlen > room) eqlen = room;
            }
            const bool ok = want && eqlen >= MINMATCH;
            const bool more = ok && eqlen == kProbe && p + kProbe < match_end && (!in_dict || didx + kProbe < dsz);
            const bool backok = kBack && ok && lane > 0 && backeq;
            const uint32_t bal = __ballot_sync(FULL_MASK, ok);
            // one-step lazy hint: the next position holds a strictly longer match
            const int nxt_len = __shfl_down_sync(FULL_MASK, eqlen, 1);
            const bool lazy = kLazy && ok && lane < 31 && ((bal >> (lane + 1)) & 1u) && (uint32_t)nxt_len > (uint32_t)eqlen + kLazyGain;
            // one word per lane for the walk: offset | length | flags
            const uint32_t pack = ((uint32_t)(p - cand) << 16) | ((uint32_t)eqlen << 8) | (lazy ? 4u : 0u) | (backok ? 2u : 0u) | (more ? 1u : 0u);

            // ---- (b) walk the measured candidates (greedy, one-step lazy): registers only, except matches longer
            // than kProbe.  Every chosen sequence goes into the shared-memory queue; flush_queue() writes them out.
            {
                int pos = max(anchor - base, 0);                  // first lane not covered yet (may be >= 32)
                uint32_t rem = bal & shl_sat(0xFFFFFFFFu, pos);
                while (rem) {
                    int f = __clz(__brev(rem));
                    uint32_t pk = __shfl_sync(FULL_MASK, pack, f);
                    if (pk & 4u) { f++; pk = __shfl_sync(FULL_MASK, pack, f) & ~2u; }
                    const uint32_t off = pk >> 16;
                    int mlen = (int)((pk >> 8) & 0xFFu);
                    if (pk & 1u) {
                        const int a0 = base + f + kProbe, c0 = a0 - (int)off;      // c0 < 0: candidate bytes are in the dictionary
                        const bool cd = kDict && c0 < 0;
                        const uint8_t* cp = cd ? dict + (c0 + dsz) : src + c0;
                        const int lim = cd ? min(match_end - a0, -c0) : match_end - a0;
                        mlen += count_equal(src + a0, cp, lim, lane);
                    }
                    const int back = (int)((pk >> 1) & 1u) & (int)(f > pos);        // one byte backwards into the literals
                    mlen += back;
                    const int mpos = base + f - back;
                    if (lane == 0) queue[qn] = make_uint4((uint32_t)anchor, (uint32_t)(mpos - anchor), (uint32_t)mlen, off);
                    qn++;
                    anchor = mpos + mlen;
                    pos = anchor - base;
                    rem &= shl_sat(0xFFFFFFFFu, pos);
                }
                if (qn >= 32) {
                    __syncwarp();
                    const int w = flush_queue(queue, 32, src, dst + op, cap - op, lane);
                    if (w < 0) return 0;
                    op += w;
                    const uint4 keep = queue[32 + (lane & (kQueueLen - 32 - 1))];    // slide the remainder (< 16) down
                    __syncwarp();
                    if (lane < kQueueLen - 32) queue[lane] = keep;
                    qn -= 32;
                    __syncwarp();
                }
            }
            if (anchor >= base + kJumpAt) {
                base = anchor;          // a long match skipped whole groups: their positions are not inserted
                v_cur = (base + lane < ld_end) ? own4(base + lane) : 0u;
                v_nxt = (base + lane + 32 < ld_end) ? own4(base + lane + 32) : 0u;
            } else {
                base += 32;
                v_prv = v_cur;
                v_cur = v_nxt;
                v_nxt = v_far;
            }
        }
    }
    if (qn > 0) {
        __syncwarp();
        const int w = flush_queue(queue, qn, src, dst + op, cap - op, lane);
        if (w < 0) return 0;
        op += w;
    }
    if (kFrag) {
        *tail_out = (uint32_t)(n - anchor);          // the stitcher owns the final literal run
        return op;
    }
    // last literals (lz4.c:1302-1329)
    {
        const int run = n - anchor;
        const int need = 1 + (run >= 15 ? ext_bytes(run - 15) : 0) + run;
        if (op + need > cap) return 0;
        uint8_t* o = dst + op;
        if (lane == 0) o[0] = (uint8_t)((run < 15 ? run : 15) << 4);
        int w = 1;
        if (run >= 15) { put_ext(o + w, run - 15, lane); w += ext_bytes(run - 15); }
        warp_copy(o + w, src + anchor, (uint32_t)run, lane);
        op += w + run;
    }
    return op;
}

// blk.CompressToBlk's framing around an encoded payload of c bytes (c == 0: it did not fit): stored fallback,
// size word, xxh32 trailer, record length (blk/blk.go:78-106); raw block API: just the length (clz4.go:40-42).
__device__ __forceinline__ void finish_record(const EncodeArgs& a, uint32_t b, const uint8_t* src, int n,
                                              uint8_t* rec, uint8_t* payload, int c, int lane)
{
    if (a.raw_blocks) {
        if (lane == 0) a.rec_len[b] = (uint32_t)c;      // 0 = does not fit
        return;
    }
    uint32_t word;
    if (c == 0) {                                       // blk/blk.go:78-92: store raw
        warp_copy(payload, src, (uint32_t)n, lane);
        c = n;
        word = (uint32_t)n | 0x80000000u;
    } else {
        word = (uint32_t)c;
    }
    if (lane == 0) store_le32(rec, word);
    uint32_t total = 4u + (uint32_t)c;
    if (a.block_checksum) {                             // blk/blk.go:98-102
        __syncwarp();
        uint32_t x = warp_xxh32(payload, (uint32_t)c, lane);
        if (lane == 0) store_le32(payload + c, x);
        total += 4;
    }
    if (lane == 0) a.rec_len[b] = total;
}

template <int kHashBits, bool kDict>
__global__ void __launch_bounds__(kEncodeWarps * 32, 7)
lz4_compress_kernel(EncodeArgs a)
{
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int kTableBytes = table_bytes(kHashBits);
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    const uint32_t b = blockIdx.x * kEncodeWarps + warp;
    if (b >= a.nblk) return;

    const uint8_t* src = a.src_base + a.src_off[b];
    const int n = (int)a.src_len[b];
    if (a.split_by_size && n > 65536) return;                     // the span kernel's block
    uint8_t* rec = a.rec_base + (uint64_t)b * a.rec_stride;
    uint8_t* payload = a.raw_blocks ? rec : rec + 4;
    uint16_t* table = reinterpret_cast<uint16_t*>(smem + warp * kTableBytes);

    __shared__ uint4 s_queue[kEncodeWarps][kQueueLen];
    int c = encode_block<kHashBits, kDict, false>(src, n, payload, (int)a.dst_cap, table, lane, a.dict, (int)a.dict_size,
                                                  a.dict_table, 0, nullptr, s_queue[warp]);
    finish_record(a, b, src, n, rec, payload, c, lane);
}

// ---------------------------------------------------------------- blocks larger than 64 KiB
//
// One warp per block would leave a 4 MiB-block frame (plz4's default) with 64 blocks of parallelism per 256 MiB.
// The encoder is free to choose its parse, so a large block is cut into 64 KiB FRAGMENTS, each encoded by its own
// warp (matches may reach back into the previous fragment: same block, same window rules), and a stitch pass joins
// the fragment streams into one LZ4 block: the literals left over at the end of fragment k become part of the first
// sequence of fragment k+1, so only that sequence's token / length bytes are rewritten; everything else is copied.
constexpr int kFragBytes = 131072;             // half the boundaries of 64 KiB fragments (a long match ends at every one)
constexpr int kHashTileWords = 4096;            // 16 KiB tiles of payload staged in shared memory for the block checksum

struct FragArgs {
    EncodeArgs e;
    uint8_t* tmp;              // [nblk * frags_per_block] slots of frag_stride bytes
    uint32_t frag_stride;
    uint32_t frags_per_block;
    uint32_t frag_bytes;       // input bytes per fragment
    int32_t* frag_len;         // encoded bytes before the final literal run; -2 = fragment does not exist
    uint32_t* frag_tail;       // literal bytes left over at the fragment's end
};

template <int kHashBits>
__global__ void __launch_bounds__(kEncodeWarps * 32, 7)
lz4_compress_frag_kernel(FragArgs a)
{
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int kTableBytes = table_bytes(kHashBits);
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    const uint32_t w = blockIdx.x * kEncodeWarps + warp;
    const uint32_t b = w / a.frags_per_block, f = w % a.frags_per_block;
    if (b >= a.e.nblk) return;
    const int n_blk = (int)a.e.src_len[b];
    const int start = (int)f * kFragBytes;
    if (start >= n_blk && f != 0) {
        if (lane == 0) a.frag_len[w] = -2;
        return;
    }
    const int n = min(kFragBytes, n_blk - start);
    const uint8_t* src = a.e.src_base + a.e.src_off[b] + start;
    uint16_t* table = reinterpret_cast<uint16_t*>(smem + warp * kTableBytes);
    __shared__ uint4 s_queue[kEncodeWarps][kQueueLen];
    uint32_t tail = 0;
    int c = encode_block<kHashBits, false, true>(src, n, a.tmp + (uint64_t)w * a.frag_stride, (int)a.frag_stride, table, lane,
                                                 nullptr, 0, nullptr, start, &tail, s_queue[warp]);
    if (lane == 0) { a.frag_len[w] = c; a.frag_tail[w] = tail; }
}

constexpr int kStitchWarps = 8;
constexpr int kMaxFrags = 66;                     // 4 MiB / 64 KiB (+ CompressBound slack on the raw block path)

// One CTA per block.  Warp 0 plans (a short serial walk over <= 64 fragments: new first-token header and output
// offset of each), all warps copy fragments in parallel, warp 0 writes the last literals and frames the record.
__global__ void __launch_bounds__(kStitchWarps * 32)
lz4_stitch_kernel(FragArgs a)
{
    __shared__ int s_out[kMaxFrags];              // output offset of the fragment's (rewritten) first token; -1 = skip
    __shared__ int s_carry[kMaxFrags];            // literals carried into the fragment
    __shared__ int s_pos[kMaxFrags];              // source position of the fragment
    __shared__ int s_end, s_tail, s_ok, s_c;
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    const uint32_t b = blockIdx.x;
    const uint8_t* src = a.e.src_base + a.e.src_off[b];
    const int n_blk = (int)a.e.src_len[b];
    if (a.e.split_by_size && n_blk <= 65536) return;          // encoded whole by the CTA kernel
    uint8_t* rec = a.e.rec_base + (uint64_t)b * a.e.rec_stride;
    uint8_t* out = a.e.raw_blocks ? rec : rec + 4;
    const int cap = (int)a.e.dst_cap;
    const int F = (int)a.frags_per_block;

    if (warp == 0) {
        int op = 0, carry = 0, pos = 0;
        bool ok = true;
        for (int f = 0; f < F; f++) {
            const uint32_t w = b * a.frags_per_block + (uint32_t)f;
            const int flen = a.frag_len[w];
            if (flen == -2) { for (int g = f + lane; g < F; g += 32) s_out[g] = -1; break; }
            const int nf = min((int)a.frag_bytes, n_blk - pos);
            if (flen <= 0) {                                            // no sequence in this fragment: all literals
                if (lane == 0) s_out[f] = -1;
                carry += nf; pos += nf;
                continue;
            }
            const uint8_t* frag = a.tmp + (uint64_t)w * a.frag_stride;
            const uint32_t tok = frag[0];
            int l0 = (int)(tok >> 4), hdr0 = 1;
            if (l0 == 15) { uint32_t s; do { s = frag[hdr0++]; l0 += (int)s; } while (s == 255); }
            const int L = carry + l0;
            const int need = 1 + (L >= 15 ? ext_bytes(L - 15) : 0) + carry + (flen - hdr0);
            if (op + need > cap) { ok = false; break; }
            if (lane == 0) { s_out[f] = op; s_carry[f] = carry; s_pos[f] = pos; }
            op += need;
            carry = (int)a.frag_tail[w];
            pos += nf;
        }
        if (pos != n_blk) ok = false;                      // fewer fragments than the block needs (never, by launch_compress)
        if (ok && op + 1 + (carry >= 15 ? ext_bytes(carry - 15) : 0) + carry > cap) ok = false;
        if (lane == 0) { s_end = op; s_tail = carry; s_ok = ok ? 1 : 0; }
    }
    __syncthreads();
    if (s_ok) {
        for (int f = warp; f < F; f += kStitchWarps) {
            int op = s_out[f];
            if (op < 0) continue;
            const uint32_t w = b * a.frags_per_block + (uint32_t)f;
            const uint8_t* frag = a.tmp + (uint64_t)w * a.frag_stride;
            const int flen = a.frag_len[w], carry = s_carry[f];
            const uint32_t tok = frag[0];
            int l0 = (int)(tok >> 4), hdr0 = 1;
            if (l0 == 15) { uint32_t s; do { s = frag[hdr0++]; l0 += (int)s; } while (s == 255); }
            const int L = carry + l0;
            // the fragment's first sequence: its literal run grows by the literals carried over
            if (lane == 0) out[op] = (uint8_t)(((L < 15 ? L : 15) << 4) | (tok & 15u));
            op += 1;
            if (L >= 15) { put_ext(out + op, L - 15, lane); op += ext_bytes(L - 15); }
            warp_copy(out + op, src + s_pos[f] - carry, (uint32_t)carry, lane);
            op += carry;
            warp_copy(out + op, frag + hdr0, (uint32_t)(flen - hdr0), lane);   // own literals, offset, length bytes, later sequences
        }
    }
    __syncthreads();
    if (warp == 0) {
    int c = 0;
    if (s_ok) {
        // last literals of the block (lz4.c:1302-1329)
        int op = s_end;
        const int carry = s_tail;
        if (lane == 0) out[op] = (uint8_t)((carry < 15 ? carry : 15) << 4);
        op += 1;
        if (carry >= 15) { put_ext(out + op, carry - 15, lane); op += ext_bytes(carry - 15); }
        warp_copy(out + op, src + n_blk - carry, (uint32_t)carry, lane);
        c = op + carry;
    }
    // ---- record framing (blk/blk.go:78-106) by the whole CTA: a block of this size is too much for one warp
    if (lane == 0) s_c = c;
    }
    __syncthreads();
    int c = s_c;
    if (a.e.raw_blocks) {
        if (threadIdx.x == 0) a.e.rec_len[b] = (uint32_t)c;              // 0 = does not fit
        return;
    }
    uint32_t word = (uint32_t)c;
    if (c == 0) {                                                        // blk/blk.go:78-92: store raw
        const int per = ((n_blk + kStitchWarps - 1) / kStitchWarps + 15) & ~15;
        const int lo = min(n_blk, warp * per), hi = min(n_blk, lo + per);
        warp_copy(out + lo, src + lo, (uint32_t)(hi - lo), lane);
        c = n_blk;
        word = (uint32_t)n_blk | 0x80000000u;
        __syncthreads();
    }
    if (threadIdx.x == 0) store_le32(rec, word);
    uint32_t total = 4u + (uint32_t)c;
    if (a.e.block_checksum) {                                            // blk/blk.go:98-102
        // xxh32 is serial along the payload, so one warp carries the chain; the other warps keep it fed: they copy
        // the payload tile by tile into shared memory (two tiles, one being filled while the other is hashed), which
        // turns every load the chain waits for from a trip to L2 / HBM into a shared-memory access
        __shared__ uint32_t s_tile[2][kHashTileWords];
        const uint32_t* pw = reinterpret_cast<const uint32_t*>(out);     // payload is 4-byte aligned (slot base + 4)
        const uint32_t nstripes = (uint32_t)c >> 4;
        const uint32_t ntiles = (nstripes * 4 + kHashTileWords - 1) / kHashTileWords;
        auto fill = [&](uint32_t t) {
            const uint32_t w0 = t * kHashTileWords, w1 = min(nstripes * 4, w0 + kHashTileWords);
            for (uint32_t i = w0 + (threadIdx.x - 32); i < w1; i += (kStitchWarps - 1) * 32) s_tile[t & 1][i - w0] = pw[i];
        };
        uint32_t acc = xxh32_init(lane);
        if (warp != 0 && ntiles) fill(0);
        __syncthreads();
        for (uint32_t t = 0; t < ntiles; t++) {
            if (warp != 0) { if (t + 1 < ntiles) fill(t + 1); }
            else acc = xxh32_consume_words(acc, s_tile[t & 1], min((uint32_t)kHashTileWords / 4, nstripes - t * (kHashTileWords / 4)), lane);
            __syncthreads();
        }
        if (warp == 0) {
            const uint32_t x = xxh32_finish(acc, out, (uint32_t)c);
            if (lane == 0) store_le32(out + c, x);
        }
        total += 4;
    }
    if (threadIdx.x == 0) a.e.rec_len[b] = total;
}

// Dictionary table: for every hash the LAST dictionary position holding it (what a block would have seen
// had the dictionary been its own first bytes), as the low 16 bits of the virtual position.
__global__ void __launch_bounds__(1024) dict_table_kernel(const uint8_t* __restrict__ dict, int dsz, int bits, uint16_t* __restrict__ table)
{
    extern __shared__ int best[];                      // highest position per hash, -1 = none
    const int entries = 1 << bits;
    for (int i = threadIdx.x; i < entries; i += blockDim.x) best[i] = -1;
    __syncthreads();
    for (int j = threadIdx.x; j + 4 <= dsz; j += blockDim.x) {
        const uint32_t v = (uint32_t)dict[j] | ((uint32_t)dict[j + 1] << 8) | ((uint32_t)dict[j + 2] << 16) | ((uint32_t)dict[j + 3] << 24);
        atomicMax(&best[table_slot(v, bits)], j);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < entries; i += blockDim.x)
        table[i] = best[i] < 0 ? (uint16_t)0xFFFF : (uint16_t)(65536 - dsz + best[i]);
}

cudaError_t launch_dict_build(const uint8_t* dict, uint32_t dict_size, int bits, uint16_t* table, cudaStream_t stream)
{
    dict_table_kernel<<<1, 1024, sizeof(int) << bits, stream>>>(dict, (int)dict_size, bits, table);
    return cudaGetLastError();
}

static int g_span_want = 0;       // fragments per span, 0 = sixteen; PLZ4CU_SPAN_WANT (experiments)
static int g_spans = 1;           // blocks above 64 KiB: spans of fragments on the CTA encoder (0: one warp per 128 KiB fragment); PLZ4CU_SPANS
static int g_cta_min = 1;         // 0: the one-warp-per-block kernels of round 1 for everything (PLZ4CU_CTA_MIN=0, measurements)
static int g_hash_bits = 12;     // 7 KiB of table per warp: 28 resident warps per SM against 12 with liblz4's 13 bits;
                                 // the one-step lazy parse more than pays the ratio back (profiles/r01_sweep.txt)

template <int kBits>
static cudaError_t set_smem_attr()
{
    cudaError_t e = cudaFuncSetAttribute(lz4_compress_kernel<kBits, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kEncodeWarps * table_bytes(kBits));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(lz4_compress_kernel<kBits, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             kEncodeWarps * table_bytes(kBits));
    if (e != cudaSuccess || kBits == 13) return e;
    return cudaFuncSetAttribute(lz4_compress_frag_kernel<(kBits == 13 ? 12 : kBits)>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                kEncodeWarps * table_bytes(kBits));
}

int compress_hash_bits(uint32_t dst_cap)
{
    // blocks above 1 MiB come in small numbers (64 per 256 MiB at 4 MiB): occupancy is not the limit there,
    // so they get liblz4's 8192-entry table back
    return (dst_cap > (1u << 20) + (1u << 20) / 255u + 16u) ? 13 : g_hash_bits;
}

cudaError_t configure_compress()
{
    // tuning knobs for experiments; the defaults are the product configuration
    if (const char* e = getenv("PLZ4CU_HASH_BITS")) {
        int v = atoi(e);
        if (v >= 11 && v <= 13) g_hash_bits = v;
    }
    if (const char* e = getenv("PLZ4CU_CTA_MIN")) g_cta_min = atoi(e);
    if (const char* e = getenv("PLZ4CU_SPANS")) g_spans = atoi(e);
    if (const char* e = getenv("PLZ4CU_SPAN_WANT")) g_span_want = atoi(e);
    if (const char* e = getenv("PLZ4CU_BACK")) {
        int v = atoi(e);
        cudaError_t err = cudaMemcpyToSymbol(g_back_dev, &v, sizeof v);
        if (err != cudaSuccess) return err;
    }
    if (const char* e = getenv("PLZ4CU_JUMP")) {
        int v = atoi(e);
        cudaError_t err = cudaMemcpyToSymbol(g_jump_dev, &v, sizeof v);
        if (err != cudaSuccess) return err;
    }
    if (const char* e = getenv("PLZ4CU_LAZY")) {
        int v = atoi(e);
        cudaError_t err = cudaMemcpyToSymbol(g_lazy_dev, &v, sizeof v);
        if (err != cudaSuccess) return err;
    }
    {
        // fragment scratch comes from the stream-ordered allocator: keep freed memory cached in the pool instead
        // of handing it back to the driver at every synchronisation
        int dev = 0;
        cudaMemPool_t pool;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    cudaError_t err = set_smem_attr<11>();
    if (err == cudaSuccess) err = set_smem_attr<12>();
    if (err == cudaSuccess) err = set_smem_attr<13>();
    return err;
}

template <int kBits>
static cudaError_t launch_frag(const FragArgs& fa, uint32_t nwarps, cudaStream_t stream)
{
    lz4_compress_frag_kernel<kBits><<<(nwarps + kEncodeWarps - 1) / kEncodeWarps, kEncodeWarps * 32, kEncodeWarps * table_bytes(kBits), stream>>>(fa);
    return cudaGetLastError();
}

// one warp per block (round 1's encoder): blocks with a dictionary up to 64 KiB, and everything under PLZ4CU_CTA_MIN=0
static cudaError_t launch_compress_warps(const EncodeArgs& a, cudaStream_t stream)
{
    dim3 grid((a.nblk + kEncodeWarps - 1) / kEncodeWarps), block(kEncodeWarps * 32);
    const int bits = compress_hash_bits(a.dst_cap);
    const size_t sm = (size_t)kEncodeWarps * table_bytes(bits);
    const bool d = a.dict_size > 0;
    if (bits == 11)      { if (d) lz4_compress_kernel<11, true><<<grid, block, sm, stream>>>(a); else lz4_compress_kernel<11, false><<<grid, block, sm, stream>>>(a); }
    else if (bits == 12) { if (d) lz4_compress_kernel<12, true><<<grid, block, sm, stream>>>(a); else lz4_compress_kernel<12, false><<<grid, block, sm, stream>>>(a); }
    else                 { if (d) lz4_compress_kernel<13, true><<<grid, block, sm, stream>>>(a); else lz4_compress_kernel<13, false><<<grid, block, sm, stream>>>(a); }
    return cudaGetLastError();
}

cudaError_t launch_compress(const EncodeArgs& a, cudaStream_t stream)
{
    if (a.nblk == 0) return cudaSuccess;
    const uint32_t max_len = a.max_src_len ? a.max_src_len : a.dst_cap;
    // No dictionary: a block of up to 64 KiB is the CTA kernel's, a larger one the span kernel's, whatever else the
    // launch holds — so a block compresses to the same bytes in any batch, stream or device (tests/test_gpu_multi.py).
    if (a.dict_size == 0 && max_len <= 65536u && g_cta_min > 0) return launch_compress_cta(a, stream);
    if (max_len > 65536u && g_spans) {
        // large blocks: spans of sixteen 64 KiB fragments, one CTA each (compress_cta.cu), then the stitch.  The cut does
        // not depend on how many blocks the launch has: a block compresses to the same bytes whatever batch, stream or
        // device it travels in (a 4 MiB block is four spans, so 64 of them — a 256 MiB file — already give every SM a CTA)
        const uint32_t nfrag = (max_len + 65535u) / 65536u;
        uint32_t span_frags = g_span_want > 0 ? (uint32_t)g_span_want : 16u;
        while ((nfrag + span_frags - 1) / span_frags > (uint32_t)kMaxFrags) span_frags *= 2;
        FragArgs fa{};
        fa.e = a;
        fa.e.split_by_size = 1;
        if (a.min_src_len <= 65536u) {                       // some block may be small (or nobody knows): those first
            cudaError_t e0 = a.dict_size ? launch_compress_warps(fa.e, stream) : launch_compress_cta(fa.e, stream);
            if (e0 != cudaSuccess) return e0;
        }
        fa.frags_per_block = (nfrag + span_frags - 1) / span_frags;
        fa.frag_bytes = span_frags * 65536u;
        fa.frag_stride = (uint32_t)(((uint64_t)fa.frag_bytes + fa.frag_bytes / 255 + 16 + 15) & ~15ull);
        const uint64_t nspan = (uint64_t)a.nblk * fa.frags_per_block;
        void* scratch = nullptr;
        cudaError_t e = cudaMallocAsync(&scratch, nspan * fa.frag_stride + nspan * 8, stream);
        if (e != cudaSuccess) return e;
        fa.tmp = static_cast<uint8_t*>(scratch);
        fa.frag_len = reinterpret_cast<int32_t*>(fa.tmp + nspan * fa.frag_stride);
        fa.frag_tail = reinterpret_cast<uint32_t*>(fa.frag_len + nspan);
        e = launch_compress_spans(fa.e, fa.tmp, fa.frag_stride, fa.frags_per_block, fa.frag_bytes, fa.frag_len, fa.frag_tail, stream);
        if (e == cudaSuccess) {
            lz4_stitch_kernel<<<a.nblk, kStitchWarps * 32, 0, stream>>>(fa);
            e = cudaGetLastError();
        }
        cudaError_t e2 = cudaFreeAsync(scratch, stream);
        return e != cudaSuccess ? e : e2;
    }
    const uint32_t frags = (max_len + kFragBytes - 1) / kFragBytes;
    if (max_len > 65536u && a.dict_size == 0 && frags <= (uint32_t)kMaxFrags) {
        // (PLZ4CU_SPANS=0) fragment-parallel encode, one warp per 128 KiB fragment, + stitch
        FragArgs fa{};
        fa.e = a;
        fa.frags_per_block = frags;
        fa.frag_bytes = kFragBytes;
        fa.frag_stride = (uint32_t)((kFragBytes + kFragBytes / 255 + 16 + 15) & ~15);
        const uint64_t nfrag = (uint64_t)a.nblk * fa.frags_per_block;
        void* scratch = nullptr;
        const uint64_t bytes = nfrag * fa.frag_stride + nfrag * 8;
        cudaError_t e = cudaMallocAsync(&scratch, bytes, stream);
        if (e != cudaSuccess) return e;
        fa.tmp = static_cast<uint8_t*>(scratch);
        fa.frag_len = reinterpret_cast<int32_t*>(fa.tmp + nfrag * fa.frag_stride);
        fa.frag_tail = reinterpret_cast<uint32_t*>(fa.frag_len + nfrag);
        const int bits = g_hash_bits == 11 ? 11 : 12;
        e = (bits == 11) ? launch_frag<11>(fa, (uint32_t)nfrag, stream) : launch_frag<12>(fa, (uint32_t)nfrag, stream);
        if (e == cudaSuccess) {
            lz4_stitch_kernel<<<a.nblk, kStitchWarps * 32, 0, stream>>>(fa);
            e = cudaGetLastError();
        }
        cudaError_t e2 = cudaFreeAsync(scratch, stream);
        return e != cudaSuccess ? e : e2;
    }
    return launch_compress_warps(a, stream);
}

}  // namespace plz4
