// decompress.cu — LZ4 block decode (sm_100a): one warp per block when a launch fills the SMs that way, two warps per
// block ("duo": a parser and a copier) when it has fewer than 32 blocks per SM, one CTA per block ("team", second half
// of this file) when it has few, large ones or only a handful.
//
// Replaces, for a whole batch of independent blocks, what plz4 does per block on a goroutine:
//   blk/frame.go:79-81      size word > block size            -> PLZ4CU_E_OVERFLOW
//   blk/frame.go:114-127    xxh32 over the payload            -> PLZ4CU_E_BLOCKHASH
//   async/reader.go:149-164 stored block: straight copy
//   compress/decompress.go:32-38,46-58 -> clz4.go:47-78 -> lz4.c:2023-2445 LZ4_decompress_generic
//
// Accept/reject behaviour and the negative return code follow liblz4's decode_full_block state
// machine exactly (derivation: DESIGN.md "decoder state machine").
//
// How the warp works.  A sequence-at-a-time decoder leaves most lanes idle (an average sequence on
// log text is ~14 output bytes) and pays the whole parse per sequence.  So the warp decodes in
// BATCHES of up to 32 sequences:
//   1. parse: the token chain is walked once (warp-uniform, broadcast byte loads); lane k keeps the
//      fields of sequence k.  Only "shortcut" sequences (lz4.c:2100-2108: literal nibble < 15, >= 17
//      input bytes left) enter a batch; anything else ends the batch and is handled by the
//      sequence-at-a-time path below, which is the literal statement of the state machine.
//   2. scan: output positions of all sequences by a warp prefix sum; the capacity / offset checks
//      that depend on the output position are evaluated for all 32 sequences at once.
//   3. copy: the batch's output range is produced 32 bytes at a time, one byte per lane.  A lane finds
//      the sequence owning its byte (a shared-memory bitmap of sequence starts + popc), computes where the byte comes
//      from — the compressed stream (literal), the dictionary, or earlier output (match) — and only
//      bytes whose source lies inside the same 32-byte chunk need lane-to-lane forwarding (pointer
//      doubling over shuffles).  All loads of a chunk are independent, so match sources that miss
//      L1/L2 overlap instead of serialising.
#include "common.cuh"
#include "kernels.h"
#include <cstdlib>

namespace plz4 {

// ---------------------------------------------------------------- one sequence at a time (state machine, literal form)

enum : int { kStepMore = 0, kStepDone = 1 };

// Decodes exactly one sequence starting at ip/op.  Returns kStepMore (ip/op advanced) or kStepDone (ret set).
template <bool kDict>
__device__ __noinline__ int decode_one(const uint8_t* __restrict__ src, int n, uint8_t* dst, int cap,
                                       const uint8_t* __restrict__ dict, int dsz, int lane, int& ip, int& op, int32_t& ret)
{
    const bool check_offset = dsz < 65536;
    const uint32_t tok = src[ip++];
    int len = (int)(tok >> 4);
    int mlen;
    uint32_t off;

    if (len != 15 && ip < n - 16 && op <= cap - 32) {
        // "shortcut" sequence: cannot be the last one, no end-of-block checks (lz4.c:2100-2108,2250-2256)
        if (lane < len) dst[op + lane] = src[ip + lane];
        op += len; ip += len;
        mlen = (int)(tok & 15);
        off = load_u16le(src + ip); ip += 2;
        if (mlen != 15 && off >= 8 && (int)off <= op) {
            mlen += MINMATCH;                      // <= 18 bytes: one round
            __syncwarp();
            if (lane < mlen) {
                int k = ((int)off >= mlen) ? lane : (lane % (int)off);
                dst[op + lane] = dst[op - (int)off + k];
            }
            op += mlen;
            return kStepMore;
        }
    } else {
        if (len == 15) {
            // read_variable_length(ip, iend-15, initial_check) lz4.c:1978-2014
            if (ip >= n - 15) { ret = -ip - 1; return kStepDone; }
            uint32_t s;
            do {
                s = src[ip++];
                len += (int)s;
                if (ip > n - 15) { ret = -ip - 1; return kStepDone; }
            } while (s == 255);
        }
        if (op + len > cap - MFLIMIT || ip + len > n - (2 + 1 + LASTLITERALS)) {
            // must be the last sequence: consume the input exactly, fit the output (lz4.c:2297-2330)
            if (ip + len != n || op + len > cap) { ret = -ip - 1; return kStepDone; }
            warp_copy(dst + op, src + ip, (uint32_t)len, lane);
            ret = op + len;
            return kStepDone;
        }
        warp_copy(dst + op, src + ip, (uint32_t)len, lane);
        op += len; ip += len;
        off = load_u16le(src + ip); ip += 2;
        mlen = (int)(tok & 15);
    }

    // general match (lz4.c:2342-2430)
    if (mlen == 15) {
        uint32_t s;
        do {
            s = src[ip++];
            mlen += (int)s;
            if (ip > n - LASTLITERALS + 1) { ret = -ip - 1; return kStepDone; }
        } while (s == 255);
    }
    mlen += MINMATCH;
    if (check_offset && op - (int)off + dsz < 0) { ret = -ip - 1; return kStepDone; }
    if (off == 0) { ret = -ip - 1; return kStepDone; }        // stated divergence: liblz4 would replay garbage
    if (op + mlen > cap - LASTLITERALS) { ret = -ip - 1; return kStepDone; }

    __syncwarp();
    {
        // virtual history = dict ++ out; byte k comes from v = op - off + (k mod off) (v < 0: dictionary)
        const int vbase = op - (int)off;
        if ((int)off >= mlen) {
            if (!kDict || vbase >= 0) {
                const uint8_t* s = dst + vbase;
                for (int k = lane; k < mlen; k += 32) dst[op + k] = s[k];
            } else {
                for (int k = lane; k < mlen; k += 32) {
                    int v = vbase + k;
                    dst[op + k] = (v < 0) ? dict[dsz + v] : dst[v];
                }
            }
        } else {
            // overlapping: periodic with period off
            int r = lane % (int)off;
            const int step = 32 % (int)off;
            for (int k = lane; k < mlen; k += 32) {
                int v = vbase + r;
                uint8_t b;
                if (kDict && v < 0) b = dict[dsz + v]; else b = dst[v];
                dst[op + k] = b;
                r += step; if (r >= (int)off) r -= (int)off;
            }
        }
    }
    op += mlen;
    __syncwarp();
    return kStepMore;
}

// ---------------------------------------------------------------- batched decode

// a byte that this kernel may have written itself (match source) or not (literal): plain global load, never the read-only path
__device__ __forceinline__ uint32_t lds_u8(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// (predicated in place: a branch around one load costs more than the load)
__device__ __forceinline__ uint32_t ld_global_u8_if(const uint8_t* p, bool need)
{
    uint32_t v = 0;
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q ld.global.u8 %0, [%1];\n\t}" : "+r"(v) : "l"(p), "r"((uint32_t)need) : "memory");
    return v;
}

// kRing: the last 64 KiB of output are mirrored in a shared-memory ring and match sources are read from there.
// Used when a launch has too few blocks to hide global-memory latency with other warps (large block sizes):
// the per-chunk round trip drops from an L2/HBM access to a shared-memory access.
// Block checksum carried along with the decode: the payload is hashed 512 bytes at a time just ahead of the parse, so a
// record comes in from DRAM once (the hash leaves its lines in L1 for the parse) instead of once per pass.
struct HashAlong {
    bool on;
    uint32_t acc;
    int done;                  // payload bytes in acc (a multiple of 512)
};

// hash the whole 512-byte chunks below `upto` (kept out of line: the decode loop has no registers to lend)
static __device__ __noinline__ void hash_along(HashAlong& H, const uint8_t* payload, int upto, int lane)
{
    const int nch = ((upto & ~511) - H.done) >> 9;
    if (nch > 0) {
        const uintptr_t pa = reinterpret_cast<uintptr_t>(payload);
        const uint32_t* wp = reinterpret_cast<const uint32_t*>(pa & ~uintptr_t(3));     // word-aligned payload and its shift
        H.acc = xxh32_consume_global<false>(H.acc, wp + (H.done >> 2), (uint32_t)(pa & 3u) * 8u, nch, lane);
        H.done += nch << 9;
    }
}

// ---- two warps per block ("duo"): one parses (walk, headers, scan, checks) and publishes batches, the other produces the
// bytes.  Both contexts hide latency, but they share ONE block: half as many 64 KiB output windows are in flight per SM as with
// a block per warp at the same occupancy, and neither warp carries the other's registers.
struct DuoSlot {
    uint32_t packA[32];         // per sequence: where its match starts (relative to out0) | offset << 16
    int litrel[32];             // per sequence: literal for the byte at batch position x is src[litrel + x]
    uint32_t bits[32];          // bitmap of sequence starts over the batch's output range
};
#ifndef PLZ4CU_DUO_RING
#define PLZ4CU_DUO_RING 2
#endif
constexpr int kDuoRing = PLZ4CU_DUO_RING;         // batches between parser and copier (a power of two)
struct DuoShared {
    DuoSlot slot[kDuoRing];
    int out0[kDuoRing], total[kDuoRing];
    volatile int published;     // batches the parser has listed
    volatile int taken;         // batches the copier has read into registers (their slots are free)
    volatile int copied;        // batches whose bytes are all written
    volatile int quit;          // the parser lists nothing more
};
constexpr int kDuoSpinLimit = 1 << 24;
#ifndef PLZ4CU_DUO_NAP
#define PLZ4CU_DUO_NAP 2000
#endif

// wait until *counter >= want (a nap between polls: a waiting warp takes no issue slots from the working ones)
__device__ __forceinline__ void duo_wait(const volatile int* counter, int want, int lane)
{
    for (int spins = 0;; spins++) {
        int v = 0;
        if (lane == 0) v = *counter;
        v = __shfl_sync(FULL_MASK, v, 0);
        if (v >= want) break;
        if (spins > kDuoSpinLimit) __trap();
        __nanosleep(PLZ4CU_DUO_NAP);
    }
    __threadfence_block();
}

template <bool kDict, bool kRing, bool kDuo = false>
__device__ __forceinline__ int32_t decode_block(const uint8_t* __restrict__ src, int n,
                                                uint8_t* dst, int cap,
                                                const uint8_t* __restrict__ dict, int dsz, int lane, uint32_t* bitmap, uint32_t* win,
                                                uint8_t* ring, HashAlong& H, DuoShared* D = nullptr)
{
    int nb = 0;                                     // kDuo: batches published
    constexpr int kBatchBytes = 1024;               // output bytes one batch may span (32 bitmap words)
    constexpr int kMaxBatchLit = 63;                // longest literal run a batched sequence may carry (6 bits)
    if (cap == 0) return (n == 1 && src[0] == 0) ? 0 : -1;
    if (n == 0) return -1;

    int ip = 0, op = 0;
    const bool check_offset = dsz < 65536;
    // word-aligned view of the compressed stream: byte q of src is byte (d4 + q) of src4
    const uint32_t d4 = (uint32_t)reinterpret_cast<uintptr_t>(src) & 3u;
    const uint32_t* __restrict__ src4 = reinterpret_cast<const uint32_t*>(src - d4);
    const uint32_t last4 = (d4 + (uint32_t)n - 1u) >> 2;         // last word holding a valid byte

    for (;;) {
        if (H.on && ip + 256 > H.done && H.done + 512 <= n) hash_along(H, src, min(ip + 1024, n), lane);
        // ---- 1. parse up to 32 shortcut sequences; lane k latches sequence k.
        // The next 128 compressed bytes are loaded one word per lane.  Every lane first computes, for each of its
        // own 4 bytes, how long a sequence header starting there would be (token + literals + offset [+ one
        // length byte]); lengths and window go to shared memory, and walking the token chain is then three
        // instructions per sequence instead of three dependent global loads.  Headers that do not fit the simple
        // shape (more than one length byte, a literal run above 63, too close to the end of the input) end the
        // batch and go through decode_one.
        int nseq = 0;
        int my_lit = 0, my_litpos = 0, my_mlen = 0;
        // input position after my sequence: literals, offset, and the one extension byte a match nibble of 15 carries
        auto my_ipn = [&]() { return my_litpos + my_lit + 2 + (my_mlen >= MINMATCH + 15 ? 1 : 0); };
        uint32_t my_off = 0;
        bool my_simple = false;                     // second shortcut stage applies on the input side (nibble != 15, offset >= 8)
        {
            const uint32_t a0 = d4 + (uint32_t)ip;              // byte address of ip relative to src4
            const uint32_t w0 = a0 >> 2;                        // first window word
            const uint32_t widx = w0 + (uint32_t)lane;
            const uint32_t w = (widx <= last4) ? src4[widx] : 0u;
            // header length if a token started at each of my 4 bytes: 3 + literals, + 1 if the match nibble is 15 — the four
            // bytes of the word at once, a nibble sum per byte (at most 19: no carry from byte to byte)
            const uint32_t wn = __shfl_down_sync(FULL_MASK, w, 1);       // lane 31 gets junk: tokens there end the batch
            const uint32_t L4 = (w >> 4) & 0x0F0F0F0Fu;
            const uint32_t M15 = (((w & 0x0F0F0F0Fu) + 0x01010101u) >> 4) & 0x01010101u;
            uint32_t dpack = 0x03030303u + L4 + M15;
            // a literal nibble of 15 takes its extension from the byte after the token (one extension byte only); rare
            const uint32_t L15 = ((L4 + 0x01010101u) >> 4) & 0x01010101u;
            if (L15) {
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    if ((L15 >> (8 * i)) & 1u) {
                        const uint32_t nxt = (i < 3 ? (w >> (8 * i + 8)) : wn) & 0xFFu;
                        const uint32_t dl = min(((dpack >> (8 * i)) & 0xFFu) + 1u + nxt, 255u);
                        dpack = (dpack & ~(0xFFu << (8 * i))) | (dl << (8 * i));
                    }
                }
            }
            // walk the chain through shared memory.  The address register IS the walk: the area is 256-byte aligned, so
            // the low byte of the address is the window offset, and a step is a byte store into the list of visited
            // positions (lane k reads entry k afterwards), a byte load, and an add that saturates below the list (byte
            // 191: nothing is ever written between the header lengths and there).  Four steps per round, none of them conditional: once the walk has left the window (offset > 108)
            // it stays out, and the entries it still writes are recognised by their value.
            bitmap[lane] = dpack;
            win[lane] = w;
            __syncwarp();
            const uint32_t area = (uint32_t)__cvta_generic_to_shared(bitmap);
            const uint32_t qlim = area + (128u - 20u);
            const uint32_t list0 = area + 192u;                 // (above the clamp: the walk never reads what it lists)
            uint32_t qa = area + (a0 & 3u);
            uint32_t lp = list0;
#pragma unroll 1
            while (lp < list0 + 32u && qa <= qlim) {
#pragma unroll
                for (uint32_t u = 0; u < 4; u++) {
                    if (lane == 0) sts_u8(lp + u, qa);          // (every lane walks the same chain; one of them keeps the list)
                    qa = min(qa + lds_u8(qa), area + 191u);
                }
                lp += 4u;
            }
            __syncwarp();
            uint32_t myq = (uint32_t)lane < lp - list0 ? lds_u8(list0 + (uint32_t)lane) : 255u;
            nseq = __popc(__ballot_sync(FULL_MASK, myq <= 128u - 20u));     // positions ascend: the valid ones are a prefix
            if (lane >= nseq) myq = 0;
            // each lane decodes its own header from the window's bytes in shared memory
            const uint32_t winb = (uint32_t)__cvta_generic_to_shared(win);
            const uint32_t tok = lds_u8(winb + myq);
            const uint32_t lext = lds_u8(winb + myq + 1u);
            const uint32_t L = tok >> 4, M = tok & 15u;
            const bool longlit = L == 15u;                      // literal run of 15 + one extension byte (lz4.c:2121-2128)
            const uint32_t lit = longlit ? 15u + lext : L;
            const uint32_t lp1 = myq + (longlit ? 2u : 1u);     // window offset of the first literal
            const bool inwin = lit <= (uint32_t)kMaxBatchLit && lp1 + lit + 2u < 128u;
            const uint32_t ob = winb + (inwin ? lp1 + lit : 0u);    // where the match offset lies
            const uint32_t off = lds_u8(ob) | (lds_u8(ob + 1u) << 8);
            const uint32_t ext = lds_u8(ob + 2u);
            const int pos = ip + (int)(myq - (a0 & 3u));        // position of my token in src
            const int ip1 = pos + (longlit ? 2 : 1);            // first literal
            const int ipn = ip1 + (int)lit + 2 + (M == 15u ? 1 : 0);
            // a long literal run is no shortcut sequence: it must pass liblz4's general-path test on the input side
            // (lz4.c:2279: ip + length <= iend - (2 + 1 + LASTLITERALS)); the output side is tested with the positions
            const bool good = lane < nseq && inwin && pos + 1 < n - 16 && (!longlit || (lext != 255u && ip1 + (int)lit <= n - 8)) &&
                              (M != 15u || (ext != 255u && ipn <= n - LASTLITERALS + 1));
            const uint32_t badseq = __ballot_sync(FULL_MASK, lane < nseq && !good);
            if (badseq) nseq = __ffs(badseq) - 1;               // decode_one takes the first one that does not fit
            my_lit = (int)lit; my_litpos = ip1; my_off = off;
            my_mlen = (int)M + MINMATCH + (M == 15u ? (int)ext : 0);
            my_simple = !longlit && (M != 15u) && off >= 8u;
        }

        if (nseq > 0) {
            // ---- 2. output positions (prefix sum) and the checks that depend on them
            const int span = (lane < nseq) ? my_lit + my_mlen : 0;
            int incl = span;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int up = __shfl_up_sync(FULL_MASK, incl, d);
                if (lane >= d) incl += up;
            }
            const int o = op + incl - span;                     // where my literals start
            const int m = o + my_lit;                           // where my match starts
            // a sequence is a shortcut only while op <= cap-32 (output side); the start bitmap below covers 1024 output
            // bytes: cut the batch at the first sequence that breaks either limit
            // ... and a long literal run must end MFLIMIT short of the capacity (lz4.c:2279: cpy <= oend - MFLIMIT)
            const uint32_t late = __ballot_sync(FULL_MASK, lane < nseq && (o > cap - 32 || o + span - op > kBatchBytes ||
                                                                            (my_lit >= 15 && m > cap - MFLIMIT)));
            if (late) nseq = __ffs(late) - 1;
            if (nseq > 0) {
                const bool mine = lane < nseq;
                const bool stage2 = my_simple && (int)my_off <= m;
                const bool bad = mine && !stage2 &&
                                 ((check_offset && m - (int)my_off + dsz < 0) || my_off == 0 || m + my_mlen > cap - LASTLITERALS);
                const uint32_t badmask = __ballot_sync(FULL_MASK, bad);
                if (badmask) return -__shfl_sync(FULL_MASK, my_ipn(), __ffs(badmask) - 1) - 1;

                // ---- 3. copy: 32 output bytes per step, one per lane
                const int out0 = op;
                const int out1 = __shfl_sync(FULL_MASK, o + span, nseq - 1);
                const int total = out1 - out0;
                const int ip_next = __shfl_sync(FULL_MASK, my_ipn(), nseq - 1);
                const int orel = mine ? o - out0 : 0x7FFFFFF;                 // sequences outside the batch start "never"
                // bitmap of sequence starts over the batch's output range: lane j ends up with bits out0+32j .. out0+32j+31
                bitmap[lane] = 0;
                __syncwarp();
                if (mine) atomicOr(&bitmap[orel >> 5], 1u << (orel & 31));
                __syncwarp();
                const uint32_t my_bits = bitmap[lane];
                // what a byte needs to know about its sequence, in two words: where the match starts (relative to out0:
                // bytes below it are literals) with the offset, and where its literals lie in the compressed block
                // (literal for the byte at batch position x: src[litrel + x])
                const uint32_t packA = (uint32_t)((orel + my_lit) & 0x7FF) | (my_off << 16);
                const int litrel = my_litpos - orel;
                if (kDuo) {
                    // hand the batch to the copier: its slot is free once the batch that was there has been taken
                    duo_wait(&D->taken, nb - (kDuoRing - 1), lane);
                    DuoSlot& sl = D->slot[nb & (kDuoRing - 1)];
                    sl.packA[lane] = packA; sl.litrel[lane] = litrel; sl.bits[lane] = my_bits;
                    if (lane == 0) { D->out0[nb & (kDuoRing - 1)] = out0; D->total[nb & (kDuoRing - 1)] = total; }
                    __syncwarp();
                    if (lane == 0) { __threadfence_block(); D->published = nb + 1; }
                    nb++;
                    ip = ip_next;
                    op = out1;
                    continue;
                }
                const uint32_t lmask = (2u << lane) - 1u;
                uint8_t* dx = dst + out0 + lane;                              // my byte of the chunk
                const uint8_t* sx = src + lane;                               // ... and where it would come from as a literal, less litrel
                int kbase = -1;                                               // sequences that start below the chunk, minus one
#pragma unroll 1
                for (int xr = lane; (xr & ~31) < total; xr += 32, dx += 32, sx += 32) {            // xr: byte position relative to out0
                    const int ln = xr & 31;
                    // owner of the byte = last sequence starting at or before it
                    const uint32_t sbits = __shfl_sync(FULL_MASK, my_bits, xr >> 5);
                    const int k = kbase + __popc(sbits & lmask);
                    kbase += __popc(sbits);
                    const uint32_t ka = __shfl_sync(FULL_MASK, packA, k);
                    const int kp = __shfl_sync(FULL_MASK, litrel, k);
                    const bool live = xr < total;
                    const bool is_lit = xr < (int)(ka & 0x7FFu);
                    const int off = (int)(ka >> 16);
                    const bool fwd = live && !is_lit && off <= ln;                                  // match source inside this chunk
                    uint32_t val = 0;
                    if (kRing) {
                        if (live && !fwd) {
                            const int s = out0 + xr - off;
                            if (is_lit) val = sx[kp];
                            else if (kDict && s < 0) val = dict[dsz + s];
                            else val = ring[s & 0xFFFF];
                        }
                    } else {
                        // one load from one pointer: the literal in the compressed block, or the match source `off` below me
                        const uint8_t* p = is_lit ? sx + kp : static_cast<const uint8_t*>(dx) - off;
                        if (kDict) {
                            const int s = out0 + xr - off;
                            if (!is_lit && s < 0) p = dict + (dsz + s);
                        }
                        val = ld_global_u8_if(p, live && !fwd);
                    }
                    if (__any_sync(FULL_MASK, fwd)) {
                        // forward values along in-chunk chains: root = the lane whose loaded value this byte finally equals
                        int root = fwd ? ln - off : ln;
#pragma unroll
                        for (int it = 0; it < 5; it++) root = __shfl_sync(FULL_MASK, root, root);
                        val = __shfl_sync(FULL_MASK, val, root);
                    }
                    if (live) {
                        *dx = (uint8_t)val;
                        if (kRing) ring[(out0 + xr) & 0xFFFF] = (uint8_t)val;
                    }
                    __syncwarp();
                }
                ip = ip_next;
                op = out1;
                continue;
            }
        }

        // ---- anything that is not a shortcut sequence: one sequence through the literal state machine
        int32_t ret = 0;
        const int op_before = op;
        if (kDuo) duo_wait(&D->copied, nb, lane);       // decode_one reads and writes the output itself: everything listed is there
        if (decode_one<kDict>(src, n, dst, cap, dict, dsz, lane, ip, op, ret) == kStepDone) return ret;
        __syncwarp();                                   // its stores may be the next batch's match sources
        if (kRing) {
            // decode_one works on global memory: mirror what it produced into the ring
            for (int k = max(op_before, op - 65536) + lane; k < op; k += 32) ring[k & 0xFFFF] = dst[k];
            __syncwarp();
        }
    }
}

// 56 warps per SM at 36 registers beat 32 warps at 64 despite the spills (measured at 8, 10, 12, 14 CTAs per SM: 14 is
// the fastest, +5 % over 12): the kernel waits on loads of match sources that miss L2, and more warps in flight hide more of them
#ifndef PLZ4CU_DEC_CTAS
#define PLZ4CU_DEC_CTAS 14
#endif
template <bool kDict, bool kRing>
__global__ void __launch_bounds__(kDecodeThreads, kRing ? 1 : PLZ4CU_DEC_CTAS)
lz4_decompress_kernel(DecodeArgs a)
{
    extern __shared__ __align__(16) uint8_t dyn_smem[];              // kRing: 64 KiB per warp
    // per warp: 256 bytes, 256-byte aligned — header lengths of the window's 128 positions (later the bitmap of sequence
    // starts), spare (the walk's steps past the window read here), the list of token positions (32 bytes from byte 192);
    // and the window's 128 bytes themselves
    __shared__ __align__(256) uint32_t s_bitmap[kDecodeThreads / 32][64];
    __shared__ uint32_t s_win[kDecodeThreads / 32][32];
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    // the token walk stores the low byte of a shared-memory address as a window offset: the area must sit on a 256-byte line
    if (((uint32_t)__cvta_generic_to_shared(&s_bitmap[0][0]) & 255u) != 0u) __trap();
    const uint32_t wpb = kRing ? 1u : (uint32_t)(kDecodeThreads / 32);
    if (kRing && warp != 0) return;
    const uint32_t b = blockIdx.x * wpb + warp;
    if (b >= a.nblk) return;

    const uint8_t* rec = a.rec_base + a.rec_off[b];
    uint8_t* out = a.dst_base + (uint64_t)b * a.dst_stride;
    // opaque from here on: with 40 registers the compiler would rather rebuild this pointer from blockIdx and the
    // arguments at every use (nine instructions per load and per store of the copy loop) than keep or spill it
    asm("" : "+l"(out));
    __builtin_assume(__isGlobal(out));
    const uint8_t* payload;
    uint32_t csize;
    bool stored = false;

    if (a.raw_blocks) {
        payload = rec;
        csize = a.raw_len[b];
    } else {
        uint32_t word = load_le32(rec);
        stored = (word & 0x80000000u) != 0;
        csize = word & 0x7FFFFFFFu;
        payload = rec + 4;
        if (csize > a.dst_cap) {                                  // blk/frame.go:79-81
            if (lane == 0) a.out_len[b] = PLZ4CU_E_OVERFLOW_;
            return;
        }
    }
    const bool hash_on = !a.raw_blocks && a.verify_checksum != 0;    // blk/frame.go:114-127
    HashAlong H{hash_on && !stored, xxh32_init(lane), 0};

    int32_t r;
    if (stored) {
        if (hash_on && load_le32(payload + csize) != warp_xxh32(payload, csize, lane)) {
            if (lane == 0) a.out_len[b] = PLZ4CU_E_BLOCKHASH_;
            return;
        }
        warp_copy(out, payload, csize, lane);
        r = (int32_t)csize;
    } else {
        r = decode_block<kDict, kRing>(payload, (int)csize, out, (int)a.dst_cap, a.dict, (int)a.dict_size, lane,
                                       s_bitmap[warp], s_win[warp], dyn_smem, H);
        if (H.on) {
            // the rest of the payload, whatever the decoder made of it: a wrong checksum outranks a decode error
            // (the reference checks it before it decodes, blk/frame.go:114-127)
            hash_along(H, payload, (int)csize, lane);
            if (load_le32(payload + csize) != xxh32_finish_global(H.acc, payload, (uint32_t)H.done, csize, lane)) r = PLZ4CU_E_BLOCKHASH_;
        }
    }
    if (lane == 0) a.out_len[b] = r;
}

// the copier of a duo: takes the batches in order and produces their bytes (decode_block step 3)
template <bool kDict>
__device__ __forceinline__ void duo_copy(DuoShared* D, const uint8_t* __restrict__ src, uint8_t* dst,
                                         const uint8_t* __restrict__ dict, int dsz, int lane)
{
    const uint32_t lmask = (2u << lane) - 1u;
    for (int k = 0;; k++) {
        for (int spins = 0;; spins++) {
            int v = 0;
            if (lane == 0) { const int q = D->quit; const int p = D->published; v = p > k ? 1 : (q ? -1 : 0); }   // quit first: it is raised last
            v = __shfl_sync(FULL_MASK, v, 0);
            if (v > 0) break;
            if (v < 0) return;
            if (spins > kDuoSpinLimit) __trap();
            __nanosleep(PLZ4CU_DUO_NAP);
        }
        __threadfence_block();
        const DuoSlot& sl = D->slot[k & (kDuoRing - 1)];
        const uint32_t packA = sl.packA[lane];
        const int litrel = sl.litrel[lane];
        const uint32_t my_bits = sl.bits[lane];
        const int out0 = D->out0[k & (kDuoRing - 1)], total = D->total[k & (kDuoRing - 1)];
        __syncwarp();
        if (lane == 0) { __threadfence_block(); D->taken = k + 1; }
        uint8_t* dx = dst + out0 + lane;                              // my byte of the chunk
        const uint8_t* sx = src + lane;                               // ... and where it would come from as a literal, less litrel
        int kbase = -1;                                               // sequences that start below the chunk, minus one
#pragma unroll 1
        for (int xr = lane; (xr & ~31) < total; xr += 32, dx += 32, sx += 32) {
            const int ln = xr & 31;
            const uint32_t sbits = __shfl_sync(FULL_MASK, my_bits, xr >> 5);
            const int q = kbase + __popc(sbits & lmask);
            kbase += __popc(sbits);
            const uint32_t ka = __shfl_sync(FULL_MASK, packA, q);
            const int kp = __shfl_sync(FULL_MASK, litrel, q);
            const bool live = xr < total;
            const bool is_lit = xr < (int)(ka & 0x7FFu);
            const int off = (int)(ka >> 16);
            const bool fwd = live && !is_lit && off <= ln;
            const uint8_t* p = is_lit ? sx + kp : static_cast<const uint8_t*>(dx) - off;
            if (kDict) {
                const int s = out0 + xr - off;
                if (!is_lit && s < 0) p = dict + (dsz + s);
            }
            uint32_t val = ld_global_u8_if(p, live && !fwd);
            if (__any_sync(FULL_MASK, fwd)) {
                int root = fwd ? ln - off : ln;
#pragma unroll
                for (int it = 0; it < 5; it++) root = __shfl_sync(FULL_MASK, root, root);
                val = __shfl_sync(FULL_MASK, val, root);
            }
            if (live) *dx = (uint8_t)val;
            __syncwarp();
        }
        if (lane == 0) { __threadfence_block(); D->copied = k + 1; }
    }
}

// two blocks per CTA, two warps per block
#ifndef PLZ4CU_DUO_CTAS
#define PLZ4CU_DUO_CTAS 16
#endif
template <bool kDict>
__global__ void __launch_bounds__(kDecodeThreads, PLZ4CU_DUO_CTAS)
lz4_decompress_duo_kernel(DecodeArgs a)
{
    constexpr int kPairs = kDecodeThreads / 64;
    __shared__ __align__(256) uint32_t s_bitmap[kPairs][64];
    __shared__ uint32_t s_win[kPairs][32];
    __shared__ DuoShared s_duo[kPairs];
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    const int pair = warp >> 1;
    const bool copier = (warp & 1) != 0;
    if (((uint32_t)__cvta_generic_to_shared(&s_bitmap[0][0]) & 255u) != 0u) __trap();
    if (threadIdx.x < kPairs) { s_duo[threadIdx.x].published = 0; s_duo[threadIdx.x].taken = 0; s_duo[threadIdx.x].copied = 0; s_duo[threadIdx.x].quit = 0; }
    __syncthreads();
    const uint32_t b = blockIdx.x * kPairs + pair;
    if (b >= a.nblk) return;
    DuoShared* D = &s_duo[pair];

    const uint8_t* rec = a.rec_base + a.rec_off[b];
    uint8_t* out = a.dst_base + (uint64_t)b * a.dst_stride;
    asm("" : "+l"(out));
    __builtin_assume(__isGlobal(out));
    const uint8_t* payload;
    uint32_t csize;
    bool stored = false;
    if (a.raw_blocks) {
        payload = rec;
        csize = a.raw_len[b];
    } else {
        uint32_t word = load_le32(rec);
        stored = (word & 0x80000000u) != 0;
        csize = word & 0x7FFFFFFFu;
        payload = rec + 4;
        if (csize > a.dst_cap) {                                  // blk/frame.go:79-81
            if (lane == 0 && !copier) a.out_len[b] = PLZ4CU_E_OVERFLOW_;
            return;
        }
    }
    if (copier) {
        if (!stored) duo_copy<kDict>(D, payload, out, a.dict, (int)a.dict_size, lane);
        return;
    }
    const bool hash_on = !a.raw_blocks && a.verify_checksum != 0;    // blk/frame.go:114-127
    HashAlong H{hash_on && !stored, xxh32_init(lane), 0};
    int32_t r;
    if (stored) {
        if (hash_on && load_le32(payload + csize) != warp_xxh32(payload, csize, lane)) {
            if (lane == 0) a.out_len[b] = PLZ4CU_E_BLOCKHASH_;
            return;
        }
        warp_copy(out, payload, csize, lane);
        r = (int32_t)csize;
    } else {
        r = decode_block<kDict, false, true>(payload, (int)csize, out, (int)a.dst_cap, a.dict, (int)a.dict_size, lane,
                                             s_bitmap[pair], s_win[pair], nullptr, H, D);
        __syncwarp();
        if (lane == 0) { __threadfence_block(); D->quit = 1; }       // (every path out of decode_block: the copier must end)
        if (H.on) {
            hash_along(H, payload, (int)csize, lane);
            if (load_le32(payload + csize) != xxh32_finish_global(H.acc, payload, (uint32_t)H.done, csize, lane)) r = PLZ4CU_E_BLOCKHASH_;
        }
    }
    if (lane == 0) a.out_len[b] = r;
}

// ---------------------------------------------------------------- team decode: one CTA per block
//
// A launch with few, large blocks (plz4's default 4 MiB block: 64 blocks per 256 MiB) cannot be filled by one warp per
// block, and a lone warp is bound by its own instruction latency: walking the token chain (one dependent shared-memory
// load per sequence) is more than a third of it, header decode and the scan most of the rest.  The team kernel gives a
// block one CTA and the work five roles, each a stage that only waits for the stage before it:
//
//  * TABLE warps look at every byte position of a superwindow (4 KiB of compressed bytes at a fixed place: 32 segments
//    of 128, one per lane) as if a token started there: literal count, match length, whether it fits the batched shape,
//    and where the chain that starts there leaves the segment — a backward pass per lane in which position q looks up
//    position q + header(q), already done.  Nothing in it depends on where the true chain runs, so superwindows are
//    prepared ahead of the parse, several at a time.
//  * the PARSER follows the true chain: one lookup per segment, then every lane lists the tokens of its own segment;
//    the list goes out in batches of 32 sequences with the output position of each batch.  Tokens that do not fit the
//    batched shape, and everything near the end of the block, go through decode_one on this warp after the others have
//    drained — so accept/reject and return codes are those of the one-warp decoder.
//  * DECODER warps take batches round-robin: offsets, output positions by prefix sum, the checks that depend on them
//    (the first failing sequence in stream order decides the block's code), the bitmap of sequence starts; they publish
//    in order into a ring of slots.
//  * COPY warps produce the bytes, 32-byte chunk by chunk, chunks dealt round-robin.  A chunk waits only for the output
//    its own matches read (`prog`: per warp, the position below which all of that warp's chunks are complete), so
//    chunks whose sources lie further back than the chunks in flight proceed in parallel; literals never wait.
//  * one warp verifies the block checksum meanwhile.
constexpr int kTeamTabs = 3;                                    // superwindow tables; table warp t owns table t and prepares
constexpr int kTeamTabWarps = kTeamTabs;                        // superwindows t, t + 3, ...: two are ahead of the parser's
constexpr int kTeamDecWarps = 3;
constexpr int kTeamCopyWarps = 16;                              // at most; a launch chooses 4, 8 or 16 (PLZ4CU_TEAM_COPY)
constexpr int kTeamFirstCopy = 1 + kTeamTabWarps + kTeamDecWarps + 1;   // warps: parser, tables, decoders, checksum, copy...
constexpr int kTeamThreads = (kTeamFirstCopy + kTeamCopyWarps) * 32;
constexpr int kTeamSeqRing = 64;                                // batches between parser and decoders
constexpr int kTeamSlots = 16;                                  // batches between decoders and copy warps
// The output window in shared memory is a ring of R bytes (64 KiB with one team per SM, 32 KiB when two teams share an SM).
// At most R/2 output bytes may be in flight, and matches reaching back R/2 or more read global memory instead of the
// ring: then no chunk in flight can overwrite a ring byte another chunk in flight still reads.
constexpr unsigned kTeamBackoffNs = 320;                        // pause between polls of a table warp that is ahead of the parse
constexpr int kTeamSpinLimit = 1 << 25;                         // watchdog: a stalled team reports PLZ4CU_E_STALL, it never hangs
constexpr int kSwBytes = 4096;                                  // superwindow
constexpr int kSwSlack = 128;                                   // bytes past it a header may touch (<= 68)
constexpr int kSwWords = (kSwBytes + kSwSlack) / 4;
constexpr int kSwMaxBatch = (kSwBytes / 3 + 3 + 31) / 32;       // a sequence header is at least 3 bytes
static_assert(kSwMaxBatch <= kTeamSeqRing, "a superwindow's batches must fit the ring");
constexpr int kTeamCapMargin = 128;                             // batched sequences end at least this far below the capacity

// per-position word of the backward pass
constexpr uint32_t kExPosMask = 0x1FFFu;        // bits 0-12: where the chain from here leaves the segment (superwindow-relative)
constexpr uint32_t kExStop = 0x2000u;           // bit 13: ... it does not: it stops at that position (a token decode_one must take)
constexpr int kExLitShift = 14;                 // bits 14-19: literal count (<= 63)
constexpr int kExMlenShift = 20;                // bits 20-28: match length (<= 273)
constexpr uint32_t kExBad = 1u << 30;           // the token here does not fit the batched shape

__device__ __forceinline__ int ex_lit(uint32_t wd) { return (int)((wd >> kExLitShift) & 63u); }
__device__ __forceinline__ int ex_mlen(uint32_t wd) { return (int)((wd >> kExMlenShift) & 0x1FFu); }
// header length: token, [literal extension], literals, offset, [match extension]
__device__ __forceinline__ int ex_dl(int lit, int mlen) { return (lit >= 15 ? 2 : 1) + lit + 2 + (mlen >= MINMATCH + 15 ? 1 : 0); }

struct TeamSeqBatch {                           // parser -> decoder
    int nseq, out0, ip0, pad;
    uint32_t rec[32];                           // token position relative to ip0 (13 bits) | literals << 13 | match length << 19
};
struct TeamSlot {                               // decoder -> copy warps
    uint32_t packA[32];                         // per sequence: start relative to out0 (10 bits) | literals (6 bits) | offset (16 bits)
    int litpos[32];                             // per sequence: position of its first literal in the compressed block
    uint32_t bits[32];                          // bitmap of sequence starts over the slot's output range (<= 1024 bytes)
    int nseq, out0, out1, pad;
};

struct TeamShared {
    uint32_t tab[kTeamTabs][128 * 32];          // per-position words, [position in segment][segment]
    uint32_t cw[kTeamTabWarps][kSwWords + kSwWords / 32 + 2];   // a table warp's superwindow bytes (one word of padding per
                                                                // 32: lane l reads word j of its segment from bank l + j)
    TeamSeqBatch sb[kTeamSeqRing];
    TeamSlot slot[kTeamSlots];
    unsigned long long err;                     // first failing sequence: (batch * 32 + lane) << 32 | code
    volatile int tab_ready[kTeamTabs];          // the superwindow a table describes
    volatile int parser_sw;                     // the superwindow the parser is in; tables of earlier ones are free
    volatile int sb_head;                       // batches listed by the parser
    volatile int dec_next[kTeamDecWarps];       // the batch a decoder will read next
    volatile int dec_done;                      // batches the decoders have dealt with, in order
    volatile int head;                          // slots published
    volatile int pend;                          // output position the published slots end at
    volatile int prog[kTeamCopyWarps];
    volatile int passed[kTeamCopyWarps];        // slots a copy warp has left behind
    volatile int quit;                          // nothing will be listed any more
    volatile int stall;                         // watchdog fired
    volatile int hash_state;                    // 0 running, 1 checksum ok, 2 mismatch
    int ring_mask;                              // R - 1
};

// Spin until the first `cnt` entries of `arr` have all reached `need`.  The decision is a warp vote: the warp stays converged.
__device__ __forceinline__ bool team_wait(TeamShared* ts, const volatile int* arr, int cnt, int need, int lane)
{
    for (int spins = 0;; spins++) {
        const int v = lane < cnt ? arr[lane] : 0x7FFFFFFF;
        if (__all_sync(FULL_MASK, v >= need)) break;
        const int st = ts->stall;
        if (__any_sync(FULL_MASK, st != 0) || spins > kTeamSpinLimit) {
            if (lane == 0) ts->stall = 1;
            return false;
        }
    }
    __threadfence_block();
    return true;
}

// Lane 0 looks at a counter and at the quit / stall flags (flags first: quit is raised after the last publication),
// everybody gets the same answer: the counter, or -1 once nothing more can come.
__device__ __forceinline__ int team_poll(TeamShared* ts, const volatile int* counter, int want_above, int lane)
{
    int r = 0;
    if (lane == 0) {
        const int flags = (ts->quit ? 1 : 0) | (ts->stall ? 2 : 0);
        const int c = *counter;
        r = c > want_above ? c : (flags ? -1 : c);
    }
    return __shfl_sync(FULL_MASK, r, 0);
}

__device__ __forceinline__ int sw_skew(int word) { return word + (word >> 5); }

// ---- a table warp: superwindows t, t + kTeamTabs, ... into table t
__device__ __forceinline__ void team_tables(TeamShared* ts, const uint8_t* __restrict__ src, int n, int t, int lane)
{
    const uintptr_t sa = reinterpret_cast<uintptr_t>(src);
    const uint32_t* __restrict__ src4 = reinterpret_cast<const uint32_t*>(sa & ~uintptr_t(3));
    const uint32_t d4 = (uint32_t)sa & 3u;
    const uint32_t last4 = (d4 + (uint32_t)n - 1u) >> 2;         // last word holding a valid byte
    uint32_t* const cw = ts->cw[t];
    const int nsw = (n + kSwBytes - 1) / kSwBytes;

    for (int k = t;; k += kTeamTabWarps) {
        // skip what the parser has left behind; wait for a free table
        for (int spins = 0;; spins++) {
            int psw = 0;
            if (lane == 0) psw = (ts->quit | ts->stall) ? -1 : ts->parser_sw;
            psw = __shfl_sync(FULL_MASK, psw, 0);
            if (psw < 0) return;
            if (k < psw) k += (psw - k + kTeamTabWarps - 1) / kTeamTabWarps * kTeamTabWarps;
            if (k >= nsw) return;
            if (k < psw + kTeamTabs) break;
            if (spins > kTeamSpinLimit) {
                if (lane == 0) ts->stall = 1;
                return;
            }
            __nanosleep(kTeamBackoffNs);            // tables are two superwindows ahead: no hurry, and a polling warp takes
                                                    // issue slots from the warps it waits for
        }
        __threadfence_block();
        uint32_t* const ex = ts->tab[k % kTeamTabs];
        const int sw_ip = k * kSwBytes;
        // ---- bytes of the superwindow into shared memory (zero past the end of the block)
        {
            const uint32_t a0 = d4 + (uint32_t)sw_ip;
            const uint32_t w0 = a0 >> 2, sh = (a0 & 3u) * 8u;
            // word i of the window = aligned words w0+i and w0+i+1 funnel-shifted; rows of 32 words, the next row's first
            // word comes along by shuffle
            uint32_t row = (w0 + (uint32_t)lane) <= last4 ? src4[w0 + lane] : 0u;
#pragma unroll 3
            for (int r = 0; r < kSwWords / 32; r++) {
                const uint32_t wi = w0 + (uint32_t)(32 * (r + 1) + lane);
                const uint32_t nrow = wi <= last4 ? src4[wi] : 0u;               // the last one reads one row past the window
                uint32_t hi = __shfl_down_sync(FULL_MASK, row, 1);
                const uint32_t n0 = __shfl_sync(FULL_MASK, nrow, 0);
                if (lane == 31) hi = n0;
                cw[sw_skew(32 * r + lane)] = sh ? __funnelshift_r(row, hi, sh) : row;
                row = nrow;
            }
            // ask for my next superwindow meanwhile
            const int ahead = sw_ip + kTeamTabWarps * kSwBytes + 128 * lane;
            if (ahead < n) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + ahead));
        }
        __syncwarp();
        // ---- backward pass: lane l owns segment l
        const int seg0 = lane * 128;
#pragma unroll 1
        for (int j = 31; j >= 0; j--) {
            const uint32_t w = cw[sw_skew(lane * 32 + j)];
            const uint32_t wn = cw[sw_skew(lane * 32 + j + 1)];
            // the four positions of this word: everything that depends on the bytes alone first ...
            uint32_t lit4[4], ml4[4], dl4[4];
            bool bad4[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t tok = (w >> (8 * i)) & 0xFFu;
                const uint32_t nxt = (i < 3 ? (w >> (8 * i + 8)) : wn) & 0xFFu;
                const int qs = seg0 + 4 * j + i;
                const uint32_t L = tok >> 4, M = tok & 15u;
                const bool longlit = L == 15u;
                const uint32_t lit = longlit ? 15u + nxt : L;
                const uint32_t hdr = (longlit ? 2u : 1u) + min(lit, 63u) + 2u;  // token .. offset
                const int eb = qs + (int)hdr;                                    // where a match extension byte would be
                const uint32_t ext = M == 15u ? (cw[sw_skew(eb >> 2)] >> ((eb & 3) * 8)) & 0xFFu : 0u;
                const uint32_t dl = hdr + (M == 15u ? 1u : 0u);
                const int pos = sw_ip + qs;
                // the batched shape and its input-side limits (decode_block: `inwin`, `good`)
                bad4[i] = lit > 63u || !(pos + 1 < n - 16) || (longlit && pos + 2 + (int)lit > n - 8) ||
                          (M == 15u && (ext == 255u || pos + (int)dl > n - LASTLITERALS + 1));
                lit4[i] = lit; ml4[i] = M + (uint32_t)MINMATCH + ext; dl4[i] = dl;
            }
            // ... then the chain lookups, last position first (a lookup may land on a position of this very word)
#pragma unroll
            for (int i = 3; i >= 0; i--) {
                const int q = 4 * j + i;
                const int t2 = q + (int)dl4[i];
                uint32_t e;
                if (bad4[i]) e = (uint32_t)(seg0 + q) | kExStop | kExBad;
                else {
                    if (t2 >= 128) e = (uint32_t)(seg0 + t2);
                    else e = ex[t2 * 32 + lane] & (kExPosMask | kExStop);
                    e |= (lit4[i] << kExLitShift) | (ml4[i] << kExMlenShift);
                }
                ex[q * 32 + lane] = e;
            }
        }
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            ts->tab_ready[k % kTeamTabs] = k;
        }
    }
}

// ---- the parser warp
template <bool kDict>
__device__ __forceinline__ int32_t team_parse(const uint8_t* __restrict__ src, int n, uint8_t* dst, int cap,
                                              const uint8_t* __restrict__ dict, int dsz, int lane, uint8_t* ring, TeamShared* ts,
                                              int nw)
{
    if (cap == 0) return (n == 1 && src[0] == 0) ? 0 : -1;
    if (n == 0) return -1;

    int ip = 0, op = 0;
    int ks = -1;                    // superwindow whose table `ex` points at
    const uint32_t* ex = nullptr;
    int batches = 0;                // batches listed
    bool tail = false;              // the output is within kTeamCapMargin of the capacity: decode_one to the end

    for (;;) {
        if (!tail) {
            if ((ip >> 12) != ks) {
                ks = ip >> 12;
                if (lane == 0) ts->parser_sw = ks;
                const volatile int* ready = &ts->tab_ready[ks % kTeamTabs];
                for (int spins = 0;; spins++) {
                    const int v = *ready;
                    const int st = ts->stall;
                    if (__all_sync(FULL_MASK, v == ks)) break;
                    if (__any_sync(FULL_MASK, st != 0) || spins > kTeamSpinLimit) {
                        if (lane == 0) ts->stall = 1;
                        return PLZ4CU_E_STALL_;
                    }
                }
                __threadfence_block();
                ex = ts->tab[ks % kTeamTabs];
            }
            const int sw_ip = ks << 12;

            // ---- the true chain: one lookup per segment from the entry position
            int cur = ip - sw_ip;
            int my_entry = -1;
            bool stopped = false;
            while (cur < kSwBytes) {
                const int s = cur >> 7;
                if (lane == s) my_entry = cur;
                const uint32_t wd = ex[(cur & 127) * 32 + s];
                cur = (int)(wd & kExPosMask);
                if (wd & kExStop) { stopped = true; break; }
            }
            // `cur`: where this stretch ends (the stopping token, or the first token past the superwindow)
            // ---- every lane measures its own segment: sequences and output bytes
            int cnt = 0, sp = 0;
            if (my_entry >= 0) {
                const int end = stopped ? min(cur, (lane + 1) * 128) : (lane + 1) * 128;
                for (int q = my_entry; q < end;) {
                    const uint32_t wd = ex[(q & 127) * 32 + lane];
                    const int lit = ex_lit(wd), mlen = ex_mlen(wd);
                    cnt++;
                    sp += lit + mlen;
                    q += ex_dl(lit, mlen);
                }
            }
            int icnt = cnt, isp = sp;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int uc = __shfl_up_sync(FULL_MASK, icnt, d), us = __shfl_up_sync(FULL_MASK, isp, d);
                if (lane >= d) { icnt += uc; isp += us; }
            }
            const int ecnt = icnt - cnt, esp = isp - sp;           // exclusive
            int total = __shfl_sync(FULL_MASK, icnt, 31);
            int total_sp = __shfl_sync(FULL_MASK, isp, 31);
            int next_ip = sw_ip + cur;
            // near the capacity liblz4's end-of-block rules apply: stop the stretch before the first segment that gets there
            const uint32_t over = __ballot_sync(FULL_MASK, my_entry >= 0 && op + isp > cap - kTeamCapMargin);
            if (over) {
                const int lc = __ffs(over) - 1;
                total = __shfl_sync(FULL_MASK, ecnt, lc);
                total_sp = __shfl_sync(FULL_MASK, esp, lc);
                next_ip = sw_ip + __shfl_sync(FULL_MASK, my_entry, lc);
                if (lane >= lc) cnt = 0;
                tail = true;
                stopped = true;
            }
            if (total > 0) {
                // ---- list the stretch: batches of 32 sequences for the decoders
                const int nb = (total + 31) >> 5;
                if (!team_wait(ts, ts->dec_next, kTeamDecWarps, batches + nb - kTeamSeqRing, lane)) return PLZ4CU_E_STALL_;
                if (cnt > 0) {
                    int g = ecnt, o = op + esp;
                    for (int q = my_entry, i = 0; i < cnt; i++, g++) {
                        const uint32_t wd = ex[(q & 127) * 32 + lane];
                        const int lit = ex_lit(wd), mlen = ex_mlen(wd);
                        TeamSeqBatch& e = ts->sb[(batches + (g >> 5)) % kTeamSeqRing];
                        e.rec[g & 31] = (uint32_t)q | ((uint32_t)lit << 13) | ((uint32_t)mlen << 19);
                        if ((g & 31) == 0) e.out0 = o;
                        o += lit + mlen;
                        q += ex_dl(lit, mlen);
                    }
                }
                for (int j = lane; j < nb; j += 32) {
                    TeamSeqBatch& e = ts->sb[(batches + j) % kTeamSeqRing];
                    e.nseq = min(32, total - 32 * j);
                    e.ip0 = sw_ip;
                }
                batches += nb;
                op += total_sp;
                __syncwarp();
                if (lane == 0) {
                    __threadfence_block();
                    ts->sb_head = batches;
                }
            }
            ip = next_ip;
            if (!stopped) continue;                                 // the chain ran off the superwindow: next one
        }

        // ---- one sequence through the literal state machine, once everything listed has been written
        if (!team_wait(ts, &ts->dec_done, 1, batches, lane)) return PLZ4CU_E_STALL_;
        if (!team_wait(ts, ts->passed, nw, ts->head, lane)) return PLZ4CU_E_STALL_;
        {
            const unsigned long long err = ts->err;
            if (err != ~0ull) return (int32_t)(uint32_t)err;
        }
        int32_t ret = 0;
        const int op_before = op;
        if (decode_one<kDict>(src, n, dst, cap, dict, dsz, lane, ip, op, ret) == kStepDone) return ret;
        __syncwarp();
        // decode_one works on global memory: mirror what it produced into the ring
        const int rmask = ts->ring_mask;
        for (int k = max(op_before, op - rmask - 1) + lane; k < op; k += 32) ring[k & rmask] = dst[k];
        __syncwarp();
    }
}

// ---- a decoder warp: batches d, d + kTeamDecWarps, ...
__device__ __forceinline__ void team_decode(TeamShared* ts, const uint8_t* __restrict__ src, int cap, int dsz, int d, int lane,
                                            int nw)
{
    const bool check_offset = dsz < 65536;
    const int window = (ts->ring_mask + 1) >> 1;
    for (int b = d;; b += kTeamDecWarps) {
        for (int spins = 0;; spins++) {
            const int c = team_poll(ts, &ts->sb_head, b, lane);
            if (c > b) break;
            if (c < 0) return;
            if (spins > kTeamSpinLimit) {
                if (lane == 0) ts->stall = 1;
                return;
            }
        }
        __threadfence_block();
        const TeamSeqBatch& e = ts->sb[b % kTeamSeqRing];
        const int nseq = e.nseq, out0 = e.out0, ip0 = e.ip0;
        const uint32_t rec = e.rec[lane];
        const bool have = lane < nseq;
        // ---- my sequence's header (decode_block step 1, from what the tables hold)
        const int my_lit = have ? (int)((rec >> 13) & 63u) : 0;
        const int my_mlen = have ? (int)(rec >> 19) : 0;
        const int my_litpos = ip0 + (int)(rec & 0x1FFFu) + (my_lit >= 15 ? 2 : 1);
        const int ob = my_litpos + my_lit;
        const uint32_t my_off = have ? load_u16le(src + ob) : 0u;
        const int my_ipn = ob + 2 + (my_mlen >= MINMATCH + 15 ? 1 : 0);
        const bool my_simple = my_lit < 15 && my_mlen < MINMATCH + 15 && my_off >= 8u;
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            ts->dec_next[d] = b + kTeamDecWarps;                // the parser may reuse the entry
        }
        // ---- output positions and the checks that depend on them (decode_block step 2)
        const int span = my_lit + my_mlen;
        int incl = span;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const int up = __shfl_up_sync(FULL_MASK, incl, s);
            if (lane >= s) incl += up;
        }
        const int o = out0 + incl - span;
        const int m = o + my_lit;
        const int out1 = __shfl_sync(FULL_MASK, out0 + incl, 31);
        const bool stage2 = my_simple && (int)my_off <= m;
        const bool bad = have && !stage2 &&
                         ((check_offset && m - (int)my_off + dsz < 0) || my_off == 0 || m + my_mlen > cap - LASTLITERALS);
        const uint32_t badmask = __ballot_sync(FULL_MASK, bad);

        // ---- my turn to publish
        if (!team_wait(ts, &ts->dec_done, 1, b, lane)) return;
        if (badmask) {
            // the first failing sequence in stream order decides the block's code; nothing of this batch is written
            const int f = __ffs(badmask) - 1;
            const int code = -__shfl_sync(FULL_MASK, my_ipn, f) - 1;
            if (lane == 0) atomicMin(&ts->err, ((unsigned long long)(uint32_t)(b * 32 + f) << 32) | (uint32_t)code);
        } else {
            // nothing may be published more than half a ring ahead of the oldest unfinished byte
            // (everything between the last slot and this batch is decode_one's, hence finished)
            const int pend = ts->pend;
            if (!team_wait(ts, ts->prog, nw, min(out1 - window, pend), lane)) return;
            // slots of at most 1024 output bytes (the start bitmap's reach)
            for (int first = 0; first < nseq;) {
                const int base = __shfl_sync(FULL_MASK, o, first);
                const uint32_t fits = __ballot_sync(FULL_MASK, lane >= first && have && o + span - base <= 1024);
                const int cnt = __popc(fits);
                const bool mine = lane >= first && lane < first + cnt;
                const int end = __shfl_sync(FULL_MASK, o + span, first + cnt - 1);
                const int h = ts->head;
                if (!team_wait(ts, ts->passed, nw, h - kTeamSlots + 1, lane)) return;
                TeamSlot& sl = ts->slot[h % kTeamSlots];
                const int orel = o - base;
                sl.bits[lane] = 0;
                __syncwarp();
                if (mine) {
                    atomicOr(&sl.bits[orel >> 5], 1u << (orel & 31));
                    sl.packA[lane - first] = (uint32_t)(orel & 0x3FF) | ((uint32_t)my_lit << 10) | (my_off << 16);
                    sl.litpos[lane - first] = my_litpos;
                }
                if (lane == 0) { sl.nseq = cnt; sl.out0 = base; sl.out1 = end; }
                __syncwarp();
                if (lane == 0) {
                    __threadfence_block();
                    ts->pend = end;
                    ts->head = h + 1;
                }
                first += cnt;
            }
        }
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            ts->dec_done = b + 1;
        }
    }
}

// ---- a copy warp: takes every published slot in order and produces the chunks dealt to it.
// Chunks are the 32-byte-aligned pieces of the OUTPUT (absolute positions), dealt round-robin: chunk i belongs to warp
// i mod nw whatever slot its bytes come in (a chunk that straddles two slots is produced in two parts by the same warp).
// So the owner of any earlier output byte is known from its position alone, and a lane waits for exactly the warp that
// produces its match source (prog[owner] past that byte) instead of for every chunk below it.
template <bool kDict>
__device__ __forceinline__ void team_copy(TeamShared* ts, const uint8_t* __restrict__ src, uint8_t* dst,
                                          const uint8_t* __restrict__ dict, int dsz, uint8_t* ring, int w, int nw, int lane, int dbg)
{
    const int rmask = ts->ring_mask;
    const uint32_t window = (uint32_t)(rmask + 1) >> 1;
    const uint32_t lmask = (2u << lane) - 1u;
    const int wmask = nw - 1;                       // nw is a power of two
    for (int k = 0;; k++) {
        for (int spins = 0;; spins++) {
            const int c = team_poll(ts, &ts->head, k, lane);
            if (c > k) break;
            if (c < 0) {
                if (lane == 0) ts->prog[w] = 0x7FFFFFFF;
                return;
            }
            if (spins > kTeamSpinLimit) {
                if (lane == 0) ts->stall = 1;
                return;
            }
        }
        __threadfence_block();
        const TeamSlot& sl = ts->slot[k % kTeamSlots];
        const int nseq = sl.nseq, out0 = sl.out0, out1 = sl.out1;
        const uint32_t packA = sl.packA[lane];
        const int my_litpos = sl.litpos[lane];
        const uint32_t my_bits = sl.bits[lane];
        const int orel = lane < nseq ? (int)(packA & 0x3FFu) : 0x7FFFFFF;   // sequences outside the slot start "never"
        const int len = out1 - out0;
        const int a = out0 & 31;                    // the slot starts this far into its first chunk
        const int chunk0 = out0 >> 5;               // absolute index of that chunk
        const int np = (a + len + 31) >> 5;         // chunks the slot touches
        int p = (w - chunk0) & wmask;               // my first one
        __syncwarp();
        if (lane == 0) ts->prog[w] = p < np ? max(out0, (chunk0 + p) << 5) : out1;
        if (dbg & 1) p = np;                         // measurements: everything but the copy
        for (; p < np; p += nw) {
            const int c = 32 * p - a;                                                       // chunk start relative to out0 (< 0: before the slot)
            const int xr = c + lane;                                                        // my byte, relative to out0
            const bool live = xr >= 0 && xr < len;
            // sequence starts inside the chunk: 32 bits of the slot's bitmap from bit c on
            const int wi = c >> 5;                                                          // floor: -1 for the straddling first chunk
            const uint32_t lo = __shfl_sync(FULL_MASK, my_bits, wi & 31), hi = __shfl_sync(FULL_MASK, my_bits, (wi + 1) & 31);
            const uint32_t sbits = __funnelshift_r(wi >= 0 ? lo : 0u, wi + 1 <= 31 ? hi : 0u, (uint32_t)(c & 31));
            const int q = __popc(__ballot_sync(FULL_MASK, orel < c)) - 1 + __popc(sbits & lmask);
            const uint32_t ka = __shfl_sync(FULL_MASK, packA, q);
            const int kp = __shfl_sync(FULL_MASK, my_litpos, q);
            const int d = xr - (int)(ka & 0x3FFu);                                          // byte index inside the sequence
            const bool is_lit = d < (int)((ka >> 10) & 63u);
            const int sr = xr - (int)(ka >> 16);                                            // match source, relative to out0
            const bool fwd = live && !is_lit && sr >= max(c, 0);                            // source inside this chunk's part of the slot
            const int s = out0 + sr;
            uint32_t val = 0;
            if (live && is_lit && !(dbg & 4)) val = src[kp + d];                            // literals wait for nobody
            // match sources that other chunks produce: wait for the warp that owns each of them
            const bool extn = live && !is_lit && !fwd && s >= 0;
            if (__any_sync(FULL_MASK, extn) && !(dbg & 2)) {
                const volatile int* theirs = &ts->prog[(s >> 5) & wmask];
                for (int spins = 0;; spins++) {
                    const bool ok = !extn || *theirs > s;
                    if (__all_sync(FULL_MASK, ok)) break;
                    const int st = ts->stall;
                    if (__any_sync(FULL_MASK, st != 0) || spins > kTeamSpinLimit) {
                        if (lane == 0) ts->stall = 1;
                        return;
                    }
                }
                __threadfence_block();
            }
            if (live && !is_lit && !fwd) {
                if (kDict && s < 0) val = dict[dsz + s];
                else val = (ka >> 16) >= window ? dst[s] : ring[s & rmask];
            }
            if (__any_sync(FULL_MASK, fwd)) {
                int root = fwd ? (sr - c) : lane;
#pragma unroll
                for (int it = 0; it < 5; it++) root = __shfl_sync(FULL_MASK, root, root);
                val = __shfl_sync(FULL_MASK, val, root);
            }
            if (live) {
                dst[out0 + xr] = (uint8_t)val;
                ring[(out0 + xr) & rmask] = (uint8_t)val;
            }
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();
                ts->prog[w] = p + nw < np ? (chunk0 + p + nw) << 5 : out1;
            }
        }
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            ts->passed[w] = k + 1;
        }
    }
}

constexpr int kTeamStateBytes = (int)((sizeof(TeamShared) + 15) & ~size_t(15));

// kPair: two teams per SM (at most 8 copy warps each, 64 registers per thread)
template <bool kDict, bool kPair>
__global__ void __launch_bounds__(kPair ? (kTeamFirstCopy + 8) * 32 : kTeamThreads, kPair ? 2 : 1)
lz4_decompress_team_kernel(DecodeArgs a, int ring_bytes, int dbg)
{
    extern __shared__ __align__(16) uint8_t dyn_smem[];              // the team's state, then the output window
    TeamShared* ts = reinterpret_cast<TeamShared*>(dyn_smem);
    uint8_t* ring = dyn_smem + kTeamStateBytes;
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    const uint32_t b = blockIdx.x;
    const int nwarps = (int)(blockDim.x >> 5);
    const int nw = nwarps - kTeamFirstCopy;                          // copy warps of this launch

    if (threadIdx.x < kTeamCopyWarps) { ts->prog[threadIdx.x] = 0; ts->passed[threadIdx.x] = 0; }
    if (threadIdx.x < kTeamDecWarps) ts->dec_next[threadIdx.x] = threadIdx.x;
    if (threadIdx.x < kTeamTabs) ts->tab_ready[threadIdx.x] = -1;
    if (threadIdx.x == 0) {
        ts->err = ~0ull; ts->parser_sw = 0; ts->sb_head = 0; ts->dec_done = 0; ts->head = 0; ts->pend = 0;
        ts->quit = 0; ts->stall = 0; ts->hash_state = 0; ts->ring_mask = ring_bytes - 1;
    }
    __syncthreads();

    const uint8_t* rec = a.rec_base + a.rec_off[b];
    uint8_t* out = a.dst_base + (uint64_t)b * a.dst_stride;
    const uint8_t* payload;
    uint32_t csize;
    bool stored = false;
    bool verify = false;
    if (a.raw_blocks) {
        payload = rec;
        csize = a.raw_len[b];
    } else {
        const uint32_t word = load_le32(rec);
        stored = (word & 0x80000000u) != 0;
        csize = word & 0x7FFFFFFFu;
        payload = rec + 4;
        if (csize > a.dst_cap) {                                      // blk/frame.go:79-81
            if (threadIdx.x == 0) a.out_len[b] = PLZ4CU_E_OVERFLOW_;
            return;
        }
        verify = a.verify_checksum != 0;
    }

    if (warp == kTeamFirstCopy - 1) {
        // checksum warp (blk/frame.go:114-127): runs beside the decode, its verdict outranks the decoder's
        if (verify) {
            const uint32_t want = load_le32(payload + csize);
            const uint32_t got = warp_xxh32(payload, csize, lane);
            __syncwarp();
            if (lane == 0) ts->hash_state = want == got ? 1 : 2;
        }
        if (!stored) return;
    }

    if (stored) {
        // straight copy (async/reader.go:149-164), a slice per warp
        if (warp != kTeamFirstCopy - 1) {
            const int cw = warp < kTeamFirstCopy ? warp : warp - 1;            // copying warps, numbered without the checksum warp
            const uint32_t piece = ((csize + (uint32_t)nwarps - 2u) / ((uint32_t)nwarps - 1u) + 511u) & ~511u;
            const uint32_t lo = min(csize, (uint32_t)cw * piece), hi = min(csize, lo + piece);
            if (hi > lo) warp_copy(out + lo, payload + lo, hi - lo, lane);
        }
        __syncthreads();
        if (threadIdx.x == 0) a.out_len[b] = (verify && ts->hash_state == 2) ? PLZ4CU_E_BLOCKHASH_ : (int32_t)csize;
        return;
    }

    if (warp == 0) {
        int32_t r = team_parse<kDict>(payload, (int)csize, out, (int)a.dst_cap, a.dict, (int)a.dict_size, lane, ring, ts, nw);
        __syncwarp();
        if (lane == 0) { __threadfence_block(); ts->quit = 1; }
        if (verify) {
            int hs = 0;
            for (int spins = 0; (hs = ts->hash_state) == 0; spins++) {
                if (spins > kTeamSpinLimit) { hs = 3; break; }
                __nanosleep(200);
            }
            if (hs == 2) r = PLZ4CU_E_BLOCKHASH_;
            else if (hs == 3) r = PLZ4CU_E_STALL_;
        }
        if (ts->stall) r = PLZ4CU_E_STALL_;
        if (lane == 0) a.out_len[b] = r;
    } else if (warp <= kTeamTabWarps) {
        if (csize > 0 && a.dst_cap > 0) team_tables(ts, payload, (int)csize, warp - 1, lane);
    } else if (warp <= kTeamTabWarps + kTeamDecWarps) {
        team_decode(ts, payload, (int)a.dst_cap, (int)a.dict_size, warp - 1 - kTeamTabWarps, lane, nw);
    } else if (warp >= kTeamFirstCopy) {
        team_copy<kDict>(ts, payload, out, a.dict, (int)a.dict_size, ring, warp - kTeamFirstCopy, nw, lane, dbg);
    }
}

constexpr uint32_t kRingBlocks = 1024;          // launches with fewer blocks than this use the ring kernel
constexpr int kRingBytes = 65536;
constexpr int kTeamSmem = 131072 + kTeamStateBytes;
int g_sm_count = 148;
int g_team_pair = -1;                           // two teams per SM: -1 when the launch has more blocks than SMs, 0 never, 1 always
int g_team_ring = 65536;                        // output window of a lone team, bytes (PLZ4CU_TEAM_RING: 65536 or 131072)
int g_team_copy = 16;                           // copy warps of a lone team (PLZ4CU_TEAM_COPY); paired teams have 8
int g_team_dbg = 0;                             // PLZ4CU_TEAM_DBG: measurement switches of the team kernel (wrong output)
int g_dec_duo = -1;                             // PLZ4CU_DEC_DUO: 1 = two warps per block (parser + copier) always, 0 = never,
                                                // -1 = when one warp per block would leave half of the SMs' warp slots empty
int g_team = 1;                                 // PLZ4CU_TEAM=0: few large blocks go back to one warp per block (measurements);
                                                // =2: every launch below kRingBlocks blocks takes the team kernel (tests)

cudaError_t configure_decompress()
{
    cudaError_t e = cudaFuncSetAttribute(lz4_decompress_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRingBytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(lz4_decompress_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRingBytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(lz4_decompress_team_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTeamSmem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(lz4_decompress_team_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTeamSmem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(lz4_decompress_team_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTeamSmem);
    if (e != cudaSuccess) return e;
    {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
    }
    if (const char* v = getenv("PLZ4CU_DEC_DUO")) g_dec_duo = atoi(v);
    if (const char* v = getenv("PLZ4CU_TEAM_PAIR")) g_team_pair = atoi(v);
    if (const char* v = getenv("PLZ4CU_TEAM_RING")) { const int r = atoi(v); if (r == 65536 || r == 131072) g_team_ring = r; }
    if (const char* v = getenv("PLZ4CU_TEAM")) g_team = atoi(v);
    if (const char* v = getenv("PLZ4CU_TEAM_DBG")) g_team_dbg = atoi(v);
    if (const char* v = getenv("PLZ4CU_TEAM_COPY")) { const int c = atoi(v); if (c == 4 || c == 8 || c == 16) g_team_copy = c; }
    return cudaFuncSetAttribute(lz4_decompress_team_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTeamSmem);
}

cudaError_t launch_decompress(const DecodeArgs& a, cudaStream_t stream)
{
    if (a.nblk == 0) return cudaSuccess;
    // (blocks beyond 32 MiB — raw-block API only — stay with one warp: a single sequence there can run for longer than the
    // team's watchdog allows its other warps to wait)
    // ... and a handful of blocks of 16-64 KiB (the per-block shims, the tail of a stream): a block's latency is what the caller
    // waits for, and a team decodes a 64 KiB block in a third of the time a warp pair needs (0.93 -> 0.32 ms per call)
    const bool few_small = a.nblk <= (uint32_t)g_sm_count && a.dst_cap >= 16384u;
    if (g_team && a.nblk < kRingBlocks && (a.dst_cap > 65536u || few_small || g_team == 2) && a.dst_cap <= (32u << 20)) {
        // few, large blocks: one CTA per block (parser warp, copy warps, checksum warp), up to 3 CTAs per SM
        const bool pair = g_team_pair < 0 ? a.nblk > (uint32_t)g_sm_count : g_team_pair != 0;
        const int ncopy = pair ? min(g_team_copy, 8) : g_team_copy;
        const int ring_bytes = pair ? 32768 : g_team_ring;
        const size_t smem = (size_t)kTeamStateBytes + (size_t)ring_bytes;
        dim3 grid(a.nblk), block((kTeamFirstCopy + ncopy) * 32);
        if (pair) {
            if (a.dict_size > 0) lz4_decompress_team_kernel<true, true><<<grid, block, smem, stream>>>(a, ring_bytes, g_team_dbg);
            else lz4_decompress_team_kernel<false, true><<<grid, block, smem, stream>>>(a, ring_bytes, g_team_dbg);
        } else {
            if (a.dict_size > 0) lz4_decompress_team_kernel<true, false><<<grid, block, smem, stream>>>(a, ring_bytes, g_team_dbg);
            else lz4_decompress_team_kernel<false, false><<<grid, block, smem, stream>>>(a, ring_bytes, g_team_dbg);
        }
        return cudaGetLastError();
    }
    if (a.nblk < kRingBlocks && a.dst_cap > 65536u) {
        // few, large blocks: one warp per CTA with a shared-memory window (up to 3 CTAs per SM)
        dim3 grid(a.nblk), block(32);
        if (a.dict_size > 0) lz4_decompress_kernel<true, true><<<grid, block, kRingBytes, stream>>>(a);
        else lz4_decompress_kernel<false, true><<<grid, block, kRingBytes, stream>>>(a);
        return cudaGetLastError();
    }
    // A launch that cannot fill the SMs with one warp per block (fewer than 32 blocks per SM) gives every block two: twice the
    // contexts to hide latency with (4096 blocks of 256 KiB: 197 -> 245 GB/s, 1024 blocks of 1 MiB: 75 -> 105); a launch
    // that can takes the one-warp kernel, which executes 13 % fewer instructions.
    if (g_dec_duo > 0 || (g_dec_duo < 0 && a.nblk <= (uint32_t)g_sm_count * 32u)) {
        const uint32_t ppb = kDecodeThreads / 64;
        dim3 grid((a.nblk + ppb - 1) / ppb), block(kDecodeThreads);
        if (a.dict_size > 0) lz4_decompress_duo_kernel<true><<<grid, block, 0, stream>>>(a);
        else lz4_decompress_duo_kernel<false><<<grid, block, 0, stream>>>(a);
        return cudaGetLastError();
    }
    const uint32_t wpb = kDecodeThreads / 32;
    dim3 grid((a.nblk + wpb - 1) / wpb), block(kDecodeThreads);
    if (a.dict_size > 0) lz4_decompress_kernel<true, false><<<grid, block, 0, stream>>>(a);
    else lz4_decompress_kernel<false, false><<<grid, block, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace plz4
