// decompress.cu — LZ4 block decode, one warp per block (sm_100a).
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

// ---------------------------------------------------------------- team decode: shared state (one CTA per block)
//
// A launch with few, large blocks (plz4's default 4 MiB block: 64 blocks per 256 MiB) cannot be filled by one warp per
// block, and a lone warp is bound by its own instruction latency.  The team kernel gives a block one CTA: warp 0 parses
// (steps 1 and 2 below, unchanged, so accept/reject and return codes stay those of the one-warp decoder) and publishes
// batches of up to 32 sequences into a ring of slots; kTeamCopyWarps warps produce the output, 32-byte chunk by chunk,
// chunks dealt round-robin over the warps across batch boundaries.  A chunk waits only for the output bytes its own
// matches read (`prog`: per warp, the output position below which all of that warp's chunks are complete), so chunks
// whose sources lie further back than the chunks in flight proceed in parallel; literals never wait.  One more warp
// checks the block checksum meanwhile.
constexpr int kTeamCopyWarps = 7;
constexpr int kTeamSlots = 8;                                   // batches published ahead of the slowest copy warp
constexpr int kTeamThreads = (kTeamCopyWarps + 2) * 32;         // parser + copy warps + checksum warp
constexpr int kTeamFar = 65536 - 16384;                         // matches reaching further back read global memory: the ring
                                                                // slots behind them may already belong to chunks in flight
                                                                // (kTeamSlots batches of <= 1 KiB)
constexpr int kTeamSpinLimit = 1 << 25;                         // watchdog: a stalled team reports PLZ4CU_E_STALL, it never hangs

struct TeamSlot {
    uint32_t packA[32];            // per sequence: start relative to out0 (10 bits) | literals (6 bits) | offset (16 bits)
    int litpos[32];                // per sequence: position of its first literal in the compressed block
    uint32_t bits[32];             // bitmap of sequence starts over the batch's output range
    int nseq, out0, out1, g0;      // g0: index of the batch's first chunk, modulo kTeamCopyWarps
};

struct TeamShared {
    TeamSlot slot[kTeamSlots];
    uint32_t window[32];                       // parser scratch (header lengths of the current window)
    volatile int prog[kTeamCopyWarps];
    volatile int passed[kTeamCopyWarps];       // batches a copy warp has left behind
    volatile int head;                         // batches published
    volatile int quit;                         // no batch will follow `head`
    volatile int stall;                        // watchdog fired
    volatile int hash_state;                   // 0 running, 1 checksum ok, 2 mismatch
};

// Spin until every copy warp's entry of `arr` has reached `need`.  The decision is a warp vote, so the warp stays converged.
__device__ __forceinline__ bool team_wait(TeamShared* ts, const volatile int* arr, int need, int lane)
{
    for (int spins = 0;; spins++) {
        const int v = lane < kTeamCopyWarps ? arr[lane] : 0x7FFFFFFF;
        const int st = ts->stall;
        if (__all_sync(FULL_MASK, v >= need)) break;
        if (__any_sync(FULL_MASK, st != 0) || spins > kTeamSpinLimit) {
            if (lane == 0) ts->stall = 1;
            return false;
        }
    }
    __threadfence_block();
    return true;
}
// all copy warps have left batch `need - 1` behind
__device__ __forceinline__ bool team_wait_passed(TeamShared* ts, int need, int lane) { return team_wait(ts, ts->passed, need, lane); }
// every output byte below `need` has been written
__device__ __forceinline__ bool team_wait_prog(TeamShared* ts, int need, int lane) { return team_wait(ts, ts->prog, need, lane); }

// ---------------------------------------------------------------- batched decode

// kRing: the last 64 KiB of output are mirrored in a shared-memory ring and match sources are read from there.
// Used when a launch has too few blocks to hide global-memory latency with other warps (large block sizes):
// the per-chunk round trip drops from an L2/HBM access to a shared-memory access.
// kTeam: the caller is the parser warp of a team (see above): batches are published to the copy warps instead of being
// copied here; decode_one still runs on this warp, after the copy warps have drained.
template <bool kDict, bool kRing, bool kTeam = false>
__device__ __forceinline__ int32_t decode_block(const uint8_t* __restrict__ src, int n,
                                                uint8_t* dst, int cap,
                                                const uint8_t* __restrict__ dict, int dsz, int lane, uint32_t* bitmap,
                                                uint8_t* ring, TeamShared* ts = nullptr)
{
    constexpr int kBatchBytes = 1024;               // output bytes one batch may span (32 bitmap words)
    constexpr int kMaxBatchLit = 63;                // longest literal run a batched sequence may carry (6 bits)
    if (cap == 0) return (n == 1 && src[0] == 0) ? 0 : -1;
    if (n == 0) return -1;

    int ip = 0, op = 0;
    int t_head = 0, t_g = 0;                        // kTeam: batches published, chunk index modulo the copy warps
    int t_pf = 0;                                   // kTeam: the compressed stream has been asked into L1 up to here
    const bool check_offset = dsz < 65536;
    // word-aligned view of the compressed stream: byte q of src is byte (d4 + q) of src4
    const uintptr_t sa = reinterpret_cast<uintptr_t>(src);
    const uint32_t* __restrict__ src4 = reinterpret_cast<const uint32_t*>(sa & ~uintptr_t(3));
    const uint32_t d4 = (uint32_t)sa & 3u;
    const uint32_t last4 = (d4 + (uint32_t)n - 1u) >> 2;         // last word holding a valid byte

    for (;;) {
        // ---- 1. parse up to 32 shortcut sequences; lane k latches sequence k.
        // The next 128 compressed bytes sit in registers (one word per lane).  Every lane first computes, for
        // each of its own 4 bytes, how long a sequence header starting there would be (token + literals +
        // offset [+ one length byte]); walking the token chain is then one shuffle per sequence instead of three
        // dependent loads.  Headers that do not fit the simple shape (literal nibble 15, more than one length
        // byte, too close to the end of the input) end the batch and go through decode_one.
        if constexpr (kTeam) {
            // the parse is a chain of dependent window loads: keep the stream 1-3 KiB ahead in L1 (the copy warps' literal
            // loads follow the same lines)
            if (ip + 1024 > t_pf) {
                if (t_pf < ip) t_pf = ip & ~127;
                const int at = t_pf + 128 * lane;
                if (lane < 16 && at < n) asm volatile("prefetch.global.L1 [%0];" ::"l"(src + at));
                t_pf += 2048;
            }
        }
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
            // header length if a token started at each of my 4 bytes: 3 + literals, + 1 if the match nibble is 15; a
            // literal nibble of 15 takes its extension from the byte after the token (one extension byte only)
            const uint32_t wn = __shfl_down_sync(FULL_MASK, w, 1);       // lane 31 gets junk: tokens there end the batch
            uint32_t dpack = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t tok = (w >> (8 * i)) & 0xFFu;
                const uint32_t nxt = (i < 3 ? (w >> (8 * i + 8)) : wn) & 0xFFu;
                uint32_t dl = 3u + (tok >> 4) + ((tok & 15u) == 15u ? 1u : 0u);
                if ((tok >> 4) == 15u) dl = min(dl + 1u + nxt, 255u);
                dpack |= dl << (8 * i);
            }
            // walk the chain through shared memory (one byte load per sequence): qb = byte offset inside the window
            bitmap[lane] = dpack;
            __syncwarp();
            const uint8_t* dl8 = reinterpret_cast<const uint8_t*>(bitmap);
            uint32_t qb = a0 & 3u;
            uint32_t myq = 0;
            while (nseq < 32 && qb <= 128u - 20u) {
                if (lane == nseq) myq = qb;
                qb += dl8[qb];
                nseq++;
            }
            __syncwarp();
            // each lane decodes its own header from the window
            auto wbyte = [&](uint32_t bo) -> uint32_t {        // byte at window offset bo (< 128)
                return (__shfl_sync(FULL_MASK, w, bo >> 2) >> ((bo & 3u) * 8u)) & 0xFFu;
            };
            const uint32_t tok = wbyte(myq);
            const uint32_t lext = wbyte(min(myq + 1u, 127u));
            const uint32_t L = tok >> 4, M = tok & 15u;
            const bool longlit = L == 15u;                      // literal run of 15 + one extension byte (lz4.c:2121-2128)
            const uint32_t lit = longlit ? 15u + lext : L;
            const uint32_t lp = myq + (longlit ? 2u : 1u);      // window offset of the first literal
            const bool inwin = lit <= (uint32_t)kMaxBatchLit && lp + lit + 2u < 128u;
            const uint32_t ob = inwin ? lp + lit : 0u;          // window offset of the match offset
            const uint32_t off = wbyte(ob) | (wbyte(ob + 1u) << 8);
            const uint32_t ext = wbyte(ob + 2u);
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

                const int out0 = op;
                const int out1 = __shfl_sync(FULL_MASK, o + span, nseq - 1);
                if constexpr (kTeam) {
                    // ---- 3'. publish the batch: the copy warps take it from here
                    TeamSlot& sl = ts->slot[t_head % kTeamSlots];
                    if (t_head >= kTeamSlots && !team_wait_passed(ts, t_head - kTeamSlots + 1, lane)) return PLZ4CU_E_STALL_;
                    const int orel = mine ? o - out0 : 0;
                    sl.bits[lane] = 0;
                    __syncwarp();
                    if (mine) atomicOr(&sl.bits[orel >> 5], 1u << (orel & 31));
                    sl.packA[lane] = (uint32_t)(orel & 0x3FF) | ((uint32_t)my_lit << 10) | (my_off << 16);
                    sl.litpos[lane] = my_litpos;
                    if (lane == 0) { sl.nseq = nseq; sl.out0 = out0; sl.out1 = out1; sl.g0 = t_g; }
                    __syncwarp();
                    if (lane == 0) { __threadfence_block(); ts->head = t_head + 1; }
                    t_head++;
                    t_g = (t_g + ((out1 - out0 + 31) >> 5)) % kTeamCopyWarps;
                    ip = __shfl_sync(FULL_MASK, my_ipn(), nseq - 1);
                    op = out1;
                    continue;
                }
                // ---- 3. copy: 32 output bytes per step, one per lane
                const int orel = mine ? o - out0 : 0x7FFFFFF;                 // sequences outside the batch start "never"
                // bitmap of sequence starts over the batch's output range: lane j ends up with bits out0+32j .. out0+32j+31
                bitmap[lane] = 0;
                __syncwarp();
                if (mine) atomicOr(&bitmap[orel >> 5], 1u << (orel & 31));
                __syncwarp();
                const uint32_t my_bits = bitmap[lane];
                // everything a byte needs to know about its sequence, in two words
                const uint32_t packA = (uint32_t)(orel & 0x3FF) | ((uint32_t)my_lit << 10) | (my_off << 16);
                uint8_t* dx = dst + out0 + lane;
                int j = 0;
                for (int c = 0; c < out1 - out0; c += 32, j++, dx += 32) {
                    const int xr = c + lane;                                                        // byte position relative to out0
                    // owner of the byte = last sequence starting at or before it
                    const uint32_t sbits = __shfl_sync(FULL_MASK, my_bits, j);
                    const int k = __popc(__ballot_sync(FULL_MASK, orel < c)) - 1 + __popc(sbits & ((2u << lane) - 1u));
                    const uint32_t ka = __shfl_sync(FULL_MASK, packA, k);
                    const int kp = __shfl_sync(FULL_MASK, my_litpos, k);
                    const bool live = xr < out1 - out0;
                    const int d = xr - (int)(ka & 0x3FFu);                                          // byte index inside the sequence
                    const bool is_lit = d < (int)((ka >> 10) & 63u);
                    const int sr = xr - (int)(ka >> 16);                                            // match source, relative to out0
                    const bool fwd = live && !is_lit && sr >= c;                                    // source inside this chunk
                    uint32_t val = 0;
                    if (live && !fwd) {
                        const int s = out0 + sr;
                        if (is_lit) val = src[kp + d];
                        else if (kDict && s < 0) val = dict[dsz + s];
                        else val = kRing ? ring[s & 0xFFFF] : dst[s];
                    }
                    if (__any_sync(FULL_MASK, fwd)) {
                        // forward values along in-chunk chains: root = the lane whose loaded value this byte finally equals
                        int root = fwd ? (sr - c) : lane;
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
                ip = __shfl_sync(FULL_MASK, my_ipn(), nseq - 1);
                op = out1;
                continue;
            }
        }

        // ---- anything that is not a shortcut sequence: one sequence through the literal state machine
        int32_t ret = 0;
        const int op_before = op;
        if constexpr (kTeam) {
            // decode_one reads earlier output from global memory: everything published must have been written
            if (!team_wait_passed(ts, t_head, lane)) return PLZ4CU_E_STALL_;
        }
        if (decode_one<kDict>(src, n, dst, cap, dict, dsz, lane, ip, op, ret) == kStepDone) return ret;
        __syncwarp();                                   // its stores may be the next batch's match sources
        if (kRing) {
            // decode_one works on global memory: mirror what it produced into the ring
            for (int k = max(op_before, op - 65536) + lane; k < op; k += 32) ring[k & 0xFFFF] = dst[k];
            __syncwarp();
        }
    }
}

template <bool kDict, bool kRing>
__global__ void __launch_bounds__(kDecodeThreads)
lz4_decompress_kernel(DecodeArgs a)
{
    extern __shared__ __align__(16) uint8_t dyn_smem[];              // kRing: 64 KiB per warp
    __shared__ uint32_t s_bitmap[kDecodeThreads / 32][32];
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    const uint32_t wpb = kRing ? 1u : (uint32_t)(kDecodeThreads / 32);
    if (kRing && warp != 0) return;
    const uint32_t b = blockIdx.x * wpb + warp;
    if (b >= a.nblk) return;

    const uint8_t* rec = a.rec_base + a.rec_off[b];
    uint8_t* out = a.dst_base + (uint64_t)b * a.dst_stride;
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
        if (a.verify_checksum) {                                  // blk/frame.go:114-127
            uint32_t want = load_le32(payload + csize);
            uint32_t got = warp_xxh32(payload, csize, lane);
            if (want != got) {
                if (lane == 0) a.out_len[b] = PLZ4CU_E_BLOCKHASH_;
                return;
            }
        }
    }

    int32_t r;
    if (stored) {
        warp_copy(out, payload, csize, lane);
        r = (int32_t)csize;
    } else {
        r = decode_block<kDict, kRing>(payload, (int)csize, out, (int)a.dst_cap, a.dict, (int)a.dict_size, lane,
                                       s_bitmap[warp], dyn_smem);
    }
    if (lane == 0) a.out_len[b] = r;
}

// ---------------------------------------------------------------- team decode: copy warps and kernel

// Copy warp `w` of a team: takes every published batch in order and produces the chunks dealt to it.
template <bool kDict>
__device__ __forceinline__ void team_copy(TeamShared* ts, const uint8_t* __restrict__ src, uint8_t* dst,
                                          const uint8_t* __restrict__ dict, int dsz, uint8_t* ring, int w, int lane, int dbg)
{
    int known = 0;                                  // every output byte below this is known to be written
    for (int k = 0;; k++) {
        for (int spins = 0;; spins++) {
            // lane 0 looks, everybody follows: quit is raised after the last batch is published, so it is read first
            uint32_t see = 0;
            if (lane == 0) {
                const uint32_t q = (uint32_t)ts->quit, st = (uint32_t)ts->stall;
                see = (uint32_t)ts->head | (q << 30) | (st << 31);
            }
            see = __shfl_sync(FULL_MASK, see, 0);
            if ((int)(see & 0x3FFFFFFFu) > k) break;
            if (see >> 30) return;                  // nothing more will come (or the watchdog fired)
            if (spins > kTeamSpinLimit) {
                if (lane == 0) ts->stall = 1;
                return;
            }
        }
        __threadfence_block();
        const TeamSlot& sl = ts->slot[k % kTeamSlots];
        const int nseq = sl.nseq, out0 = sl.out0, out1 = sl.out1, g0 = sl.g0;
        const uint32_t packA = sl.packA[lane];
        const int my_litpos = sl.litpos[lane];
        const uint32_t my_bits = sl.bits[lane];
        const int orel = lane < nseq ? (int)(packA & 0x3FFu) : 0x7FFFFFF;   // sequences outside the batch start "never"
        const int len = out1 - out0;
        const int nch = (len + 31) >> 5;
        int j = w - g0;                             // my first chunk of this batch
        if (j < 0) j += kTeamCopyWarps;
        __syncwarp();
        if (lane == 0) ts->prog[w] = j < nch ? out0 + 32 * j : out1;
        if (dbg & 1) j = nch;                        // measurements: the parser alone
        for (; j < nch; j += kTeamCopyWarps) {
            const int c = 32 * j;
            const int xr = c + lane;                                                        // byte position relative to out0
            const uint32_t sbits = __shfl_sync(FULL_MASK, my_bits, j);
            const int q = __popc(__ballot_sync(FULL_MASK, orel < c)) - 1 + __popc(sbits & ((2u << lane) - 1u));
            const uint32_t ka = __shfl_sync(FULL_MASK, packA, q);
            const int kp = __shfl_sync(FULL_MASK, my_litpos, q);
            const bool live = xr < len;
            const int d = xr - (int)(ka & 0x3FFu);                                          // byte index inside the sequence
            const bool is_lit = d < (int)((ka >> 10) & 63u);
            const int sr = xr - (int)(ka >> 16);                                            // match source, relative to out0
            const bool fwd = live && !is_lit && sr >= c;                                    // source inside this chunk
            const int s = out0 + sr;
            uint32_t val = 0;
            if (live && is_lit) val = src[kp + d];                                          // literals wait for nobody
            // the match sources of this chunk that other chunks produce
            const bool ext = live && !is_lit && !fwd && s >= 0;
            const int need = __reduce_max_sync(FULL_MASK, ext ? s + 1 : 0);
            if (need > known && !(dbg & 2)) {     // dbg 2: measurements, no waiting for sources
                if (!team_wait_prog(ts, need, lane)) return;
                known = need;
            }
            if (live && !is_lit && !fwd) {
                if (kDict && s < 0) val = dict[dsz + s];
                else val = (ka >> 16) > (uint32_t)kTeamFar ? dst[s] : ring[s & 0xFFFF];
            }
            if (__any_sync(FULL_MASK, fwd)) {
                int root = fwd ? (sr - c) : lane;
#pragma unroll
                for (int it = 0; it < 5; it++) root = __shfl_sync(FULL_MASK, root, root);
                val = __shfl_sync(FULL_MASK, val, root);
            }
            if (live) {
                dst[out0 + xr] = (uint8_t)val;
                ring[(out0 + xr) & 0xFFFF] = (uint8_t)val;
            }
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();
                ts->prog[w] = j + kTeamCopyWarps < nch ? out0 + 32 * (j + kTeamCopyWarps) : out1;
            }
        }
        __syncwarp();
        if (lane == 0) ts->passed[w] = k + 1;
    }
}

template <bool kDict>
__global__ void __launch_bounds__(kTeamThreads)
lz4_decompress_team_kernel(DecodeArgs a, int dbg)
{
    extern __shared__ __align__(16) uint8_t dyn_smem[];              // 64 KiB output window, then the team's state
    uint8_t* ring = dyn_smem;
    TeamShared* ts = reinterpret_cast<TeamShared*>(dyn_smem + 65536);
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    const uint32_t b = blockIdx.x;

    if (threadIdx.x < kTeamCopyWarps) { ts->prog[threadIdx.x] = 0; ts->passed[threadIdx.x] = 0; }
    if (threadIdx.x == 0) { ts->head = 0; ts->quit = 0; ts->stall = 0; ts->hash_state = 0; }
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

    if (warp == kTeamCopyWarps + 1) {
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
        if (warp <= kTeamCopyWarps) {
            const uint32_t piece = ((csize + kTeamCopyWarps) / (kTeamCopyWarps + 1) + 511u) & ~511u;
            const uint32_t lo = min(csize, (uint32_t)warp * piece), hi = min(csize, lo + piece);
            if (hi > lo) warp_copy(out + lo, payload + lo, hi - lo, lane);
        }
        __syncthreads();
        if (threadIdx.x == 0) a.out_len[b] = (verify && ts->hash_state == 2) ? PLZ4CU_E_BLOCKHASH_ : (int32_t)csize;
        return;
    }

    if (warp == 0) {
        int32_t r = decode_block<kDict, true, true>(payload, (int)csize, out, (int)a.dst_cap, a.dict, (int)a.dict_size, lane,
                                                    ts->window, ring, ts);
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
    } else {
        team_copy<kDict>(ts, payload, out, a.dict, (int)a.dict_size, ring, warp - 1, lane, dbg);
    }
}

constexpr uint32_t kRingBlocks = 1024;          // launches with fewer blocks than this use the ring kernel
constexpr int kRingBytes = 65536;
constexpr int kTeamSmem = 65536 + (int)sizeof(TeamShared);
int g_team_dbg = 0;                             // PLZ4CU_TEAM_DBG: measurement switches of the team kernel (wrong output)
int g_team = 1;                                 // PLZ4CU_TEAM=0: few large blocks go back to one warp per block (measurements);
                                                // =2: every launch below kRingBlocks blocks takes the team kernel (tests)

cudaError_t configure_decompress()
{
    cudaError_t e = cudaFuncSetAttribute(lz4_decompress_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRingBytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(lz4_decompress_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRingBytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(lz4_decompress_team_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTeamSmem);
    if (e != cudaSuccess) return e;
    if (const char* v = getenv("PLZ4CU_TEAM")) g_team = atoi(v);
    if (const char* v = getenv("PLZ4CU_TEAM_DBG")) g_team_dbg = atoi(v);
    return cudaFuncSetAttribute(lz4_decompress_team_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTeamSmem);
}

cudaError_t launch_decompress(const DecodeArgs& a, cudaStream_t stream)
{
    if (a.nblk == 0) return cudaSuccess;
    if (g_team && a.nblk < kRingBlocks && (a.dst_cap > 65536u || g_team == 2)) {
        // few, large blocks: one CTA per block (parser warp, copy warps, checksum warp), up to 3 CTAs per SM
        dim3 grid(a.nblk), block(kTeamThreads);
        if (a.dict_size > 0) lz4_decompress_team_kernel<true><<<grid, block, kTeamSmem, stream>>>(a, g_team_dbg);
        else lz4_decompress_team_kernel<false><<<grid, block, kTeamSmem, stream>>>(a, g_team_dbg);
        return cudaGetLastError();
    }
    if (a.nblk < kRingBlocks && a.dst_cap > 65536u) {
        // few, large blocks: one warp per CTA with a shared-memory window (up to 3 CTAs per SM)
        dim3 grid(a.nblk), block(32);
        if (a.dict_size > 0) lz4_decompress_kernel<true, true><<<grid, block, kRingBytes, stream>>>(a);
        else lz4_decompress_kernel<false, true><<<grid, block, kRingBytes, stream>>>(a);
        return cudaGetLastError();
    }
    const uint32_t wpb = kDecodeThreads / 32;
    dim3 grid((a.nblk + wpb - 1) / wpb), block(kDecodeThreads);
    if (a.dict_size > 0) lz4_decompress_kernel<true, false><<<grid, block, 0, stream>>>(a);
    else lz4_decompress_kernel<false, false><<<grid, block, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace plz4
