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
//       candidate in parallel — two aligned 16-byte loads of the candidate, own bytes from the
//       neighbours' registers by shuffle — giving a match length of up to 16 and a backward byte, then
//   (b) walks the measured candidates greedily with no memory access on the common path: first
//       match at/after the anchor wins; only matches longer than 16 go back to memory; short
//       sequences (the common case) leave the warp as a single predicated byte store.
// The output is a valid LZ4 block obeying the end-of-block rules liblz4's decoder enforces
// (last match starts <= n-12 and ends <= n-5, lz4.c:245-246,963-964); bytes differ from liblz4's,
// size stays within the tolerance pinned by tests/test_gpu_compress.py.
#include "common.cuh"
#include "kernels.h"

#include <cstdlib>

namespace plz4 {

constexpr int kEncodeWarps = 4;                 // blocks per CTA
__constant__ int g_lazy_dev = 1;                // 0 = greedy; k>0 = take p+1 if its match is longer by >= k

// bytes needed for the 255-run extension of a length whose nibble saturated
__device__ __forceinline__ int ext_bytes(int rest) { return rest / 255 + 1; }

__device__ __forceinline__ void put_ext(uint8_t* o, int rest, int lane)
{
    int nb = ext_bytes(rest);
    for (int k = lane; k < nb; k += 32) o[k] = (k == nb - 1) ? (uint8_t)(rest - 255 * (nb - 1)) : (uint8_t)255;
}

// Long-match tail: equal bytes of src[a..] vs src[b..] with a < limit, 32 per ballot.
__device__ __forceinline__ int count_equal(const uint8_t* __restrict__ src, int a, int b, int limit, int lane)
{
    int total = 0;
    for (;;) {
        int k = a + total + lane;
        bool eq = (k < limit) && (src[k] == src[b + total + lane]);
        uint32_t ne = __ballot_sync(FULL_MASK, !eq);
        if (ne) return total + (__ffs(ne) - 1);
        total += 32;
    }
}

// General (rare) sequence emit: literal run >= 15 or match >= 19.
__device__ __noinline__ int emit_long(uint8_t* o, const uint8_t* __restrict__ lits, int lit, uint32_t off, int mlen, int lane)
{
    const int mrest = mlen - MINMATCH - 15;
    if (lane == 0) o[0] = (uint8_t)(((lit < 15 ? lit : 15) << 4) | (mrest >= 0 ? 15 : mlen - MINMATCH));
    int w = 1;
    if (lit >= 15) { put_ext(o + w, lit - 15, lane); w += ext_bytes(lit - 15); }
    warp_copy(o + w, lits, (uint32_t)lit, lane);
    w += lit;
    if (lane == 0) { o[w] = (uint8_t)off; o[w + 1] = (uint8_t)(off >> 8); }
    w += 2;
    if (mrest >= 0) { put_ext(o + w, mrest, lane); w += ext_bytes(mrest); }
    return w;
}

template <typename TabT, int kHashBits>
__device__ __forceinline__ int encode_block(const uint8_t* __restrict__ src, int n, uint8_t* dst,
                                            int cap, TabT* table, int lane)
{
    constexpr TabT kEmpty = (TabT)~(TabT)0;
    constexpr int kProbe = 16;                       // bytes of every candidate examined in parallel
    const bool kLazy = g_lazy_dev != 0;
    const uint32_t kLazyGain = (uint32_t)g_lazy_dev - 1u;
    {
        uint4 fill = make_uint4(~0u, ~0u, ~0u, ~0u);
        uint4* t4 = reinterpret_cast<uint4*>(table);
        constexpr int kVecs = (int)(sizeof(TabT) << kHashBits) / 16;
        for (int i = lane; i < kVecs; i += 32) t4[i] = fill;
    }
    __syncwarp();

    int op = 0, anchor = 0;
    if (n >= MFLIMIT + 1) {
        const int mf_end = n - MFLIMIT + 1;          // a match may start at p < mf_end
        const int match_end = n - LASTLITERALS;      // and must end at or before match_end
        const int ld_end = n - 3;                    // 4 bytes can be read at q < ld_end
        const uintptr_t src_end = reinterpret_cast<uintptr_t>(src + n);
        int base = 0;
        uint32_t v_cur = (lane < ld_end) ? load_u32_unaligned(src + lane) : 0u;
        uint32_t v_nxt = (lane + 32 < ld_end) ? load_u32_unaligned(src + lane + 32) : 0u;
        uint32_t tail_byte = 0;                      // byte at base-1 (only meaningful when base advanced by 32)
        while (base < mf_end) {
            const int p = base + lane;
            const bool valid = p < mf_end;
            const uint32_t v = v_cur;
            // bytes two groups ahead go in flight now; they are consumed at the bottom of the loop
            const uint32_t v_far = (p + 64 < ld_end) ? load_u32_unaligned(src + p + 64) : 0u;

            // ---- (a1) hash, table lookup, same-group duplicates, table update
            uint32_t h = 0x80000000u | (uint32_t)lane;
            int cand = -1;
            if (valid) {
                h = (v * 2654435761u) >> (32 - kHashBits);
                TabT c = table[h];
                if (c != kEmpty) cand = (int)c;
            }
            const uint32_t same = __match_any_sync(FULL_MASK, h);
            const uint32_t lower = same & ((1u << lane) - 1u);
            if (lower) cand = base + 31 - __clz(lower);
            if (valid && (same >> lane) == 1u) table[h] = (TabT)p;   // most recent occurrence wins
            __syncwarp();

            // ---- (a2) every lane measures its own candidate: 16 bytes forwards, 1 byte backwards.
            // positions already covered by the previous match were inserted above but need no candidate
            const bool want = valid && p >= anchor && cand >= 0 && (uint32_t)(p - cand) <= MAX_DISTANCE;
            uint4 r0 = make_uint4(0, 0, 0, 0), r1 = make_uint4(0, 0, 0, 0);
            uint32_t cback = 0x100;                  // never equals a byte
            uint32_t t = 0;
            if (want) {
                const uintptr_t ca = reinterpret_cast<uintptr_t>(src + cand);
                const uint4* q = reinterpret_cast<const uint4*>(ca & ~uintptr_t(15));
                t = (uint32_t)ca & 15u;
                r0 = q[0];
                if (reinterpret_cast<uintptr_t>(q + 1) < src_end) r1 = q[1];
                if (cand > 0) cback = src[cand - 1];
            }
            // own bytes p+4 .. p+15 come from the neighbours' registers
            uint32_t o1, o2, o3;
            {
                const int s1 = lane + 4, s2 = lane + 8, s3 = lane + 12;
                const uint32_t a1 = __shfl_sync(FULL_MASK, v_cur, s1 & 31), b1 = __shfl_sync(FULL_MASK, v_nxt, s1 & 31);
                const uint32_t a2 = __shfl_sync(FULL_MASK, v_cur, s2 & 31), b2 = __shfl_sync(FULL_MASK, v_nxt, s2 & 31);
                const uint32_t a3 = __shfl_sync(FULL_MASK, v_cur, s3 & 31), b3 = __shfl_sync(FULL_MASK, v_nxt, s3 & 31);
                o1 = (s1 < 32) ? a1 : b1; o2 = (s2 < 32) ? a2 : b2; o3 = (s3 < 32) ? a3 : b3;
            }
            uint32_t oback = __shfl_up_sync(FULL_MASK, v_cur, 1) & 0xFFu;
            if (lane == 0) oback = tail_byte;
            const uint32_t next_tail = __shfl_sync(FULL_MASK, v_cur, 31) & 0xFFu;

            int eqlen;
            {
                // barrel-select the 5 words that hold candidate bytes t .. t+19 of the 32 fetched
                const bool w1 = (t & 4u) != 0, w2 = (t & 8u) != 0;
                const uint32_t T0 = w1 ? r0.y : r0.x, T1 = w1 ? r0.z : r0.y, T2 = w1 ? r0.w : r0.z, T3 = w1 ? r1.x : r0.w,
                               T4 = w1 ? r1.y : r1.x, T5 = w1 ? r1.z : r1.y, T6 = w1 ? r1.w : r1.z;
                const uint32_t S0 = w2 ? T2 : T0, S1 = w2 ? T3 : T1, S2 = w2 ? T4 : T2, S3 = w2 ? T5 : T3, S4 = w2 ? T6 : T4;
                const uint32_t bs = (t & 3u) * 8u;
                const uint32_t x0 = __funnelshift_r(S0, S1, bs) ^ v;
                const uint32_t x1 = __funnelshift_r(S1, S2, bs) ^ o1;
                const uint32_t x2 = __funnelshift_r(S2, S3, bs) ^ o2;
                const uint32_t x3 = __funnelshift_r(S3, S4, bs) ^ o3;
                // first differing byte: ctz(x)>>3, and ctz(0) == 32 conveniently means "all four equal"
                const int b0 = __clz(__brev(x0)) >> 3, b1 = __clz(__brev(x1)) >> 3;
                const int b2 = __clz(__brev(x2)) >> 3, b3 = __clz(__brev(x3)) >> 3;
                const int tail = (b2 < 4) ? b2 : 4 + b3;
                const int mid = (b1 < 4) ? b1 : 4 + tail;
                eqlen = (b0 < 4) ? b0 : 4 + mid;
                const int room = match_end - p;
                if (eqlen > room) eqlen = room;
            }
            const bool ok = want && eqlen >= MINMATCH;
            const bool more = ok && eqlen == kProbe && p + kProbe < match_end;
            const bool backok = ok && p > anchor && oback == cback;
            // one word per lane for the greedy walk: offset | length | flags
            const uint32_t pack = ((uint32_t)(p - cand) << 16) | ((uint32_t)eqlen << 8) | (backok ? 2u : 0u) | (more ? 1u : 0u);
            uint32_t bal = __ballot_sync(FULL_MASK, ok);

            // ---- (b) greedy walk over the verified candidates: no loads unless a match is long
            while (bal) {
                int f = __ffs(bal) - 1;
                uint32_t pk = __shfl_sync(FULL_MASK, pack, f);
                if (kLazy && f < 31) {
                    // one-step lazy parse: a strictly longer match one byte later beats this one
                    const uint32_t pk1 = __shfl_sync(FULL_MASK, pack, f + 1);
                    if (((bal >> (f + 1)) & 1u) && ((pk1 >> 8) & 0xFFu) > ((pk >> 8) & 0xFFu) + kLazyGain) { f++; pk = pk1 & ~2u; }
                }
                int mpos = base + f;
                const uint32_t off = pk >> 16;
                int mlen = (int)((pk >> 8) & 0xFFu);
                if (pk & 1u) mlen += count_equal(src, mpos + kProbe, mpos - (int)off + kProbe, match_end, lane);
                if ((pk & 2u) && mpos > anchor) {
                    mpos--; mlen++;
                    while (mpos > anchor && mpos - (int)off > 0 && src[mpos - 1] == src[mpos - (int)off - 1]) { mpos--; mlen++; }
                }
                const int lit = mpos - anchor;
                uint8_t* o = dst + op;
                if (lit < 15 && mlen < 19) {
                    // whole sequence = token + lit literals + offset <= 17 bytes: one store per lane
                    const int need = lit + 3;
                    if (op + need > cap) return 0;
                    uint32_t byte;
                    if (lane == 0) byte = (uint32_t)((lit << 4) | (mlen - MINMATCH));
                    else if (lane <= lit) byte = src[anchor + lane - 1];
                    else byte = (lane == lit + 1) ? off : (off >> 8);
                    if (lane < need) o[lane] = (uint8_t)byte;
                    op += need;
                } else {
                    const int mrest = mlen - MINMATCH - 15;
                    const int need = 1 + (lit >= 15 ? ext_bytes(lit - 15) : 0) + lit + 2 + (mrest >= 0 ? ext_bytes(mrest) : 0);
                    if (op + need > cap) return 0;
                    op += emit_long(o, src + anchor, lit, off, mlen, lane);
                }
                anchor = mpos + mlen;
                const int d = anchor - base;
                bal = (d >= 32) ? 0u : (bal & ~((1u << d) - 1u));
            }
            if (anchor > base + 32) {
                base = anchor;          // a long match skipped whole groups: their positions are not inserted
                v_cur = (base + lane < ld_end) ? load_u32_unaligned(src + base + lane) : 0u;
                v_nxt = (base + lane + 32 < ld_end) ? load_u32_unaligned(src + base + lane + 32) : 0u;
            } else {
                base += 32;
                v_cur = v_nxt;
                v_nxt = v_far;
                tail_byte = next_tail;
            }
        }
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

template <int kHashBits>
__global__ void __launch_bounds__(kEncodeWarps * 32)
lz4_compress_kernel(EncodeArgs a)
{
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int kTableBytes = 2 << kHashBits;      // u16[1<<bits] (n <= 64 KiB) or u32[1<<(bits-1)]
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    const uint32_t b = blockIdx.x * kEncodeWarps + warp;
    if (b >= a.nblk) return;

    const uint8_t* src = a.src_base + a.src_off[b];
    const int n = (int)a.src_len[b];
    uint8_t* rec = a.rec_base + (uint64_t)b * a.rec_stride;
    uint8_t* payload = a.raw_blocks ? rec : rec + 4;
    void* table = smem + warp * kTableBytes;

    int c;
    if (n <= 65536) c = encode_block<uint16_t, kHashBits>(src, n, payload, (int)a.dst_cap, (uint16_t*)table, lane);
    else            c = encode_block<uint32_t, kHashBits - 1>(src, n, payload, (int)a.dst_cap, (uint32_t*)table, lane);

    if (a.raw_blocks) {
        if (lane == 0) a.rec_len[b] = (uint32_t)c;      // 0 = does not fit (clz4.go:40-42)
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

static int g_hash_bits = 12;     // 8 KiB of table per warp: twice the resident warps of liblz4's 13 bits; the one-step
                                 // lazy parse more than pays the ratio back (profiles/r01_sweep.txt)

cudaError_t configure_compress()
{
    if (const char* e = getenv("PLZ4CU_HASH_BITS")) {          // tuning knob: 12 halves the table (more warps/SM)
        int v = atoi(e);
        if (v >= 11 && v <= 13) g_hash_bits = v;
    }
    if (const char* e = getenv("PLZ4CU_LAZY")) {
        int v = atoi(e);
        cudaError_t err = cudaMemcpyToSymbol(g_lazy_dev, &v, sizeof v);
        if (err != cudaSuccess) return err;
    }
    {
        cudaError_t err = cudaFuncSetAttribute(lz4_compress_kernel<11>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               kEncodeWarps * (2 << 11));
        if (err != cudaSuccess) return err;
    }
    cudaError_t err = cudaFuncSetAttribute(lz4_compress_kernel<13>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           kEncodeWarps * (2 << 13));
    if (err != cudaSuccess) return err;
    return cudaFuncSetAttribute(lz4_compress_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                kEncodeWarps * (2 << 12));
}

cudaError_t launch_compress(const EncodeArgs& a, cudaStream_t stream)
{
    if (a.nblk == 0) return cudaSuccess;
    dim3 grid((a.nblk + kEncodeWarps - 1) / kEncodeWarps), block(kEncodeWarps * 32);
    if (g_hash_bits == 11) lz4_compress_kernel<11><<<grid, block, kEncodeWarps * (2 << 11), stream>>>(a);
    else if (g_hash_bits == 12) lz4_compress_kernel<12><<<grid, block, kEncodeWarps * (2 << 12), stream>>>(a);
    else                   lz4_compress_kernel<13><<<grid, block, kEncodeWarps * (2 << 13), stream>>>(a);
    return cudaGetLastError();
}

}  // namespace plz4
