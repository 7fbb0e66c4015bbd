// compress.cu — LZ4 level-1 block encode, one warp per block (sm_100a).
//
// Replaces, for a whole batch of independent blocks, what plz4 does per block on a goroutine:
//   async/writer.go:232-282 compressLoop -> blk/blk.go:69-109 CompressToBlk
//     -> compress/indie.go:66-74 -> clz4.go:31-45 -> lz4.c:930-1338 LZ4_compress_generic_validated
//   plus the record framing (size word / stored fallback / xxh32 trailer) of blk/blk.go:87-106.
//
// It is NOT liblz4's serial parse.  Per 32-position group the warp
//   (a) hashes all 32 positions at once, looks every one up in a per-warp shared-memory table and
//       resolves same-group duplicates with match.any, so each position sees its most recent earlier
//       occurrence (what liblz4's mutating table gives it one position at a time), then
//   (b) walks the verified candidates greedily: first match at/after the anchor wins, is extended
//       backwards (uniform) and forwards (32 bytes per ballot), and is emitted cooperatively.
// The output is a valid LZ4 block obeying the end-of-block rules liblz4's decoder enforces
// (last match starts <= n-12 and ends <= n-5, lz4.c:245-246,963-964); bytes differ from liblz4's,
// size stays within the tolerance pinned by tests/test_compress_gpu.py.
#include "common.cuh"
#include "kernels.h"

namespace plz4 {

constexpr int kEncodeWarps = 4;                 // blocks per CTA
constexpr int kTableBytes = 16384;              // per warp: u16[8192] (n <= 64 KiB) or u32[4096]

__device__ __forceinline__ int count_equal(const uint8_t* __restrict__ src, int a, int b, int limit, int lane)
{
    // number of equal bytes src[a+i] == src[b+i], a+i < limit
    int total = 0;
    for (;;) {
        int k = a + total + lane;
        bool eq = (k < limit) && (src[k] == src[b + total + lane]);
        uint32_t ne = __ballot_sync(FULL_MASK, !eq);
        if (ne) return total + (__ffs(ne) - 1);
        total += 32;
    }
}

// bytes needed for the 255-run extension of a length whose nibble saturated
__device__ __forceinline__ int ext_bytes(int rest) { return rest / 255 + 1; }

__device__ __forceinline__ void put_ext(uint8_t* o, int rest, int lane)
{
    int nb = ext_bytes(rest);
    for (int k = lane; k < nb; k += 32) o[k] = (k == nb - 1) ? (uint8_t)(rest - 255 * (nb - 1)) : (uint8_t)255;
}

template <typename TabT, int kHashBits>
__device__ __forceinline__ int encode_block(const uint8_t* __restrict__ src, int n, uint8_t* dst,
                                            int cap, TabT* table, int lane)
{
    constexpr TabT kEmpty = (TabT)~(TabT)0;
    {
        uint4 fill = make_uint4(~0u, ~0u, ~0u, ~0u);
        uint4* t4 = reinterpret_cast<uint4*>(table);
        for (int i = lane; i < kTableBytes / 16; i += 32) t4[i] = fill;
    }
    __syncwarp();

    int op = 0, anchor = 0;
    if (n >= MFLIMIT + 1) {
        const int mf_end = n - MFLIMIT + 1;          // a match may start at p < mf_end
        const int match_end = n - LASTLITERALS;      // and must end at or before match_end
        int base = 0;
        while (base < mf_end) {
            const int p = base + lane;
            const bool valid = p < mf_end;
            uint32_t v = 0, h = 0x80000000u | (uint32_t)lane;
            int cand = -1;
            if (valid) {
                v = load_u32_unaligned(src + p);
                h = (v * 2654435761u) >> (32 - kHashBits);
                TabT c = table[h];
                if (c != kEmpty) cand = (int)c;
            }
            const uint32_t same = __match_any_sync(FULL_MASK, h);
            const uint32_t lower = same & ((1u << lane) - 1u);
            if (lower) cand = base + 31 - __clz(lower);
            if (valid && (same >> lane) == 1u) table[h] = (TabT)p;   // most recent occurrence wins
            __syncwarp();
            const bool ok = valid && cand >= 0 && (uint32_t)(p - cand) <= MAX_DISTANCE &&
                            load_u32_unaligned(src + cand) == v;
            uint32_t bal = __ballot_sync(FULL_MASK, ok);
            if (anchor > base) bal = (anchor - base >= 32) ? 0u : (bal & ~((1u << (anchor - base)) - 1u));

            while (bal) {
                const int f = __ffs(bal) - 1;
                int mpos = base + f;
                int mc = __shfl_sync(FULL_MASK, cand, f);
                while (mpos > anchor && mc > 0 && src[mpos - 1] == src[mc - 1]) { mpos--; mc--; }
                const int mlen = MINMATCH + count_equal(src, mpos + MINMATCH, mc + MINMATCH, match_end, lane);
                const int lit = mpos - anchor;
                const int mrest = mlen - MINMATCH - 15;
                const int need = 1 + (lit >= 15 ? ext_bytes(lit - 15) : 0) + lit + 2 + (mrest >= 0 ? ext_bytes(mrest) : 0);
                if (op + need > cap) return 0;
                uint8_t* o = dst + op;
                if (lane == 0) o[0] = (uint8_t)(((lit < 15 ? lit : 15) << 4) | (mrest >= 0 ? 15 : mlen - MINMATCH));
                int w = 1;
                if (lit >= 15) { put_ext(o + w, lit - 15, lane); w += ext_bytes(lit - 15); }
                warp_copy(o + w, src + anchor, (uint32_t)lit, lane);
                w += lit;
                if (lane == 0) { uint32_t off = (uint32_t)(mpos - mc); o[w] = (uint8_t)off; o[w + 1] = (uint8_t)(off >> 8); }
                w += 2;
                if (mrest >= 0) { put_ext(o + w, mrest, lane); w += ext_bytes(mrest); }
                op += w;
                anchor = mpos + mlen;
                const int d = anchor - base;
                bal = (d >= 32) ? 0u : (bal & ~((1u << d) - 1u));
            }
            base = (anchor > base + 32) ? anchor : base + 32;
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

__global__ void __launch_bounds__(kEncodeWarps * 32)
lz4_compress_kernel(EncodeArgs a)
{
    extern __shared__ __align__(16) uint8_t smem[];
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
    if (n <= 65536) c = encode_block<uint16_t, 13>(src, n, payload, (int)a.dst_cap, (uint16_t*)table, lane);
    else            c = encode_block<uint32_t, 12>(src, n, payload, (int)a.dst_cap, (uint32_t*)table, lane);

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

cudaError_t configure_compress()
{
    return cudaFuncSetAttribute(lz4_compress_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                kEncodeWarps * kTableBytes);
}

cudaError_t launch_compress(const EncodeArgs& a, cudaStream_t stream)
{
    if (a.nblk == 0) return cudaSuccess;
    dim3 grid((a.nblk + kEncodeWarps - 1) / kEncodeWarps), block(kEncodeWarps * 32);
    lz4_compress_kernel<<<grid, block, kEncodeWarps * kTableBytes, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace plz4
