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
// Work split inside the warp: the sequence parse is warp-uniform (every lane reads the same token /
// offset bytes, which the LSU serves as one broadcast), the byte movement is lane-parallel.  A match
// never needs lane-to-lane ordering inside itself because byte k of a match with offset `off` is
// out[op - off + (k mod off)] — always a byte that was final before the match started — so the only
// synchronisation is one __syncwarp() before a match reads what earlier sequences wrote.
#include "common.cuh"
#include "kernels.h"

namespace plz4 {

struct DecodeOut { int32_t ret; };

template <bool kDict>
__device__ __forceinline__ int32_t decode_block(const uint8_t* __restrict__ src, int n,
                                                uint8_t* dst, int cap,
                                                const uint8_t* __restrict__ dict, int dsz, int lane)
{
    if (cap == 0) return (n == 1 && src[0] == 0) ? 0 : -1;
    if (n == 0) return -1;

    int ip = 0, op = 0;
    const bool check_offset = dsz < 65536;

    for (;;) {
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
                continue;
            }
        } else {
            if (len == 15) {
                // read_variable_length(ip, iend-15, initial_check) lz4.c:1978-2014
                if (ip >= n - 15) return -ip - 1;
                uint32_t s;
                do {
                    s = src[ip++];
                    len += (int)s;
                    if (ip > n - 15) return -ip - 1;
                } while (s == 255);
            }
            if (op + len > cap - MFLIMIT || ip + len > n - (2 + 1 + LASTLITERALS)) {
                // must be the last sequence: consume the input exactly, fit the output (lz4.c:2297-2330)
                if (ip + len != n || op + len > cap) return -ip - 1;
                warp_copy(dst + op, src + ip, (uint32_t)len, lane);
                return op + len;
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
                if (ip > n - LASTLITERALS + 1) return -ip - 1;
            } while (s == 255);
        }
        mlen += MINMATCH;
        if (check_offset && op - (int)off + dsz < 0) return -ip - 1;
        if (off == 0) return -ip - 1;        // stated divergence: liblz4 would replay garbage
        if (op + mlen > cap - LASTLITERALS) return -ip - 1;

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
    }
}

template <bool kDict>
__global__ void __launch_bounds__(kDecodeThreads)
lz4_decompress_kernel(DecodeArgs a)
{
    const int lane = lane_id();
    const uint32_t b = blockIdx.x * (kDecodeThreads / 32) + (threadIdx.x >> 5);
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
        r = decode_block<kDict>(payload, (int)csize, out, (int)a.dst_cap, a.dict, (int)a.dict_size, lane);
    }
    if (lane == 0) a.out_len[b] = r;
}

cudaError_t launch_decompress(const DecodeArgs& a, cudaStream_t stream)
{
    if (a.nblk == 0) return cudaSuccess;
    const uint32_t wpb = kDecodeThreads / 32;
    dim3 grid((a.nblk + wpb - 1) / wpb), block(kDecodeThreads);
    if (a.dict_size > 0) lz4_decompress_kernel<true><<<grid, block, 0, stream>>>(a);
    else lz4_decompress_kernel<false><<<grid, block, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace plz4
