// misc.cu — small support kernels: record packing, batched xxh32, synthetic log text.
#include "common.cuh"
#include "kernels.h"
#include "logtext.h"

namespace plz4 {

// ---------------------------------------------------------------- pack records

// Exclusive prefix sum of rec_len into packed_off[0..nblk] by ONE CTA (nblk is ~1e5: microseconds).
__global__ void __launch_bounds__(1024) scan_lengths_kernel(const uint32_t* __restrict__ len, uint32_t nblk,
                                                            uint64_t* __restrict__ off)
{
    __shared__ uint64_t warp_sum[32];
    __shared__ uint64_t carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nblk; base += 1024) {
        uint32_t i = base + threadIdx.x;
        uint64_t v = (i < nblk) ? len[i] : 0;
        uint64_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint64_t t = __shfl_up_sync(FULL_MASK, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) warp_sum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint64_t s = warp_sum[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint64_t t = __shfl_up_sync(FULL_MASK, s, d);
                if (lane >= d) s += t;
            }
            warp_sum[lane] = s;
        }
        __syncthreads();
        uint64_t carry = carry_s;
        uint64_t excl = carry + (warp ? warp_sum[warp - 1] : 0) + incl - v;
        if (i < nblk) off[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_sum[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) off[nblk] = carry_s;
}

__global__ void __launch_bounds__(256) pack_records_kernel(const uint8_t* __restrict__ rec_base, uint32_t rec_stride,
                                                           const uint32_t* __restrict__ rec_len, uint32_t nblk,
                                                           uint8_t* __restrict__ packed,
                                                           const uint64_t* __restrict__ off)
{
    const int lane = threadIdx.x & 31;
    const uint32_t b = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (b >= nblk) return;
    warp_copy(packed + off[b], rec_base + (uint64_t)b * rec_stride, rec_len[b], lane);
}

cudaError_t launch_scan_u32(const uint32_t* len, uint32_t n, uint64_t* off, cudaStream_t stream)
{
    scan_lengths_kernel<<<1, 1024, 0, stream>>>(len, n, off);
    return cudaGetLastError();
}

cudaError_t launch_pack(const uint8_t* rec_base, uint32_t rec_stride, const uint32_t* rec_len, uint32_t nblk,
                        uint8_t* packed, uint64_t* packed_off, cudaStream_t stream)
{
    scan_lengths_kernel<<<1, 1024, 0, stream>>>(rec_len, nblk, packed_off);
    if (nblk) pack_records_kernel<<<(nblk + 7) / 8, 256, 0, stream>>>(rec_base, rec_stride, rec_len, nblk, packed, packed_off);
    return cudaGetLastError();
}

// ---------------------------------------------------------------- batched xxh32

__global__ void __launch_bounds__(128) xxh32_kernel(const uint8_t* __restrict__ base, const uint64_t* __restrict__ off,
                                                    const uint32_t* __restrict__ len, uint32_t nblk,
                                                    uint32_t* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const uint32_t b = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (b >= nblk) return;
    uint32_t h = warp_xxh32(base + off[b], len[b], lane);
    if (lane == 0) out[b] = h;
}

cudaError_t launch_xxh32(const uint8_t* base, const uint64_t* off, const uint32_t* len, uint32_t nblk,
                         uint32_t* out, cudaStream_t stream)
{
    if (nblk) xxh32_kernel<<<(nblk + 3) / 4, 128, 0, stream>>>(base, off, len, nblk, out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------- synthetic log text

__global__ void __launch_bounds__(64) gen_logtext_kernel(uint32_t seed, uint64_t first_seg, uint8_t* dst, uint64_t n)
{
    uint64_t s = (uint64_t)blockIdx.x * 64 + threadIdx.x;
    uint64_t begin = s * LOGTEXT_SEG;
    if (begin >= n) return;
    uint32_t len = (n - begin < LOGTEXT_SEG) ? (uint32_t)(n - begin) : LOGTEXT_SEG;
    lt_fill_segment(seed, first_seg + s, dst + begin, len);
}

cudaError_t launch_gen_logtext(uint32_t seed, uint64_t first_seg, uint8_t* dst, uint64_t n, cudaStream_t stream)
{
    uint64_t nseg = (n + LOGTEXT_SEG - 1) / LOGTEXT_SEG;
    if (nseg) gen_logtext_kernel<<<(unsigned)((nseg + 63) / 64), 64, 0, stream>>>(seed, first_seg, dst, n);
    return cudaGetLastError();
}

}  // namespace plz4
