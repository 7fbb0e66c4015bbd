// kernels.h — launch interfaces between the C-ABI layer (engine.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace plz4 {

constexpr int32_t PLZ4CU_E_BLOCKHASH_ = -0x7F000001;
constexpr int32_t PLZ4CU_E_OVERFLOW_ = -0x7F000002;
constexpr int32_t PLZ4CU_E_STALL_ = -0x7F000003;       // a decode team's watchdog fired (never expected)

constexpr int kDecodeThreads = 128;    // 4 warps = 4 blocks per CTA

struct DecodeArgs {
    const uint8_t* rec_base;
    const uint64_t* rec_off;
    const uint32_t* raw_len;      // raw_blocks only
    uint32_t nblk;
    uint32_t dst_cap;
    int verify_checksum;
    int raw_blocks;
    const uint8_t* dict;          // last <=64 KiB of the dictionary, or nullptr
    uint32_t dict_size;
    uint8_t* dst_base;
    uint64_t dst_stride;
    int32_t* out_len;
};
cudaError_t launch_decompress(const DecodeArgs& a, cudaStream_t stream);
cudaError_t configure_decompress();   // one-time function attributes (opt-in shared memory)

struct EncodeArgs {
    const uint8_t* src_base;
    const uint64_t* src_off;
    const uint32_t* src_len;
    uint32_t nblk;
    uint32_t dst_cap;
    int block_checksum;
    int raw_blocks;
    const uint8_t* dict;          // dictionary bytes (device) or nullptr
    uint32_t dict_size;
    const uint16_t* dict_table;   // the dictionary's table for compress_hash_bits(dst_cap) (device) or nullptr
    uint8_t* rec_base;
    uint32_t rec_stride;
    uint32_t* rec_len;
    uint32_t max_src_len;         // upper bound of src_len[] (exact on the host paths); picks the kernels and the span count
    uint32_t min_src_len;         // lower bound of src_len[] (0 = unknown): spares the launch of a kernel no block needs
    int split_by_size;            // set by launch_compress: blocks above 64 KiB belong to the span kernel, the rest to the CTA
                                  // kernel, each kernel leaving the other's blocks alone (a block's bytes never depend on its batch)
};
cudaError_t launch_compress(const EncodeArgs& a, cudaStream_t stream);
// one CTA per block of <= 64 KiB, block staged in shared memory by TMA (compress_cta.cu); no dictionary
cudaError_t launch_compress_cta(const EncodeArgs& a, cudaStream_t stream);
cudaError_t configure_compress_cta();
// blocks above 64 KiB: one CTA per span of 64 KiB fragments; the spans' streams (final literal run left out) go to tmp slots
cudaError_t launch_compress_spans(const EncodeArgs& a, uint8_t* tmp, uint32_t slot_stride, uint32_t spans_per_block, uint32_t span_bytes,
                                  int32_t* span_len, uint32_t* span_tail, cudaStream_t stream);
cudaError_t configure_compress();     // one-time function attributes (opt-in shared memory)

// dictionary table build (hash -> last position) for a given table size, device side
cudaError_t launch_dict_build(const uint8_t* dict, uint32_t dict_size, int bits, uint16_t* table, cudaStream_t stream);
int compress_hash_bits(uint32_t dst_cap);          // table size the encoder will use for this capacity

cudaError_t launch_pack(const uint8_t* rec_base, uint32_t rec_stride, const uint32_t* rec_len, uint32_t nblk,
                        uint8_t* packed, uint64_t* packed_off, cudaStream_t stream);
cudaError_t launch_scan_u32(const uint32_t* len, uint32_t n, uint64_t* off, cudaStream_t stream);   // exclusive, off[n] = total

// Block boundaries of a device-resident frame body (frame_index.cu).  Synchronises `stream`.
struct FrameIndexResult {
    uint64_t nblk;         // data blocks found before the walk stopped
    uint32_t why;          // 1 EndMark reached, 2 size word above the block size, 3 record runs past the body,
                           // 4 body ends where a size word should be
    uint64_t stop_off;     // body offset where it stopped (the EndMark when why == 1)
};
cudaError_t launch_frame_index(const uint8_t* body, uint64_t len, uint32_t bsz, int blk_check, uint64_t* rec_off, uint32_t cap,
                               FrameIndexResult* out, uint64_t* launches, cudaStream_t stream);
cudaError_t launch_xxh32(const uint8_t* base, const uint64_t* off, const uint32_t* len, uint32_t nblk,
                         uint32_t* out, cudaStream_t stream);
cudaError_t launch_gen_logtext(uint32_t seed, uint64_t first_seg, uint8_t* dst, uint64_t n, cudaStream_t stream);

}  // namespace plz4
