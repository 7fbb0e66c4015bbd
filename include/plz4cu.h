/*
 * plz4cu.h — C ABI of the B200-native LZ4 block engine behind plz4's independent-block path.
 *
 * This is the drop-in boundary: what a cgo shim replacing plz4's internal/pkg/clz4/clz4.go
 * (and the per-block fan-out in internal/pkg/async) binds to.  Plain pointers and sizes only.
 * Reference interface each entry point replaces is cited as file:line relative to the plz4 tree.
 *
 * Memory contract: the kernels fetch whole aligned 4- and 16-byte words.  Device input buffers (sources, records,
 * dictionaries) must therefore be readable up to the next 16-byte boundary after their last byte — true for every
 * CUDA allocation; nothing past that boundary is touched (profiles/r01_sanitizer.txt).  Device outputs are written
 * exactly (a record slot up to its record length, an output slot up to the decoded length).  The HOST decompress entry
 * points copy whole dst_cap-wide slots back: bytes of a slot beyond out_len[b], and the whole slot of a failed block,
 * are unspecified (they come from a scratch buffer the engine reuses) — read out_len[b] bytes, no more.
 * Host buffers may be page-locked (plz4cu_host_alloc, cudaHostAlloc, cudaHostRegister: copied by DMA in place, ~49 GB/s) or
 * ordinary pageable memory (staged through the engine's own pinned slabs by several host threads, ~26-32 GB/s).
 *
 * Block record layout (identical to blk.CompressToBlk, internal/pkg/blk/blk.go:87-106):
 *     [ LE32 size | bit31 = stored uncompressed ][ payload ][ LE32 xxh32(payload) if block checksum ]
 *
 * Error convention
 *   - Infrastructure failures (CUDA error, bad argument) : functions return a negative
 *     PLZ4CU_ERR_* code; plz4cu_last_error() has the text.  A shim must map these onto a
 *     non-ErrCompress error so the stream aborts (blk/blk.go:82-85) instead of storing raw.
 *   - Per-block codec results are reported per block, never as a call failure:
 *       compress   : rec_len[b] is the record length; an incompressible block is stored raw
 *                    with bit 31 set, exactly what blk.go:78-92 does on ErrCompress.
 *       decompress : out_len[b] >= 0 is the decoded size; < 0 is
 *                      -(byte offset)-1        the LZ4_decompress_safe code (clz4.go:47-60, lz4.c:2443)
 *                      PLZ4CU_E_BLOCKHASH      xxh32 mismatch          (blk/frame.go:114-127)
 *                      PLZ4CU_E_OVERFLOW       size word > block size  (blk/frame.go:79-81)
 *                      PLZ4CU_E_STALL          engine fault (a decode team's watchdog fired): map it like a
 *                                              PLZ4CU_ERR_* failure, not like corrupted data
 */
#ifndef PLZ4CU_H
#define PLZ4CU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PLZ4CU_API __attribute__((visibility("default")))
#else
#define PLZ4CU_API
#endif

#define PLZ4CU_OK                 0
#define PLZ4CU_ERR_CUDA          -1   /* a CUDA runtime call failed                              */
#define PLZ4CU_ERR_ARG           -2   /* invalid argument                                        */
#define PLZ4CU_ERR_NOMEM         -3   /* allocation failed                                       */
#define PLZ4CU_ERR_NODEVICE      -4   /* no usable GPU: the engine never falls back to the CPU   */

#define PLZ4CU_E_BLOCKHASH   ((int32_t)-0x7F000001)
#define PLZ4CU_E_OVERFLOW    ((int32_t)-0x7F000002)
#define PLZ4CU_E_STALL       ((int32_t)-0x7F000003)   /* internal: a decode team's watchdog fired; never expected */

#define PLZ4CU_STORED_BIT    0x80000000u
#define PLZ4CU_REC_OVERHEAD  8u          /* size word + checksum, blk/pool.go:15 szOverhead */
#define PLZ4CU_DICT_MAX      65536u      /* compress/dict.go:3 lz4DictSz */

typedef struct plz4cu_dict plz4cu_dict_t;
typedef void* plz4cu_stream_t;           /* a cudaStream_t; NULL = the legacy default stream */

/* ---------------------------------------------------------------- lifecycle */

/* Number of visible CUDA devices, or PLZ4CU_ERR_NODEVICE. */
PLZ4CU_API int plz4cu_device_count(void);
/* Bind the calling thread to `device` and warm the context.  0 or a negative PLZ4CU_ERR_*. */
PLZ4CU_API int plz4cu_init(int device);
/* Register the devices ONE stream may spread its batches over (whole independent blocks are sharded
 * across the GPUs of a box, no collective): warms every context and leaves the calling thread on devs[0].  A writer or
 * reader opened with opts.n_devices > 1 (or -1: all of them) then cuts every batch into contiguous runs of blocks, one
 * run per device, each on its own pipeline; the dictionary is replicated per device and the bytes produced are those of
 * the one-GPU stream.  The reference's fan-out over N workers (async/writer.go:439-467 via opts.WorkerPool.Submit,
 * opts/opts.go:43-45) is the model.  0 or a negative PLZ4CU_ERR_*. */
PLZ4CU_API int plz4cu_init_devices(int ndev, const int* devs);
/* Devices registered by plz4cu_init_devices (0 if it was never called). */
PLZ4CU_API int plz4cu_registered_devices(void);
/* Text of the last failure on this thread ("" if none). */
PLZ4CU_API const char* plz4cu_last_error(void);
/* "plz4cu <version> sm_100a" */
PLZ4CU_API const char* plz4cu_version(void);
/* Kernels launched by this process so far (bench.py's gpu_launches). */
PLZ4CU_API uint64_t plz4cu_launch_count(void);

/* ---------------------------------------------------------------- sizes */

/* compress.CompressBound (compress/compress.go:83-85) == LZ4_COMPRESSBOUND (clz4/lz4.h:215). */
PLZ4CU_API size_t plz4cu_compress_bound(size_t n);

/* ---------------------------------------------------------------- memory */

/* Pinned host slabs: what blk.BorrowBlk/ReturnBlk (blk/pool.go:35-69) hand out in the GPU build. */
PLZ4CU_API void* plz4cu_host_alloc(size_t n);
PLZ4CU_API void plz4cu_host_free(void* p);
/* Release the slabs the pool keeps cached for reuse (freed slabs are otherwise kept pinned, like a sync.Pool). */
PLZ4CU_API void plz4cu_host_trim(void);
/* Outstanding pinned slabs — the blk.CntBorrowed() leak gauge (blk/pool.go:29-33). */
PLZ4CU_API int64_t plz4cu_host_outstanding(void);
PLZ4CU_API void* plz4cu_device_alloc(size_t n);
PLZ4CU_API void plz4cu_device_free(void* p);

/* ---------------------------------------------------------------- dictionary */

/* clz4.NewDictCtx (clz4/clz4.go:101-120) + compress.NewDictT (compress/dict.go:10-16,43-56):
 * keeps the last 64 KiB of `d` on the device and builds the match table once.  Host pointer. */
PLZ4CU_API plz4cu_dict_t* plz4cu_dict_create(const void* d, size_t n);
PLZ4CU_API void plz4cu_dict_destroy(plz4cu_dict_t* dict);

/* ---------------------------------------------------------------- device-resident batches
 * All pointers below are DEVICE pointers.  Calls are asynchronous on `stream`.
 *
 * Replaces the body of async/writer.go:232-282 compressLoop -> blk.CompressToBlk (blk/blk.go:69-109)
 * -> Compressor.Compress (compress/indie.go:66-74, :27-35) for a whole batch of independent blocks.
 *   src_base + src_off[b], src_len[b]   : block b's bytes (src_len[b] <= dst_cap... see below)
 *   dst_cap                             : room given to the compressor (bsz on the frame path, blk.go:73;
 *                                         CompressBound on the raw path, plz4_block.go:105-107)
 *   rec_base + b*rec_stride             : slot receiving block b's record; rec_stride >= max(dst_cap,
 *                                         max src_len) + 8 and a multiple of 16
 *   rec_len[b]                          : record length written (4 + n [+ 4])
 *   raw_blocks != 0                     : raw block API (plz4_block.go:96-119): no size word, no checksum,
 *                                         the slot holds just the LZ4 block and rec_len[b] is its size,
 *                                         or 0 when it does not fit dst_cap (clz4.go:40-42 ErrLz4Compress)
 */
PLZ4CU_API int plz4cu_compress_batch_device(plz4cu_stream_t stream,
                                 const void* src_base, const uint64_t* src_off, const uint32_t* src_len,
                                 uint32_t nblk, uint32_t dst_cap, int block_checksum, int raw_blocks,
                                 const plz4cu_dict_t* dict,
                                 void* rec_base, uint32_t rec_stride, uint32_t* rec_len);

/* Replaces async/reader.go:192-221 _decompressLoop -> BlkT.Decompress (blk/blk.go:50-61) ->
 * Decompressor.Decompress (compress/decompress.go:32-38, :46-58), plus the block-hash check and
 * size-overflow check of blk/frame.go:79-81,114-127, for a whole batch.
 *   rec_base + rec_off[b]               : block b's record ([size][payload][xxh]) when raw_blocks == 0,
 *                                         or its bare LZ4 block of raw_len[b] bytes when raw_blocks != 0
 *   raw_len                             : only read when raw_blocks != 0
 *   dst_base + b*dst_stride, dst_cap    : output slot and capacity (bsz on the frame path, blk.go:51)
 *   out_len[b]                          : see "Error convention" above
 */
PLZ4CU_API int plz4cu_decompress_batch_device(plz4cu_stream_t stream,
                                   const void* rec_base, const uint64_t* rec_off, const uint32_t* raw_len,
                                   uint32_t nblk, uint32_t dst_cap, int verify_checksum, int raw_blocks,
                                   const plz4cu_dict_t* dict,
                                   void* dst_base, uint64_t dst_stride, int32_t* out_len);

/* Pack fixed-stride record slots into one contiguous run in block order (what writeLoop's in-order
 * wr.Write does, async/writer.go:316-348).  packed_off[b] receives each record's start (the dstMark
 * a progress callback reports, async/writer.go:329); packed_off[nblk] the total.  Device pointers. */
PLZ4CU_API int plz4cu_pack_records_device(plz4cu_stream_t stream,
                               const void* rec_base, uint32_t rec_stride, const uint32_t* rec_len,
                               uint32_t nblk, void* packed, uint64_t* packed_off);

/* ---------------------------------------------------------------- device-resident frames
 * The frame walk of blk/frame.go:54-112 for a frame BODY that already sits in device memory:
 * the offsets of all block records are found on the device, in parallel, so that a whole frame can be decoded
 * without its bytes ever visiting the host.  `body` points at the first block's size word, `len` is the number of
 * bytes available from there (it may run past the frame's end).  rec_off (device, `cap` entries) receives the
 * body-relative offset of every record, ready for plz4cu_decompress_batch_device(rec_base = body).
 * On return *nblk is the number of data blocks and *end_off the body offset just past the EndMark word.
 * Returns 0, PLZ4CU_Z_BLOCK_SIZE_OVERFLOW, PLZ4CU_Z_BLOCK_READ (a record runs past `len`), PLZ4CU_Z_BLOCK_SIZE_READ
 * (`len` ends where a size word should be), PLZ4CU_ERR_ARG (more than `cap` blocks) or PLZ4CU_ERR_CUDA.
 * Synchronises `stream`. */
PLZ4CU_API int plz4cu_frame_index_device(plz4cu_stream_t stream, const void* body, uint64_t len, uint32_t block_size,
                              int block_checksum, uint64_t* rec_off, uint32_t cap, uint64_t* nblk, uint64_t* end_off);

typedef struct plz4cu_frame_info {
    uint32_t block_size;         /* bytes (descriptor/index.go:26-38)                                   */
    uint32_t header_len;         /* bytes before the first block                                        */
    int32_t  block_checksum, content_checksum, has_content_size, has_dict_id;
    uint32_t dict_id;
    uint32_t content_hash;       /* trailer value when content_checksum != 0 (NOT verified: serial xxh32) */
    uint64_t content_size;
    uint64_t nblk;               /* data blocks                                                         */
    uint64_t frame_len;          /* bytes the frame occupies: header .. EndMark (+ content checksum)    */
    uint64_t out_bytes;          /* decoded bytes in total                                              */
    int32_t  contiguous;         /* 1: every block but the last is full, so dst holds the plain stream  */
    int32_t  reserved0;
} plz4cu_frame_info_t;

/* Decode one whole LZ4 frame from device memory to device memory: header (header/read.go:26-119, fetched with one
 * small copy), device-side frame walk, batched block decode with block-checksum verification.
 *   frame, frame_len        : device pointer to the magic number; bytes available (may extend past the frame)
 *   dst, dst_cap            : device output, slot b = dst + b*block_size; needs nblk*block_size <= dst_cap
 *   rec_off, out_len, cap   : device scratch the caller owns (cap entries each); out_len[b] as for
 *                             plz4cu_decompress_batch_device
 * Returns 0 or the first error in stream order as a PLZ4CU_Z_* code (header, block hash / decode of the blocks
 * before a broken walk, then the walk's own error; info->out_bytes counts what was decoded before it),
 * PLZ4CU_Z_UNSUPPORTED for linked blocks, PLZ4CU_ERR_ARG when cap / dst_cap are too small: info (block size, block
 * count, frame length) is filled in all the same, so a first call with cap == 0 sizes the buffers for the second.  The content checksum and content size are reported, not
 * checked.  Synchronous. */
PLZ4CU_API int plz4cu_decompress_frame_device(plz4cu_stream_t stream, const void* frame, uint64_t frame_len,
                                   const plz4cu_dict_t* dict, void* dst, uint64_t dst_cap,
                                   uint64_t* rec_off, int32_t* out_len, uint32_t cap, plz4cu_frame_info_t* info);

/* The writer's side of the same: n device-resident bytes become ONE complete LZ4 frame in device memory (header,
 * block records in order as async/writer.go:316-348 would write them, EndMark) — byte for byte what THIS library's
 * NewWriter produces from the same bytes and options, and decodable by any LZ4 frame reader.  It is NOT bit-exact with
 * plz4 / liblz4 output: the parse differs (sizes within the stated tolerance), so do not compare or deduplicate
 * compressed bytes across the two.  Honoured options: block size, block checksum, content size, dictionary id; level
 * must be 1 and blocks independent.  A content checksum is refused with PLZ4CU_Z_UNSUPPORTED: it is a serial xxh32 of
 * the whole input and stays a host-side job.  frame_cap: header + ceil(n/bsz)*(bsz+8) + 4 always suffices; returns
 * PLZ4CU_ERR_ARG when the frame does not fit.  Synchronous. */
struct plz4cu_opts;                                   /* plz4cu_opts_t, defined with the frame streams below */
PLZ4CU_API int plz4cu_compress_frame_device(plz4cu_stream_t stream, const void* src, uint64_t n, const struct plz4cu_opts* opts,
                                 const plz4cu_dict_t* dict, void* frame, uint64_t frame_cap, uint64_t* frame_len);

/* Synthetic benchmark input (SURVEY.md §8d logtext): fills n bytes of stream `seed` starting at
 * 64 KiB segment `first_seg`.  Device pointer / host pointer flavours produce identical bytes. */
PLZ4CU_API int plz4cu_gen_logtext_device(plz4cu_stream_t stream, uint32_t seed, uint64_t first_seg, void* dst, uint64_t n);
PLZ4CU_API int plz4cu_gen_logtext_host(uint32_t seed, uint64_t first_seg, void* dst, uint64_t n);

/* ---------------------------------------------------------------- host-resident batches
 * Same contracts with HOST pointers (pinned for full speed): the engine stages H2D, runs the
 * kernels and copies results back, pipelined over internal streams.  Synchronous.
 * compress : records are returned PACKED in block order in `packed` (capacity packed_cap bytes,
 *            worst case nblk*(max_len+8)); packed_off has nblk+1 entries.
 */
PLZ4CU_API int plz4cu_compress_batch_host(const void* src, const uint64_t* src_off, const uint32_t* src_len,
                               uint32_t nblk, uint32_t dst_cap, int block_checksum, int raw_blocks,
                               const plz4cu_dict_t* dict,
                               void* packed, uint64_t packed_cap, uint64_t* packed_off);
PLZ4CU_API int plz4cu_decompress_batch_host(const void* recs, uint64_t recs_bytes, const uint64_t* rec_off, const uint32_t* raw_len,
                                 uint32_t nblk, uint32_t dst_cap, int verify_checksum, int raw_blocks,
                                 const plz4cu_dict_t* dict,
                                 void* dst, uint64_t dst_stride, int32_t* out_len);

/* ---------------------------------------------------------------- per-block shims
 * Exact int-return semantics of the C symbols clz4.go binds, so the existing one-block-at-a-time
 * Compressor / Decompressor implementations keep working (batch of one; host pointers):
 *   plz4cu_compress_fast          LZ4_compress_fast(src,dst,n,cap,1)              clz4.go:31-45
 *   plz4cu_compress_fast_dict     resetStream_fast+attach_dictionary+fast_continue clz4.go:160-179
 *   plz4cu_decompress_safe        LZ4_decompress_safe                             clz4.go:47-60
 *   plz4cu_decompress_safe_dict   LZ4_decompress_safe_usingDict                   clz4.go:62-78
 * Return: compress 0 = does not fit; decompress < 0 = corrupt input.  Infrastructure failures
 * return INT32_MIN (check plz4cu_last_error()).
 */
PLZ4CU_API int plz4cu_compress_fast(const void* src, int n, void* dst, int cap);
PLZ4CU_API int plz4cu_compress_fast_dict(const plz4cu_dict_t* dict, const void* src, int n, void* dst, int cap);
PLZ4CU_API int plz4cu_decompress_safe(const void* src, int n, void* dst, int cap);
PLZ4CU_API int plz4cu_decompress_safe_dict(const plz4cu_dict_t* dict, const void* src, int n, void* dst, int cap);

/* xxh32.ChecksumZero (xxh32/xxh32zero.go:238-280) of nblk device buffers, one warp each. */
PLZ4CU_API int plz4cu_xxh32_batch_device(plz4cu_stream_t stream, const void* base, const uint64_t* off,
                              const uint32_t* len, uint32_t nblk, uint32_t* out);


/* ---------------------------------------------------------------- frame streams (host side of the path)
 *
 * C mirror of plz4.NewWriter / plz4.NewReader (plz4_writer.go:14-53, plz4_reader.go:12-33) with the
 * option set of plz4_opts.go:70-255.  The frame header / descriptor / trailer logic, block slicing,
 * Flush barrier, progress callbacks, WithReadOffset, skip frames, frame concatenation, content checksum
 * (serial xxh32 on a host core) and the sticky error state live here on the host; every block goes
 * through the batched GPU engine above.  What replaces what:
 *   writer : internal/pkg/sync/writer.go + internal/pkg/async/writer.go (batcher instead of compressLoop)
 *   reader : internal/pkg/rdr/rdr.go + internal/pkg/blk/frame.go:54-139 + internal/pkg/async/reader.go
 *   header : internal/pkg/header/{write,read,skip}.go, descriptor/*.go, trailer/trailer.go
 * Not supported by this engine (reported as PLZ4CU_Z_UNSUPPORTED, never silently degraded):
 *   levels 2-12 (lz4hc.c) and linked blocks (compress/linked.go) — the reference runs those on CPU cores.
 *
 * I/O goes through callbacks so any io.Reader / io.Writer can sit behind them:
 *   write: return bytes written (== n) or a negative value to signal an I/O error
 *   read : return bytes read (short reads allowed), 0 at end of stream, negative on I/O error
 *   seek : optional (may be NULL): skip `delta` bytes forward; return 0, or negative if it cannot seek
 *
 * Threading (mirrors the reference's sync / async packages, opts.NParallel): with n_parallel == 0 everything,
 * callbacks included, runs on the calling thread, one block per engine call.  With n_parallel != 0 the stream is
 * staged over helper threads it owns: a writer's sink and progress callbacks are invoked from one internal thread,
 * in block order, and an I/O or engine error surfaces on a later Write / Flush / Close, once (async/writer.go:
 * 175-190); a reader's source callback is invoked from one internal thread that reads and decodes one batch ahead.
 * Callbacks of one stream never run concurrently with each other; a writer's have all returned when Flush / Close
 * return, a reader's source callback may run between Read calls (read-ahead) and never after Close returned.
 * A stream object itself is not thread-safe (one caller at a time, as in the reference); distinct streams are
 * independent and may be used from different threads.  A writer starts its helper threads only once the stream
 * outgrows 8 MiB; until then (and for every small stream) it behaves like the synchronous flavour.
 */
typedef int64_t (*plz4cu_write_fn)(void* ctx, const void* data, size_t n);
typedef int64_t (*plz4cu_read_fn)(void* ctx, void* buf, size_t n);
typedef int     (*plz4cu_seek_fn)(void* ctx, int64_t delta);
/* opts.WorkerPool.Submit (internal/pkg/opts/opts.go:43-45): run task(arg) on a worker; 0 = accepted. */
typedef int     (*plz4cu_submit_fn)(void* ctx, void (*task)(void*), void* arg);
/* opts.ProgressFuncT (plz4_opts.go:113-124): (src_block_offset, dst_block_offset) at every block boundary. */
typedef void    (*plz4cu_progress_fn)(void* ctx, int64_t src_off, int64_t dst_off);
/* opts.SkipCallbackT: called with a skippable frame's payload (already read: at most sz bytes). */
typedef int     (*plz4cu_skip_fn)(void* ctx, uint8_t nibble, const void* payload, uint32_t sz);
/* opts.DictCallbackT: may return a dictionary for the dictionary id found in a frame header. */
typedef int     (*plz4cu_dict_fn)(void* ctx, uint32_t dict_id, const void** dict, size_t* dict_len);

typedef struct plz4cu_opts {
    int32_t  level;              /* WithLevel: only 1 is implemented by this engine                       */
    int32_t  n_parallel;         /* WithParallel: 0 = synchronous, one block per engine call; != 0 = staged */
    int32_t  pending_size;       /* WithPendingSize: bytes of blocks per engine call (<= 0: auto, 64 MiB;   */
                                 /*   128 MiB for >= 1 MiB blocks, readers then keep up to 7 batches decoding) */
    int32_t  block_size_idx;     /* WithBlockSize: 4..7 (64 KiB, 256 KiB, 1 MiB, 4 MiB)                    */
    int32_t  block_checksum;     /* WithBlockChecksum                                                      */
    int32_t  content_checksum;   /* WithContentChecksum                                                    */
    int32_t  block_linked;       /* WithBlockLinked: unsupported (PLZ4CU_Z_UNSUPPORTED)                    */
    int32_t  has_content_size;   /* WithContentSize                                                        */
    uint64_t content_size;
    int32_t  has_dict_id;        /* WithDictionaryId                                                       */
    uint32_t dict_id;
    const void* dict;            /* WithDictionary (last 64 KiB used)                                      */
    size_t   dict_len;
    int64_t  read_offset;        /* WithReadOffset                                                         */
    int32_t  content_size_check; /* WithContentSizeCheck                                                   */
    int32_t  n_devices;          /* 0 / 1: the calling thread's device; N > 1: the first N devices registered by   */
                                 /*   plz4cu_init_devices; -1: all of them                                          */
    plz4cu_progress_fn progress; void* progress_ctx;   /* WithProgress      */
    plz4cu_skip_fn     skip_cb;  void* skip_ctx;       /* WithSkipCallback  */
    plz4cu_dict_fn     dict_cb;  void* dict_ctx;       /* WithDictCallback  */
    plz4cu_submit_fn   submit;   void* submit_ctx;     /* WithWorkerPool (plz4_opts.go:107, opts/opts.go:43-45,97-104): the    */
                                 /*   stream's long-running stages (engine, sink / source, content hash) are handed to    */
                                 /*   submit(submit_ctx, task, arg), which must run task(arg) on some thread and may       */
                                 /*   return at once; NULL = the stream starts its own threads (opts.StubWorkerPool)      */
} plz4cu_opts_t;

/* parseOpts defaults (plz4_opts.go:238-255): level 1, parallel 1, 4 MiB blocks, content checksum on. */
PLZ4CU_API void plz4cu_opts_default(plz4cu_opts_t* o);

/* Stream error codes: the reference's sentinel errors (internal/pkg/zerr/zerr.go:11-41). */
#define PLZ4CU_Z_CLOSED               -101
#define PLZ4CU_Z_HEADER_HASH          -102   /* corrupted */
#define PLZ4CU_Z_BLOCK_HASH           -103   /* corrupted */
#define PLZ4CU_Z_CONTENT_HASH         -104   /* corrupted */
#define PLZ4CU_Z_HEADER_READ          -105
#define PLZ4CU_Z_HEADER_WRITE         -106
#define PLZ4CU_Z_MAGIC                -107   /* corrupted */
#define PLZ4CU_Z_VERSION              -108
#define PLZ4CU_Z_BLOCK_SIZE_READ      -109
#define PLZ4CU_Z_BLOCK_READ           -110
#define PLZ4CU_Z_BLOCK_SIZE_OVERFLOW  -111   /* corrupted */
#define PLZ4CU_Z_DECOMPRESS           -112   /* corrupted */
#define PLZ4CU_Z_RESERVE_BIT          -113   /* corrupted */
#define PLZ4CU_Z_BLOCK_DESCRIPTOR     -114   /* corrupted */
#define PLZ4CU_Z_CONTENT_HASH_READ    -115
#define PLZ4CU_Z_CONTENT_SIZE         -116   /* corrupted */
#define PLZ4CU_Z_READ_OFFSET          -117
#define PLZ4CU_Z_READ_OFFSET_LINKED   -118
#define PLZ4CU_Z_SKIP                 -119
#define PLZ4CU_Z_NIBBLE               -120
#define PLZ4CU_Z_UNSUPPORTED          -121
#define PLZ4CU_Z_WRITE                -122   /* the write callback failed (the Go code returns the io error itself) */
#define PLZ4CU_Z_ENGINE               -123   /* CUDA / engine failure: see plz4cu_last_error() */
/* plz4.Lz4Corrupted (plz4_err.go:43-45). */
PLZ4CU_API int plz4cu_err_corrupted(int code);
PLZ4CU_API const char* plz4cu_strerror(int code);

typedef struct plz4cu_writer plz4cu_writer_t;
typedef struct plz4cu_reader plz4cu_reader_t;

/* plz4.NewWriter (plz4_writer.go:40-53). */
PLZ4CU_API plz4cu_writer_t* plz4cu_writer_new(plz4cu_write_fn wr, void* wr_ctx, const plz4cu_opts_t* opts);
/* Writer.Write: bytes consumed (== n) or a negative PLZ4CU_Z_* code. */
PLZ4CU_API int64_t plz4cu_writer_write(plz4cu_writer_t* w, const void* src, size_t n);
/* Writer.ReadFrom: bytes consumed from `rd` until it reports end of stream, or a negative code. */
PLZ4CU_API int64_t plz4cu_writer_read_from(plz4cu_writer_t* w, plz4cu_read_fn rd, void* rd_ctx);
/* Writer.Flush: synchronous barrier, emits the pending partial block (async/writer.go:109-133). */
PLZ4CU_API int plz4cu_writer_flush(plz4cu_writer_t* w);
/* Writer.Close: flush + EndMark (+ content checksum); 0 or a negative code (async/writer.go:135-191). */
PLZ4CU_API int plz4cu_writer_close(plz4cu_writer_t* w);
PLZ4CU_API void plz4cu_writer_free(plz4cu_writer_t* w);

/* plz4.NewReader (plz4_reader.go:28-33).  Lazy: nothing is read until the first Read / WriteTo. */
PLZ4CU_API plz4cu_reader_t* plz4cu_reader_new(plz4cu_read_fn rd, plz4cu_seek_fn seek, void* rd_ctx, const plz4cu_opts_t* opts);
/* Reader.Read: > 0 bytes produced, 0 = end of stream (io.EOF), negative = PLZ4CU_Z_* (rdr/rdr.go:39-87). */
PLZ4CU_API int64_t plz4cu_reader_read(plz4cu_reader_t* r, void* dst, size_t n);
/* Reader.WriteTo: total bytes written, or a negative code (rdr/rdr.go:139-174). */
PLZ4CU_API int64_t plz4cu_reader_write_to(plz4cu_reader_t* r, plz4cu_write_fn wr, void* wr_ctx);
PLZ4CU_API int plz4cu_reader_close(plz4cu_reader_t* r);
PLZ4CU_API void plz4cu_reader_free(plz4cu_reader_t* r);

/* plz4.WriteSkipFrameHeader (plz4_writer.go:56-62, header/skip.go:18-34): 8 bytes. */
PLZ4CU_API int plz4cu_write_skip_frame_header(plz4cu_write_fn wr, void* wr_ctx, uint8_t nibble, uint32_t sz);

/* In-memory endpoints for the stream callbacks (what bytes.Reader / bytes.Buffer are to the Go tests): a membuf
 * wraps caller memory, never copies or frees it.  Reading consumes [pos, len); writing appends at len up to cap.
 * Pass the membuf as the callback context together with plz4cu_membuf_read / _write / _seek. */
typedef struct plz4cu_membuf plz4cu_membuf_t;
PLZ4CU_API plz4cu_membuf_t* plz4cu_membuf_new(void* data, size_t len, size_t cap);
PLZ4CU_API void plz4cu_membuf_free(plz4cu_membuf_t* m);
PLZ4CU_API size_t plz4cu_membuf_len(const plz4cu_membuf_t* m);
PLZ4CU_API int64_t plz4cu_membuf_read(void* ctx, void* buf, size_t n);
PLZ4CU_API int64_t plz4cu_membuf_write(void* ctx, const void* data, size_t n);
PLZ4CU_API int plz4cu_membuf_seek(void* ctx, int64_t delta);

/* header/write.go:23-73: the frame header these options produce (7..19 bytes incl. the HC byte); returns its length. */
PLZ4CU_API int plz4cu_frame_header(const plz4cu_opts_t* opts, uint8_t out[19]);

/* xxh32.ChecksumZero of a host buffer, computed on the host (header HC byte, content checksum). */
PLZ4CU_API uint32_t plz4cu_xxh32_host(const void* p, size_t n);

#ifdef __cplusplus
}
#endif
#endif /* PLZ4CU_H */
