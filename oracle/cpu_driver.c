/*
 * cpu_driver.c — CPU BASELINE DRIVER (test/bench infrastructure, NOT product code).
 *
 * A pthread fan-out over independent blocks that mirrors the reference's worker loops:
 *   compress:   async/writer.go:232-282 compressLoop -> blk.CompressToBlk (blk/blk.go:69-109)
 *   decompress: async/reader.go:192-221 _decompressLoop -> BlkT.Decompress (blk/blk.go:50-61)
 *               with the block-hash check of blk/frame.go:114-127 done per block.
 * The codec entry points are passed in as function pointers, so the same driver times either
 * oracle/_ref/libreflz4.so (the reference's own liblz4, kind="reference") or liborc.so (kind="port").
 * Workers pull block indices from an atomic counter, the moral equivalent of the Go channel.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdatomic.h>

#define DRV_API __attribute__((visibility("default")))

typedef int (*lz4_compress_fast_fn)(const char* src, char* dst, int n, int cap, int accel);
typedef int (*lz4_decompress_safe_fn)(const char* src, char* dst, int n, int cap);
typedef uint32_t (*xxh32_fn)(const uint8_t* p, size_t n);

typedef struct {
    /* shared, read-only */
    int mode;                       /* 0 = compress, 1 = decompress */
    lz4_compress_fast_fn cfn;
    lz4_decompress_safe_fn dfn;
    xxh32_fn xfn;
    const uint8_t* src;             /* compress: raw bytes; decompress: records at rec_off[] */
    uint64_t total;                 /* compress: total raw bytes */
    const uint64_t* rec_off;        /* decompress: record offsets */
    const uint32_t* rec_len;        /* decompress: record lengths */
    int bsz;
    int checksum;
    uint32_t nblk;
    uint8_t* dst;                   /* compress: nblk slots of (bsz+8); decompress: nblk slots of bsz */
    uint32_t* out_len;              /* compress: record length; decompress: decoded length or <0 */
    atomic_uint next;
    atomic_int errors;
} job_t;

static void* worker(void* arg)
{
    job_t* j = (job_t*)arg;
    for (;;) {
        uint32_t b = atomic_fetch_add(&j->next, 1);
        if (b >= j->nblk) break;
        if (j->mode == 0) {
            uint64_t off = (uint64_t)b * (uint64_t)j->bsz;
            int n = (int)((j->total - off < (uint64_t)j->bsz) ? (j->total - off) : (uint64_t)j->bsz);
            uint8_t* rec = j->dst + (uint64_t)b * (uint64_t)(j->bsz + 8);
            int c = j->cfn((const char*)j->src + off, (char*)rec + 4, n, j->bsz, 1);
            uint32_t word;
            if (c == 0) { memcpy(rec + 4, j->src + off, (size_t)n); c = n; word = (uint32_t)n | 0x80000000u; }
            else word = (uint32_t)c;
            memcpy(rec, &word, 4);
            if (j->checksum) { uint32_t x = j->xfn(rec + 4, (size_t)c); memcpy(rec + 4 + c, &x, 4); c += 4; }
            j->out_len[b] = (uint32_t)(c + 4);
        } else {
            const uint8_t* rec = j->src + j->rec_off[b];
            uint32_t word; int c, r;
            uint8_t* out = j->dst + (uint64_t)b * (uint64_t)j->bsz;
            memcpy(&word, rec, 4);
            c = (int)(word & 0x7FFFFFFFu);
            if (j->checksum) {
                uint32_t want; memcpy(&want, rec + 4 + c, 4);
                if (j->xfn(rec + 4, (size_t)c) != want) { j->out_len[b] = (uint32_t)-2; atomic_fetch_add(&j->errors, 1); continue; }
            }
            if (word & 0x80000000u) { memcpy(out, rec + 4, (size_t)c); r = c; }
            else r = j->dfn((const char*)rec + 4, (char*)out, c, j->bsz);
            if (r < 0) atomic_fetch_add(&j->errors, 1);
            j->out_len[b] = (uint32_t)r;
        }
    }
    return NULL;
}

static int run(job_t* j, int nthreads)
{
    pthread_t th[256];
    int i;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    atomic_init(&j->next, 0);
    atomic_init(&j->errors, 0);
    for (i = 1; i < nthreads; i++) pthread_create(&th[i], NULL, worker, j);
    worker(j);
    for (i = 1; i < nthreads; i++) pthread_join(th[i], NULL);
    return atomic_load(&j->errors);
}

/* Compress `total` bytes cut into bsz-sized blocks; records land in fixed slots of bsz+8 bytes. */
DRV_API int drv_compress_blocks(void* compress_fast, void* xxh32, const uint8_t* src, uint64_t total,
                                int bsz, int checksum, uint8_t* dst, uint32_t* rec_len, int nthreads)
{
    job_t j;
    memset(&j, 0, sizeof j);
    j.mode = 0; j.cfn = (lz4_compress_fast_fn)compress_fast; j.xfn = (xxh32_fn)xxh32;
    j.src = src; j.total = total; j.bsz = bsz; j.checksum = checksum;
    j.nblk = (uint32_t)((total + (uint64_t)bsz - 1) / (uint64_t)bsz);
    j.dst = dst; j.out_len = rec_len;
    return run(&j, nthreads);
}

/* Decode nblk records ([size][payload][xxh]) into fixed slots of bsz bytes. */
DRV_API int drv_decompress_blocks(void* decompress_safe, void* xxh32, const uint8_t* recs,
                                  const uint64_t* rec_off, const uint32_t* rec_len, uint32_t nblk,
                                  int bsz, int checksum, uint8_t* dst, uint32_t* out_len, int nthreads)
{
    job_t j;
    memset(&j, 0, sizeof j);
    j.mode = 1; j.dfn = (lz4_decompress_safe_fn)decompress_safe; j.xfn = (xxh32_fn)xxh32;
    j.src = recs; j.rec_off = rec_off; j.rec_len = rec_len; j.nblk = nblk;
    j.bsz = bsz; j.checksum = checksum; j.dst = dst; j.out_len = out_len;
    return run(&j, nthreads);
}

/* ------------------------------------------------------------------ dictionary path (BASELINE configs[3])
 * plz4.CompressBlock(src, WithBlockDictionary(d)) per payload, amortised form: the dictionary context is built
 * once (clz4.NewDictCtx, clz4.go:101-120) and every worker keeps one LZ4_stream_t, doing exactly
 * StreamIndieCtx.Compress (clz4.go:160-179): resetStream_fast + attach_dictionary + compress_fast_continue. */
typedef void (*lz4_reset_fast_fn)(void* strm);
typedef void (*lz4_attach_fn)(void* strm, const void* dict_strm);
typedef int (*lz4_continue_fn)(void* strm, const char* src, char* dst, int n, int cap, int accel);

typedef struct {
    lz4_reset_fast_fn reset; lz4_attach_fn attach; lz4_continue_fn cont;
    const void* dict_strm;
    const uint8_t* src; const uint64_t* off; uint32_t msg, nmsg, slot;
    uint8_t* dst; uint32_t* out_len;
    atomic_uint next;
} djob_t;

static void* dworker(void* arg)
{
    djob_t* j = (djob_t*)arg;
    /* LZ4_stream_t is 16416 bytes (lz4.h:719-727) */
    static __thread long long strm[(16416 + 7) / 8];
    memset(strm, 0, sizeof strm);
    for (;;) {
        uint32_t b = atomic_fetch_add(&j->next, 64);
        if (b >= j->nmsg) break;
        uint32_t e = b + 64 < j->nmsg ? b + 64 : j->nmsg;
        for (; b < e; b++) {
            j->reset(strm);
            j->attach(strm, j->dict_strm);
            j->out_len[b] = (uint32_t)j->cont(strm, (const char*)j->src + j->off[b], (char*)j->dst + (uint64_t)b * j->slot,
                                             (int)j->msg, (int)j->slot, 1);
        }
    }
    return NULL;
}

DRV_API int drv_compress_dict_msgs(void* reset_fast, void* attach, void* cont, const void* dict_strm,
                                   const uint8_t* src, const uint64_t* off, uint32_t msg_len, uint32_t nmsg,
                                   uint8_t* dst, uint32_t slot, uint32_t* out_len, int nthreads)
{
    djob_t j;
    pthread_t th[256];
    int i;
    memset(&j, 0, sizeof j);
    j.reset = (lz4_reset_fast_fn)reset_fast; j.attach = (lz4_attach_fn)attach; j.cont = (lz4_continue_fn)cont;
    j.dict_strm = dict_strm; j.src = src; j.off = off; j.msg = msg_len; j.nmsg = nmsg; j.dst = dst; j.slot = slot;
    j.out_len = out_len;
    atomic_init(&j.next, 0);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    for (i = 1; i < nthreads; i++) pthread_create(&th[i], NULL, dworker, &j);
    dworker(&j);
    for (i = 1; i < nthreads; i++) pthread_join(th[i], NULL);
    return 0;
}

/* ---- benchmark input on the host without touching the product library: the workload's definition header
 * (plz4_b200/csrc/logtext.h, plain C) compiled here, one worker per run of segments. */
#include "../plz4_b200/csrc/logtext.h"

typedef struct { uint32_t seed; uint64_t first_seg; uint8_t* dst; uint64_t n; uint64_t nseg; atomic_ullong next; } gen_t;

static void* gen_worker(void* arg)
{
    gen_t* g = (gen_t*)arg;
    for (;;) {
        uint64_t s = atomic_fetch_add(&g->next, 16);
        if (s >= g->nseg) break;
        for (uint64_t k = s; k < s + 16 && k < g->nseg; k++) {
            uint64_t pos = k * LOGTEXT_SEG;
            uint32_t len = (g->n - pos < LOGTEXT_SEG) ? (uint32_t)(g->n - pos) : LOGTEXT_SEG;
            lt_fill_segment(g->seed, g->first_seg + k, g->dst + pos, len);
        }
    }
    return 0;
}

DRV_API int drv_gen_logtext(uint32_t seed, uint64_t first_seg, uint8_t* dst, uint64_t n, int threads)
{
    gen_t g;
    pthread_t th[256];
    g.seed = seed; g.first_seg = first_seg; g.dst = dst; g.n = n; g.nseg = (n + LOGTEXT_SEG - 1) / LOGTEXT_SEG;
    atomic_init(&g.next, 0);
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    for (int i = 0; i < threads; i++) pthread_create(&th[i], 0, gen_worker, &g);
    for (int i = 0; i < threads; i++) pthread_join(th[i], 0);
    return 0;
}
