/*
 * lz4_port.c — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A from-scratch, index-based C restatement of the arithmetic that sits behind plz4's
 * independent-block hot path.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this file's library (liborc.so); the
 * product (plz4_b200/libplz4cu.so) never links, loads or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks every function below
 * byte-for-byte / return-code-for-return-code against oracle/_ref/libreflz4.so, which
 * is the reference's own vendored liblz4 1.10.0 (internal/pkg/clz4/lz4.c) compiled by
 * oracle/Makefile, and tests/test_golden.py checks it against the golden vectors in
 * the reference's own tests (G1..G10).
 *
 * What each function restates (reference file:line, relative to /root/reference):
 *   orc_compress_bound      internal/pkg/clz4/lz4.h:215            LZ4_COMPRESSBOUND
 *   orc_compress_fast       internal/pkg/clz4/lz4.c:1453,1382-1403 LZ4_compress_fast(acc=1) table choice
 *                           internal/pkg/clz4/lz4.c:930-1338       LZ4_compress_generic_validated
 *                           internal/pkg/clz4/lz4.c:777-795        LZ4_hash4 / LZ4_hash5
 *                           internal/pkg/clz4/lz4.c:680-703        LZ4_count
 *   orc_dict_*              internal/pkg/clz4/lz4.c:1587-1646      LZ4_loadDict_internal(_ld_slow)
 *                           internal/pkg/clz4/clz4.go:101-120      NewDictCtx
 *   orc_compress_dict       internal/pkg/clz4/clz4.go:160-179      StreamIndieCtx.Compress
 *                           internal/pkg/clz4/lz4.c:1658-1683      LZ4_attach_dictionary
 *                           internal/pkg/clz4/lz4.c:1707-1783      LZ4_compress_fast_continue
 *   orc_decompress_safe     internal/pkg/clz4/lz4.c:2023-2445,2451 LZ4_decompress_generic / _safe
 *   orc_decompress_dict     internal/pkg/clz4/lz4.c:2719-2732,2523 LZ4_decompress_safe_usingDict -> forceExtDict
 *   orc_xxh32               internal/pkg/xxh32/xxh32zero.go:238-280 ChecksumZero
 *   orc_block_record        internal/pkg/blk/blk.go:69-109         CompressToBlk
 *
 * Style note: liblz4 works on raw pointers with speculative wide copies; this file works
 * on byte indices with exact copies.  The *results* (bytes and return codes) are identical;
 * that is what the tests pin.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>

#define ORC_API __attribute__((visibility("default")))

enum {
    MINMATCH = 4,
    LASTLITERALS = 5,
    MFLIMIT = 12,
    MIN_INPUT_FOR_MATCH = MFLIMIT + 1,    /* lz4.c:249 LZ4_minLength */
    MAX_DISTANCE = 65535,                 /* lz4.h:674  LZ4_DISTANCE_MAX */
    LIMIT_64K = 65536 + (MFLIMIT - 1),    /* lz4.c:710  LZ4_64Klimit */
    WINDOW = 65536,
    HASH_LOG = 12,                        /* lz4.h LZ4_HASHLOG = LZ4_MEMORY_USAGE-2 */
    MAX_INPUT = 0x7E000000
};

static inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline uint32_t rd16le(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }

/* ------------------------------------------------------------------ bound */

ORC_API int orc_compress_bound(int n)
{
    if (n < 0 || (unsigned)n > (unsigned)MAX_INPUT) return 0;
    return n + n / 255 + 16;
}

/* ------------------------------------------------------------------ xxh32 */

#define XP1 2654435761u
#define XP2 2246822519u
#define XP3 3266489917u
#define XP4 668265263u
#define XP5 374761393u
static inline uint32_t rol32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

ORC_API uint32_t orc_xxh32(const uint8_t* p, size_t n)
{
    /* seed 0, xxh32zero.go:238-280 */
    size_t i = 0;
    uint32_t h;
    if (n >= 16) {
        uint32_t a = XP1 + XP2, b = XP2, c = 0, d = 0u - XP1;
        for (; i + 16 <= n; i += 16) {
            a = rol32(a + rd32(p + i) * XP2, 13) * XP1;
            b = rol32(b + rd32(p + i + 4) * XP2, 13) * XP1;
            c = rol32(c + rd32(p + i + 8) * XP2, 13) * XP1;
            d = rol32(d + rd32(p + i + 12) * XP2, 13) * XP1;
        }
        h = rol32(a, 1) + rol32(b, 7) + rol32(c, 12) + rol32(d, 18);
    } else {
        h = XP5;
    }
    h += (uint32_t)n;
    for (; i + 4 <= n; i += 4) h = rol32(h + rd32(p + i) * XP3, 17) * XP4;
    for (; i < n; i++) h = rol32(h + p[i] * XP5, 11) * XP1;
    h ^= h >> 15; h *= XP2; h ^= h >> 13; h *= XP3; h ^= h >> 16;
    return h;
}

/* ------------------------------------------------------------------ compressor */

typedef enum { TAB_U16 = 0, TAB_U32 = 1 } tab_kind_t;
typedef enum {
    HIST_NONE = 0,        /* LZ4_compress_fast: fresh ctx, indices start at 0                     */
    HIST_PREFIX_EMPTY,    /* continue() with no usable dict: withPrefix64k + dictSmall, dictSize 0 */
    HIST_EXTDICT          /* continue() with attached dict ctx (usingDictCtx == usingExtDict here)  */
} hist_kind_t;

typedef struct {
    uint8_t  bytes[WINDOW];   /* last <=64 KiB of the user dictionary (compress/dict.go:43-56) */
    uint32_t size;            /* dictSize as seen by liblz4 (0 if the dict was < 8 bytes)       */
    uint32_t raw_size;        /* bytes kept (for the decoder, which has no 8-byte floor)       */
    uint32_t table[1 << HASH_LOG];
} orc_dict_t;

static inline uint32_t hash4(uint32_t v, tab_kind_t k)
{
    return (v * 2654435761u) >> (k == TAB_U16 ? (32 - (HASH_LOG + 1)) : (32 - HASH_LOG));
}
static inline uint32_t hash5(uint64_t v)
{
    return (uint32_t)(((v << 24) * 889523592379ull) >> (64 - HASH_LOG));
}
static inline uint32_t hash_at(const uint8_t* p, tab_kind_t k)
{
    /* 64-bit build: byU16 hashes 4 bytes, everything else hashes 5 (lz4.c:797-806) */
    return (k == TAB_U16) ? hash4(rd32(p), k) : hash5(rd64(p));
}

/* common-prefix length of a[0..] and b[0..], a bounded by a_end (lz4.c:680-703) */
static uint32_t common_len(const uint8_t* a, const uint8_t* b, const uint8_t* a_end)
{
    const uint8_t* s = a;
    while (a < a_end && *a == *b) { a++; b++; }
    return (uint32_t)(a - s);
}

static int emit_len_bytes(uint8_t* dst, int op, uint32_t rest)
{
    while (rest >= 255) { dst[op++] = 255; rest -= 255; }
    dst[op++] = (uint8_t)rest;
    return op;
}

/*
 * The greedy single-probe parse.  Indices: position i of src has table index base_idx+i,
 * dict byte j (0-based inside dict->bytes[0..size)) has index 65536 - size + j.
 */
static int compress_core(const uint8_t* src, int n, uint8_t* dst, int cap, int limited,
                         tab_kind_t kind, hist_kind_t hist, const orc_dict_t* dict)
{
    uint32_t tab32[1 << HASH_LOG];
    uint16_t tab16[1 << (HASH_LOG + 1)];
    const uint32_t base_idx = (hist == HIST_NONE) ? 0u : (uint32_t)WINDOW;   /* startIndex */
    const uint32_t dsize = (hist == HIST_EXTDICT) ? dict->size : 0u;
    const uint8_t* const dbytes = (hist == HIST_EXTDICT) ? dict->bytes : NULL;
    int op = 0, anchor = 0, ip;

#define TGET(h)     ((kind == TAB_U16) ? (uint32_t)tab16[h] : tab32[h])
#define TPUT(h, v)  do { if (kind == TAB_U16) tab16[h] = (uint16_t)(v); else tab32[h] = (v); } while (0)

    if (kind == TAB_U16) memset(tab16, 0, sizeof tab16);
    else if (hist == HIST_EXTDICT) memcpy(tab32, dict->table, sizeof tab32);
    else memset(tab32, 0, sizeof tab32);

    if (n < MIN_INPUT_FOR_MATCH) goto tail;

    {
        const int mf_end = n - MFLIMIT + 1;      /* mflimitPlusOne */
        const int match_end = n - LASTLITERALS;  /* matchlimit     */
        uint32_t fwd_hash;

        TPUT(hash_at(src, kind), base_idx);
        ip = 1;
        fwd_hash = hash_at(src + ip, kind);

        for (;;) {
            /* cand_idx: table index of the candidate; in_dict: candidate lies in the dictionary */
            uint32_t cand_idx = 0;
            int in_dict = 0;
            int tok;
            /* ---- search (lz4.c:1040-1101) ---- */
            {
                int probe = ip, step = 1, tries = 1 << 6;
                for (;;) {
                    uint32_t h = fwd_hash;
                    uint32_t cur = base_idx + (uint32_t)probe;
                    cand_idx = TGET(h);
                    ip = probe;
                    probe += step;
                    step = (tries++) >> 6;
                    if (probe > mf_end) goto tail;
                    fwd_hash = hash_at(src + probe, kind);
                    TPUT(h, cur);
                    in_dict = (hist == HIST_EXTDICT) && (cand_idx < base_idx);
                    if (hist == HIST_PREFIX_EMPTY && cand_idx < base_idx) continue;     /* dictSmall */
                    if (kind != TAB_U16 && cand_idx + MAX_DISTANCE < cur) continue;     /* too far   */
                    {
                        const uint8_t* m = in_dict ? dbytes + (cand_idx - (WINDOW - dsize))
                                                   : src + (cand_idx - base_idx);
                        if (rd32(m) == rd32(src + ip)) break;
                    }
                }
            }
            /* ---- extend backwards (lz4.c:1104-1109) ---- */
            {
                uint32_t low = in_dict ? (uint32_t)(WINDOW - dsize) : base_idx;
                while (ip > anchor && cand_idx > low) {
                    uint8_t mb = in_dict ? dbytes[cand_idx - 1 - (WINDOW - dsize)]
                                         : src[cand_idx - 1 - base_idx];
                    if (src[ip - 1] != mb) break;
                    ip--; cand_idx--;
                }
            }
            /* ---- literals (lz4.c:1112-1136) ---- */
            {
                uint32_t lit = (uint32_t)(ip - anchor);
                tok = op++;
                if (limited && (int64_t)op + lit + (2 + 1 + LASTLITERALS) + lit / 255 > cap) return 0;
                if (lit >= 15) {
                    dst[tok] = 0xF0;
                    op = emit_len_bytes(dst, op, lit - 15);
                } else {
                    dst[tok] = (uint8_t)(lit << 4);
                }
                memcpy(dst + op, src + anchor, lit);
                op += (int)lit;
            }
            for (;;) {
                /* ---- offset + match length (lz4.c:1155-1226) ---- */
                uint32_t cur = base_idx + (uint32_t)ip;
                uint32_t off = cur - cand_idx;
                uint32_t mlen;   /* beyond MINMATCH */
                dst[op++] = (uint8_t)off; dst[op++] = (uint8_t)(off >> 8);
                if (in_dict) {
                    uint32_t mpos = cand_idx - (WINDOW - dsize);           /* offset inside dict bytes */
                    int lim = ip + (int)(dsize - mpos);
                    if (lim > match_end) lim = match_end;
                    mlen = common_len(src + ip + MINMATCH, dbytes + mpos + MINMATCH, src + lim);
                    ip += (int)mlen + MINMATCH;
                    if (ip == lim) {   /* ran off the end of the dictionary: continue in the block itself */
                        uint32_t more = common_len(src + lim, src, src + match_end);
                        mlen += more; ip += (int)more;
                    }
                } else {
                    mlen = common_len(src + ip + MINMATCH, src + (cand_idx - base_idx) + MINMATCH,
                                      src + match_end);
                    ip += (int)mlen + MINMATCH;
                }
                if (limited && (int64_t)op + (1 + LASTLITERALS) + (mlen + 240) / 255 > cap) return 0;
                if (mlen >= 15) {
                    dst[tok] += 15;
                    op = emit_len_bytes(dst, op, mlen - 15);
                } else {
                    dst[tok] += (uint8_t)mlen;
                }
                anchor = ip;
                if (ip >= mf_end) goto tail;
                /* ---- refill + immediate re-probe (lz4.c:1236-1295) ---- */
                TPUT(hash_at(src + ip - 2, kind), base_idx + (uint32_t)(ip - 2));
                {
                    uint32_t h = hash_at(src + ip, kind);
                    const uint8_t* m;
                    cur = base_idx + (uint32_t)ip;
                    cand_idx = TGET(h);
                    TPUT(h, cur);
                    in_dict = (hist == HIST_EXTDICT) && (cand_idx < base_idx);
                    if (hist == HIST_PREFIX_EMPTY && cand_idx < base_idx) break;
                    if (kind != TAB_U16 && cand_idx + MAX_DISTANCE < cur) break;
                    m = in_dict ? dbytes + (cand_idx - (WINDOW - dsize)) : src + (cand_idx - base_idx);
                    if (rd32(m) != rd32(src + ip)) break;
                    tok = op++;
                    dst[tok] = 0;
                }
            }
            ip++;
            fwd_hash = hash_at(src + ip, kind);
        }
    }

tail:
    /* ---- last literals (lz4.c:1302-1329) ---- */
    {
        uint32_t run = (uint32_t)(n - anchor);
        if (limited && (int64_t)op + run + 1 + (run + 255 - 15) / 255 > cap) return 0;
        if (run >= 15) {
            dst[op++] = 0xF0;
            op = emit_len_bytes(dst, op, run - 15);
        } else {
            dst[op++] = (uint8_t)(run << 4);
        }
        memcpy(dst + op, src + anchor, run);
        op += (int)run;
    }
    return op;
#undef TGET
#undef TPUT
}

/* LZ4_compress_fast(src, dst, n, cap, 1) — clz4.go:31-45 */
ORC_API int orc_compress_fast(const uint8_t* src, int n, uint8_t* dst, int cap)
{
    int limited;
    if (n < 0 || (unsigned)n > (unsigned)MAX_INPUT) return 0;
    limited = !(cap >= orc_compress_bound(n));
    if (n == 0) {
        if (limited && cap <= 0) return 0;
        dst[0] = 0;
        return 1;
    }
    return compress_core(src, n, dst, cap, limited, (n < LIMIT_64K) ? TAB_U16 : TAB_U32, HIST_NONE, NULL);
}

/* NewDictCtx(dict) — clz4.go:101-120 + LZ4_loadDictSlow (lz4.c:1587-1646) */
ORC_API orc_dict_t* orc_dict_create(const uint8_t* d, size_t n)
{
    orc_dict_t* dc = (orc_dict_t*)calloc(1, sizeof *dc);
    uint32_t first_idx, j;
    if (!dc) return NULL;
    if (n > WINDOW) { d += n - WINDOW; n = WINDOW; }
    if (n) memcpy(dc->bytes, d, n);
    dc->raw_size = (uint32_t)n;
    if (n < 8) { dc->size = 0; return dc; }     /* dictSize < HASH_UNIT: context stays empty */
    dc->size = (uint32_t)n;
    first_idx = WINDOW - dc->size;
    for (j = 0; j + 8 <= dc->size; j += 3)            /* pass 1: every third position, last wins */
        dc->table[hash5(rd64(dc->bytes + j))] = first_idx + j;
    for (j = 0; j + 8 <= dc->size; j++) {             /* pass 2: every position, only into empty slots */
        uint32_t h = hash5(rd64(dc->bytes + j));
        if (dc->table[h] == 0) dc->table[h] = first_idx + j;
    }
    return dc;
}
ORC_API void orc_dict_destroy(orc_dict_t* dc) { free(dc); }
ORC_API const uint8_t* orc_dict_bytes(const orc_dict_t* dc) { return dc->bytes; }
ORC_API uint32_t orc_dict_size(const orc_dict_t* dc) { return dc->raw_size; }

/* StreamIndieCtx.Compress — clz4.go:160-179 (resetStream_fast + attach_dictionary + fast_continue) */
ORC_API int orc_compress_dict(const orc_dict_t* dc, const uint8_t* src, int n, uint8_t* dst, int cap)
{
    if (n < 0 || (unsigned)n > (unsigned)MAX_INPUT) return 0;
    if (n == 0) {                /* continue() is always limitedOutput */
        if (cap <= 0) return 0;
        dst[0] = 0;
        return 1;
    }
    if (dc == NULL || dc->size == 0)
        return compress_core(src, n, dst, cap, 1, TAB_U32, HIST_PREFIX_EMPTY, NULL);
    return compress_core(src, n, dst, cap, 1, TAB_U32, HIST_EXTDICT, dc);
}

/* ------------------------------------------------------------------ decompressor */

/*
 * Exact accept/reject behaviour (and return codes) of LZ4_decompress_generic in
 * decode_full_block mode, for noDict and usingExtDict.  liblz4 has a fast loop and a safe
 * loop; both implement the state machine below (derivation in DESIGN.md §oracle):
 *
 *   "shortcut" sequence  : literal nibble != 15, at least 17 input bytes left after the token
 *                          and at least 32 output bytes left  -> no end-of-block checks at all;
 *                          its match is copied unchecked iff nibble != 15, offset >= 8 and the
 *                          match starts inside the block.
 *   general literal run  : last sequence iff (op+L > cap-12) or (ip+L > n-8); a last sequence
 *                          must consume the input exactly and fit the output.
 *   general match        : offset must reach real data; copies ending past cap-5 are rejected.
 *
 * dst bytes are written exactly (no wild copies), so on success dst[0..ret) is identical.
 * Divergence stated once: offset==0 is rejected here (-(ip)-1); liblz4 copies from itself and
 * produces garbage/zeroes (lz4.c:500,2407) — no valid encoder emits it and no reference test pins it.
 * The ref-vs-port fuzz test excludes streams with offset 0 for that reason.
 */
/* test hook: how many times the offset==0 divergence fired (lets the ref-vs-port fuzz skip those inputs) */
static _Thread_local uint64_t g_zero_offset_hits;
ORC_API uint64_t orc_dbg_zero_offset_hits(void) { return g_zero_offset_hits; }

static int decompress_core(const uint8_t* src, int n, uint8_t* dst, int cap,
                           const uint8_t* dict, uint32_t dsz, int ext_dict)
{
    int64_t ip = 0, op = 0;
    const int check_offset = dsz < (uint32_t)WINDOW;

    if (src == NULL || cap < 0) return -1;
    if (cap == 0) return (n == 1 && src[0] == 0) ? 0 : -1;
    if (n == 0) return -1;

    for (;;) {
        uint32_t tok = src[ip++];
        int64_t len = tok >> 4;
        uint32_t off;
        int64_t mlen;

        if (len != 15 && ip < (int64_t)n - 16 && op <= (int64_t)cap - 32) {
            /* shortcut stage 1 (lz4.c:2100-2108 / 2250-2256) */
            memcpy(dst + op, src + ip, (size_t)len);
            op += len; ip += len;
            mlen = tok & 15;
            off = rd16le(src + ip); ip += 2;
            if (mlen != 15 && off >= 8 && (int64_t)off <= op) {
                /* shortcut stage 2 (lz4.c:2148-2160 / 2266-2276) */
                int64_t k;
                mlen += MINMATCH;
                for (k = 0; k < mlen; k++) dst[op + k] = dst[op - off + k];
                op += mlen;
                continue;
            }
        } else {
            if (len == 15) {
                /* read_variable_length(ip, iend-15, initial_check=1) lz4.c:1978-2014 */
                uint32_t s;
                if (ip >= (int64_t)n - 15) return (int)(-ip - 1);
                do {
                    s = src[ip++];
                    len += s;
                    if (ip > (int64_t)n - 15) return (int)(-ip - 1);
                } while (s == 255);
            }
            if (op + len > (int64_t)cap - MFLIMIT || ip + len > (int64_t)n - (2 + 1 + LASTLITERALS)) {
                /* must be the last sequence (lz4.c:2297-2330) */
                if (ip + len != n || op + len > cap) return (int)(-ip - 1);
                memmove(dst + op, src + ip, (size_t)len);
                op += len;
                return (int)op;
            }
            memcpy(dst + op, src + ip, (size_t)len);
            op += len; ip += len;
            off = rd16le(src + ip); ip += 2;
            mlen = tok & 15;
        }
        /* general match (lz4.c:2342-2430) */
        if (mlen == 15) {
            uint32_t s;
            do {
                s = src[ip++];
                mlen += s;
                if (ip > (int64_t)n - LASTLITERALS + 1) return (int)(-ip - 1);
            } while (s == 255);
        }
        mlen += MINMATCH;
        if (check_offset && op - (int64_t)off + (int64_t)dsz < 0) return (int)(-ip - 1);
        if (off == 0) { g_zero_offset_hits++; return (int)(-ip - 1); }   /* stated divergence */
        if (ext_dict && (int64_t)off > op) {
            int64_t from_dict = (int64_t)off - op;                 /* bytes available before block start */
            int64_t k;
            if (op + mlen > (int64_t)cap - LASTLITERALS) return (int)(-ip - 1);
            if (mlen <= from_dict) {
                memmove(dst + op, dict + dsz - from_dict, (size_t)mlen);
                op += mlen;
            } else {
                int64_t rest = mlen - from_dict;
                memcpy(dst + op, dict + dsz - from_dict, (size_t)from_dict);
                op += from_dict;
                for (k = 0; k < rest; k++) dst[op + k] = dst[k];
                op += rest;
            }
            continue;
        }
        if (op + mlen > (int64_t)cap - LASTLITERALS) return (int)(-ip - 1);
        {
            int64_t k;
            for (k = 0; k < mlen; k++) dst[op + k] = dst[op - off + k];
            op += mlen;
        }
    }
}

/* LZ4_decompress_safe — clz4.go:47-60 */
ORC_API int orc_decompress_safe(const uint8_t* src, int n, uint8_t* dst, int cap)
{
    return decompress_core(src, n, dst, cap, NULL, 0, 0);
}

/* LZ4_decompress_safe_usingDict with a dictionary that is NOT contiguous with dst — clz4.go:62-78 */
ORC_API int orc_decompress_dict(const uint8_t* src, int n, uint8_t* dst, int cap,
                                const uint8_t* dict, int dsz)
{
    if (dsz == 0) return decompress_core(src, n, dst, cap, NULL, 0, 0);
    return decompress_core(src, n, dst, cap, dict, (uint32_t)dsz, 1);
}

/* ------------------------------------------------------------------ frame block record */

/*
 * blk.CompressToBlk (blk/blk.go:69-109): rec = [LE32 size | bit31 stored][payload][LE32 xxh32(payload)]
 * The compressor gets exactly bsz bytes of room; a return of 0 means "store raw".
 * `rec` must have room for bsz + 8 bytes.  Returns the record length.
 */
ORC_API int orc_block_record(const orc_dict_t* dc, const uint8_t* src, int n, int bsz,
                             int block_checksum, uint8_t* rec)
{
    int c = dc ? orc_compress_dict(dc, src, n, rec + 4, bsz) : orc_compress_fast(src, n, rec + 4, bsz);
    uint32_t word;
    if (c == 0) {
        memcpy(rec + 4, src, (size_t)n);
        c = n;
        word = (uint32_t)n | 0x80000000u;
    } else {
        word = (uint32_t)c;
    }
    rec[0] = (uint8_t)word; rec[1] = (uint8_t)(word >> 8); rec[2] = (uint8_t)(word >> 16); rec[3] = (uint8_t)(word >> 24);
    if (block_checksum) {
        uint32_t x = orc_xxh32(rec + 4, (size_t)c);
        rec[4 + c] = (uint8_t)x; rec[5 + c] = (uint8_t)(x >> 8); rec[6 + c] = (uint8_t)(x >> 16); rec[7 + c] = (uint8_t)(x >> 24);
        return c + 8;
    }
    return c + 4;
}

/* ------------------------------------------------------------------ liblz4-shaped adapters
 * Same argument order as LZ4_compress_fast / LZ4_decompress_safe so oracle/cpu_driver.c can time
 * the port through the very function-pointer types it uses for oracle/_ref/libreflz4.so. */
ORC_API int orc_lz4_compress_fast(const char* src, char* dst, int n, int cap, int accel)
{
    (void)accel;
    return orc_compress_fast((const uint8_t*)src, n, (uint8_t*)dst, cap);
}
ORC_API int orc_lz4_decompress_safe(const char* src, char* dst, int n, int cap)
{
    return orc_decompress_safe((const uint8_t*)src, n, (uint8_t*)dst, cap);
}
