// logtext.h — deterministic synthetic log-text generator (benchmark workload).
//
// The stream is defined per 64 KiB *segment*: segment s of stream `seed` depends only on
// (seed, s), so the same bytes can be produced on the host (tests, CPU baseline) and on the
// device (one thread per segment; 8 GiB never has to cross PCIe).  Lines look like
//   "1700000123.004512 INFO [gateway] pid=417 tid=23 request id=48211 completed in 87 ms status=200\n"
// and the last line of a segment is cut at the segment boundary.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define LT_HD __host__ __device__ __forceinline__
#else
#define LT_HD static inline
#endif

#define LOGTEXT_SEG 65536u
#define LOGTEXT_DEFAULT_SEED 0x504C5A34u   /* "PLZ4" */

typedef struct lt_rng { uint64_t s; } lt_rng;

LT_HD uint64_t lt_mix(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
LT_HD uint32_t lt_next(lt_rng* r)
{
    r->s = lt_mix(r->s);
    return (uint32_t)(r->s >> 32);
}
LT_HD uint32_t lt_below(lt_rng* r, uint32_t n) { return (uint32_t)(((uint64_t)lt_next(r) * n) >> 32); }

LT_HD uint32_t lt_put_str(uint8_t* b, uint32_t p, const char* s)
{
    while (*s) b[p++] = (uint8_t)*s++;
    return p;
}
LT_HD uint32_t lt_put_uint(uint8_t* b, uint32_t p, uint32_t v, int min_digits)
{
    char tmp[10];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n < min_digits) tmp[n++] = '0';
    while (n) b[p++] = (uint8_t)tmp[--n];
    return p;
}

// Build one line into `b` (needs 192 bytes of room); returns its length.
LT_HD uint32_t lt_line(lt_rng* r, uint64_t* clock_us, uint8_t* b)
{
    const char* const levels[5] = {"INFO", "DEBUG", "WARN", "ERROR", "TRACE"};
    const char* const svcs[8] = {"auth", "api", "db", "cache", "queue", "worker", "gateway", "billing"};
    uint32_t p = 0;
    uint32_t svc = lt_below(r, 8);
    uint32_t lv = lt_below(r, 16);
    uint32_t tpl = lt_below(r, 6);
    *clock_us += 1 + lt_below(r, 900);
    p = lt_put_uint(b, p, (uint32_t)(*clock_us / 1000000ull), 10);
    b[p++] = '.';
    p = lt_put_uint(b, p, (uint32_t)(*clock_us % 1000000ull), 6);
    b[p++] = ' ';
    p = lt_put_str(b, p, levels[lv < 9 ? 0 : (lv < 12 ? 1 : (lv < 14 ? 2 : (lv < 15 ? 3 : 4)))]);
    p = lt_put_str(b, p, " [");
    p = lt_put_str(b, p, svcs[svc]);
    p = lt_put_str(b, p, "] pid=");
    p = lt_put_uint(b, p, 400 + svc * 7, 3);
    p = lt_put_str(b, p, " tid=");
    p = lt_put_uint(b, p, 1 + lt_below(r, 64), 1);
    b[p++] = ' ';
    switch (tpl) {
    case 0:
        p = lt_put_str(b, p, "request id=");
        p = lt_put_uint(b, p, lt_below(r, 65536), 1);
        p = lt_put_str(b, p, " completed in ");
        p = lt_put_uint(b, p, lt_below(r, 500), 1);
        p = lt_put_str(b, p, " ms status=");
        p = lt_put_uint(b, p, (lt_below(r, 10) < 8) ? 200u : 404u + 96u * lt_below(r, 2), 1);
        break;
    case 1:
        p = lt_put_str(b, p, "connection from 10.");
        p = lt_put_uint(b, p, lt_below(r, 4), 1);
        b[p++] = '.';
        p = lt_put_uint(b, p, lt_below(r, 256), 1);
        b[p++] = '.';
        p = lt_put_uint(b, p, lt_below(r, 256), 1);
        b[p++] = ':';
        p = lt_put_uint(b, p, 1024 + lt_below(r, 64512), 1);
        p = lt_put_str(b, p, " accepted");
        break;
    case 2:
        p = lt_put_str(b, p, "cache miss key=user:");
        p = lt_put_uint(b, p, lt_below(r, 65536), 1);
        p = lt_put_str(b, p, " shard=");
        p = lt_put_uint(b, p, lt_below(r, 32), 1);
        p = lt_put_str(b, p, " fallback=origin");
        break;
    case 3:
        p = lt_put_str(b, p, "flushed ");
        p = lt_put_uint(b, p, lt_below(r, 65536), 1);
        p = lt_put_str(b, p, " records to segment ");
        p = lt_put_uint(b, p, lt_below(r, 4096), 1);
        p = lt_put_str(b, p, " in ");
        p = lt_put_uint(b, p, lt_below(r, 65536), 1);
        p = lt_put_str(b, p, " us");
        break;
    case 4:
        p = lt_put_str(b, p, "retry attempt ");
        p = lt_put_uint(b, p, 1 + lt_below(r, 5), 1);
        p = lt_put_str(b, p, " for job ");
        p = lt_put_uint(b, p, lt_below(r, 65536), 1);
        p = lt_put_str(b, p, " backoff=");
        p = lt_put_uint(b, p, 100u << lt_below(r, 6), 1);
        p = lt_put_str(b, p, " ms");
        break;
    default:
        p = lt_put_str(b, p, "gc pause ");
        p = lt_put_uint(b, p, lt_below(r, 20000), 1);
        p = lt_put_str(b, p, " us heap=");
        p = lt_put_uint(b, p, 32768 + lt_below(r, 32768), 1);
        p = lt_put_str(b, p, " KB live=");
        p = lt_put_uint(b, p, lt_below(r, 32768), 1);
        p = lt_put_str(b, p, " KB");
        break;
    }
    {   // request-scoped trace id: 8 hex digits of fresh entropy per line
        uint32_t t = lt_next(r);
        p = lt_put_str(b, p, " trace=");
        for (int i = 0; i < 8; i++) { uint32_t d = (t >> (4 * i)) & 15u; b[p++] = (uint8_t)(d < 10 ? '0' + d : 'a' + d - 10); }
    }
    b[p++] = '\n';
    return p;
}

// Fill bytes [0, len) of segment `seg` (len <= LOGTEXT_SEG) into out.
LT_HD void lt_fill_segment(uint32_t seed, uint64_t seg, uint8_t* out, uint32_t len)
{
    lt_rng r;
    uint8_t line[192];
    uint64_t clock_us = 1700000000ull * 1000000ull + seg * 30000000ull;
    uint32_t pos = 0;
    r.s = lt_mix(((uint64_t)seed << 32) ^ seg);
    while (pos < len) {
        uint32_t n = lt_line(&r, &clock_us, line);
        uint32_t k = (len - pos < n) ? (len - pos) : n;
        for (uint32_t i = 0; i < k; i++) out[pos + i] = line[i];
        pos += k;
    }
}
