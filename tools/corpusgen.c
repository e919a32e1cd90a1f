/*
 * corpusgen.c -- synthetic corpus generator for the BASELINE.json configs (bench + tests).
 *
 * Generates seeded content and compresses it with the system libzstd (dlopen("libzstd.so.1"),
 * hand-declared prototypes: SURVEY.md Appendix G).  libzstd is used ONLY as the generator of
 * .zst inputs -- it is never on the decode path.
 *
 * Content kinds (cg_fill):
 *   0 TEXT      Zipf(1/(rank+1)) draws over a 5 000-word vocabulary of random 2-9 letter lowercase
 *               words joined by spaces (config 2; frame i uses seed base+i)
 *   1 SKEWED    i.i.d. bytes, weight 1/(1+(v mod 64)) (config 4: Huffman literals, ~0 sequences)
 *   2 RANDOM    incompressible bytes (-> Raw blocks)
 *   3 CONSTANT  runs of one byte (-> RLE blocks)
 *   4 LONGRANGE text plus paragraphs re-emitted from up to `lr_window` bytes back (config 3)
 *   5 MIXED     random | zeros | text | random segments (config 5 piece)
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- rng ---- */
typedef struct { uint64_t s; } rng_t;
static inline uint64_t rng_next(rng_t *r) { /* splitmix64 */
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline double rng_unit(rng_t *r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }

/* ---- vocabulary ---- */
#define VOCAB 5000
static char vocab_words[VOCAB][10];
static uint8_t vocab_len[VOCAB];
static double vocab_cdf[VOCAB];
static pthread_once_t vocab_once = PTHREAD_ONCE_INIT;
static void vocab_init(void) {
    rng_t r = {0x5EEDF00Dull};
    double sum = 0;
    for (int i = 0; i < VOCAB; i++) {
        int n = 2 + (int)(rng_next(&r) % 8);
        for (int k = 0; k < n; k++) vocab_words[i][k] = (char)('a' + rng_next(&r) % 26);
        vocab_len[i] = (uint8_t)n;
        sum += 1.0 / (double)(i + 1);
        vocab_cdf[i] = sum;
    }
    for (int i = 0; i < VOCAB; i++) vocab_cdf[i] /= sum;
}
static inline int zipf_draw(rng_t *r) {
    double u = rng_unit(r);
    int lo = 0, hi = VOCAB - 1;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (vocab_cdf[mid] < u) lo = mid + 1; else hi = mid;
    }
    return lo;
}

static void fill_text(rng_t *r, uint8_t *dst, size_t n) {
    size_t p = 0;
    while (p < n) {
        int w = zipf_draw(r);
        for (int k = 0; k < vocab_len[w] && p < n; k++) dst[p++] = (uint8_t)vocab_words[w][k];
        if (p < n) dst[p++] = ' ';
    }
}

static void fill_skewed(rng_t *r, uint8_t *dst, size_t n) {
    static double cdf[256];
    static int ready = 0;
    if (!ready) { /* benign race: every thread computes the same table */
        double s = 0, t[256];
        for (int v = 0; v < 256; v++) { s += 1.0 / (double)(1 + (v % 64)); t[v] = s; }
        for (int v = 0; v < 256; v++) cdf[v] = t[v] / s;
        __atomic_store_n(&ready, 1, __ATOMIC_RELEASE);
    }
    for (size_t i = 0; i < n; i++) {
        double u = rng_unit(r);
        int lo = 0, hi = 255;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (cdf[mid] < u) lo = mid + 1; else hi = mid; }
        dst[i] = (uint8_t)lo;
    }
}

static void fill_random(rng_t *r, uint8_t *dst, size_t n) {
    size_t i = 0;
    for (; i + 8 <= n; i += 8) { uint64_t v = rng_next(r); memcpy(dst + i, &v, 8); }
    for (; i < n; i++) dst[i] = (uint8_t)rng_next(r);
}

static void fill_constant(rng_t *r, uint8_t *dst, size_t n) {
    size_t p = 0;
    while (p < n) {
        size_t run = 1000 + rng_next(r) % 400000;
        if (run > n - p) run = n - p;
        memset(dst + p, (int)(rng_next(r) & 0xFF), run);
        p += run;
    }
}

static void fill_longrange(rng_t *r, uint8_t *dst, size_t n, size_t window) {
    size_t p = 0;
    while (p < n) {
        if (p > 4096 && (rng_next(r) % 100) < 35) { /* re-emit an earlier paragraph */
            size_t back = 64 + rng_next(r) % (window < p ? window - 64 : p - 64);
            size_t len = 40 + rng_next(r) % 600;
            if (back < len) back = len;
            if (back > p) back = p;
            if (len > n - p) len = n - p;
            memmove(dst + p, dst + p - back, len);
            p += len;
        } else {
            size_t len = 200 + rng_next(r) % 800;
            if (len > n - p) len = n - p;
            fill_text(r, dst + p, len);
            p += len;
        }
    }
}

static void fill_mixed(rng_t *r, uint8_t *dst, size_t n) {
    size_t q = n / 4;
    fill_random(r, dst, q);
    memset(dst + q, 0, q);
    fill_text(r, dst + 2 * q, q);
    fill_random(r, dst + 3 * q, n - 3 * q);
}

void cg_fill(int kind, uint64_t seed, uint8_t *dst, size_t n, size_t lr_window) {
    pthread_once(&vocab_once, vocab_init);
    rng_t r = {seed * 0x9E3779B97F4A7C15ull + 0x1234567ull};
    switch (kind) {
    case 0: fill_text(&r, dst, n); break;
    case 1: fill_skewed(&r, dst, n); break;
    case 2: fill_random(&r, dst, n); break;
    case 3: fill_constant(&r, dst, n); break;
    case 4: fill_longrange(&r, dst, n, lr_window ? lr_window : (8u << 20)); break;
    default: fill_mixed(&r, dst, n); break;
    }
}

/* ---- hashing (size-independent parity property: checksum of checksums) ---- */
uint64_t cg_hash(const uint8_t *p, size_t n) {
    uint64_t h = 0x9E3779B97F4A7C15ull ^ (uint64_t)n;
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        uint64_t w;
        memcpy(&w, p + i, 8);
        h = (h ^ w) * 0xFF51AFD7ED558CCDull;
        h = (h << 31) | (h >> 33);
    }
    uint64_t tail = 0;
    for (int k = 0; i < n; i++, k++) tail |= (uint64_t)p[i] << (8 * k);
    h = (h ^ tail) * 0xC4CEB9FE1A85EC53ull;
    return h ^ (h >> 29);
}

typedef struct {
    const uint8_t *base;
    const uint64_t *off, *len;
    uint64_t *out;
    uint32_t n;
    volatile uint32_t *next;
} hash_job;
static void *hash_worker(void *arg) {
    hash_job *j = (hash_job *)arg;
    for (;;) {
        uint32_t i = __atomic_fetch_add(j->next, 64, __ATOMIC_RELAXED);
        if (i >= j->n) break;
        uint32_t e = i + 64 < j->n ? i + 64 : j->n;
        for (; i < e; i++) j->out[i] = cg_hash(j->base + j->off[i], (size_t)j->len[i]);
    }
    return NULL;
}
void cg_hash_frames(const uint8_t *base, const uint64_t *off, const uint64_t *len, uint32_t n, uint64_t *out, int nthreads) {
    volatile uint32_t next = 0;
    hash_job job = {base, off, len, out, n, &next};
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    int started = 0;
    for (int t = 0; t < nthreads; t++) if (pthread_create(&th[t], NULL, hash_worker, &job) == 0) started++; else break;
    if (!started) hash_worker(&job);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
}

/* ---- libzstd via dlopen ---- */
typedef struct { const void *src; size_t size; size_t pos; } zin_t;
typedef struct { void *dst; size_t size; size_t pos; } zout_t;
static struct {
    void *h;
    size_t (*compressBound)(size_t);
    unsigned (*isError)(size_t);
    void *(*createCCtx)(void);
    size_t (*freeCCtx)(void *);
    size_t (*setParameter)(void *, int, int);
    size_t (*compress2)(void *, void *, size_t, const void *, size_t);
    size_t (*compressStream2)(void *, zout_t *, zin_t *, int);
    size_t (*decompress)(void *, size_t, const void *, size_t);
    unsigned (*versionNumber)(void);
    size_t (*compressUsingDict)(void *, void *, size_t, const void *, size_t, const void *, size_t, int);
    size_t (*decompressUsingDict)(void *, void *, size_t, const void *, size_t, const void *, size_t);
    void *(*createDCtx)(void);
    size_t (*freeDCtx)(void *);
    size_t (*trainFromBuffer)(void *, size_t, const void *, const size_t *, unsigned);
    unsigned (*zdictIsError)(size_t);
} Z;
static pthread_once_t z_once = PTHREAD_ONCE_INIT;
static void z_init(void) {
    Z.h = dlopen("libzstd.so.1", RTLD_NOW);
    if (!Z.h) return;
    Z.compressBound = dlsym(Z.h, "ZSTD_compressBound");
    Z.isError = dlsym(Z.h, "ZSTD_isError");
    Z.createCCtx = dlsym(Z.h, "ZSTD_createCCtx");
    Z.freeCCtx = dlsym(Z.h, "ZSTD_freeCCtx");
    Z.setParameter = dlsym(Z.h, "ZSTD_CCtx_setParameter");
    Z.compress2 = dlsym(Z.h, "ZSTD_compress2");
    Z.compressStream2 = dlsym(Z.h, "ZSTD_compressStream2");
    Z.decompress = dlsym(Z.h, "ZSTD_decompress");
    Z.versionNumber = dlsym(Z.h, "ZSTD_versionNumber");
    Z.compressUsingDict = dlsym(Z.h, "ZSTD_compress_usingDict");
    Z.decompressUsingDict = dlsym(Z.h, "ZSTD_decompress_usingDict");
    Z.createDCtx = dlsym(Z.h, "ZSTD_createDCtx");
    Z.freeDCtx = dlsym(Z.h, "ZSTD_freeDCtx");
    Z.trainFromBuffer = dlsym(Z.h, "ZDICT_trainFromBuffer");
    Z.zdictIsError = dlsym(Z.h, "ZDICT_isError");
}
int cg_zstd_available(void) {
    pthread_once(&z_once, z_init);
    return Z.h && Z.compress2 && Z.compressStream2 ? (int)Z.versionNumber() : 0;
}
size_t cg_compress_bound(size_t n) {
    pthread_once(&z_once, z_init);
    return Z.compressBound ? Z.compressBound(n) : 0;
}

/* One-shot compress of a caller buffer (level, checksum flag).  Returns size or 0 on error. */
size_t cg_compress(const uint8_t *src, size_t n, uint8_t *dst, size_t cap, int level, int checksum) {
    if (!cg_zstd_available()) return 0;
    void *c = Z.createCCtx();
    Z.setParameter(c, 100, level);
    Z.setParameter(c, 201, checksum);
    size_t r = Z.compress2(c, dst, cap, src, n);
    Z.freeCCtx(c);
    return Z.isError(r) ? 0 : r;
}

/* Streaming compress without a pledged size: window descriptor, no Frame_Content_Size
 * (config 3; SURVEY.md Appendix C caveats).  Returns size or 0. */
size_t cg_compress_stream(const uint8_t *src, size_t n, uint8_t *dst, size_t cap, int level, int window_log, int checksum) {
    if (!cg_zstd_available()) return 0;
    void *c = Z.createCCtx();
    Z.setParameter(c, 100, level);
    if (window_log) Z.setParameter(c, 101, window_log);
    Z.setParameter(c, 201, checksum);
    zout_t o = {dst, cap, 0};
    size_t chunk = 1u << 20, p = 0;
    int ok = 1;
    while (p < n && ok) {
        size_t take = n - p < chunk ? n - p : chunk;
        zin_t in = {src + p, take, 0};
        while (in.pos < in.size) {
            size_t r = Z.compressStream2(c, &o, &in, 0);
            if (Z.isError(r) || o.pos == o.size) { ok = 0; break; }
        }
        p += take;
    }
    if (ok) {
        zin_t in = {src, 0, 0};
        for (;;) {
            size_t r = Z.compressStream2(c, &o, &in, 2);
            if (Z.isError(r)) { ok = 0; break; }
            if (r == 0) break;
            if (o.pos == o.size) { ok = 0; break; }
        }
    }
    Z.freeCCtx(c);
    return ok ? o.pos : 0;
}

/* ---- dictionaries: inputs for the dictionary tests only ---- */
/* One-shot compress WITH a dictionary (raw content, or a formatted one from cg_train_dictionary).  Returns size or 0. */
size_t cg_compress_dict(const uint8_t *src, size_t n, const uint8_t *dict, size_t dict_len, uint8_t *dst, size_t cap, int level) {
    if (!cg_zstd_available() || !Z.compressUsingDict) return 0;
    void *c = Z.createCCtx();
    size_t r = Z.compressUsingDict(c, dst, cap, src, n, dict, dict_len, level);
    Z.freeCCtx(c);
    return Z.isError(r) ? 0 : r;
}
/* libzstd's own decoder with the dictionary: the known answer the oracle's dictionary extension is pinned with. */
size_t cg_zstd_decompress_dict(const uint8_t *src, size_t n, const uint8_t *dict, size_t dict_len, uint8_t *dst, size_t cap) {
    if (!cg_zstd_available() || !Z.decompressUsingDict) return (size_t)-1;
    void *d = Z.createDCtx();
    size_t r = Z.decompressUsingDict(d, dst, cap, src, n, dict, dict_len);
    Z.freeDCtx(d);
    return Z.isError(r) ? (size_t)-1 : r;
}
/* ZDICT_trainFromBuffer: a FORMATTED dictionary (magic, id, entropy tables, repeat offsets, content) from samples laid out
 * back to back.  Returns its size or 0. */
size_t cg_train_dictionary(uint8_t *dict, size_t cap, const uint8_t *samples, const size_t *sizes, unsigned nsamples) {
    if (!cg_zstd_available() || !Z.trainFromBuffer) return 0;
    size_t r = Z.trainFromBuffer(dict, cap, samples, sizes, nsamples);
    return Z.zdictIsError(r) ? 0 : r;
}

size_t cg_zstd_decompress(const uint8_t *src, size_t n, uint8_t *dst, size_t cap) {
    if (!cg_zstd_available()) return (size_t)-1;
    size_t r = Z.decompress(dst, cap, src, n);
    return Z.isError(r) ? (size_t)-1 : r;
}

/* libzstd over a batch of independent frames, one frame per task, nthreads workers; each worker decodes into its own
 * scratch buffer (the output is discarded, as in the reference arm).  Returns the number of frames that failed.
 * Only used by bench.py for the "industrial CPU" line next to the reference port. */
typedef struct {
    const uint8_t *base;
    const uint64_t *off, *len, *raw;
    uint32_t n;
    volatile uint32_t *next;
    volatile uint32_t *failed;
} zdec_job;
static void *zdec_worker(void *arg) {
    zdec_job *j = (zdec_job *)arg;
    size_t cap = 1 << 16;
    uint8_t *buf = (uint8_t *)malloc(cap);
    for (;;) {
        uint32_t i = __atomic_fetch_add(j->next, 16, __ATOMIC_RELAXED);
        if (i >= j->n || !buf) break;
        uint32_t e = i + 16 < j->n ? i + 16 : j->n;
        for (; i < e; i++) {
            size_t need = (size_t)j->raw[i] + 64;
            if (need > cap) {
                free(buf);
                cap = need;
                buf = (uint8_t *)malloc(cap);
                if (!buf) break;
            }
            size_t r = Z.decompress(buf, cap, j->base + j->off[i], (size_t)j->len[i]);
            if (Z.isError(r) || r != (size_t)j->raw[i]) __atomic_fetch_add(j->failed, 1, __ATOMIC_RELAXED);
        }
    }
    free(buf);
    return NULL;
}
uint32_t cg_zstd_decompress_batch_mt(const uint8_t *base, const uint64_t *off, const uint64_t *len, const uint64_t *raw, uint32_t n,
                                     int nthreads) {
    if (!cg_zstd_available()) return n;
    volatile uint32_t next = 0, failed = 0;
    zdec_job job = {base, off, len, raw, n, &next, &failed};
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    int started = 0;
    for (int t = 0; t < nthreads; t++) if (pthread_create(&th[t], NULL, zdec_worker, &job) == 0) started++; else break;
    if (!started) zdec_worker(&job);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
    return failed;
}

/* ---- batch of independent frames, multi-threaded ---- */
typedef struct {
    const int32_t *kind;      /* per frame */
    const uint64_t *seed;     /* per frame */
    const uint64_t *size;     /* per frame uncompressed size */
    uint32_t n;
    int level, checksum;
    uint8_t *dst;
    uint64_t dst_cap;
    uint64_t *frame_off, *frame_len, *hash;
    volatile uint64_t *dst_used;
    volatile uint32_t *next;
    volatile int *failed;
} gen_job;

static void *gen_worker(void *arg) {
    gen_job *j = (gen_job *)arg;
    void *c = Z.createCCtx();
    Z.setParameter(c, 100, j->level);
    Z.setParameter(c, 201, j->checksum);
    uint8_t *raw = NULL, *comp = NULL;
    size_t raw_cap = 0, comp_cap = 0;
    for (;;) {
        uint32_t i = __atomic_fetch_add(j->next, 1, __ATOMIC_RELAXED);
        if (i >= j->n || *j->failed) break;
        size_t n = (size_t)j->size[i];
        if (n > raw_cap || !raw) { free(raw); raw = malloc(n ? n : 1); raw_cap = n; }
        size_t bound = Z.compressBound(n);
        if (bound > comp_cap || !comp) { free(comp); comp = malloc(bound ? bound : 64); comp_cap = bound; }
        if (!raw || !comp) { *j->failed = 1; break; }
        cg_fill(j->kind[i], j->seed[i], raw, n, 0);
        size_t r = Z.compress2(c, comp, comp_cap, raw, n);
        if (Z.isError(r)) { *j->failed = 1; break; }
        uint64_t at = __atomic_fetch_add(j->dst_used, (uint64_t)r, __ATOMIC_RELAXED);
        if (at + r > j->dst_cap) { *j->failed = 2; break; }
        memcpy(j->dst + at, comp, r);
        j->frame_off[i] = at;
        j->frame_len[i] = r;
        if (j->hash) j->hash[i] = cg_hash(raw, n);
    }
    free(raw);
    free(comp);
    Z.freeCCtx(c);
    return NULL;
}

/* Generates and compresses n frames (one zstd frame each).  Frames land in dst in completion
 * order; frame_off/frame_len say where.  Returns bytes used in dst, 0 on failure. */
uint64_t cg_generate_frames(const int32_t *kind, const uint64_t *seed, const uint64_t *size, uint32_t n, int level,
                            int checksum, int nthreads, uint8_t *dst, uint64_t dst_cap, uint64_t *frame_off,
                            uint64_t *frame_len, uint64_t *hash) {
    if (!cg_zstd_available()) return 0;
    pthread_once(&vocab_once, vocab_init);
    volatile uint64_t used = 0;
    volatile uint32_t next = 0;
    volatile int failed = 0;
    gen_job job = {kind, seed, size, n, level, checksum, dst, dst_cap, frame_off, frame_len, hash, &used, &next, &failed};
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    int started = 0;
    for (int t = 0; t < nthreads; t++) if (pthread_create(&th[t], NULL, gen_worker, &job) == 0) started++; else break;
    if (!started) gen_worker(&job);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
    return failed ? 0 : used;
}
