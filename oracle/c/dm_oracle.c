/*
 * dm_oracle.c -- plain-C restatement of the detector hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Same rules as oracle/rtok.py (R-tok L1-L6) and oracle/nvd.py (R-spec 1-4), written
 * sequentially, byte by byte, with EXACT string sets (no fingerprints), so that it can
 * check the CUDA path on inputs too large for the pure-Python oracle and serve as the
 * native CPU baseline of bench.py.  It restates, for raw-line input, what the reference
 * does per message in Engine._run_loop -> Service.process -> NewValueDetector
 * (/root/reference/src/service/features/engine.py:153-217, src/service/core.py:176-206;
 * detector arithmetic in the un-vendored detectmatelibrary 0.1.0 @ ecdda558,
 * uv.lock:240-251 -- PARITY UNPINNED, see oracle/__init__.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this library.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define DMO_MAX_KEYS 32
#define DMO_MAX_KEYLEN 64

typedef struct {
    uint64_t hash;
    uint64_t off;   /* into arena */
    uint32_t len;
    uint32_t used;
} dmo_slot;

typedef struct {
    dmo_slot* slots;
    uint64_t cap;   /* power of two */
    uint64_t count;
    uint8_t* arena;
    uint64_t arena_len, arena_cap;
} dmo_set;

typedef struct dmo {
    int n_keys;
    uint8_t keys[DMO_MAX_KEYS][DMO_MAX_KEYLEN];
    uint32_t key_len[DMO_MAX_KEYS];
    dmo_set sets[DMO_MAX_KEYS];
    uint64_t n_lines, n_anomalies;
    uint64_t unknown_per_key[DMO_MAX_KEYS];
} dmo;

static uint64_t fnv1a(const uint8_t* p, uint32_t n) {
    uint64_t h = 1469598103934665603ull;
    for (uint32_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}

static void set_init(dmo_set* s) {
    s->cap = 1024; s->count = 0;
    s->slots = (dmo_slot*)calloc(s->cap, sizeof(dmo_slot));
    s->arena_cap = 1 << 16; s->arena_len = 0;
    s->arena = (uint8_t*)malloc(s->arena_cap);
}

static int set_find(const dmo_set* s, const uint8_t* v, uint32_t n, uint64_t h, uint64_t* idx_out) {
    uint64_t i = h & (s->cap - 1);
    for (;;) {
        const dmo_slot* sl = &s->slots[i];
        if (!sl->used) { *idx_out = i; return 0; }
        if (sl->hash == h && sl->len == n && memcmp(s->arena + sl->off, v, n) == 0) { *idx_out = i; return 1; }
        i = (i + 1) & (s->cap - 1);
    }
}

static void set_grow(dmo_set* s) {
    uint64_t ocap = s->cap; dmo_slot* old = s->slots;
    s->cap = ocap * 2;
    s->slots = (dmo_slot*)calloc(s->cap, sizeof(dmo_slot));
    for (uint64_t j = 0; j < ocap; j++) if (old[j].used) {
        uint64_t i = old[j].hash & (s->cap - 1);
        while (s->slots[i].used) i = (i + 1) & (s->cap - 1);
        s->slots[i] = old[j];
    }
    free(old);
}

static void set_insert(dmo_set* s, const uint8_t* v, uint32_t n) {
    uint64_t h = fnv1a(v, n), idx;
    if (set_find(s, v, n, h, &idx)) return;
    if (s->arena_len + n > s->arena_cap) {
        while (s->arena_len + n > s->arena_cap) s->arena_cap *= 2;
        s->arena = (uint8_t*)realloc(s->arena, s->arena_cap);
    }
    memcpy(s->arena + s->arena_len, v, n);
    s->slots[idx].hash = h; s->slots[idx].off = s->arena_len; s->slots[idx].len = n; s->slots[idx].used = 1;
    s->arena_len += n; s->count++;
    if (s->count * 2 > s->cap) set_grow(s);
}

static int set_contains(const dmo_set* s, const uint8_t* v, uint32_t n) {
    uint64_t idx;
    return set_find(s, v, n, fnv1a(v, n), &idx);
}

dmo* dmo_create(int n_keys, const uint8_t* keys_blob, const uint32_t* key_lens) {
    if (n_keys < 0 || n_keys > DMO_MAX_KEYS) return NULL;
    dmo* d = (dmo*)calloc(1, sizeof(dmo));
    d->n_keys = n_keys;
    uint64_t off = 0;
    for (int k = 0; k < n_keys; k++) {
        if (key_lens[k] == 0 || key_lens[k] > DMO_MAX_KEYLEN) { free(d); return NULL; }
        memcpy(d->keys[k], keys_blob + off, key_lens[k]);
        d->key_len[k] = key_lens[k];
        off += key_lens[k];
        set_init(&d->sets[k]);
    }
    return d;
}

void dmo_destroy(dmo* d) {
    if (!d) return;
    for (int k = 0; k < d->n_keys; k++) { free(d->sets[k].slots); free(d->sets[k].arena); }
    free(d);
}

/* One line: R-tok L2-L6 then R-spec 1/2/4.  Returns the unknown-field mask. */
static uint32_t do_line(dmo* d, const uint8_t* line, uint64_t n, int train) {
    uint32_t seen = 0, unknown = 0;
    int inq = 0;
    for (uint64_t p = 0; p < n; p++) {
        uint8_t c = line[p];
        int start = !inq && (p == 0 || line[p - 1] == 0x20 || line[p - 1] == 0x27);
        if (start) {
            uint64_t q = p; int ok = 0;
            while (q < n) {
                uint8_t b = line[q];
                if (b == 0x3D) { ok = q > p; break; }
                if (b == 0x20 || b == 0x22 || b == 0x27) break;
                q++;
            }
            if (ok) {
                uint32_t klen = (uint32_t)(q - p);
                int k;
                for (k = 0; k < d->n_keys; k++)
                    if (d->key_len[k] == klen && memcmp(d->keys[k], line + p, klen) == 0) break;
                if (k < d->n_keys && !(seen & (1u << k))) {        /* L6: first occurrence wins */
                    seen |= 1u << k;
                    uint64_t e = q + 1; int vq = 0;
                    while (e < n) {
                        uint8_t b = line[e];
                        if (b == 0x20 && !vq) break;
                        if (b == 0x22) vq ^= 1;
                        e++;
                    }
                    const uint8_t* v = line + q + 1; uint32_t vlen = (uint32_t)(e - q - 1);
                    if (train) set_insert(&d->sets[k], v, vlen);
                    else if (!set_contains(&d->sets[k], v, vlen)) { unknown |= 1u << k; d->unknown_per_key[k]++; }
                }
            }
        }
        if (c == 0x22) inq ^= 1;
    }
    return unknown;
}

/* Process one message of raw lines.  The first n_train_lines records of THIS call are
 * training records (flag 0, score 0), the rest are detected.  flags/scores/masks may be
 * NULL.  Returns the number of records (R-tok L1). */
uint64_t dmo_process(dmo* d, const uint8_t* buf, uint64_t n, uint64_t n_train_lines,
                     uint8_t* flags, float* scores, uint32_t* masks) {
    uint64_t li = 0, s = 0;
    while (s < n) {
        const uint8_t* nl = (const uint8_t*)memchr(buf + s, '\n', n - s);
        uint64_t e = nl ? (uint64_t)(nl - buf) : n;
        int train = li < n_train_lines;
        uint32_t m = do_line(d, buf + s, e - s, train);
        int cnt = __builtin_popcount(m);
        if (flags) flags[li] = cnt > 0;
        if (scores) scores[li] = (float)cnt;
        if (masks) masks[li] = m;
        if (cnt) d->n_anomalies++;
        li++;
        s = e + 1;
    }
    d->n_lines += li;
    return li;
}

uint64_t dmo_known_count(const dmo* d, int k) { return (k >= 0 && k < d->n_keys) ? d->sets[k].count : 0; }
uint64_t dmo_total_lines(const dmo* d) { return d->n_lines; }
uint64_t dmo_total_anomalies(const dmo* d) { return d->n_anomalies; }
uint64_t dmo_unknown_count(const dmo* d, int k) { return (k >= 0 && k < d->n_keys) ? d->unknown_per_key[k] : 0; }

/* Copy the known values of field k out as [u32 len][bytes]... ; returns bytes needed. */
uint64_t dmo_export_known(const dmo* d, int k, uint8_t* out, uint64_t cap) {
    if (k < 0 || k >= d->n_keys) return 0;
    const dmo_set* s = &d->sets[k];
    uint64_t need = 0;
    for (uint64_t i = 0; i < s->cap; i++) if (s->slots[i].used) {
        uint32_t len = s->slots[i].len;
        if (out && need + 4 + len <= cap) { memcpy(out + need, &len, 4); memcpy(out + need + 4, s->arena + s->slots[i].off, len); }
        need += 4 + len;
    }
    return need;
}

/* dm_fp64 restated in C (definition: oracle/fingerprint.py) for cross-checking. */
static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
uint64_t dmo_fp64(const uint8_t* v, uint32_t n) {
    uint32_t h = 0x9747B28Cu, g = 0x165667B1u;
    for (uint32_t i = 0; i < n; i += 4) {
        uint32_t w = 0;
        for (uint32_t j = 0; j < 4 && i + j < n; j++) w |= (uint32_t)v[i + j] << (8 * j);
        uint32_t k = rotl32(w * 0xCC9E2D51u, 15) * 0x1B873593u;
        h = rotl32(h ^ k, 13) * 5u + 0xE6546B64u;
        g = rotl32(g + w * 0x85EBCA77u, 13) * 0x9E3779B1u;
    }
    h ^= n; h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    g ^= n; g ^= g >> 15; g *= 0x85EBCA77u; g ^= g >> 13; g *= 0xC2B2AE3Du; g ^= g >> 16;
    uint64_t fp = ((uint64_t)h << 32) | g;
    return fp ? fp : 1;
}
