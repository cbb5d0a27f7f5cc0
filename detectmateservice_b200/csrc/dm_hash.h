// dm_hash.h -- dm_fp64 value fingerprint and table-key derivation (host + device).
//
// The reference's NewValueDetector keeps Python sets of value strings (detectmatelibrary,
// un-vendored; call site /root/reference/src/service/core.py:201-203).  The device path keeps
// 64-bit keys instead; the definition is fixed in DESIGN.md ("dm_fp64") and restated
// independently by the test oracle.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define DM_HD __host__ __device__ __forceinline__
#else
#define DM_HD static inline
#endif

DM_HD uint32_t dm_rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

struct DmHashState {
    uint32_t h, g;
};

DM_HD void dm_hash_init(DmHashState& s) {
    s.h = 0x9747B28Cu;
    s.g = 0x165667B1u;
}

// One little-endian 32-bit word of the (zero padded) value.
DM_HD void dm_hash_word(DmHashState& s, uint32_t w) {
    uint32_t k = dm_rotl32(w * 0xCC9E2D51u, 15) * 0x1B873593u;
    s.h = dm_rotl32(s.h ^ k, 13) * 5u + 0xE6546B64u;
    s.g = dm_rotl32(s.g + w * 0x85EBCA77u, 13) * 0x9E3779B1u;
}

DM_HD uint64_t dm_hash_final(const DmHashState& s, uint32_t n) {
    uint32_t h = s.h ^ n, g = s.g ^ n;
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    g ^= g >> 15; g *= 0x85EBCA77u; g ^= g >> 13; g *= 0xC2B2AE3Du; g ^= g >> 16;
    uint64_t fp = ((uint64_t)h << 32) | g;
    return fp ? fp : 1ull;
}

DM_HD uint64_t dm_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

DM_HD uint64_t dm_field_salt(uint32_t field) { return dm_splitmix64((uint64_t)field + 1ull); }

DM_HD uint64_t dm_make_key(uint64_t fp, uint64_t salt) {
    uint64_t k = fp ^ salt;
    return k ? k : 1ull;
}

// Whole-value fingerprint of a byte string (host helper and slow device paths).
DM_HD uint64_t dm_fp64_bytes(const uint8_t* v, uint32_t n) {
    DmHashState s;
    dm_hash_init(s);
    for (uint32_t i = 0; i < n; i += 4) {
        uint32_t w = 0;
        for (uint32_t j = 0; j < 4 && i + j < n; j++) w |= (uint32_t)v[i + j] << (8 * j);
        dm_hash_word(s, w);
    }
    return dm_hash_final(s, n);
}
