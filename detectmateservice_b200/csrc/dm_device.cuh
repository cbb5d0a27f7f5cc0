// dm_device.cuh -- device-side data structures and helpers shared by the kernels.
//
// Compiles with nvcc for sm_100a (the product) and, with -DDM_EMU, with g++ against
// tests/emu/cuda_emu.h (CPU emulation used only by the test tier).
#pragma once
#ifdef DM_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include "../../include/dmdetect.h"
#include "dm_hash.h"

#define DM_STATS_WORDS (8 + DM_MAX_KEYS)   // layout of dm_stats_t as uint64 words

// Monitored keys, as the kernels see them (copied to shared memory per CTA).
struct DmKeys {
    uint32_t n;
    uint32_t len[DM_MAX_KEYS];
    uint64_t salt[DM_MAX_KEYS];
    uint8_t bytes[DM_MAX_KEYS][DM_MAX_KEYLEN];
    // fused-kernel key identification.  With w_c = bytes q-4..q-1 and w_b = bytes q-8..q-5 in
    // front of an '=' at q (later byte in the higher bits), key k matches its last
    // min(len,8) bytes iff ((w_c ^ tailbits) & tailmask) | ((w_b ^ midbits) & midmask) == 0.
    // The field-start delimiter (byte q-len-1) sits in word dsel (0 = w_c, 1 = w_b, 2 = w_a =
    // bytes q-12..q-9) at bit offset dshift, for keys shorter than 12 bytes.
    uint32_t tailbits[DM_MAX_KEYS];
    uint32_t tailmask[DM_MAX_KEYS];
    uint32_t midbits[DM_MAX_KEYS];
    uint32_t midmask[DM_MAX_KEYS];
    uint32_t dsel[DM_MAX_KEYS];
    uint32_t dshift[DM_MAX_KEYS];
    // the same patterns packed for one 16-byte shared-memory load per key, in the order the
    // fused kernel tries the keys: longest first, so that a key that is a suffix of another
    // ("res" / "xres") is only considered when the longer one does not match.
    uint32_t order[DM_MAX_KEYS];
    alignas(16) uint32_t pat[DM_MAX_KEYS][4];      // [slot] = {tailbits, tailmask, midbits, midmask} of key order[slot]
};

static inline void dm_keys_finalize_host(DmKeys* k) {
    for (uint32_t i = 0; i < k->n; ++i) {
        const uint32_t len = k->len[i];
        uint32_t tb = 0, tm = 0, mb = 0, mm = 0;
        for (uint32_t d = 1; d <= len && d <= 8; ++d) {            // d = distance from the '='
            const uint32_t byte = k->bytes[i][len - d];
            if (d <= 4) { tb |= byte << (8 * (4 - d)); tm |= 0xFFu << (8 * (4 - d)); }
            else { mb |= byte << (8 * (8 - d)); mm |= 0xFFu << (8 * (8 - d)); }
        }
        k->tailbits[i] = tb; k->tailmask[i] = tm; k->midbits[i] = mb; k->midmask[i] = mm;
        const uint32_t dd = len + 1;                                // distance of the delimiter
        k->dsel[i] = dd <= 12 ? (dd - 1) / 4 : 3;                   // 3 = not in the 12-byte window
        k->dshift[i] = dd <= 12 ? 8 * (4 * ((dd - 1) / 4 + 1) - dd) : 0;
    }
    for (uint32_t i = 0; i < k->n; ++i) k->order[i] = i;
    for (uint32_t i = 1; i < k->n; ++i) {                           // insertion sort, longest first, stable
        const uint32_t v = k->order[i];
        uint32_t j = i;
        while (j > 0 && k->len[k->order[j - 1]] < k->len[v]) { k->order[j] = k->order[j - 1]; --j; }
        k->order[j] = v;
    }
    for (uint32_t s = 0; s < k->n; ++s) {
        const uint32_t i = k->order[s];
        k->pat[s][0] = k->tailbits[i]; k->pat[s][1] = k->tailmask[i];
        k->pat[s][2] = k->midbits[i]; k->pat[s][3] = k->midmask[i];
    }
}

// Per-batch header written by the kernels, read back by the host.
struct DmBatchHeader {
    unsigned long long n_newlines;
    unsigned long long n_lines;        // records in the batch (R-tok L1)
    unsigned long long n_anomalies;    // records with score > 0
    unsigned int anomaly_list_count;   // entries appended to the anomaly list (may exceed cap)
    unsigned int error;                // DM_DEVERR_* bits
};

#define DM_DEVERR_TABLE_FULL 1u
#define DM_DEVERR_TOO_MANY_LINES 2u
#define DM_DEVERR_NOVEL_OVERFLOW 4u

// Known-set table: open addressing, linear probing, 0 = empty slot.
struct DmTable {
    unsigned long long* slots;
    uint32_t mask;                     // capacity - 1
    uint32_t limit;                    // max entries (load limit)
    unsigned long long* count;         // entries
    unsigned long long* novel;         // keys inserted since creation, in insertion order
    unsigned long long* novel_count;
    uint32_t novel_cap;
};

__device__ __forceinline__ uint32_t dm_slot_of(uint64_t key, uint32_t mask) {
    return ((uint32_t)key ^ (uint32_t)(key >> 32)) & mask;
}

// Read-only probe (detection launches: the table does not change while they run).
__device__ __forceinline__ bool dm_table_contains(const DmTable& t, uint64_t key) {
    uint32_t i = dm_slot_of(key, t.mask);
    for (uint32_t probes = 0; probes <= t.mask; ++probes) {
        unsigned long long v = __ldg(t.slots + i);
        if (v == key) return true;
        if (v == 0ull) return false;
        i = (i + 1) & t.mask;
    }
    return false;
}

// Probe that tolerates concurrent inserts (training launches).
__device__ __forceinline__ bool dm_table_contains_volatile(const DmTable& t, uint64_t key) {
    uint32_t i = dm_slot_of(key, t.mask);
    for (uint32_t probes = 0; probes <= t.mask; ++probes) {
        unsigned long long v = *((volatile unsigned long long*)(t.slots + i));
        if (v == key) return true;
        if (v == 0ull) return false;
        i = (i + 1) & t.mask;
    }
    return false;
}

// Returns true if this call inserted the key (it was not present).
__device__ __forceinline__ bool dm_table_insert(const DmTable& t, uint64_t key, unsigned int* err) {
    uint32_t i = dm_slot_of(key, t.mask);
    for (uint32_t probes = 0; probes <= t.mask; ++probes) {
        unsigned long long v = atomicCAS(t.slots + i, 0ull, (unsigned long long)key);
        if (v == 0ull) {
            unsigned long long c = atomicAdd(t.count, 1ull);
            if (c + 1 > t.limit) atomicOr(err, DM_DEVERR_TABLE_FULL);
            unsigned long long j = atomicAdd(t.novel_count, 1ull);
            if (j < t.novel_cap) t.novel[j] = key; else atomicOr(err, DM_DEVERR_NOVEL_OVERFLOW);
            return true;
        }
        if (v == key) return false;
        i = (i + 1) & t.mask;
    }
    atomicOr(err, DM_DEVERR_TABLE_FULL);
    return false;
}

#ifndef DM_EMU
// 4-bit mask of the bytes of w equal to the replicated byte pattern pat (bit j = byte j).
__device__ __forceinline__ uint32_t dm_nib_eq(uint32_t w, uint32_t pat) {
    uint32_t m = __vcmpeq4(w, pat) & 0x08040201u;
    return (m * 0x01010101u) >> 24;
}

// 16-bit mask over a 16-byte chunk.
__device__ __forceinline__ uint32_t dm_mask16_eq(uint4 v, uint32_t pat) {
    return dm_nib_eq(v.x, pat) | (dm_nib_eq(v.y, pat) << 4) | (dm_nib_eq(v.z, pat) << 8) |
           (dm_nib_eq(v.w, pat) << 12);
}
#endif

__device__ __forceinline__ uint32_t dm_lanemask_lt() {
#ifdef DM_EMU
    return (1u << (threadIdx.x & 31)) - 1u;
#else
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
#endif
}

// Named barriers (PTX bar.sync / bar.arrive; counts are in threads).
__device__ __forceinline__ void dm_bar_sync(int id, int count) {
#ifdef DM_EMU
    emu_bar_sync(id, (unsigned)count);
#else
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
#endif
}
__device__ __forceinline__ void dm_bar_arrive(int id, int count) {
#ifdef DM_EMU
    emu_bar_arrive(id, (unsigned)count);
#else
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
#endif
}

// Arguments of the warp-per-record kernels (dm_kernels_v1.cuh, dm_kernels_format.cuh).
struct DmKeys;
struct DmDetectArgs {
    const uint8_t* buf;
    const uint32_t* line_start;
    const DmBatchHeader* hdr_in;      // n_lines
    DmBatchHeader* hdr;
    const DmKeys* keys;
    DmTable table;
    uint8_t* flags;
    float* scores;
    uint64_t out_cap;
    dm_anomaly_t* anomalies;
    uint32_t anomaly_cap;
    unsigned long long* stats;
    uint64_t line_lo, line_hi;        // records [line_lo, min(line_hi, n_lines)) are processed
    uint64_t nbytes;                  // message size
    const void* combos;               // lanes kernel with combinations: const DmMonitors* (else NULL)
};
