// dm_kernels_records.cuh -- record mode on the GPU: a batch of serialized ParserSchema
// records (the reference's real wire format of the detector's input,
// /root/reference/container/fluentout/schemas_pb.rb:8; one record per NNG message today,
// engine.py:163-187) walked, matched against the configured monitors, hashed and probed on
// the device.  One thread per record.
//
//   ParserSchema   4 EventID int32 (varint)            6 variables repeated string
//                 10 logFormatVariables map<string,string> (entry: 1 key, 2 value)
//   every other field is skipped by wire type.
//
// Monitor semantics (R-spec 3): a global monitor applies to every record, an
// event monitor only when record.EventID matches; header monitors read
// logFormatVariables[KEY] (last entry wins, as in protobuf maps), variable monitors read
// variables[i]; a missing field skips the monitor.
#pragma once
#include "dm_device.cuh"

struct DmMonitor {
    int32_t event_id;
    uint32_t has_event;      // 0 = global scope
    uint32_t source;         // 0 = header variable (by key), 1 = variable (by index)
    uint32_t var_index;
    uint32_t key_len;
    uint8_t key[64];
};

// Combination monitors (NewValueComboDetector, DESIGN.md R-combo): combo c is the ordered
// tuple of the values of monitors combo_members[combo_off[c] .. combo_off[c+1]); its table
// field (salt) and output-mask bit are n + c.  A monitor whose member_only bit is set does
// not alert on its own.
#define DM_MAX_COMBO_MEMBERS 64
struct DmMonitors {
    uint32_t n;
    uint32_t n_combos;
    uint32_t member_only;
    uint32_t pad_;
    uint32_t combo_off[DM_MAX_KEYS + 1];
    uint8_t combo_members[DM_MAX_COMBO_MEMBERS];
    uint32_t pad2_[3];
    DmMonitor m[DM_MAX_KEYS];
};

// Fingerprint of an ordered tuple of value fingerprints (order-sensitive fold).
DM_HD uint64_t dm_combo_fold(uint64_t acc, uint64_t fp) { return dm_splitmix64(acc ^ fp); }
DM_HD uint64_t dm_combo_seed(uint32_t n_members) { return 0xC0B0C0B0ull + 0x9E3779B97F4A7C15ull * (uint64_t)n_members; }

struct DmRecordsArgs {
    const uint8_t* buf;
    const uint32_t* rec_off;       // n_records + 1 offsets of the record payloads (after the length prefix)
    const uint32_t* rec_len;       // payload lengths
    uint32_t n_records;
    uint32_t n_train_records;
    const DmMonitors* mons;
    DmTable table;
    uint8_t* flags;
    float* scores;
    uint32_t* masks;
    DmBatchHeader* hdr;
    unsigned long long* stats;
};

__device__ __forceinline__ bool dm_pb_varint(const uint8_t* __restrict__ buf, uint32_t& pos, uint32_t end, uint64_t& out) {
    uint64_t v = 0;
    for (uint32_t shift = 0; shift < 70; shift += 7) {
        if (pos >= end) return false;
        const uint32_t b = buf[pos++];
        v |= (uint64_t)(b & 0x7Fu) << shift;
        if (!(b & 0x80u)) { out = v; return true; }
    }
    return false;
}

__global__ void __launch_bounds__(128) dm_k_records(DmRecordsArgs a, int phase) {
    // phase 0: training records insert; phase 1: the others are scored (separate launches so
    // that detection sees every insert)
    __shared__ DmMonitors sm;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(a.mons);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&sm);
        for (uint32_t i = threadIdx.x; i < sizeof(DmMonitors) / 4; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const uint8_t* __restrict__ buf = a.buf;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < a.n_records; r += gridDim.x * blockDim.x) {
        const bool train = r < a.n_train_records;
        if (train != (phase == 0)) continue;
        uint32_t vp[DM_MAX_KEYS], vl[DM_MAX_KEYS];
        uint32_t present = 0;
        uint32_t pos = a.rec_off[r];
        const uint32_t end = pos + a.rec_len[r];
        int32_t eid = 0;
        bool has_eid = false;
        uint32_t var_idx = 0;
        bool ok = true;
        while (ok && pos < end) {
            uint64_t tag;
            if (!dm_pb_varint(buf, pos, end, tag)) { ok = false; break; }
            const uint32_t field = (uint32_t)(tag >> 3), wt = (uint32_t)(tag & 7u);
            if (wt == 0) {
                uint64_t v;
                if (!dm_pb_varint(buf, pos, end, v)) { ok = false; break; }
                if (field == 4) { eid = (int32_t)(uint32_t)v; has_eid = true; }
            } else if (wt == 2) {
                uint64_t len64;
                if (!dm_pb_varint(buf, pos, end, len64) || len64 > (uint64_t)(end - pos)) { ok = false; break; }
                const uint32_t len = (uint32_t)len64;
                if (field == 6) {
                    for (uint32_t k = 0; k < sm.n; ++k)
                        if (sm.m[k].source == 1 && sm.m[k].var_index == var_idx) { vp[k] = pos; vl[k] = len; present |= 1u << k; }
                    ++var_idx;
                } else if (field == 10) {
                    uint32_t p = pos, kp = 0, kl = 0, xp = pos, xl = 0;      // key, value (absent = empty)
                    const uint32_t e = pos + len;
                    bool eok = true;
                    while (p < e) {
                        uint64_t t2, l2;
                        if (!dm_pb_varint(buf, p, e, t2) || (t2 & 7u) != 2u || !dm_pb_varint(buf, p, e, l2) || l2 > (uint64_t)(e - p)) { eok = false; break; }
                        if ((t2 >> 3) == 1) { kp = p; kl = (uint32_t)l2; }
                        else if ((t2 >> 3) == 2) { xp = p; xl = (uint32_t)l2; }
                        p += (uint32_t)l2;
                    }
                    if (!eok) { ok = false; break; }
                    for (uint32_t k = 0; k < sm.n; ++k) {
                        if (sm.m[k].source != 0 || sm.m[k].key_len != kl) continue;
                        bool eq = true;
                        for (uint32_t i = 0; i < kl; ++i)
                            if (buf[kp + i] != sm.m[k].key[i]) { eq = false; break; }
                        if (eq) { vp[k] = xp; vl[k] = xl; present |= 1u << k; }
                    }
                }
                pos += len;
            } else if (wt == 1) {
                pos += 8;
            } else if (wt == 5) {
                pos += 4;
            } else {
                ok = false;
            }
        }
        if (!ok) { atomicAdd(a.stats + 7, 1ull); continue; }     // malformed record: counted, not scored
        uint32_t unknown = 0, have = 0;
        uint64_t fpv[DM_MAX_KEYS];
        for (uint32_t k = 0; k < sm.n; ++k) {
            if (!((present >> k) & 1u)) continue;
            if (sm.m[k].has_event && !(has_eid && eid == sm.m[k].event_id)) continue;
            const uint64_t fp = dm_fp64_bytes(buf + vp[k], vl[k]);
            fpv[k] = fp;
            have |= 1u << k;
            if ((sm.member_only >> k) & 1u) continue;
            const uint64_t key = dm_make_key(fp, dm_field_salt(k));
            if (train) dm_table_insert(a.table, key, &a.hdr->error);
            else if (!dm_table_contains(a.table, key)) unknown |= 1u << k;
        }
        for (uint32_t c = 0; c < sm.n_combos; ++c) {
            const uint32_t lo = sm.combo_off[c], hi = sm.combo_off[c + 1];
            uint64_t acc = dm_combo_seed(hi - lo);
            bool all = true;
            for (uint32_t j = lo; j < hi; ++j) {
                const uint32_t k = sm.combo_members[j];
                if (!((have >> k) & 1u)) { all = false; break; }
                acc = dm_combo_fold(acc, fpv[k]);
            }
            if (!all) continue;                                   // a missing member skips the combination
            const uint64_t key = dm_make_key(acc ? acc : 1ull, dm_field_salt(sm.n + c));
            if (train) dm_table_insert(a.table, key, &a.hdr->error);
            else if (!dm_table_contains(a.table, key)) unknown |= 1u << (sm.n + c);
        }
        if (!train) {
            const uint32_t cnt = (uint32_t)__popc(unknown);
            a.flags[r] = cnt ? 1 : 0;
            a.scores[r] = (float)cnt;
            a.masks[r] = unknown;
            if (cnt) {
                atomicAdd(&a.hdr->n_anomalies, 1ull);
                atomicAdd(a.stats + 3, 1ull);
                atomicAdd(a.stats + 4, (unsigned long long)cnt);
                uint32_t m = unknown;
                while (m) { const int k = __ffs(m) - 1; m &= m - 1; atomicAdd(a.stats + 8 + k, 1ull); }
            }
        } else {
            a.flags[r] = 0; a.scores[r] = 0.0f; a.masks[r] = 0;
        }
    }
}
