// dm_kernels_values.cuh -- record mode: the detector fed with already-extracted values.
//
// In the reference pipeline the detector receives ONE ParserSchema per message
// (/root/reference/src/service/features/engine.py:163-187) whose monitored values sit in
// logFormatVariables / variables (docs/interfaces.md:138-204).  The host decodes the
// protobuf framing; hashing, set membership and scoring stay on the device, through the
// same dm_fp64 / table code as the raw-line kernels.
#pragma once
#include "dm_device.cuh"

struct DmValuesArgs {
    const uint8_t* blob;
    const uint32_t* offsets;      // n_values + 1
    const uint32_t* fields;       // monitored field of value i
    const uint32_t* record_of;    // record index of value i
    uint32_t n_values;
    uint32_t n_records;
    uint32_t n_train_records;     // records [0, n_train_records) are training data
    const DmKeys* keys;
    DmTable table;
    uint8_t* flags;               // n_records, zero-filled by the caller
    float* scores;
    uint32_t* masks;
    DmBatchHeader* hdr;
    unsigned long long* stats;
};

__global__ void __launch_bounds__(256) dm_k_values(DmValuesArgs a, int phase) {
    // phase 0: training records insert; phase 1: detection records probe (separate launches,
    // so that detection sees every insert -- R-spec 1/2 ordering)
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n_values; i += gridDim.x * blockDim.x) {
        const uint32_t rec = a.record_of[i];
        const bool train = rec < a.n_train_records;
        if (train != (phase == 0)) continue;
        const uint32_t k = a.fields[i];
        const uint32_t b = a.offsets[i], e = a.offsets[i + 1];
        const uint64_t key = dm_make_key(dm_fp64_bytes(a.blob + b, e - b), dm_field_salt(k));
        if (train) {
            dm_table_insert(a.table, key, &a.hdr->error);
        } else if (!dm_table_contains(a.table, key)) {
            const float old = atomicAdd(a.scores + rec, 1.0f);
            a.flags[rec] = 1;
            atomicOr(a.masks + rec, 1u << k);
            atomicAdd(a.stats + 8 + k, 1ull);
            atomicAdd(a.stats + 4, 1ull);
            if (old == 0.0f) {
                atomicAdd(&a.hdr->n_anomalies, 1ull);
                atomicAdd(a.stats + 3, 1ull);
            }
        }
    }
}
