// dm_kernels_tile.cuh -- fused single-pass tile kernel (placeholder until it lands).
#pragma once
#include "dm_device.cuh"

struct DmTileScratch { int unused; };

static inline int dm_tile_scratch_create(DmTileScratch*, uint64_t, int) { return DM_OK; }
static inline void dm_tile_scratch_destroy(DmTileScratch*) {}
static inline int dm_tile_launch(DmTileScratch*, const uint8_t*, uint64_t, uint64_t, const DmKeys*, DmTable, uint8_t*,
                                 float*, uint64_t, dm_anomaly_t*, uint32_t, DmBatchHeader*, unsigned long long*,
                                 uint64_t, cudaStream_t) { return DM_ERR_STATE; }   // >= 0: kernels launched
