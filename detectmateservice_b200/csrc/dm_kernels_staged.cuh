// dm_kernels_staged.cuh -- the "staged" variant: the rows decomposition with its three
// stages split into kernels that each map ONE item to ONE thread, so every lane has work.
//
// Why: in dm_k_rows a warp owns a few rows and drains its private queues when they hold 32
// entries; with 16 MiB messages a warp only ever sees ~100 '=' and ~40 fields, so the
// expensive stages run half empty (ncu: 12-25 of 32 lanes active).  Here the stages talk
// through compacted lists in global memory (L2-resident: ~7 MB + ~4 MB per 16 MiB message):
//
//   K_A  dm_k_rowindex (dm_kernels_rows.cuh)  row prefix of '\n', zero-fill, header, counters = 0
//   K_1  dm_k_stage1   per 512-byte row: classify '\n' / '=', record index of every '=',
//                      append {position, record} of every '=' to the candidate list
//                      (one atomicAdd per CTA of 32 rows)
//   K_2  dm_k_stage2   one candidate per thread: which monitored key stands in front of the
//                      '=' (12-byte window against packed key patterns); append
//                      {value position, record, key} to the field list (one atomicAdd per CTA)
//   K_3  dm_k_stage3   one field per thread: dm_fp64 of the value, table probe; the rare
//                      unknown values are re-checked exactly by the whole warp, then alert /
//                      insert.  Runs twice on a message that holds training AND detection
//                      records (train pass, then detect pass), filtered by record index.
//
// Same device helpers and same rules as the other variants; parity-tested against the
// oracle like them (DM_KERNEL=staged).
#pragma once
#include "dm_kernels_rows.cuh"

#define DMS_THREADS 256
#define DMS_WARPS (DMS_THREADS / 32)
#define DMS_ROWS_PER_WARP 4u
#define DMS_ROWS_PER_CTA (DMS_WARPS * DMS_ROWS_PER_WARP)

struct DmCand { uint32_t q; uint32_t g; };
struct DmField { uint32_t vpos; uint32_t g; uint32_t k; };

struct DmStagedArgs {
    DmRowsArgs r;                 // message, row prefix, keys, table, outputs, header, statistics
    DmCand* cand;
    DmField* fields;
    unsigned int* counts;         // [0] candidates, [1] fields (sized for the worst case: cannot overflow)
    uint32_t cand_cap, field_cap;
};

// ---------------------------------------------------------------------------------------
// K_1: rows -> candidate list
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DMS_THREADS) dm_k_stage1(DmStagedArgs a) {
    __shared__ uint32_t s_cnt[DMS_ROWS_PER_CTA];
    __shared__ uint32_t s_off[DMS_ROWS_PER_CTA];
    __shared__ uint32_t s_base;
    const uint8_t* __restrict__ buf = a.r.buf;
    const uint64_t nbytes = a.r.nbytes;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt = dm_lanemask_lt();

    for (uint32_t tile = blockIdx.x; tile * DMS_ROWS_PER_CTA < a.r.n_rows; tile += gridDim.x) {
        uint32_t eq16[DMS_ROWS_PER_WARP], nl16[DMS_ROWS_PER_WARP], gch[DMS_ROWS_PER_WARP], excl[DMS_ROWS_PER_WARP];
        uint4 v[DMS_ROWS_PER_WARP];
        const uint32_t row0 = tile * DMS_ROWS_PER_CTA + warp * DMS_ROWS_PER_WARP;
#pragma unroll
        for (uint32_t i = 0; i < DMS_ROWS_PER_WARP; ++i) {
            const uint64_t off = (uint64_t)(row0 + i) * DMR_ROW + (uint64_t)lane * 16;
            v[i] = make_uint4(0, 0, 0, 0);
            if (off < nbytes) v[i] = __ldg(reinterpret_cast<const uint4*>(buf + off));
        }
#pragma unroll
        for (uint32_t i = 0; i < DMS_ROWS_PER_WARP; ++i) {
            const uint32_t row = row0 + i;
            const uint64_t off = (uint64_t)row * DMR_ROW + (uint64_t)lane * 16;
            eq16[i] = 0; nl16[i] = 0; gch[i] = 0;
            if (off < nbytes) {
                nl16[i] = dm_row_nl_mask(v[i], off, nbytes);
                const uint32_t e0 = dm_eqflags(v[i].x, 0x3D3D3D3Du), e1 = dm_eqflags(v[i].y, 0x3D3D3D3Du);
                const uint32_t e2 = dm_eqflags(v[i].z, 0x3D3D3D3Du), e3 = dm_eqflags(v[i].w, 0x3D3D3D3Du);
                if (e0 | e1 | e2 | e3) {
                    eq16[i] = dm_flags_to_nib(e0) | (dm_flags_to_nib(e1) << 4) | (dm_flags_to_nib(e2) << 8) | (dm_flags_to_nib(e3) << 12);
                    if (off + 16 > nbytes) eq16[i] &= (1u << (uint32_t)(nbytes - off)) - 1u;
                }
            }
            // record index in front of this lane's chunk
            const uint32_t b_nl = __ballot_sync(0xffffffffu, nl16[i] != 0);
            uint32_t pre = 0;
            if (b_nl) {
                const uint32_t my = (uint32_t)__popc(nl16[i]);
                if (__ballot_sync(0xffffffffu, my > 1) == 0) {
                    pre = __popc(b_nl & lt);
                } else {
                    uint32_t incl = my;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
                        if ((int)lane >= d) incl += y;
                    }
                    pre = incl - my;
                }
            }
            if (row < a.r.n_rows) gch[i] = a.r.row_prefix[row] + pre;
            // exclusive prefix of the '=' counts over the lanes
            const uint32_t my_eq = (uint32_t)__popc(eq16[i]);
            uint32_t incl = my_eq;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
                if ((int)lane >= d) incl += y;
            }
            excl[i] = incl - my_eq;
            if (lane == 31) s_cnt[warp * DMS_ROWS_PER_WARP + i] = incl;
        }
        __syncthreads();
        if (warp == 0) {
            // 32 row counts -> offsets inside the CTA's block of the list; one atomicAdd for the CTA
            const uint32_t c = s_cnt[lane];
            uint32_t incl = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
                if ((int)lane >= d) incl += y;
            }
            s_off[lane] = incl - c;
            if (lane == 31) {
                const uint32_t total = incl;
                uint32_t base = 0xFFFFFFFFu;
                if (total) {
                    base = atomicAdd(a.counts + 0, total);
                    if (base + total > a.cand_cap) { atomicOr(&a.r.hdr->error, DM_DEVERR_TOO_MANY_LINES); base = 0xFFFFFFFFu; }
                }
                s_base = base;
            }
        }
        __syncthreads();
        const uint32_t base = s_base;
        if (base != 0xFFFFFFFFu) {
#pragma unroll
            for (uint32_t i = 0; i < DMS_ROWS_PER_WARP; ++i) {
                uint32_t m = eq16[i];
                uint32_t slot = base + s_off[warp * DMS_ROWS_PER_WARP + i] + excl[i];
                const uint32_t off32 = (uint32_t)((uint64_t)(row0 + i) * DMR_ROW) + lane * 16u;
                while (m) {
                    const uint32_t j = (uint32_t)__ffs(m) - 1;
                    m &= m - 1;
                    DmCand e;
                    e.q = off32 + j;
                    e.g = gch[i] + (uint32_t)__popc(nl16[i] & ((1u << j) - 1u));
                    a.cand[slot++] = e;
                }
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// K_2: candidates -> fields
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DMS_THREADS) dm_k_stage2(DmStagedArgs a) {
    __shared__ DmKeys sk;
    __shared__ uint32_t s_woff[DMS_WARPS];
    __shared__ uint32_t s_base;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(a.r.keys);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&sk);
        for (uint32_t i = threadIdx.x; i < sizeof(DmKeys) / 4; i += DMS_THREADS) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const uint8_t* __restrict__ buf = a.r.buf;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt = dm_lanemask_lt();
    uint32_t n = a.counts[0];
    if (n > a.cand_cap) n = a.cand_cap;
    for (uint32_t start = blockIdx.x * DMS_THREADS; start < n; start += gridDim.x * DMS_THREADS) {
        const uint32_t i = start + threadIdx.x;
        bool matched = false;
        DmField f;
        f.vpos = 0; f.g = 0; f.k = 0;
        if (i < n) {
            const DmCand c = a.cand[i];
            const int k = dm_key_identify(buf, (uint64_t)c.q, sk);
            if (k >= 0) { matched = true; f.vpos = c.q + 1; f.g = c.g; f.k = (uint32_t)k; }
        }
        const uint32_t mb = __ballot_sync(0xffffffffu, matched);
        if (lane == 0) s_woff[warp] = (uint32_t)__popc(mb);
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t tot = 0;
#pragma unroll
            for (int w = 0; w < DMS_WARPS; ++w) { const uint32_t c = s_woff[w]; s_woff[w] = tot; tot += c; }
            uint32_t base = 0xFFFFFFFFu;
            if (tot) {
                base = atomicAdd(a.counts + 1, tot);
                if (base + tot > a.field_cap) { atomicOr(&a.r.hdr->error, DM_DEVERR_TOO_MANY_LINES); base = 0xFFFFFFFFu; }
            }
            s_base = base;
        }
        __syncthreads();
        if (matched && s_base != 0xFFFFFFFFu) a.fields[s_base + s_woff[warp] + (uint32_t)__popc(mb & lt)] = f;
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// K_3: fields -> alerts / inserts
// ---------------------------------------------------------------------------------------
template <bool TRAIN>
__global__ void __launch_bounds__(DMS_THREADS) dm_k_stage3(DmStagedArgs a) {
    __shared__ DmKeys sk;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(a.r.keys);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&sk);
        for (uint32_t i = threadIdx.x; i < sizeof(DmKeys) / 4; i += DMS_THREADS) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const uint8_t* __restrict__ buf = a.r.buf;
    const uint64_t nbytes = a.r.nbytes;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lt = dm_lanemask_lt();
    uint32_t n = a.counts[1];
    if (n > a.field_cap) n = a.field_cap;
    const uint32_t warps_total = gridDim.x * DMS_WARPS;
    const uint32_t warp_id = blockIdx.x * DMS_WARPS + (threadIdx.x >> 5);
    for (uint32_t start = warp_id * 32u; start < n; start += warps_total * 32u) {
        const uint32_t i = start + lane;
        bool cand = false;
        uint64_t ckey = 0;
        uint32_t cq = 0, ck = 0, cg = 0;
        if (i < n) {
            const DmField f = a.fields[i];
            if ((uint64_t)f.g >= a.r.line_lo && (uint64_t)f.g < a.r.line_hi) {
                const uint64_t fp = dm_hash_value(buf, nbytes, f.vpos);
                const uint64_t key = dm_make_key(fp, sk.salt[f.k]);
                const bool known = TRAIN ? dm_table_contains_volatile(a.r.table, key) : dm_table_contains(a.r.table, key);
                cand = !known;
                ckey = key; cq = f.vpos - 1; ck = f.k; cg = f.g;
            }
        }
        uint32_t cb = __ballot_sync(0xffffffffu, cand);
        while (cb) {
            const int L = __ffs(cb) - 1;
            cb &= cb - 1;
            const uint32_t vq = __shfl_sync(0xffffffffu, cq, L);
            const uint32_t vk = __shfl_sync(0xffffffffu, ck, L);
            if (TRAIN) {
                const int still_new = ((int)lane == L) ? (dm_table_contains_volatile(a.r.table, ckey) ? 0 : 1) : 0;
                if (!__shfl_sync(0xffffffffu, still_new, L)) continue;
            }
            uint32_t ls = 0;
            const bool ok = dm_verify_field_warp(buf, vq, vk, sk, lane, lt, &ls);
            if ((int)lane == L && ok) {
                if (TRAIN) {
                    dm_table_insert(a.r.table, ckey, &a.r.hdr->error);
                } else {
                    bool first = false;
                    if (cg < a.r.out_cap) {
                        const float old = atomicAdd(a.r.scores + cg, 1.0f);
                        a.r.flags[cg] = 1;
                        first = old == 0.0f;
                    }
                    atomicAdd(a.r.stats + 8 + ck, 1ull);
                    atomicAdd(a.r.stats + 4, 1ull);
                    if (first) { atomicAdd(&a.r.hdr->n_anomalies, 1ull); atomicAdd(a.r.stats + 3, 1ull); }
                    const unsigned int idx = atomicAdd(&a.r.hdr->anomaly_list_count, 1u);
                    if (idx < a.r.anomaly_cap) {
                        dm_anomaly_t r;
                        r.line = cg; r.mask = 1u << ck; r.offset = ls;
                        a.r.anomalies[idx] = r;
                    }
                }
            }
        }
    }
}

#ifndef DM_EMU
// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
struct DmStagedScratch {
    DmCand* d_cand = nullptr;
    DmField* d_fields = nullptr;
    unsigned int* d_counts = nullptr;
    uint32_t cand_cap = 0, field_cap = 0;
    int grid = 0;
};

static inline int dm_staged_scratch_create(DmStagedScratch* s, uint64_t max_batch_bytes, int sm_count) {
    // worst cases: every byte a '=' (candidates); every other byte (a field is at least "k=")
    s->cand_cap = (uint32_t)(max_batch_bytes + 1024);
    s->field_cap = (uint32_t)(max_batch_bytes / 2 + 1024);
    if (cudaMalloc(&s->d_cand, (uint64_t)s->cand_cap * sizeof(DmCand)) != cudaSuccess) return DM_ERR_CUDA;
    if (cudaMalloc(&s->d_fields, (uint64_t)s->field_cap * sizeof(DmField)) != cudaSuccess) return DM_ERR_CUDA;
    if (cudaMalloc(&s->d_counts, 4 * sizeof(unsigned int)) != cudaSuccess) return DM_ERR_CUDA;
    if (cudaMemset(s->d_counts, 0, 4 * sizeof(unsigned int)) != cudaSuccess) return DM_ERR_CUDA;
    s->grid = sm_count * 8;
    return DM_OK;
}

static inline void dm_staged_scratch_destroy(DmStagedScratch* s) {
    cudaFree(s->d_cand); cudaFree(s->d_fields); cudaFree(s->d_counts);
    s->d_cand = nullptr; s->d_fields = nullptr; s->d_counts = nullptr;
}

// Enqueue the staged kernels for one message.  Returns kernels launched or < 0.
static inline int dm_staged_launch(DmStagedScratch* s, DmRowsScratch* rs, const uint8_t* d_buf, uint64_t nbytes,
                                   uint64_t n_train_lines, const DmKeys* d_keys, DmTable table, uint8_t* d_flags,
                                   float* d_scores, uint64_t out_cap, dm_anomaly_t* d_anoms, uint32_t anomaly_cap,
                                   DmBatchHeader* d_hdr, unsigned long long* d_stats, uint64_t max_lines, cudaStream_t st,
                                   void (*mark)(void*, cudaStream_t, int), void* mark_ctx) {
    const uint32_t n_rows = (uint32_t)((nbytes + DMR_ROW - 1) / DMR_ROW);
    if (n_rows == 0) return 0;
    if (n_rows > rs->max_rows) return DM_ERR_CAPACITY;
    DmStagedArgs a;
    a.r.buf = d_buf; a.r.nbytes = nbytes; a.r.n_rows = n_rows;
    a.r.n_tiles = (n_rows + DMR_TILE_ROWS - 1) / DMR_TILE_ROWS;
    a.r.row_prefix = rs->d_row_prefix; a.r.tile_state = rs->d_tile_state;
    rs->epoch = (rs->epoch % 0x3FFFFFFEu) + 1u;
    a.r.epoch = rs->epoch;
    a.r.keys = d_keys; a.r.table = table; a.r.flags = d_flags; a.r.scores = d_scores; a.r.out_cap = out_cap;
    a.r.anomalies = d_anoms; a.r.anomaly_cap = anomaly_cap; a.r.hdr = d_hdr; a.r.stats = d_stats;
    a.r.row_ctr = rs->d_row_ctr; a.r.ctr_base = rs->ctr_base; a.r.n_train_lines = n_train_lines; a.r.max_lines = max_lines;
    a.r.line_lo = 0; a.r.line_hi = ~0ull;
    a.cand = s->d_cand; a.fields = s->d_fields; a.counts = s->d_counts; a.cand_cap = s->cand_cap; a.field_cap = s->field_cap;
    int launched = 0;
    a.r.line_start = nullptr; a.r.group = DMR_GROUP; a.r.static_rows = 0; a.r.timeline = nullptr;
    a.r.aux_counts = s->d_counts;                      // cleared by tile 0 of K_A
    dm_k_rowindex<<<a.r.n_tiles, DMR_A_THREADS, 0, st>>>(a.r);
    const uint32_t tiles1 = (n_rows + DMS_ROWS_PER_CTA - 1) / DMS_ROWS_PER_CTA;
    const int grid1 = (int)(tiles1 < (uint32_t)s->grid ? tiles1 : (uint32_t)s->grid);
    if (mark) mark(mark_ctx, st, 0);
    dm_k_stage1<<<grid1, DMS_THREADS, 0, st>>>(a);
    if (mark) mark(mark_ctx, st, 1);
    dm_k_stage2<<<s->grid, DMS_THREADS, 0, st>>>(a);
    launched += 3;
    if (n_train_lines > 0) {
        a.r.line_lo = 0; a.r.line_hi = n_train_lines;
        dm_k_stage3<true><<<s->grid, DMS_THREADS, 0, st>>>(a);
        ++launched;
    }
    a.r.line_lo = n_train_lines; a.r.line_hi = ~0ull;
    dm_k_stage3<false><<<s->grid, DMS_THREADS, 0, st>>>(a);
    ++launched;
    if (cudaGetLastError() != cudaSuccess) return DM_ERR_CUDA;
    return launched;
}
#endif  // DM_EMU
