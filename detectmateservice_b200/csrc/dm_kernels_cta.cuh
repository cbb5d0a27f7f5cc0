// dm_kernels_cta.cuh -- the "cta" variant of the main kernel (K_B): same rows decomposition
// as dm_kernels_rows.cuh (K_A builds the row prefix first), but the two expensive stages run
// on CTA-wide queues in shared memory so that they always work on FULL 256-entry chunks.
//
// Why: in dm_k_rows each warp drains private queues; a 16 MiB message gives a warp only ~100
// '=' and ~40 fields, so key identification runs with 25 and value hashing with 12 of 32
// lanes active (ncu, profiles/r01g).  Here a CTA of 8 warps proceeds in rounds, strictly
// barrier-synchronous (no spinning, nothing to deadlock):
//
//   round:  each warp takes one 512-byte row, classifies '\n' / '=' and appends up to 128 of
//           its '=' to queue 1                                             __syncthreads
//           while queue 1 holds >= 256: one '=' per THREAD -> key identification, matches
//           appended to queue 2                                            __syncthreads
//               while queue 2 holds >= 256: one field per THREAD -> dm_fp64, probe, exact
//               re-check of unknown values (per warp), alert / insert      __syncthreads
//   leftovers (< 256 entries) stay queued for the next round; the last round flushes them.
//
// Queue bounds: queue 1 <= 255 + 8 * 128, queue 2 <= 255 + 256; rows with more than 128 '='
// take several rounds.  Same helpers, same rules, same parity tests as the other variants
// (DM_KERNEL=cta).
#pragma once
#include "dm_kernels_rows.cuh"

#define DMC_THREADS 256
#define DMC_WARPS (DMC_THREADS / 32)
#define DMC_PUSH_MAX 128u
#define DMC_Q1CAP 2048u
#define DMC_Q2CAP 512u

template <bool TRAIN, bool RANGE>
__global__ void __launch_bounds__(DMC_THREADS) dm_k_cta(DmRowsArgs a) {
    __shared__ DmKeys sk;
    __shared__ DmRQ1Entry s_q1[DMC_Q1CAP];
    __shared__ DmRQ2Entry s_q2[DMC_Q2CAP];
    __shared__ unsigned int s_q1_tail, s_q2_tail;
    __shared__ unsigned long long s_row0;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(a.keys);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&sk);
        for (uint32_t i = threadIdx.x; i < sizeof(DmKeys) / 4; i += DMC_THREADS) dst[i] = __ldg(src + i);
        if (threadIdx.x == 0) { s_q1_tail = 0; s_q2_tail = 0; }
    }
    __syncthreads();
    const uint8_t* __restrict__ buf = a.buf;
    const uint64_t nbytes = a.nbytes;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt = dm_lanemask_lt();
    uint32_t q1_head = 0, q2_head = 0;            // identical in every thread

    // stage 2 on one chunk: thread t handles queue-2 entry q2_head + t (if < n)
    auto stage2 = [&](uint32_t n) {
        bool cand = false;
        uint64_t ckey = 0;
        uint32_t cq = 0, ck = 0, cg = 0;
        if (threadIdx.x < n) {
            const DmRQ2Entry e = s_q2[(q2_head + threadIdx.x) & (DMC_Q2CAP - 1)];
            const uint64_t fp = dm_hash_value(buf, nbytes, e.vpos);
            const uint64_t key = dm_make_key(fp, sk.salt[e.k]);
            const bool known = TRAIN ? dm_table_contains_volatile(a.table, key) : dm_table_contains(a.table, key);
            cand = !known;
            ckey = key; cq = e.vpos - 1; ck = e.k; cg = e.g;
        }
        uint32_t cb = __ballot_sync(0xffffffffu, cand);
        while (cb) {
            const int L = __ffs(cb) - 1;
            cb &= cb - 1;
            const uint32_t vq = __shfl_sync(0xffffffffu, cq, L);
            const uint32_t vk = __shfl_sync(0xffffffffu, ck, L);
            if (TRAIN) {
                const int still_new = ((int)lane == L) ? (dm_table_contains_volatile(a.table, ckey) ? 0 : 1) : 0;
                if (!__shfl_sync(0xffffffffu, still_new, L)) continue;
            }
            uint32_t ls = 0;
            const bool ok = dm_verify_field_warp(buf, vq, vk, sk, lane, lt, &ls);
            if ((int)lane == L && ok) {
                if (TRAIN) {
                    dm_table_insert(a.table, ckey, &a.hdr->error);
                } else {
                    bool first = false;
                    if (cg < a.out_cap) {
                        const float old = atomicAdd(a.scores + cg, 1.0f);
                        a.flags[cg] = 1;
                        first = old == 0.0f;
                    }
                    atomicAdd(a.stats + 8 + ck, 1ull);
                    atomicAdd(a.stats + 4, 1ull);
                    if (first) { atomicAdd(&a.hdr->n_anomalies, 1ull); atomicAdd(a.stats + 3, 1ull); }
                    const unsigned int idx = atomicAdd(&a.hdr->anomaly_list_count, 1u);
                    if (idx < a.anomaly_cap) {
                        dm_anomaly_t r;
                        r.line = cg; r.mask = 1u << ck; r.offset = ls;
                        a.anomalies[idx] = r;
                    }
                }
            }
        }
        q2_head += n;
    };

    // stage 1 on one chunk: thread t handles queue-1 entry q1_head + t (if < n); matches go to queue 2
    auto stage1 = [&](uint32_t n) {
        bool matched = false;
        DmRQ2Entry qe;
        qe.vpos = 0; qe.g = 0; qe.k = 0;
        if (threadIdx.x < n) {
            const DmRQ1Entry e = s_q1[(q1_head + threadIdx.x) & (DMC_Q1CAP - 1)];
            if (!RANGE || e.g != 0xFFFFFFFFu) {
                const int k = dm_key_identify(buf, (uint64_t)e.q, sk);
                if (k >= 0) { matched = true; qe.vpos = e.q + 1; qe.g = e.g; qe.k = (uint32_t)k; }
            }
        }
        const uint32_t mb = __ballot_sync(0xffffffffu, matched);
        if (mb) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&s_q2_tail, (unsigned int)__popc(mb));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (matched) s_q2[(base + __popc(mb & lt)) & (DMC_Q2CAP - 1)] = qe;
        }
        q1_head += n;
        __syncthreads();                                   // queue 2 entries visible
        const uint32_t t2 = s_q2_tail;                     // snapshot ...
        __syncthreads();                                   // ... taken by everybody before anyone pushes again
        while (t2 - q2_head >= DMC_THREADS) {
            stage2(DMC_THREADS);
            __syncthreads();                               // chunk consumed before its slots are reused
        }
    };

    for (;;) {
        if (threadIdx.x == 0) s_row0 = atomicAdd(a.row_ctr, (unsigned long long)DMC_WARPS) - a.ctr_base;
        __syncthreads();
        const unsigned long long row0 = s_row0;
        if (row0 >= a.n_rows) break;
        const uint32_t row = (uint32_t)row0 + warp;
        // ---- classify this warp's row ----
        uint32_t nl16 = 0, m = 0, g_chunk = 0;
        const uint64_t off = (uint64_t)row * DMR_ROW + (uint64_t)lane * 16;
        if (row < a.n_rows) {
            if (off < nbytes) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(buf + off));
                nl16 = dm_row_nl_mask(v, off, nbytes);
                const uint32_t e0 = dm_eqflags(v.x, 0x3D3D3D3Du), e1 = dm_eqflags(v.y, 0x3D3D3D3Du);
                const uint32_t e2 = dm_eqflags(v.z, 0x3D3D3D3Du), e3 = dm_eqflags(v.w, 0x3D3D3D3Du);
                if (e0 | e1 | e2 | e3) {
                    m = dm_flags_to_nib(e0) | (dm_flags_to_nib(e1) << 4) | (dm_flags_to_nib(e2) << 8) | (dm_flags_to_nib(e3) << 12);
                    if (off + 16 > nbytes) m &= (1u << (uint32_t)(nbytes - off)) - 1u;
                }
            }
            const uint32_t b_nl = __ballot_sync(0xffffffffu, nl16 != 0);
            uint32_t pre = 0;
            if (b_nl) {
                const uint32_t my = (uint32_t)__popc(nl16);
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
            g_chunk = __ldg(a.row_prefix + row) + pre;
            if (RANGE && m) {
                const uint32_t g_lo = g_chunk, g_hi = g_chunk + (uint32_t)__popc(nl16);
                if ((uint64_t)g_hi < a.line_lo || (uint64_t)g_lo >= a.line_hi) m = 0;
            }
        }
        // ---- rounds: push up to 128 '=' per warp, then run the stages on full chunks ----
        int more;
        do {
            const uint32_t my_eq = (uint32_t)__popc(m);
            uint32_t incl = my_eq;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
                if ((int)lane >= d) incl += y;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            const uint32_t take = total < DMC_PUSH_MAX ? total : DMC_PUSH_MAX;
            if (take) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(&s_q1_tail, take);
                base = __shfl_sync(0xffffffffu, base, 0);
                uint32_t idx = incl - my_eq;                   // rank of this lane's first '=' in the row
                while (m && idx < take) {
                    const uint32_t j = (uint32_t)__ffs(m) - 1;
                    m &= m - 1;
                    const uint32_t g = g_chunk + (uint32_t)__popc(nl16 & ((1u << j) - 1u));
                    DmRQ1Entry e;
                    e.q = (uint32_t)off + j;
                    e.g = (!RANGE || ((uint64_t)g >= a.line_lo && (uint64_t)g < a.line_hi)) ? g : 0xFFFFFFFFu;
                    s_q1[(base + idx) & (DMC_Q1CAP - 1)] = e;
                    ++idx;
                }
            }
            more = __syncthreads_or(m != 0);                   // also publishes the queue-1 entries
            const uint32_t t1 = s_q1_tail;                     // snapshot, frozen by the barrier below:
            __syncthreads();                                   // control flow must not depend on a tail that moves
            while (t1 - q1_head >= DMC_THREADS) stage1(DMC_THREADS);
        } while (more);
    }
    // ---- flush: whatever is left in the queues ----
    {
        const uint32_t n1 = s_q1_tail - q1_head;               // < 256
        stage1(n1);                                            // (also drains full queue-2 chunks)
        const uint32_t n2 = s_q2_tail - q2_head;               // < 256
        if (n2) stage2(n2);
    }
}

#ifndef DM_EMU
struct DmCtaScratch { int grid = 0; };

static inline int dm_cta_scratch_create(DmCtaScratch* s, int sm_count) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dm_k_cta<false, false>, DMC_THREADS, 0) != cudaSuccess) return DM_ERR_CUDA;
    if (per_sm < 1) per_sm = 1;
    const char* cap = getenv("DM_CTA_PER_SM");
    if (cap && atoi(cap) > 0 && atoi(cap) < per_sm) per_sm = atoi(cap);
    s->grid = sm_count * per_sm;
    return DM_OK;
}

// Enqueue K_A and the cta main kernel(s) for one message.  Returns kernels launched or < 0.
static inline int dm_cta_launch(DmCtaScratch* cs, DmRowsScratch* s, const uint8_t* d_buf, uint64_t nbytes, uint64_t n_train_lines,
                                const DmKeys* d_keys, DmTable table, uint8_t* d_flags, float* d_scores, uint64_t out_cap,
                                dm_anomaly_t* d_anoms, uint32_t anomaly_cap, DmBatchHeader* d_hdr, unsigned long long* d_stats,
                                uint64_t max_lines, cudaStream_t st, void (*mark)(void*, cudaStream_t, int), void* mark_ctx) {
    const uint32_t n_rows = (uint32_t)((nbytes + DMR_ROW - 1) / DMR_ROW);
    if (n_rows == 0) return 0;
    if (n_rows > s->max_rows) return DM_ERR_CAPACITY;
    DmRowsArgs a;
    a.buf = d_buf; a.nbytes = nbytes; a.n_rows = n_rows;
    a.n_tiles = (n_rows + DMR_TILE_ROWS - 1) / DMR_TILE_ROWS;
    a.row_prefix = s->d_row_prefix; a.tile_state = s->d_tile_state;
    s->epoch = (s->epoch % 0x3FFFFFFEu) + 1u;
    a.epoch = s->epoch;
    a.keys = d_keys; a.table = table; a.flags = d_flags; a.scores = d_scores; a.out_cap = out_cap;
    a.anomalies = d_anoms; a.anomaly_cap = anomaly_cap; a.hdr = d_hdr; a.stats = d_stats;
    a.row_ctr = s->d_row_ctr; a.n_train_lines = n_train_lines; a.max_lines = max_lines;
    a.line_lo = 0; a.line_hi = ~0ull; a.ctr_base = s->ctr_base; a.aux_counts = nullptr; a.line_start = nullptr; a.group = DMR_GROUP; a.static_rows = 0; a.timeline = nullptr;
    int launched = 0;
    dm_k_rowindex<<<a.n_tiles, DMR_A_THREADS, 0, st>>>(a);
    ++launched;
    const uint32_t rounds = (n_rows + DMC_WARPS - 1) / DMC_WARPS;
    int grid = (int)(rounds < (uint32_t)cs->grid ? rounds : (uint32_t)cs->grid);
    // every CTA ends with one failing fetch of DMC_WARPS rows
    const unsigned long long per_launch = (unsigned long long)rounds * DMC_WARPS + (unsigned long long)grid * DMC_WARPS;
    if (n_train_lines > 0) {
        a.line_lo = 0; a.line_hi = n_train_lines; a.ctr_base = s->ctr_base;
        dm_k_cta<true, true><<<grid, DMC_THREADS, 0, st>>>(a);
        s->ctr_base += per_launch;
        ++launched;
    }
    a.line_lo = n_train_lines; a.line_hi = ~0ull; a.ctr_base = s->ctr_base;
    if (mark) mark(mark_ctx, st, 0);
    if (n_train_lines > 0) dm_k_cta<false, true><<<grid, DMC_THREADS, 0, st>>>(a);
    else dm_k_cta<false, false><<<grid, DMC_THREADS, 0, st>>>(a);
    if (mark) mark(mark_ctx, st, 1);
    s->ctr_base += per_launch;
    ++launched;
    if (cudaGetLastError() != cudaSuccess) return DM_ERR_CUDA;
    return launched;
}
#endif  // DM_EMU
