// dm_kernels_rows.cuh -- the "rows" variant of the fused tokenizer + detector (sm_100a).
//
// Same rules and same device helpers as dm_kernels_tile.cuh (key identification, dm_fp64 of
// a value, warp-cooperative exact re-check); different decomposition:
//
//   K_A  dm_k_rowindex   streams the message once from HBM (16-byte loads), counts the '\n'
//        of every 512-byte row, turns the counts into a global exclusive prefix
//        (row_prefix[r] = number of '\n' in front of row r; tiles of 128 rows, decoupled
//        look-back with 128-wide windows), zero-fills flags / scores of the records that end
//        in its tile and writes the batch header.  The record index of ANY byte is then
//        row_prefix[row] + (number of '\n' in front of it inside its row): no ownership of
//        records by warps, no extension rows.
//   K_B  dm_k_rows       every warp takes a static share of the rows and then fetches groups of
//        rows from an atomic counter (the message is still in L2: 16 MiB << 126 MB) and treats
//        each row on its own: classify '\n' / '=', compact the '=' into queue 1, identify keys
//        (stage 1), hash + probe values (stage 2), re-check the rare unknown values exactly,
//        apply alerts with atomics.  No block-level synchronisation after the key tables are
//        loaded.
//   The kernels of consecutive steps overlap through programmatic dependent launch (dm_pdl_wait):
//   K_B is scheduled while K_A runs, the next step's K_A streams its message in while K_B drains.
#pragma once
#include "dm_kernels_tile.cuh"

#define DMR_ROW 512u
#define DMR_TILE_ROWS 64u
#define DMR_A_THREADS 256
#define DMR_B_WARPS 8
#define DMR_B_THREADS (DMR_B_WARPS * 32)
#define DMR_GROUP 2u          // rows fetched per atomic (after the static share)
#define DMR_Q1CAP 128u
#define DMR_Q2CAP 64u

struct DmRowsArgs {
    const uint8_t* buf;
    uint64_t nbytes;
    uint32_t n_rows;
    uint32_t n_tiles;
    uint32_t* row_prefix;             // n_rows entries
    unsigned long long* tile_state;   // look-back words of K_A
    uint32_t epoch;
    const DmKeys* keys;
    DmTable table;
    uint8_t* flags;
    float* scores;
    uint64_t out_cap;
    dm_anomaly_t* anomalies;
    uint32_t anomaly_cap;
    DmBatchHeader* hdr;
    unsigned long long* stats;
    unsigned long long* row_ctr;      // monotonically increasing across launches
    unsigned long long ctr_base;
    uint64_t line_lo, line_hi;        // K_B handles records with index in [lo, hi)
    uint64_t n_train_lines;
    uint64_t max_lines;
    unsigned int* aux_counts;         // staged variant: list counters cleared by K_A (else NULL)
    uint32_t group;                   // rows fetched per atomic by a K_B warp
    uint32_t static_rows;             // rows every K_B warp takes without asking (warp w: rows [w*S, (w+1)*S))
    unsigned long long* timeline;     // diagnostics (DM_ROWS_TIMELINE): per K_B warp {smid, t_first, t_work_end, t_exit} in ns, else NULL
    uint32_t* line_start;             // lanes variant: K_A also writes the record index (else NULL):
                                      // line_start[g] = first byte of record g, line_start[n] = end sentinel
};

// Programmatic dependent launch (PDL).  The two kernels of a step and the first kernel of the
// next step are launched with cudaLaunchAttributeProgrammaticStreamSerialization: a kernel may
// be scheduled while its predecessor in the stream is still running, does whatever does not
// depend on it (K_B: key tables to shared memory; K_A: stream the rows in and count), and calls
// dm_pdl_wait() before it touches anything the predecessor reads or writes.  Launched without the
// attribute both calls are no-ops.
__device__ __forceinline__ void dm_pdl_wait() {
#ifndef DM_EMU
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
__device__ __forceinline__ void dm_pdl_launch_dependents() {
#ifndef DM_EMU
    asm volatile("griddepcontrol.launch_dependents;");
#endif
}

// 16-bit '\n' mask of this lane's chunk of a row, slack bytes behind the message dropped
__device__ __forceinline__ uint32_t dm_row_nl_mask(const uint4& v, uint64_t off, uint64_t nbytes) {
    const uint32_t f0 = dm_eqflags(v.x, 0x0A0A0A0Au), f1 = dm_eqflags(v.y, 0x0A0A0A0Au);
    const uint32_t f2 = dm_eqflags(v.z, 0x0A0A0A0Au), f3 = dm_eqflags(v.w, 0x0A0A0A0Au);
    if ((f0 | f1 | f2 | f3) == 0) return 0;
    uint32_t m = dm_flags_to_nib(f0) | (dm_flags_to_nib(f1) << 4) | (dm_flags_to_nib(f2) << 8) | (dm_flags_to_nib(f3) << 12);
    if (off + 16 > nbytes) m &= (1u << (uint32_t)(nbytes - off)) - 1u;
    return m;
}

// ---------------------------------------------------------------------------------------
// K_A: row index
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DMR_A_THREADS) dm_k_rowindex(DmRowsArgs a) {
    __shared__ uint32_t s_rowcnt[DMR_TILE_ROWS];
    __shared__ uint32_t s_rowpre[DMR_TILE_ROWS];      // '\n' in front of each row, inside the tile
    __shared__ unsigned long long s_base;
    __shared__ uint32_t s_total;
    const uint8_t* __restrict__ buf = a.buf;
    const uint64_t nbytes = a.nbytes;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t tile = blockIdx.x;

    dm_pdl_launch_dependents();                       // the detect kernel of this step may be scheduled now
    // newline count of each row of the tile: 8 warps x 8 rows, all 8 loads of a warp in flight
    uint32_t nlm[8];                                  // this lane's '\n' masks (kept for the record index)
    {
        const uint32_t rr = warp * 8;
        uint4 v[8];
        uint64_t off[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            off[i] = ((uint64_t)tile * DMR_TILE_ROWS + rr + i) * DMR_ROW + (uint64_t)lane * 16;
            v[i] = make_uint4(0, 0, 0, 0);
            if (off[i] < nbytes) v[i] = __ldg(reinterpret_cast<const uint4*>(buf + off[i]));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint32_t m = off[i] < nbytes ? dm_row_nl_mask(v[i], off[i], nbytes) : 0u;
            nlm[i] = m;
            const uint32_t c = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(m));
            if (lane == 0) s_rowcnt[rr + i] = c;
        }
    }
    // Everything up to and including the look-back below only reads the message and touches
    // tile_state, which no other kernel uses (epoch-tagged): under PDL it runs while the previous
    // step's detect kernel is still busy.  dm_pdl_wait() comes right before the first write to
    // anything that kernel reads or writes (row_prefix, header, outputs).
    __syncthreads();

    if (warp == 0) {
        // exclusive prefix of the 64 row counts (2 per lane), then the tile's global base
        uint32_t c[2], lane_sum = 0;
#pragma unroll
        for (int i = 0; i < 2; ++i) { c[i] = s_rowcnt[lane * 2 + i]; lane_sum += c[i]; }
        uint32_t incl = lane_sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
            if ((int)lane >= d) incl += y;
        }
        const uint32_t agg = __shfl_sync(0xffffffffu, incl, 31);
        const unsigned long long tag = (unsigned long long)a.epoch << 34;
        unsigned long long excl = 0;
        if (tile > 0) {
            if (lane == 0) atomicExch(a.tile_state + tile, tag | (DMT_ST_AGG << 32) | agg);
            long long hi = (long long)tile - 1;
            bool done = false;
            while (!done) {
                unsigned long long st[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const long long idx = hi - 32 * j - (long long)lane;
                    st[j] = idx >= 0 ? *((volatile unsigned long long*)(a.tile_state + idx)) : 0ull;
                }
                int consumed = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (done || consumed < j) continue;
                    const long long idx = hi - 32 * j - (long long)lane;
                    unsigned long long w = st[j];
                    if (idx >= 0 && (w >> 34) != a.epoch) w = 0;
                    const uint32_t status = idx >= 0 ? (uint32_t)((w >> 32) & 3u) : (uint32_t)DMT_ST_PREFIX;
                    const uint32_t not_ready = __ballot_sync(0xffffffffu, status == 0);
                    const uint32_t is_pref = __ballot_sync(0xffffffffu, status == DMT_ST_PREFIX);
                    const uint32_t first_pref = is_pref ? (uint32_t)(__ffs(is_pref) - 1) : 32u;
                    const uint32_t upto = first_pref < 32u ? first_pref : 31u;
                    const uint32_t win = upto == 31u ? 0xffffffffu : ((2u << upto) - 1u);
                    if (not_ready & win) continue;
                    const uint32_t val = (lane <= upto && idx >= 0) ? (uint32_t)w : 0u;
                    excl += __reduce_add_sync(0xffffffffu, val);
                    consumed = j + 1;
                    if (first_pref < 32u) done = true;
                }
                hi -= 32 * consumed;
                if (!done && consumed < 4) __nanosleep(20);
            }
        }
        if (lane == 0) atomicExch(a.tile_state + tile, tag | (DMT_ST_PREFIX << 32) | (unsigned long long)((uint32_t)excl + agg));
        dm_pdl_wait();
        uint32_t run = (uint32_t)excl + (incl - lane_sum);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const uint32_t row = tile * DMR_TILE_ROWS + lane * 2 + i;
            if (row < a.n_rows) a.row_prefix[row] = run;
            s_rowpre[lane * 2 + i] = run - (uint32_t)excl;
            run += c[i];
        }
        if (lane == 0) { s_base = excl; s_total = agg; }
        if (tile == a.n_tiles - 1 && lane == 0) {
            const unsigned long long nl = excl + agg;
            const bool tail = nbytes > 0 && buf[nbytes - 1] != 0x0Au;
            unsigned long long n_lines = nl + (tail ? 1ull : 0ull);
            // the last tile (the only one that knows the totals) initialises the whole batch header
            a.hdr->n_anomalies = 0; a.hdr->anomaly_list_count = 0;
            a.hdr->error = (n_lines > a.max_lines || n_lines > a.out_cap) ? DM_DEVERR_TOO_MANY_LINES : 0u;
            a.hdr->n_newlines = nl;
            a.hdr->n_lines = n_lines;
            if (a.aux_counts) { a.aux_counts[0] = 0; a.aux_counts[1] = 0; a.aux_counts[2] = 0; }
            const unsigned long long tr = a.n_train_lines < n_lines ? a.n_train_lines : n_lines;
            a.stats[0] += n_lines;
            a.stats[1] += tr;
            a.stats[2] += n_lines - tr;
            a.stats[5] += nbytes;
            if (tail) s_total = agg + 1;            // the unterminated last record is zero-filled here too
        }
    }
    __syncthreads();
    dm_pdl_wait();                                    // (warp 0 has waited already; the other warps write below)
    const unsigned long long base = s_base;
    if (a.line_start) {
        // lanes variant: the record index.  The '\n' at byte x with k '\n' in front of it ends
        // record k, so record k+1 starts at x+1.  (Every record's outputs are written by its own
        // lane there, no zero-fill.)
        if (tile == 0 && threadIdx.x == 0) a.line_start[0] = 0;
        if (tile == a.n_tiles - 1 && threadIdx.x == 0 && nbytes > 0 && buf[nbytes - 1] != 0x0Au) {
            const unsigned long long n_lines = base + s_total;          // s_total counts the unterminated record
            if (n_lines <= a.max_lines) a.line_start[n_lines] = (uint32_t)(nbytes + 1);
        }
        const uint32_t rr = warp * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint32_t m = nlm[i];
            // records in front of this lane's chunk: tile base + row prefix + lower lanes of the row
            uint32_t incl = (uint32_t)__popc(m);
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
                if ((int)lane >= d) incl += y;
            }
            unsigned long long k = base + s_rowpre[rr + i] + (incl - (uint32_t)__popc(m));
            const uint64_t off = ((uint64_t)tile * DMR_TILE_ROWS + rr + i) * DMR_ROW + (uint64_t)lane * 16;
            while (m) {
                const uint32_t b = (uint32_t)__ffs(m) - 1u;
                m &= m - 1;
                ++k;
                if (k <= a.max_lines) a.line_start[k] = (uint32_t)(off + b + 1);
            }
        }
        return;
    }
    // zero-fill the outputs of the records that end in this tile
    const uint32_t total = s_total;
    for (uint32_t i = threadIdx.x; i < total; i += DMR_A_THREADS) {
        const unsigned long long g = base + i;
        if (g < a.out_cap) { a.flags[g] = 0; a.scores[g] = 0.0f; }
    }
}

// ---------------------------------------------------------------------------------------
// K_B: rows
// ---------------------------------------------------------------------------------------
struct DmRQ1Entry { uint32_t q; uint32_t g; };
struct DmRQ2Entry { uint32_t vpos; uint32_t g; uint32_t k; };

template <bool TRAIN, bool RANGE>
__global__ void __launch_bounds__(DMR_B_THREADS) dm_k_rows(DmRowsArgs a) {
    __shared__ DmKeys sk;   // 16-byte aligned through its alignas(16) member
    __shared__ DmRQ1Entry s_q1[DMR_B_WARPS][DMR_Q1CAP];
    __shared__ DmRQ2Entry s_q2[DMR_B_WARPS][DMR_Q2CAP];
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(a.keys);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&sk);
        for (uint32_t i = threadIdx.x; i < sizeof(DmKeys) / 4; i += DMR_B_THREADS) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    dm_pdl_wait();                                    // K_A (and, for the detect pass, the training pass) are complete
    const uint8_t* __restrict__ buf = a.buf;
    const uint64_t nbytes = a.nbytes;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt = dm_lanemask_lt();
    DmRQ1Entry* q1 = s_q1[warp];
    DmRQ2Entry* q2 = s_q2[warp];
    uint32_t q1h = 0, q1n = 0, q2h = 0, q2n = 0;

    // stage 2: one identified field per lane -- fingerprint, probe, exact re-check when unknown
    auto drain2 = [&](uint32_t n) {
        bool cand = false;
        uint64_t ckey = 0;
        uint32_t cq = 0, ck = 0, cg = 0;
        if (lane < n) {
            const DmRQ2Entry e = q2[(q2h + lane) & (DMR_Q2CAP - 1)];
            const uint64_t fp = dm_hash_value(buf, nbytes, e.vpos);
            const uint64_t key = dm_make_key(fp, sk.salt[e.k]);
            const bool known = TRAIN ? dm_table_contains_volatile(a.table, key) : dm_table_contains(a.table, key);
            cand = !known;
            ckey = key; cq = e.vpos - 1; ck = e.k; cg = e.g;
        }
        q2h += n;
        q2n -= n;
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
                    // an alert: unknown value in monitored field ck of record cg
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
        __syncwarp();
    };

    // stage 1: one '=' per lane -- which monitored key (if any) stands in front of it
    auto drain1 = [&](uint32_t n) {
        bool matched = false;
        DmRQ2Entry qe;
        qe.vpos = 0; qe.g = 0; qe.k = 0;
        if (lane < n) {
            const DmRQ1Entry e = q1[(q1h + lane) & (DMR_Q1CAP - 1)];
            if (!RANGE || e.g != 0xFFFFFFFFu) {            // 0xFFFFFFFF: record outside this launch's range
                const int k = dm_key_identify(buf, (uint64_t)e.q, sk);
                if (k >= 0) { matched = true; qe.vpos = e.q + 1; qe.g = e.g; qe.k = (uint32_t)k; }
            }
        }
        q1h += n;
        q1n -= n;
        const uint32_t mb = __ballot_sync(0xffffffffu, matched);
        if (mb) {
            if (matched) q2[(q2h + q2n + __popc(mb & lt)) & (DMR_Q2CAP - 1)] = qe;
            q2n += __popc(mb);
            __syncwarp();
            if (q2n >= 32u) drain2(32u);
        }
        __syncwarp();
    };

    // Row distribution.  A 64k-record message is only ~4.6 rows per warp of the grid, so every
    // warp first takes a STATIC share (a.static_rows contiguous rows, no atomics, no start-up
    // convoy on the counter); the rest is handed out dynamically in groups of a.group rows, and
    // the fetch for the following group is issued BEFORE the current one is processed, which
    // takes the same-address atomic's latency off the critical path.  Every warp does exactly
    // (groups it received) + 1 atomics, so the host can account for the counter.
    const uint32_t wid = blockIdx.x * DMR_B_WARPS + warp;
#ifndef DM_EMU
    unsigned long long t_first = 0;
    if (a.timeline) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_first));
#endif
    const unsigned long long dyn_base = (unsigned long long)gridDim.x * DMR_B_WARPS * a.static_rows;
    unsigned long long first_row = (unsigned long long)wid * a.static_rows;
    uint32_t take = a.static_rows;
    for (;;) {
        unsigned long long next = 0;
        if (lane == 0) next = dyn_base + (atomicAdd(a.row_ctr, (unsigned long long)a.group) - a.ctr_base);
        const uint32_t r_end = (uint32_t)(first_row + take < a.n_rows ? first_row + take : a.n_rows);
        for (uint32_t row = (uint32_t)(first_row < a.n_rows ? first_row : a.n_rows); row < r_end; ++row) {
            const uint64_t off = (uint64_t)row * DMR_ROW + (uint64_t)lane * 16;
            uint32_t nl16 = 0, eq16 = 0;
            if (off < nbytes) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(buf + off));
                nl16 = dm_row_nl_mask(v, off, nbytes);
                const uint32_t e0 = dm_eqflags(v.x, 0x3D3D3D3Du), e1 = dm_eqflags(v.y, 0x3D3D3D3Du);
                const uint32_t e2 = dm_eqflags(v.z, 0x3D3D3D3Du), e3 = dm_eqflags(v.w, 0x3D3D3D3Du);
                if (e0 | e1 | e2 | e3) {
                    eq16 = dm_flags_to_nib(e0) | (dm_flags_to_nib(e1) << 4) | (dm_flags_to_nib(e2) << 8) | (dm_flags_to_nib(e3) << 12);
                    if (off + 16 > nbytes) eq16 &= (1u << (uint32_t)(nbytes - off)) - 1u;
                }
            }
            // record index in front of this lane's chunk
            const uint32_t b_nl = __ballot_sync(0xffffffffu, nl16 != 0);
            uint32_t pre = 0;
            if (b_nl) {
                const uint32_t my = (uint32_t)__popc(nl16);
                const uint32_t b_multi = __ballot_sync(0xffffffffu, my > 1);
                if (b_multi == 0) {
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
            const uint32_t g_chunk = a.row_prefix[row] + pre;

            // every '=' of the row (whose record is in this launch's range) goes to queue 1
            uint32_t m = eq16;
            // drop the '=' whose record lies outside [line_lo, line_hi): cheap pre-filter per lane
            if (RANGE && m) {
                const uint32_t g_lo = g_chunk, g_hi = g_chunk + (uint32_t)__popc(nl16);
                if ((uint64_t)g_hi < a.line_lo || (uint64_t)g_lo >= a.line_hi) m = 0;
            }
            const uint32_t my_eq = (uint32_t)__popc(m);
            uint32_t eincl = my_eq;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, eincl, d);
                if ((int)lane >= d) eincl += y;
            }
            const uint32_t row_eq = __shfl_sync(0xffffffffu, eincl, 31);
            if (row_eq == 0) continue;
            if (q1n + row_eq <= DMR_Q1CAP) {
                uint32_t slot = q1h + q1n + (eincl - my_eq);
                while (m) {
                    const uint32_t j = (uint32_t)__ffs(m) - 1;
                    m &= m - 1;
                    const uint32_t g = g_chunk + (uint32_t)__popc(nl16 & ((1u << j) - 1u));
                    DmRQ1Entry e;
                    e.q = (uint32_t)off + j;
                    e.g = (!RANGE || ((uint64_t)g >= a.line_lo && (uint64_t)g < a.line_hi)) ? g : 0xFFFFFFFFu;
                    q1[slot & (DMR_Q1CAP - 1)] = e;
                    ++slot;
                }
                q1n += row_eq;
                __syncwarp();
                while (q1n >= 32u) drain1(32u);
            } else {
                // a row dense in '=' (more than the queue holds): one per lane per round
                while (__ballot_sync(0xffffffffu, m != 0)) {
                    bool has = false;
                    DmRQ1Entry e;
                    e.q = 0; e.g = 0;
                    if (m) {
                        const uint32_t j = (uint32_t)__ffs(m) - 1;
                        m &= m - 1;
                        const uint32_t g = g_chunk + (uint32_t)__popc(nl16 & ((1u << j) - 1u));
                        has = true;
                        e.q = (uint32_t)off + j;
                        e.g = (!RANGE || ((uint64_t)g >= a.line_lo && (uint64_t)g < a.line_hi)) ? g : 0xFFFFFFFFu;
                    }
                    const uint32_t hb = __ballot_sync(0xffffffffu, has);
                    if (has) q1[(q1h + q1n + __popc(hb & lt)) & (DMR_Q1CAP - 1)] = e;
                    q1n += __popc(hb);
                    __syncwarp();
                    while (q1n >= 32u) drain1(32u);
                }
            }
        }
        next = __shfl_sync(0xffffffffu, next, 0);
        if (next >= a.n_rows) break;
        first_row = next;
        take = a.group;
    }
    dm_pdl_launch_dependents();                       // next step's K_A may start streaming its message in
#ifndef DM_EMU
    unsigned long long t_work = 0;
    if (a.timeline) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_work));
#endif
    while (q1n) drain1(q1n < 32u ? q1n : 32u);
    while (q2n) drain2(q2n < 32u ? q2n : 32u);
#ifndef DM_EMU
    if (a.timeline && lane == 0) {
        unsigned long long t_exit;
        uint32_t smid;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_exit));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        unsigned long long* o = a.timeline + 4ull * wid;
        o[0] = smid; o[1] = t_first; o[2] = t_work; o[3] = t_exit;
    }
#endif
}

#ifndef DM_EMU
// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
struct DmRowsScratch {
    uint32_t* d_row_prefix = nullptr;
    unsigned long long* d_tile_state = nullptr;
    unsigned long long* d_row_ctr = nullptr;
    uint64_t max_rows = 0, max_tiles = 0;
    unsigned long long ctr_base = 0;
    uint32_t epoch = 0;
    int grid_b = 0;
    uint32_t group = DMR_GROUP;       // DM_ROWS_GROUP
    uint32_t static_pct = 65;         // DM_ROWS_STATIC: share of the rows distributed statically
    bool pdl = true;                  // DM_PDL=0 switches programmatic dependent launch off
    unsigned long long* d_timeline = nullptr;   // DM_ROWS_TIMELINE=1: 4 words per K_B warp of the last launch
    int last_grid = 0;
};

static inline int dm_rows_scratch_create(DmRowsScratch* s, uint64_t max_batch_bytes, int sm_count) {
    s->max_rows = (max_batch_bytes + DMR_ROW - 1) / DMR_ROW + 1;
    s->max_tiles = (s->max_rows + DMR_TILE_ROWS - 1) / DMR_TILE_ROWS + 1;
    if (cudaMalloc(&s->d_row_prefix, s->max_rows * sizeof(uint32_t)) != cudaSuccess) return DM_ERR_CUDA;
    if (cudaMalloc(&s->d_tile_state, s->max_tiles * sizeof(unsigned long long)) != cudaSuccess) return DM_ERR_CUDA;
    if (cudaMemset(s->d_tile_state, 0, s->max_tiles * sizeof(unsigned long long)) != cudaSuccess) return DM_ERR_CUDA;
    if (cudaMalloc(&s->d_row_ctr, sizeof(unsigned long long)) != cudaSuccess) return DM_ERR_CUDA;
    if (cudaMemset(s->d_row_ctr, 0, sizeof(unsigned long long)) != cudaSuccess) return DM_ERR_CUDA;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dm_k_rows<false, false>, DMR_B_THREADS, 0) != cudaSuccess) return DM_ERR_CUDA;
    if (per_sm < 1) per_sm = 1;
    const char* pd = getenv("DM_PDL");
    if (pd && atoi(pd) == 0) s->pdl = false;
    const char* tl = getenv("DM_ROWS_TIMELINE");           // diagnostics: per-warp start / end times of K_B
    if (tl && atoi(tl) == 1) {
        if (cudaMalloc(&s->d_timeline, (size_t)sm_count * 16 * DMR_B_WARPS * 4 * sizeof(unsigned long long)) != cudaSuccess) return DM_ERR_CUDA;
    }
    const char* sp = getenv("DM_ROWS_STATIC");             // tuning knob: % of the rows distributed statically
    if (sp && atoi(sp) >= 0 && atoi(sp) <= 100) s->static_pct = (uint32_t)atoi(sp);
    const char* grp = getenv("DM_ROWS_GROUP");             // tuning knob: rows per atomic fetch
    if (grp && atoi(grp) >= 1 && atoi(grp) <= 64) s->group = (uint32_t)atoi(grp);
    const char* cap = getenv("DM_ROWS_CTAS_PER_SM");       // tuning knob: fewer, longer-lived warps
    if (cap && atoi(cap) > 0 && atoi(cap) < per_sm) per_sm = atoi(cap);
    s->grid_b = sm_count * per_sm;
    return DM_OK;
}

static inline void dm_rows_scratch_destroy(DmRowsScratch* s) {
    cudaFree(s->d_row_prefix);
    cudaFree(s->d_tile_state);
    cudaFree(s->d_row_ctr);
    s->d_row_prefix = nullptr; s->d_tile_state = nullptr; s->d_row_ctr = nullptr;
}

// Enqueue K_A and K_B for one message.  Returns the number of kernels launched or < 0.
// <<<>>> with the programmatic-stream-serialization attribute (see dm_pdl_wait)
template <typename... Args>
static inline void dm_launch_pdl_smem(void (*kernel)(Args...), unsigned grid, unsigned block, size_t smem, cudaStream_t st,
                                      bool pdl, Args... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid, 1, 1); cfg.blockDim = dim3(block, 1, 1); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, args...);
}
template <typename Arg>
static inline void dm_launch_pdl(void (*kernel)(Arg), unsigned grid, unsigned block, cudaStream_t st, bool pdl, Arg arg) {
    dm_launch_pdl_smem(kernel, grid, block, 0, st, pdl, arg);
}

static inline int dm_rows_launch(DmRowsScratch* s, const uint8_t* d_buf, uint64_t nbytes, uint64_t n_train_lines,
                                 const DmKeys* d_keys, DmTable table, uint8_t* d_flags, float* d_scores,
                                 uint64_t out_cap, dm_anomaly_t* d_anoms, uint32_t anomaly_cap, DmBatchHeader* d_hdr,
                                 unsigned long long* d_stats, uint64_t max_lines, cudaStream_t st,
                                 void (*mark)(void*, cudaStream_t, int), void* mark_ctx) {
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
    a.line_lo = 0; a.line_hi = ~0ull; a.ctr_base = s->ctr_base; a.aux_counts = nullptr; a.line_start = nullptr;
    int launched = 0;
    dm_launch_pdl(dm_k_rowindex, a.n_tiles, DMR_A_THREADS, st, s->pdl, a);
    ++launched;
    const uint32_t G = s->group;
    a.group = G;
    a.timeline = s->d_timeline;
    int grid = (int)((n_rows + DMR_B_WARPS - 1) / DMR_B_WARPS);
    if (grid > s->grid_b) grid = s->grid_b;
    if (grid < 1) grid = 1;
    s->last_grid = grid;
    const unsigned long long W = (unsigned long long)grid * DMR_B_WARPS;
    // static share: s->static_pct % of the rows, rounded down to whole rows per warp
    const uint32_t S = (uint32_t)(((unsigned long long)n_rows * s->static_pct / 100ull) / W);
    a.static_rows = S;
    const unsigned long long dyn_rows = n_rows > W * S ? n_rows - W * S : 0;
    const unsigned long long dyn_groups = (dyn_rows + G - 1) / G;
    // every warp's last fetch fails; all fetches advance the counter by G
    const unsigned long long per_launch = (dyn_groups + W) * G;
    if (n_train_lines > 0) {
        a.line_lo = 0; a.line_hi = n_train_lines; a.ctr_base = s->ctr_base;
        dm_launch_pdl(dm_k_rows<true, true>, (unsigned)grid, DMR_B_THREADS, st, s->pdl, a);
        s->ctr_base += per_launch;
        ++launched;
    }
    a.line_lo = n_train_lines; a.line_hi = ~0ull; a.ctr_base = s->ctr_base;
    if (mark) mark(mark_ctx, st, 0);
    if (n_train_lines > 0) dm_launch_pdl(dm_k_rows<false, true>, (unsigned)grid, DMR_B_THREADS, st, s->pdl, a);
    else dm_launch_pdl(dm_k_rows<false, false>, (unsigned)grid, DMR_B_THREADS, st, s->pdl, a);
    if (mark) mark(mark_ctx, st, 1);
    s->ctr_base += per_launch;
    ++launched;
    if (cudaGetLastError() != cudaSuccess) return DM_ERR_CUDA;
    return launched;
}
#endif  // DM_EMU
