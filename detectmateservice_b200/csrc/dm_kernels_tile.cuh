// dm_kernels_tile.cuh -- the fused single-pass tokenizer + detector kernel (sm_100a).
//
// Replaces, per record: MatcherParser field extraction + NewValueDetector.train/.detect of
// the un-vendored detectmatelibrary, which the reference drives one record at a time from
// Service.process (/root/reference/src/service/core.py:201-203).  Rules: DESIGN.md R-tok
// L1-L6 and R-spec 1-4.
//
// Shape of the computation (DESIGN.md "Fused kernel"):
//   * the message is cut into 32 KiB tiles; a CTA = 8 worker warps + 1 scanner warp takes
//     tiles from an atomic counter (persistent, in order);
//   * worker warp w owns the records that START inside its 4 KiB segment.  It reads 512-byte
//     rows with one coalesced 16-byte load per lane.  Pass 1 counts its record starts (the
//     rows stay in L1), pass 2 classifies '\n' and '=' bytes with SIMD-in-register compares
//     and compacts every '=' into a per-warp queue.  The queues keep all 32 lanes busy in the
//     expensive stages: stage 1 takes one '=' per lane and compares the 12 bytes in front of
//     it against every monitored key (word-parallel), stage 2 takes one identified field per
//     lane: dm_fp64 of its value, probe of the known-set table;
//   * quote parity (R-tok L2-L4) and first-occurrence-wins (L6) are NOT tracked on that
//     fast path.  They can only change the outcome for a value that is not in the table
//     (an alert, or an insert while training), so exactly those -- rare -- candidates are
//     re-checked by a sequential re-tokenisation of their record (dm_verify_field);
//   * the scanner warp turns the 8 record counts into global record indices with a
//     decoupled look-back over per-tile state words, while the workers are in pass 2;
//   * alerts are applied with atomics after the warp has zero-filled the flags / scores of
//     its own records.
#pragma once
#include "dm_device.cuh"

#define DMT_WARPS 4
#define DMT_THREADS (DMT_WARPS * 32 + 32)
#define DMT_ROW 512u
#define DMT_SEG_ROWS 8u
#define DMT_SEG (DMT_ROW * DMT_SEG_ROWS)
#define DMT_TILE (DMT_SEG * DMT_WARPS)
#define DMT_Q1CAP 128u     // '=' positions waiting for key identification (circular)
#define DMT_Q2CAP 64u      // identified fields waiting for value hashing (circular)
#define DMT_PCAP 64u       // alerts waiting for the record-index base (circular)

#define DMT_ST_AGG 1ull
#define DMT_ST_PREFIX 2ull

struct DmFusedArgs {
    const uint8_t* buf;
    uint64_t nbytes;
    uint32_t n_tiles;
    const DmKeys* keys;
    DmTable table;
    uint8_t* flags;
    float* scores;
    uint64_t out_cap;
    dm_anomaly_t* anomalies;
    uint32_t anomaly_cap;
    DmBatchHeader* hdr;
    unsigned long long* stats;
    unsigned long long* tile_state;   // one word per tile: epoch<<34 | status<<32 | value
    uint32_t epoch;
    unsigned long long* tile_ctr;     // counts the dynamically fetched tiles, across launches
    unsigned long long ctr_base;      // value of *tile_ctr when this launch starts
    uint64_t line_lo, line_hi;        // this launch handles records with index in [lo, hi)
    uint64_t n_train_lines;
    uint64_t max_lines;
    int range_check;                  // honour [line_lo, line_hi) (message holds training AND detection records)
    int zero_fill;                    // this launch zero-fills flags / scores
    int finalize;                     // this launch writes the batch header and the statistics
};

struct DmQ1Entry { uint32_t q; int32_t ln; };
struct DmQ2Entry { uint32_t vpos; int32_t ln; uint32_t k; };
struct DmPEntry { uint32_t ln; uint32_t k; uint32_t lstart; };

// 0x80 in every byte of w that equals the byte replicated in pat (pat bytes < 0x80).
__device__ __forceinline__ uint32_t dm_eqflags(uint32_t w, uint32_t pat) {
    const uint32_t x = w ^ pat;
    const uint32_t a = (x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
    return ~(a | x) & 0x80808080u;
}
// 0x80 flags (bits 7,15,23,31) -> 4-bit mask
__device__ __forceinline__ uint32_t dm_flags_to_nib(uint32_t f) { return (f * 0x00204081u) >> 28; }

__device__ __forceinline__ uint32_t dm_ld8(const uint8_t* __restrict__ buf, uint64_t p) { return __ldg(buf + p); }
__device__ __forceinline__ uint32_t dm_ld32(const uint8_t* __restrict__ buf, uint64_t p_aligned) {
    return __ldg(reinterpret_cast<const uint32_t*>(buf + p_aligned));
}

// Byte-wise check that key k ends right before the '=' at q and starts at a field start
// (R-tok L4 delimiter; quote parity is re-checked later, see file header).  The last
// `skip_tail` key bytes are known to match already.
__device__ __forceinline__ bool dm_key_check(const uint8_t* __restrict__ buf, uint64_t q, uint32_t k, const DmKeys& sk,
                                             uint32_t skip_tail) {
    const uint32_t len = sk.len[k];
    if (q < len) return false;
    const uint64_t st = q - len;
    if (st > 0) {
        const uint32_t c = dm_ld8(buf, st - 1);
        if (c != 0x20u && c != 0x27u && c != 0x0Au) return false;
    }
    const uint32_t n = len - skip_tail;
    for (uint32_t i = 0; i < n; ++i)
        if (dm_ld8(buf, st + i) != sk.bytes[k][i]) return false;
    return true;
}

// Which monitored key (if any) ends right before the '=' at q?  Word-parallel: the 12 bytes
// in front of q are compared against every key's precomputed patterns.
__device__ __forceinline__ int dm_key_identify(const uint8_t* __restrict__ buf, uint64_t q, const DmKeys& sk) {
    if (q < 12) {
        for (uint32_t k = 0; k < sk.n; ++k)
            if (dm_key_check(buf, q, k, sk, 0)) return (int)k;
        return -1;
    }
    const uint64_t a0 = (q - 12) & ~3ull;
    const uint32_t sh = (uint32_t)((q - 12) & 3) * 8;
    const uint32_t x0 = dm_ld32(buf, a0), x1 = dm_ld32(buf, a0 + 4), x2 = dm_ld32(buf, a0 + 8), x3 = dm_ld32(buf, a0 + 12);
    const uint32_t w_a = __funnelshift_r(x0, x1, sh);    // bytes q-12 .. q-9
    const uint32_t w_b = __funnelshift_r(x1, x2, sh);    // bytes q-8 .. q-5
    const uint32_t w_c = __funnelshift_r(x2, x3, sh);    // bytes q-4 .. q-1
    // first (= longest) key whose last min(len,8) bytes stand in front of the '='
    uint32_t slot = 0xFFFFFFFFu;
    for (uint32_t s = 0; s < sk.n; ++s) {
        const uint4 p = *reinterpret_cast<const uint4*>(sk.pat[s]);
        const uint32_t diff = ((w_c ^ p.x) & p.y) | ((w_b ^ p.z) & p.w);
        if (diff == 0 && slot == 0xFFFFFFFFu) slot = s;
    }
    if (slot == 0xFFFFFFFFu) return -1;
    const uint32_t k = sk.order[slot];
    const uint32_t len = sk.len[k];
    if (len <= 8) {
        // field-start delimiter (R-tok L4); a shorter key that is a suffix of this one cannot
        // be a field start here either, its delimiter position holds a byte of this key
        const uint32_t sel = sk.dsel[k];
        const uint32_t dw = sel == 0 ? w_c : (sel == 1 ? w_b : w_a);
        const uint32_t d = (dw >> sk.dshift[k]) & 0xFFu;
        return (d == 0x20u || d == 0x27u || d == 0x0Au) ? (int)k : -1;
    }
    // keys longer than the 8 compared bytes: finish byte by byte, then fall back to the others
    if (dm_key_check(buf, q, k, sk, 8)) return (int)k;
    for (uint32_t s = slot + 1; s < sk.n; ++s) {
        const uint32_t k2 = sk.order[s];
        if (dm_key_check(buf, q, k2, sk, 0)) return (int)k2;
    }
    return -1;
}

// dm_fp64 of the value that starts at vpos: ends at the first space outside double quotes
// (parity counted from the value start, R-tok L5), at '\n', or at the end of the message.
// Works on 16-byte blocks; quote parity is carried with shift/xor prefix tricks, no branches
// on the data inside a block.
__device__ __forceinline__ uint64_t dm_hash_value(const uint8_t* __restrict__ buf, uint64_t nbytes, uint64_t vpos) {
    DmHashState st;
    dm_hash_init(st);
    uint32_t n = 0, in_q = 0;                 // in_q: 0 or 0x80808080
    uint64_t a = vpos & ~3ull;
    const uint32_t sh = (uint32_t)(vpos & 3) * 8;
    uint32_t lo = (a < nbytes) ? dm_ld32(buf, a) : 0u;
    uint64_t pos = vpos;
    for (;;) {
        const uint32_t x1 = (a + 4 < nbytes) ? dm_ld32(buf, a + 4) : 0u;
        const uint32_t x2 = (a + 8 < nbytes) ? dm_ld32(buf, a + 8) : 0u;
        const uint32_t x3 = (a + 12 < nbytes) ? dm_ld32(buf, a + 12) : 0u;
        const uint32_t x4 = (a + 16 < nbytes) ? dm_ld32(buf, a + 16) : 0u;
        uint32_t w[4];
        w[0] = __funnelshift_r(lo, x1, sh); w[1] = __funnelshift_r(x1, x2, sh);
        w[2] = __funnelshift_r(x2, x3, sh); w[3] = __funnelshift_r(x3, x4, sh);
        lo = x4;
        a += 16;
        uint32_t nvtot = 16;
        uint32_t q_state = in_q;
        uint32_t term[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t nl = dm_eqflags(w[i], 0x0A0A0A0Au);
            const uint32_t sp = dm_eqflags(w[i], 0x20202020u);
            const uint32_t dq = dm_eqflags(w[i], 0x22222222u);
            uint32_t incl = dq ^ (dq << 8);
            incl ^= incl << 16;                                     // quote parity up to and including each byte
            const uint32_t before = (incl << 8) ^ q_state;          // in-quote state in front of each byte
            term[i] = nl | (sp & ~before);
            q_state ^= (uint32_t)((int32_t)incl >> 31) & 0x80808080u;   // parity of the whole word
        }
#pragma unroll
        for (int i = 3; i >= 0; --i)
            if (term[i]) nvtot = 4u * i + ((uint32_t)(__ffs(term[i]) - 1) >> 3);
        const uint64_t rem = nbytes > pos ? nbytes - pos : 0;
        if (rem < nvtot) nvtot = (uint32_t)rem;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (nvtot > 4u * i) {
                const uint32_t nb = nvtot - 4u * i;                 // valid bytes of this word (>= 1)
                dm_hash_word(st, nb >= 4u ? w[i] : (w[i] & ((1u << (8 * nb)) - 1u)));
            }
        }
        n += nvtot;
        if (nvtot < 16) break;
        in_q = q_state;
        pos += 16;
    }
    return dm_hash_final(st, n);
}

// 16-bit mask of the bytes of a 16-byte chunk equal to the byte replicated in pat
__device__ __forceinline__ uint32_t dm_chunk_mask(const uint4& v, uint32_t pat) {
    return dm_flags_to_nib(dm_eqflags(v.x, pat)) | (dm_flags_to_nib(dm_eqflags(v.y, pat)) << 4) |
           (dm_flags_to_nib(dm_eqflags(v.z, pat)) << 8) | (dm_flags_to_nib(dm_eqflags(v.w, pat)) << 12);
}

// Exact re-check of one candidate, by the whole warp (all lanes pass the same q and k and get
// the same answer): is the '=' at q the FIRST true field (R-tok L2-L6) with key k of its
// record?  Also returns the record's first byte.  Each lane takes one 16-byte chunk of a
// 512-byte window; quote parity is a ballot prefix, the first true field a warp minimum.
__device__ __forceinline__ bool dm_verify_field_warp(const uint8_t* __restrict__ buf, uint32_t q, uint32_t k,
                                                     const DmKeys& sk, uint32_t lane, uint32_t lt,
                                                     uint32_t* line_start) {
    // ---- the record's first byte: last '\n' before q ----
    uint32_t s = 0;
    for (long long cq = (long long)(q >> 4);; cq -= 32) {
        const long long c = cq - 31 + (long long)lane;
        uint32_t m = 0;
        if (c >= 0) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(buf + c * 16));
            m = dm_chunk_mask(v, 0x0A0A0A0Au);
            if (c == (long long)(q >> 4)) m &= (1u << (q & 15u)) - 1u;       // only bytes in front of q
        }
        const uint32_t b = __ballot_sync(0xffffffffu, m != 0);
        if (b) {
            const uint32_t L = 31u - (uint32_t)__clz((int)b);
            const uint32_t pos = (uint32_t)(c * 16) + (31u - (uint32_t)__clz((int)m)) + 1u;
            s = __shfl_sync(0xffffffffu, pos, (int)L);
            break;
        }
        if (cq - 31 <= 0) break;                                             // reached the start of the message
    }
    *line_start = s;
    // ---- first true field with key k in [s, q] ----
    const uint32_t len = sk.len[k];
    uint32_t carry = 0;
    for (uint32_t wbase = s & ~15u; wbase <= q; wbase += 512u) {
        const uint32_t off = wbase + lane * 16u;
        uint32_t dq16 = 0, eq16 = 0;
        if (off <= q) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(buf + off));
            uint32_t keep = 0xFFFFu;
            if (off < s) keep &= ~((1u << (s - off)) - 1u);                  // bytes of the previous record
            if (q - off < 15u) keep &= (2u << (q - off)) - 1u;                // bytes behind q
            dq16 = dm_chunk_mask(v, 0x22222222u) & keep;
            eq16 = dm_chunk_mask(v, 0x3D3D3D3Du) & keep;
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, __popc(dq16) & 1);
        const uint32_t inq_lane = carry ^ ((uint32_t)__popc(bal & lt) & 1u);
        uint32_t best = 0xFFFFFFFFu;
        uint32_t m = eq16;
        while (m && best == 0xFFFFFFFFu) {
            const uint32_t j = (uint32_t)__ffs(m) - 1;
            m &= m - 1;
            const uint32_t p = off + j;
            const uint32_t inq = inq_lane ^ ((uint32_t)__popc(dq16 & ((1u << j) - 1u)) & 1u);
            if (!inq && p >= s + len) {
                const uint32_t p0 = p - len;
                bool ok = true;
                if (p0 > s) {
                    const uint32_t d = dm_ld8(buf, p0 - 1);
                    ok = d == 0x20u || d == 0x27u;
                }
                for (uint32_t i = 0; ok && i < len; ++i) ok = dm_ld8(buf, p0 + i) == sk.bytes[k][i];
                if (ok) best = p;
            }
        }
        const uint32_t mn = __reduce_min_sync(0xffffffffu, best);
        if (mn != 0xFFFFFFFFu) return mn == q;
        carry ^= (uint32_t)__popc(bal) & 1u;
    }
    return false;
}

// ---------------------------------------------------------------------------------------
template <bool TRAIN>
__global__ void __launch_bounds__(DMT_THREADS) dm_k_tile(DmFusedArgs a) {
    __shared__ DmKeys sk;   // 16-byte aligned through its alignas(16) member
    __shared__ uint32_t s_cnt[DMT_WARPS];
    __shared__ unsigned long long s_base[DMT_WARPS];
    __shared__ long long s_tile;
    __shared__ DmQ1Entry s_q1[DMT_WARPS][DMT_Q1CAP];
    __shared__ DmQ2Entry s_q2[DMT_WARPS][DMT_Q2CAP];
    __shared__ DmPEntry s_pend[DMT_WARPS][DMT_PCAP];

    const uint8_t* __restrict__ buf = a.buf;
    const uint64_t nbytes = a.nbytes;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lt = dm_lanemask_lt();

    // The first tile of a CTA is its block index (the host launches at most as many CTAs as
    // are co-resident, and never more than there are tiles); further tiles come from the
    // atomic counter.
    long long tile = (long long)blockIdx.x;
    bool keys_ready = false;                                   // workers: sk is valid (barrier 3 passed)
    if (warp == DMT_WARPS) {
        // the scanner warp brings the key tables into shared memory while the workers are
        // already counting record starts (pass 1 does not need them)
        const uint32_t* src = reinterpret_cast<const uint32_t*>(a.keys);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&sk);
        for (uint32_t i = lane; i < sizeof(DmKeys) / 4; i += 32) dst[i] = __ldg(src + i);
        __threadfence_block();
        dm_bar_arrive(3, DMT_THREADS);
    }
    for (;;) {
        if (tile >= (long long)a.n_tiles) break;

        if (warp == DMT_WARPS) {
            // ------------------------------ scanner warp ------------------------------
            // fetch this CTA's next tile now; the answer is only needed after this tile
            if (lane == 0) s_tile = (long long)gridDim.x + (long long)(atomicAdd(a.tile_ctr, 1ull) - a.ctr_base);
            dm_bar_sync(1, DMT_THREADS);                       // the record counts are in s_cnt
            const uint32_t c = lane < DMT_WARPS ? s_cnt[lane] : 0u;
            uint32_t incl = c;
#pragma unroll
            for (int d = 1; d < DMT_WARPS; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
                if ((int)lane >= d) incl += y;
            }
            const uint32_t agg = __shfl_sync(0xffffffffu, incl, DMT_WARPS - 1);
            const unsigned long long tag = (unsigned long long)a.epoch << 34;
            unsigned long long excl = 0;
            if (tile > 0) {
                if (lane == 0) atomicExch(a.tile_state + tile, tag | (DMT_ST_AGG << 32) | agg);
                // Decoupled look-back.  All tiles of a message are usually in flight at once, so
                // hardly any predecessor has published a PREFIX yet: waiting for one would make
                // tile t walk t/32 dependent round trips.  Instead 128 predecessors are loaded
                // per step (4 independent loads per lane) and only their AGGREGATES are needed;
                // a PREFIX, where one is found, just ends the walk early.
                long long hi = tile - 1;
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
                        if (idx >= 0 && (w >> 34) != a.epoch) w = 0;       // stale word of an earlier launch
                        const uint32_t status = idx >= 0 ? (uint32_t)((w >> 32) & 3u) : (uint32_t)DMT_ST_PREFIX;
                        const uint32_t not_ready = __ballot_sync(0xffffffffu, status == 0);
                        const uint32_t is_pref = __ballot_sync(0xffffffffu, status == DMT_ST_PREFIX);
                        const uint32_t first_pref = is_pref ? (uint32_t)(__ffs(is_pref) - 1) : 32u;
                        const uint32_t upto = first_pref < 32u ? first_pref : 31u;
                        const uint32_t win = upto == 31u ? 0xffffffffu : ((2u << upto) - 1u);
                        if (not_ready & win) continue;                      // retry from this window
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
            if (lane < DMT_WARPS) s_base[lane] = excl + (incl - c);
            if (a.finalize && tile == (long long)a.n_tiles - 1 && lane == 0) {
                unsigned long long n_lines = excl + agg;
                if (n_lines > a.max_lines || n_lines > a.out_cap) atomicOr(&a.hdr->error, DM_DEVERR_TOO_MANY_LINES);
                a.hdr->n_lines = n_lines;
                const unsigned long long tr = a.n_train_lines < n_lines ? a.n_train_lines : n_lines;
                a.stats[0] += n_lines;
                a.stats[1] += tr;
                a.stats[2] += n_lines - tr;
                a.stats[5] += nbytes;
            }
            __threadfence_block();
            dm_bar_arrive(2, DMT_THREADS);                     // bases are in s_base
        } else {
            // ------------------------------ worker warp -------------------------------
            const uint64_t seg_start = (uint64_t)tile * DMT_TILE + (uint64_t)warp * DMT_SEG;
            uint64_t seg_end = seg_start + DMT_SEG;
            if (seg_end > nbytes) seg_end = nbytes;
            uint32_t n_owned = 0, is_start0 = 0, in_seg_rows = 0;

            // ---- pass 1: how many records start inside [seg_start, seg_end) ----
            if (seg_start < nbytes) {
                in_seg_rows = (uint32_t)((seg_end - seg_start + DMT_ROW - 1) / DMT_ROW);
                is_start0 = (seg_start == 0 || dm_ld8(buf, seg_start - 1) == 0x0Au) ? 1u : 0u;
                uint32_t cnt = 0;
                for (uint32_t r = 0; r < in_seg_rows; ++r) {
                    const uint64_t off = seg_start + (uint64_t)r * DMT_ROW + (uint64_t)lane * 16;
                    if (off < seg_end) {
                        const uint4 v = __ldg(reinterpret_cast<const uint4*>(buf + off));
                        uint32_t f0 = dm_eqflags(v.x, 0x0A0A0A0Au), f1 = dm_eqflags(v.y, 0x0A0A0A0Au);
                        uint32_t f2 = dm_eqflags(v.z, 0x0A0A0A0Au), f3 = dm_eqflags(v.w, 0x0A0A0A0Au);
                        if (f0 | f1 | f2 | f3) {
                            if (off + 16 > seg_end) {            // only the last chunk of the message
                                const uint32_t vb = (uint32_t)(seg_end - off);
                                uint32_t m = dm_flags_to_nib(f0) | (dm_flags_to_nib(f1) << 4) |
                                             (dm_flags_to_nib(f2) << 8) | (dm_flags_to_nib(f3) << 12);
                                cnt += __popc(m & ((1u << vb) - 1u));
                            } else {
                                cnt += __popc(f0) + __popc(f1) + __popc(f2) + __popc(f3);
                            }
                        }
                    }
                }
                const uint32_t total = __reduce_add_sync(0xffffffffu, cnt);
                const uint32_t last_nl = dm_ld8(buf, seg_end - 1) == 0x0Au ? 1u : 0u;
                n_owned = is_start0 + total - last_nl;
            }
            if (lane == 0) s_cnt[warp] = n_owned;
            __threadfence_block();
            dm_bar_arrive(1, DMT_THREADS);

            if (!keys_ready) { dm_bar_sync(3, DMT_THREADS); keys_ready = true; }
            bool have_base = false, zero_done = false;
            unsigned long long base = 0;
            if (a.range_check) { dm_bar_sync(2, DMT_THREADS); have_base = true; base = s_base[warp]; }

            uint32_t q1h = 0, q1n = 0, q2h = 0, q2n = 0, ph = 0, pn = 0;   // circular queues: head, count
            DmQ1Entry* q1 = s_q1[warp];
            DmQ2Entry* q2 = s_q2[warp];
            DmPEntry* pend = s_pend[warp];

            // apply up to 32 pending alerts (needs the global record index of this warp's records)
            auto flush = [&]() {
                if (!have_base) { dm_bar_sync(2, DMT_THREADS); have_base = true; base = s_base[warp]; }
                if (!zero_done) {
                    if (a.zero_fill) {
                        for (uint32_t i = lane; i < n_owned; i += 32) {
                            const unsigned long long g = base + i;
                            if (g < a.out_cap) { a.flags[g] = 0; a.scores[g] = 0.0f; }
                        }
                    }
                    zero_done = true;
                    __syncwarp();
                }
                const uint32_t take = pn < 32u ? pn : 32u;
                if (take) {
                    bool first = false;
                    if (lane < take) {
                        const DmPEntry e = pend[(ph + lane) & (DMT_PCAP - 1)];
                        const unsigned long long g = base + e.ln;
                        if (g < a.out_cap) {
                            const float old = atomicAdd(a.scores + g, 1.0f);
                            a.flags[g] = 1;
                            first = old == 0.0f;
                        }
                        atomicAdd(a.stats + 8 + e.k, 1ull);
                        const unsigned int idx = atomicAdd(&a.hdr->anomaly_list_count, 1u);
                        if (idx < a.anomaly_cap) {
                            dm_anomaly_t r;
                            r.line = (uint32_t)g; r.mask = 1u << e.k; r.offset = e.lstart;
                            a.anomalies[idx] = r;
                        }
                    }
                    const uint32_t nf = __popc(__ballot_sync(0xffffffffu, first));
                    if (lane == 0) {
                        if (nf) { atomicAdd(&a.hdr->n_anomalies, (unsigned long long)nf); atomicAdd(a.stats + 3, (unsigned long long)nf); }
                        atomicAdd(a.stats + 4, (unsigned long long)take);
                    }
                    ph += take;
                    pn -= take;
                    __syncwarp();
                }
            };

            // stage 2: one identified field per lane -- fingerprint, probe, exact re-check when unknown
            auto drain2 = [&](uint32_t n) {
                bool unk = false, cand = false;
                uint64_t ckey = 0;
                uint32_t cq = 0, ck = 0, cln = 0;
                DmPEntry pe;
                pe.ln = 0; pe.k = 0; pe.lstart = 0;
                if (lane < n) {
                    const DmQ2Entry e = q2[(q2h + lane) & (DMT_Q2CAP - 1)];
                    bool in_range = true;
                    if (a.range_check) {
                        const unsigned long long g = base + (unsigned long long)e.ln;
                        in_range = g >= a.line_lo && g < a.line_hi;
                    }
                    if (in_range) {
                        const uint64_t fp = dm_hash_value(buf, nbytes, e.vpos);
                        const uint64_t key = dm_make_key(fp, sk.salt[e.k]);
                        const bool known = TRAIN ? dm_table_contains_volatile(a.table, key) : dm_table_contains(a.table, key);
                        cand = !known;
                        ckey = key;
                        cq = e.vpos - 1;
                        ck = e.k;
                        cln = (uint32_t)e.ln;
                    }
                }
                // values that are not in the table: exact re-check, one candidate at a time, by the whole warp
                uint32_t cb = __ballot_sync(0xffffffffu, cand);
                while (cb) {
                    const int L = __ffs(cb) - 1;
                    cb &= cb - 1;
                    const uint32_t vq = __shfl_sync(0xffffffffu, cq, L);
                    const uint32_t vk = __shfl_sync(0xffffffffu, ck, L);
                    if (TRAIN) {
                        // another lane (or warp) may have inserted this very value meanwhile
                        const int still_new = ((int)lane == L) ? (dm_table_contains_volatile(a.table, ckey) ? 0 : 1) : 0;
                        if (!__shfl_sync(0xffffffffu, still_new, L)) continue;
                    }
                    uint32_t ls = 0;
                    const bool ok = dm_verify_field_warp(buf, vq, vk, sk, lane, lt, &ls);
                    if ((int)lane == L && ok) {
                        if (TRAIN) {
                            dm_table_insert(a.table, ckey, &a.hdr->error);
                        } else {
                            unk = true;
                            pe.ln = cln; pe.k = ck; pe.lstart = ls;
                        }
                    }
                }
                q2h += n;
                q2n -= n;
                if (!TRAIN) {
                    const uint32_t ub = __ballot_sync(0xffffffffu, unk);
                    if (ub) {
                        if (unk) pend[(ph + pn + __popc(ub & lt)) & (DMT_PCAP - 1)] = pe;
                        pn += __popc(ub);
                        __syncwarp();
                        if (pn >= 32u) flush();
                    }
                }
                __syncwarp();
            };

            // stage 1: one '=' per lane -- which monitored key (if any) stands in front of it
            auto drain1 = [&](uint32_t n) {
                bool matched = false;
                DmQ2Entry qe;
                qe.vpos = 0; qe.ln = 0; qe.k = 0;
                if (lane < n) {
                    const DmQ1Entry e = q1[(q1h + lane) & (DMT_Q1CAP - 1)];
                    if (e.ln >= 0 && e.ln < (int32_t)n_owned) {
                        const int k = dm_key_identify(buf, (uint64_t)e.q, sk);
                        if (k >= 0) { matched = true; qe.vpos = e.q + 1; qe.ln = e.ln; qe.k = (uint32_t)k; }
                    }
                }
                q1h += n;
                q1n -= n;
                const uint32_t mb = __ballot_sync(0xffffffffu, matched);
                if (mb) {
                    if (matched) q2[(q2h + q2n + __popc(mb & lt)) & (DMT_Q2CAP - 1)] = qe;
                    q2n += __popc(mb);
                    __syncwarp();
                    if (q2n >= 32u) drain2(32u);
                }
                __syncwarp();
            };

            // ---- pass 2: classify the rows, queue every '=' ----
            if (n_owned > 0) {
                const int32_t ln_off = (int32_t)is_start0 - 1;
                const uint32_t need_nl = n_owned - is_start0 + 1;   // newlines until the last owned record is closed
                uint32_t nl_seen = 0;
                for (uint32_t row = 0;; ++row) {
                    const uint64_t row_off = seg_start + (uint64_t)row * DMT_ROW;
                    if (row_off >= nbytes) break;
                    const uint64_t off = row_off + (uint64_t)lane * 16;
                    uint32_t nlF0 = 0, nlF1 = 0, nlF2 = 0, nlF3 = 0, eq16 = 0;
                    if (off < nbytes) {
                        const uint4 v = __ldg(reinterpret_cast<const uint4*>(buf + off));
                        nlF0 = dm_eqflags(v.x, 0x0A0A0A0Au); nlF1 = dm_eqflags(v.y, 0x0A0A0A0Au);
                        nlF2 = dm_eqflags(v.z, 0x0A0A0A0Au); nlF3 = dm_eqflags(v.w, 0x0A0A0A0Au);
                        const uint32_t e0 = dm_eqflags(v.x, 0x3D3D3D3Du), e1 = dm_eqflags(v.y, 0x3D3D3D3Du);
                        const uint32_t e2 = dm_eqflags(v.z, 0x3D3D3D3Du), e3 = dm_eqflags(v.w, 0x3D3D3D3Du);
                        if (e0 | e1 | e2 | e3)
                            eq16 = dm_flags_to_nib(e0) | (dm_flags_to_nib(e1) << 4) | (dm_flags_to_nib(e2) << 8) |
                                   (dm_flags_to_nib(e3) << 12);
                        if (off + 16 > nbytes) {                  // last chunk of the message: drop the slack bytes
                            const uint32_t vb = (uint32_t)(nbytes - off);
                            eq16 &= (1u << vb) - 1u;
                            const uint32_t keep0 = vb >= 4 ? 0xFFFFFFFFu : ((1u << (8 * vb)) - 1u);
                            const uint32_t keep1 = vb >= 8 ? 0xFFFFFFFFu : (vb > 4 ? ((1u << (8 * (vb - 4))) - 1u) : 0u);
                            const uint32_t keep2 = vb >= 12 ? 0xFFFFFFFFu : (vb > 8 ? ((1u << (8 * (vb - 8))) - 1u) : 0u);
                            const uint32_t keep3 = vb > 12 ? ((1u << (8 * (vb - 12))) - 1u) : 0u;
                            nlF0 &= keep0; nlF1 &= keep1; nlF2 &= keep2; nlF3 &= keep3;
                        }
                    }
                    const bool nl_any = (nlF0 | nlF1 | nlF2 | nlF3) != 0;
                    const uint32_t b_nl = __ballot_sync(0xffffffffu, nl_any);
                    uint32_t pre = 0, row_nl = 0;
                    if (b_nl) {
                        const uint32_t my = nl_any ? (uint32_t)(__popc(nlF0) + __popc(nlF1) + __popc(nlF2) + __popc(nlF3)) : 0u;
                        const uint32_t b_multi = __ballot_sync(0xffffffffu, my > 1);
                        if (b_multi == 0) {
                            pre = __popc(b_nl & lt);
                            row_nl = __popc(b_nl);
                        } else {
                            uint32_t incl = my;
#pragma unroll
                            for (int d = 1; d < 32; d <<= 1) {
                                const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
                                if ((int)lane >= d) incl += y;
                            }
                            pre = incl - my;
                            row_nl = __shfl_sync(0xffffffffu, incl, 31);
                        }
                    }
                    const int32_t ln_chunk = ln_off + (int32_t)(nl_seen + pre);

                    // every '=' of the row goes to queue 1, in position order
                    const uint32_t my_eq = (uint32_t)__popc(eq16);
                    uint32_t eincl = my_eq;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t y = __shfl_up_sync(0xffffffffu, eincl, d);
                        if ((int)lane >= d) eincl += y;
                    }
                    const uint32_t row_eq = __shfl_sync(0xffffffffu, eincl, 31);
                    if (row_eq) {
                        uint32_t m = eq16;
                        if (q1n + row_eq <= DMT_Q1CAP) {
                            uint32_t slot = q1h + q1n + (eincl - my_eq);
                            while (m) {
                                const uint32_t j = (uint32_t)__ffs(m) - 1;
                                m &= m - 1;
                                int32_t ln = ln_chunk;
                                if (nl_any) {
                                    const uint32_t wj = j >> 2, bm = (1u << ((j & 3) * 8)) - 1u;
                                    ln += __popc(nlF0 & (wj > 0 ? 0xFFFFFFFFu : bm));
                                    ln += __popc(nlF1 & (wj > 1 ? 0xFFFFFFFFu : (wj == 1 ? bm : 0u)));
                                    ln += __popc(nlF2 & (wj > 2 ? 0xFFFFFFFFu : (wj == 2 ? bm : 0u)));
                                    ln += __popc(nlF3 & (wj == 3 ? bm : 0u));
                                }
                                DmQ1Entry e;
                                e.q = (uint32_t)(off + j); e.ln = ln;
                                q1[slot & (DMT_Q1CAP - 1)] = e;
                                ++slot;
                            }
                            q1n += row_eq;
                            __syncwarp();
                            while (q1n >= 32u) drain1(32u);
                        } else {
                            // a row dense in '=' (more than the queue holds): one per lane per round
                            while (__ballot_sync(0xffffffffu, m != 0)) {
                                bool has = false;
                                DmQ1Entry e;
                                e.q = 0; e.ln = 0;
                                if (m) {
                                    const uint32_t j = (uint32_t)__ffs(m) - 1;
                                    m &= m - 1;
                                    int32_t ln = ln_chunk;
                                    if (nl_any) {
                                        const uint32_t wj = j >> 2, bm = (1u << ((j & 3) * 8)) - 1u;
                                        ln += __popc(nlF0 & (wj > 0 ? 0xFFFFFFFFu : bm));
                                        ln += __popc(nlF1 & (wj > 1 ? 0xFFFFFFFFu : (wj == 1 ? bm : 0u)));
                                        ln += __popc(nlF2 & (wj > 2 ? 0xFFFFFFFFu : (wj == 2 ? bm : 0u)));
                                        ln += __popc(nlF3 & (wj == 3 ? bm : 0u));
                                    }
                                    has = true; e.q = (uint32_t)(off + j); e.ln = ln;
                                }
                                const uint32_t hb = __ballot_sync(0xffffffffu, has);
                                if (has) q1[(q1h + q1n + __popc(hb & lt)) & (DMT_Q1CAP - 1)] = e;
                                q1n += __popc(hb);
                                __syncwarp();
                                while (q1n >= 32u) drain1(32u);
                            }
                        }
                    }
                    nl_seen += row_nl;
                    if (row + 1 >= in_seg_rows && nl_seen >= need_nl) break;
                }
                while (q1n) drain1(q1n < 32u ? q1n : 32u);
                while (q2n) drain2(q2n < 32u ? q2n : 32u);
            }
            while (pn > 32u) flush();
            flush();
        }
        __syncthreads();
        tile = s_tile;
        __syncthreads();
    }
}

#ifndef DM_EMU
// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
struct DmTileScratch {
    unsigned long long* d_tile_state = nullptr;
    unsigned long long* d_tile_ctr = nullptr;
    uint64_t max_tiles = 0;
    unsigned long long ctr_base = 0;    // host mirror of *d_tile_ctr between launches
    uint32_t epoch = 0;
    int grid = 0;
};

static inline int dm_tile_scratch_create(DmTileScratch* s, uint64_t max_batch_bytes, int sm_count) {
    s->max_tiles = (max_batch_bytes + DMT_TILE - 1) / DMT_TILE + 1;
    if (cudaMalloc(&s->d_tile_state, s->max_tiles * sizeof(unsigned long long)) != cudaSuccess) return DM_ERR_CUDA;
    if (cudaMemset(s->d_tile_state, 0, s->max_tiles * sizeof(unsigned long long)) != cudaSuccess) return DM_ERR_CUDA;
    if (cudaMalloc(&s->d_tile_ctr, sizeof(unsigned long long)) != cudaSuccess) return DM_ERR_CUDA;
    if (cudaMemset(s->d_tile_ctr, 0, sizeof(unsigned long long)) != cudaSuccess) return DM_ERR_CUDA;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dm_k_tile<false>, DMT_THREADS, 0) != cudaSuccess) return DM_ERR_CUDA;
    if (per_sm < 1) per_sm = 1;
    s->grid = sm_count * per_sm;
    return DM_OK;
}

static inline void dm_tile_scratch_destroy(DmTileScratch* s) {
    cudaFree(s->d_tile_state);
    cudaFree(s->d_tile_ctr);
    s->d_tile_state = nullptr;
    s->d_tile_ctr = nullptr;
}

// Enqueue the fused kernel(s) for one message.  Returns the number of kernels launched
// (>= 0) or a negative DM_ERR_* code.
static inline int dm_tile_launch(DmTileScratch* s, const uint8_t* d_buf, uint64_t nbytes, uint64_t n_train_lines,
                                 const DmKeys* d_keys, DmTable table, uint8_t* d_flags, float* d_scores,
                                 uint64_t out_cap, dm_anomaly_t* d_anoms, uint32_t anomaly_cap, DmBatchHeader* d_hdr,
                                 unsigned long long* d_stats, uint64_t max_lines, cudaStream_t st) {
    const uint32_t n_tiles = (uint32_t)((nbytes + DMT_TILE - 1) / DMT_TILE);
    if (n_tiles == 0) return 0;
    if (n_tiles > s->max_tiles) return DM_ERR_CAPACITY;
    DmFusedArgs a;
    a.buf = d_buf; a.nbytes = nbytes; a.n_tiles = n_tiles; a.keys = d_keys; a.table = table;
    a.flags = d_flags; a.scores = d_scores; a.out_cap = out_cap; a.anomalies = d_anoms; a.anomaly_cap = anomaly_cap;
    a.hdr = d_hdr; a.stats = d_stats; a.tile_state = s->d_tile_state; a.tile_ctr = s->d_tile_ctr;
    a.n_train_lines = n_train_lines; a.max_lines = max_lines;
    const int grid = (int)(n_tiles < (uint32_t)s->grid ? n_tiles : (uint32_t)s->grid);
    int launched = 0;
    if (n_train_lines > 0) {
        s->epoch = (s->epoch % 0x3FFFFFFEu) + 1u;
        a.epoch = s->epoch; a.ctr_base = s->ctr_base;
        a.line_lo = 0; a.line_hi = n_train_lines; a.range_check = 1; a.zero_fill = 1; a.finalize = 0;
        dm_k_tile<true><<<grid, DMT_THREADS, 0, st>>>(a);
        s->ctr_base += (unsigned long long)n_tiles;
        ++launched;
    }
    s->epoch = (s->epoch % 0x3FFFFFFEu) + 1u;
    a.epoch = s->epoch; a.ctr_base = s->ctr_base;
    a.line_lo = n_train_lines; a.line_hi = ~0ull; a.range_check = n_train_lines > 0 ? 1 : 0;
    a.zero_fill = n_train_lines > 0 ? 0 : 1; a.finalize = 1;
    dm_k_tile<false><<<grid, DMT_THREADS, 0, st>>>(a);
    s->ctr_base += (unsigned long long)n_tiles;
    ++launched;
    if (cudaGetLastError() != cudaSuccess) return DM_ERR_CUDA;
    return launched;
}
#endif  // DM_EMU
