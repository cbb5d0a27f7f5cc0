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
//     rows stay in L1), pass 2 classifies '\n' and '=' bytes with SIMD-in-register compares,
//     filters every '=' against the monitored keys by the bytes in front of it, and queues
//     the matches; a full queue is drained one value per lane: dm_fp64 of the value, probe
//     of the known-set table;
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

#define DMT_WARPS 8
#define DMT_THREADS (DMT_WARPS * 32 + 32)
#define DMT_ROW 512u
#define DMT_SEG_ROWS 8u
#define DMT_SEG (DMT_ROW * DMT_SEG_ROWS)
#define DMT_TILE (DMT_SEG * DMT_WARPS)
#define DMT_QCAP 64
#define DMT_PCAP 64

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
    unsigned long long* tile_ctr;     // monotonically increasing across launches
    unsigned long long ctr_base;      // value of *tile_ctr when this launch's tile 0 is taken
    uint64_t line_lo, line_hi;        // this launch handles records with index in [lo, hi)
    uint64_t n_train_lines;
    uint64_t max_lines;
    int range_check;                  // honour [line_lo, line_hi) (message holds training AND detection records)
    int zero_fill;                    // this launch zero-fills flags / scores
    int finalize;                     // this launch writes the batch header and the statistics
};

struct DmQEntry { uint32_t vpos; int32_t ln; uint32_t k; };
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

// Does monitored key k end right before the '=' at q (bytes + field-start delimiter)?
// Quote parity is deliberately not checked here (see file header).
__device__ __forceinline__ bool dm_key_check(const uint8_t* __restrict__ buf, uint64_t q, uint32_t k, const DmKeys& sk,
                                             uint32_t skip_tail) {
    const uint32_t len = sk.len[k];
    if (q < len) return false;
    const uint64_t st = q - len;
    if (st > 0) {
        const uint32_t c = dm_ld8(buf, st - 1);
        if (c != 0x20u && c != 0x27u && c != 0x0Au) return false;
    }
    const uint32_t n = len - skip_tail;     // the last skip_tail bytes were compared already
    for (uint32_t i = 0; i < n; ++i)
        if (dm_ld8(buf, st + i) != sk.bytes[k][i]) return false;
    return true;
}

__device__ __forceinline__ int dm_key_filter(const uint8_t* __restrict__ buf, uint64_t q, const DmKeys& sk) {
    if (q < 4) {
        for (uint32_t k = 0; k < sk.n; ++k)
            if (dm_key_check(buf, q, k, sk, 0)) return (int)k;
        return -1;
    }
    const uint64_t a = (q - 4) & ~3ull;
    const uint32_t w0 = dm_ld32(buf, a), w1 = dm_ld32(buf, a + 4);
    const uint32_t t4 = __funnelshift_r(w0, w1, (uint32_t)((q - 4) & 3) * 8);   // bytes q-4 .. q-1
    const uint32_t h = dm_tail_hash(t4 >> 16);
    if (!((sk.bitmap[h >> 5] >> (h & 31)) & 1u)) return -1;
    for (uint32_t k = 0; k < sk.n; ++k) {
        if (((t4 ^ sk.tailbits[k]) & sk.tailmask[k]) == 0) {
            const uint32_t len = sk.len[k];
            if (dm_key_check(buf, q, k, sk, len < 4 ? len : 4)) return (int)k;
        }
    }
    return -1;
}

// dm_fp64 of the value that starts at vpos: ends at the first space outside double quotes
// (parity counted from the value start, R-tok L5), at '\n', or at the end of the message.
__device__ __forceinline__ uint64_t dm_hash_value(const uint8_t* __restrict__ buf, uint64_t nbytes, uint64_t vpos) {
    DmHashState st;
    dm_hash_init(st);
    uint32_t n = 0, in_q = 0;
    uint64_t a = vpos & ~3ull;
    const uint32_t sh = (uint32_t)(vpos & 3) * 8;
    uint32_t lo = (a < nbytes) ? dm_ld32(buf, a) : 0u;
    uint64_t pos = vpos;
    for (;;) {
        a += 4;
        const uint32_t hi = (a < nbytes) ? dm_ld32(buf, a) : 0u;
        uint32_t w = __funnelshift_r(lo, hi, sh);
        lo = hi;
        // bytes at or beyond the end of the message terminate the value
        const uint64_t rem = nbytes > pos ? nbytes - pos : 0;
        uint32_t term;
        const uint32_t dq = dm_eqflags(w, 0x22222222u);
        if (dq == 0) {
            const uint32_t nl = dm_eqflags(w, 0x0A0A0A0Au);
            term = in_q ? nl : (nl | dm_eqflags(w, 0x20202020u));
        } else {
            term = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const uint32_t c = (w >> (8 * b)) & 0xFFu;
                if (term == 0) {
                    if (c == 0x0Au || (c == 0x20u && !in_q)) term = 0x80u << (8 * b);
                    else if (c == 0x22u) in_q ^= 1u;
                }
            }
        }
        uint32_t nv = term ? (uint32_t)(__ffs(term) - 1) >> 3 : 4u;
        if (rem < nv) nv = (uint32_t)rem;
        if (nv < 4) {
            if (nv) { dm_hash_word(st, w & ((1u << (8 * nv)) - 1u)); n += nv; }
            break;
        }
        dm_hash_word(st, w);
        n += 4;
        pos += 4;
    }
    return dm_hash_final(st, n);
}

// Slow, exact re-check of one candidate: is the '=' at q the FIRST true field (R-tok L2-L6)
// with key k of its record?  Also returns the record's first byte.
__device__ __forceinline__ bool dm_verify_field(const uint8_t* __restrict__ buf, uint64_t q, uint32_t k, const DmKeys& sk,
                                                uint64_t* line_start) {
    uint64_t s = q;
    while (s > 0 && dm_ld8(buf, s - 1) != 0x0Au) --s;
    *line_start = s;
    const uint32_t len = sk.len[k];
    uint32_t inq = 0, prev = 0x20u;
    for (uint64_t p = s; p + len <= q; ++p) {
        const uint32_t c = dm_ld8(buf, p);
        if (!inq && (p == s || prev == 0x20u || prev == 0x27u)) {
            if (dm_ld8(buf, p + len) == 0x3Du) {
                bool eq = true;
                for (uint32_t i = 0; i < len; ++i)
                    if (dm_ld8(buf, p + i) != sk.bytes[k][i]) { eq = false; break; }
                if (eq) return p + len == q;
            }
        }
        if (c == 0x22u) inq ^= 1u;
        prev = c;
    }
    return false;
}

// ---------------------------------------------------------------------------------------
template <bool TRAIN>
__global__ void __launch_bounds__(DMT_THREADS) dm_k_tile(DmFusedArgs a) {
    __shared__ DmKeys sk;
    __shared__ uint32_t s_cnt[DMT_WARPS];
    __shared__ unsigned long long s_base[DMT_WARPS];
    __shared__ long long s_tile;
    __shared__ DmQEntry s_queue[DMT_WARPS][DMT_QCAP];
    __shared__ DmPEntry s_pend[DMT_WARPS][DMT_PCAP];

    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(a.keys);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&sk);
        for (uint32_t i = threadIdx.x; i < sizeof(DmKeys) / 4; i += DMT_THREADS) dst[i] = src[i];
    }
    const uint8_t* __restrict__ buf = a.buf;
    const uint64_t nbytes = a.nbytes;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lt = dm_lanemask_lt();

    for (;;) {
        if (threadIdx.x == DMT_WARPS * 32) s_tile = (long long)(atomicAdd(a.tile_ctr, 1ull) - a.ctr_base);
        __syncthreads();
        const long long tile = s_tile;
        if (tile >= (long long)a.n_tiles) break;

        if (warp == DMT_WARPS) {
            // ------------------------------ scanner warp ------------------------------
            dm_bar_sync(1, DMT_THREADS);                       // the 8 record counts are in s_cnt
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
                // decoupled look-back, 32 predecessors per step
                long long hi = tile - 1;
                for (;;) {
                    const long long idx = hi - (long long)lane;
                    unsigned long long st = 0;
                    if (idx >= 0) {
                        st = *((volatile unsigned long long*)(a.tile_state + idx));
                        if ((st >> 34) != a.epoch) st = 0;             // stale word of an earlier launch
                    }
                    const uint32_t status = idx >= 0 ? (uint32_t)((st >> 32) & 3u) : (uint32_t)DMT_ST_PREFIX;
                    const uint32_t not_ready = __ballot_sync(0xffffffffu, status == 0);
                    const uint32_t is_pref = __ballot_sync(0xffffffffu, status == DMT_ST_PREFIX);
                    // usable window: lanes up to (and including) the first prefix, none of them not-ready
                    const uint32_t first_pref = is_pref ? (uint32_t)(__ffs(is_pref) - 1) : 32u;
                    const uint32_t upto = first_pref < 32u ? first_pref : 31u;
                    const uint32_t win = upto == 31u ? 0xffffffffu : ((2u << upto) - 1u);
                    if (not_ready & win) { __nanosleep(40); continue; }
                    const uint32_t val = (lane <= upto && idx >= 0) ? (uint32_t)st : 0u;
                    excl += __reduce_add_sync(0xffffffffu, val);
                    if (first_pref < 32u) break;
                    hi -= 32;
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

            bool have_base = false, zero_done = false;
            unsigned long long base = 0;
            if (a.range_check) { dm_bar_sync(2, DMT_THREADS); have_base = true; base = s_base[warp]; }

            uint32_t qn = 0, pn = 0;
            DmQEntry* queue = s_queue[warp];
            DmPEntry* pend = s_pend[warp];

            // apply the pending alerts (needs the global record index of this warp's records)
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
                if (pn) {
                    bool first = false;
                    if (lane < pn) {
                        const DmPEntry e = pend[lane];
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
                        atomicAdd(a.stats + 4, (unsigned long long)(pn < 32u ? pn : 32u));
                    }
                    // entries beyond the first 32 move to the front
                    DmPEntry mv;
                    const bool has_mv = pn > 32u && lane < pn - 32u;
                    if (has_mv) mv = pend[32 + lane];
                    __syncwarp();
                    if (has_mv) pend[lane] = mv;
                    pn = pn > 32u ? pn - 32u : 0u;
                    __syncwarp();
                }
            };

            // one queued value per lane: fingerprint, probe, and the exact re-check when unknown
            auto drain = [&](uint32_t n) {
                bool unk = false;
                DmPEntry pe;
                pe.ln = 0; pe.k = 0; pe.lstart = 0;
                if (lane < n) {
                    const DmQEntry e = queue[lane];
                    bool in_range = true;
                    if (a.range_check) {
                        const unsigned long long g = base + (unsigned long long)e.ln;
                        in_range = g >= a.line_lo && g < a.line_hi;
                    }
                    if (in_range) {
                        const uint64_t fp = dm_hash_value(buf, nbytes, e.vpos);
                        const uint64_t key = dm_make_key(fp, sk.salt[e.k]);
                        const bool known = TRAIN ? dm_table_contains_volatile(a.table, key) : dm_table_contains(a.table, key);
                        if (!known) {
                            uint64_t ls;
                            if (dm_verify_field(buf, (uint64_t)e.vpos - 1, e.k, sk, &ls)) {
                                if (TRAIN) {
                                    dm_table_insert(a.table, key, &a.hdr->error);
                                } else {
                                    unk = true;
                                    pe.ln = (uint32_t)e.ln; pe.k = e.k; pe.lstart = (uint32_t)ls;
                                }
                            }
                        }
                    }
                }
                // compact the queue
                DmQEntry mv;
                const bool has_mv = qn > n && lane < qn - n;
                if (has_mv) mv = queue[n + lane];
                __syncwarp();
                if (has_mv) queue[lane] = mv;
                qn -= n;
                if (!TRAIN) {
                    const uint32_t ub = __ballot_sync(0xffffffffu, unk);
                    if (ub) {
                        if (unk) pend[pn + __popc(ub & lt)] = pe;
                        pn += __popc(ub);
                        __syncwarp();
                        if (pn >= 32u) flush();
                    }
                }
                __syncwarp();
            };

            // ---- pass 2: classify, filter, queue ----
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
                    const int32_t limit = row >= in_seg_rows ? (int32_t)n_owned : 0x7fffffff;

                    uint32_t m = eq16;
                    while (__ballot_sync(0xffffffffu, m != 0)) {
                        bool matched = false;
                        DmQEntry qe;
                        qe.vpos = 0; qe.ln = 0; qe.k = 0;
                        if (m) {
                            const uint32_t j = (uint32_t)__ffs(m) - 1;
                            m &= m - 1;
                            const uint64_t q = off + j;
                            int32_t ln = ln_chunk;
                            if (nl_any) {
                                const uint32_t wj = j >> 2, bm = (1u << ((j & 3) * 8)) - 1u;
                                ln += __popc(nlF0 & (wj > 0 ? 0xFFFFFFFFu : bm));
                                ln += __popc(nlF1 & (wj > 1 ? 0xFFFFFFFFu : (wj == 1 ? bm : 0u)));
                                ln += __popc(nlF2 & (wj > 2 ? 0xFFFFFFFFu : (wj == 2 ? bm : 0u)));
                                ln += __popc(nlF3 & (wj == 3 ? bm : 0u));
                            }
                            if (ln >= 0 && ln < limit) {
                                const int k = dm_key_filter(buf, q, sk);
                                if (k >= 0) { matched = true; qe.vpos = (uint32_t)(q + 1); qe.ln = ln; qe.k = (uint32_t)k; }
                            }
                        }
                        const uint32_t mb = __ballot_sync(0xffffffffu, matched);
                        if (mb) {
                            if (matched) queue[qn + __popc(mb & lt)] = qe;
                            qn += __popc(mb);
                            __syncwarp();
                            if (qn >= 32u) drain(32u);
                        }
                    }
                    nl_seen += row_nl;
                    if (row + 1 >= in_seg_rows && nl_seen >= need_nl) break;
                }
                if (qn) drain(qn);
            }
            if (pn > 32u) flush();
            flush();
        }
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
        s->ctr_base += (unsigned long long)n_tiles + (unsigned long long)grid;
        ++launched;
    }
    s->epoch = (s->epoch % 0x3FFFFFFEu) + 1u;
    a.epoch = s->epoch; a.ctr_base = s->ctr_base;
    a.line_lo = n_train_lines; a.line_hi = ~0ull; a.range_check = n_train_lines > 0 ? 1 : 0;
    a.zero_fill = n_train_lines > 0 ? 0 : 1; a.finalize = 1;
    dm_k_tile<false><<<grid, DMT_THREADS, 0, st>>>(a);
    s->ctr_base += (unsigned long long)n_tiles + (unsigned long long)grid;
    ++launched;
    if (cudaGetLastError() != cudaSuccess) return DM_ERR_CUDA;
    return launched;
}
#endif  // DM_EMU
