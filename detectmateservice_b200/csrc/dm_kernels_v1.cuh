// dm_kernels_v1.cuh -- first correct CUDA path: line index (3 small kernels) + a
// warp-per-record tokenizer/detector.  Kept as the reference implementation the fused
// tile kernel (dm_kernels_tile.cuh) is checked against on the GPU.
//
// What it replaces: per record, MatcherParser (header/fields) + NewValueDetector.train /
// .detect from the un-vendored detectmatelibrary, driven by
// /root/reference/src/service/core.py:201-203.  Rules: DESIGN.md R-tok L1-L6, R-spec 1-4.
#pragma once
#include "dm_device.cuh"

#define DM_TILE_BYTES 4096
#define DM_TILE_THREADS 256

// ---------------------------------------------------------------------------------------
// K1: newline count per 4 KiB tile (16-byte vector loads, one chunk per thread).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t dm_chunk_newlines(const uint8_t* __restrict__ buf, uint64_t nbytes,
                                                      uint64_t chunk_off) {
    if (chunk_off >= nbytes) return 0;
    uint4 v = __ldg(reinterpret_cast<const uint4*>(buf + chunk_off));
    uint32_t m = dm_mask16_eq(v, 0x0A0A0A0Au);
    uint64_t rem = nbytes - chunk_off;
    if (rem < 16) m &= (1u << rem) - 1u;
    return m;
}

__global__ void __launch_bounds__(DM_TILE_THREADS)
dm_k_count_newlines(const uint8_t* __restrict__ buf, uint64_t nbytes, uint32_t n_tiles,
                    uint32_t* __restrict__ tile_counts) {
    __shared__ uint32_t warp_sums[DM_TILE_THREADS / 32];
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        uint64_t off = (uint64_t)tile * DM_TILE_BYTES + (uint64_t)threadIdx.x * 16;
        uint32_t c = __popc(dm_chunk_newlines(buf, nbytes, off));
        c = __reduce_add_sync(0xffffffffu, c);
        if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = c;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t s = 0;
            for (int w = 0; w < DM_TILE_THREADS / 32; ++w) s += warp_sums[w];
            tile_counts[tile] = s;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// K2: exclusive scan of the tile counts (one CTA), batch header, statistics bookkeeping.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
dm_k_scan_tiles(const uint8_t* __restrict__ buf, uint64_t nbytes, uint32_t n_tiles,
                const uint32_t* __restrict__ tile_counts, uint32_t* __restrict__ tile_base,
                uint32_t* __restrict__ line_start, uint64_t max_lines, uint64_t n_train_lines,
                DmBatchHeader* __restrict__ hdr, unsigned long long* __restrict__ stats) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_tiles; base += 1024) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = (i < n_tiles) ? tile_counts[i] : 0;
        uint32_t x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
            if ((threadIdx.x & 31) >= d) x += y;
        }
        if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t t = warp_tot[threadIdx.x];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t y = __shfl_up_sync(0xffffffffu, t, d);
                if (threadIdx.x >= d) t += y;
            }
            warp_tot[threadIdx.x] = t;   // inclusive over warps
        }
        __syncthreads();
        uint32_t warp_excl = (threadIdx.x >> 5) ? warp_tot[(threadIdx.x >> 5) - 1] : 0;
        uint32_t carry = carry_s;
        if (i < n_tiles) tile_base[i] = carry + warp_excl + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_excl + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        unsigned long long nl = carry_s;
        bool tail = nbytes > 0 && buf[nbytes - 1] != 0x0A;
        unsigned long long n_lines = nl + (tail ? 1ull : 0ull);
        hdr->n_newlines = nl;
        unsigned int err = 0;
        if (n_lines > max_lines) { err |= DM_DEVERR_TOO_MANY_LINES; n_lines = max_lines; }
        hdr->n_lines = n_lines;
        if (err) atomicOr(&hdr->error, err);
        line_start[0] = 0;
        if (tail && nl + 1 <= max_lines) line_start[nl + 1] = (uint32_t)(nbytes + 1);
        unsigned long long tr = n_train_lines < n_lines ? n_train_lines : n_lines;
        stats[0] += n_lines;
        stats[1] += tr;
        stats[2] += n_lines - tr;
        stats[5] += nbytes;
    }
}

// ---------------------------------------------------------------------------------------
// K3: line_start[i] = byte offset of record i (ordered compaction of the newline masks).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DM_TILE_THREADS)
dm_k_line_starts(const uint8_t* __restrict__ buf, uint64_t nbytes, uint32_t n_tiles,
                 const uint32_t* __restrict__ tile_base, uint32_t* __restrict__ line_start,
                 uint64_t max_lines) {
    __shared__ uint32_t warp_tot[DM_TILE_THREADS / 32];
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        uint64_t off = (uint64_t)tile * DM_TILE_BYTES + (uint64_t)threadIdx.x * 16;
        uint32_t m = dm_chunk_newlines(buf, nbytes, off);
        uint32_t c = __popc(m);
        uint32_t x = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
            if ((threadIdx.x & 31) >= d) x += y;
        }
        if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = x;
        __syncthreads();
        uint32_t excl = x - c;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) excl += warp_tot[w];
        uint64_t idx = (uint64_t)tile_base[tile] + excl + 1;   // record that starts after this newline
        while (m) {
            uint32_t b = __ffs(m) - 1;
            m &= m - 1;
            if (idx <= max_lines) line_start[idx] = (uint32_t)(off + b + 1);
            ++idx;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// K4: warp-per-record tokenizer + detector.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int dm_match_key_g(const uint8_t* __restrict__ buf, uint64_t s, uint64_t q,
                                              const DmKeys& sk) {
    for (uint32_t k = 0; k < sk.n; ++k) {
        uint32_t len = sk.len[k];
        if (q - s < len) continue;
        uint64_t st = q - len;
        if (buf[q - 1] != sk.bytes[k][len - 1]) continue;
        if (st > s) {
            uint8_t c = buf[st - 1];
            if (c != 0x20 && c != 0x27) continue;
        }
        bool ok = true;
        for (uint32_t i = 0; i + 1 < len; ++i)
            if (buf[st + i] != sk.bytes[k][i]) { ok = false; break; }
        if (ok) return (int)k;
    }
    return -1;
}

// Fingerprint of the value that starts at v (R-tok L5: ends at the first space outside
// double quotes, or at the record end e).
__device__ __forceinline__ uint64_t dm_hash_value_g(const uint8_t* __restrict__ buf, uint64_t v, uint64_t e) {
    DmHashState st;
    dm_hash_init(st);
    uint32_t w = 0, nb = 0, n = 0, vq = 0;
    for (uint64_t p = v; p < e; ++p) {
        uint32_t c = buf[p];
        if (c == 0x20 && !vq) break;
        if (c == 0x22) vq ^= 1;
        w |= c << (8 * nb);
        ++nb; ++n;
        if (nb == 4) { dm_hash_word(st, w); w = 0; nb = 0; }
    }
    if (nb) dm_hash_word(st, w);
    return dm_hash_final(st, n);
}

// (DmDetectArgs lives in dm_device.cuh: shared with dm_kernels_format.cuh)

template <bool TRAIN>
__global__ void __launch_bounds__(256) dm_k_detect_lines(DmDetectArgs a) {
    __shared__ DmKeys sk;
    __shared__ unsigned int s_unk[DM_MAX_KEYS];
    __shared__ unsigned long long s_anom, s_score;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(a.keys);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&sk);
        for (uint32_t i = threadIdx.x; i < sizeof(DmKeys) / 4; i += blockDim.x) dst[i] = src[i];
        if (threadIdx.x < DM_MAX_KEYS) s_unk[threadIdx.x] = 0;
        if (threadIdx.x == 0) { s_anom = 0; s_score = 0; }
    }
    __syncthreads();

    const uint8_t* __restrict__ buf = a.buf;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warps_total = (uint64_t)gridDim.x * (blockDim.x >> 5);
    const uint64_t warp_id = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    uint64_t n_lines = a.hdr_in->n_lines;
    uint64_t hi = a.line_hi < n_lines ? a.line_hi : n_lines;
    const uint32_t lt = dm_lanemask_lt();

    for (uint64_t line = a.line_lo + warp_id; line < hi; line += warps_total) {
        const uint64_t s = a.line_start[line];
        const uint64_t e = (uint64_t)a.line_start[line + 1] - 1;   // the '\n' (or nbytes)
        uint32_t seen = 0;        // warp-uniform: keys already resolved in this record (L6)
        uint32_t unknown = 0;     // lane-local, OR-reduced at the end
        uint32_t carry_inq = 0;   // warp-uniform: in_quote at the start of the current step

        for (uint64_t base = s & ~3ull; base < e; base += 128) {
            const uint64_t pos = base + (uint64_t)lane * 4;
            uint32_t w = 0, valid = 0;
            if (pos < e) {
                w = __ldg(reinterpret_cast<const uint32_t*>(buf + pos));
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (pos + j >= s && pos + j < e) valid |= 1u << j;
            }
            const uint32_t dq = dm_nib_eq(w, 0x22222222u) & valid;
            const uint32_t eq = dm_nib_eq(w, 0x3D3D3D3Du) & valid;
            const uint32_t bal = __ballot_sync(0xffffffffu, __popc(dq) & 1);
            const uint32_t inq_lane = carry_inq ^ (__popc(bal & lt) & 1);
            carry_inq ^= __popc(bal) & 1;

            // candidate '=' positions outside quotes whose preceding bytes spell a monitored key
            int kidx[4] = {-1, -1, -1, -1};
            uint32_t matchbits = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if ((eq >> j) & 1) {
                    uint32_t inq = inq_lane ^ (__popc(dq & ((1u << j) - 1u)) & 1);
                    if (!inq) {
                        int k = dm_match_key_g(buf, s, pos + j, sk);
                        if (k >= 0) { kidx[j] = k; matchbits |= 1u << j; }
                    }
                }
            }
            // resolve first-occurrence-wins in position order (lane-major, then byte)
            uint32_t any = __ballot_sync(0xffffffffu, matchbits != 0);
            uint32_t winbits = 0;
            while (any) {
                const int L = __ffs(any) - 1;
                any &= any - 1;
                const uint32_t mb = __shfl_sync(0xffffffffu, matchbits, L);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int k = __shfl_sync(0xffffffffu, kidx[j], L);
                    if ((mb >> j) & 1) {
                        if (!((seen >> k) & 1)) {
                            seen |= 1u << k;
                            if ((int)lane == L) winbits |= 1u << j;
                        }
                    }
                }
            }
            // winners hash their value and consult the known-set table
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if ((winbits >> j) & 1) {
                    const int k = kidx[j];
                    const uint64_t fp = dm_hash_value_g(buf, pos + j + 1, e);
                    const uint64_t key = dm_make_key(fp, sk.salt[k]);
                    if (TRAIN) {
                        dm_table_insert(a.table, key, &a.hdr->error);
                    } else if (!dm_table_contains(a.table, key)) {
                        unknown |= 1u << k;
                    }
                }
            }
        }
        unknown = __reduce_or_sync(0xffffffffu, unknown);
        if (lane == 0) {
            const uint32_t cnt = __popc(unknown);
            if (line < a.out_cap) {
                if (a.flags) a.flags[line] = cnt ? 1 : 0;
                if (a.scores) a.scores[line] = (float)cnt;
            }
            if (cnt) {
                atomicAdd(&s_anom, 1ull);
                atomicAdd(&s_score, (unsigned long long)cnt);
                uint32_t m = unknown;
                while (m) { int k = __ffs(m) - 1; m &= m - 1; atomicAdd(&s_unk[k], 1u); }
                unsigned int idx = atomicAdd(&a.hdr->anomaly_list_count, 1u);
                if (idx < a.anomaly_cap) {
                    dm_anomaly_t r; r.line = (uint32_t)line; r.mask = unknown; r.offset = s;
                    a.anomalies[idx] = r;
                }
            }
        }
    }
    __syncthreads();
    if (!TRAIN) {
        if (threadIdx.x == 0 && s_anom) {
            atomicAdd(&a.hdr->n_anomalies, s_anom);
            atomicAdd(&a.stats[3], s_anom);
            atomicAdd(&a.stats[4], s_score);
        }
        if (threadIdx.x < DM_MAX_KEYS && s_unk[threadIdx.x])
            atomicAdd(&a.stats[8 + threadIdx.x], (unsigned long long)s_unk[threadIdx.x]);
    }
}
