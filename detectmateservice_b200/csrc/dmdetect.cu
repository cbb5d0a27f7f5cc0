// dmdetect.cu -- host side of libdmdetect.so: the C ABI of include/dmdetect.h.
//
// Replaces, for one message of raw records, the Python call chain
//   Service.process -> CoreComponent.process -> NewValueDetector.train/detect
// (/root/reference/src/service/core.py:176-206; detector in the un-vendored
// detectmatelibrary).  No CPU path: every entry point that computes needs the device.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "dm_device.cuh"
#include "dm_kernels_index.cuh"
#include "dm_kernels_stream.cuh"
#include "dm_kernels_values.cuh"
#include "dm_kernels_records.cuh"
#include "dm_kernels_format.cuh"
#include "dm_kernels_lanes.cuh"
#include "dm_format_host.h"

static void (*g_nccl_destroy_hook)(void*) = nullptr;   // set once NCCL is loaded (dm_nccl_load)

// ---------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int dm_fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define DM_CUDA(call)                                                                         \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess)                                                               \
            return dm_fail(DM_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                           __FILE__, __LINE__);                                               \
    } while (0)

extern "C" const char* dm_last_error(void) { return g_err; }
extern "C" int dm_abi_version(void) { return DM_ABI_VERSION; }

extern "C" uint64_t dm_table_key(uint32_t field, const uint8_t* value, uint32_t len) {
    return dm_make_key(dm_fp64_bytes(value, len), dm_field_salt(field));
}

// ---------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------
struct dm_handle {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;       // the handle's own stream
    cudaStream_t last_stream = nullptr;  // stream of the most recent enqueue
    uint64_t max_batch_bytes = 0, max_lines = 0;
    uint32_t table_log2 = 0;
    uint32_t n_keys = 0;
    int kernel_variant = 6;              // 6 = stream (default), 5 = lanes (one thread per record; DM_KERNEL=lanes)

    DmKeys h_keys;
    DmKeys* d_keys = nullptr;
    uint8_t* d_in = nullptr;             // staging for host input
    uint32_t* d_line_start = nullptr;
    uint8_t* d_flags = nullptr;
    float* d_scores = nullptr;
    DmBatchHeader* d_hdr = nullptr;
    DmBatchHeader* h_hdr = nullptr;      // pinned
    dm_anomaly_t* d_anoms = nullptr;
    uint32_t anomaly_cap = 0;
    unsigned long long* d_stats = nullptr;         // DM_STATS_WORDS
    unsigned long long* d_stats_exported = nullptr;
    unsigned long long* d_stats_global = nullptr;
    unsigned long long* h_stats = nullptr;         // pinned
    DmTable table;
    uint64_t novel_exported = 0;         // novel keys already shipped in a window
    uint64_t novel_pending = 0;          // novel keys the last window could not carry (more than DM_WINDOW_KEYS learnt)
    DmRowsScratch rows;                  // row / record index scratch (K_A: lanes and log_format kernels)
    DmxScratch dmx;                      // stream-variant scratch (default kernel)
    uint64_t last_nbytes = 0;
    uint32_t* d_vals = nullptr;          // record mode: offsets / fields / record_of
    uint64_t vals_cap = 0;
    uint32_t* d_masks = nullptr;
    DmMonitors* d_mons = nullptr;        // record mode on the device
    bool mons_set = false;
    DmMonitors h_mons;             // host copy (dm_set_monitors + dm_set_combos)
    DmFormat* d_fmt = nullptr;     // log_format + templates (dm_set_format)
    uint8_t* d_norm = nullptr;     // R-norm: the message's normalised Content text (dm_set_format_ex with flags)
    uint32_t fmt_norm = 0;
    void* nccl_comm[2] = {nullptr, nullptr};   // own NCCL communicators (dm_nccl_init): windows alternate between them
    uint32_t nccl_n = 0, nccl_rank = 0, nccl_world = 1;
    unsigned long long* d_win[2] = {nullptr, nullptr};   // window exchange buffers of dm_window_allreduce
    cudaEvent_t ev_win_done[2] = {nullptr, nullptr};
    uint64_t win_seq = 0;
    bool fmt_set = false;
    uint32_t fmt_slots = 1;        // DmFormat.max_slots (dynamic shared memory of the thread-per-record kernel)
    // pipelined host path: two slots
    struct Slot {
        uint8_t* d_in = nullptr; uint8_t* d_flags = nullptr; float* d_scores = nullptr;
        DmBatchHeader* d_hdr = nullptr; DmBatchHeader* h_hdr = nullptr; dm_anomaly_t* d_anoms = nullptr;
        uint8_t* h_flags = nullptr; float* h_scores = nullptr;
        cudaEvent_t ev_in = nullptr, ev_comp = nullptr, ev_hdr = nullptr;
        cudaStream_t st_out = nullptr;
        bool busy = false;
    } slots[2];
    cudaStream_t st_in = nullptr;
    bool slots_ready = false;
    // measurement support
    bool profile = false;
    std::vector<cudaEvent_t> ev;         // pairs: ev[2i] start, ev[2i+1] stop
    size_t ev_used = 0;
    uint64_t launches = 0;               // kernels launched since dm_create
};

static const size_t DM_PROFILE_PAIRS = 4096;

static void dm_prof_mark(dm_handle* h, cudaStream_t st, int which) {
    if (!h->profile || h->ev_used / 2 >= DM_PROFILE_PAIRS) return;
    if (which == 0 && (h->ev_used & 1)) return;
    if (which == 1 && !(h->ev_used & 1)) return;
    cudaEventRecord(h->ev[h->ev_used], st);
    h->ev_used++;
}

static void dm_prof_mark_cb(void* ctx, cudaStream_t st, int which) { dm_prof_mark((dm_handle*)ctx, st, which); }

static const uint32_t DM_WINDOW_KEYS = 1u << 16;    // keys one rank can ship per window

static int dm_pick_stream(dm_handle* h, void* stream, cudaStream_t* out) {
    *out = stream ? (cudaStream_t)stream : h->stream;
    h->last_stream = *out;
    return DM_OK;
}

// Anything this library enqueues that is not a stream-variant detect launch ends the chain of
// overlapping launches on that stream: the next detect launch is then ordered behind it in full.
static void dm_break_chain(dm_handle* h, cudaStream_t st) {
    if (h->dmx.chain_stream == st) h->dmx.chain_stream = nullptr;
}

extern "C" int dm_destroy(dm_handle* h);
// Everything dm_create allocates; on failure dm_create frees whatever was allocated so far (dm_destroy).
static int dm_create_fill(dm_handle* h, int device, const cudaDeviceProp& prop, uint32_t n_keys, const uint8_t* keys_blob,
                          const uint32_t* key_lens, uint64_t max_batch_bytes, uint64_t max_lines, uint32_t table_log2_slots) {
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    h->max_batch_bytes = max_batch_bytes;
    h->max_lines = max_lines ? max_lines : std::max<uint64_t>(max_batch_bytes / 8, 1024);
    h->table_log2 = table_log2_slots;
    h->n_keys = n_keys;
    memset(&h->h_keys, 0, sizeof(DmKeys));
    h->h_keys.n = n_keys;
    uint64_t off = 0;
    for (uint32_t k = 0; k < n_keys; ++k) {
        uint32_t len = key_lens[k];
        if (len == 0 || len > DM_MAX_KEYLEN) { return dm_fail(DM_ERR_ARG, "key %u: length %u not in 1..%d", k, len, DM_MAX_KEYLEN); }
        for (uint32_t i = 0; i < len; ++i) {
            uint8_t c = keys_blob[off + i];
            if (c == 0x20 || c == 0x22 || c == 0x27 || c == 0x3D || c == 0x0A) { return dm_fail(DM_ERR_ARG, "key %u holds a separator byte 0x%02x", k, c); }
            h->h_keys.bytes[k][i] = c;
        }
        for (uint32_t j = 0; j < k; ++j)
            if (h->h_keys.len[j] == len && memcmp(h->h_keys.bytes[j], h->h_keys.bytes[k], len) == 0) { return dm_fail(DM_ERR_ARG, "key %u duplicates key %u", k, j); }
        h->h_keys.len[k] = len;
        h->h_keys.salt[k] = dm_field_salt(k);
        off += len;
    }
    dm_keys_finalize_host(&h->h_keys);

    DM_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->last_stream = h->stream;
    DM_CUDA(cudaMalloc(&h->d_keys, sizeof(DmKeys)));
    DM_CUDA(cudaMemcpy(h->d_keys, &h->h_keys, sizeof(DmKeys), cudaMemcpyHostToDevice));
    DM_CUDA(cudaMalloc(&h->d_in, max_batch_bytes + 256));
    DM_CUDA(cudaMemset(h->d_in, 0, max_batch_bytes + 256));
    DM_CUDA(cudaMalloc(&h->d_line_start, (h->max_lines + 2) * sizeof(uint32_t)));
    DM_CUDA(cudaMalloc(&h->d_flags, h->max_lines + 16));
    DM_CUDA(cudaMalloc(&h->d_scores, (h->max_lines + 4) * sizeof(float)));
    DM_CUDA(cudaMalloc(&h->d_hdr, sizeof(DmBatchHeader)));
    DM_CUDA(cudaMemset(h->d_hdr, 0, sizeof(DmBatchHeader)));
    DM_CUDA(cudaHostAlloc(&h->h_hdr, sizeof(DmBatchHeader), cudaHostAllocDefault));
    memset(h->h_hdr, 0, sizeof(DmBatchHeader));
    h->anomaly_cap = (uint32_t)std::min<uint64_t>(h->max_lines, 1u << 20);
    DM_CUDA(cudaMalloc(&h->d_anoms, (uint64_t)h->anomaly_cap * sizeof(dm_anomaly_t)));
    DM_CUDA(cudaMalloc(&h->d_stats, 3 * DM_STATS_WORDS * sizeof(unsigned long long)));
    DM_CUDA(cudaMemset(h->d_stats, 0, 3 * DM_STATS_WORDS * sizeof(unsigned long long)));
    h->d_stats_exported = h->d_stats + DM_STATS_WORDS;
    h->d_stats_global = h->d_stats + 2 * DM_STATS_WORDS;
    DM_CUDA(cudaHostAlloc(&h->h_stats, DM_STATS_WORDS * sizeof(unsigned long long), cudaHostAllocDefault));

    const uint64_t cap = 1ull << table_log2_slots;
    h->table.mask = (uint32_t)(cap - 1);
    h->table.limit = (uint32_t)(cap / 2);
    h->table.novel_cap = (uint32_t)(cap / 2);
    DM_CUDA(cudaMalloc(&h->table.slots, cap * sizeof(unsigned long long)));
    DM_CUDA(cudaMemset(h->table.slots, 0, cap * sizeof(unsigned long long)));
    DM_CUDA(cudaMalloc(&h->table.novel, (cap / 2) * sizeof(unsigned long long)));
    DM_CUDA(cudaMalloc(&h->table.count, 2 * sizeof(unsigned long long)));
    DM_CUDA(cudaMemset(h->table.count, 0, 2 * sizeof(unsigned long long)));
    h->table.novel_count = h->table.count + 1;

    int rc = dm_rows_scratch_create(&h->rows, max_batch_bytes, h->sm_count);
    if (rc != DM_OK) { return dm_fail(DM_ERR_CUDA, "row index scratch allocation failed: %s", cudaGetErrorString(cudaGetLastError())); }
    rc = dmx_scratch_create(&h->dmx, h->h_keys, max_batch_bytes, h->anomaly_cap, h->sm_count);
    if (rc == DM_ERR_ARG) return dm_fail(DM_ERR_ARG, "DM_STREAM_RECHECK=%s: thread, chain or auto", getenv("DM_STREAM_RECHECK"));
    if (rc != DM_OK) { return dm_fail(DM_ERR_CUDA, "stream scratch allocation failed: %s", cudaGetErrorString(cudaGetLastError())); }
    { const char* ov = getenv("DM_OVERLAP"); if (ov) h->dmx.overlap = atoi(ov) != 0; }
    const char* env = getenv("DM_KERNEL");
    if (env && env[0]) {
        if (strcmp(env, "lanes") == 0) h->kernel_variant = 5;
        else if (strcmp(env, "stream") == 0) h->kernel_variant = 6;
        else return dm_fail(DM_ERR_ARG, "DM_KERNEL=%s: the key=value kernels are 'stream' (default) and 'lanes'", env);
    }
    DM_CUDA(cudaDeviceSynchronize());
    return DM_OK;
}


extern "C" int dm_create(int device, uint32_t n_keys, const uint8_t* keys_blob, const uint32_t* key_lens,
                         uint64_t max_batch_bytes, uint64_t max_lines, uint32_t table_log2_slots,
                         dm_handle** out) {
    if (!out) return dm_fail(DM_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (n_keys > DM_MAX_KEYS) return dm_fail(DM_ERR_ARG, "n_keys %u > %d", n_keys, DM_MAX_KEYS);
    if (n_keys && (!keys_blob || !key_lens)) return dm_fail(DM_ERR_ARG, "keys missing");
    if (max_batch_bytes == 0 || max_batch_bytes > 0xFFFFFF00ull)
        return dm_fail(DM_ERR_ARG, "max_batch_bytes must be in 1..2^32-256");
    if (table_log2_slots < 10 || table_log2_slots > 28)
        return dm_fail(DM_ERR_ARG, "table_log2_slots must be in 10..28");
    int n_dev = 0;
    cudaError_t ce = cudaGetDeviceCount(&n_dev);
    if (ce != cudaSuccess || n_dev <= 0)
        return dm_fail(DM_ERR_NO_DEVICE, "no CUDA device (%s); this library has no CPU path",
                       ce == cudaSuccess ? "device count 0" : cudaGetErrorString(ce));
    if (device < 0 || device >= n_dev) return dm_fail(DM_ERR_ARG, "device %d out of range (0..%d)", device, n_dev - 1);
    cudaDeviceProp prop;
    DM_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return dm_fail(DM_ERR_NO_DEVICE, "device %d is sm_%d%d; libdmdetect is built for sm_100a only", device,
                       prop.major, prop.minor);
    DM_CUDA(cudaSetDevice(device));

    dm_handle* h = new (std::nothrow) dm_handle();
    if (!h) return dm_fail(DM_ERR_ARG, "out of host memory");
    const int rc = dm_create_fill(h, device, prop, n_keys, keys_blob, key_lens, max_batch_bytes, max_lines, table_log2_slots);
    if (rc != DM_OK) {
        char keep[sizeof(g_err)];
        memcpy(keep, g_err, sizeof(keep));
        dm_destroy(h);
        memcpy(g_err, keep, sizeof(keep));
        return rc;
    }
    *out = h;
    return DM_OK;
}

extern "C" int dm_set_overlap(dm_handle* h, int on) {
    if (!h) return dm_fail(DM_ERR_ARG, "handle is NULL");
    h->dmx.overlap = on != 0;
    h->dmx.chain_stream = nullptr;
    return DM_OK;
}

extern "C" int dm_profile_enable(dm_handle* h, int on) {
    if (!h) return dm_fail(DM_ERR_ARG, "handle is NULL");
    DM_CUDA(cudaSetDevice(h->device));
    if (on && h->ev.empty()) {
        h->ev.resize(2 * DM_PROFILE_PAIRS);
        for (auto& e : h->ev) DM_CUDA(cudaEventCreate(&e));
    }
    h->profile = on != 0;
    h->ev_used = 0;
    return DM_OK;
}

extern "C" int dm_profile_read(dm_handle* h, double* kernel_ms_sum, uint64_t* n_timed, uint64_t* kernels_launched_total) {
    if (!h) return dm_fail(DM_ERR_ARG, "handle is NULL");
    DM_CUDA(cudaSetDevice(h->device));
    DM_CUDA(cudaStreamSynchronize(h->last_stream));
    double sum = 0.0;
    const size_t pairs = h->ev_used / 2;
    for (size_t i = 0; i < pairs; ++i) {
        float ms = 0.f;
        DM_CUDA(cudaEventElapsedTime(&ms, h->ev[2 * i], h->ev[2 * i + 1]));
        sum += ms;
    }
    h->ev_used = 0;
    if (kernel_ms_sum) *kernel_ms_sum = sum;
    if (n_timed) *n_timed = pairs;
    if (kernels_launched_total) *kernels_launched_total = h->launches;
    return DM_OK;
}

extern "C" int dm_destroy(dm_handle* h) {
    if (!h) return DM_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (auto& e : h->ev) cudaEventDestroy(e);
    dm_rows_scratch_destroy(&h->rows);
    dmx_scratch_destroy(&h->dmx);
    cudaFree(h->d_keys); cudaFree(h->d_in);
    cudaFree(h->d_line_start); cudaFree(h->d_flags); cudaFree(h->d_scores); cudaFree(h->d_hdr);
    cudaFreeHost(h->h_hdr); cudaFree(h->d_anoms); cudaFree(h->d_stats); cudaFreeHost(h->h_stats);
    cudaFree(h->table.slots); cudaFree(h->table.novel); cudaFree(h->table.count);
    cudaFree(h->d_vals); cudaFree(h->d_masks); cudaFree(h->d_mons); cudaFree(h->d_fmt); cudaFree(h->d_norm);
    for (int k = 0; k < 2; ++k) {
        cudaFree(h->d_win[k]);
        if (h->ev_win_done[k]) cudaEventDestroy(h->ev_win_done[k]);
        if (h->nccl_comm[k] && g_nccl_destroy_hook) g_nccl_destroy_hook(h->nccl_comm[k]);
    }
    for (auto& sl : h->slots) {
        cudaFree(sl.d_in); cudaFree(sl.d_flags); cudaFree(sl.d_scores); cudaFree(sl.d_hdr); cudaFree(sl.d_anoms);
        cudaFreeHost(sl.h_hdr); cudaFreeHost(sl.h_flags); cudaFreeHost(sl.h_scores);
        if (sl.ev_in) cudaEventDestroy(sl.ev_in);
        if (sl.ev_comp) cudaEventDestroy(sl.ev_comp);
        if (sl.ev_hdr) cudaEventDestroy(sl.ev_hdr);
        if (sl.st_out) cudaStreamDestroy(sl.st_out);
    }
    if (h->st_in) cudaStreamDestroy(h->st_in);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return DM_OK;
}

static int dm_check_device_errors(dm_handle* h) {
    unsigned int err = h->h_hdr->error;
    if (h->h_hdr->anomaly_list_count > h->anomaly_cap) err |= DM_DEVERR_ANOMALY_OVERFLOW;
    if (err & DM_DEVERR_TABLE_FULL)
        return dm_fail(DM_ERR_TABLE_FULL, "known-set table over its load limit (2^%u slots); recreate with a larger table_log2_slots", h->table_log2);
    if (err & DM_DEVERR_NOVEL_OVERFLOW)
        return dm_fail(DM_ERR_TABLE_FULL, "more than %u distinct values learnt; novel-key list full", h->table.novel_cap);
    if (err & DM_DEVERR_ANOMALY_OVERFLOW)
        return dm_fail(DM_ERR_CAPACITY, "more than %u unknown values in one batch; the anomaly list is full", h->anomaly_cap);
    if (err & DM_DEVERR_TOO_MANY_LINES)
        return dm_fail(DM_ERR_CAPACITY, "batch holds more than max_lines=%llu records", (unsigned long long)h->max_lines);
    return DM_OK;
}

// ---------------------------------------------------------------------------------------
// the hot path
// ---------------------------------------------------------------------------------------
extern "C" int dm_process_lines(dm_handle* h, const uint8_t* buf, uint64_t nbytes, int buf_on_device,
                                uint64_t n_train_lines, uint8_t* flags_out, float* scores_out,
                                uint64_t out_cap_lines, int out_on_device, uint64_t* n_lines_out,
                                uint64_t* n_anomalies_out, void* stream_) {
    if (!h) return dm_fail(DM_ERR_ARG, "handle is NULL");
    if (nbytes > h->max_batch_bytes)
        return dm_fail(DM_ERR_CAPACITY, "message of %llu bytes exceeds max_batch_bytes=%llu", (unsigned long long)nbytes, (unsigned long long)h->max_batch_bytes);
    if (nbytes && !buf) return dm_fail(DM_ERR_ARG, "buf is NULL");
    if (buf_on_device && (((uintptr_t)buf) & 15)) return dm_fail(DM_ERR_ARG, "device buf must be 16-byte aligned");
    DM_CUDA(cudaSetDevice(h->device));
    cudaStream_t st;
    dm_pick_stream(h, stream_, &st);

    const uint8_t* d_buf = buf;
    if (!buf_on_device) {
        if (nbytes) DM_CUDA(cudaMemcpyAsync(h->d_in, buf, nbytes, cudaMemcpyHostToDevice, st));
        DM_CUDA(cudaMemsetAsync(h->d_in + nbytes, 0, 64, st));
        d_buf = h->d_in;
    }
    const bool host_out = !out_on_device && (flags_out || scores_out);
    uint8_t* d_flags = (out_on_device && flags_out) ? flags_out : h->d_flags;
    float* d_scores = (out_on_device && scores_out) ? scores_out : h->d_scores;
    const uint64_t dev_cap = out_on_device ? ((flags_out || scores_out) ? out_cap_lines : h->max_lines) : h->max_lines;
    // when only one of the two outputs is a caller device buffer the other stays internal
    const uint64_t cap_flags = (out_on_device && flags_out) ? out_cap_lines : h->max_lines;
    const uint64_t cap_scores = (out_on_device && scores_out) ? out_cap_lines : h->max_lines;
    const uint64_t out_cap = std::min(cap_flags, cap_scores);
    (void)dev_cap;

    bool stream_kernel = false;
    // (the stream kernel and K_A write the per-batch header themselves)
    if (nbytes == 0) DM_CUDA(cudaMemsetAsync(h->d_hdr, 0, sizeof(DmBatchHeader), st));
    h->last_nbytes = nbytes;

    if (h->fmt_set && h->mons_set && h->h_mons.n_combos > 0)
        return dm_fail(DM_ERR_STATE, "combination monitors are not evaluated in log_format mode (use key=value records or ParserSchema input)");
    if (h->fmt_set) {
        // log_format / template mode: K_A writes the record index, then one THREAD per record
        const uint32_t n_rows = (uint32_t)((nbytes + DMR_ROW - 1) / DMR_ROW);
        if (n_rows > h->rows.max_rows) return dm_fail(DM_ERR_CAPACITY, "message too large for the row index");
        if (n_rows > 0) {
            DmRowsArgs ra;
            ra.buf = d_buf; ra.nbytes = nbytes; ra.n_rows = n_rows;
            ra.n_tiles = (n_rows + DMR_TILE_ROWS - 1) / DMR_TILE_ROWS;
            ra.row_prefix = h->rows.d_row_prefix; ra.tile_state = h->rows.d_tile_state;
            h->rows.epoch = (h->rows.epoch % 0x3FFFFFFEu) + 1u;
            ra.epoch = h->rows.epoch;
            ra.keys = h->d_keys; ra.table = h->table; ra.flags = d_flags; ra.scores = d_scores; ra.out_cap = out_cap;
            ra.anomalies = h->d_anoms; ra.anomaly_cap = h->anomaly_cap; ra.hdr = h->d_hdr; ra.stats = h->d_stats;
            ra.row_ctr = h->rows.d_row_ctr; ra.n_train_lines = n_train_lines; ra.max_lines = h->max_lines;
            ra.line_lo = 0; ra.line_hi = ~0ull; ra.ctr_base = h->rows.ctr_base; ra.aux_counts = nullptr;
            ra.line_start = h->d_line_start; ra.group = DMR_GROUP; ra.static_rows = 0; ra.timeline = nullptr;
            dm_launch_pdl(dm_k_rowindex, ra.n_tiles, DMR_A_THREADS, st, h->rows.pdl, ra);
            DmDetectArgs a;
            a.buf = d_buf; a.line_start = h->d_line_start; a.hdr_in = h->d_hdr; a.hdr = h->d_hdr;
            a.keys = h->d_keys; a.table = h->table; a.flags = d_flags; a.scores = d_scores; a.out_cap = out_cap;
            a.anomalies = h->d_anoms; a.anomaly_cap = h->anomaly_cap; a.stats = h->d_stats; a.nbytes = nbytes; a.combos = nullptr;
            const uint64_t max_recs = std::min<uint64_t>(nbytes / 2 + 1, h->max_lines);
            const uint64_t want = (max_recs + DM_FMTL_THREADS - 1) / DM_FMTL_THREADS;
            const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(want, (uint64_t)h->sm_count * 32));
            const size_t smem = (size_t)2 * h->fmt_slots * DM_FMTL_THREADS * sizeof(uint2);
            const bool norm = h->fmt_norm != 0;
            if (n_train_lines > 0) {
                a.line_lo = 0; a.line_hi = n_train_lines;
                dm_launch_pdl_smem(norm ? dm_k_format_lanes<true, true> : dm_k_format_lanes<true, false>, (unsigned)grid,
                                   DM_FMTL_THREADS, smem, st, false, a, (const DmFormat*)h->d_fmt, h->d_norm);
            }
            a.line_lo = n_train_lines; a.line_hi = ~0ull;
            dm_prof_mark(h, st, 0);
            dm_launch_pdl_smem(norm ? dm_k_format_lanes<false, true> : dm_k_format_lanes<false, false>, (unsigned)grid,
                               DM_FMTL_THREADS, smem, st, false, a, (const DmFormat*)h->d_fmt, h->d_norm);
            dm_prof_mark(h, st, 1);
            h->launches += 2 + (n_train_lines > 0 ? 1 : 0);
        }
    } else if (h->kernel_variant == 5 || (h->mons_set && h->h_mons.n_combos > 0)) {
        // one thread per record; the only raw-line kernel that evaluates combination monitors
        const bool combos = h->mons_set && h->h_mons.n_combos > 0;
        const int rc = dm_lanes_launch(&h->rows, h->d_line_start, d_buf, nbytes, n_train_lines, h->d_keys, h->table, d_flags,
                                       d_scores, out_cap, h->d_anoms, h->anomaly_cap, h->d_hdr, h->d_stats, h->max_lines,
                                       h->sm_count, combos ? (const void*)h->d_mons : nullptr, h->n_keys, st, dm_prof_mark_cb, h);
        if (rc < 0) return dm_fail(DM_ERR_CUDA, "lanes kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        h->launches += (uint64_t)rc;
    } else {
        stream_kernel = true;
        const int rc = dmx_launch(&h->dmx, d_buf, nbytes, n_train_lines, h->table, d_flags, d_scores, out_cap, h->d_anoms,
                                  h->anomaly_cap, h->d_hdr, h->d_stats, h->max_lines, st, h->dmx.overlap, dm_prof_mark_cb, h);
        if (rc < 0) return dm_fail(DM_ERR_CUDA, "stream kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        h->launches += (uint64_t)rc;
    }
    DM_CUDA(cudaGetLastError());
    if (!stream_kernel) h->dmx.chain_stream = nullptr;     // (a launch that is not the stream kernel's ends the overlap chain)

    const bool want_sync = host_out || n_lines_out || n_anomalies_out;
    if (!want_sync) return DM_OK;

    DM_CUDA(cudaMemcpyAsync(h->h_hdr, h->d_hdr, sizeof(DmBatchHeader), cudaMemcpyDeviceToHost, st));
    DM_CUDA(cudaStreamSynchronize(st));
    int rc = dm_check_device_errors(h);
    if (rc != DM_OK) return rc;
    const uint64_t n_lines = h->h_hdr->n_lines;
    if (n_lines_out) *n_lines_out = n_lines;
    if (n_anomalies_out) *n_anomalies_out = h->h_hdr->n_anomalies;
    if (host_out) {
        if (n_lines > out_cap_lines)
            return dm_fail(DM_ERR_CAPACITY, "batch holds %llu records, output capacity is %llu", (unsigned long long)n_lines, (unsigned long long)out_cap_lines);
        if (flags_out && n_lines) DM_CUDA(cudaMemcpyAsync(flags_out, h->d_flags, n_lines, cudaMemcpyDeviceToHost, st));
        if (scores_out && n_lines) DM_CUDA(cudaMemcpyAsync(scores_out, h->d_scores, n_lines * sizeof(float), cudaMemcpyDeviceToHost, st));
        DM_CUDA(cudaStreamSynchronize(st));
    }
    return DM_OK;
}

// ---------------------------------------------------------------------------------------
// record mode
// ---------------------------------------------------------------------------------------
extern "C" int dm_process_values(dm_handle* h, const uint8_t* blob, uint64_t blob_bytes, const uint32_t* offsets,
                                 const uint32_t* fields, const uint32_t* record_of, uint32_t n_values,
                                 uint32_t n_records, uint32_t n_train_records, uint64_t record_bytes,
                                 uint8_t* flags_out, float* scores_out, uint32_t* masks_out,
                                 uint64_t* n_anomalies_out) {
    if (!h) return dm_fail(DM_ERR_ARG, "handle is NULL");
    if (n_values && (!blob && blob_bytes)) return dm_fail(DM_ERR_ARG, "blob is NULL");
    if (n_values && (!offsets || !fields || !record_of)) return dm_fail(DM_ERR_ARG, "value arrays missing");
    if (blob_bytes > h->max_batch_bytes) return dm_fail(DM_ERR_CAPACITY, "value blob of %llu bytes exceeds max_batch_bytes", (unsigned long long)blob_bytes);
    if (n_records > h->max_lines) return dm_fail(DM_ERR_CAPACITY, "%u records exceed max_lines=%llu", n_records, (unsigned long long)h->max_lines);
    for (uint32_t i = 0; i < n_values; ++i) {
        if (fields[i] >= h->n_keys || record_of[i] >= n_records || offsets[i] > offsets[i + 1] || offsets[i + 1] > blob_bytes)
            return dm_fail(DM_ERR_ARG, "value %u: field/record/offset out of range", i);
    }
    DM_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    h->last_stream = st;
    dm_break_chain(h, st);
    if (n_values > h->vals_cap) {
        if (h->d_vals) DM_CUDA(cudaFree(h->d_vals));
        h->vals_cap = std::max<uint64_t>(2ull * n_values, 4096);
        DM_CUDA(cudaMalloc(&h->d_vals, (3 * h->vals_cap + 1) * sizeof(uint32_t)));
    }
    if (!h->d_masks) DM_CUDA(cudaMalloc(&h->d_masks, (h->max_lines + 4) * sizeof(uint32_t)));
    uint32_t* d_off = h->d_vals;
    uint32_t* d_fld = d_off + h->vals_cap + 1;
    uint32_t* d_rec = d_fld + h->vals_cap;
    DM_CUDA(cudaMemsetAsync(h->d_hdr, 0, sizeof(DmBatchHeader), st));
    if (n_values) {
        if (blob_bytes) DM_CUDA(cudaMemcpyAsync(h->d_in, blob, blob_bytes, cudaMemcpyHostToDevice, st));
        DM_CUDA(cudaMemcpyAsync(d_off, offsets, (uint64_t)(n_values + 1) * 4, cudaMemcpyHostToDevice, st));
        DM_CUDA(cudaMemcpyAsync(d_fld, fields, (uint64_t)n_values * 4, cudaMemcpyHostToDevice, st));
        DM_CUDA(cudaMemcpyAsync(d_rec, record_of, (uint64_t)n_values * 4, cudaMemcpyHostToDevice, st));
    }
    if (n_records) {
        DM_CUDA(cudaMemsetAsync(h->d_flags, 0, n_records, st));
        DM_CUDA(cudaMemsetAsync(h->d_scores, 0, (uint64_t)n_records * 4, st));
        DM_CUDA(cudaMemsetAsync(h->d_masks, 0, (uint64_t)n_records * 4, st));
    }
    if (n_values) {
        DmValuesArgs a;
        a.blob = h->d_in; a.offsets = d_off; a.fields = d_fld; a.record_of = d_rec; a.n_values = n_values;
        a.n_records = n_records; a.n_train_records = n_train_records; a.keys = h->d_keys; a.table = h->table;
        a.flags = h->d_flags; a.scores = h->d_scores; a.masks = h->d_masks; a.hdr = h->d_hdr; a.stats = h->d_stats;
        const int grid = (int)std::min<uint32_t>((n_values + 255) / 256, (uint32_t)h->sm_count * 8);
        if (n_train_records > 0) { dm_k_values<<<grid, 256, 0, st>>>(a, 0); h->launches++; }
        dm_k_values<<<grid, 256, 0, st>>>(a, 1);
        h->launches++;
        DM_CUDA(cudaGetLastError());
    }
    DM_CUDA(cudaMemcpyAsync(h->h_hdr, h->d_hdr, sizeof(DmBatchHeader), cudaMemcpyDeviceToHost, st));
    if (flags_out && n_records) DM_CUDA(cudaMemcpyAsync(flags_out, h->d_flags, n_records, cudaMemcpyDeviceToHost, st));
    if (scores_out && n_records) DM_CUDA(cudaMemcpyAsync(scores_out, h->d_scores, (uint64_t)n_records * 4, cudaMemcpyDeviceToHost, st));
    if (masks_out && n_records) DM_CUDA(cudaMemcpyAsync(masks_out, h->d_masks, (uint64_t)n_records * 4, cudaMemcpyDeviceToHost, st));
    DM_CUDA(cudaStreamSynchronize(st));
    // statistics that do not need the device: record and byte counts
    {
        const unsigned long long tr = std::min<uint32_t>(n_train_records, n_records);
        unsigned long long add[6] = {n_records, tr, n_records - tr, 0, 0, record_bytes};
        unsigned long long cur[6];
        DM_CUDA(cudaMemcpy(cur, h->d_stats, sizeof(cur), cudaMemcpyDeviceToHost));
        cur[0] += add[0]; cur[1] += add[1]; cur[2] += add[2]; cur[5] += add[5];
        unsigned long long upd[3] = {cur[0], cur[1], cur[2]};
        DM_CUDA(cudaMemcpy(h->d_stats, upd, sizeof(upd), cudaMemcpyHostToDevice));
        DM_CUDA(cudaMemcpy(h->d_stats + 5, &cur[5], sizeof(unsigned long long), cudaMemcpyHostToDevice));
    }
    h->h_hdr->n_lines = n_records;
    int rc = dm_check_device_errors(h);
    if (rc != DM_OK) return rc;
    if (n_anomalies_out) *n_anomalies_out = h->h_hdr->n_anomalies;
    return DM_OK;
}

// ---------------------------------------------------------------------------------------
// record mode on the device (batches of length-delimited ParserSchema records)
// ---------------------------------------------------------------------------------------
static_assert(sizeof(dm_monitor_t) == sizeof(DmMonitor), "dm_monitor_t and DmMonitor must have the same layout");

extern "C" int dm_set_monitors(dm_handle* h, uint32_t n_monitors, const dm_monitor_t* monitors) {
    if (!h || (n_monitors && !monitors)) return dm_fail(DM_ERR_ARG, "NULL argument");
    if (n_monitors != h->n_keys) return dm_fail(DM_ERR_ARG, "%u monitors given, the handle was created with %u fields", n_monitors, h->n_keys);
    DmMonitors& hm = h->h_mons;
    memset(&hm, 0, sizeof(hm));
    hm.n = n_monitors;
    for (uint32_t i = 0; i < n_monitors; ++i) {
        if (monitors[i].source > 1) return dm_fail(DM_ERR_ARG, "monitor %u: source must be 0 (header) or 1 (variable)", i);
        if (monitors[i].source == 0 && (monitors[i].key_len == 0 || monitors[i].key_len > 64))
            return dm_fail(DM_ERR_ARG, "monitor %u: header key length %u not in 1..64", i, monitors[i].key_len);
        memcpy(&hm.m[i], &monitors[i], sizeof(DmMonitor));
    }
    DM_CUDA(cudaSetDevice(h->device));
    if (!h->d_mons) DM_CUDA(cudaMalloc(&h->d_mons, sizeof(DmMonitors)));
    DM_CUDA(cudaStreamSynchronize(h->last_stream));
    DM_CUDA(cudaMemcpy(h->d_mons, &hm, sizeof(DmMonitors), cudaMemcpyHostToDevice));
    h->mons_set = true;
    return DM_OK;
}

extern "C" int dm_set_combos(dm_handle* h, uint32_t n_combos, const uint32_t* member_off, const uint32_t* members,
                             uint32_t member_only_mask) {
    if (!h) return dm_fail(DM_ERR_ARG, "handle is NULL");
    if (!h->mons_set) return dm_fail(DM_ERR_STATE, "dm_set_monitors has not been called");
    if (n_combos && (!member_off || !members)) return dm_fail(DM_ERR_ARG, "NULL argument");
    DmMonitors& hm = h->h_mons;
    if (hm.n + n_combos > DM_MAX_KEYS)
        return dm_fail(DM_ERR_ARG, "%u monitors + %u combinations exceed the %d output-mask bits", hm.n, n_combos, DM_MAX_KEYS);
    if (n_combos && member_off[0] != 0) return dm_fail(DM_ERR_ARG, "member_off[0] must be 0");
    if (n_combos && member_off[n_combos] > DM_MAX_COMBO_MEMBERS)
        return dm_fail(DM_ERR_ARG, "%u combination members in total, at most %d", member_off[n_combos], DM_MAX_COMBO_MEMBERS);
    for (uint32_t c = 0; c < n_combos; ++c) {
        if (member_off[c + 1] <= member_off[c]) return dm_fail(DM_ERR_ARG, "combination %u is empty or member_off is not increasing", c);
        for (uint32_t j = member_off[c]; j < member_off[c + 1]; ++j)
            if (members[j] >= hm.n) return dm_fail(DM_ERR_ARG, "combination %u: member %u is not a monitor index (< %u)", c, members[j], hm.n);
    }
    if (hm.n < 32 && (member_only_mask >> hm.n)) return dm_fail(DM_ERR_ARG, "member_only_mask has bits beyond the %u monitors", hm.n);
    hm.n_combos = n_combos;
    hm.member_only = member_only_mask;
    memset(hm.combo_off, 0, sizeof(hm.combo_off));
    memset(hm.combo_members, 0, sizeof(hm.combo_members));
    for (uint32_t c = 0; c <= n_combos && n_combos; ++c) hm.combo_off[c] = member_off[c];
    for (uint32_t j = 0; n_combos && j < member_off[n_combos]; ++j) hm.combo_members[j] = (uint8_t)members[j];
    DM_CUDA(cudaSetDevice(h->device));
    DM_CUDA(cudaStreamSynchronize(h->last_stream));
    DM_CUDA(cudaMemcpy(h->d_mons, &hm, sizeof(DmMonitors), cudaMemcpyHostToDevice));
    return DM_OK;
}

// ---------------------------------------------------------------------------------------
// log_format + templates (MatcherParser fused in front of the detector)
// ---------------------------------------------------------------------------------------

extern "C" int dm_set_format(dm_handle* h, const char* log_format, const char* content_name, uint32_t n_templates,
                             const char* const* templates) {
    return dm_set_format_ex(h, log_format, content_name, n_templates, templates, 0);
}

extern "C" int dm_set_format_ex(dm_handle* h, const char* log_format, const char* content_name, uint32_t n_templates,
                                const char* const* templates, uint32_t norm_flags) {
    if (!h) return dm_fail(DM_ERR_ARG, "handle is NULL");
    if (!log_format) {                                             // back to key=value tokenisation
        h->fmt_set = false;
        return DM_OK;
    }
    if (!h->mons_set) return dm_fail(DM_ERR_STATE, "dm_set_monitors has not been called");
    if (n_templates && !templates) return dm_fail(DM_ERR_ARG, "templates is NULL");
    DmFormat* f = new (std::nothrow) DmFormat;
    if (!f) return dm_fail(DM_ERR_CUDA, "out of host memory");
    std::string err;
    int rc = DM_OK;
    if (!dm_format_build(log_format, content_name, n_templates, templates, norm_flags, h->h_mons, f, &err)) {
        rc = dm_fail(DM_ERR_ARG, "%s", err.c_str());
    } else {
        cudaError_t e = cudaSetDevice(h->device);
        if (e == cudaSuccess && !h->d_fmt) e = cudaMalloc(&h->d_fmt, sizeof(DmFormat));
        if (e == cudaSuccess && norm_flags && n_templates && !h->d_norm) e = cudaMalloc(&h->d_norm, h->max_batch_bytes + 256);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->last_stream);
        if (e == cudaSuccess) e = cudaMemcpy(h->d_fmt, f, sizeof(DmFormat), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) rc = dm_fail(DM_ERR_CUDA, "dm_set_format: %s", cudaGetErrorString(e));
        else {
            h->fmt_set = true;
            h->fmt_slots = f->max_slots;
            h->fmt_norm = n_templates ? norm_flags : 0;             // (without templates there is nothing to normalise)
        }
    }
    delete f;
    return rc;
}

extern "C" int dm_stream_recheck_chained(dm_handle* h) {
    if (!h) return dm_fail(DM_ERR_ARG, "handle is NULL");
    return h->dmx.last_chained ? 1 : 0;
}

// diagnostics: per-CTA time stamps of the last stream-kernel launch (DM_STREAM_TIMELINE=1)
extern "C" int dm_debug_rows_timeline(dm_handle* h, unsigned long long* out, uint64_t cap_words, uint32_t* n_warps_out) {
    if (!h || !n_warps_out) return dm_fail(DM_ERR_ARG, "NULL argument");
    if (h->dmx.d_timeline) {
        // stream variant: per CTA {smid, t_start, t_rows_done, t_exit}, then 8 epilogue stamps
        *n_warps_out = h->dmx.last_grid;
        if (!out) return DM_OK;
        DM_CUDA(cudaSetDevice(h->device));
        DM_CUDA(cudaStreamSynchronize(h->last_stream));
        const uint64_t words = std::min<uint64_t>(cap_words, 4ull * h->dmx.last_grid + 8);
        DM_CUDA(cudaMemcpy(out, h->dmx.d_timeline, words * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        return DM_OK;
    }
    return dm_fail(DM_ERR_STATE, "create the handle with DM_STREAM_TIMELINE=1");
}

extern "C" int dm_process_records(dm_handle* h, const uint8_t* buf, uint64_t nbytes, uint32_t n_train_records,
                                  uint8_t* flags_out, float* scores_out, uint32_t* masks_out, uint64_t out_cap,
                                  uint64_t* n_records_out, uint64_t* n_anomalies_out) {
    if (!h) return dm_fail(DM_ERR_ARG, "handle is NULL");
    if (!h->mons_set) return dm_fail(DM_ERR_STATE, "dm_set_monitors has not been called");
    if (nbytes && !buf) return dm_fail(DM_ERR_ARG, "buf is NULL");
    if (nbytes > h->max_batch_bytes) return dm_fail(DM_ERR_CAPACITY, "message of %llu bytes exceeds max_batch_bytes=%llu", (unsigned long long)nbytes, (unsigned long long)h->max_batch_bytes);
    // framing: varint length, payload, varint length, payload ...
    std::vector<uint32_t> off, len;
    uint64_t pos = 0;
    while (pos < nbytes) {
        uint64_t l = 0;
        int shift = 0;
        for (;;) {
            if (pos >= nbytes || shift > 63) return dm_fail(DM_ERR_ARG, "truncated length prefix at byte %llu", (unsigned long long)pos);
            const uint8_t b = buf[pos++];
            l |= (uint64_t)(b & 0x7F) << shift;
            if (!(b & 0x80)) break;
            shift += 7;
        }
        if (l > nbytes - pos) return dm_fail(DM_ERR_ARG, "record of %llu bytes at byte %llu runs past the end of the message", (unsigned long long)l, (unsigned long long)pos);
        off.push_back((uint32_t)pos);
        len.push_back((uint32_t)l);
        pos += l;
    }
    const uint32_t n_records = (uint32_t)off.size();
    if (n_records > h->max_lines) return dm_fail(DM_ERR_CAPACITY, "%u records exceed max_lines=%llu", n_records, (unsigned long long)h->max_lines);
    if ((flags_out || scores_out || masks_out) && n_records > out_cap)
        return dm_fail(DM_ERR_CAPACITY, "batch holds %u records, output capacity is %llu", n_records, (unsigned long long)out_cap);
    DM_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    h->last_stream = st;
    dm_break_chain(h, st);
    if (2ull * n_records > 3ull * h->vals_cap) {
        if (h->d_vals) DM_CUDA(cudaFree(h->d_vals));
        h->vals_cap = std::max<uint64_t>(2ull * n_records, 4096);
        DM_CUDA(cudaMalloc(&h->d_vals, (3 * h->vals_cap + 1) * sizeof(uint32_t)));
    }
    if (!h->d_masks) DM_CUDA(cudaMalloc(&h->d_masks, (h->max_lines + 4) * sizeof(uint32_t)));
    uint32_t* d_off = h->d_vals;
    uint32_t* d_len = h->d_vals + n_records;
    DM_CUDA(cudaMemsetAsync(h->d_hdr, 0, sizeof(DmBatchHeader), st));
    if (n_records) {
        // (the kernel skips malformed records without touching their outputs: they read as "no alert")
        DM_CUDA(cudaMemsetAsync(h->d_flags, 0, n_records, st));
        DM_CUDA(cudaMemsetAsync(h->d_scores, 0, (uint64_t)n_records * 4, st));
        DM_CUDA(cudaMemsetAsync(h->d_masks, 0, (uint64_t)n_records * 4, st));
        DM_CUDA(cudaMemcpyAsync(h->d_in, buf, nbytes, cudaMemcpyHostToDevice, st));
        DM_CUDA(cudaMemcpyAsync(d_off, off.data(), (uint64_t)n_records * 4, cudaMemcpyHostToDevice, st));
        DM_CUDA(cudaMemcpyAsync(d_len, len.data(), (uint64_t)n_records * 4, cudaMemcpyHostToDevice, st));
        DmRecordsArgs a;
        a.buf = h->d_in; a.rec_off = d_off; a.rec_len = d_len; a.n_records = n_records; a.n_train_records = n_train_records;
        a.mons = h->d_mons; a.table = h->table; a.flags = h->d_flags; a.scores = h->d_scores; a.masks = h->d_masks;
        a.hdr = h->d_hdr; a.stats = h->d_stats;
        const int grid = (int)std::min<uint32_t>((n_records + 127) / 128, (uint32_t)h->sm_count * 16);
        if (n_train_records > 0) { dm_k_records<<<grid, 128, 0, st>>>(a, 0); h->launches++; }
        if (n_train_records < n_records) {
            dm_prof_mark(h, st, 0);
            dm_k_records<<<grid, 128, 0, st>>>(a, 1);
            dm_prof_mark(h, st, 1);
            h->launches++;
        }
        DM_CUDA(cudaGetLastError());
    }
    DM_CUDA(cudaMemcpyAsync(h->h_hdr, h->d_hdr, sizeof(DmBatchHeader), cudaMemcpyDeviceToHost, st));
    if (flags_out && n_records) DM_CUDA(cudaMemcpyAsync(flags_out, h->d_flags, n_records, cudaMemcpyDeviceToHost, st));
    if (scores_out && n_records) DM_CUDA(cudaMemcpyAsync(scores_out, h->d_scores, (uint64_t)n_records * 4, cudaMemcpyDeviceToHost, st));
    if (masks_out && n_records) DM_CUDA(cudaMemcpyAsync(masks_out, h->d_masks, (uint64_t)n_records * 4, cudaMemcpyDeviceToHost, st));
    DM_CUDA(cudaStreamSynchronize(st));
    {
        const unsigned long long tr = std::min<uint32_t>(n_train_records, n_records);
        unsigned long long cur[6];
        DM_CUDA(cudaMemcpy(cur, h->d_stats, sizeof(cur), cudaMemcpyDeviceToHost));
        unsigned long long upd[3] = {cur[0] + n_records, cur[1] + tr, cur[2] + (n_records - tr)};
        const unsigned long long nb = cur[5] + nbytes;
        DM_CUDA(cudaMemcpy(h->d_stats, upd, sizeof(upd), cudaMemcpyHostToDevice));
        DM_CUDA(cudaMemcpy(h->d_stats + 5, &nb, sizeof(nb), cudaMemcpyHostToDevice));
    }
    h->h_hdr->n_lines = n_records;
    int rc = dm_check_device_errors(h);
    if (rc != DM_OK) return rc;
    if (n_records_out) *n_records_out = n_records;
    if (n_anomalies_out) *n_anomalies_out = h->h_hdr->n_anomalies;
    return DM_OK;
}

// ---------------------------------------------------------------------------------------
// pipelined host path
// ---------------------------------------------------------------------------------------
static int dm_slots_init(dm_handle* h) {
    if (h->slots_ready) return DM_OK;
    DM_CUDA(cudaStreamCreateWithFlags(&h->st_in, cudaStreamNonBlocking));
    for (auto& sl : h->slots) {
        DM_CUDA(cudaMalloc(&sl.d_in, h->max_batch_bytes + 256));
        DM_CUDA(cudaMemset(sl.d_in, 0, h->max_batch_bytes + 256));
        DM_CUDA(cudaMalloc(&sl.d_flags, h->max_lines + 16));
        DM_CUDA(cudaMalloc(&sl.d_scores, (h->max_lines + 4) * sizeof(float)));
        DM_CUDA(cudaMalloc(&sl.d_hdr, sizeof(DmBatchHeader)));
        DM_CUDA(cudaMemset(sl.d_hdr, 0, sizeof(DmBatchHeader)));
        DM_CUDA(cudaMalloc(&sl.d_anoms, (uint64_t)h->anomaly_cap * sizeof(dm_anomaly_t)));
        DM_CUDA(cudaHostAlloc(&sl.h_hdr, sizeof(DmBatchHeader), cudaHostAllocDefault));
        DM_CUDA(cudaHostAlloc(&sl.h_flags, h->max_lines + 16, cudaHostAllocDefault));
        DM_CUDA(cudaHostAlloc(&sl.h_scores, (h->max_lines + 4) * sizeof(float), cudaHostAllocDefault));
        DM_CUDA(cudaEventCreateWithFlags(&sl.ev_in, cudaEventDisableTiming));
        DM_CUDA(cudaEventCreateWithFlags(&sl.ev_comp, cudaEventDisableTiming));
        DM_CUDA(cudaEventCreateWithFlags(&sl.ev_hdr, cudaEventDisableTiming));
        DM_CUDA(cudaStreamCreateWithFlags(&sl.st_out, cudaStreamNonBlocking));
    }
    h->slots_ready = true;
    return DM_OK;
}

extern "C" int dm_submit_lines(dm_handle* h, const uint8_t* host_buf, uint64_t nbytes, uint64_t n_train_lines, uint32_t slot) {
    if (!h) return dm_fail(DM_ERR_ARG, "handle is NULL");
    if (slot > 1) return dm_fail(DM_ERR_ARG, "slot must be 0 or 1");
    if (nbytes > h->max_batch_bytes)
        return dm_fail(DM_ERR_CAPACITY, "message of %llu bytes exceeds max_batch_bytes=%llu", (unsigned long long)nbytes, (unsigned long long)h->max_batch_bytes);
    if (nbytes && !host_buf) return dm_fail(DM_ERR_ARG, "host_buf is NULL");
    if (h->kernel_variant != 6) return dm_fail(DM_ERR_STATE, "the pipelined path runs the stream kernel (DM_KERNEL=lanes is set)");
    if (h->fmt_set) return dm_fail(DM_ERR_STATE, "the pipelined path tokenises key=value records; with a log_format use dm_process_lines");
    if (h->mons_set && h->h_mons.n_combos > 0) return dm_fail(DM_ERR_STATE, "combination monitors run in dm_process_lines / dm_process_records, not in the pipelined path");
    DM_CUDA(cudaSetDevice(h->device));
    int rc = dm_slots_init(h);
    if (rc != DM_OK) return rc;
    dm_handle::Slot& sl = h->slots[slot];
    if (sl.busy) return dm_fail(DM_ERR_STATE, "slot %u was submitted and not collected yet", slot);
    if (nbytes) DM_CUDA(cudaMemcpyAsync(sl.d_in, host_buf, nbytes, cudaMemcpyHostToDevice, h->st_in));
    DM_CUDA(cudaMemsetAsync(sl.d_in + nbytes, 0, 64, h->st_in));
    DM_CUDA(cudaEventRecord(sl.ev_in, h->st_in));
    cudaStream_t st = h->stream;                       // ONE compute stream: submission order = processing order
    h->last_stream = st;
    DM_CUDA(cudaStreamWaitEvent(st, sl.ev_in, 0));
    if (nbytes == 0) DM_CUDA(cudaMemsetAsync(sl.d_hdr, 0, sizeof(DmBatchHeader), st));
    const int launched = dmx_launch(&h->dmx, sl.d_in, nbytes, n_train_lines, h->table, sl.d_flags, sl.d_scores, h->max_lines, sl.d_anoms,
                                    h->anomaly_cap, sl.d_hdr, h->d_stats, h->max_lines, st, true, dm_prof_mark_cb, h);
    if (launched < 0) return dm_fail(DM_ERR_CUDA, "stream kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    h->launches += (uint64_t)launched;
    DM_CUDA(cudaEventRecord(sl.ev_comp, st));
    DM_CUDA(cudaStreamWaitEvent(sl.st_out, sl.ev_comp, 0));
    DM_CUDA(cudaMemcpyAsync(sl.h_hdr, sl.d_hdr, sizeof(DmBatchHeader), cudaMemcpyDeviceToHost, sl.st_out));
    DM_CUDA(cudaEventRecord(sl.ev_hdr, sl.st_out));
    sl.busy = true;
    return DM_OK;
}

extern "C" int dm_collect(dm_handle* h, uint32_t slot, const uint8_t** flags_out, const float** scores_out,
                          uint64_t* n_lines_out, uint64_t* n_anomalies_out) {
    if (!h) return dm_fail(DM_ERR_ARG, "handle is NULL");
    if (slot > 1 || !h->slots_ready || !h->slots[slot].busy) return dm_fail(DM_ERR_STATE, "slot %u has nothing to collect", slot);
    DM_CUDA(cudaSetDevice(h->device));
    dm_handle::Slot& sl = h->slots[slot];
    DM_CUDA(cudaEventSynchronize(sl.ev_hdr));
    sl.busy = false;
    *h->h_hdr = *sl.h_hdr;
    int rc = dm_check_device_errors(h);
    if (rc != DM_OK) return rc;
    const uint64_t n = sl.h_hdr->n_lines;
    if (n && (flags_out || scores_out)) {
        if (flags_out) DM_CUDA(cudaMemcpyAsync(sl.h_flags, sl.d_flags, n, cudaMemcpyDeviceToHost, sl.st_out));
        if (scores_out) DM_CUDA(cudaMemcpyAsync(sl.h_scores, sl.d_scores, n * sizeof(float), cudaMemcpyDeviceToHost, sl.st_out));
        DM_CUDA(cudaStreamSynchronize(sl.st_out));
    }
    if (flags_out) *flags_out = sl.h_flags;
    if (scores_out) *scores_out = sl.h_scores;
    if (n_lines_out) *n_lines_out = n;
    if (n_anomalies_out) *n_anomalies_out = sl.h_hdr->n_anomalies;
    return DM_OK;
}

static int dm_fetch_anomalies(dm_handle* h, const dm_anomaly_t* d_list, uint32_t total, dm_anomaly_t* out, uint32_t cap, uint32_t* n_out) {
    const uint32_t have = std::min(total, h->anomaly_cap);
    *n_out = 0;
    if (have == 0) return DM_OK;
    std::vector<dm_anomaly_t> tmp(have);
    DM_CUDA(cudaMemcpy(tmp.data(), d_list, (uint64_t)have * sizeof(dm_anomaly_t), cudaMemcpyDeviceToHost));
    std::sort(tmp.begin(), tmp.end(), [](const dm_anomaly_t& a, const dm_anomaly_t& b) { return a.line < b.line; });
    uint32_t m = 0;
    for (uint32_t i = 0; i < have; ++i) {
        if (m && tmp[m - 1].line == tmp[i].line) { tmp[m - 1].mask |= tmp[i].mask; continue; }
        tmp[m++] = tmp[i];
    }
    *n_out = m;
    if (out && cap) memcpy(out, tmp.data(), (uint64_t)std::min(m, cap) * sizeof(dm_anomaly_t));
    return DM_OK;
}

extern "C" int dm_collect_anomalies(dm_handle* h, uint32_t slot, dm_anomaly_t* out, uint32_t cap, uint32_t* n_out) {
    if (!h || !n_out) return dm_fail(DM_ERR_ARG, "NULL argument");
    if (slot > 1 || !h->slots_ready) return dm_fail(DM_ERR_STATE, "slot %u was never used", slot);
    DM_CUDA(cudaSetDevice(h->device));
    dm_handle::Slot& sl = h->slots[slot];
    if (sl.busy) return dm_fail(DM_ERR_STATE, "slot %u is still in flight: dm_collect it first", slot);
    return dm_fetch_anomalies(h, sl.d_anoms, sl.h_hdr->anomaly_list_count, out, cap, n_out);
}

extern "C" int dm_sync(dm_handle* h, uint64_t* n_lines_out, uint64_t* n_anomalies_out) {
    if (!h) return dm_fail(DM_ERR_ARG, "handle is NULL");
    DM_CUDA(cudaSetDevice(h->device));
    DM_CUDA(cudaMemcpyAsync(h->h_hdr, h->d_hdr, sizeof(DmBatchHeader), cudaMemcpyDeviceToHost, h->last_stream));
    DM_CUDA(cudaStreamSynchronize(h->last_stream));
    int rc = dm_check_device_errors(h);
    if (rc != DM_OK) return rc;
    if (n_lines_out) *n_lines_out = h->h_hdr->n_lines;
    if (n_anomalies_out) *n_anomalies_out = h->h_hdr->n_anomalies;
    return DM_OK;
}

extern "C" int dm_get_anomalies(dm_handle* h, dm_anomaly_t* out, uint32_t cap, uint32_t* n_out) {
    if (!h || !n_out) return dm_fail(DM_ERR_ARG, "NULL argument");
    DM_CUDA(cudaSetDevice(h->device));
    int rc = dm_sync(h, nullptr, nullptr);
    if (rc != DM_OK) return rc;
    const uint32_t total = h->h_hdr->anomaly_list_count;
    *n_out = std::min<uint64_t>(total, h->h_hdr->n_anomalies ? h->h_hdr->n_anomalies : total);
    const uint32_t have = std::min(total, h->anomaly_cap);
    if (!out || cap == 0 || have == 0) return DM_OK;
    std::vector<dm_anomaly_t> tmp(have);
    DM_CUDA(cudaMemcpy(tmp.data(), h->d_anoms, (uint64_t)have * sizeof(dm_anomaly_t), cudaMemcpyDeviceToHost));
    std::sort(tmp.begin(), tmp.end(), [](const dm_anomaly_t& a, const dm_anomaly_t& b) { return a.line < b.line; });
    // the fused kernel appends one entry per unknown field: merge the entries of a record
    uint32_t m = 0;
    for (uint32_t i = 0; i < have; ++i) {
        if (m && tmp[m - 1].line == tmp[i].line) { tmp[m - 1].mask |= tmp[i].mask; continue; }
        tmp[m++] = tmp[i];
    }
    *n_out = m;
    memcpy(out, tmp.data(), (uint64_t)std::min(m, cap) * sizeof(dm_anomaly_t));
    return DM_OK;
}

static void dm_words_to_stats(const unsigned long long* w, dm_stats_t* out) {
    out->lines = w[0]; out->train_lines = w[1]; out->detect_lines = w[2]; out->anomalies = w[3];
    out->score_sum = w[4]; out->bytes = w[5]; out->known_keys = w[6]; out->bad_records = w[7];
    for (int k = 0; k < DM_MAX_KEYS; ++k) out->unknown_per_key[k] = w[8 + k];
}

extern "C" int dm_get_stats(dm_handle* h, dm_stats_t* out) {
    if (!h || !out) return dm_fail(DM_ERR_ARG, "NULL argument");
    DM_CUDA(cudaSetDevice(h->device));
    DM_CUDA(cudaStreamSynchronize(h->last_stream));
    DM_CUDA(cudaMemcpy(h->h_stats, h->d_stats, DM_STATS_WORDS * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    dm_words_to_stats(h->h_stats, out);
    unsigned long long cnt = 0;
    DM_CUDA(cudaMemcpy(&cnt, h->table.count, sizeof(cnt), cudaMemcpyDeviceToHost));
    out->known_keys = cnt;
    return DM_OK;
}

extern "C" int dm_get_global_stats(dm_handle* h, dm_stats_t* out) {
    if (!h || !out) return dm_fail(DM_ERR_ARG, "NULL argument");
    DM_CUDA(cudaSetDevice(h->device));
    DM_CUDA(cudaStreamSynchronize(h->last_stream));
    DM_CUDA(cudaMemcpy(h->h_stats, h->d_stats_global, DM_STATS_WORDS * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    dm_words_to_stats(h->h_stats, out);
    return DM_OK;
}

// ---------------------------------------------------------------------------------------
// known-set export / import / reset
// ---------------------------------------------------------------------------------------
__global__ void dm_k_insert_keys(DmTable t, const unsigned long long* __restrict__ keys, uint64_t n, unsigned int* err) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        unsigned long long k = keys[i];
        if (k) dm_table_insert(t, k, err, false);
    }
}

extern "C" int dm_export_known(dm_handle* h, uint64_t* keys_out, uint64_t cap, uint64_t* n_out) {
    if (!h || !n_out) return dm_fail(DM_ERR_ARG, "NULL argument");
    DM_CUDA(cudaSetDevice(h->device));
    DM_CUDA(cudaStreamSynchronize(h->last_stream));
    // every key of the table (own records, peers, dm_import_known): the slots themselves are read back
    const uint64_t slots = (uint64_t)h->table.mask + 1;
    std::vector<unsigned long long> tmp(slots);
    DM_CUDA(cudaMemcpy(tmp.data(), h->table.slots, slots * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    uint64_t n = 0;
    for (uint64_t i = 0; i < slots; ++i)
        if (tmp[i]) tmp[n++] = tmp[i];
    std::sort(tmp.begin(), tmp.begin() + n);
    *n_out = n;
    if (keys_out && cap) memcpy(keys_out, tmp.data(), std::min<uint64_t>(n, cap) * sizeof(uint64_t));
    return DM_OK;
}

extern "C" int dm_import_known(dm_handle* h, const uint64_t* keys, uint64_t n) {
    if (!h || (n && !keys)) return dm_fail(DM_ERR_ARG, "NULL argument");
    if (n == 0) return DM_OK;
    DM_CUDA(cudaSetDevice(h->device));
    unsigned long long* d_tmp = nullptr;
    DM_CUDA(cudaMalloc(&d_tmp, n * sizeof(unsigned long long)));
    DM_CUDA(cudaMemcpy(d_tmp, keys, n * sizeof(unsigned long long), cudaMemcpyHostToDevice));
    DM_CUDA(cudaMemsetAsync(h->d_hdr, 0, sizeof(DmBatchHeader), h->last_stream));
    dm_break_chain(h, h->last_stream);
    dm_k_insert_keys<<<std::max(1, (int)std::min<uint64_t>((n + 255) / 256, 1024)), 256, 0, h->last_stream>>>(h->table, d_tmp, n, &h->d_hdr->error);
    DM_CUDA(cudaGetLastError());
    int rc = dm_sync(h, nullptr, nullptr);
    cudaFree(d_tmp);
    return rc;
}

extern "C" int dm_reset(dm_handle* h) {
    if (!h) return dm_fail(DM_ERR_ARG, "handle is NULL");
    DM_CUDA(cudaSetDevice(h->device));
    DM_CUDA(cudaStreamSynchronize(h->last_stream));
    DM_CUDA(cudaMemset(h->table.slots, 0, ((uint64_t)h->table.mask + 1) * sizeof(unsigned long long)));
    DM_CUDA(cudaMemset(h->table.count, 0, 2 * sizeof(unsigned long long)));
    DM_CUDA(cudaMemset(h->d_stats, 0, 3 * DM_STATS_WORDS * sizeof(unsigned long long)));
    DM_CUDA(cudaMemset(h->d_hdr, 0, sizeof(DmBatchHeader)));
    h->novel_exported = 0;
    h->novel_pending = 0;
    return DM_OK;
}

// ---------------------------------------------------------------------------------------
// multi-GPU window exchange
//   buffer layout (uint64 words): [0, DM_STATS_WORDS) statistics delta, summed over ranks;
//   with_keys: then per rank r a segment [count, key_0 .. key_{DM_WINDOW_KEYS-1}] that only
//   rank r fills, so the SUM all-reduce concatenates the segments.
// ---------------------------------------------------------------------------------------
extern "C" uint64_t dm_window_words(dm_handle* h, uint32_t world, int with_keys) {
    (void)h;
    return (uint64_t)DM_STATS_WORDS + (with_keys ? (uint64_t)world * (1ull + DM_WINDOW_KEYS) : 0ull);
}

__global__ void dm_k_window_export(unsigned long long* __restrict__ out, uint64_t n_words,
                                   unsigned long long* __restrict__ stats, unsigned long long* __restrict__ exported,
                                   const unsigned long long* __restrict__ table_count,
                                   const unsigned long long* __restrict__ novel, uint64_t novel_from, uint64_t novel_to,
                                   uint64_t seg_off, int with_keys) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t nth = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = tid; i < n_words; i += nth) {
        unsigned long long v = 0;
        if (i < DM_STATS_WORDS) {
            unsigned long long cur = (i == 6) ? 0ull : stats[i];
            v = cur - atomicExch(exported + i, cur);      // telescopes mod 2^64 whatever the order of in-flight windows
        } else if (with_keys && i >= seg_off && i < seg_off + 1 + DM_WINDOW_KEYS) {
            const uint64_t j = i - seg_off;
            if (j == 0) v = novel_to - novel_from;
            else if (novel_from + j - 1 < novel_to) v = novel[novel_from + j - 1];
        }
        out[i] = v;
    }
    (void)table_count;
}

__global__ void dm_k_window_import(const unsigned long long* __restrict__ in, uint32_t world, uint32_t self_rank,
                                   unsigned long long* __restrict__ global_stats, DmTable t, unsigned int* err,
                                   int with_keys) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t nth = (uint64_t)gridDim.x * blockDim.x;
    if (tid < DM_STATS_WORDS) atomicAdd(global_stats + tid, in[tid]);     // (two windows may be in flight)
    if (!with_keys) return;
    for (uint32_t r = 0; r < world; ++r) {
        if (r == self_rank) continue;            // own keys are already in the local table
        const uint64_t seg = (uint64_t)DM_STATS_WORDS + (uint64_t)r * (1ull + DM_WINDOW_KEYS);
        const uint64_t cnt = in[seg];
        for (uint64_t j = tid; j < cnt && j < DM_WINDOW_KEYS; j += nth) {
            unsigned long long k = in[seg + 1 + j];
            if (k) dm_table_insert(t, k, err, false);
        }
    }
}

extern "C" int dm_window_export(dm_handle* h, uint64_t* dev_buf, uint32_t rank, uint32_t world, int with_keys, void* stream_) {
    if (!h || !dev_buf || rank >= world) return dm_fail(DM_ERR_ARG, "bad window arguments");
    DM_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = stream_ ? (cudaStream_t)stream_ : h->stream;     // (not recorded as the compute stream)
    dm_break_chain(h, st);
    uint64_t novel_to = h->novel_exported;
    if (with_keys) {
        // the number of keys learnt so far is needed on the host to bound the segment
        unsigned long long cnt[2];
        DM_CUDA(cudaMemcpyAsync(cnt, h->table.count, sizeof(cnt), cudaMemcpyDeviceToHost, st));
        DM_CUDA(cudaStreamSynchronize(st));
        novel_to = std::min<uint64_t>(cnt[1], h->table.novel_cap);
        // at most DM_WINDOW_KEYS keys travel per window; the rest goes with the next one (dm_window_pending_keys).
        // Failing here instead would leave the other ranks waiting in the all-reduce.
        if (novel_to - h->novel_exported > DM_WINDOW_KEYS) novel_to = h->novel_exported + DM_WINDOW_KEYS;
        h->novel_pending = std::min<uint64_t>(cnt[1], h->table.novel_cap) - novel_to;
    }
    const uint64_t n_words = dm_window_words(h, world, with_keys);
    const uint64_t seg_off = (uint64_t)DM_STATS_WORDS + (uint64_t)rank * (1ull + DM_WINDOW_KEYS);
    dm_k_window_export<<<std::max(1, (int)std::min<uint64_t>((n_words + 255) / 256, 2048)), 256, 0, st>>>(
        (unsigned long long*)dev_buf, n_words, h->d_stats, h->d_stats_exported, h->table.count, h->table.novel,
        h->novel_exported, novel_to, seg_off, with_keys);
    DM_CUDA(cudaGetLastError());
    h->novel_exported = novel_to;
    return DM_OK;
}

extern "C" int dm_window_pending_keys(dm_handle* h, uint64_t* n_out) {
    if (!h || !n_out) return dm_fail(DM_ERR_ARG, "NULL argument");
    *n_out = h->novel_pending;
    return DM_OK;
}

extern "C" int dm_window_import(dm_handle* h, const uint64_t* dev_buf, uint32_t rank, uint32_t world, int with_keys, void* stream_) {
    if (!h || !dev_buf || rank >= world) return dm_fail(DM_ERR_ARG, "bad window arguments");
    DM_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = stream_ ? (cudaStream_t)stream_ : h->stream;     // (not recorded as the compute stream)
    dm_break_chain(h, st);
    dm_k_window_import<<<with_keys ? 256 : 1, 256, 0, st>>>((const unsigned long long*)dev_buf, world, rank,
                                                          h->d_stats_global, h->table, &h->d_hdr->error, with_keys);
    DM_CUDA(cudaGetLastError());
    // (keys received from peers are inserted without joining the novel list: they are not shipped again)
    return DM_OK;
}

// ---------------------------------------------------------------------------------------
// The window exchange done natively: export -> ncclAllReduce(sum, uint64) -> import, all
// enqueued by ONE call (the torch.distributed route costs ~30 us of host time per window,
// which is what bounds the multi-GPU step).  NCCL is resolved at run time (dlopen of the
// libnccl.so.2 the process already has -- PyTorch's -- so there is no link-time dependency).
// ---------------------------------------------------------------------------------------
#include <dlfcn.h>
namespace {
struct DmNcclUniqueId { char internal[128]; };
typedef int (*fn_ncclGetUniqueId)(DmNcclUniqueId*);
typedef int (*fn_ncclCommInitRank)(void**, int, DmNcclUniqueId, int);
typedef int (*fn_ncclAllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_ncclCommDestroy)(void*);
typedef const char* (*fn_ncclGetErrorString)(int);
struct DmNccl {
    void* lib = nullptr;
    fn_ncclGetUniqueId get_id = nullptr;
    fn_ncclCommInitRank init_rank = nullptr;
    fn_ncclAllReduce all_reduce = nullptr;
    fn_ncclCommDestroy destroy = nullptr;
    fn_ncclGetErrorString err = nullptr;
};
DmNccl g_nccl;
const int DM_NCCL_UINT64 = 5, DM_NCCL_SUM = 0;      // ncclUint64, ncclSum (stable across NCCL 2.x)

int dm_nccl_load() {
    if (g_nccl.lib) return DM_OK;
    const char* names[] = {getenv("DM_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* n : names) {
        if (!n) continue;
        lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) return dm_fail(DM_ERR_STATE, "libnccl.so.2 not found (set DM_NCCL_LIB): %s", dlerror());
    g_nccl.get_id = (fn_ncclGetUniqueId)dlsym(lib, "ncclGetUniqueId");
    g_nccl.init_rank = (fn_ncclCommInitRank)dlsym(lib, "ncclCommInitRank");
    g_nccl.all_reduce = (fn_ncclAllReduce)dlsym(lib, "ncclAllReduce");
    g_nccl.destroy = (fn_ncclCommDestroy)dlsym(lib, "ncclCommDestroy");
    g_nccl.err = (fn_ncclGetErrorString)dlsym(lib, "ncclGetErrorString");
    if (!g_nccl.get_id || !g_nccl.init_rank || !g_nccl.all_reduce || !g_nccl.destroy || !g_nccl.err)
        return dm_fail(DM_ERR_STATE, "libnccl lacks an expected symbol");
    g_nccl.lib = lib;
    g_nccl_destroy_hook = [](void* c) { g_nccl.destroy(c); };
    return DM_OK;
}
}  // namespace

extern "C" int dm_nccl_unique_id(uint8_t* out128) {
    if (!out128) return dm_fail(DM_ERR_ARG, "NULL argument");
    int rc = dm_nccl_load();
    if (rc != DM_OK) return rc;
    DmNcclUniqueId id;
    const int r = g_nccl.get_id(&id);
    if (r != 0) return dm_fail(DM_ERR_CUDA, "ncclGetUniqueId: %s", g_nccl.err(r));
    memcpy(out128, &id, sizeof(id));
    return DM_OK;
}

extern "C" int dm_nccl_init(dm_handle* h, const uint8_t* ids, uint32_t n_comms, uint32_t rank, uint32_t world) {
    if (!h || !ids || world == 0 || rank >= world || n_comms < 1 || n_comms > 2) return dm_fail(DM_ERR_ARG, "bad arguments");
    if (h->nccl_n) return dm_fail(DM_ERR_STATE, "the handle already has its communicators");
    int rc = dm_nccl_load();
    if (rc != DM_OK) return rc;
    DM_CUDA(cudaSetDevice(h->device));
    for (uint32_t k = 0; k < n_comms; ++k) {
        DmNcclUniqueId id;
        memcpy(&id, ids + 128 * k, sizeof(id));
        void* comm = nullptr;
        const int r = g_nccl.init_rank(&comm, (int)world, id, (int)rank);
        if (r != 0) return dm_fail(DM_ERR_CUDA, "ncclCommInitRank: %s", g_nccl.err(r));
        h->nccl_comm[k] = comm;
        DM_CUDA(cudaMalloc(&h->d_win[k], dm_window_words(h, world, 1) * sizeof(unsigned long long)));
        DM_CUDA(cudaEventCreateWithFlags(&h->ev_win_done[k], cudaEventDisableTiming));
    }
    h->nccl_n = n_comms; h->nccl_rank = rank; h->nccl_world = world;
    return DM_OK;
}

extern "C" int dm_window_allreduce(dm_handle* h, int with_keys, void* stream_) {
    if (!h) return dm_fail(DM_ERR_ARG, "handle is NULL");
    if (!h->nccl_n) return dm_fail(DM_ERR_STATE, "dm_nccl_init has not been called");
    // consecutive windows alternate between the communicators (and their buffers): called on
    // alternating streams, two all-reduces are in flight, which matters at 8 ranks where one
    // export + all-reduce + import chain is longer than a step
    const uint32_t k = (uint32_t)(h->win_seq++ % h->nccl_n);
    cudaStream_t st = stream_ ? (cudaStream_t)stream_ : h->stream;     // (not recorded as the compute stream)
    DM_CUDA(cudaSetDevice(h->device));
    DM_CUDA(cudaStreamWaitEvent(st, h->ev_win_done[k], 0));           // the buffer's previous window is through
    int rc = dm_window_export(h, (uint64_t*)h->d_win[k], h->nccl_rank, h->nccl_world, with_keys, stream_);
    if (rc != DM_OK) return rc;
    const int r = g_nccl.all_reduce(h->d_win[k], h->d_win[k], (size_t)dm_window_words(h, h->nccl_world, with_keys),
                                    DM_NCCL_UINT64, DM_NCCL_SUM, h->nccl_comm[k], st);
    if (r != 0) return dm_fail(DM_ERR_CUDA, "ncclAllReduce: %s", g_nccl.err(r));
    rc = dm_window_import(h, (const uint64_t*)h->d_win[k], h->nccl_rank, h->nccl_world, with_keys, stream_);
    if (rc != DM_OK) return rc;
    DM_CUDA(cudaEventRecord(h->ev_win_done[k], st));
    return DM_OK;
}

// ---------------------------------------------------------------------------------------
// Host utility: write a host buffer back and evict it from the CPU caches.  A pinned buffer
// the CPU has just written sits in its caches as dirty lines, and the GPU's DMA reads of such
// lines are snooped out of the caches at a fraction of the DRAM streaming rate (measured:
// 16 MiB in 500-1500 us instead of 313 us, scripts/pinned_numa_diag.py).  bench.py flushes its
// pinned inputs once after filling them so that the end-to-end number does not depend on
// what the set-up code left in the L3.
// ---------------------------------------------------------------------------------------
#if defined(__x86_64__)
#include <emmintrin.h>
extern "C" int dm_host_cache_flush(const void* p, uint64_t nbytes) {
    if (!p && nbytes) return dm_fail(DM_ERR_ARG, "NULL argument");
    const char* c = (const char*)((uintptr_t)p & ~(uintptr_t)63);
    const char* e = (const char*)p + nbytes;
    _mm_mfence();
    for (; c < e; c += 64) _mm_clflush(c);
    _mm_mfence();
    return DM_OK;
}
#else
extern "C" int dm_host_cache_flush(const void* p, uint64_t nbytes) { (void)p; (void)nbytes; return DM_OK; }
#endif
