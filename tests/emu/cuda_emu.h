// cuda_emu.h -- a tiny CUDA-on-CPU shim, TEST INFRASTRUCTURE ONLY.
//
// Lets tests/ compile the *same* kernel source as the product (dm_kernels_tile.cuh) with
// g++ and run ONE thread block on OS threads, so the tokenizer/detector logic can be
// checked against the oracle in the CPU test tier (there is no GPU in the build
// container).  One OS thread per CUDA thread; warp collectives and block barriers are
// real barriers, so every lane of a warp must reach a collective (the kernels only use
// full-mask collectives at convergent points).  Nothing here is linked into
// libdmdetect.so and nothing in the product calls it.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#define DM_EMU 1
#define DMX_KEYTAB_LOOP 1      // (no bulk-copy engine in the emulator)
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define DM_NOINLINE __attribute__((noinline))
#define __shared__ static
#define __restrict__
#define __launch_bounds__(...)

struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };
struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
static inline uint2 make_uint2(uint32_t a, uint32_t b) { return uint2{a, b}; }
extern std::vector<uint8_t> g_emu_dyn_smem;           // stands in for `extern __shared__` (blocks run one at a time)
static inline uint4 make_uint4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return uint4{a, b, c, d}; }

extern thread_local emu_dim3 threadIdx;
extern thread_local emu_dim3 blockIdx;
extern emu_dim3 blockDim, gridDim;

// ---- barriers -------------------------------------------------------------------------
struct EmuBarrier {
    std::mutex m; std::condition_variable cv; unsigned arrived = 0; unsigned gen = 0;
    void sync(unsigned count) {
        std::unique_lock<std::mutex> lk(m);
        unsigned g = gen;
        if (++arrived >= count) { arrived = 0; ++gen; cv.notify_all(); return; }
        cv.wait(lk, [&] { return gen != g; });
    }
    void arrive(unsigned count) {
        std::unique_lock<std::mutex> lk(m);
        if (++arrived >= count) { arrived = 0; ++gen; cv.notify_all(); }
    }
};
struct EmuWarp {
    EmuBarrier bar;
    uint64_t slot[32];
};
struct EmuBlock {
    EmuBarrier block_bar;
    EmuBarrier named[16];
    std::vector<EmuWarp> warps;
};
extern EmuBlock* g_emu_block;

static inline EmuWarp& emu_warp() { return g_emu_block->warps[threadIdx.x >> 5]; }
static inline unsigned emu_lane() { return threadIdx.x & 31; }

static inline void __syncthreads() { g_emu_block->block_bar.sync(blockDim.x); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu_warp().bar.sync(32); }
extern std::atomic<int> g_emu_or_flag[2];
static inline int __syncthreads_or(int pred) {
    // two flags used alternately so that a fast thread's next call cannot clobber this one
    static thread_local int phase = 0;
    const int p = phase;
    phase ^= 1;
    if (pred) g_emu_or_flag[p].store(1);
    g_emu_block->block_bar.sync(blockDim.x);
    const int r = g_emu_or_flag[p].load();
    g_emu_block->block_bar.sync(blockDim.x);
    if (threadIdx.x == 0) g_emu_or_flag[p].store(0);
    g_emu_block->block_bar.sync(blockDim.x);
    return r;
}
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __nanosleep(unsigned) { std::this_thread::yield(); }
// named barriers: counts are in threads, as in PTX bar.sync / bar.arrive
static inline void emu_bar_sync(int id, unsigned count) { g_emu_block->named[id].sync(count); }
static inline void emu_bar_arrive(int id, unsigned count) { g_emu_block->named[id].arrive(count); }

// ---- warp collectives --------------------------------------------------------------------
template <typename T> static inline T emu_exchange(T v, unsigned src) {
    EmuWarp& w = emu_warp();
    uint64_t raw = 0; std::memcpy(&raw, &v, sizeof(T));
    w.slot[emu_lane()] = raw;
    w.bar.sync(32);
    uint64_t got = w.slot[src & 31];
    w.bar.sync(32);
    T out; std::memcpy(&out, &got, sizeof(T));
    return out;
}
template <typename T> static inline T __shfl_sync(unsigned, T v, int src) { return emu_exchange(v, (unsigned)src); }
template <typename T> static inline T __shfl_up_sync(unsigned, T v, unsigned d) {
    unsigned l = emu_lane();
    T got = emu_exchange(v, l >= d ? l - d : l);
    return l >= d ? got : v;
}
template <typename T> static inline T __shfl_down_sync(unsigned, T v, unsigned d) {
    unsigned l = emu_lane();
    T got = emu_exchange(v, l + d < 32 ? l + d : l);
    return l + d < 32 ? got : v;
}
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return emu_exchange(v, emu_lane() ^ (unsigned)m); }
static inline unsigned __ballot_sync(unsigned, int pred) {
    EmuWarp& w = emu_warp();
    w.slot[emu_lane()] = pred ? 1 : 0;
    w.bar.sync(32);
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= (unsigned)(w.slot[i] & 1) << i;
    w.bar.sync(32);
    return r;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }
static inline unsigned emu_reduce(unsigned v, int op) {
    EmuWarp& w = emu_warp();
    w.slot[emu_lane()] = v;
    w.bar.sync(32);
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r = op ? (r | (unsigned)w.slot[i]) : (r + (unsigned)w.slot[i]);
    w.bar.sync(32);
    return r;
}
static inline unsigned __reduce_add_sync(unsigned, unsigned v) { return emu_reduce(v, 0); }
static inline unsigned __reduce_min_sync(unsigned, unsigned v) {
    EmuWarp& w = emu_warp();
    w.slot[emu_lane()] = v;
    w.bar.sync(32);
    unsigned r = 0xFFFFFFFFu;
    for (int i = 0; i < 32; ++i) r = (unsigned)w.slot[i] < r ? (unsigned)w.slot[i] : r;
    w.bar.sync(32);
    return r;
}
static inline unsigned __reduce_or_sync(unsigned, unsigned v) { return emu_reduce(v, 1); }
static inline unsigned __reduce_xor_sync(unsigned, unsigned v) {
    EmuWarp& w = emu_warp();
    w.slot[emu_lane()] = v;
    w.bar.sync(32);
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r ^= (unsigned)w.slot[i];
    w.bar.sync(32);
    return r;
}
static inline unsigned __reduce_max_sync(unsigned, unsigned v) {
    EmuWarp& w = emu_warp();
    w.slot[emu_lane()] = v;
    w.bar.sync(32);
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r = (unsigned)w.slot[i] > r ? (unsigned)w.slot[i] : r;
    w.bar.sync(32);
    return r;
}

// ---- scalar intrinsics -------------------------------------------------------------------
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
static inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
    uint64_t v = ((uint64_t)hi << 32) | lo;
    return (unsigned)(v >> (sh & 31));
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh) {
    uint64_t v = ((uint64_t)hi << 32) | lo;
    return (unsigned)((v << (sh & 31)) >> 32);
}
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T __ldcg(const T* p) { return *p; }
static inline uint4 __ldg(const uint4* p) { uint4 r; std::memcpy(&r, p, 16); return r; }

// ---- atomics (GCC builtins on plain memory) ----------------------------------------------
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline float atomicAdd(float* p, float v) {
    uint32_t* ip = reinterpret_cast<uint32_t*>(p);
    uint32_t old = __atomic_load_n(ip, __ATOMIC_SEQ_CST);
    for (;;) {
        float f; std::memcpy(&f, &old, 4);
        float nf = f + v; uint32_t ni; std::memcpy(&ni, &nf, 4);
        if (__atomic_compare_exchange_n(ip, &old, ni, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) return f;
    }
}
static inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicCAS(unsigned long long* p, unsigned long long cmp, unsigned long long val) {
    __atomic_compare_exchange_n(p, &cmp, val, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return cmp;
}
static inline unsigned long long atomicExch(unsigned long long* p, unsigned long long v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }

// ---- launch ------------------------------------------------------------------------------
// Runs ONE block of `threads` threads (grid = 1).
void emu_launch(unsigned threads, const std::function<void()>& kernel_body);
