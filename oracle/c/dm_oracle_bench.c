/*
 * dm_oracle_bench.c -- threaded throughput driver of the C restatement (dm_oracle.c).
 * TEST / MEASUREMENT INFRASTRUCTURE ONLY: bench.py's CPU legs (`cpu_baseline`, `--impl reference`).
 *
 * The reference runs ONE Python thread per service (/root/reference/src/service/features/
 * engine.py:80-82); a box is used by running one service per core.  This driver is the strong
 * version of that: one pinned worker thread per core, each with its own trained detector (the
 * reference's services share nothing either), each scanning its own shard of the same synthetic
 * messages, in C.  Timing is done here, not in Python: all workers start together behind a
 * barrier and scan their shard over and over until the main thread calls time after
 * `min_seconds`; a sample's rate = records scanned by all workers / wall time.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

typedef struct dmo dmo;
dmo* dmo_create(int n_keys, const uint8_t* keys_blob, const uint32_t* key_lens);
void dmo_destroy(dmo* d);
uint64_t dmo_process(dmo* d, const uint8_t* buf, uint64_t n, uint64_t n_train_lines, uint8_t* flags, float* scores, uint32_t* masks);

typedef struct {
    /* what the worker needs to build ITS detector on ITS core (first touch: the state lands on the worker's NUMA node) */
    int n_keys;
    const uint8_t* keys_blob;
    const uint32_t* key_lens;
    const uint8_t* train;
    uint64_t train_bytes;
    dmo* det;
    const uint8_t* shard;
    uint64_t shard_bytes;
    uint8_t* flags;
    float* scores;
    int cpu;
    volatile int* stop;
    pthread_barrier_t* start;
    uint64_t lines;      /* records scanned in this sample */
    uint64_t anomalies;
} dmo_worker;

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void* worker_main(void* arg) {
    dmo_worker* w = (dmo_worker*)arg;
    if (w->cpu >= 0) {
        cpu_set_t set;
        CPU_ZERO(&set);
        CPU_SET(w->cpu, &set);
        pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
    }
    if (!w->det) {
        w->det = dmo_create(w->n_keys, w->keys_blob, w->key_lens);
        if (w->train && w->train_bytes) dmo_process(w->det, w->train, w->train_bytes, ~0ull, NULL, NULL, NULL);
        w->flags = (uint8_t*)malloc((256u << 10) + 4096);
        w->scores = (float*)malloc(((256u << 10) + 4096) * sizeof(float));
        memset(w->flags, 0, (256u << 10) + 4096);
        memset(w->scores, 0, ((256u << 10) + 4096) * sizeof(float));
    }
    pthread_barrier_wait(w->start);
    uint64_t lines = 0, anomalies = 0;
    /* pieces of 256 KiB so that `stop` is seen promptly */
    const uint64_t piece = 256u << 10;
    while (!*w->stop) {
        uint64_t off = 0;
        while (off < w->shard_bytes && !*w->stop) {
            uint64_t end = off + piece < w->shard_bytes ? off + piece : w->shard_bytes;
            if (end < w->shard_bytes) {                       /* cut at a record boundary */
                const uint8_t* nl = (const uint8_t*)memchr(w->shard + end, '\n', w->shard_bytes - end);
                end = nl ? (uint64_t)(nl - w->shard) + 1 : w->shard_bytes;
            }
            const uint64_t n = dmo_process(w->det, w->shard + off, end - off, 0, w->flags, w->scores, NULL);
            for (uint64_t i = 0; i < n; i++) anomalies += w->flags[i];
            lines += n;
            off = end;
        }
    }
    w->lines = lines;
    w->anomalies = anomalies;
    return NULL;
}

/* Returns the number of worker threads used (<= n_threads: the CPUs this process may run on), or < 0.
 *   train / detect : one training message (all records train) and one detection message; worker t scans the
 *                    t-th of n_threads equal shards of `detect` (cut at record boundaries).
 *   rates_out[n_samples] : records/s of every sample; anomalies_out: alerts seen (sanity). */
int dmo_bench_threads(int n_keys, const uint8_t* keys_blob, const uint32_t* key_lens, const uint8_t* train, uint64_t train_bytes,
                      const uint8_t* detect, uint64_t detect_bytes, int n_threads, double min_seconds, int n_samples,
                      double* rates_out, uint64_t* anomalies_out) {
    if (n_threads < 1 || n_samples < 1 || !detect || !detect_bytes) return -1;
    cpu_set_t allowed;
    CPU_ZERO(&allowed);
    int cpus[4096], n_cpus = 0;
    if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0)
        for (int c = 0; c < CPU_SETSIZE && n_cpus < 4096; c++)
            if (CPU_ISSET(c, &allowed)) cpus[n_cpus++] = c;
    if (n_cpus > 0 && n_threads > n_cpus) n_threads = n_cpus;
    dmo_worker* w = (dmo_worker*)calloc((size_t)n_threads, sizeof(dmo_worker));
    pthread_t* th = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));
    if (!w || !th) return -2;
    volatile int stop = 0;
    pthread_barrier_t start;
    uint64_t cut = 0;
    for (int t = 0; t < n_threads; t++) {
        w[t].n_keys = n_keys; w[t].keys_blob = keys_blob; w[t].key_lens = key_lens;
        w[t].train = train; w[t].train_bytes = train_bytes;
        uint64_t end = t + 1 == n_threads ? detect_bytes : detect_bytes * (uint64_t)(t + 1) / (uint64_t)n_threads;
        if (end < detect_bytes) {
            const uint8_t* nl = (const uint8_t*)memchr(detect + end, '\n', detect_bytes - end);
            end = nl ? (uint64_t)(nl - detect) + 1 : detect_bytes;
        }
        if (end < cut) end = cut;
        w[t].shard = detect + cut;
        w[t].shard_bytes = end - cut;
        cut = end;
        w[t].cpu = n_cpus > 0 ? cpus[t % n_cpus] : -1;
        w[t].stop = &stop;
        w[t].start = &start;
    }
    uint64_t anomalies = 0;
    for (int s = 0; s < n_samples; s++) {
        stop = 0;
        pthread_barrier_init(&start, NULL, (unsigned)n_threads + 1);
        for (int t = 0; t < n_threads; t++) pthread_create(&th[t], NULL, worker_main, &w[t]);
        pthread_barrier_wait(&start);              /* (in the first sample the workers build and train their detectors first) */
        const double t0 = now_s();
        while (now_s() - t0 < min_seconds) usleep(2000);
        stop = 1;
        uint64_t lines = 0;
        for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
        const double dt = now_s() - t0;
        for (int t = 0; t < n_threads; t++) { lines += w[t].lines; anomalies += w[t].anomalies; }
        rates_out[s] = (double)lines / dt;
        pthread_barrier_destroy(&start);
    }
    if (anomalies_out) *anomalies_out = anomalies;
    for (int t = 0; t < n_threads; t++) { dmo_destroy(w[t].det); free(w[t].flags); free(w[t].scores); }
    free(w); free(th);
    return n_threads;
}
