/*
 * lwo_bench.c -- CPU baseline driver for bench.py (cpu_baseline / --impl reference).
 * TEST/BENCH INFRASTRUCTURE ONLY (see lewton_oracle.h).
 *
 * Work per chain and packet, mirroring the steady state of audio.rs:1043-1154
 * for long/long blocks: inverse_mdct (imdct.rs:291), left-half window/OLA
 * against the previous block's saved right half, save the right half, emit n/2
 * f32 samples.  Scratch is preallocated (the reference's per-call vec!,
 * imdct.rs:302, is left out, in the CPU's favour).  Static partition of chains
 * over `threads` pthreads; returns wall seconds (CLOCK_MONOTONIC) of the
 * parallel region.
 */
#define _GNU_SOURCE
#include "lewton_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* defined in lewton_oracle.c */
void lwo_inverse_mdct_with_scratch(const lwo_tables *t, float *buffer, float *buf2);

typedef struct {
    const lwo_tables *t;
    int c0, c1, packets, reps;
    const float *spectrum;
    float *out;
} job_t;

static void *worker(void *arg)
{
    job_t *j = (job_t *)arg;
    const int n = j->t->n, n2 = n >> 1;
    const float *w = j->t->window;
    float *x = (float *)malloc(sizeof(float) * n);
    float *scratch = (float *)malloc(sizeof(float) * n2);
    float *prev = (float *)malloc(sizeof(float) * n2);
    for (int rep = 0; rep < j->reps; rep++)
    for (int c = j->c0; c < j->c1; c++) {
        int has = 0;
        float *o = j->out + (size_t)c * j->packets * n2;
        for (int p = 0; p < j->packets; p++) {
            const float *s = j->spectrum + ((size_t)c * j->packets + p) * n2;
            memcpy(x, s, sizeof(float) * n2);
            lwo_inverse_mdct_with_scratch(j->t, x, scratch);
            if (has) {
                for (int i = 0; i < n2; i++)
                    o[i] = (x[i] * w[i]) + (prev[i] * w[n2 - 1 - i]);
                o += n2;
            }
            memcpy(prev, x + n2, sizeof(float) * n2);
            has = 1;
        }
    }
    free(x); free(scratch); free(prev);
    return NULL;
}

double lwo_bench_chains(int bs, int chains, int packets, const float *spectrum,
                        float *out, int threads, int reps)
{
    lwo_tables *t = lwo_tables_new(bs);
    if (!t) return -1.0;
    if (threads < 1) threads = 1;
    if (threads > chains) threads = chains;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * threads);
    job_t *jobs = (job_t *)malloc(sizeof(job_t) * threads);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int i = 0; i < threads; i++) {
        jobs[i].t = t;
        jobs[i].c0 = (int)((long)chains * i / threads);
        jobs[i].c1 = (int)((long)chains * (i + 1) / threads);
        jobs[i].packets = packets;
        jobs[i].reps = reps < 1 ? 1 : reps;
        jobs[i].spectrum = spectrum;
        jobs[i].out = out;
        pthread_create(&th[i], NULL, worker, &jobs[i]);
    }
    for (int i = 0; i < threads; i++) pthread_join(th[i], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(th); free(jobs);
    lwo_tables_free(t);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
