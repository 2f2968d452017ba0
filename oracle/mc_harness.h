/*
 * oracle/mc_harness.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Frame loops around a single-frame decoder with the reference signature
 * (/root/reference/message_passing.c:7).  The decoder is passed as a function
 * pointer so the same loops drive (a) the restatement in mp_oracle.c and
 * (b) the UNMODIFIED reference object compiled into oracle/_ref (ref_harness.c).
 *
 * The accumulation restates the reference drivers:
 *   parallel_simulator_expurgated.py:233-243  (ensemble loop, expurgation filter)
 *   parallel_simulator_expurgated.py:367-379  (fixed-code loop, no filter == expurgation -1)
 *   parallel_simulator_expurgated.py:165      (errors = [initial_count] + errors)
 * with the all-zero codeword (:233) and the erasure patterns of bec_rng.h in
 * place of numpy's unseeded generator (channels.py:24-26).
 *
 * Counter block per erasure point (uint64, additive across frame ranges):
 *   [0 .. its]   sum over COUNTED frames of [initial_count, errors[0..its-1]]
 *   [its+1]      block errors   (counted frames with errors[-1] != 0)
 *   [its+2]      bit errors     (sum of errors[-1] over counted frames)
 *   [its+3]      frames         (all frames, the reference's num_tests denominator)
 *   [its+4]      counted frames (errors[-1] > expurgation)
 *   [its+5]      executed iterations summed over ALL frames (ret+1 on
 *                success, else first skipped index, else its) -- edge-updates = this * n*dv
 */
#ifndef ORACLE_MC_HARNESS_H
#define ORACLE_MC_HARNESS_H
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <time.h>
#include "bec_rng.h"

typedef int (*orc_decode_fn)(int *Mvc, int iterations, int *v2c, int *c2v, int *errors,
                             int n, int k, int dv, int dc);

#define ORC_NCOUNTERS(its) ((its) + 6)

/* iterations the reference actually executed for one frame (work metric) */
static inline int orc_executed_iterations(int ret, const int *errors, int iterations) {
    if (ret < iterations) return ret + 1;
    for (int it = 2; it < iterations; it++)
        if (errors[it - 1] == errors[it - 2]) return it;
    return iterations;
}

/* the 32-frame words of word index q for all n variables */
static inline void orc_draw_words(uint32_t *words, int n, uint64_t seed, uint32_t point, uint64_t q, uint64_t T) {
    for (int i = 0; i < n; i++) words[i] = orc_bec_word(seed, point, q, (uint32_t)i, T);
}

/* erasure pattern of one frame -> Mvc (0 / 2), all-zero codeword; words/cached_q cache word q */
static inline void orc_draw_frame(int *Mvc, int n, uint64_t seed, uint32_t point, uint64_t frame, uint64_t T,
                                  uint32_t *words, uint64_t *cached_q) {
    uint64_t q = frame >> 5;
    int bit = (int)(frame & 31);
    if (*cached_q != q) { orc_draw_words(words, n, seed, point, q, T); *cached_q = q; }
    for (int i = 0; i < n; i++) Mvc[i] = ((words[i] >> bit) & 1u) ? 2 : 0;
}

static inline double orc_now(void) {
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

typedef struct {
    orc_decode_fn fn;
    int *v2c, *c2v;
    int n, k, dv, dc, iterations, expurgation;
    uint64_t seed, T, first_frame, num_frames;
    uint32_t point;
    uint64_t *counters; /* ORC_NCOUNTERS, private to the thread */
    double decode_seconds; /* out: wall time spent inside fn() only */
} orc_mc_job;

static void *orc_mc_worker(void *arg) {
    orc_mc_job *jb = (orc_mc_job *)arg;
    int its = jb->iterations, n = jb->n;
    int *Mvc = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    int *errors = (int *)malloc(sizeof(int) * (size_t)(its > 0 ? its : 1));
    uint32_t *words = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n > 0 ? n : 1));
    uint64_t cached_q = ~(uint64_t)0;
    uint64_t *C = jb->counters;
    double tdec = 0.0;
    for (uint64_t f = jb->first_frame; f < jb->first_frame + jb->num_frames; f++) {
        orc_draw_frame(Mvc, n, jb->seed, jb->point, f, jb->T, words, &cached_q);
        int initial = 0;
        for (int i = 0; i < n; i++) initial += (Mvc[i] == 2);
        memset(errors, 0, sizeof(int) * (size_t)its);
        double t0 = orc_now();
        int ret = jb->fn(Mvc, its, jb->v2c, jb->c2v, errors, n, jb->k, jb->dv, jb->dc);
        tdec += orc_now() - t0;
        int last = its > 0 ? errors[its - 1] : initial;
        C[its + 3] += 1;
        C[its + 5] += (uint64_t)orc_executed_iterations(ret, errors, its);
        if (last > jb->expurgation) {
            C[0] += (uint64_t)initial;
            for (int it = 0; it < its; it++) C[1 + it] += (uint64_t)errors[it];
            C[its + 1] += (last != 0);
            C[its + 2] += (uint64_t)last;
            C[its + 4] += 1;
        }
    }
    jb->decode_seconds = tdec;
    free(Mvc); free(errors); free(words);
    return NULL;
}

/* counters: [npoints][ORC_NCOUNTERS(iterations)], overwritten */
static int orc_mc_run_with(orc_decode_fn fn, const int *v2c, const int *c2v, int n, int k, int dv, int dc,
                           const double *eps, int npoints, uint64_t seed, uint64_t first_frame,
                           uint64_t num_frames, int iterations, int expurgation, int nthreads,
                           uint64_t *counters, double *decode_seconds /* may be NULL: max over threads, summed over points */) {
    if (nthreads < 1) nthreads = 1;
    int nc = ORC_NCOUNTERS(iterations);
    memset(counters, 0, sizeof(uint64_t) * (size_t)nc * (size_t)npoints);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    orc_mc_job *jobs = (orc_mc_job *)malloc(sizeof(orc_mc_job) * (size_t)nthreads);
    uint64_t *priv = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)nc * (size_t)nthreads);
    double tsum = 0.0;
    for (int p = 0; p < npoints; p++) {
        double tmax = 0.0;
        memset(priv, 0, sizeof(uint64_t) * (size_t)nc * (size_t)nthreads);
        uint64_t T = orc_bec_threshold(eps[p]);
        for (int t = 0; t < nthreads; t++) {
            uint64_t lo = num_frames * (uint64_t)t / (uint64_t)nthreads;
            uint64_t hi = num_frames * (uint64_t)(t + 1) / (uint64_t)nthreads;
            orc_mc_job jb = { fn, (int *)v2c, (int *)c2v, n, k, dv, dc, iterations, expurgation,
                              seed, T, first_frame + lo, hi - lo, (uint32_t)p, priv + (size_t)t * nc, 0.0 };
            jobs[t] = jb;
            pthread_create(&th[t], NULL, orc_mc_worker, &jobs[t]);
        }
        for (int t = 0; t < nthreads; t++) {
            pthread_join(th[t], NULL);
            if (jobs[t].decode_seconds > tmax) tmax = jobs[t].decode_seconds;
            for (int c = 0; c < nc; c++) counters[(size_t)p * nc + c] += priv[(size_t)t * nc + c];
        }
        tsum += tmax;
    }
    if (decode_seconds) *decode_seconds = tsum;
    free(th); free(jobs); free(priv);
    return 0;
}

/* Batched decode: Mvc [frames][n] int8 in {0,1,2}, in place; errors [frames][iterations]
 * (the C array, without the prepended initial count); ret [frames]. */
static int orc_decode_batch_with(orc_decode_fn fn, const int *v2c, const int *c2v, int n, int k, int dv, int dc,
                                 int8_t *Mvc, long frames, int iterations, int *errors, int *ret) {
    int *buf = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    int *err = (int *)malloc(sizeof(int) * (size_t)(iterations > 0 ? iterations : 1));
    for (long f = 0; f < frames; f++) {
        for (int i = 0; i < n; i++) buf[i] = Mvc[(size_t)f * n + i];
        memset(err, 0, sizeof(int) * (size_t)iterations);
        ret[f] = fn(buf, iterations, (int *)v2c, (int *)c2v, err, n, k, dv, dc);
        for (int i = 0; i < n; i++) Mvc[(size_t)f * n + i] = (int8_t)buf[i];
        for (int it = 0; it < iterations; it++) errors[(size_t)f * iterations + it] = err[it];
    }
    free(buf); free(err);
    return 0;
}

/* Draw patterns for frames [first_frame, first_frame+frames) as int8 0/2, [frames][n] */
static void orc_draw_patterns_i8(int8_t *out, int n, double eps, uint64_t seed, uint32_t point,
                                 uint64_t first_frame, long frames) {
    uint64_t T = orc_bec_threshold(eps);
    int *tmp = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    uint32_t *words = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n > 0 ? n : 1));
    uint64_t cached_q = ~(uint64_t)0;
    for (long f = 0; f < frames; f++) {
        orc_draw_frame(tmp, n, seed, point, first_frame + (uint64_t)f, T, words, &cached_q);
        for (int i = 0; i < n; i++) out[(size_t)f * n + i] = (int8_t)tmp[i];
    }
    free(tmp); free(words);
}

#endif
