/*
 * oracle/ref_harness.c -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Frame loops (mc_harness.h) around the UNMODIFIED reference decoder.  The Makefile
 * compiles /root/reference/message_passing.c where it lies and links the object into
 * oracle/_ref/libref_harness.so together with this file; no reference source is
 * copied into the repository.  `message_passing` below is the reference symbol
 * (/root/reference/message_passing.c:7).
 *
 * Used as: the parity checker for the fused Monte-Carlo CUDA path (same erasure
 * patterns, counters must match bit-for-bit) and as the CPU baseline
 * (cpu_baseline.kind == "reference") with one thread per host core.
 */
#include "mc_harness.h"

extern int message_passing(int *Mvc, int iterations, int *variable_to_check_list, int *check_to_variable_list,
                           int *errors, int n, int k, int dv, int dc);

int ref_mc_run(const int *v2c, const int *c2v, int n, int k, int dv, int dc,
               const double *eps, int npoints, uint64_t seed, uint64_t first_frame, uint64_t num_frames,
               int iterations, int expurgation, int nthreads, uint64_t *counters, double *decode_seconds) {
    return orc_mc_run_with(message_passing, v2c, c2v, n, k, dv, dc, eps, npoints, seed, first_frame,
                           num_frames, iterations, expurgation, nthreads, counters, decode_seconds);
}

int ref_decode_batch(const int *v2c, const int *c2v, int n, int k, int dv, int dc,
                     int8_t *Mvc, long frames, int iterations, int *errors, int *ret) {
    return orc_decode_batch_with(message_passing, v2c, c2v, n, k, dv, dc, Mvc, frames, iterations, errors, ret);
}
