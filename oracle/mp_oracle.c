/*
 * oracle/mp_oracle.c -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product path (libldpc_b200.so) never does.
 *
 * Plain-C restatement of the reference BEC flooding decoder
 *     /root/reference/message_passing.c:7-82
 * written from its behavioural contract, NOT a copy: where the
 * reference materialises one message per check socket (Mcv[(n-k)*dc], :24-45) and
 * then searches sockets per erased variable (:52-65), this restatement keeps two
 * numbers per check -- how many of its sockets are erased and the XOR of its known
 * sockets -- which determine every outgoing message.
 *
 * PARITY PINNING: the reference has no tests or golden vectors for this path
 * (SURVEY.md 4).  This restatement is pinned (tests/test_oracle.py) against
 *   (a) the known-answer vectors of SURVEY.md 8c (obtained from the compiled reference),
 *   (b) tests/golden/*.npz, outputs of the UNMODIFIED reference C compiled in the build
 *       container (script: tests/golden/make_golden.py), and
 *   (c) when oracle/_ref exists, live differential runs against the compiled reference.
 */
#include "mc_harness.h"
#include "graph_twin.h"

/*
 * Same signature and semantics as message_passing.c:7.
 *   Mvc[n] in {0,1,2}, 2 = erased, updated in place          (:54-60)
 *   errors[it] += #erased after iteration it (caller zeroes)  (:70-75)
 *   returns it of first erasure-free iteration, else iterations (:76-82)
 */
int orc_message_passing(int *Mvc, int iterations, int *v2c, int *c2v, int *errors,
                        int n, int k, int dv, int dc) {
    int m = n - k;
    int *n_erased = (int *)malloc(sizeof(int) * (size_t)(m > 0 ? m : 1));
    int *parity = (int *)malloc(sizeof(int) * (size_t)(m > 0 ? m : 1));
    int it;
    for (it = 0; it < iterations; it++) {
        /* stall rule, message_passing.c:16-19: only from the third iteration on */
        if (it >= 2 && errors[it - 1] == errors[it - 2]) {
            errors[it] = errors[it - 1];
            continue;
        }
        /* check summary from the PRE-iteration state (flooding schedule) */
        for (int c = 0; c < m; c++) {
            int ne = 0, par = 0;
            for (int s = 0; s < dc; s++) {
                int val = Mvc[c2v[c * dc + s]];
                if (val == 2) ne++; else par ^= (val & 1);
            }
            n_erased[c] = ne;
            parity[c] = par;
        }
        /* An erased variable i is resolved by a check that lists it (by socket, :58) and whose
         * only erased socket is that one; the value is the XOR of the other sockets (:35-42).
         * Candidates are visited in variable_to_check order, then socket order, and the last
         * one wins because the reference has no break (:56-61). */
        for (int i = 0; i < n; i++) {
            if (Mvc[i] != 2) continue;
            int resolved = 0, value = 0;
            for (int t = 0; t < dv; t++) {
                int c = v2c[i * dv + t];
                if (n_erased[c] != 1) continue;      /* another socket (maybe i again) is erased */
                for (int s = 0; s < dc; s++)
                    if (c2v[c * dc + s] == i) { resolved = 1; value = parity[c]; }
            }
            if (resolved) Mvc[i] = value;
        }
        int count = 0;
        for (int i = 0; i < n; i++) count += (Mvc[i] == 2);
        errors[it] += count;
        if (count == 0) break;
    }
    free(n_erased); free(parity);
    return it;
}

int orc_executed(int ret, const int *errors, int iterations) {
    return orc_executed_iterations(ret, errors, iterations);
}

int orc_mc_run(const int *v2c, const int *c2v, int n, int k, int dv, int dc,
               const double *eps, int npoints, uint64_t seed, uint64_t first_frame, uint64_t num_frames,
               int iterations, int expurgation, int nthreads, uint64_t *counters, double *decode_seconds) {
    return orc_mc_run_with(orc_message_passing, v2c, c2v, n, k, dv, dc, eps, npoints, seed, first_frame,
                           num_frames, iterations, expurgation, nthreads, counters, decode_seconds);
}

int orc_decode_batch(const int *v2c, const int *c2v, int n, int k, int dv, int dc,
                     int8_t *Mvc, long frames, int iterations, int *errors, int *ret) {
    return orc_decode_batch_with(orc_message_passing, v2c, c2v, n, k, dv, dc, Mvc, frames, iterations, errors, ret);
}

void orc_draw_patterns(int8_t *out, int n, double eps, uint64_t seed, uint32_t point,
                       uint64_t first_frame, long frames) {
    orc_draw_patterns_i8(out, n, eps, seed, point, first_frame, frames);
}

uint32_t orc_bec_word_export(uint64_t seed, uint32_t point, uint64_t q, uint32_t var, double eps) {
    return orc_bec_word(seed, point, q, var, orc_bec_threshold(eps));
}

void orc_philox_export(uint32_t *ctr4, uint32_t k0, uint32_t k1) { orc_philox4x32_10(ctr4, k0, k1); }

int orc_generate_graph(int n, int dv, int dc, uint64_t seed, int *variable_lookup, int *check_lookup) {
    return orc_generate_graph_twin(n, dv, dc, seed, variable_lookup, check_lookup);
}
