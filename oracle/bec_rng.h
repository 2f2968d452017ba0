/*
 * oracle/bec_rng.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Host restatement of the counter-based BEC erasure-pattern generator that the
 * CUDA path draws in-kernel.  It replaces, for the GPU path, the reference's
 * channel  BEC.new_transmit  (/root/reference/channels.py:24-26):
 *     random_vector = np.random.rand(n); where(random_vector < eps, 2, x)
 * i.e. i.i.d. erasures with probability eps, strict '<'.
 *
 * The reference draws from numpy's unseeded global generator, so there is no
 * reference bit-stream to match; what must match bit-for-bit is this file
 * against the device implementation (iib_project_ldpc_codes_b200/csrc/bec_rng.cuh),
 * so that the reference C decoder can be fed exactly the patterns the GPU
 * decoded.  The two files are written independently from the definition below.
 *
 * Definition
 * ----------
 *   Philox4x32-10 (Salmon et al., SC'11), key = (seed_lo, seed_hi).
 *   For erasure point p (< 2^24), 32-frame word index q = frame >> 5 (64 bit),
 *   variable i (< 2^32), block b (< 8):
 *       (r[4b], r[4b+1], r[4b+2], r[4b+3]) = philox(ctr = (i, q_lo, q_hi, p<<8 | b))
 *   Threshold  T = ceil(eps * 2^32)  clipped to [0, 2^32]  (u < eps  <=>  U < T
 *   for U uniform on 32-bit integers, u = U / 2^32).
 *   The 32 frames of word q compare their private uniform integers, MSB first,
 *   against T in bit-sliced form: random word r[j] carries bit (31-j) of all 32
 *   uniform integers.
 *       lt = 0, eq = ~0
 *       for j = 0..31:  t = bit (31-j) of T
 *           if t: lt |= eq & ~r[j]; eq &= r[j]      else: eq &= ~r[j]
 *           stop when eq == 0 or when all lower bits of T are zero
 *   Frame f = 32 q + bit erases variable i  <=>  bit (f & 31) of lt is set.
 *   T == 2^32 -> every bit erased, T == 0 -> none.
 */
#ifndef ORACLE_BEC_RNG_H
#define ORACLE_BEC_RNG_H
#include <stdint.h>
#include <math.h>

static inline void orc_philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int round = 0; round < 10; round++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

/* threshold: returns T in [0, 2^32] as uint64 */
static inline uint64_t orc_bec_threshold(double eps) {
    if (!(eps > 0.0)) return 0;
    if (eps >= 1.0) return (uint64_t)1 << 32;
    double t = ceil(eps * 4294967296.0);
    if (t >= 4294967296.0) return (uint64_t)1 << 32;
    return (uint64_t)t;
}

/* 32 Bernoulli(T/2^32) bits for (seed, point, word q, variable i) */
static inline uint32_t orc_bec_word(uint64_t seed, uint32_t point, uint64_t q, uint32_t var, uint64_t T) {
    if (T == 0) return 0u;
    if (T >= ((uint64_t)1 << 32)) return 0xFFFFFFFFu;
    uint32_t t = (uint32_t)T, lt = 0, eq = 0xFFFFFFFFu;
    uint32_t r[4];
    for (int j = 0; j < 32; j++) {
        if ((j & 3) == 0) {
            r[0] = var; r[1] = (uint32_t)q; r[2] = (uint32_t)(q >> 32);
            r[3] = (point << 8) | (uint32_t)(j >> 2);
            orc_philox4x32_10(r, (uint32_t)seed, (uint32_t)(seed >> 32));
        }
        uint32_t w = r[j & 3];
        if ((t >> (31 - j)) & 1u) { lt |= eq & ~w; eq &= w; }
        else                      { eq &= ~w; }
        if (eq == 0) break;
        if (j == 31 || (uint32_t)(t << (j + 1)) == 0) break;
    }
    return lt;
}

#endif
