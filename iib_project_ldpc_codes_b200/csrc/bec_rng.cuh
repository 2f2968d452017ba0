// bec_rng.cuh -- in-kernel BEC erasure-pattern generator (device side).
//
// Replaces the reference channel  BEC.new_transmit  (/root/reference/channels.py:24-26:
// `np.where(np.random.rand(n) < eps, 2, x)`) for the fused Monte-Carlo path: erasures are
// i.i.d. Bernoulli(eps), strict '<', drawn from a counter-based generator keyed by
// (seed, erasure point, 32-frame word index, variable) so that the pattern of frame f
// does not depend on the GPU count, the launch geometry or the words-per-group choice.
//
//   key      = (seed_lo, seed_hi)                                  Philox4x32-10
//   counter  = (variable, q_lo, q_hi, point << 8 | block)          q = frame >> 5
//   block b yields the random words r[4b .. 4b+3]; r[j] carries bit (31-j) of the 32
//   per-frame uniform integers U; frame (32q + bit) erases the variable iff U < T,
//   T = ceil(eps * 2^32)  (computed on the host, see ldpc_b200.cu).
//
// The comparison runs MSB-first in bit-sliced form and stops as soon as every frame of
// the word is decided, or all remaining bits of T are zero.
#pragma once
#include <stdint.h>

namespace ldpc {

__device__ __forceinline__ void philox4x32_10(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3,
                                              uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

// The same rounds with the 20 round keys READ from `rk` (k0 + r*0x9E3779B9, k1 + r*0xBB67AE85 for r = 0..9, made on
// the host: philox_round_keys).  When `rk` is a kernel parameter the keys are constant-bank operands of the LOP3s
// and occupy no registers, which lets the scheduler interleave more independent chains.
__device__ __forceinline__ void philox4x32_10_rk(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3,
                                                 const uint32_t (&rk)[20]) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ rk[2 * r];
        const uint32_t n2 = hi0 ^ c3 ^ rk[2 * r + 1];
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
}
__host__ __device__ inline void philox_round_keys(uint32_t seed_lo, uint32_t seed_hi, uint32_t (&rk)[20]) {
    for (int r = 0; r < 10; r++) { rk[2 * r] = seed_lo + (uint32_t)r * 0x9E3779B9u; rk[2 * r + 1] = seed_hi + (uint32_t)r * 0xBB67AE85u; }
}

struct BecParams {
    uint32_t seed_lo, seed_hi;
    uint32_t threshold;  // low 32 bits of T
    uint32_t mode;       // 0: none erased (T == 0), 1: compare against threshold, 2: all erased (T == 2^32)
};

// 32 Bernoulli bits for (point, word q, variable var)
__device__ __forceinline__ uint32_t bec_word(const BecParams& bp, uint32_t point, uint64_t q, uint32_t var) {
    if (bp.mode != 1) return bp.mode == 2 ? 0xFFFFFFFFu : 0u;
    const uint32_t t = bp.threshold;
    uint32_t lt = 0, eq = 0xFFFFFFFFu;
#pragma unroll 1
    for (int b = 0; b < 8; b++) {
        uint32_t r0 = var, r1 = (uint32_t)q, r2 = (uint32_t)(q >> 32), r3 = (point << 8) | (uint32_t)b;
        philox4x32_10(r0, r1, r2, r3, bp.seed_lo, bp.seed_hi);
        const uint32_t r[4] = {r0, r1, r2, r3};
        // Four levels without branches: m = all-ones iff the threshold bit of the level is set (warp-uniform).
        //   bit set:   lt |= eq & ~r ; eq &= r        bit clear:   eq &= ~r
        // Levels past the point where every frame is decided (eq == 0) or where no lower threshold bit is
        // set cannot change lt, so the exit tests only need to run once per block.
        const uint32_t tb = t << (4 * b);
#pragma unroll
        for (int s = 0; s < 4; s++) {
            const uint32_t m = (uint32_t)((int32_t)(tb << s) >> 31);
            lt |= eq & ~r[s] & m;
            eq &= ~(r[s] ^ m);
        }
        if (eq == 0u || b == 7 || (t << (4 * b + 4)) == 0u) break;
    }
    return lt;
}

// One frame of one variable: bit `bit` of bec_word(bp, point, q, var), without deciding the other 31 frames of the
// word (1.07 Philox blocks on average instead of 3.2).  Used to re-draw single frames (expurgation re-count).
__device__ __forceinline__ uint32_t bec_lane_bit(const BecParams& bp, uint32_t point, uint64_t q, uint32_t var, uint32_t bit) {
    if (bp.mode != 1) return bp.mode == 2 ? 1u : 0u;
    const uint32_t t = bp.threshold;
#pragma unroll 1
    for (int b = 0; b < 8; b++) {
        uint32_t r0 = var, r1 = (uint32_t)q, r2 = (uint32_t)(q >> 32), r3 = (point << 8) | (uint32_t)b;
        philox4x32_10(r0, r1, r2, r3, bp.seed_lo, bp.seed_hi);
        const uint32_t r[4] = {r0, r1, r2, r3};
#pragma unroll
        for (int s = 0; s < 4; s++) {   // level 4b+s compares bit 31-(4b+s) of the frame's uniform integer with the threshold's
            const uint32_t tb = (t >> (31 - (4 * b + s))) & 1u, ub = (r[s] >> bit) & 1u;
            if (tb != ub) return tb;    // first differing bit decides: U < T iff the threshold has the 1
        }
        if ((t << (4 * b + 4)) == 0u) return 0u;   // all remaining threshold bits are zero: U >= T
    }
    return 0u;   // U == T
}

// ---- the same generator, split for the team-level draw of the E1 kernels (e1_kernels.cuh: team_draw) ----
// One Philox block = four comparison levels.  bec_levels4 advances (lt, eq) by the four levels of block b.
__device__ __forceinline__ void bec_block(const BecParams& bp, uint32_t point, uint64_t q, uint32_t var, int b,
                                          uint32_t& lt, uint32_t& eq) {
    uint32_t r0 = var, r1 = (uint32_t)q, r2 = (uint32_t)(q >> 32), r3 = (point << 8) | (uint32_t)b;
    philox4x32_10(r0, r1, r2, r3, bp.seed_lo, bp.seed_hi);
    const uint32_t r[4] = {r0, r1, r2, r3};
    const uint32_t tb = bp.threshold << (4 * b);
#pragma unroll
    for (int s = 0; s < 4; s++) {
        const uint32_t m = (uint32_t)((int32_t)(tb << s) >> 31);
        lt |= eq & ~r[s] & m;
        eq &= ~(r[s] ^ m);
    }
}

__device__ __forceinline__ void bec_block_rk(const BecParams& bp, const uint32_t (&rk)[20], uint32_t point, uint64_t q,
                                             uint32_t var, int b, uint32_t& lt, uint32_t& eq) {
    uint32_t r0 = var, r1 = (uint32_t)q, r2 = (uint32_t)(q >> 32), r3 = (point << 8) | (uint32_t)b;
    philox4x32_10_rk(r0, r1, r2, r3, rk);
    const uint32_t r[4] = {r0, r1, r2, r3};
    const uint32_t tb = bp.threshold << (4 * b);
#pragma unroll
    for (int s = 0; s < 4; s++) {
        const uint32_t m = (uint32_t)((int32_t)(tb << s) >> 31);
        lt |= eq & ~r[s] & m;
        eq &= ~(r[s] ^ m);
    }
}

// true when block b was the last one bec_word would have executed for this word
__device__ __forceinline__ bool bec_done_after(const BecParams& bp, int b, uint32_t eq) {
    return eq == 0u || b == 7 || (bp.threshold << (4 * b + 4)) == 0u;
}

// level masks of the first two blocks (hoisted out of the per-variable loop by the caller)
struct BecMasks8 { uint32_t m[8]; };
__device__ __forceinline__ BecMasks8 bec_masks8(const BecParams& bp) {
    BecMasks8 M;
#pragma unroll
    for (int s = 0; s < 8; s++) M.m[s] = (uint32_t)((int32_t)(bp.threshold << s) >> 31);
    return M;
}

// blocks 0 and 1 without any exit test: identical (lt) to bec_word's first two blocks -- levels past the
// point where bec_word would have stopped cannot change lt (eq == 0, or every remaining mask is zero).
__device__ __forceinline__ void bec_first8(const BecParams& bp, const BecMasks8& M, uint32_t point, uint64_t q,
                                           uint32_t var, uint32_t& lt, uint32_t& eq) {
    lt = 0u; eq = 0xFFFFFFFFu;
#pragma unroll
    for (int b = 0; b < 2; b++) {
        uint32_t r0 = var, r1 = (uint32_t)q, r2 = (uint32_t)(q >> 32), r3 = (point << 8) | (uint32_t)b;
        philox4x32_10(r0, r1, r2, r3, bp.seed_lo, bp.seed_hi);
        const uint32_t r[4] = {r0, r1, r2, r3};
#pragma unroll
        for (int s = 0; s < 4; s++) {
            const uint32_t m = M.m[4 * b + s];
            lt |= eq & ~r[s] & m;
            eq &= ~(r[s] ^ m);
        }
    }
}

// bec_first8 with host-made round keys
__device__ __forceinline__ void bec_first8_rk(const BecMasks8& M, const uint32_t (&rk)[20], uint32_t point, uint64_t q,
                                              uint32_t var, uint32_t& lt, uint32_t& eq) {
    lt = 0u; eq = 0xFFFFFFFFu;
#pragma unroll
    for (int b = 0; b < 2; b++) {
        uint32_t r0 = var, r1 = (uint32_t)q, r2 = (uint32_t)(q >> 32), r3 = (point << 8) | (uint32_t)b;
        philox4x32_10_rk(r0, r1, r2, r3, rk);
        const uint32_t r[4] = {r0, r1, r2, r3};
#pragma unroll
        for (int s = 0; s < 4; s++) {
            const uint32_t m = M.m[4 * b + s];
            lt |= eq & ~r[s] & m;
            eq &= ~(r[s] ^ m);
        }
    }
}

}  // namespace ldpc
