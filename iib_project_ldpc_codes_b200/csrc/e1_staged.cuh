// e1_staged.cuh -- staged variant of the E1 Monte-Carlo kernel: survivors are re-packed between stages.
//
// A bit-sliced frame group runs in lock step until its slowest frame stops; at n = 1000, eps = 0.40 the
// reference executes 16.5 iterations per frame on average but the slowest of 128 frames needs 44
// (SURVEY.md 7.4 "lane inflation", measured 2.68x in profiles/r1a_*).  Frames are independent, so the
// iteration range [0, cap) is cut into stages [c_s, c_{s+1}) per erasure point: a group that reaches
// c_{s+1} with frames still running writes those frames (state column, bit-packed frame-major) to an HBM
// spill buffer and takes new work; the next launch packs 32*WORDS survivors of the SAME age into a dense
// group and continues.  Everything a frame contributes (per-iteration erasure sums, block/bit errors,
// executed iterations) is additive, and all frames of a group always have the same age, so the cheap
// horizontal popcount per iteration stays valid.  Results are bit-identical to the unstaged kernel.
//
// 32x32 bit transposes (warp shuffles, 5 butterfly steps) convert between the bit-sliced shared-memory
// layout (word per node, bit per frame) and the frame-major spill records (word per 32 nodes).
#pragma once
#include "e1_kernels.cuh"

namespace ldpc {

// lane l holds row l; afterwards lane b holds column b  (out_b bit l = in_l bit b).  Warp-collective.
// Five butterfly stages; stage j exchanges the off-diagonal j x j blocks with lane ^ j.  Every lane sends its word
// rotated so that the bits its partner needs already sit in the partner's target positions (left by j from lanes
// with bit j set, right by j from the others), and the receiver merges with one bit-select: 3 instructions per
// stage (funnel shift, shuffle, LOP3) instead of shift/mask/select chains -- the re-packing between stages is
// transposes more than anything else.
__device__ __forceinline__ uint32_t transpose32(uint32_t x, int lane) {
#pragma unroll
    for (int j = 16; j >= 1; j >>= 1) {
        const uint32_t m = j == 16 ? 0x0000FFFFu : j == 8 ? 0x00FF00FFu : j == 4 ? 0x0F0F0F0Fu : j == 2 ? 0x33333333u : 0x55555555u;
        const bool hi = (lane & j) != 0;
        const uint32_t keep = hi ? (m << j) : m;                       // the bits this lane keeps
        const uint32_t send = __funnelshift_l(x, x, hi ? j : 32 - j);  // rotate: left by j (hi) / right by j (lo)
        const uint32_t p = __shfl_xor_sync(0xFFFFFFFFu, send, j);
        x = (x & keep) | (p & ~keep);
    }
    return x;
}

struct StageParams {
    int stage, nstages;
    const int* cuts;                      // [npoints][nstages+1] ages: stage s covers [cuts[s], cuts[s+1])
    const uint32_t* spill_in;             // records of the previous stage
    uint32_t* spill_out;
    const unsigned long long* cnt_in;     // [npoints] records available per point
    unsigned long long* cnt_out;          // [npoints] records written per point (atomic)
    unsigned long long point_stride;      // record capacity per point
    int rec_words;                        // 32-bit words per record = ceil(n/32) rounded up to 4
    unsigned long long* stop_hist;        // optional [npoints][cap+1]: frames by executed-iteration count
    // Expurgated run (parallel_simulator_expurgated.py:238: only frames whose final residual exceeds the threshold
    // contribute to any statistic), pass 1: the staged pipeline adds the denominators only (frames, executed
    // iterations) and LISTS every frame that ends with erasures left; pass 2 (e1_mc_kernel, recount) re-draws exactly
    // those frames, packs them into dense groups and accumulates their statistics.  Every spill record carries its
    // frame's chunk-relative id in its last word (rec_words >= ceil(n/32) + 1).
    int exp_mode;
    uint32_t* id_list;                    // [npoints][point_stride]
    unsigned long long* id_cnt;           // [npoints] (atomic)
    unsigned long long* stats;            // handle statistics: [2] += spilled records, [3] += listed frames
};

// Everything after a group's iterations [lo, age): which frames are deferred to the next stage, the residual of the
// finished ones, the flush of the stage's statistics, and the spill of the deferred frames (bit-packed,
// frame-major) for the next launch.  Shared by the staged kernel and the overlapped stage-0 kernel (e1_overlap.cuh).
template <int WORDS>
__device__ __forceinline__ void stage_finish_group(const E1Params& P, const StageParams& S, const TeamCtx& t, VecT<WORDS>* x,
                                                   const uint32_t* hbuf, uint32_t* work_slot, uint32_t point, int lo, int hi,
                                                   int age, const uint32_t (&valid)[WORDS], const uint32_t (&running)[WORDS],
                                                   const uint32_t (&any)[WORDS], uint32_t cnt, unsigned long long exec_sum,
                                                   const uint32_t* fid /* shared: chunk-relative frame id per lane */) {
    constexpr int W = 32 * WORDS;
    unsigned long long* C = P.counters + (size_t)point * raw_stride(P.cap);
    // frames still running at a stage boundary below the cap are deferred to the next stage
    uint32_t D[WORDS], F[WORDS], ndef = 0;
    LDPC_FORW {
        D[w] = (age == hi && hi < P.cap) ? running[w] : 0u;
        F[w] = valid[w] & ~D[w];
        ndef += __popc(D[w]);
    }
    if (S.stop_hist && t.tid == 0 && hi >= P.cap) {
        uint32_t capped = 0;
        LDPC_FORW capped += __popc(running[w]);
        red_add(S.stop_hist + (size_t)point * (P.cap + 1) + P.cap, capped);   // ran every iteration up to the cap
    }
    uint32_t resid = cnt;   // erased count of the finished frames
    if (ndef && !S.exp_mode) {
        uint32_t none[1] = {0u};
        resid = 0;
        for (int i = t.tid; i < P.n; i += t.nthr) {
            const VecT<WORDS> xi = x[i];
            LDPC_FORW resid += __popc(xi.a[w] & F[w]);
        }
        team_reduce<1>(t, none, resid);
    }
    // ---- flush the statistics of this stage
    team_sync(t);
    const int first_bin = S.stage == 0 ? 0 : 1;
    if (!S.exp_mode)
        for (int b = first_bin + t.tid; b <= age - lo; b += t.nthr) red_add(C + lo + b, hbuf[b]);
    if (t.tid == 0) {
        uint32_t blk = 0, nv = 0;
        LDPC_FORW { blk += __popc(any[w] & F[w]); nv += __popc(valid[w]); }
        if (!S.exp_mode) {
            if (age < P.cap) red_add(C + P.cap + 6 + age + 1, resid);   // bins after `age` keep the finished frames' residual
            red_add(C + P.cap + 1, blk);
            red_add(C + P.cap + 2, resid);
            if (S.stage == 0) red_add(C + P.cap + 4, nv);
        }
        if (S.stage == 0) red_add(C + P.cap + 3, nv);
        red_add(C + P.cap + 5, exec_sum);
        red_add(C + 2 * P.cap + 8, (unsigned long long)(age - lo) * W);
    }
    // ---- expurgated run: list the finished frames that still have erasures (they are re-drawn and counted in pass 2)
    if (S.exp_mode) {
        uint32_t E[WORDS], nemit = 0;
        LDPC_FORW { E[w] = any[w] & F[w]; nemit += __popc(E[w]); }
        if (nemit) {   // team-uniform
            if (t.tid == 0) {
                const unsigned long long v = atomicAdd(S.id_cnt + point, (unsigned long long)nemit);
                work_slot[0] = (uint32_t)v; work_slot[1] = (uint32_t)(v >> 32);
                if (S.stats) atomicAdd(S.stats + 3, (unsigned long long)nemit);
            }
            team_sync(t);
            const unsigned long long ebase = (unsigned long long)work_slot[0] | ((unsigned long long)work_slot[1] << 32);
            team_sync(t);   // work_slot is reused for the spill base below
            if (t.warp == 0) {
                uint32_t before = 0;
                LDPC_FORW {
                    if ((E[w] >> t.lane) & 1u)
                        S.id_list[(size_t)point * S.point_stride + ebase + before + __popc(E[w] & ((1u << t.lane) - 1u))] = fid[w * 32 + t.lane];
                    before += __popc(E[w]);
                }
            }
        }
    }
    // ---- spill the deferred frames (bit-packed, frame-major) for the next stage
    if (ndef) {
        unsigned long long sbase;
        if (t.team_warps == 1) {
            unsigned long long v = 0;
            if (t.lane == 0) {
                v = atomicAdd(S.cnt_out + point, (unsigned long long)ndef);
                if (S.stats) atomicAdd(S.stats + 2, (unsigned long long)ndef);
            }
            sbase = __shfl_sync(0xFFFFFFFFu, v, 0);
        } else {
            if (t.tid == 0) {
                unsigned long long v = atomicAdd(S.cnt_out + point, (unsigned long long)ndef);
                work_slot[0] = (uint32_t)v; work_slot[1] = (uint32_t)(v >> 32);
                if (S.stats) atomicAdd(S.stats + 2, (unsigned long long)ndef);
            }
            team_sync(t);
            sbase = (unsigned long long)work_slot[0] | ((unsigned long long)work_slot[1] << 32);
        }
        uint32_t* out = S.spill_out + (size_t)point * S.point_stride * S.rec_words;
        for (int j0 = t.warp * 4; j0 < S.rec_words; j0 += t.team_warps * 4) {
            uint32_t before = 0;
            const int s0 = 32 * j0 + t.lane;
            VecT<WORDS> v[4];   // whole-vector loads of variables s0, s0+32, s0+64, s0+96 (no bank conflicts)
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (s0 + 32 * k < P.n) v[k] = x[s0 + 32 * k];
                else { LDPC_FORW v[k].a[w] = 0u; }
            }
            LDPC_FORW {
                if (D[w] == 0u) continue;   // team-uniform: no survivor in this word, nothing to transpose
                uint4 r;
                r.x = transpose32(v[0].a[w], t.lane);
                r.y = transpose32(v[1].a[w], t.lane);
                r.z = transpose32(v[2].a[w], t.lane);
                r.w = transpose32(v[3].a[w], t.lane);
                if (S.exp_mode && j0 + 4 == S.rec_words) r.w = fid[w * 32 + t.lane];   // the record's last word: its frame id
                if ((D[w] >> t.lane) & 1u) {
                    const unsigned long long rec = sbase + before + __popc(D[w] & ((1u << t.lane) - 1u));
                    *reinterpret_cast<uint4*>(out + (size_t)rec * S.rec_words + j0) = r;
                }
                before += __popc(D[w]);
            }
        }
    }
}

// Reload of one dense group of survivors: frame-major spill records of `point`, records [gi*W, gi*W + W), back into the
// bit-sliced state `x` (32x32 bit transposes).  Team-cooperative (every warp of `t` takes slices of the record).
// Shared by the staged kernel and the producers of the overlapped kernel (e1_overlap.cuh).
template <int WORDS>
__device__ __forceinline__ void stage_reload_group(const E1Params& P, const StageParams& S, const TeamCtx& t, VecT<WORDS>* x,
                                                   uint32_t point, unsigned long long gi, const uint32_t (&valid)[WORDS]) {
    constexpr int W = 32 * WORDS;
    const uint32_t* recs = S.spill_in + ((size_t)point * S.point_stride + gi * W) * S.rec_words;
    for (int j0 = t.warp * 4; j0 < S.rec_words; j0 += t.team_warps * 4) {
        VecT<WORDS> o[4];   // state vectors of variables s0, s0+32, s0+64, s0+96: whole-vector stores (no bank conflicts)
        LDPC_FORW {
            uint4 r = make_uint4(0, 0, 0, 0);
            if ((valid[w] >> t.lane) & 1u)
                r = *reinterpret_cast<const uint4*>(recs + (size_t)(w * 32 + t.lane) * S.rec_words + j0);
            if (valid[w] != 0u) {   // team-uniform
                o[0].a[w] = transpose32(r.x, t.lane); o[1].a[w] = transpose32(r.y, t.lane);
                o[2].a[w] = transpose32(r.z, t.lane); o[3].a[w] = transpose32(r.w, t.lane);
            } else {
                o[0].a[w] = 0u; o[1].a[w] = 0u; o[2].a[w] = 0u; o[3].a[w] = 0u;
            }
        }
        const int s0 = 32 * j0 + t.lane;
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (s0 + 32 * k < P.n) x[s0 + 32 * k] = o[k];
    }
}

template <int WORDS, int DV, int DC, bool TM = false>
__global__ void __launch_bounds__(e1_threads(WORDS), 1)
e1_mc_staged_kernel(const __grid_constant__ E1Params P, const __grid_constant__ StageParams S) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint32_t tm_slot;
    uint16_t* s_c2v = reinterpret_cast<uint16_t*>(smem + P.sm_c2v);
    uint16_t* s_v2c = reinterpret_cast<uint16_t*>(smem + P.sm_v2c);
    if (!TM) {
        const int dc = DC ? DC : P.dc, dv = DV ? DV : P.dv;
        const int nc = dc * P.m_pad / 8, nv = dv * P.n_pad / 8;
        const uint4* g0 = reinterpret_cast<const uint4*>(P.c2v);
        const uint4* g1 = reinterpret_cast<const uint4*>(P.v2c);
        uint4* d0 = reinterpret_cast<uint4*>(s_c2v);
        uint4* d1 = reinterpret_cast<uint4*>(s_v2c);
        for (int i = threadIdx.x; i < nc; i += blockDim.x) d0[i] = __ldg(g0 + i);
        if (!P.v2c_global) for (int i = threadIdx.x; i < nv; i += blockDim.x) d1[i] = __ldg(g1 + i);
        stage_v2c_resident(P, smem);
    }
    __syncthreads();

    TeamCtx t;
    t.team_warps = P.team_warps;
    t.nthr = P.team_warps * 32;
    const int team = threadIdx.x / t.nthr;
    t.tid = threadIdx.x - team * t.nthr;
    t.lane = threadIdx.x & 31;
    t.warp = t.tid >> 5;
    t.bar_id = 1 + team;
    unsigned char* base = smem + P.sm_team0 + (size_t)team * P.sm_team_stride;
    VecT<WORDS>* x = reinterpret_cast<VecT<WORDS>*>(base + P.sm_x);
    VecT<WORDS>* two = reinterpret_cast<VecT<WORDS>*>(base + P.sm_t);
    t.sa_x = (uint32_t)__cvta_generic_to_shared(x); t.sa_t = (uint32_t)__cvta_generic_to_shared(two); t.sm = smem;
    uint32_t* xs = reinterpret_cast<uint32_t*>(x);
    uint32_t* hbuf = reinterpret_cast<uint32_t*>(base + P.sm_hbuf);
    t.scr = reinterpret_cast<uint32_t*>(base + P.sm_scr);
    uint32_t* work_slot = t.scr + 32 * (2 * WORDS + 1);
    uint32_t* fid = work_slot + 2 + 32 * WORDS + 8;   // chunk-relative frame id per lane (expurgated runs)
    uint32_t tm_base = 0;
    if constexpr (TM) {   // socket lists -> tensor memory (e1_tmem.cuh)
        tm_base = tm_alloc_all(&tm_slot);
        tm_fill_lists<WORDS, DV, DC>(P, t, tm_base);
    }

    constexpr int W = 32 * WORDS;
    const int stride = raw_stride(P.cap);
    const int nst1 = S.nstages + 1;

    unsigned long long total = 0;
    if (S.stage == 0) total = (unsigned long long)P.npoints * P.groups_per_point;
    else for (int p = 0; p < P.npoints; p++) total += (S.cnt_in[p] + W - 1) / W;

    for (;;) {
        unsigned long long item;
        if (t.team_warps == 1) {
            unsigned long long v = 0;
            if (t.lane == 0) v = atomicAdd(P.work_counter, 1ull);
            item = __shfl_sync(0xFFFFFFFFu, v, 0);
        } else {
            if (t.tid == 0) {
                unsigned long long v = atomicAdd(P.work_counter, 1ull);
                work_slot[0] = (uint32_t)v; work_slot[1] = (uint32_t)(v >> 32);
            }
            team_sync(t);
            item = (unsigned long long)work_slot[0] | ((unsigned long long)work_slot[1] << 32);
            team_sync(t);   // work_slot is reused for the spill base below
        }
        if (item >= total) break;

        uint32_t point;
        unsigned long long gi;
        uint32_t valid[WORDS];
        if (S.stage == 0) {
            point = (uint32_t)(item / P.groups_per_point);
            gi = item - (unsigned long long)point * P.groups_per_point;
            LDPC_FORW {
                const unsigned long long f0 = (gi * WORDS + w) * 32ull;
                valid[w] = f0 >= P.num_frames ? 0u
                           : (P.num_frames - f0 >= 32ull ? 0xFFFFFFFFu : ((1u << (uint32_t)(P.num_frames - f0)) - 1u));
            }
        } else {
            unsigned long long rest = item;
            int p = 0;
            for (; p < P.npoints; p++) {
                const unsigned long long gp = (S.cnt_in[p] + W - 1) / W;
                if (rest < gp) break;
                rest -= gp;
            }
            point = (uint32_t)p;
            gi = rest;
            const unsigned long long left = S.cnt_in[p] - gi * W;
            LDPC_FORW {
                const unsigned long long f0 = (unsigned long long)w * 32ull;
                valid[w] = f0 >= left ? 0u : (left - f0 >= 32ull ? 0xFFFFFFFFu : ((1u << (uint32_t)(left - f0)) - 1u));
            }
        }
        const int lo = S.cuts[point * nst1 + S.stage], hi = S.cuts[point * nst1 + S.stage + 1];
        unsigned long long* C = P.counters + (size_t)point * stride;

        uint32_t any[WORDS], cnt = 0;
        // ---- initial state of the group
        if (S.stage == 0) {
            const BecParams bp = P.bec[point];
            const unsigned long long q0 = P.first_word + gi * WORDS;
            uint32_t red[WORDS];
            team_draw<WORDS, false, TM ? WORDS : 1, TM>(P, t, bp, point + P.point0, q0, valid, valid, x, reinterpret_cast<uint32_t*>(two),
                                    t.scr + 32 * (2 * WORDS + 1) + 2 + 32 * WORDS, red, cnt);
            if (S.exp_mode)
                for (int f = t.tid; f < W; f += t.nthr) fid[f] = (uint32_t)(gi * W + f);
            team_reduce<WORDS>(t, red, cnt);
            LDPC_FORW any[w] = red[w];
            if (t.tid == 0) hbuf[0] = cnt;
        } else {
            const uint32_t* recs = S.spill_in + ((size_t)point * S.point_stride + gi * W) * S.rec_words;
            if (S.exp_mode)
                for (int f = t.tid; f < W; f += t.nthr)
                    fid[f] = (unsigned long long)f < S.cnt_in[point] - gi * W ? recs[(size_t)f * S.rec_words + S.rec_words - 1] : 0u;
            stage_reload_group<WORDS>(P, S, t, x, point, gi, valid);
            LDPC_FORW any[w] = valid[w];
            team_sync(t);
        }

        // ---- iterations [lo, hi) in lock step (message_passing.c:15-78 per frame)
        uint32_t running[WORDS];
        LDPC_FORW running[w] = valid[w];
        unsigned long long exec_sum = 0;
        int age = lo;
        for (; age < hi; age++) {
            uint32_t r = 0u, pc = 0u;
            LDPC_FORW { r |= running[w]; pc += __popc(running[w]); }
            if (r == 0u) break;
            exec_sum += pc;
            uint32_t oc[2 * WORDS], a2[WORDS], c2[WORDS];
            if constexpr (TM) {
                check_phase_tm<WORDS, DC>(P, t);
                team_sync(t);
                var_phase_tm<WORDS, DV, false>(P, t, valid, a2, c2, cnt);
            } else {
                check_phase<WORDS, DV, DC, false>(P, t, s_c2v, x, two, nullptr, nullptr);
                team_sync(t);
                mc_var_phase<WORDS, DV, DC, false>(P, t, s_v2c, reinterpret_cast<const uint16_t*>(smem + P.sm_vres), x, two, valid, a2, c2, cnt);
            }
            LDPC_FORW { oc[w] = a2[w]; oc[WORDS + w] = c2[w]; }
            team_reduce<2 * WORDS>(t, oc, cnt);
            if (t.tid == 0) hbuf[age - lo + 1] = cnt;
            uint32_t stopped = 0;
            LDPC_FORW {
                any[w] = oc[w];
                const uint32_t nr = running[w] & oc[w] & (age == 0 ? 0xFFFFFFFFu : oc[WORDS + w]);   // :16-19, :76-78
                stopped += __popc(running[w] & ~nr);
                running[w] = nr;
            }
            if (S.stop_hist && t.tid == 0) red_add(S.stop_hist + (size_t)point * (P.cap + 1) + age + 1, stopped);
        }
        stage_finish_group<WORDS>(P, S, t, x, hbuf, work_slot, point, lo, hi, age, valid, running, any, cnt, exec_sum, fid);
        team_sync(t);   // shared state is rewritten by the next group
    }
    if constexpr (TM) tm_free_all(tm_base);
}

}  // namespace ldpc
