// e1_overlap.cuh -- the staged E1 Monte-Carlo pipeline with the channel draw (stage 0) and the survivor reload (later
// stages) OVERLAPPED with the decoding (warp-specialised producer / consumer CTA).
//
// Why: in the staged kernel (e1_staged.cuh) a team first draws its group's erasure pattern (channels.py:24-26;
// Philox + bit-sliced comparison: ALU work only, the shared-memory pipe idles) and then decodes it
// (message_passing.c:24-65; gathers: the shared-memory pipe is saturated, the ALU half idle).  When one team owns the
// SM (n = 10^4: state 120 KB) nothing overlaps the two: ncu (profiles/r2a_*) attributes 22 % of the stage-0 warp
// samples to team_draw with 4 % of the shared-memory wavefronts, while check_phase_tm + var_phase_tm run the pipe at
// 95 %.  With the socket lists in tensor memory (e1_tmem.cuh) 100 KB of shared memory are free, so:
//
//   * warps [0, NC)        "consumers": one team that decodes group after group, exactly the code of the staged kernel;
//   * warps [NC, NC + NP)  "producers": draw the NEXT group's pattern into a staging buffer (their own queue, scratch
//                          and named barrier), reduce its initial erasure count / any-erased masks and post them with
//                          the work item in a mailbox;
//   * hand-off through two named barriers (PTX bar.arrive / bar.sync, the producer-consumer pattern of the ISA
//     manual): FULL  -- producers arrive after the pattern is written, consumers wait;
//              EMPTY -- every consumer arrives after it has copied its part of the pattern into x[] (and read the
//                       mailbox), producers wait before drawing again.
//     The copy costs n/8 bytes per frame through the shared-memory pipe once per group (< 1 % of a group's traffic)
//     and frees the staging buffer at once, so a single buffer is enough and the gather addresses kept in tensor
//     memory stay absolute.
//
// Results are bit-identical to e1_mc_staged_kernel (same generator keys, same decoder, same epilogue).
#pragma once
#include "e1_staged.cuh"

namespace ldpc {

constexpr int OV_BAR_CONS = 1, OV_BAR_PROD = 2, OV_BAR_FULL = 3, OV_BAR_EMPTY = 4;

__device__ __forceinline__ void ov_bar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
// clock read that stays where it is written (statistics around the barrier waits)
__device__ __forceinline__ uint32_t ov_clock() {
    uint32_t c;
    asm volatile("mov.u32 %0, %%clock;" : "=r"(c)::"memory");
    return c;
}
__device__ __forceinline__ void ov_bar_arrive(int id, int count) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// mailbox (32-bit words): [0] end flag, [1] point, [2..3] group index, [4] initial erased count, [5 .. 5+WORDS) any-erased masks

// ---- channel draw of the producer warps (channels.py:24-26), warp-autonomous ------------------------------------
// Same generator as bec_word / team_draw, bit for bit (every word takes Philox blocks 0, 1, 2, ... until all its frames
// are decided), scheduled for a few warps that run NEXT to a team saturating the shared-memory pipe: a round trip
// through that pipe costs ~600 cycles there, so team_draw's team-wide queues (atomic slot counters, a barrier per
// round, read-modify-write of the pattern words) serialise on it.  Here a warp owns its variables and a small private
// queue: unfinished words are appended with positions from a ballot (stores only, no atomics, no barriers); whenever
// 32 are waiting the warp reads them back in ONE round trip and gives each another block on all 32 lanes; late bits
// reach the pattern with a fire-and-forget shared-memory atomic OR.  2.15 blocks per word as in team_draw.
constexpr int OV_WQ = 96;   // private queue entries per producer warp (12 bytes each)
__host__ __device__ constexpr uint32_t ov_wq_bytes() { return OV_WQ * 12u; }

template <int WORDS>
__device__ __forceinline__ void warp_draw(const E1Params& P, int lane, int warp, int nwarps, const BecParams& bp_, uint32_t point,
                                          unsigned long long q0, const uint32_t (&valid)[WORDS], VecT<WORDS>* stg,
                                          unsigned char* wq_raw, uint32_t (&red)[WORDS], uint32_t& cnt) {
    LDPC_FORW red[w] = 0u;
    cnt = 0u;
    BecParams bp = bp_;
    bp.seed_lo = P.seed_lo; bp.seed_hi = P.seed_hi;   // same values, but provably warp-uniform (kernel parameters)
    const int nthr = nwarps * 32, n = P.n;
    if (bp.mode != 1) {   // eps == 0 or eps == 1
        for (int i = warp * 32 + lane; i < n; i += nthr) {
            VecT<WORDS> xi;
            LDPC_FORW { xi.a[w] = bp.mode == 2 ? valid[w] : 0u; red[w] |= xi.a[w]; cnt += __popc(xi.a[w]); }
            stg[i] = xi;
        }
        return;
    }
    uint32_t* xs = reinterpret_cast<uint32_t*>(stg);
    uint2* wq = reinterpret_cast<uint2*>(wq_raw);                       // (word index | next block << 24, undecided frames)
    uint32_t* wqv = reinterpret_cast<uint32_t*>(wq_raw + OV_WQ * 8);    // original variable id of the entry
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t qn = 0;   // entries waiting (warp-uniform)
    const BecMasks8 M = bec_masks8(bp);
    const bool deeper = (bp.threshold << 8) != 0u;

    // append the lanes with need == true; a full queue makes the lane finish its word itself (cold)
    auto push = [&](bool need, uint32_t idx, int b, uint32_t var, uint32_t eq, uint32_t& lt_extra) {
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, need);
        if (m == 0u) return;
        const uint32_t pos = qn + __popc(m & lt_mask);
        if (need) {
            if (pos < (uint32_t)OV_WQ) { wq[pos] = make_uint2(idx | ((uint32_t)b << 24), eq); wqv[pos] = var; }
            else lt_extra |= bec_finish(bp, point, q0 + idx % WORDS, var, b, eq);
        }
        qn = min(qn + (uint32_t)__popc(m), (uint32_t)OV_WQ);
    };
    // one more block for up to 32 waiting words (taken from the top of the queue)
    auto drain = [&](bool all) {
        while (qn >= 32u || (all && qn > 0u)) {
            const uint32_t take = min(qn, 32u), base = qn - take;
            __syncwarp();
            const bool have = (uint32_t)lane < take;
            uint2 ent = make_uint2(0u, 0u);
            uint32_t var = 0u;
            if (have) { ent = wq[base + lane]; var = wqv[base + lane]; }
            qn = base;
            __syncwarp();
            const uint32_t idx = ent.x & 0xFFFFFFu;
            const int b = (int)(ent.x >> 24);
            const uint32_t w = idx % WORDS;
            uint32_t lt = 0u, eq = ent.y;
            bool more = false;
            if (have) {
                bec_block_rk(bp, P.rk, point, q0 + w, var, b, lt, eq);
                more = !bec_done_after(bp, b, eq);
            }
            uint32_t extra = 0u;
            push(more, idx, b + 1, var, eq, extra);
            lt |= extra;
            if (lt != 0u) {   // a subset of the entry's undecided frames, which were masked by `valid`
                atomicOr(&xs[idx], lt);
#pragma unroll
                for (int k = 0; k < WORDS; k++) {
                    const uint32_t l = (w == (uint32_t)k) ? lt : 0u;
                    red[k] |= l;
                    cnt += __popc(l);
                }
            }
        }
    };

    // the original id of the NEXT variable is fetched (L2) one trip ahead
    int i = warp * 32 + lane;
    uint32_t var = 0u;
    if (i < n) var = P.var_orig ? (uint32_t)P.var_orig[i] : (uint32_t)i;
    for (int base = warp * 32; base < n; base += nthr, i += nthr) {
        const bool act = i < n;
        uint32_t var_next = 0u;
        if (i + nthr < n) var_next = P.var_orig ? (uint32_t)P.var_orig[i + nthr] : (uint32_t)(i + nthr);
        uint32_t lt[WORDS], eq[WORDS];
        LDPC_FORW {
            bec_first8_rk(M, P.rk, point, q0 + w, var, lt[w], eq[w]);
            eq[w] &= act ? valid[w] : 0u;
        }
        VecT<WORDS> xi;
        LDPC_FORW {
            uint32_t extra = 0u;
            push(eq[w] != 0u && deeper, (uint32_t)i * WORDS + w, 2, var, eq[w], extra);
            xi.a[w] = (lt[w] | extra) & valid[w];
            if (act) { red[w] |= xi.a[w]; cnt += __popc(xi.a[w]); }
        }
        if (act) stg[i] = xi;
        drain(false);
        var = var_next;
    }
    drain(true);
}

// The producer loop reads the kernel parameters in place (constant bank): the run seed must stay a kernel parameter for
// the Philox round keys to live in uniform registers -- a by-value copy of the fields (tried, also as a non-inlined
// function) put 20 round keys into vector registers and spilled inside the rounds.
template <int WORDS>
__device__ __forceinline__ void ov_producer_loop(const E1Params& P, const StageParams& S, unsigned char* smem, int NC, int NP) {
    TeamCtx t;
    t.team_warps = NP;
    t.nthr = NP * 32;
    t.tid = (int)threadIdx.x - NC * 32;
    t.lane = threadIdx.x & 31;
    t.warp = t.tid >> 5;
    const int all = (NC + NP) * 32;
    constexpr int W = 32 * WORDS;
    unsigned long long total = 0;   // stage 0: every group of every point; later stages: dense groups of the spilled survivors
    if (S.stage == 0) total = (unsigned long long)P.npoints * P.groups_per_point;
    else for (int p = 0; p < P.npoints; p++) total += (S.cnt_in[p] + W - 1) / W;
    t.bar_id = OV_BAR_PROD;
    t.sm = smem; t.sa_x = 0; t.sa_t = 0; t.ta_c = 0; t.ta_v = 0;
    t.scr = reinterpret_cast<uint32_t*>(smem + P.sm_pscr);
    uint32_t* work_slot = t.scr + 32 * (2 * WORDS + 1);
    VecT<WORDS>* stg = reinterpret_cast<VecT<WORDS>*>(smem + P.sm_stage);
    volatile uint32_t* mbox = reinterpret_cast<volatile uint32_t*>(smem + P.sm_mbox);
    bool first = true;
    uint32_t pw = 0;   // time spent waiting for the staging buffer, in units of 64 cycles (statistics)
    const long long t_begin = clock64();
    for (;;) {
        if (t.tid == 0) {
            const unsigned long long v = atomicAdd(P.work_counter, 1ull);
            work_slot[0] = (uint32_t)v; work_slot[1] = (uint32_t)(v >> 32);
        }
        team_sync(t);
        const unsigned long long item = (unsigned long long)work_slot[0] | ((unsigned long long)work_slot[1] << 32);
        if (!first) {   // the consumers have taken the previous pattern and mailbox
            const uint32_t c0 = ov_clock();
            ov_bar_sync(OV_BAR_EMPTY, all);
            pw += (ov_clock() - c0) >> 6;
        }
        first = false;
        if (item >= total) {
            if (t.tid == 0) mbox[0] = 1u;
            ov_bar_arrive(OV_BAR_FULL, all);
            break;
        }
        uint32_t point;
        unsigned long long gi;
        uint32_t valid[WORDS], red[WORDS], cnt = 0u;
        if (S.stage == 0) {
            point = (uint32_t)(item / P.groups_per_point);
            gi = item - (unsigned long long)point * P.groups_per_point;
            LDPC_FORW {
                const unsigned long long f0 = (gi * WORDS + w) * 32ull;
                valid[w] = f0 >= P.num_frames ? 0u
                           : (P.num_frames - f0 >= 32ull ? 0xFFFFFFFFu : ((1u << (uint32_t)(P.num_frames - f0)) - 1u));
            }
            const BecParams bp = P.bec[point];
            warp_draw<WORDS>(P, t.lane, t.warp, NP, bp, point + P.point0, P.first_word + gi * WORDS, valid, stg,
                             smem + P.sm_q + (uint32_t)t.warp * ov_wq_bytes(), red, cnt);
        } else {
            // later stages: the producers RELOAD the next dense group of survivors (frame-major records -> bit-sliced
            // state through 32x32 transposes) while the consumers decode the current one
            unsigned long long rest = item;
            int p = 0;
            for (; p < P.npoints; p++) {
                const unsigned long long gp = (S.cnt_in[p] + W - 1) / W;
                if (rest < gp) break;
                rest -= gp;
            }
            point = (uint32_t)p; gi = rest;
            const unsigned long long left = S.cnt_in[p] - gi * W;
            LDPC_FORW {
                const unsigned long long f0 = (unsigned long long)w * 32ull;
                valid[w] = f0 >= left ? 0u : (left - f0 >= 32ull ? 0xFFFFFFFFu : ((1u << (uint32_t)(left - f0)) - 1u));
                red[w] = valid[w];
            }
            stage_reload_group<WORDS>(P, S, t, stg, point, gi, valid);
        }
        team_reduce<WORDS>(t, red, cnt);   // its barrier also completes the producers' pattern writes
        if (t.tid == 0) {
            mbox[0] = 0u; mbox[1] = point; mbox[2] = (uint32_t)gi; mbox[3] = (uint32_t)(gi >> 32); mbox[4] = cnt;
            LDPC_FORW mbox[5 + w] = red[w];
        }
        ov_bar_arrive(OV_BAR_FULL, all);
    }
    if (t.tid == 0 && S.stats) {
        atomicAdd(S.stats + 5, (unsigned long long)pw * 64ull);
        atomicAdd(S.stats + 6, (unsigned long long)(clock64() - t_begin) - (unsigned long long)pw * 64ull);
    }
}

// Threads of the overlapped kernel (a template parameter: the register budget follows from it).  At W = 64 the decoder
// fits 72 registers, which buys 4 more warps than the staged kernel's 768 threads x 80 registers: 20 consumer + 8
// producer warps at n = 10^4 (measured 8.20e7 frames/s against 8.16e7 with 16 + 8; 24 + 8 at 64 registers: 7.8e7).
__host__ __device__ constexpr int ov_threads(int words) { return words == 2 ? 896 : e1_threads(words); }

template <int WORDS, int DV, int DC, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
e1_mc_overlap_kernel(const __grid_constant__ E1Params P, const __grid_constant__ StageParams S) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint32_t tm_slot;
    constexpr int W = 32 * WORDS;
    const int NC = P.team_warps, NP = P.ov_np;
    const int all = (NC + NP) * 32;
    const bool consumer = (int)threadIdx.x < NC * 32;
    const unsigned long long total = (unsigned long long)P.npoints * P.groups_per_point;

    // socket lists -> tensor memory (consumer threads only own columns; the allocation is CTA-wide)
    const uint32_t tm_base = tm_alloc_all(&tm_slot);

    if (!consumer) {
        // ------------------------------------------------ producers: channel draw (channels.py:24-26) ----
        ov_producer_loop<WORDS>(P, S, smem, NC, NP);
    } else {
        // ------------------------------------------------ consumers: decode (message_passing.c:15-78) ----
        TeamCtx t;
        t.team_warps = NC;
        t.nthr = NC * 32;
        t.tid = (int)threadIdx.x;
        t.lane = threadIdx.x & 31;
        t.warp = t.tid >> 5;
        t.bar_id = OV_BAR_CONS;
        unsigned char* base = smem + P.sm_team0;
        VecT<WORDS>* x = reinterpret_cast<VecT<WORDS>*>(base + P.sm_x);
        VecT<WORDS>* two = reinterpret_cast<VecT<WORDS>*>(base + P.sm_t);
        const VecT<WORDS>* stg = reinterpret_cast<const VecT<WORDS>*>(smem + P.sm_stage);
        volatile uint32_t* mbox = reinterpret_cast<volatile uint32_t*>(smem + P.sm_mbox);
        t.sa_x = (uint32_t)__cvta_generic_to_shared(x); t.sa_t = (uint32_t)__cvta_generic_to_shared(two); t.sm = smem;
        uint32_t* hbuf = reinterpret_cast<uint32_t*>(base + P.sm_hbuf);
        t.scr = reinterpret_cast<uint32_t*>(base + P.sm_scr);
        uint32_t* work_slot = t.scr + 32 * (2 * WORDS + 1);
        uint32_t* fid = work_slot + 2 + 32 * WORDS + 8;
        tm_fill_lists<WORDS, DV, DC>(P, t, tm_base);
        const int nst1 = S.nstages + 1;
        const int nvec4 = (int)(((size_t)P.n * sizeof(VecT<WORDS>) + 15) / 16);
        const uint4* s4 = reinterpret_cast<const uint4*>(stg);
        uint4* x4 = reinterpret_cast<uint4*>(x);
        uint32_t cw = 0;   // time spent waiting for a pattern, in units of 64 cycles (statistics)
        const long long t_begin = clock64();
        for (;;) {
            {
                const uint32_t c0 = ov_clock();
                ov_bar_sync(OV_BAR_FULL, all);
                cw += (ov_clock() - c0) >> 6;
            }
            if (mbox[0] != 0u) break;
            const uint32_t point = mbox[1];
            const unsigned long long gi = (unsigned long long)mbox[2] | ((unsigned long long)mbox[3] << 32);
            uint32_t cnt = mbox[4];
            uint32_t any[WORDS], valid[WORDS];
            LDPC_FORW {
                any[w] = mbox[5 + w];
                if (S.stage == 0) {
                    const unsigned long long f0 = (gi * WORDS + w) * 32ull;
                    valid[w] = f0 >= P.num_frames ? 0u
                               : (P.num_frames - f0 >= 32ull ? 0xFFFFFFFFu : ((1u << (uint32_t)(P.num_frames - f0)) - 1u));
                } else valid[w] = any[w];   // survivors: all of them still hold erasures
            }
            for (int i = t.tid; i < nvec4; i += t.nthr) x4[i] = s4[i];
            ov_bar_arrive(OV_BAR_EMPTY, all);   // this thread's part of the pattern and the mailbox are taken
            if (S.exp_mode) {
                if (S.stage == 0) {
                    for (int f = t.tid; f < W; f += t.nthr) fid[f] = (uint32_t)(gi * W + f);
                } else {   // the frame id travels in the last word of the spill record
                    const uint32_t* recs = S.spill_in + ((size_t)point * S.point_stride + gi * W) * S.rec_words;
                    for (int f = t.tid; f < W; f += t.nthr)
                        fid[f] = (unsigned long long)f < S.cnt_in[point] - gi * W ? recs[(size_t)f * S.rec_words + S.rec_words - 1] : 0u;
                }
            }
            if (t.tid == 0) hbuf[0] = cnt;
            team_sync(t);   // x complete

            const int lo = S.cuts[point * nst1 + S.stage], hi = S.cuts[point * nst1 + S.stage + 1];
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
                check_phase_tm<WORDS, DC>(P, t);
                team_sync(t);
                var_phase_tm<WORDS, DV, false>(P, t, valid, a2, c2, cnt);
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
        if (t.tid == 0 && S.stats) {
            atomicAdd(S.stats + 4, (unsigned long long)cw * 64ull);
            atomicAdd(S.stats + 7, (unsigned long long)(clock64() - t_begin) - (unsigned long long)cw * 64ull);
        }
    }
    tm_free_all(tm_base);
}

}  // namespace ldpc
