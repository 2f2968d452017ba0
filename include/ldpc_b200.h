/*
 * ldpc_b200.h -- C ABI of libldpc_b200.so: the B200 (sm_100a) implementation of the BEC
 * message-passing Monte-Carlo hot path of roryhighnam/iib_project_ldpc_codes.
 *
 * Plain C, plain pointers and sizes; no torch / CUDA types in any signature.  All entry points are
 * GPU-only: there is no CPU fallback.  A call that cannot run on the GPU returns a negative status
 * and ldpc_b200_last_error() says why.
 *
 * Reference interfaces replaced (file:line relative to the reference repository):
 *   message_passing            message_passing.c:7        (called from parallel_simulator.py:162-164,
 *                                                          parallel_simulator_expurgated.py:162-164)
 *   generate_random_code       random_code_generator.c:21 (called from parallel_simulator_expurgated.py:223,
 *                                                          peeling_decoder.py:42)
 *   ldpc_b200_decode_batch     a loop of message_passing.c:7 calls over frames, i.e. the body of
 *                              regular_LDPC_code.message_pass_decode (parallel_simulator.py:131-166)
 *   ldpc_b200_mc_run / _launch the frame loops run_simulation (parallel_simulator_expurgated.py:200-256)
 *                              and run_simulation_fixed_ldpc (:367-392) including the channel draw
 *                              (channels.py:24-26) and the expurgation predicate (:238)
 */
#ifndef LDPC_B200_H
#define LDPC_B200_H

#include <stdint.h>
#include <stdbool.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LDPC_B200_OK            0
#define LDPC_B200_EINVAL      (-1)  /* bad argument / malformed graph / symbol outside {0,1,2} */
#define LDPC_B200_ECUDA       (-2)  /* CUDA runtime failure (no device, launch error, out of memory) */
#define LDPC_B200_EUNSUPPORTED (-3) /* configuration not supported by this build */
#define LDPC_B200_ENOMEM      (-4)  /* out of host memory (no C++ exception ever crosses this ABI) */

typedef struct ldpc_b200_graph ldpc_b200_graph;

typedef struct ldpc_b200_mc_stats {
    uint64_t phys_frame_iterations; /* frame-iterations the kernel physically ran (lock-step groups)       */
    uint64_t second_pass_groups;    /* frame groups re-decoded for the expurgation filter                  */
    int32_t engine;                 /* 1 = shared-memory resident (E1), 2 = HBM/L2 streamed (E2)          */
    int32_t words;                  /* 32-bit words per node and frame group (frames per group = 32*words)*/
    int32_t teams;                  /* frame groups resident per CTA                                       */
    int32_t team_warps;             /* warps cooperating on one frame group                                */
    int32_t grid;                   /* CTAs launched                                                       */
    int32_t smem_bytes;             /* dynamic shared memory per CTA                                       */
    int32_t stages;                 /* stages of the last launch (survivor re-packing, 1 = unstaged)       */
    int32_t kernel_launches;        /* kernels launched since the previous statistics read                 */
    uint64_t spilled_records;       /* frames written to (and read back from) the HBM spill buffers between stages */
    uint64_t recounted_frames;      /* expurgated runs: frames listed by pass 1 and re-drawn by the re-count pass  */
    int32_t record_bytes;           /* bytes per spill record: DRAM traffic = spilled_records * record_bytes * 2   */
    int32_t tmem_lists;             /* 1 = socket lists in tensor memory (tcgen05.ld-supplied gather addresses)    */
    int32_t overlap_producers;      /* > 0: stage 0 ran as the producer/consumer kernel with this many channel-draw warps */
    int32_t reserved_;
    uint64_t overlap_cycles[4];     /* overlapped stage 0, SM cycles summed over CTAs: consumers waiting for a pattern,
                                       producers waiting for the staging buffer, producers drawing, consumers decoding */
} ldpc_b200_mc_stats;

/* ---- reference drop-ins ------------------------------------------------------------------- */

/* message_passing.c:7 -- identical arguments, in-place outputs and return value.
 * Mvc[n] in {0,1,2} (2 = erased) is decoded in place; errors[it] += erased count after iteration it
 * (caller zeroes, as parallel_simulator.py:147 does); returns the iteration index at which no erasure
 * was left, else `iterations`.  Runs on the GPU (device 0 unless LDPC_B200_DEVICE is set); device
 * copies of the graph are cached by content.  Returns a negative LDPC_B200_E* on failure (no valid
 * reference result is negative). */
int message_passing(int *Mvc, int iterations, int *variable_to_check_list, int *check_to_variable_list,
                    int *errors, int n, int k, int dv, int dc);

/* random_code_generator.c:21 -- same arguments and outputs (host code; the graph is built on the CPU
 * in the reference too).  Differences, all supersets: parity_check may be NULL (the dense (n-k) x n
 * matrix is impossible at n = 10^6), variable_lookup is built sparsely, and retries are iterative. */
int generate_random_code(int n, int dv, int dc, int *variable_lookup, int *check_lookup, int *sequence,
                         bool *parity_check, int iterations, bool first_run, int seed);

/* ---- library ------------------------------------------------------------------------------ */

const char *ldpc_b200_last_error(void);
int ldpc_b200_device_count(void);

/* Seeded, sparse generator with the (variable_lookup, check_lookup) contract of
 * random_code_generator.c:34-36,53-64: socket permutation + rejection of graphs with a repeated
 * variable in a check; check_lookup unsorted within a check, variable_lookup in ascending check order.
 * mode 0: pure rejection (the reference's distribution); mode 1: repair repeated variables by socket
 * swaps (needed for (4,8) and large n where rejection takes > 10^4 shuffles).  Returns 1 on success. */
int ldpc_b200_generate_graph(int n, int dv, int dc, uint64_t seed, int mode, int *variable_lookup, int *check_lookup);

/* Graph `index` of the in-kernel ensemble generator (gen_mode 2 of ldpc_b200_mc_run_ensemble) for `graph_seed`:
 * Philox keys -> sort -> listed repeats repaired, bit-identical to what the kernel builds in shared memory.
 * n*dv must be < 65536. */
int ldpc_b200_generate_graph_indexed(int n, int dv, int dc, uint64_t graph_seed, uint64_t index,
                                     int *variable_lookup, int *check_lookup);

/* Upload a Tanner graph: variable_to_check_list [n][dv], check_to_variable_list [n-k][dc], int32,
 * the layouts of message_passing.c:7.  k must equal n - n*dv/dc. */
int ldpc_b200_graph_create(const int *variable_to_check_list, const int *check_to_variable_list,
                           int n, int k, int dv, int dc, int device, ldpc_b200_graph **out);
void ldpc_b200_graph_destroy(ldpc_b200_graph *g);

/* Counter block per erasure point (uint64, additive over disjoint frame ranges and over GPUs):
 *   [0 .. iterations]  sum over COUNTED frames of [initial_count, errors[0 .. iterations-1]]
 *   [iterations+1]     block errors      [iterations+2]  bit errors
 *   [iterations+3]     frames (all)      [iterations+4]  counted frames (errors[-1] > expurgation)
 *   [iterations+5]     executed iterations summed over all frames (edge updates = this * n * dv) */
int ldpc_b200_counters_len(int iterations);

/* Fused Monte-Carlo over frames [first_frame, first_frame + num_frames) of every erasure point:
 * draw erasures on the device, decode, accumulate.  first_frame must be a multiple of 128.
 * expurgation < 0 counts every frame (parallel_simulator.py:227-231).
 * counters: HOST buffer [npoints][ldpc_b200_counters_len(iterations)], overwritten. stats may be NULL. */
int ldpc_b200_mc_run(ldpc_b200_graph *g, const double *eps, int npoints, uint64_t seed,
                     uint64_t first_frame, uint64_t num_frames, int iterations, int expurgation,
                     uint64_t *counters, ldpc_b200_mc_stats *stats);

/* Same, asynchronous on `stream` (a cudaStream_t, or NULL), ADDING into the DEVICE buffer d_counters
 * of the same shape.  Nothing is copied to the host. */
int ldpc_b200_mc_launch(ldpc_b200_graph *g, const double *eps, int npoints, uint64_t seed,
                        uint64_t first_frame, uint64_t num_frames, int iterations, int expurgation,
                        uint64_t *d_counters, void *stream);
/* The same launch for the erasure points [first_point, first_point + npoints) of a longer list: point p of this
 * call is keyed as first_point + p by the channel generator, so a run can be split by POINT across GPUs or calls
 * ((point, block) partitioning; the reference runs one OS process per point and merges the CSVs with
 * tools/combine_data.py:64-95) and still decodes exactly the frames of the unsplit run.  d_counters holds the rows of
 * THESE points. */
int ldpc_b200_mc_launch_points(ldpc_b200_graph *g, const double *eps, int npoints, uint32_t first_point, uint64_t seed,
                               uint64_t first_frame, uint64_t num_frames, int iterations, int expurgation,
                               uint64_t *d_counters, void *stream);
/* statistics of the launches since the last call (call after synchronising the stream) */
int ldpc_b200_mc_stats_read(ldpc_b200_graph *g, ldpc_b200_mc_stats *stats);
/* device durations (ms, CUDA events on the launch stream) of the decode kernels launched since the
 * last call; returns how many were written (<= max) */
int ldpc_b200_kernel_times(ldpc_b200_graph *g, float *ms, int max);

/* Ensemble Monte-Carlo (run_simulation, parallel_simulator_expurgated.py:200-256): a fresh random simple
 * (dv,dc)-regular graph per `frames_per_graph` frames.  Graph G (global index first_graph + g) is
 * ldpc_b200_generate_graph(n, dv, dc, graph_seed*1000003 + G, gen_mode); its frames are the global frames
 * [G*F, G*F + frames_per_graph) with F = frame_stride (a multiple of 128, >= frames_per_graph; 0 = frames_per_graph
 * rounded up to 128), so a run is reproducible and additive over graph ranges, and a call with fewer frames per
 * graph and the same frame_stride decodes a PREFIX of every graph's frames (the drivers use that to stop exactly at
 * the 200th block error, parallel_simulator_expurgated.py:200).  counters: HOST [npoints][iterations+6],
 * overwritten.  n < 65536. */
int ldpc_b200_mc_run_ensemble(int device, int n, int dv, int dc, uint64_t graph_seed, int gen_mode,
                              uint64_t first_graph, uint64_t num_graphs, uint32_t frames_per_graph,
                              uint32_t frame_stride, const double *eps, int npoints, uint64_t seed, int iterations,
                              int expurgation, uint64_t *counters, ldpc_b200_mc_stats *stats);

/* Batched decode: Mvc HOST [frames][n] int8 in {0,1,2}, decoded in place; errors HOST
 * [frames][iterations] receives what message_passing.c leaves in a zeroed errors[]; ret [frames]. */
int ldpc_b200_decode_batch(ldpc_b200_graph *g, int8_t *Mvc, uint64_t frames, int iterations,
                           int *errors, int *ret);

/* Erasure patterns of the in-kernel generator for frames [first_frame, first_frame+frames), first_frame
 * a multiple of 32: out HOST [frames][n] int8 in {0,2}.  Computed on the GPU with the kernel's code. */
int ldpc_b200_draw_patterns(int device, int n, double eps, uint64_t seed, uint32_t point,
                            uint64_t first_frame, uint64_t frames, int8_t *out);

/* Measured shared-memory read bandwidth of the whole GPU (conflict-free 128-bit loads from every SM),
 * in GB/s: the roofline denominator of the shared-memory-resident engine. */
int ldpc_b200_smem_probe(int device, double *gbytes_per_s);

/* Shared-memory placement of the graph for the Monte-Carlo kernel (csrc/relabel.hpp): effort 0 keeps
 * the caller's node numbering, >= 1 searches a bank-conflict-free numbering + socket order (default 2,
 * or LDPC_B200_PLACE_EFFORT).  Results do not depend on it.  placement_info reports the gather
 * wavefronts per flooding iteration of the natural / chosen / conflict-free layouts. */
int ldpc_b200_set_place_effort(ldpc_b200_graph *g, int effort);
int ldpc_b200_placement_info(ldpc_b200_graph *g, int words, double *natural, double *placed, double *ideal);
/* The bank-conflict-aware placement is searched on background host threads (or read from the verified disk cache,
 * $LDPC_B200_CACHE_DIR / ~/.cache/ldpc_b200) while the first launches already decode on the natural layout -- same
 * results, ~1.5x slower kernels.  This call blocks until no search of the handle is running and every result is on
 * the device (benchmarks call it before their timed region).  source (may be NULL) receives for `words` (1, 2, 4;
 * 0 = none): 0 = no placement yet, 1 = searched in this process, 2 = read from the disk cache. */
int ldpc_b200_placement_wait(ldpc_b200_graph *g, int words, int *source);

/* Engine override: 0 automatic (shared-memory engine when the graph fits, else streamed), 1 = E1, 2 = E2. */
int ldpc_b200_set_engine(ldpc_b200_graph *g, int engine);

/* Tuning overrides (0 = automatic): words in {1,2,4}, teams per CTA, warps per team. */
int ldpc_b200_set_tuning(ldpc_b200_graph *g, int words, int teams, int team_warps);

#ifdef __cplusplus
}
#endif
#endif
