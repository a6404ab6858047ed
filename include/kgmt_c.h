/* kgmt_c.h — C ABI of libkgmt_b200.so: the KGMT tree-expansion hot path of
 * nipe1783/cudaSBMP, written from scratch for NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary.  The reference has no FFI layer: its boundary
 * is the C++ class `KGMT` (include/planners/KGMT.cuh:23-109) constructed and
 * called by demos/main.cu:30,62.  Every entry point below names the reference
 * interface it replaces.  The reference-compatible C++ facade
 * (cudasbmp_b200/include/planners/KGMT.cuh) is a thin header over this ABI, so
 * demos/main.cu compiles and runs unchanged against it.
 *
 * Conventions: plain C types only; every function returns KGMT_OK (0) or a
 * negative kgmt_status; nothing calls exit() or throws across the boundary
 * (the reference printf+exit(1)s, include/helper/helper.cuh:19-27); a context
 * is owned by the caller, bound to one CUDA device and one stream, not
 * thread-safe; distinct contexts are independent.  There is NO CPU fallback:
 * without a CUDA device kgmt_create fails with KGMT_ERR_CUDA.
 */
#ifndef KGMT_C_H
#define KGMT_C_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KGMT_ABI_VERSION 2
#define KGMT_SAMPLE_DIM 7            /* (x, y, theta, v, a, steering, duration): KGMT.cu:5, State.h:9-20 */

typedef enum kgmt_status {
    KGMT_OK = 0,
    KGMT_ERR_INVALID = -1,           /* bad argument */
    KGMT_ERR_CUDA = -2,              /* CUDA runtime error; see kgmt_last_error */
    KGMT_ERR_STATE = -3,             /* call out of order (e.g. iterate before begin) */
    KGMT_ERR_NOMEM = -4,
    KGMT_ERR_COMM = -5               /* NCCL error */
} kgmt_status;

/* why a plan stopped (KGMT.cu:118,252-259; canonical: SURVEY.md App. B #11) */
typedef enum kgmt_stop {
    KGMT_RUNNING = 0,
    KGMT_SOLVED = 1,                 /* costToGoal != 0                      KGMT.cu:252 */
    KGMT_TREE_FULL = 2,              /* treeSize >= maxTreeSize              KGMT.cu:255 */
    KGMT_ITER_LIMIT = 3,             /* itr == numIterations                 KGMT.cu:118 */
    KGMT_FRONTIER_EMPTY = 4,         /* nothing accepted: the reference spins to numIterations */
    KGMT_PEER_SOLVED = 5             /* kgmt_peer_race: another GPU of the race reached the goal first */
} kgmt_stop;

typedef enum kgmt_collide {
    KGMT_COLLIDE_GRID = 0,           /* uniform-grid culled; same flags as brute force, fewer tests */
    KGMT_COLLIDE_BRUTE = 1           /* every step bbox vs every obstacle (collisionCheck.cu:16-28), smem tiles */
} kgmt_collide;

/* The nine constructor arguments of KGMT::KGMT (KGMT.cuh:28, KGMT.cu:10-11) plus
 * what the reference hard-codes or leaves to chance. */
typedef struct kgmt_params {
    float    width, height;          /* workspace                              main.cu:20-21 */
    int      N, n;                   /* R1 grid N x N, R2 grid n x n per cell  main.cu:22-23 */
    int      num_iterations;         /*                                        main.cu:24    */
    int      max_tree_size;          /*                                        main.cu:25    */
    int      num_disc;               /* Euler steps per edge                   main.cu:26    */
    float    agent_length;           /* wheelbase L                            main.cu:27    */
    float    goal_threshold;         /* goal disc radius                       main.cu:28    */
    uint32_t seed;                   /* Philox key; iteration i uses key (seed+i, 0). Reference: time(NULL), KGMT.cu:111 */
    int      device;                 /* CUDA ordinal, -1 = current device */
    int      max_candidates;         /* candidate slots per iteration; 0 = max_tree_size (as the reference, KGMT.cu:28) */
    int      collision_mode;         /* kgmt_collide */
    int      record_candidates;      /* 1: also keep per-candidate valid/r1/r2/u3/accept arrays for export */
    int      cull_cells;             /* cull grid is cull_cells x cull_cells; 0 = choose from the obstacle set */
    int      reserved[5];            /* [0]: shared-memory staging budget for the collision data in bytes (0 = 66 KB);
                                        [1]: resident CTAs per SM of the persistent kernel (0 = all that fit);
                                        rest 0 */
    /* The car model ("systems/car.yaml" of the reference is an empty file; its dynamics and control ranges are literals
     * in statePropagator.cu:17-19).  Control c is drawn as lo + u * (hi - lo), u = curand_uniform in (0, 1]:
     *   a        = fmaf(u0, (float)(accel_max - accel_min), (float)accel_min)              :17  u0 * 10.0f - 5.0f
     *   steering = (float) fma((double)u1, steer_max - steer_min, steer_min)               :18  u1 * 2.0f * M_PI - M_PI
     *   duration = fmaf(u2, (float)(duration_max - duration_min), (float)duration_min)     :19  u2 * 1.0f + 0.05f
     * With the defaults (kgmt_default_params) these are bit for bit the reference's expressions as nvcc compiles them
     * for sm_100a (one FFMA; one DFMA on (double)(u1+u1) == 2*(double)u1; u2 * 1 + 0.05f).  Doubles because the
     * reference's steering bounds are the double constant M_PI. */
    double   accel_min, accel_max;           /* default -5, 5 */
    double   steer_min, steer_max;           /* default -M_PI, M_PI */
    double   duration_min, duration_max;     /* default (double)0.05f, (double)0.05f + 1 */
} kgmt_params;

typedef struct kgmt_iter_stats {
    int       iteration;             /* 1-based index of the step just executed */
    int       mode;                  /* 1 = 32 children/node (propagateG), 2 = floor(remaining/active) (propagateGV2), 3 = prefix */
    int       children;
    int       frontier;              /* nodes expanded */
    int       candidates;            /* edges propagated and collision-checked */
    int       accepted;              /* nodes inserted */
    int       tree_size;             /* after insertion */
    int       stop;                  /* kgmt_stop */
    float     cost_to_goal;
    int       goal_index;
} kgmt_iter_stats;

typedef struct kgmt_result {
    int       stop;                  /* kgmt_stop */
    int       iterations;
    int       tree_size;             /* KGMT::treeSize_ */
    float     cost_to_goal;          /* KGMT::costToGoal_ (0 = none) */
    int       goal_index;            /* tree index of the goal node, -1 = none */
    long long expansions;            /* candidate edges checked over the whole plan */
    float     device_ms;             /* CUDA-event time of the expansion loop (kgmt_plan_batch: of the whole launch) */
    int       kernel_launches;       /* launches of this library's kernels inside the loop */
    float     done_ms;               /* kgmt_plan_batch: device time from the start of the launch to this query's last iteration
                                        (its time-to-solution inside the batch, queueing included); otherwise = device_ms */
    float     service_ms;            /* kgmt_plan_batch: device time this query occupied its cluster; otherwise = device_ms */
} kgmt_result;

/* Array ids for kgmt_export / kgmt_import.  0..12 are the thirteen CSV dumps of
 * KGMT.cu:299-311, in that order, in the reference's own element layout
 * (AoS rows of 7 floats, int, bool-as-uint8).  Row counts: max_tree_size for
 * tree/candidate arrays, N*N for R1*, N*N*n*n for R2*. */
typedef enum kgmt_array {
    KGMT_ARR_SAMPLES = 0,            /* float [maxTree][7]   samples.csv            d_treeSamples_      */
    KGMT_ARR_UNEXPLORED = 1,         /* float [maxCand][7]   unexploredSamples.csv  d_unexploredSamples_*/
    KGMT_ARR_PARENT = 2,             /* int   [maxTree]      parentRelations.csv    d_treeParentIdx_    */
    KGMT_ARR_U_PARENT = 3,           /* int   [maxCand]      uParentIdx.csv         d_uParentIdx_       */
    KGMT_ARR_G = 4,                  /* u8    [maxTree]      G.csv                  d_G_                */
    KGMT_ARR_R2AVAIL = 5,            /* int   [N*N*n*n]      R2Avail.csv */
    KGMT_ARR_R1AVAIL = 6,            /* int   [N*N]          R1Avail.csv */
    KGMT_ARR_R1VALID = 7,
    KGMT_ARR_R2VALID = 8,
    KGMT_ARR_R1INVALID = 9,
    KGMT_ARR_R2INVALID = 10,
    KGMT_ARR_R1SCORE = 11,           /* float [N*N]          R1Score.csv */
    KGMT_ARR_R1 = 12,
    KGMT_ARR_R2 = 13,                /* int   [N*N*n*n]      d_R2_ (not dumped by the reference) */
    KGMT_ARR_COSTS = 14,             /* float [maxTree]      d_costs_ */
    /* per-candidate records of the last iteration (record_candidates = 1), [maxCand] */
    KGMT_ARR_U_VALID = 15,           /* u8    */
    KGMT_ARR_U_R1 = 16,              /* int   */
    KGMT_ARR_U_R2 = 17,              /* int   */
    KGMT_ARR_U_U3 = 18,              /* float accept uniform (KGMT.cu:395) */
    KGMT_ARR_U_ACCEPT = 19,          /* u8    GNew as decided this iteration (KGMT.cu:397) */
    KGMT_ARR_COUNT = 20
} kgmt_array;

typedef struct kgmt_ctx kgmt_ctx;

/* ---- lifetime ------------------------------------------------------------------------------ */
int  kgmt_abi_version(void);
void kgmt_default_params(kgmt_params* p);                       /* main.cu:19-28 literals + statePropagator.cu:17-19 control ranges */
/* Overlay `p` (already holding defaults or the caller's values) with the keys found in a car-model file in the
 * reference's systems/ directory (systems/car.yaml is EMPTY upstream: an empty or missing-key file changes nothing).
 * Flat "key: value" YAML subset, '#' comments, optional one-level nesting ("controls:" / "planner:" blocks are
 * flattened).  Keys: wheelbase | agent_length, num_disc, accel_min, accel_max, steer_min, steer_max, duration_min,
 * duration_max, width, height, N, n, num_iterations, max_tree_size, goal_threshold, seed.  Unknown keys are an error
 * (KGMT_ERR_INVALID; *bad_line receives the 1-based line number when not NULL). */
int  kgmt_params_from_yaml(const char* path, kgmt_params* p, int* bad_line);
int  kgmt_create(const kgmt_params* p, kgmt_ctx** out);         /* replaces KGMT::KGMT, KGMT.cu:10-78 */
void kgmt_destroy(kgmt_ctx* ctx);                               /* replaces ~KGMT + KGMT.cu:314-316 */
const char* kgmt_last_error(const kgmt_ctx* ctx);               /* replaces CUDA_ERROR_CHECK's printf, helper.cuh:19-27 */
int  kgmt_set_seed(kgmt_ctx* ctx, uint32_t seed);               /* Philox key of the next plan (reference: time(NULL), KGMT.cu:111) */
int  kgmt_reset(kgmt_ctx* ctx);                                 /* re-plan without re-allocating (reference: single-shot, App. B #12) */

/* ---- obstacles: float[K][4] = (minx, miny, maxx, maxy), obstacles.csv:1-5 ------------------ */
int  kgmt_set_obstacles(kgmt_ctx* ctx, const float* d_aabb, int K);      /* DEVICE pointer, as plan()'s d_obstacles (KGMT.cuh:31, main.cu:60-62); copied */
int  kgmt_set_obstacles_host(kgmt_ctx* ctx, const float* h_aabb, int K); /* HOST pointer (what readObstaclesFromCSV returns, helper.cu:11-34);
                                                                            asynchronous: staged in pinned memory, the cull grid is built on the
                                                                            device behind the upload, nothing waits before the next plan */

/* ---- planning ------------------------------------------------------------------------------ */
/* KGMT::plan (KGMT.cuh:31, KGMT.cu:80-317) without the CSV dump: initial/goal are HOST float[7].  Returns when the planner's
 * scalars are on the host; the back-trace of the goal path runs behind it on the same stream (kgmt_extract_path below). */
int  kgmt_plan(kgmt_ctx* ctx, const float* initial7, const float* goal7, kgmt_result* out);
/* the same, split: root insertion (KGMT.cu:85-114) then one while-loop body (KGMT.cu:118-259) per call */
int  kgmt_begin(kgmt_ctx* ctx, const float* initial7, const float* goal7);
int  kgmt_expand_iteration(kgmt_ctx* ctx, kgmt_iter_stats* out);
/* up to `count` loop bodies in one launch (stops early when the planner stops); `out` = the last one executed */
int  kgmt_expand_iterations(kgmt_ctx* ctx, int count, kgmt_iter_stats* out);
int  kgmt_get_result(kgmt_ctx* ctx, kgmt_result* out);
/* back-trace of the parent links from the goal node (node < 0) or any node to the root: rows of 7 floats,
 * root first.  Returns the path length (may exceed max_rows; only max_rows are written).  For the goal node of the
 * kgmt_plan that has just returned (paths up to 128 nodes) the rows are already on their way: an event wait and a host
 * copy, no kernel. */
int  kgmt_extract_path(kgmt_ctx* ctx, int node, float* h_rows7, int max_rows);

/* ---- batched planning (BASELINE config 4) ----------------------------------------------------------------------
 * Q independent queries (HOST rows of 7 floats; seed per query) on the context's map and parameters in ONE launch: a
 * thread-block cluster of cluster_size CTAs (1, 2, 4 or 8) plans one query at a time and pulls the next from a ticket.
 * Each query gives exactly the result kgmt_plan gives for the same (init, goal, seed).  out[Q]; h_paths7 (optional)
 * receives [Q][max_path][7] solution rows, root first, h_path_len[Q] their lengths.  Returns the number of
 * concurrent workspaces used (> 0) or a negative status.  cluster_size = 0 chooses it (kgmt_batch_cluster_size). */
int  kgmt_plan_batch(kgmt_ctx* ctx, const float* h_inits7, const float* h_goals7, const uint32_t* h_seeds, int Q,
                     int cluster_size, kgmt_result* out, float* h_paths7, int max_path, int* h_path_len, float* device_ms);
/* the cluster size kgmt_plan_batch uses for Q queries when asked with 0: the largest of 8, 4, 2 whose Q clusters are all
 * resident at once (few queries: more CTAs per query shorten every query), else 2 (many queries: the measured optimum,
 * workspaces recycled by ticket). */
int  kgmt_batch_cluster_size(kgmt_ctx* ctx, int Q);

/* ---- sharded expansion (BASELINE config 5; SURVEY.md §8e "sharded expansion") --------------------------------------
 * ONE iteration's candidates are split over `world` ranks; tree and maps are replicated on every GPU and stay
 * bit-identical to a single-GPU run (the reference has no multi-GPU mode; this replaces one while-loop body,
 * KGMT.cu:118-259, per round of the three calls below).  The collectives in between are the caller's (NCCL):
 *
 *   kgmt_shard_expand   stages 2-5a on this rank's contiguous candidate range; region-counter increments go to the
 *                       zero-initialised DEVICE slab d_delta (kgmt_shard_delta_ints ints: R1,R1Valid,R1Invalid,scratch
 *                       [N*N] | R2,R2Valid,R2Invalid,first-reached flags [N*N*n*n]); reports the rows it accepted
 *        -> all-gather(accepted_local) ; cap_rows = max over ranks rounded up to a multiple of 4
 *   kgmt_shard_pack     this rank's accepted rows, candidate order, into DEVICE d_send:
 *                       float4 state[cap_rows] | float4 (a, steering, duration, cost)[cap_rows] | int32 slot[cap_rows]
 *        -> all-gather(d_send, 36*cap_rows bytes per rank) into d_recv ; all-reduce SUM(d_delta)
 *   kgmt_shard_commit   inserts every rank's rows in rank order (= global candidate order), adds the reduced deltas to
 *                       the maps (the slab is zeroed again), advances the planner, scores the next iteration
 * Mixing these with kgmt_expand_iteration(s) on the same context is allowed between complete rounds. */
typedef struct kgmt_shard_info {
    int iteration, candidates, children, frontier;   /* the pending iteration (identical on every rank) */
    int chunk_lo, chunk_hi;                          /* this rank's 32-candidate chunks [lo, hi) */
    int accepted_local;                              /* rows this rank contributes */
    int stop;                                        /* kgmt_stop before this iteration */
} kgmt_shard_info;
size_t kgmt_shard_delta_ints(const kgmt_ctx* ctx);
int  kgmt_shard_expand(kgmt_ctx* ctx, int rank, int world, int* d_delta, kgmt_shard_info* out);
int  kgmt_shard_pack(kgmt_ctx* ctx, void* d_send, int cap_rows);
int  kgmt_shard_commit(kgmt_ctx* ctx, const void* d_recv, int cap_rows, const int* h_counts, int world, int* d_delta,
                       kgmt_iter_stats* out);
/* ---- the same sharded expansion with the exchange done by the kernels themselves over PEER MEMORY (NVLink /
 * NVSwitch), no NCCL on the data path: accepted rows are written straight into every rank's tree, the counter deltas
 * are all-reduced as reduce-scatter + all-gather over peer loads/stores, counts and the goal candidate travel through
 * device mailboxes with system-scope release/acquire.  Setup: every rank exports kgmt_peer_handle_bytes() bytes of
 * cudaIpc handles (kgmt_peer_export), the host program gathers them (any transport) and every rank attaches
 * (kgmt_peer_attach).  Then one kgmt_peer_expand_begin + kgmt_peer_expand_end per iteration on every rank; results are
 * bit-identical to kgmt_expand_iteration on one GPU.  A peer that does not arrive within 5 s gives KGMT_ERR_COMM
 * instead of a hung GPU; the replicas of an aborted exchange are undefined — restart the plan (kgmt_begin) on every
 * rank.  kgmt_peer_attach_local wires contexts of one process (tests). */
size_t kgmt_peer_handle_bytes(void);
int  kgmt_peer_export(kgmt_ctx* ctx, void* out_handles, size_t bytes);
int  kgmt_peer_attach(kgmt_ctx* ctx, int rank, int world, const void* all_handles);
int  kgmt_peer_attach_local(kgmt_ctx* ctx, int rank, int world, kgmt_ctx* const* peers);
int  kgmt_peer_expand_begin(kgmt_ctx* ctx);
int  kgmt_peer_expand_end(kgmt_ctx* ctx, kgmt_iter_stats* out);
int  kgmt_peer_detach(kgmt_ctx* ctx);
/* The same exchange FUSED with the expansion into one persistent cooperative kernel per rank (compute + collective in one
 * launch, the grid barrier as the only local synchronisation, no host round trip between iterations): up to `count`
 * sharded iterations, or a whole plan (root insertion on every rank + every iteration until the planner stops — KGMT::plan,
 * KGMT.cu:80-317, with each iteration's candidates split over the ranks).  Every attached rank makes the same call;
 * every rank ends with the same scalars and a tree bit-identical to the single-GPU one. */
int  kgmt_peer_expand_iterations(kgmt_ctx* ctx, int count, kgmt_iter_stats* out);
int  kgmt_peer_plan(kgmt_ctx* ctx, const float* initial7, const float* goal7, kgmt_result* out);
/* portfolio race between the attached ranks: same query, own seed per rank (kgmt_set_seed), ONE launch per rank; the
 * first rank to reach the goal stops the others through a word in their memory; they return KGMT_PEER_SOLVED.
 * race_id > 0, growing from race to race (KGMT::plan has no counterpart: KGMT.cu:118-259 is a single-GPU loop). */
int  kgmt_peer_race(kgmt_ctx* ctx, const float* initial7, const float* goal7, int race_id, kgmt_result* out);
/* ---- multi-GPU communicator: one process per GPU on one NVSwitch box, host side in C/C++ (no Python needed) ---------
 * The reference is single-GPU (KGMT::plan, KGMT.cu:80-317, called by demos/main.cu:62); these calls are what its main()
 * would use with one process per GPU.  NCCL (loaded with dlopen on first use) carries the setup and the result
 * collectives; the data path of the sharded expansion runs over peer memory inside the library's own kernels.
 *   kgmt_comm_unique_id   rank 0: a 128-byte NCCL unique id; the host program hands it to every rank (pipe, file, MPI ...)
 *   kgmt_comm_init        NCCL communicator on the context's device; all-gather of the peer-memory handles; attach;
 *                         barrier.  After it every kgmt_peer_* call and the calls below are available.
 *   kgmt_plan_batch_sharded  BASELINE config 4: Q independent queries, contiguous shards over the ranks, every rank plans
 *                         its shard with kgmt_plan_batch; out_all[Q] (identical on every rank) by one NCCL all-gather
 *   kgmt_plan_portfolio   the SAME query with seed base_seed + rank on every rank, ONE launch per rank; the first rank
 *                         to reach the goal stops the others through peer memory (kgmt_peer_race); the lowest-cost
 *                         solution's result block and path are broadcast to every rank (NCCL) — first-solution broadcast
 *   kgmt_expand_sharded   BASELINE config 5: ONE iteration's candidates split over the ranks.  exchange selects how the
 *                         accepted rows / counters travel; ms3 (optional) = {compute ms, exchange ms, exchange bytes}
 *                         (the fused kernel cannot separate them: {total ms, 0, 0})
 *   kgmt_plan_sharded     whole plans with sharded iterations in one persistent kernel per rank (= kgmt_peer_plan) */
typedef enum kgmt_exchange {
    KGMT_EXCHANGE_FUSED = 0,          /* compute + exchange in one persistent kernel per rank, peer memory (NVLink) */
    KGMT_EXCHANGE_PEER_LAUNCHES = 1,  /* the same exchange as a sequence of kernels (kgmt_peer_expand_begin / _end) */
    KGMT_EXCHANGE_NCCL = 2            /* all-gather(counts), all-gather(rows), all-reduce(counter deltas) by NCCL */
} kgmt_exchange;
#define KGMT_COMM_ID_BYTES 128
int  kgmt_comm_unique_id(void* out_id128);
int  kgmt_comm_init(kgmt_ctx* ctx, int rank, int world, const void* nccl_unique_id128);
int  kgmt_comm_destroy(kgmt_ctx* ctx);
int  kgmt_comm_barrier(kgmt_ctx* ctx);
int  kgmt_comm_rank(const kgmt_ctx* ctx);
int  kgmt_comm_world(const kgmt_ctx* ctx);
int  kgmt_plan_batch_sharded(kgmt_ctx* ctx, const float* h_inits7, const float* h_goals7, const uint32_t* h_seeds, int Q,
                             int cluster_size, kgmt_result* out_all, float* device_ms_max);
int  kgmt_plan_portfolio(kgmt_ctx* ctx, const float* initial7, const float* goal7, uint32_t base_seed, int race_id,
                         kgmt_result* out_winner, int* winner_rank, float* h_path7, int max_rows, int* path_len);
int  kgmt_expand_sharded(kgmt_ctx* ctx, int exchange, kgmt_iter_stats* out, float* ms3);
int  kgmt_plan_sharded(kgmt_ctx* ctx, const float* initial7, const float* goal7, kgmt_result* out);
/* launch on the caller's CUDA stream (cudaStream_t) instead of the context's own; NULL restores it.  Lets the calls
 * above order with NCCL collectives enqueued on the same stream without extra synchronisation. */
int  kgmt_set_stream(kgmt_ctx* ctx, void* cuda_stream);

/* ---- stage-level entry points (parity tests, throughput sweeps) ----------------------------- */
/* Stage 1: R1 scores from the current maps (updateR1, KGMT.cu:487-538). */
int  kgmt_stage_scores(kgmt_ctx* ctx);
/* Stages 2-4 only: candidate s (0 <= s < P*children) expands h_parents7[s / children] with stream
 * (key0, slot0 + s); fills the candidate arrays (UNEXPLORED, U_VALID, U_R1, U_R2, U_U3); maps and tree untouched.
 * (propagateAndCheck, statePropagator.cu:5-76 + getR1/getR2, KGMT.cu:602-629) */
int  kgmt_stage_propagate(kgmt_ctx* ctx, const float* h_parents7, int P, int children,
                          uint32_t key0, uint32_t slot0, float* device_ms);
/* Stage 5a alone on CALLER-SUPPLIED candidates (tail of propagateG / propagateGV2, KGMT.cu:390-411 / :458-480): the M
 * candidates (HOST: rows of 7 floats as the reference's unexploredSamples, valid flags 0/1, accept uniforms, parent tree
 * indices) replace the pending iteration's candidates: region indices, R1/R2 counters, accept test against the CURRENT
 * maps and scores (set them with kgmt_import) on the iteration-start snapshot, ballots and staging for the insertion.
 * The per-candidate results are exported as KGMT_ARR_U_R1 / U_R2 / U_ACCEPT; needs record_candidates = 1 and a begun
 * context (kgmt_begin / kgmt_seed_frontier); M <= max_candidates. */
int  kgmt_stage_update_maps(kgmt_ctx* ctx, const float* h_cand7, const unsigned char* h_valid, const float* h_u3,
                            const int* h_parent, int M);
/* Stage 5b alone (scan(GNew) + findInd + updateG, KGMT.cu:222-245, :540-593, then :249-259): ordered insertion of the
 * candidates accepted by the preceding kgmt_stage_update_maps, parent links from its h_parent, cost = cost[parent] +
 * duration, goal test; advances the planner scalars and scores the next iteration. */
int  kgmt_stage_insert(kgmt_ctx* ctx, kgmt_iter_stats* out);
/* Load `count` nodes (HOST rows of 7 floats) as tree[0,count), all of them frontier, costs 0, root cells
 * marked like KGMT.cu:88-97 for every node; then kgmt_expand_iteration steps from there. */
int  kgmt_seed_frontier(kgmt_ctx* ctx, const float* h_nodes7, int count, const float* goal7);
/* > 0: that many children per frontier node in every later iteration (throughput sweeps, BASELINE config 5);
 * 0: the reference policy (32, or floor(remaining/active) when the tree is nearly full, KGMT.cu:151-158) */
int  kgmt_set_children(kgmt_ctx* ctx, int children);
/* device-side checkpoint / restore of maps + scalars (tree rows above the checkpointed size are dead) */
int  kgmt_checkpoint(kgmt_ctx* ctx);
int  kgmt_restore(kgmt_ctx* ctx);

/* ---- data exchange -------------------------------------------------------------------------- */
int  kgmt_export(kgmt_ctx* ctx, int array_id, void* h_dst, size_t bytes);        /* copyAndWriteVectorToCSV's D2H half, helper.cuh:74-79 */
int  kgmt_import(kgmt_ctx* ctx, int array_id, const void* h_src, size_t bytes);  /* maps only (ids 5..13) */
size_t kgmt_array_bytes(const kgmt_ctx* ctx, int array_id);
/* the reference's 13 CSV files (KGMT.cu:299-311, "%.10f" fixed, helper.cuh:53-72) into `dir` */
int  kgmt_dump_csv(kgmt_ctx* ctx, const char* dir);

/* ---- introspection -------------------------------------------------------------------------- */
int  kgmt_tree_size(const kgmt_ctx* ctx);
float kgmt_cost_to_goal(const kgmt_ctx* ctx);
float kgmt_r1_size(const kgmt_ctx* ctx);                         /* KGMT::R1Size_, KGMT.cu:13 */
float kgmt_r2_size(const kgmt_ctx* ctx);                         /* KGMT::R2Size_, KGMT.cu:14 */
void* kgmt_stream(const kgmt_ctx* ctx);                          /* cudaStream_t the context launches on */
long long kgmt_launch_count(const kgmt_ctx* ctx);                /* kernels of this library launched so far */
/* work done by the recording kernels (record_candidates = 1) since the last kgmt_begin / kgmt_plan / kgmt_seed_frontier:
 * out4 = {Euler steps executed (statePropagator.cu:31 loop trips), (step bbox, obstacle AABB) overlap tests executed
 * (collisionCheck.cu:6-14 calls; with the culled back end: entries read from the cell lists, padding included),
 * candidate edges, 0}.  The non-recording kernels do not count (the counters cost instructions). */
int  kgmt_work_counters(kgmt_ctx* ctx, unsigned long long* out4);
/* the bounds-checked build of the library (libkgmt_b200_check.so, -DKGMT_BOUNDS_CHECK): out4 = {id of the first index
 * check that failed on the device (0 = none), number of failures, offending value, its limit} since the last call.
 * The product build returns KGMT_ERR_STATE. */
int  kgmt_debug_checks(kgmt_ctx* ctx, int* out4);
/* diagnostics: enable != 0 turns on per-iteration device timestamps; out8 rows (8 x u64) = {globaltimer ns when the
 * iteration was finalized, candidates << 32 | accepted, then CTA 0's globaltimer at: iteration start, phase A done,
 * first grid barrier passed, phase B done, last grid barrier passed, 0} of the last plan; returns rows written */
int  kgmt_iteration_log(kgmt_ctx* ctx, int enable, unsigned long long* out8, int max_rows);
/* out8 = {collision back end (0 grid/smem, 1 grid/L1, 2 exhaustive/smem, 3 exhaustive/L1, 4 exhaustive/TMA-streamed tiles), cull cells per side,
 *         cull grid items, dynamic shared memory bytes, persistent grid size, SM count, R1 smem histograms, K} */
int  kgmt_get_config(const kgmt_ctx* ctx, int* out8);

#ifdef __cplusplus
}
#endif
#endif /* KGMT_C_H */
