/* kgmt_kernels.cuh — the sm_100a kernels of the KGMT tree-expansion step.
 *
 * One persistent cooperative kernel replaces the reference's per-iteration
 *   scan(R1Avail)+updateR1, scan(G)+findInd, propagateG|propagateGV2,
 *   scan(GNew)+findInd, updateG      (src/planners/KGMT.cu:118-259)
 * and the whole host loop of KGMT::plan around them.
 *
 *   stage 1  frontier  = the contiguous node range appended by the previous
 *            iteration (no scan: KGMT.cu:378,451,568,582 make G exactly that);
 *            R1 scores are produced by the CTA that finishes an iteration last.
 *   phase A  (warps fully independent, no CTA barrier) per 32-candidate chunk:
 *     stage 2  stateless Philox4x32-10 per candidate slot (kgmt_device.cuh)
 *     stage 3  Euler car dynamics on SoA float4 node storage
 *     stage 4  step-bbox vs obstacle AABBs staged in shared memory by bulk async
 *              copies (cp.async.bulk + mbarrier): exhaustive, or culled through a
 *              uniform grid (identical flags)
 *     stage 5a region counters (R1 family in per-CTA shared-memory histograms, R2
 *              family with global reductions), accept test on the iteration-start
 *              snapshot; accepted candidates are ballot-compacted inside the chunk
 *              into a staging row, the ballot mask is the chunk's scan input
 *   ONE grid barrier per iteration
 *   phase B  stage 5b ordered insertion: block sums + a 256-wide scan of the chunk
 *            popcounts give every accepted candidate its tree slot in candidate
 *            order; rows move staging -> tree with float4 loads/stores.
 *            Every CTA advances the planner scalars itself (same inputs, same result),
 *            so there is no second barrier: the next phase A starts at once and a
 *            chunk waits (acquire) only for the 256-row tree segment holding its
 *            parent and for the score buffer of its iteration.
 *
 * Also in this file: batch_kernel (one thread-block cluster per query, config 4), shard_* (one iteration split over ranks, NCCL
 * exchange by the caller) and peer_* (the same with the exchange done over peer memory), propagate_only_kernel
 * (stages 2-4 alone), setup / export kernels.
 *
 * Canonical semantics where the reference races: SURVEY.md Appendix B.
 */
#pragma once
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kgmt_device.cuh"

namespace kgmt {
namespace cg = cooperative_groups;

constexpr int TILE = 256;                 /* threads per CTA */
constexpr int WARPS = TILE / 32;
constexpr int CHUNK = 32;                 /* candidates per chunk == one warp */
constexpr int BLK_CHUNKS = 256;           /* chunks per scan block (one per thread of a CTA in phase B) */

enum { STOP_RUNNING = 0, STOP_SOLVED = 1, STOP_TREE_FULL = 2, STOP_ITER_LIMIT = 3, STOP_FRONTIER_EMPTY = 4,
       STOP_PEER_SOLVED = 5 /* portfolio race: another GPU reached the goal first */ };
enum { COL_GRID_SMEM = 0, COL_GRID_GLOBAL = 1, COL_BRUTE_SMEM = 2, COL_BRUTE_GLOBAL = 3, COL_BRUTE_STREAM = 4 };
constexpr int STREAM_TILE = 1024;         /* obstacles per streamed shared-memory tile (16 KB), double buffered */
enum { FLAG_VALID = 1, FLAG_ACCEPT = 2 };

/* planner scalars.  The first COPIED_WORDS ints are advanced identically by every CTA in shared memory and written
 * back by CTA 0; the live fields after them are only ever touched in global memory (atomics / flags). */
struct DevState {
    int treeSize, frontierStart, frontierCount, itr;           /* itr = iteration about to run (1-based) */
    int stop, forceChildren; float costToGoal; float R1Threshold;
    int mode, children, M, numChunks;                           /* shape of the iteration about to run */
    int lastMode, lastChildren, lastFrontier, lastM;
    int lastAccepted, lastItr, iterationsDone, goalSlot;        /* goalSlot: candidate slot of the goal node in its iteration */
    int blocksTotal, pad0;                                      /* scan blocks of all finished iterations (see insertDone) */
    long long expansions;
    /* ---- live */
    /* goalBest / peerSolved are indexed by ITERATION PARITY: after the grid barrier of iteration i every CTA reads slot
     * i & 1 while CTAs that are already in phase A of iteration i+1 write slot (i+1) & 1 — one shared word would let a
     * lagging CTA see an i+1 goal, stop alone and hang the rest at the next barrier.  A slot written in iteration i is
     * next written in i+2, which no CTA enters before all have passed barrier i+1, i.e. after all have read slot i & 1.
     * A goal stops the plan, so a slot never needs clearing inside a plan. */
    /* each group of live words has its own 128-byte line: thousands of warps poll insertDone / scoreReady while CTA 0
     * rewrites the copied words and thread 0 of every CTA reads goalBest right after the barrier */
    alignas(128) unsigned long long goalBest[2];                /* (cost bits << 32) | candidate slot, ~0 = none (atomicMin) */
    int goalIdx;                                                /* tree index of the goal node (written by its inserter) */
    int peerSolved[2];                                          /* portfolio race: a peer's win as seen before this iteration's barrier */
    int pad2;
    /* work counters (diagnostics, kgmt_work_counters): Euler steps executed, (step, obstacle) AABB tests executed */
    unsigned long long stepsDone, pairsTested;
    alignas(128) int scoreReady;                                /* scores of this iteration are complete (release/acquire) */
    alignas(128) int insertDone;                                /* scan blocks inserted so far, whole plan (release counter) */
};
constexpr int COPIED_WORDS = 24;
constexpr int THRESHOLD_WORD = 7;         /* R1Threshold: written by scores_block straight to global memory, never copied back */
static_assert(offsetof(DevState, goalBest) == 128 && COPIED_WORDS * 4 <= 128, "DevState layout");
static_assert(offsetof(DevState, R1Threshold) == THRESHOLD_WORD * 4, "DevState layout");

struct KArgs {
    /* tree, SoA */
    float4* treeState;            /* (x, y, theta, v) */
    float4* treeCtrl;             /* (a, steering, duration, cost) */
    int*    treeParent;
    /* occupancy maps */
    int *R1, *R1Valid, *R1Invalid, *R1Avail, *R1Cov; float* R1Score[2];   /* scores: buffer itr&1 */
    int *R2, *R2Valid, *R2Invalid; unsigned* R2Stamp;
    int* R2StampDelta;            /* sharded expansion only: cells this rank reached for the first time (0/1) */
    /* per-candidate records (null unless recording) */
    float4* candState; float4* candCtrl; int* candParent; int* candR1; int* candR2; unsigned char* candFlags;
    /* ordered insertion: everything indexed by iteration parity / iteration mod 3 */
    unsigned* chunkMask;          /* [2][chunksCap] accept ballot of each 32-candidate chunk */
    int* blockSum;                /* [3][blocksCap] accepted candidates per scan block */
    unsigned* ticket;             /* [3] next chunk to hand out */
    float4* stageState; float4* stageCtrl;   /* [2][maxCand] accepted rows, compacted inside their chunk */
    int chunksCap, blocksCap, maxCand, totalWarps;
    /* portfolio race over peer memory (kgmt_peer_race): raceId > 0 switches it on */
    int raceId, raceWorld, raceRank;
    int* const* raceFlags;        /* device table [raceWorld]: every rank's 'a plan of race N was solved' word, mapped here */
    DevState* st;
    /* collision */
    const float4* obstacles; int K;
    const int* cellStart; const float4* cellItems; int cullC; float cullInvX, cullInvY; int cellStartInts; int numItems;
    int obsTile;                  /* obstacles per shared-memory tile (stream mode) */
    unsigned long long* iterLog;  /* [256][8]: per iteration {advance ns, M<<32|accepted, CTA0: start, phase A done,
                                     barrier passed, phase B done, 0, 0}; null = off */
    /* parameters */
    float W, H, L, R1Size, R2Size, goalX, goalY, goalR;
    int N, n, c1, numDisc, maxTree, numIterations, useHist;
    uint32_t seed;
    CarRanges car;                /* control ranges (statePropagator.cu:17-19 literals by default) */
};

/* ------------------------------------------------------------------ small PTX helpers -- */
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
/* 1-D bulk async copy global -> shared (TMA engine; SASS: UBLKCP), completion on an mbarrier */
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
/* issue a copy of any 16-byte-multiple size in <= 32 KB pieces; caller has already posted expect_tx */
__device__ __forceinline__ void bulk_g2s_chunked(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    for (uint32_t off = 0; off < bytes; off += 32768u) {
        const uint32_t nb = min(32768u, bytes - off);
        bulk_g2s((char*)dst + off, (const char*)src + off, nb, bar);
    }
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

/* Pin a warp-uniform value in a register.  Values the compiler can re-derive from the kernel parameters (constant bank)
 * or special registers are re-derived inside the hot loops whenever that saves a register — a few extra issue slots per
 * Euler step and per cull-grid row, eleven per row for the shared-memory base of the CSR.  A value that comes back from
 * a shuffle (addresses built from special registers) or from a volatile shared-memory load (kernel parameters: the
 * assembler folds a shuffle of a constant) is opaque to that and has to be kept. */
__device__ __forceinline__ uint32_t pin_u32(uint32_t v) {
    uint32_t r;
    asm volatile("shfl.sync.idx.b32 %0, %1, 0, 0x1f, 0xffffffff;" : "=r"(r) : "r"(v));
    return r;
}
__device__ __forceinline__ uint32_t lds_pinned(const uint32_t* p) {
    uint32_t r;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(r) : "r"((uint32_t)__cvta_generic_to_shared(p)));
    return r;
}

/* ------------------------------------------------------------------------- stage 1 ----
 * R1 scores, updateR1 (KGMT.cu:487-538) for any N, by one CTA of TILE threads.
 * covR uses the running count of available R2 cells (R1Cov) instead of re-summing
 * n*n flags (:510-514) — the same integer.  The sum order is fixed (the reference's
 * cub::BlockReduce order is unspecified): p[t] = sum_k score[t+1024k], then a
 * stride-halving tree (DESIGN.md "scores"; the CPU checker restates the same order). */
__device__ void scores_block(const KArgs& A, float* p /* smem [1024] */, float* out /* [c1] */) {
    const int tid = threadIdx.x, c1 = A.c1;
    const float nn = (float)(A.n * A.n);
    int availLocal = 0;
    for (int t = tid; t < 1024; t += TILE) {
        float acc = 0.0f;
        for (int c = t; c < c1; c += 1024) {
            float score = 0.0f;
            if (__ldcg(&A.R1Avail[c]) != 0) {
                ++availLocal;
                const float covR = __fdiv_rn((float)__ldcg(&A.R1Cov[c]), nn);
                const float nV = (float)__ldcg(&A.R1Valid[c]), nI = (float)__ldcg(&A.R1Invalid[c]);
                const float num = __fadd_rn(0.01f, nV);
                const float freeVol = __fdiv_rn(num, __fadd_rn(num, nI));
                const double f2 = __dmul_rn((double)freeVol, (double)freeVol);
                const double f4 = __dmul_rn(f2, f2);
                const double r = (double)__ldcg(&A.R1[c]);
                const double den = __dmul_rn((double)__fadd_rn(1.0f, covR), __dadd_rn(1.0, __dmul_rn(r, r)));
                score = __double2float_rn(__ddiv_rn(f4, den));
            }
            out[c] = score;                             /* raw; normalised below */
            acc = __fadd_rn(acc, score);
        }
        p[t] = acc;
    }
    __shared__ int sAvail;
    if (tid == 0) sAvail = 0;
    __syncthreads();
    if (availLocal) atomicAdd(&sAvail, availLocal);
    for (int stride = 512; stride >= 1; stride >>= 1) {
        __syncthreads();
        for (int t = tid; t < stride; t += TILE) p[t] = __fadd_rn(p[t], p[t + stride]);
    }
    __syncthreads();
    const float total = p[0];
    if (tid == 0) A.st->R1Threshold = sAvail ? __fdiv_rn(total, (float)sAvail) : 0.0f;
    for (int c = tid; c < c1; c += TILE)
        out[c] = (__ldcg(&A.R1Avail[c]) == 0) ? 1.0f : __fdiv_rn(out[c], total);
    __syncthreads();
}

/* expansion policy, KGMT.cu:151-158 (canonical prefix mode: SURVEY.md App. B #7).  An iteration never produces more
 * candidates than fit the remaining tree rows AND the candidate staging (maxCand): forced children (mode 4) that do
 * not fit fall back to the reference policy, and the reference policy itself is evaluated on min(remaining, maxCand)
 * — identical to the reference whenever maxCand >= remaining, which is the default (max_candidates = 0). */
__host__ __device__ __forceinline__ void expansion_shape(int active, int treeSize, int maxTree, int maxCand, int forceChildren,
                                                         int& mode, int& children, int& M) {
    const long long remaining = (long long)maxTree - treeSize;
    const long long cap = remaining < (long long)maxCand ? remaining : (long long)maxCand;
    if (forceChildren > 0 && (long long)active * forceChildren <= cap) {
        mode = 4; children = forceChildren; M = active * forceChildren; return;
    }
    if (32LL * active > cap) {
        const int it = (cap == remaining) ? (int)((float)remaining / (float)active)      /* KGMT.cu:153, float division */
                                          : (int)(cap / active);
        if (it >= 1) { mode = 2; children = it; M = active * it; }
        else         { mode = 3; children = 1;  M = (int)cap; }
    } else { mode = 1; children = 32; M = 32 * active; }
}

__device__ __forceinline__ int ld_acquire_s32(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_s32(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_relaxed_s32(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_s32(int* p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys_s32(const int* p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_acq_rel() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
/* wait until *p >= want: relaxed polling with back-off, one acquire fence at the end */
__device__ __forceinline__ void wait_ge(const int* p, int want) {
    unsigned ns = 64;
    while (ld_relaxed_s32(p) < want) { __nanosleep(ns); if (ns < 1024) ns <<= 1; }
    fence_acq_rel();
}

/* sum of v over the CTA (all TILE threads call; result in every thread) */
__device__ __forceinline__ int block_sum(int v, int* sRed /* [WARPS] */) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sRed[threadIdx.x >> 5] = v;
    __syncthreads();
    int t = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) t += sRed[w];
    return t;
}

/* two sums at once (same barriers as one) */
__device__ __forceinline__ int2 block_sum2(int a, int b, int* sRed /* [2 * WARPS] */) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { sRed[threadIdx.x >> 5] = a; sRed[WARPS + (threadIdx.x >> 5)] = b; }
    __syncthreads();
    int ta = 0, tb = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) { ta += sRed[w]; tb += sRed[WARPS + w]; }
    return make_int2(ta, tb);
}

/* Ordered insertion of one scan block can be SPLIT over several CTAs (each takes every n-th group of 32 rows): small
 * iterations have one or two blocks with thousands of accepted rows, and one CTA moving them is a chain of dependent L2
 * round trips (4-10 us of a 20 us iteration on config 1).  Pure function of the iteration shape, so every CTA agrees. */
__host__ __device__ __forceinline__ int insert_split(int numBlocks, int groupCtas) {
    const int s = groupCtas / (numBlocks > 0 ? numBlocks : 1);
    return s < 1 ? 1 : (s > 8 ? 8 : s);
}

/* end of an iteration, KGMT.cu:249-259 + the next iteration's :119: pure function of (S, accepted, goalBest),
 * evaluated by thread 0 of EVERY CTA on its own copy. */
__device__ __forceinline__ void advance_state(const KArgs& A, DevState& S, int accepted, unsigned long long gb, int peerSolved = 0) {
    S.lastMode = S.mode; S.lastChildren = S.children; S.lastFrontier = S.frontierCount;
    S.lastM = S.M; S.lastAccepted = accepted; S.lastItr = S.itr;
    S.iterationsDone += 1;
    {
        const int nb = (S.numChunks + BLK_CHUNKS - 1) / BLK_CHUNKS;
        S.blocksTotal += nb * insert_split(nb, A.totalWarps / WARPS);       /* insertion work items of the iteration */
    }
    S.expansions += S.M;
    S.frontierStart = S.treeSize;
    S.frontierCount = accepted;
    S.treeSize += accepted;                                                         /* :249 */
    if (gb != ~0ull && S.costToGoal == 0.0f) {                                      /* canonical min-cost goal (App. B #5) */
        S.costToGoal = __uint_as_float((unsigned)(gb >> 32));
        S.goalSlot = (int)(unsigned)gb;
    }
    int stop = STOP_RUNNING;
    if (S.costToGoal != 0.0f)              stop = STOP_SOLVED;                      /* :252 */
    else if (S.treeSize >= A.maxTree)      stop = STOP_TREE_FULL;                   /* :255 */
    else if (accepted == 0)                stop = STOP_FRONTIER_EMPTY;
    else if (S.itr >= A.numIterations)     stop = STOP_ITER_LIMIT;                  /* :118 */
    else if (peerSolved)                   stop = STOP_PEER_SOLVED;
    S.stop = stop;
    if (stop == STOP_RUNNING) {
        S.itr += 1;                                                                 /* :119 */
        int mode, children, Mn;
        expansion_shape(accepted, S.treeSize, A.maxTree, A.maxCand, S.forceChildren, mode, children, Mn);
        S.mode = mode; S.children = children; S.M = Mn; S.numChunks = (Mn + CHUNK - 1) / CHUNK;
    }
}

/* per-iteration view, identical on every thread */
struct IterView {
    int itr, treeSize, frontierStart, children, M, numChunks; uint32_t key0;
    const float* score;                      /* R1 scores of this iteration */
    unsigned* chunkMask; int* blockSum; float4* stageState; float4* stageCtrl;   /* this iteration's buffers */
    int goalSlot;                            /* >= 0: this iteration produced the goal node at that candidate slot */
    const int* parentOf;                     /* null: parent = frontierStart + slot / children; else explicit (kgmt_stage_*) */
};

__device__ __forceinline__ IterView make_view(const KArgs& A, const DevState& S) {
    IterView it;
    it.itr = S.itr; it.treeSize = S.treeSize; it.frontierStart = S.frontierStart; it.children = S.children;
    it.M = S.M; it.numChunks = S.numChunks; it.key0 = A.seed + (uint32_t)S.itr;
    it.score = A.R1Score[S.itr & 1];
    it.chunkMask = A.chunkMask + (size_t)(S.itr & 1) * A.chunksCap;
    it.blockSum = A.blockSum + (size_t)(S.itr % 3) * A.blocksCap;
    it.stageState = A.stageState + (size_t)(S.itr & 1) * A.maxCand;
    it.stageCtrl = A.stageCtrl + (size_t)(S.itr & 1) * A.maxCand;
    it.goalSlot = -1;
    it.parentOf = nullptr;
    return it;
}

/* ---------------------------------------------------------------- phase A: one chunk ---
 * 32 candidates, one per lane: stages 2-5a.  No communication outside the warp.
 * scoresOk (warp-uniform) remembers that this iteration's score buffer has been seen complete. */
/* one candidate of a chunk between its stages */
struct ChunkCand {
    float4 x; Controls u; int parent; float parentCost; bool live, valid;
};

/* ------------------------------------------------ parents straight from the staging rows ---
 * Phase B moves the accepted rows of iteration i from the staging buffers into the tree, and phase A of iteration i+1
 * reads its parents from the tree: a chunk that starts right after the barrier would wait for the whole of phase B
 * (block sums -> ballots -> rows -> fence -> flag -> poll: 4-5 us of dependent L2 round trips, every iteration).
 * Until the tree rows are announced a chunk whose 32 candidates share ONE parent (32 children per node, the reference
 * policy while the tree has room) fetches that parent from where phase A left it: the q-th new node is the r-th
 * accepted candidate of the scan block whose prefix range holds q; the block comes from the per-CTA prefix of the
 * block sums (already loaded for advance_state), the chunk and the rank inside it from the block's 256 ballots
 * (eight per lane, one warp scan).  Same row as the tree will hold, so the same bits. */
constexpr int BLK_PRE_CAP = TILE;         /* scan blocks whose prefix a CTA keeps (one per thread) */
struct StagedParents {
    const int* blkPre;                    /* shared memory: exclusive prefix of the previous iteration's block sums */
    int numBlocks, numChunks;             /* of the previous iteration */
    const unsigned* mask; const float4* stageState; const float4* stageCtrl;
};
__device__ __forceinline__ void staged_parent(const StagedParents& sp, int q, int lane, float4& x, float& cost) {
    int lo = 0, hi = sp.numBlocks - 1;
    while (lo < hi) {                     /* last block whose prefix is <= q (empty blocks share a prefix with their successor) */
        const int mid = (lo + hi + 1) >> 1;
        if (sp.blkPre[mid] <= q) lo = mid; else hi = mid - 1;
    }
    const int r = q - sp.blkPre[lo];      /* rank of the row inside the block */
    const int c0 = lo * BLK_CHUNKS + lane * 8;
    const uint4* mp = reinterpret_cast<const uint4*>(sp.mask + c0);
    const uint4 ma = __ldcg(mp), mb = __ldcg(mp + 1);
    unsigned m[8] = {ma.x, ma.y, ma.z, ma.w, mb.x, mb.y, mb.z, mb.w};      /* fully unrolled below: stays in registers */
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { if (c0 + k >= sp.numChunks) m[k] = 0u; cnt += __popc(m[k]); }   /* ballots past the end are stale */
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    const unsigned owners = __ballot_sync(0xffffffffu, incl > r);
    const int owner = owners ? (__ffs(owners) - 1) : 31;
    /* every lane resolves (chunk, rank) as if it were the owner; the owner's answer is broadcast */
    int rr = r - (incl - cnt), ci = c0;
#pragma unroll
    for (int k = 0; k < 7; ++k) { const int pc = __popc(m[k]); if (ci == c0 + k && rr >= pc) { rr -= pc; ci = c0 + k + 1; } }
    ci = __shfl_sync(0xffffffffu, ci, owner);
    rr = __shfl_sync(0xffffffffu, rr, owner);
    KGMT_CHECK_RANGE(210, ci, sp.numChunks); KGMT_CHECK_RANGE(211, rr, CHUNK);
    x = __ldcg(&sp.stageState[ci * CHUNK + rr]);
    cost = __ldcg(&sp.stageCtrl[ci * CHUNK + rr]).w;
}

/* stage 2 + the parent read of stage 3 for lane `lane` of chunk c */
__device__ __forceinline__ ChunkCand chunk_setup(const KArgs& A, const IterView& it, int c, int lane,
                                                 bool staged = false, const StagedParents& sp = StagedParents{}) {
    float4 sx = make_float4(0.f, 0.f, 0.f, 0.f); float sc = 0.f;
    if (staged) staged_parent(sp, c, lane, sx, sc);               /* whole warp: one parent (children == 32) */
    ChunkCand cc;
    const int s = c * CHUNK + lane;
    cc.live = s < it.M;
    cc.x = make_float4(0.f, 0.f, 0.f, 0.f);
    cc.u = Controls{0.f, 0.f, 0.f, 0.f};
    cc.parent = -1; cc.parentCost = 0.f; cc.valid = false;
    if (cc.live) {
        cc.parent = it.frontierStart + (it.children == 32 ? (s >> 5) : s / it.children);   /* KGMT.cu:374-376 / :454 */
        KGMT_CHECK_RANGE(201, cc.parent, it.treeSize);
        if (staged) { cc.x = sx; cc.parentCost = sc; }
        else {
            cc.x = __ldcg(&A.treeState[cc.parent]);                        /* L2-coherent: written by other SMs */
            cc.parentCost = __ldcg(&A.treeCtrl[cc.parent]).w;
        }
        cc.u = sample_controls((uint32_t)s, it.key0, A.car);
    }
    return cc;
}

/* stage 5a for the chunk: region indices, counters, accept on the iteration-start snapshot, ballot compaction */
template <bool RECORD, bool SHARD>
__device__ __forceinline__ void chunk_finish(const KArgs& A, const IterView& it, const ChunkCand& cc, int c, int lane,
                                             int* hV, int* hI, bool& scoresOk) {
    const int s = c * CHUNK + lane;
    const bool live = cc.live, valid = cc.valid;
    const float4 x = cc.x;
    const Controls u = cc.u;
    int r1 = -1, r2 = -1;
    bool accept = false;
    unsigned casOld = 1u; bool casDone = false;        /* first-reached R2 cell: the CAS round trip overlaps the rest of the stage */
    KGMT_CHECK_RANGE(202, c, A.chunksCap);
    if (live) KGMT_CHECK_RANGE(203, s, A.maxCand);
    if (live) {
        const RegionCell rc = region_r1_cell(x.x, x.y, A.R1Size, A.N);     /* KGMT.cu:390 */
        r1 = rc.r1;
        r2 = region_r2_cell(x.x, x.y, rc, A.R1Size, A.R2Size, A.n);        /* KGMT.cu:391 */
    }
    if (!scoresOk) {                /* scores of this iteration (and hence the maps of the previous one) are final */
        wait_ge(&A.st->scoreReady, it.itr);
        scoresOk = true;
    }
    if (live) { KGMT_CHECK_RANGE(204, r1 + 1, A.c1 + 1); KGMT_CHECK_RANGE(205, r2 + 1, (long long)A.c1 * A.n * A.n + 1); }
    if (live && r1 >= 0) {          /* maps + accept, KGMT.cu:392-411 on the iteration-start snapshot (App. B #1,#2) */
        if (valid) {
            accept = u.u3 <= __ldcg(&it.score[r1]);
            if (r2 >= 0) {
                const unsigned stamp = __ldcg(&A.R2Stamp[r2]);
                if (stamp == 0u || stamp > (unsigned)it.itr) accept = true;        /* unavailable at iteration start */
                if (stamp == 0u) {
                    if (SHARD) A.R2StampDelta[r2] = 1;      /* the stamp itself is set at commit, after the all-reduce */
                    else { casOld = atomicCAS(&A.R2Stamp[r2], 0u, (unsigned)it.itr + 1u); casDone = true; }   /* answer used last */
                }
                atomicAdd(&A.R2Valid[r2], 1);
            }
        } else if (r2 >= 0) {
            atomicAdd(&A.R2Invalid[r2], 1);
        }
        if (r2 >= 0) atomicAdd(&A.R2[r2], 1);
        if (A.useHist) {
            atomicAdd(valid ? &hV[r1] : &hI[r1], 1);
        } else {
            atomicAdd(&A.R1[r1], 1);
            if (valid) { atomicAdd(&A.R1Valid[r1], 1); A.R1Avail[r1] = 1; }
            else atomicAdd(&A.R1Invalid[r1], 1);
        }
    }
    /* ballot compaction inside the chunk; the mask is the input of the ordered scan (phase B) */
    const unsigned bal = __ballot_sync(0xffffffffu, accept);
    if (accept) {
        const int at = c * CHUNK + __popc(bal & ((1u << lane) - 1u));
        KGMT_CHECK_RANGE(206, at, A.maxCand);
        const float cost = __fadd_rn(cc.parentCost, u.duration);                   /* :585-586, :631-633 */
        __stcg(&it.stageState[at], x);
        __stcg(&it.stageCtrl[at], make_float4(u.a, u.steering, u.duration, cost));
        if (!SHARD && in_goal(x.x, x.y, A.goalX, A.goalY, A.goalR))               /* :589; min cost, then first in order */
            atomicMin(&A.st->goalBest[it.itr & 1], ((unsigned long long)__float_as_uint(cost) << 32) | (unsigned)s);
    }
    if (lane == 0) {
        __stcg(&it.chunkMask[c], bal);
        if (bal) atomicAdd(&it.blockSum[c / BLK_CHUNKS], __popc(bal));
    }
    if (casDone && casOld == 0u) atomicAdd(&A.R1Cov[r1], 1);     /* this candidate stamped the cell: covR of updateR1 (KGMT.cu:510-514) */
    if (RECORD && live) {
        A.candState[s] = x;
        A.candCtrl[s] = make_float4(u.a, u.steering, u.duration, u.u3);
        A.candParent[s] = cc.parent;
        A.candR1[s] = r1; A.candR2[s] = r2;
        A.candFlags[s] = (unsigned char)((valid ? FLAG_VALID : 0) | (accept ? FLAG_ACCEPT : 0));
    }
}

/* ---------------------------------------------------------------- phase A: one chunk ---
 * 32 candidates, one per lane: stages 2-5a.  No communication outside the warp.
 * scoresOk (warp-uniform) remembers that this iteration's score buffer has been seen complete. */
template <class Collide, bool RECORD, bool SHARD = false>
__device__ __forceinline__ void expand_chunk(const KArgs& A, const IterView& it, const DynParams& dyn,
                                             const Collide& col, int c, int lane, int* hV, int* hI, bool& scoresOk,
                                             bool staged = false, const StagedParents& sp = StagedParents{}) {
    ChunkCand cc = chunk_setup(A, it, c, lane, staged, sp);
    if (RECORD) {                   /* the recording kernels also count the work (kgmt_work_counters) */
        EdgeWork wk{0u, 0u};
        if (cc.live) cc.valid = propagate_edge(cc.x, cc.u, dyn, col, &wk);
        unsigned st = wk.steps, pr = wk.pairs;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { st += __shfl_xor_sync(0xffffffffu, st, o); pr += __shfl_xor_sync(0xffffffffu, pr, o); }
        if (lane == 0) { atomicAdd(&A.st->stepsDone, (unsigned long long)st); atomicAdd(&A.st->pairsTested, (unsigned long long)pr); }
    } else {
        if (cc.live) cc.valid = propagate_edge(cc.x, cc.u, dyn, col);
    }
    chunk_finish<RECORD, SHARD>(A, it, cc, c, lane, hV, hI, scoresOk);
}

/* ------------------------------------------- phase A with the obstacles STREAMED in tiles ----
 * Exhaustive test when the obstacle set exceeds the shared-memory staging budget (BASELINE config 3).  The obstacle
 * array cycles through two 16 KB shared-memory tiles filled by the TMA engine (cp.async.bulk + mbarrier); the CTA takes
 * WARPS consecutive chunks at a time and all its warps walk the tile sequence together: wait(full) -> one pass of every
 * edge over the tile (edge_tile_pass) -> the last warp to finish with a buffer re-arms it and issues the next load, so
 * the copy of tile t+2 overlaps the tests against tile t+1. */
struct TileStream {
    float4* buf0; uint64_t* full; int* cnt;        /* two tiles back to back; full[2] mbarriers; cnt[2] warps done with a tile */
    const float4* obstacles; int T;                /* global array padded to T * STREAM_TILE entries */
    unsigned g;                                    /* tiles consumed so far by this CTA (uniform over its threads) */
};

__device__ __forceinline__ void tile_issue(const TileStream& ts, unsigned g) {
    const int b = (int)(g & 1u);
    KGMT_CHECK_RANGE(209, (int)(g % (unsigned)ts.T), ts.T);
    mbar_expect_tx(&ts.full[b], STREAM_TILE * 16u);
    bulk_g2s(ts.buf0 + b * STREAM_TILE, ts.obstacles + (size_t)(g % (unsigned)ts.T) * STREAM_TILE, STREAM_TILE * 16u, &ts.full[b]);
}

template <bool RECORD>
__device__ __forceinline__ void stream_group(const KArgs& A, const IterView& it, const DynParams& dyn, TileStream& ts,
                                             int base, int lane, int warp, int* hV, int* hI, bool& scoresOk) {
    const int c = base + warp;
    const bool chunkLive = c < it.numChunks;
    ChunkCand cc = chunkLive ? chunk_setup(A, it, c, lane) : ChunkCand{};
    const float dt = __fdiv_rn(cc.u.duration, (float)dyn.numDisc);
    const float tanS = tanf(cc.u.steering);
    const float4 s0 = cc.x;
    EdgeExit e{s0, 0, false};
    for (int t = 0; t < ts.T; ++t) {
        const int b = (int)(ts.g & 1u);
        mbar_wait(&ts.full[b], (ts.g >> 1) & 1u);
        if (chunkLive && cc.live) {
            if (t == 0) edge_tile_pass<true>(s0, cc.u, dyn, dt, tanS, ts.buf0 + b * STREAM_TILE, STREAM_TILE, e);
            else if (e.step > 0) edge_tile_pass<false>(s0, cc.u, dyn, dt, tanS, ts.buf0 + b * STREAM_TILE, STREAM_TILE, e);
        }
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            if (atomicAdd(&ts.cnt[b], 1) == WARPS - 1) {       /* every warp is done reading this buffer: refill it */
                ts.cnt[b] = 0;
                __threadfence_block();
                tile_issue(ts, ts.g + 2u);
            }
        }
        ts.g += 1u;
    }
    if (chunkLive) {
        cc.x = e.s; cc.valid = e.valid;
        chunk_finish<RECORD, false>(A, it, cc, c, lane, hV, hI, scoresOk);
    }
}

/* position of the r-th (0-based) set bit of m: five popcount halvings */
__device__ __forceinline__ int nth_set_bit(unsigned m, int r) {
    int pos = 0;
#pragma unroll
    for (int w = 16; w > 0; w >>= 1) {
        const unsigned lo = m & ((1u << w) - 1u);
        const int c = __popc(lo);
        if (r >= c) { r -= c; m >>= w; pos += w; } else { m = lo; }
    }
    return pos;
}

/* -------------------------------------------------------- phase B: one scan block ------
 * updateG (KGMT.cu:555-591) for the accepted candidates of 256 consecutive chunks.  Thread t
 * loads the ballot of chunk blk*256+t; a CTA-wide scan of the popcounts orders the rows; each
 * warp then moves the rows of its 32 chunks with all lanes busy: row q of the warp belongs to
 * the chunk found by a 5-step shuffle search over the inclusive counts.
 * base = accepted candidates in all earlier blocks. */
__device__ __forceinline__ void insert_block(const KArgs& A, const IterView& it, int blk, int base, int* sScan /* [WARPS] */,
                                             int slice = 0, int nslices = 1, bool haveMask = false, unsigned mask0 = 0u) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = blk * BLK_CHUNKS + tid;
    const unsigned mask = haveMask ? mask0 : ((c < it.numChunks) ? __ldcg(&it.chunkMask[c]) : 0u);
    const int cnt = __popc(mask);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    __syncthreads();
    if (lane == 31) sScan[warp] = incl;
    __syncthreads();
    int warpBase = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) if (w < warp) warpBase += sScan[w];
    const int W = __shfl_sync(0xffffffffu, incl, 31);          /* rows of this warp's 32 chunks */
    const int c0 = blk * BLK_CHUNKS + warp * 32;
    const int dst0 = it.treeSize + base + warpBase;
    for (int q0 = 32 * slice; q0 < W; q0 += 32 * nslices) {
        const int q = q0 + lane;
        int i = 0;                                             /* number of chunks whose inclusive count <= q */
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            const int v = __shfl_sync(0xffffffffu, incl, i + step - 1);
            if (v <= q) i += step;
        }
        i = min(i, 31);
        const unsigned m = __shfl_sync(0xffffffffu, mask, i);
        const int excl = __shfl_sync(0xffffffffu, incl - cnt, i);
        if (q < W) {
            const int r = q - excl;
            const int bit = nth_set_bit(m, r);
            const int ci = c0 + i;
            const int slot = ci * CHUNK + bit;
            const float4 x = __ldcg(&it.stageState[ci * CHUNK + r]);
            const float4 u = __ldcg(&it.stageCtrl[ci * CHUNK + r]);
            const int dst = dst0 + q;
            KGMT_CHECK_RANGE(207, dst, A.maxTree); KGMT_CHECK_RANGE(208, ci * CHUNK + r, A.maxCand);
            A.treeState[dst] = x;
            A.treeCtrl[dst] = u;
            A.treeParent[dst] = it.parentOf ? __ldcg(&it.parentOf[slot]) : it.frontierStart + slot / it.children;
            if (slot == it.goalSlot) A.st->goalIdx = dst;
        }
    }
    /* announce the block: the next phase A starts without a barrier and waits on this counter */
    __threadfence();
    __syncthreads();
    if (tid == 0) atomicAdd(&A.st->insertDone, 1);
}

/* the group of CTAs that plans one query together */
struct GridGroup {
    cg::grid_group g; int rank, size;
    __device__ GridGroup() : g(cg::this_grid()), rank((int)blockIdx.x), size((int)gridDim.x) {}
    __device__ __forceinline__ void sync() { g.sync(); }
};

/* ------------------------------------------------------------------ the planner loop ---
 * Runs up to maxIters expansion iterations (1 = kgmt_expand_iteration, all = kgmt_plan) or until
 * the planner stops, on the CTAs of `grp` (all co-resident). */
/* the collision structure as the kernels see it, staged once per launch */
struct ColSet {
    CollideGridS gridS; CollideGrid gridG;
    CollideSmemAll allS, allG;
    int* hV; int* hI;                       /* per-CTA R1 histograms (shared memory) */
    TileStream stream;                      /* COL_BRUTE_STREAM only */
    float W, H; int numDisc;                /* workspace size and step count, pinned in registers */
};

/* carve dynamic shared memory [R1 histograms][collision data] and stage the collision structure into it with
 * bulk async copies (TMA engine) completing on an mbarrier.  All threads of the CTA call. */
template <int COL>
__device__ __forceinline__ ColSet stage_collision(const KArgs& A, unsigned char* smem_raw, uint64_t* sBar) {
    const int tid = threadIdx.x;
    ColSet cs;
    cs.stream = TileStream{};
    /* the grid resolution, the workspace size and the step count take a round trip through shared memory (pin_u32) */
    __shared__ uint32_t sPin[8];
    if (tid == 0) {
        sPin[0] = (uint32_t)A.cullC; sPin[1] = __float_as_uint(A.cullInvX); sPin[2] = __float_as_uint(A.cullInvY);
        sPin[3] = __float_as_uint(A.W); sPin[4] = __float_as_uint(A.H); sPin[5] = (uint32_t)A.numDisc;
    }
    __syncthreads();
    cs.W = __uint_as_float(lds_pinned(&sPin[3])); cs.H = __uint_as_float(lds_pinned(&sPin[4])); cs.numDisc = (int)lds_pinned(&sPin[5]);
    cs.hV = reinterpret_cast<int*>(smem_raw);
    cs.hI = cs.hV + (A.useHist ? A.c1 : 0);
    unsigned char* colBase = smem_raw + (A.useHist ? ((2 * A.c1 * 4 + 15) & ~15) : 0);
    const float4* sObs = nullptr; const int* sCellStart = nullptr; const float4* sItems = nullptr;
    if (COL == COL_GRID_SMEM || COL == COL_BRUTE_SMEM) {
        uint32_t bytes0 = 0, bytes1 = 0;
        if (COL == COL_GRID_SMEM) { bytes0 = (uint32_t)A.cellStartInts * 4u; bytes1 = (uint32_t)A.numItems * 16u; }
        else                      { bytes0 = (uint32_t)A.K * 16u; }
        if (tid == 0) { mbar_init(sBar, 1); mbar_fence_init(); }
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(sBar, bytes0 + bytes1);
            if (COL == COL_GRID_SMEM) {
                bulk_g2s_chunked(colBase, A.cellStart, bytes0, sBar);
                bulk_g2s_chunked(colBase + bytes0, A.cellItems, bytes1, sBar);
            } else {
                bulk_g2s_chunked(colBase, A.obstacles, bytes0, sBar);
            }
        }
        mbar_wait(sBar, 0);
        if (COL == COL_GRID_SMEM) {
            sCellStart = reinterpret_cast<const int*>(colBase);
            sItems = reinterpret_cast<const float4*>(colBase + bytes0);
        } else {
            sObs = reinterpret_cast<const float4*>(colBase);
        }
    }
    if (COL == COL_BRUTE_STREAM) {
        /* two tiles + two 'full' mbarriers + two arrival counters; the first two loads go out at once */
        __shared__ __align__(8) uint64_t sFull[2];
        __shared__ int sCnt[2];
        TileStream& ts = cs.stream;
        ts.buf0 = reinterpret_cast<float4*>(colBase);
        ts.full = sFull; ts.cnt = sCnt; ts.obstacles = A.obstacles; ts.T = A.obsTile; ts.g = 0u;
        if (tid == 0) { mbar_init(&sFull[0], 1); mbar_init(&sFull[1], 1); sCnt[0] = 0; sCnt[1] = 0; mbar_fence_init(); }
        __syncthreads();
        if (tid == 0) { tile_issue(ts, 0u); tile_issue(ts, 1u); }
    }
    {
        /* the two shared-window addresses of the staged CSR go through a warp shuffle: a value that comes out of a
         * convergent operation cannot be re-derived inside the (divergent) row loop, so it stays in a register instead of
         * being recomputed from the kernel parameters and special registers in every trip */
        uint32_t sa = sCellStart ? smem_u32(sCellStart) : 0u, ia = sItems ? smem_u32(sItems) : 0u;
        sa = pin_u32(sa); ia = pin_u32(ia);
        cs.gridS = CollideGridS{sa, ia, (int)lds_pinned(&sPin[0]), __uint_as_float(lds_pinned(&sPin[1])),
                                __uint_as_float(lds_pinned(&sPin[2])), A.cellStartInts, A.numItems};
    }
    cs.gridG = CollideGrid{A.cellStart, A.cellItems, A.cullC, A.cullInvX, A.cullInvY, A.cellStartInts, A.numItems};
    cs.allS = CollideSmemAll{sObs, A.K};
    cs.allG = CollideSmemAll{A.obstacles, A.K};
    return cs;
}

/* a CTA must not exit with bulk copies still landing in its shared memory: wait for the two tiles in flight */
__device__ __forceinline__ void stream_drain(const TileStream& ts) {
    mbar_wait(&ts.full[ts.g & 1u], (ts.g >> 1) & 1u);
    mbar_wait(&ts.full[(ts.g + 1u) & 1u], ((ts.g + 1u) >> 1) & 1u);
}

template <int COL, bool RECORD, class Group>
__device__ void run_plan(const KArgs& A, int maxIters, Group& grp, ColSet& cs) {
    __shared__ int sRed[2 * WARPS];
    __shared__ float sP[1024];
    __shared__ int sBlkPre[BLK_PRE_CAP + 1];
    __shared__ DevState S;
    __shared__ int sAccepted;
    __shared__ unsigned long long sGoalBest;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    DevState* st = A.st;
    int* hV = cs.hV; int* hI = cs.hI;
    /* workspace size and step count pinned in registers (see stage_collision) */
    /* (not for the tile-streamed back end: its edge passes are re-run per tile and the three extra live registers cost
     * 17 % there — 2.41 s against 2.82 s per config-3 plan) */
    const DynParams dyn = (COL == COL_BRUTE_STREAM) ? DynParams{A.W, A.H, A.L, A.numDisc} : DynParams{cs.W, cs.H, A.L, cs.numDisc};
    const CollideGridS& colGridS = cs.gridS; const CollideGrid& colGridG = cs.gridG;
    const CollideSmemAll& colAllS = cs.allS; const CollideSmemAll& colAllG = cs.allG;

    /* every CTA keeps its own copy of the planner scalars */
    __syncthreads();                       /* a previous call's readers of S are done */
    if (tid < COPIED_WORDS) reinterpret_cast<int*>(&S)[tid] = __ldcg(reinterpret_cast<const int*>(st) + tid);
    __syncthreads();

    const int totalWarps = grp.size * WARPS;
    const int gw = grp.rank * WARPS + warp;
    const int scoreRank = grp.size - 1;
    const int resetRank = grp.size > 1 ? grp.size - 2 : 0;
    /* insertion items go to the CTAs from the top of the group down (below the two housekeeping CTAs): the first chunks of
     * an iteration belong to the low ranks, which start them at once from the staging rows */
    const int insertFirst = (3 * grp.size - 3 - grp.rank) % grp.size;
    StagedParents prev{sBlkPre, 0, 0, nullptr, nullptr, nullptr};
    bool prevOk = false;                   /* sBlkPre / prev describe the iteration before the one about to run */

    for (int iter = 0; iter < maxIters; ++iter) {
        if (S.stop != STOP_RUNNING) break;
        IterView it = make_view(A, S);
        const int logRow = S.iterationsDone;
        auto stamp = [&](int colIdx) {
            if (A.iterLog && grp.rank == 0 && tid == 0 && logRow < 255) {
                unsigned long long t;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                A.iterLog[8 * logRow + colIdx] = t;
            }
        };
        stamp(2);

        /* ---- phase A: first chunk by position, further chunks by ticket (prefetched behind the current chunk) */
        if (A.useHist) for (int c = tid; c < 2 * A.c1; c += TILE) hV[c] = 0;
        __syncthreads();
        if (COL == COL_BRUTE_STREAM) {
            /* CTA-level tickets: WARPS consecutive chunks per fetch (the counter starts at totalWarps, as below) */
            __shared__ int sBase;
            bool scoresOk = false;
            unsigned* ticket = A.ticket + (it.itr % 3);
            int base = grp.rank * WARPS;
            if (base < it.numChunks) wait_ge(&st->insertDone, S.blocksTotal);
            while (base < it.numChunks) {
                if (tid == 0) sBase = (int)atomicAdd(ticket, (unsigned)WARPS);
                stream_group<RECORD>(A, it, dyn, cs.stream, base, lane, warp, hV, hI, scoresOk);
                __syncthreads();
                base = sBase;
                __syncthreads();
            }
        } else {
            bool scoresOk = false;
            unsigned* ticket = A.ticket + (it.itr % 3);
            int c = gw;
            int t = 0;
            /* parents were inserted by the previous phase B, possibly still running on other CTAs: until its last block is
             * announced, take the parent from the staging rows (staged_parent) when the chunk has a single one */
            bool inserted = false;
            const bool canStage = prevOk && it.children == CHUNK;
            while (c < it.numChunks) {
                if (lane == 0) t = (int)atomicAdd(ticket, 1u);
                bool staged = false;
                if (!inserted) {
                    int done = (lane == 0) ? ld_relaxed_s32(&st->insertDone) : 0;       /* one lane reads, all agree */
                    done = __shfl_sync(0xffffffffu, done, 0);
                    if (done >= S.blocksTotal) { fence_acq_rel(); inserted = true; }
                    else if (canStage) staged = true;
                    else { wait_ge(&st->insertDone, S.blocksTotal); inserted = true; }
                }
                if (COL == COL_GRID_SMEM)        expand_chunk<CollideGridS, RECORD>(A, it, dyn, colGridS, c, lane, hV, hI, scoresOk, staged, prev);
                else if (COL == COL_GRID_GLOBAL) expand_chunk<CollideGrid, RECORD>(A, it, dyn, colGridG, c, lane, hV, hI, scoresOk, staged, prev);
                else if (COL == COL_BRUTE_SMEM)  expand_chunk<CollideSmemAll, RECORD>(A, it, dyn, colAllS, c, lane, hV, hI, scoresOk, staged, prev);
                else if (COL == COL_BRUTE_GLOBAL) expand_chunk<CollideSmemAll, RECORD>(A, it, dyn, colAllG, c, lane, hV, hI, scoresOk, staged, prev);
                c = __shfl_sync(0xffffffffu, t, 0);
            }
        }
        __syncthreads();
        if (A.useHist) {               /* R1 = R1Valid + R1Invalid increments, KGMT.cu:392,406,409 */
            for (int c = tid; c < A.c1; c += TILE) {
                const int v = hV[c], iv = hI[c];
                if (v | iv) {
                    atomicAdd(&A.R1[c], v + iv);
                    if (v) { atomicAdd(&A.R1Valid[c], v); A.R1Avail[c] = 1; }
                    if (iv) atomicAdd(&A.R1Invalid[c], iv);
                }
            }
        }
        stamp(3);
        /* portfolio race: ONE thread looks at the 'a peer has solved' word before the barrier, so that every CTA takes the
         * same decision after it */
        if (A.raceId > 0 && grp.rank == 0 && tid == 0)
            st->peerSolved[it.itr & 1] = (ld_acquire_sys_s32(A.raceFlags[A.raceRank]) == A.raceId) ? 1 : 0;
        grp.sync();                    /* the one barrier: all ballots, block sums, maps and the goal minimum are final */
        stamp(4);

        /* ---- every CTA advances the planner scalars on its own copy.  The same pass over the block sums leaves their
         * exclusive prefix in shared memory — the insertion base of this CTA's items and the key of staged_parent in the
         * next phase A — and the ballots of this CTA's first insertion item are fetched alongside: one L2 round trip
         * after the barrier instead of three (the items and the split depend only on the iteration's shape). */
        const int numBlocks = (it.numChunks + BLK_CHUNKS - 1) / BLK_CHUNKS;
        const int split = insert_split(numBlocks, A.totalWarps / WARPS);
        const int blk0 = (insertFirst < numBlocks * split) ? insertFirst / split : -1;
        const bool preOk = numBlocks <= BLK_PRE_CAP;
        unsigned mask0 = 0u;
        int base0 = 0;
        {
            if (blk0 >= 0) {
                const int c0 = blk0 * BLK_CHUNKS + tid;
                if (c0 < it.numChunks) mask0 = __ldcg(&it.chunkMask[c0]);
            }
            unsigned long long gbNow = ~0ull; int peerNow = 0;     /* thread 0: in flight with the block sums */
            if (tid == 0) {
                gbNow = *(volatile unsigned long long*)&st->goalBest[it.itr & 1];
                if (A.raceId > 0) peerNow = *(volatile int*)&st->peerSolved[it.itr & 1];
            }
            int accepted;
            if (preOk) {
                const int v = (tid < numBlocks) ? __ldcg(&it.blockSum[tid]) : 0;
                int incl = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
                __syncthreads();               /* readers of sRed / sBlkPre from the previous round are done */
                if (lane == 31) sRed[warp] = incl;
                __syncthreads();
                int wb = 0, tot = 0;
#pragma unroll
                for (int w2 = 0; w2 < WARPS; ++w2) { const int x = sRed[w2]; if (w2 < warp) wb += x; tot += x; }
                sBlkPre[tid] = wb + incl - v;
                if (tid == 0) sBlkPre[BLK_PRE_CAP] = tot;
                accepted = tot;
                __syncthreads();
                if (blk0 >= 0) base0 = sBlkPre[blk0];
            } else {
                int mine = 0, mineBase = 0;
                for (int b = tid; b < numBlocks; b += TILE) {
                    const int v = __ldcg(&it.blockSum[b]);
                    mine += v;
                    if (b < blk0) mineBase += v;
                }
                const int2 sums = block_sum2(mine, mineBase, sRed);
                accepted = sums.x;
                base0 = sums.y;
            }
            prev.numBlocks = numBlocks; prev.numChunks = it.numChunks;
            prev.mask = it.chunkMask; prev.stageState = it.stageState; prev.stageCtrl = it.stageCtrl;
            prevOk = preOk;
            if (tid == 0) {
                sGoalBest = gbNow;
                const bool hadGoal = S.costToGoal != 0.0f;
                advance_state(A, S, accepted, sGoalBest, peerNow);
                sAccepted = (!hadGoal && S.costToGoal != 0.0f) ? S.goalSlot : -1;
                if (A.raceId > 0 && grp.rank == 0 && S.stop == STOP_SOLVED) {
                    /* first solution: tell every other GPU of the race (system-scope release over NVLink) */
                    for (int p = 0; p < A.raceWorld; ++p)
                        if (p != A.raceRank) st_release_sys_s32(A.raceFlags[p], A.raceId);
                }
                if (grp.rank == 0 && A.iterLog && logRow < 255) {
                    unsigned long long t;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                    A.iterLog[8 * logRow] = t;
                    A.iterLog[8 * logRow + 1] = ((unsigned long long)(unsigned)it.M << 32) | (unsigned)accepted;
                }
            }
            __syncthreads();
            it.goalSlot = sAccepted;
        }

        /* ---- housekeeping spread over otherwise idle CTAs (all of it is consumed after a later barrier or flag) */
        if (grp.rank == scoreRank && S.stop == STOP_RUNNING) {
            /* next iteration's R1 scores from the now-final maps, then publish them */
            scores_block(A, sP, A.R1Score[S.itr & 1]);
            __threadfence();
            __syncthreads();
            if (tid == 0) st_release_s32(&st->scoreReady, S.itr);
        }
        if (grp.rank == resetRank) {
            /* recycle the ticket and block sums last used by the previous iteration (next used two iterations from now) */
            const int r = (it.itr + 2) % 3;
            for (int b = tid; b < A.blocksCap; b += TILE) A.blockSum[(size_t)r * A.blocksCap + b] = 0;
            if (tid == 0) A.ticket[r] = (unsigned)totalWarps;
        }
        if (grp.rank == 0 && tid < COPIED_WORDS && tid != THRESHOLD_WORD)   /* for the host and for the next launch */
            reinterpret_cast<int*>(st)[tid] = reinterpret_cast<const int*>(&S)[tid];

        /* ---- phase B: ordered insertion, scan-block slices strided over the CTAs (first item: base and ballots already here) */
        for (int v = insertFirst; v < numBlocks * split; v += grp.size) {
            const int blk = v / split;
            if (v == insertFirst) {
                insert_block(A, it, blk, base0, sRed, v - blk * split, split, true, mask0);
            } else {
                int base;
                if (preOk) base = sBlkPre[blk];
                else {
                    int mine = 0;
                    for (int b = tid; b < blk; b += TILE) mine += __ldcg(&it.blockSum[b]);
                    base = block_sum(mine, sRed);
                }
                insert_block(A, it, blk, base, sRed, v - blk * split, split);
            }
        }
        stamp(5);
    }
}

#ifndef KGMT_EXPAND_MIN_CTAS
#define KGMT_EXPAND_MIN_CTAS 3        /* register budget of the cooperative kernel: 65536 / (256 * 3) = 85 per thread */
#endif
#ifdef KGMT_EXPAND_MAXNREG
#define KGMT_EXPAND_BOUNDS __maxnreg__(KGMT_EXPAND_MAXNREG)
#else
#define KGMT_EXPAND_BOUNDS __launch_bounds__(TILE, KGMT_EXPAND_MIN_CTAS)
#endif
template <int COL, bool RECORD>
__global__ void KGMT_EXPAND_BOUNDS expand_kernel(const KArgs A, int maxIters) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t sBar;
    GridGroup grp;
    ColSet cs = stage_collision<COL>(A, smem_raw, &sBar);
    run_plan<COL, RECORD, GridGroup>(A, maxIters, grp, cs);
    if (COL == COL_BRUTE_STREAM) stream_drain(cs.stream);
}

/* ------------------------------------------------------- batched planning (config 4) ----
 * Many independent queries on one map: a thread-block CLUSTER (1, 2, 4 or 8 CTAs, hardware barrier) plans one
 * query at a time in its own workspace and pulls the next query from a ticket.  Same run_plan, same results as
 * kgmt_plan on the same (init, goal, seed). */
struct ClusterGroup {
    int rank, size;
    __device__ ClusterGroup() {
        cg::cluster_group c = cg::this_cluster();
        rank = (int)c.block_rank(); size = (int)c.num_blocks();
    }
    __device__ __forceinline__ void sync() {
        if (size == 1) __syncthreads(); else cg::this_cluster().sync();
    }
};

struct BatchArgs {
    KArgs base;                   /* map, parameters; the per-query pointers below override its arrays */
    int Q, numWorkspaces;
    const float4* initState; const float4* initCtrl; const float2* goalXY; const uint32_t* seeds;   /* [Q] */
    DevState* states;             /* [Q] results */
    float* paths; int* pathLen; int maxPath;                     /* [Q][maxPath][7], [Q]; null = no paths */
    int* queryTicket; int* wsQuery;                              /* [1], [numWorkspaces] */
    unsigned long long* launchT0;                                /* [1] globaltimer when the first CTA of the launch started */
    /* workspace strides (elements) */
    size_t treeStride, mapIntsStride, chunkStride, blockStride, stageStride;
    int* mapSlab;                 /* [ws][mapIntsStride]: R1,R1Valid,R1Invalid,R1Avail,R1Cov,R1Score x2 | R2,R2Valid,R2Invalid,R2Stamp */
    size_t c2;
};

__device__ void begin_block(const KArgs& A, float4 rootState, float4 rootCtrl, float* sP, int forceChildren);

#ifndef KGMT_BATCH_MIN_CTAS
#define KGMT_BATCH_MIN_CTAS 4         /* many small trees: occupancy (concurrent workspaces) matters more than registers */
#endif
template <int COL>
__global__ void __launch_bounds__(TILE, KGMT_BATCH_MIN_CTAS) batch_kernel(const BatchArgs B) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t sBar;
    __shared__ float sPb[1024];
    ClusterGroup grp;
    const int tid = threadIdx.x;
    const int ws = (int)blockIdx.x / grp.size;
    ColSet cs = stage_collision<COL>(B.base, smem_raw, &sBar);

    /* this workspace's arrays */
    KArgs A = B.base;
    A.treeState = B.base.treeState + (size_t)ws * B.treeStride;
    A.treeCtrl = B.base.treeCtrl + (size_t)ws * B.treeStride;
    A.treeParent = B.base.treeParent + (size_t)ws * B.treeStride;
    int* m = B.mapSlab + (size_t)ws * B.mapIntsStride;
    const size_t c1 = (size_t)A.c1, c2 = B.c2;
    A.R1 = m; m += c1; A.R1Valid = m; m += c1; A.R1Invalid = m; m += c1; A.R1Avail = m; m += c1; A.R1Cov = m; m += c1;
    A.R1Score[0] = reinterpret_cast<float*>(m); m += c1; A.R1Score[1] = reinterpret_cast<float*>(m); m += c1;
    A.R2 = m; m += c2; A.R2Valid = m; m += c2; A.R2Invalid = m; m += c2; A.R2Stamp = reinterpret_cast<unsigned*>(m);
    A.chunkMask = B.base.chunkMask + (size_t)ws * B.chunkStride;
    A.blockSum = B.base.blockSum + (size_t)ws * B.blockStride;
    A.ticket = B.base.ticket + (size_t)ws * 4;
    A.stageState = B.base.stageState + (size_t)ws * B.stageStride;
    A.stageCtrl = B.base.stageCtrl + (size_t)ws * B.stageStride;
    A.totalWarps = grp.size * WARPS;
    A.iterLog = nullptr;

    const int nThreads = grp.size * TILE, gtid = grp.rank * TILE + tid;
    if (blockIdx.x == 0 && tid == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        *B.launchT0 = t;
    }
    for (;;) {
        if (grp.rank == 0 && tid == 0) B.wsQuery[ws] = atomicAdd(B.queryTicket, 1);
        __threadfence();
        grp.sync();
        const int q = __ldcg(&B.wsQuery[ws]);
        if (q >= B.Q) break;
        A.st = B.states + q;
        A.goalX = B.goalXY[q].x; A.goalY = B.goalXY[q].y; A.seed = B.seeds[q];
        /* reset the workspace: maps and scan block sums (tree rows and staging are overwritten as they are used) */
        {
            int4* z = reinterpret_cast<int4*>(B.mapSlab + (size_t)ws * B.mapIntsStride);
            const size_t n4 = B.mapIntsStride / 4;
            for (size_t i = gtid; i < n4; i += nThreads) z[i] = make_int4(0, 0, 0, 0);
            for (size_t i = gtid; i < B.blockStride; i += nThreads) A.blockSum[i] = 0;
        }
        __threadfence();
        grp.sync();
        if (grp.rank == 0) begin_block(A, B.initState[q], B.initCtrl[q], sPb, 0);
        if (grp.rank == 0 && tid == 0) {        /* per-query device clock: start (pairsTested) and end (stepsDone) of its service */
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            A.st->pairsTested = t;
        }
        __threadfence();
        grp.sync();
        run_plan<COL, false, ClusterGroup>(A, 0x7fffffff, grp, cs);
        __threadfence();
        grp.sync();                             /* every insertion of the last iteration has landed */
        if (grp.rank == 0 && tid == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            A.st->stepsDone = t;
        }
        if (B.paths && grp.rank == 0 && tid == 0) {
            /* solution back-trace, root first */
            const DevState* r = A.st;
            int len = 0;
            const int goalIdx = *(volatile const int*)&r->goalIdx;
            if (*(volatile const int*)&r->stop == STOP_SOLVED && goalIdx >= 0) {
                for (int v = goalIdx; v >= 0 && len <= A.maxTree; v = __ldcg(&A.treeParent[v])) ++len;
                int at = len - 1;
                for (int v = goalIdx; v >= 0 && at >= 0; v = __ldcg(&A.treeParent[v]), --at) {
                    if (at < B.maxPath) {
                        const float4 x = __ldcg(&A.treeState[v]), u = __ldcg(&A.treeCtrl[v]);
                        float* o = B.paths + ((size_t)q * B.maxPath + at) * 7;
                        o[0] = x.x; o[1] = x.y; o[2] = x.z; o[3] = x.w; o[4] = u.x; o[5] = u.y; o[6] = u.z;
                    }
                }
            }
            B.pathLen[q] = len;
        }
        grp.sync();                             /* the workspace may be recycled */
    }
}

/* ------------------------------------------ sharded expansion (config 5) ----
 * The candidates of ONE iteration are split over G ranks (tree and maps replicated on every GPU):
 *   shard_expand   rank g runs stages 2-5a on its contiguous chunk range; counter increments go to a zeroed DELTA slab
 *                  (A.R1.., A.R2.. point into it), first-reached R2 cells are flagged in R2StampDelta; accepted rows are
 *                  ballot-compacted into the staging buffers exactly as in the single-GPU kernel
 *   shard_prefix   exclusive prefix of the rank's scan-block sums + its accepted count
 *   shard_pack     accepted rows, in candidate order, into the send buffer [state f4[cap] | ctrl f4[cap] | slot i32[cap]]
 *   -- NCCL: all-gather(counts), all-gather(rows), all-reduce SUM(delta slab) --
 *   shard_insert   every rank inserts all rows in rank-major (= global candidate) order: the tree, parent links and
 *                  costs are bit-identical to the single-GPU result
 *   shard_apply    maps += reduced deltas, stamps, R1Cov, R1Avail; the slab is zeroed for the next iteration
 *   shard_finalize goal node, planner scalars (advance_state), next iteration's scores */
__global__ void shard_reset_kernel(const KArgs A, int totalWarps) {
    const DevState* st = A.st;
    int* bs = A.blockSum + (size_t)(st->itr % 3) * A.blocksCap;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < A.blocksCap; b += gridDim.x * blockDim.x) bs[b] = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) A.ticket[3] = (unsigned)totalWarps;
}

template <int COL>
__global__ void __launch_bounds__(TILE) shard_expand_kernel(const KArgs A, int chunkLo, int chunkHi) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t sBar;
    __shared__ DevState S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const ColSet cs = stage_collision<COL>(A, smem_raw, &sBar);
    if (tid < COPIED_WORDS) reinterpret_cast<int*>(&S)[tid] = __ldcg(reinterpret_cast<const int*>(A.st) + tid);
    int* hV = cs.hV; int* hI = cs.hI;
    if (A.useHist) for (int c = tid; c < 2 * A.c1; c += TILE) hV[c] = 0;
    __syncthreads();
    if (S.stop != STOP_RUNNING) return;
    const IterView it = make_view(A, S);
    const DynParams dyn{cs.W, cs.H, A.L, cs.numDisc};      /* pinned in registers by stage_collision */
    bool scoresOk = true;                  /* the scores were produced by the previous commit, stream-ordered */
    int c = chunkLo + (int)blockIdx.x * WARPS + warp;
    int t = 0;
    const int hi = min(chunkHi, it.numChunks);
    while (c < hi) {
        if (lane == 0) t = (int)atomicAdd(&A.ticket[3], 1u);
        if (COL == COL_GRID_SMEM)        expand_chunk<CollideGridS, false, true>(A, it, dyn, cs.gridS, c, lane, hV, hI, scoresOk);
        else if (COL == COL_GRID_GLOBAL) expand_chunk<CollideGrid, false, true>(A, it, dyn, cs.gridG, c, lane, hV, hI, scoresOk);
        else if (COL == COL_BRUTE_SMEM)  expand_chunk<CollideSmemAll, false, true>(A, it, dyn, cs.allS, c, lane, hV, hI, scoresOk);
        else                             expand_chunk<CollideSmemAll, false, true>(A, it, dyn, cs.allG, c, lane, hV, hI, scoresOk);
        c = chunkLo + __shfl_sync(0xffffffffu, t, 0);
    }
    __syncthreads();
    if (A.useHist) {
        for (int r = tid; r < A.c1; r += TILE) {
            const int v = hV[r], iv = hI[r];
            if (v | iv) {
                atomicAdd(&A.R1[r], v + iv);
                if (v) atomicAdd(&A.R1Valid[r], v);
                if (iv) atomicAdd(&A.R1Invalid[r], iv);
            }
        }
    }
}

/* one CTA: prefix[b - blkLo] = accepted rows of the rank's scan blocks before b; *total = the rank's accepted count */
__global__ void __launch_bounds__(TILE) shard_prefix_kernel(const KArgs A, int blkLo, int blkHi, int* prefix, int* total) {
    __shared__ int sRed[WARPS];
    __shared__ int sCarry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int* bs = A.blockSum + (size_t)(A.st->itr % 3) * A.blocksCap;
    if (tid == 0) sCarry = 0;
    __syncthreads();
    for (int b0 = blkLo; b0 < blkHi; b0 += TILE) {
        const int b = b0 + tid;
        const int v = (b < blkHi) ? __ldcg(&bs[b]) : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) sRed[warp] = incl;
        __syncthreads();
        int wb = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) if (w < warp) wb += sRed[w];
        const int carry = sCarry;
        if (b < blkHi) prefix[b - blkLo] = carry + wb + incl - v;
        __syncthreads();
        if (tid == TILE - 1) sCarry = carry + wb + incl;
        __syncthreads();
    }
    if (tid == 0) *total = sCarry;
}

/* CTA per scan block: the block's accepted rows, in candidate order, into the send sections (same scan as insert_block) */
__global__ void __launch_bounds__(TILE) shard_pack_kernel(const KArgs A, int blkLo, int blkHi, const int* prefix,
                                                          float4* sendState, float4* sendCtrl, int* sendSlot) {
    __shared__ int sScan[WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const DevState* st = A.st;
    const int itr = st->itr, numChunks = st->numChunks;
    const unsigned* chunkMask = A.chunkMask + (size_t)(itr & 1) * A.chunksCap;
    const float4* stageState = A.stageState + (size_t)(itr & 1) * A.maxCand;
    const float4* stageCtrl = A.stageCtrl + (size_t)(itr & 1) * A.maxCand;
    for (int blk = blkLo + (int)blockIdx.x; blk < blkHi; blk += (int)gridDim.x) {
        const int c = blk * BLK_CHUNKS + tid;
        const unsigned mask = (c < numChunks) ? __ldcg(&chunkMask[c]) : 0u;
        const int cnt = __popc(mask);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        __syncthreads();
        if (lane == 31) sScan[warp] = incl;
        __syncthreads();
        int warpBase = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) if (w < warp) warpBase += sScan[w];
        const int W = __shfl_sync(0xffffffffu, incl, 31);
        const int c0 = blk * BLK_CHUNKS + warp * 32;
        const int dst0 = __ldcg(&prefix[blk - blkLo]) + warpBase;
        for (int q0 = 0; q0 < W; q0 += 32) {
            const int q = q0 + lane;
            int i = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const int v = __shfl_sync(0xffffffffu, incl, i + step - 1);
                if (v <= q) i += step;
            }
            i = min(i, 31);
            const unsigned m = __shfl_sync(0xffffffffu, mask, i);
            const int excl = __shfl_sync(0xffffffffu, incl - cnt, i);
            if (q < W) {
                const int r = q - excl;
                const int bit = nth_set_bit(m, r);
                const int ci = c0 + i;
                sendState[dst0 + q] = __ldcg(&stageState[ci * CHUNK + r]);
                sendCtrl[dst0 + q] = __ldcg(&stageCtrl[ci * CHUNK + r]);
                sendSlot[dst0 + q] = ci * CHUNK + bit;
            }
        }
    }
}

struct ShardCommit {
    const unsigned char* recv;    /* [world] segments of cap rows: state f4[cap] | ctrl f4[cap] | slot i32[cap] */
    int cap, world;
    int prefix[17];               /* prefix[g] = rows of ranks < g; prefix[world] = all accepted rows */
};

__global__ void __launch_bounds__(TILE) shard_insert_kernel(const KArgs A, const ShardCommit C) {
    const DevState* st = A.st;
    const int treeSize = st->treeSize, frontierStart = st->frontierStart, children = st->children;
    const int total = C.prefix[C.world];
    const size_t segBytes = (size_t)C.cap * 36;
    for (int q = blockIdx.x * TILE + threadIdx.x; q < total; q += gridDim.x * TILE) {
        int g = 0;
        while (q >= C.prefix[g + 1]) ++g;
        const int j = q - C.prefix[g];
        const unsigned char* seg = C.recv + (size_t)g * segBytes;
        const float4 x = reinterpret_cast<const float4*>(seg)[j];
        const float4 u = reinterpret_cast<const float4*>(seg + (size_t)C.cap * 16)[j];
        const int slot = reinterpret_cast<const int*>(seg + (size_t)C.cap * 32)[j];
        const int dst = treeSize + q;
        A.treeState[dst] = x;                                                      /* updateG, KGMT.cu:555-591 */
        A.treeCtrl[dst] = u;
        A.treeParent[dst] = frontierStart + slot / children;
        if (in_goal(x.x, x.y, A.goalX, A.goalY, A.goalR))
            atomicMin(&A.st->goalBest[st->itr & 1], ((unsigned long long)__float_as_uint(u.w) << 32) | (unsigned)q);   /* keyed by row */
    }
}

/* maps += all-reduced deltas (KGMT.cu:392-411 summed over the ranks); the slab is left zeroed.
 * delta layout: R1, R1Valid, R1Invalid, (R1Avail scratch) [c1] | R2, R2Valid, R2Invalid, R2StampDelta [c2] */
__global__ void shard_apply_kernel(const KArgs A, int* delta, size_t c2) {
    const size_t c1 = (size_t)A.c1;
    const unsigned stampNew = (unsigned)A.st->itr + 1u;
    const int nn = A.n * A.n;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    /* leave this iteration's scan block sums clean, as the cooperative kernel does (a later launch of it may follow) */
    int* bs = A.blockSum + (size_t)(A.st->itr % 3) * A.blocksCap;
    for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < (size_t)A.blocksCap; b += stride) bs[b] = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < c2; i += stride) {
        int* d = delta + 4 * c1;
        const int a = d[i], v = d[c2 + i], iv = d[2 * c2 + i], sd = d[3 * c2 + i];
        if (a) { A.R2[i] += a; d[i] = 0; }
        if (v) { A.R2Valid[i] += v; d[c2 + i] = 0; }
        if (iv) { A.R2Invalid[i] += iv; d[2 * c2 + i] = 0; }
        if (sd) {
            if (A.R2Stamp[i] == 0u) { A.R2Stamp[i] = stampNew; atomicAdd(&A.R1Cov[i / nn], 1); }
            d[3 * c2 + i] = 0;
        }
        if (i < c1) {
            const int a1 = delta[i], v1 = delta[c1 + i], i1 = delta[2 * c1 + i];
            if (a1) { A.R1[i] += a1; delta[i] = 0; }
            if (v1) { A.R1Valid[i] += v1; A.R1Avail[i] = 1; delta[c1 + i] = 0; }
            if (i1) { A.R1Invalid[i] += i1; delta[2 * c1 + i] = 0; }
            delta[3 * c1 + i] = 0;
        }
    }
}

__global__ void __launch_bounds__(TILE) shard_finalize_kernel(const KArgs A, const ShardCommit C) {
    __shared__ float sP[1024];
    __shared__ DevState S;
    const int tid = threadIdx.x;
    DevState* st = A.st;
    if (tid < COPIED_WORDS) reinterpret_cast<int*>(&S)[tid] = __ldcg(reinterpret_cast<const int*>(st) + tid);
    __syncthreads();
    const int treeSize0 = S.treeSize;
    if (tid == 0) {
        /* shard_insert keys the goal minimum by ROW (rank-major row order == candidate order, so ties break the same
         * way); turn it back into the (cost, candidate slot) key advance_state expects and record the tree index */
        unsigned long long gb = *(volatile unsigned long long*)&st->goalBest[S.itr & 1];
        if (gb != ~0ull && S.costToGoal == 0.0f) {
            const int q = (int)(unsigned)gb;
            int g = 0;
            while (q >= C.prefix[g + 1]) ++g;
            const int* slots = reinterpret_cast<const int*>(C.recv + (size_t)g * C.cap * 36 + (size_t)C.cap * 32);
            gb = (gb & 0xffffffff00000000ull) | (unsigned)slots[q - C.prefix[g]];
            st->goalIdx = treeSize0 + q;
        }
        advance_state(A, S, C.prefix[C.world], gb);
    }
    __syncthreads();
    if (S.stop == STOP_RUNNING) scores_block(A, sP, A.R1Score[S.itr & 1]);
    __syncthreads();
    if (tid < COPIED_WORDS && tid != THRESHOLD_WORD) reinterpret_cast<int*>(st)[tid] = reinterpret_cast<const int*>(&S)[tid];
    if (tid == 0) { st->scoreReady = S.itr; st->insertDone = S.blocksTotal; }
}

/* ------------------------- sharded expansion over PEER MEMORY (NVLink / NVSwitch, no NCCL) ----
 * The same split as above, but the exchange is done by the kernels themselves through pointers into the other GPUs'
 * memory (cudaIpc handles, or the same process in the tests): per iteration
 *   shard_expand + shard_prefix            local, as above (deltas into this rank's slab)
 *   peer_counts      one warp: writes this rank's accepted count into every peer's mailbox (system-scope release),
 *                    waits for every peer's count -> base = rows of lower ranks, total = rows of all ranks
 *   peer_pack        accepted rows of this rank, candidate order, written STRAIGHT INTO EVERY RANK'S TREE at
 *                    treeSize + base + local position (float4 state, float4 ctrl+cost, parent link): the all-gather and
 *                    the insertion are one pass of stores over NVLink
 *   peer_reduce      all-reduce of the counter deltas as reduce-scatter + all-gather over peer loads/stores: this rank
 *                    owns a contiguous 1/world of the cells, sums every rank's delta for them and writes the new counter
 *                    values (and first-reached stamps) into every rank's maps
 *   peer_barrier     one warp: system fence, posts this rank's goal candidate, waits for every peer -> global goal
 *   recount_cov, peer_finalize             local: R1Cov from the stamps, planner scalars, next scores
 * Replicas stay bit-identical to the single-GPU run for the same reason as with NCCL: rank-major = candidate order. */
constexpr int PEER_MAX = 16;
struct PeerMail { unsigned long long goal; int count; int seq1; int seq2; int pad; };      /* 24 B -> padded to 32 */
struct PeerPlan { unsigned long long goalLocal, goalGlobal; int base, total, err, pad; int counts[PEER_MAX];
                  int ready1, ready2, pad2[2]; };   /* fused kernel: exchange seq whose counts / goal are in (release/acquire, gpu scope) */
struct PeerArgs {
    int rank, world, seq;
    float4* treeState[PEER_MAX]; float4* treeCtrl[PEER_MAX]; int* treeParent[PEER_MAX];
    int* mapSlab[PEER_MAX]; int* delta[PEER_MAX]; PeerMail* mail[PEER_MAX];     /* [p] = rank p's arrays as mapped on this device */
    PeerPlan* plan;                                                             /* local */
};

/* wait until *p == want, at most ~5 s (a peer that never arrives must not hang the GPU); false on timeout */
__device__ __forceinline__ bool peer_wait(const int* p, int want) {
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    unsigned ns = 64;
    while (ld_acquire_sys_s32(p) != want) {
        __nanosleep(ns); if (ns < 2048) ns <<= 1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > 5000000000ull) return false;
    }
    return true;
}

__global__ void peer_counts_kernel(const PeerArgs P, const int* localTotal) {
    const int lane = threadIdx.x;
    PeerPlan* plan = P.plan;
    const int cnt = __ldcg(localTotal);
    if (lane == 0) { plan->goalLocal = ~0ull; plan->goalGlobal = ~0ull; }
    bool ok = true;
    int c = 0;
    if (lane < P.world) {
        PeerMail* m = P.mail[lane] + P.rank;
        m->count = cnt;
        __threadfence_system();
        st_release_sys_s32(&m->seq1, P.seq);
        const PeerMail* in = P.mail[P.rank] + lane;
        ok = peer_wait(&in->seq1, P.seq);
        c = *(volatile const int*)&in->count;
    }
    if (__any_sync(0xffffffffu, !ok)) { if (lane == 0) plan->err = 1; return; }
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane < P.world) plan->counts[lane] = c;
    if (lane == P.rank) plan->base = incl - c;
    if (lane == 31) plan->total = incl;
}

/* CTA per scan block: this rank's accepted rows into every rank's tree */
__global__ void __launch_bounds__(TILE) peer_pack_kernel(const KArgs A, const PeerArgs P, int blkLo, int blkHi, const int* prefix) {
    __shared__ int sScan[WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const DevState* st = A.st;
    const PeerPlan* plan = P.plan;
    if (*(volatile const int*)&plan->err) return;
    const int itr = st->itr, numChunks = st->numChunks, treeSize = st->treeSize, frontierStart = st->frontierStart, children = st->children;
    const int base = plan->base;
    const unsigned* chunkMask = A.chunkMask + (size_t)(itr & 1) * A.chunksCap;
    const float4* stageState = A.stageState + (size_t)(itr & 1) * A.maxCand;
    const float4* stageCtrl = A.stageCtrl + (size_t)(itr & 1) * A.maxCand;
    for (int blk = blkLo + (int)blockIdx.x; blk < blkHi; blk += (int)gridDim.x) {
        const int c = blk * BLK_CHUNKS + tid;
        const unsigned mask = (c < numChunks) ? __ldcg(&chunkMask[c]) : 0u;
        const int cnt = __popc(mask);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        __syncthreads();
        if (lane == 31) sScan[warp] = incl;
        __syncthreads();
        int warpBase = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) if (w < warp) warpBase += sScan[w];
        const int W = __shfl_sync(0xffffffffu, incl, 31);
        const int c0 = blk * BLK_CHUNKS + warp * 32;
        const int q0g = base + __ldcg(&prefix[blk - blkLo]) + warpBase;      /* global row index of this warp's first row */
        for (int q0 = 0; q0 < W; q0 += 32) {
            const int q = q0 + lane;
            int i = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const int v = __shfl_sync(0xffffffffu, incl, i + step - 1);
                if (v <= q) i += step;
            }
            i = min(i, 31);
            const unsigned m = __shfl_sync(0xffffffffu, mask, i);
            const int excl = __shfl_sync(0xffffffffu, incl - cnt, i);
            if (q < W) {
                const int r = q - excl;
                const int bit = nth_set_bit(m, r);
                const int ci = c0 + i;
                const int slot = ci * CHUNK + bit;
                const float4 x = __ldcg(&stageState[ci * CHUNK + r]);
                const float4 u = __ldcg(&stageCtrl[ci * CHUNK + r]);
                const int row = q0g + q;
                const int dst = treeSize + row;
                const int parent = frontierStart + slot / children;
                for (int p = 0; p < P.world; ++p) {                            /* updateG on every replica, KGMT.cu:555-591 */
                    P.treeState[p][dst] = x;
                    P.treeCtrl[p][dst] = u;
                    P.treeParent[p][dst] = parent;
                }
                if (in_goal(x.x, x.y, A.goalX, A.goalY, A.goalR))
                    atomicMin(&P.plan->goalLocal, ((unsigned long long)__float_as_uint(u.w) << 32) | (unsigned)row);
            }
        }
    }
    __threadfence_system();
}

/* this rank's share [lo, hi) of the delta index space: sum over the ranks, new values into every replica's maps.
 * delta layout as in shard_apply: R1, R1Valid, R1Invalid, scratch [c1] | R2, R2Valid, R2Invalid, first-reached [c2];
 * map slab layout: R1, R1Valid, R1Invalid, R1Avail, R1Cov, R1Score x2 [c1] | R2, R2Valid, R2Invalid, R2Stamp [c2] */
__global__ void peer_reduce_kernel(const KArgs A, const PeerArgs P, size_t c2, size_t lo, size_t hi) {
    if (*(volatile const int*)&P.plan->err) return;
    const size_t c1 = (size_t)A.c1;
    const unsigned stampNew = (unsigned)A.st->itr + 1u;
    const int* mine = P.mapSlab[P.rank];
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t idx = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < hi; idx += stride) {
        int sum = 0;
        for (int p = 0; p < P.world; ++p) sum += __ldcg(&P.delta[p][idx]);
        if (sum == 0) continue;
        size_t at; int val; size_t at2 = (size_t)-1; int val2 = 0;
        if (idx < 4 * c1) {
            const size_t k = idx / c1, i = idx - k * c1;
            if (k == 3) continue;                                  /* scratch section */
            at = k * c1 + i; val = mine[at] + sum;                 /* R1 / R1Valid / R1Invalid */
            if (k == 1) { at2 = 3 * c1 + i; val2 = 1; }            /* R1Avail, KGMT.cu:400 */
        } else {
            const size_t j = idx - 4 * c1, k = j / c2, i = j - k * c2;
            at = 7 * c1 + k * c2 + i;
            if (k == 3) { if (mine[at] != 0) continue; val = (int)stampNew; }      /* first reached in this iteration */
            else val = mine[at] + sum;
        }
        for (int p = 0; p < P.world; ++p) {
            P.mapSlab[p][at] = val;
            if (at2 != (size_t)-1) P.mapSlab[p][at2] = val2;
        }
    }
    __threadfence_system();
}

__global__ void peer_barrier_kernel(const PeerArgs P) {
    const int lane = threadIdx.x;
    PeerPlan* plan = P.plan;
    if (*(volatile const int*)&plan->err) return;
    __threadfence_system();
    bool ok = true;
    unsigned long long g = ~0ull;
    if (lane < P.world) {
        PeerMail* m = P.mail[lane] + P.rank;
        m->goal = *(volatile unsigned long long*)&plan->goalLocal;
        __threadfence_system();
        st_release_sys_s32(&m->seq2, P.seq);
        const PeerMail* in = P.mail[P.rank] + lane;
        ok = peer_wait(&in->seq2, P.seq);
        g = *(volatile const unsigned long long*)&in->goal;
    }
    if (__any_sync(0xffffffffu, !ok)) { if (lane == 0) plan->err = 1; return; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, g, o); g = t < g ? t : g; }
    if (lane == 0) plan->goalGlobal = g;
}

__global__ void __launch_bounds__(TILE) peer_finalize_kernel(const KArgs A, const PeerArgs P) {
    __shared__ float sP[1024];
    __shared__ DevState S;
    const int tid = threadIdx.x;
    DevState* st = A.st;
    const PeerPlan* plan = P.plan;
    if (*(volatile const int*)&plan->err) return;
    if (tid < COPIED_WORDS) reinterpret_cast<int*>(&S)[tid] = __ldcg(reinterpret_cast<const int*>(st) + tid);
    /* leave this iteration's scan block sums clean (see shard_apply_kernel) */
    int* bs = A.blockSum + (size_t)(__ldcg(&st->itr) % 3) * A.blocksCap;
    for (int b = tid; b < A.blocksCap; b += TILE) bs[b] = 0;
    __syncthreads();
    const int treeSize0 = S.treeSize;
    if (tid == 0) {
        const unsigned long long gb = plan->goalGlobal;          /* (cost bits << 32) | global row: ties break in candidate order */
        if (gb != ~0ull && S.costToGoal == 0.0f) st->goalIdx = treeSize0 + (int)(unsigned)gb;
        advance_state(A, S, plan->total, gb);
    }
    __syncthreads();
    if (S.stop == STOP_RUNNING) scores_block(A, sP, A.R1Score[S.itr & 1]);
    __syncthreads();
    if (tid < COPIED_WORDS && tid != THRESHOLD_WORD) reinterpret_cast<int*>(st)[tid] = reinterpret_cast<const int*>(&S)[tid];
    if (tid == 0) { st->scoreReady = S.itr; st->insertDone = S.blocksTotal; }
}

/* ------------------ sharded PLANS in one persistent kernel per rank: compute + exchange fused ------
 * The multi-launch sequence above costs nine launches and two host-visible round trips per iteration.  Here ONE
 * cooperative launch per rank runs any number of iterations (kgmt_peer_plan: the whole plan) and does the exchange
 * itself over peer memory (NVLink / NVSwitch), with the grid barrier as the only local synchronisation:
 *
 *   phase A        stages 2-5a on this rank's contiguous range of scan blocks (chunks by ticket inside the range);
 *                  counter increments into this rank's delta slab, accepted rows ballot-compacted into staging
 *   grid barrier 1
 *   counts         CTA 0 / warp 0: system fence, this rank's accepted count into every peer's mailbox (release.sys), wait
 *                  for every peer's count -> base row, total; released to the other CTAs through plan->ready1
 *   pack           accepted rows of this rank, candidate order, stored STRAIGHT INTO EVERY RANK'S TREE at
 *                  treeSize + base + local position — all-gather and insertion are one pass of stores over NVLink
 *   reduce         this rank's 1/world share of the counter cells: sum of every rank's delta (peer loads), new counter
 *                  values and first-reached stamps into every rank's maps (reduce-scatter + all-gather, peer stores)
 *   grid barrier 2
 *   goal           CTA 0 / warp 0: system fence, this rank's goal candidate to every peer, wait for every peer ->
 *                  every pack / reduce store of every rank has landed here; global goal minimum; plan->ready2
 *   finish         all CTAs: zero the delta slab, recount R1Cov from the stamps, advance the planner scalars (every
 *                  CTA on its own copy, as in run_plan)
 *   grid barrier 3, then one CTA scores the next iteration while the others already run its phase A
 *
 * A peer is never more than one exchange ahead (it needs this rank's count to finish exchange i + 1, and this rank
 * posts that only after it has consumed everything of exchange i), so one mailbox slot per peer suffices; the sequence
 * numbers only grow.  A peer that does not arrive within 5 s sets plan->err and every CTA leaves the loop together.
 * Replicas stay bit-identical to a single-GPU run: rank-major row order == candidate order (tests/test_gpu_sharded.py). */
__device__ __forceinline__ bool peer_wait_ge(const int* p, int want) {
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    unsigned ns = 32;
    while ((int)(ld_acquire_sys_s32(p) - want) < 0) {
        __nanosleep(ns); if (ns < 1024) ns <<= 1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > 5000000000ull) return false;
    }
    return true;
}
/* every thread of the CTA waits until the local word has reached `want`.  The word is written by CTA 0 / warp 0 with a
 * gpu-scope release AFTER that warp acquired the peers' words at system scope, so the chain
 *     peer stores -> peer release.sys -> CTA 0 acquire.sys -> CTA 0 release.gpu -> this acquire.gpu -> CTA barrier
 * orders the peers' stores before every thread of this CTA (causality is transitive across scopes); peer data is read
 * with ld.cg, never from L1.  System-scope fences are kept to ONE thread per hand-shake: executed by every thread (or by
 * one thread of each of 444 CTAs) they cost ~10 us per hand-shake on B200 (profiles/r02d_fused_timeline.txt). */
__device__ __forceinline__ void cta_wait_flag(const int* p, int want) {
    if (threadIdx.x == 0) {
        unsigned ns = 32;
        while ((int)(ld_acquire_s32(p) - want) < 0) { __nanosleep(ns); if (ns < 256) ns <<= 1; }
    }
    __syncthreads();
}

template <int COL>
__global__ void __launch_bounds__(TILE, 3) expand_sharded_kernel(const KArgs A, const KArgs As, const PeerArgs P, int maxIters, int seq0, size_t c2) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t sBar;
    __shared__ int sRed[WARPS];
    __shared__ int sScan[WARPS];
    __shared__ float sP[1024];
    __shared__ DevState S;
    __shared__ int sGoalSlot;
    cg::grid_group grid = cg::this_grid();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nCta = (int)gridDim.x, cta = (int)blockIdx.x;
    ColSet cs = stage_collision<COL>(A, smem_raw, &sBar);
    int* hV = cs.hV; int* hI = cs.hI;
    DevState* st = A.st;
    PeerPlan* plan = P.plan;
    const DynParams dyn{cs.W, cs.H, A.L, cs.numDisc};      /* pinned in registers by stage_collision */
    const int rank = P.rank, world = P.world;
    if (tid < COPIED_WORDS) reinterpret_cast<int*>(&S)[tid] = __ldcg(reinterpret_cast<const int*>(st) + tid);
    __syncthreads();
    const int totalWarps = nCta * WARPS;
    const int gw = cta * WARPS + warp;
    const size_t c1 = (size_t)A.c1;
    const size_t deltaInts = 4 * c1 + 4 * c2;
    const int nn = A.n * A.n;

    for (int iter = 0; iter < maxIters; ++iter) {
        if (S.stop != STOP_RUNNING) break;
        const int seq = seq0 + iter + 1;
        IterView it = make_view(A, S);
        const int logRow = S.iterationsDone;
        auto stamp = [&](int colIdx) {        /* kgmt_iteration_log: CTA 0's clock at the phase boundaries */
            if (A.iterLog && cta == 0 && tid == 0 && logRow < 255) {
                unsigned long long t;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                A.iterLog[8 * logRow + colIdx] = t;
            }
        };
        stamp(2);
        const int numBlocks = (it.numChunks + BLK_CHUNKS - 1) / BLK_CHUNKS;
        const int bBase = numBlocks / world, bRem = numBlocks % world;
        const int bLo = rank * bBase + min(rank, bRem), bHi = bLo + bBase + (rank < bRem ? 1 : 0);
        const int cLo = bLo * BLK_CHUNKS, cHi = min(bHi * BLK_CHUNKS, it.numChunks);

        /* ---- phase A on [cLo, cHi): first chunk by position, the rest by ticket */
        if (A.useHist) for (int c = tid; c < 2 * A.c1; c += TILE) hV[c] = 0;
        __syncthreads();
        {
            bool scoresOk = false;
            unsigned* ticket = A.ticket + (it.itr % 3);
            int c = cLo + gw, t = 0;
            while (c < cHi) {
                if (lane == 0) t = (int)atomicAdd(ticket, 1u);
                if (COL == COL_GRID_SMEM)        expand_chunk<CollideGridS, false, true>(As, it, dyn, cs.gridS, c, lane, hV, hI, scoresOk);
                else if (COL == COL_GRID_GLOBAL) expand_chunk<CollideGrid, false, true>(As, it, dyn, cs.gridG, c, lane, hV, hI, scoresOk);
                else if (COL == COL_BRUTE_SMEM)  expand_chunk<CollideSmemAll, false, true>(As, it, dyn, cs.allS, c, lane, hV, hI, scoresOk);
                else                             expand_chunk<CollideSmemAll, false, true>(As, it, dyn, cs.allG, c, lane, hV, hI, scoresOk);
                c = cLo + __shfl_sync(0xffffffffu, t, 0);
            }
        }
        __syncthreads();
        if (A.useHist) {
            for (int r = tid; r < A.c1; r += TILE) {
                const int v = hV[r], iv = hI[r];
                if (v | iv) {
                    atomicAdd(&As.R1[r], v + iv);
                    if (v) atomicAdd(&As.R1Valid[r], v);
                    if (iv) atomicAdd(&As.R1Invalid[r], iv);
                }
            }
        }
        stamp(3);
        grid.sync();                                                    /* 1: ballots, block sums, deltas of this rank are final */

        /* ---- counts: CTA 0 / warp 0 talks to the peers */
        if (cta == 0 && warp == 0) {
            int mine = 0;
            for (int b = bLo + lane; b < bHi; b += 32) mine += __ldcg(&it.blockSum[b]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
            if (lane == 0) { plan->goalLocal = ~0ull; plan->goalGlobal = ~0ull; }
            bool ok = true;
            int cnt = 0;
            if (lane < world) {
                PeerMail* m = P.mail[lane] + rank;
                m->count = mine;
                /* release at system scope: cumulative over everything the grid barrier ordered before this thread —
                 * every CTA's delta increments of this iteration */
                st_release_sys_s32(&m->seq1, seq);
                const PeerMail* in = P.mail[rank] + lane;
                ok = peer_wait_ge(&in->seq1, seq);
                cnt = *(volatile const int*)&in->count;
            }
            const bool bad = __any_sync(0xffffffffu, !ok);
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t2 = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t2; }
            if (lane == rank) plan->base = incl - cnt;
            if (lane == 31) plan->total = incl;
            if (bad && lane == 0) plan->err = 1;
            __threadfence();
            __syncwarp();
            if (lane == 0) st_release_s32(&plan->ready1, seq);
        }
        cta_wait_flag(&plan->ready1, seq);                              /* peers' deltas are read below */
        if (*(volatile const int*)&plan->err) break;
        stamp(4);
        const int base = *(volatile const int*)&plan->base, total = *(volatile const int*)&plan->total;

        /* ---- pack: this rank's accepted rows into every replica's tree (updateG on every replica, KGMT.cu:555-591) */
        const int packSplit = insert_split(bHi - bLo, nCta);            /* few blocks: several CTAs share one (every n-th group of 32 rows) */
        for (int v = cta; v < (bHi - bLo) * packSplit; v += nCta) {
            const int blk = bLo + v / packSplit, slice = v - (v / packSplit) * packSplit;
            int m2 = 0;
            for (int b = bLo + tid; b < blk; b += TILE) m2 += __ldcg(&it.blockSum[b]);
            const int prefix = block_sum(m2, sRed);
            const int c = blk * BLK_CHUNKS + tid;
            const unsigned mask = (c < it.numChunks) ? __ldcg(&it.chunkMask[c]) : 0u;
            const int cnt = __popc(mask);
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t2 = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t2; }
            __syncthreads();
            if (lane == 31) sScan[warp] = incl;
            __syncthreads();
            int warpBase = 0;
#pragma unroll
            for (int w2 = 0; w2 < WARPS; ++w2) if (w2 < warp) warpBase += sScan[w2];
            const int W = __shfl_sync(0xffffffffu, incl, 31);
            const int c0 = blk * BLK_CHUNKS + warp * 32;
            const int q0g = base + prefix + warpBase;
            for (int q0 = 32 * slice; q0 < W; q0 += 32 * packSplit) {
                const int q = q0 + lane;
                int i = 0;
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    const int v2 = __shfl_sync(0xffffffffu, incl, i + step - 1);
                    if (v2 <= q) i += step;
                }
                i = min(i, 31);
                const unsigned m = __shfl_sync(0xffffffffu, mask, i);
                const int excl = __shfl_sync(0xffffffffu, incl - cnt, i);
                if (q < W) {
                    const int r = q - excl;
                    const int bit = nth_set_bit(m, r);
                    const int ci = c0 + i;
                    const int slot = ci * CHUNK + bit;
                    const float4 x = __ldcg(&it.stageState[ci * CHUNK + r]);
                    const float4 u = __ldcg(&it.stageCtrl[ci * CHUNK + r]);
                    const int row = q0g + q;
                    const int dst = it.treeSize + row;
                    const int parent = it.frontierStart + slot / it.children;
                    KGMT_CHECK_RANGE(210, dst, A.maxTree); KGMT_CHECK_RANGE(211, parent, it.treeSize);
                    for (int p = 0; p < world; ++p) {
                        P.treeState[p][dst] = x;
                        P.treeCtrl[p][dst] = u;
                        P.treeParent[p][dst] = parent;
                    }
                    if (in_goal(x.x, x.y, A.goalX, A.goalY, A.goalR))
                        atomicMin(&plan->goalLocal, ((unsigned long long)__float_as_uint(u.w) << 32) | (unsigned)row);
                }
            }
            __syncthreads();
        }

        /* ---- reduce: this rank's share of the delta index space (in 16-byte vectors: c1, c2 and so every section are
         * multiples of 4 ints or handled by the scalar tail), new values into every replica's maps.  The loads of one
         * vector from every rank are independent and issued together: over NVLink the pass is latency bound. */
        {
            const size_t nVec = deltaInts / 4;                           /* deltaInts = 4 (c1 + c2) */
            const size_t perV = nVec / world, extraV = nVec % world;
            const size_t loV = rank * perV + min((size_t)rank, extraV), hiV = loV + perV + ((size_t)rank < extraV ? 1 : 0);
            const unsigned stampNew = (unsigned)it.itr + 1u;
            const int* mine = P.mapSlab[rank];
            for (size_t vIdx = loV + (size_t)cta * TILE + tid; vIdx < hiV; vIdx += (size_t)nCta * TILE) {
                int4 acc = make_int4(0, 0, 0, 0);
#pragma unroll 4
                for (int p = 0; p < world; ++p) {
                    const int4 d = __ldcg(reinterpret_cast<const int4*>(P.delta[p]) + vIdx);
                    acc.x += d.x; acc.y += d.y; acc.z += d.z; acc.w += d.w;
                }
                if ((acc.x | acc.y | acc.z | acc.w) == 0) continue;
                const int sums[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int sum = sums[e];
                    if (sum == 0) continue;
                    const size_t idx = vIdx * 4 + e;
                    size_t at; int val; size_t at2 = (size_t)-1;
                    if (idx < 4 * c1) {
                        const size_t k = idx / c1, i = idx - k * c1;
                        if (k == 3) continue;
                        at = k * c1 + i; val = __ldcg(&mine[at]) + sum;
                        if (k == 1) at2 = 3 * c1 + i;                          /* R1Avail, KGMT.cu:400 */
                    } else {
                        const size_t j = idx - 4 * c1, k = j / c2, i = j - k * c2;
                        at = 7 * c1 + k * c2 + i;
                        if (k == 3) { if (__ldcg(&mine[at]) != 0) continue; val = (int)stampNew; }
                        else val = __ldcg(&mine[at]) + sum;
                    }
                    KGMT_CHECK_RANGE(212, at, 7 * c1 + 4 * c2);
                    for (int p = 0; p < world; ++p) {
                        P.mapSlab[p][at] = val;
                        if (at2 != (size_t)-1) P.mapSlab[p][at2] = 1;
                    }
                }
            }
        }
        stamp(5);
        grid.sync();                                                    /* 2: every pack / reduce store of this rank is issued */

        /* ---- goal exchange = the barrier across the GPUs */
        if (cta == 0 && warp == 0) {
            bool ok = true;
            unsigned long long g = ~0ull;
            if (lane < world) {
                PeerMail* m = P.mail[lane] + rank;
                m->goal = *(volatile unsigned long long*)&plan->goalLocal;
                /* system-scope release, cumulative over the grid barrier: every row and map value this rank stored into
                 * the peers' memory is visible to a peer that acquires this word */
                st_release_sys_s32(&m->seq2, seq);
                const PeerMail* in = P.mail[rank] + lane;
                ok = peer_wait_ge(&in->seq2, seq);
                g = *(volatile const unsigned long long*)&in->goal;
            }
            const bool bad = __any_sync(0xffffffffu, !ok);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { const unsigned long long t2 = __shfl_xor_sync(0xffffffffu, g, o); g = t2 < g ? t2 : g; }
            if (lane == 0) { plan->goalGlobal = g; if (bad) plan->err = 1; }
            __threadfence();
            __syncwarp();
            if (lane == 0) st_release_s32(&plan->ready2, seq);
        }
        cta_wait_flag(&plan->ready2, seq);                              /* rows and map values written by the peers are read from here on */
        if (*(volatile const int*)&plan->err) break;
        stamp(6);

        /* ---- finish: zero the delta slab, recount R1Cov, advance the planner scalars */
        {
            int4* z = reinterpret_cast<int4*>(P.delta[rank]);
            const size_t n4 = deltaInts / 4;                            /* c1 and c2 are multiples of 1: deltaInts = 4 (c1 + c2) */
            for (size_t i = (size_t)cta * TILE + tid; i < n4; i += (size_t)nCta * TILE) z[i] = make_int4(0, 0, 0, 0);
            for (int c = cta * WARPS + warp; c < A.c1; c += totalWarps) {
                int cntS = 0;
                for (int i = lane; i < nn; i += 32) cntS += (__ldcg(&A.R2Stamp[(size_t)c * nn + i]) != 0u);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) cntS += __shfl_xor_sync(0xffffffffu, cntS, o);
                if (lane == 0) A.R1Cov[c] = cntS;
            }
        }
        if (tid == 0) {
            const unsigned long long gb = *(volatile unsigned long long*)&plan->goalGlobal;   /* (cost bits << 32) | global row */
            const bool hadGoal = S.costToGoal != 0.0f;
            const int treeSize0 = S.treeSize;
            advance_state(A, S, total, gb);
            sGoalSlot = (!hadGoal && S.costToGoal != 0.0f) ? treeSize0 + (int)(unsigned)gb : -1;
            if (cta == 0 && sGoalSlot >= 0) st->goalIdx = sGoalSlot;
        }
        __syncthreads();
        /* recycle the ticket / block sums used two iterations from now; scalars to global memory */
        if (cta == (nCta > 1 ? nCta - 2 : 0)) {
            const int r = (it.itr + 2) % 3;
            for (int b = tid; b < A.blocksCap; b += TILE) A.blockSum[(size_t)r * A.blocksCap + b] = 0;
            if (tid == 0) A.ticket[r] = (unsigned)totalWarps;
        }
        if (cta == 0 && tid < COPIED_WORDS && tid != THRESHOLD_WORD)
            reinterpret_cast<int*>(st)[tid] = reinterpret_cast<const int*>(&S)[tid];
        __threadfence();
        if (cta == 0 && tid == 0 && A.iterLog && logRow < 255) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            A.iterLog[8 * logRow] = t;
            A.iterLog[8 * logRow + 1] = ((unsigned long long)(unsigned)it.M << 32) | (unsigned)total;
        }
        grid.sync();                                                    /* 3: delta zeroed, R1Cov recounted */
        stamp(7);
        if (cta == nCta - 1 && S.stop == STOP_RUNNING) {
            scores_block(A, sP, A.R1Score[S.itr & 1]);
            __threadfence();
            __syncthreads();
            if (tid == 0) st_release_s32(&st->scoreReady, S.itr);
        }
    }
    /* the insertion bookkeeping of the single-GPU loop is not used here: leave it consistent for a later kgmt_expand_* */
    if (cta == 0 && tid == 0) st->insertDone = S.blocksTotal;
}

/* -------------------------------------------- stages 2-4 alone (parity / sweeps) -------
 * candidate s expands parents[s / children] with stream (key0, slot0 + s); writes the
 * candidate records only. */
template <int COL>
__global__ void __launch_bounds__(TILE) propagate_only_kernel(const KArgs A, const float4* parents, long long M,
                                                              int children, uint32_t key0, uint32_t slot0) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t sBar;
    const int tid = threadIdx.x;
    const float4* sObs = nullptr; const int* sCellStart = nullptr; const float4* sItems = nullptr;
    if (COL == COL_GRID_SMEM || COL == COL_BRUTE_SMEM) {
        uint32_t bytes0 = 0, bytes1 = 0;
        if (COL == COL_GRID_SMEM) { bytes0 = (uint32_t)A.cellStartInts * 4u; bytes1 = (uint32_t)A.numItems * 16u; }
        else                      { bytes0 = (uint32_t)A.K * 16u; }
        if (tid == 0) { mbar_init(&sBar, 1); mbar_fence_init(); }
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(&sBar, bytes0 + bytes1);
            if (COL == COL_GRID_SMEM) {
                bulk_g2s_chunked(smem_raw, A.cellStart, bytes0, &sBar);
                bulk_g2s_chunked(smem_raw + bytes0, A.cellItems, bytes1, &sBar);
            } else {
                bulk_g2s_chunked(smem_raw, A.obstacles, bytes0, &sBar);
            }
        }
        mbar_wait(&sBar, 0);
        if (COL == COL_GRID_SMEM) {
            sCellStart = reinterpret_cast<const int*>(smem_raw);
            sItems = reinterpret_cast<const float4*>(smem_raw + bytes0);
        } else {
            sObs = reinterpret_cast<const float4*>(smem_raw);
        }
    }
    const DynParams dyn{A.W, A.H, A.L, A.numDisc};
    /* shared-window addresses of the staged CSR, pinned in registers (see stage_collision) */
    const uint32_t sStartAddr = pin_u32(sCellStart ? smem_u32(sCellStart) : 0u);
    const uint32_t sItemAddr = pin_u32(sItems ? smem_u32(sItems) : 0u);
    for (long long s = (long long)blockIdx.x * TILE + tid; s < M; s += (long long)gridDim.x * TILE) {
        float4 x = __ldg(&parents[s / children]);
        const Controls u = sample_controls(slot0 + (uint32_t)s, key0, A.car);
        bool valid;
        if (COL == COL_GRID_SMEM) {
            const CollideGridS col{sStartAddr, sItemAddr, A.cullC, A.cullInvX, A.cullInvY, A.cellStartInts, A.numItems};
            valid = propagate_edge(x, u, dyn, col);
        } else if (COL == COL_GRID_GLOBAL) {
            const CollideGrid col{A.cellStart, A.cellItems, A.cullC, A.cullInvX, A.cullInvY, A.cellStartInts, A.numItems};
            valid = propagate_edge(x, u, dyn, col);
        } else if (COL == COL_BRUTE_SMEM) {
            const CollideSmemAll col{sObs, A.K};
            valid = propagate_edge(x, u, dyn, col);
        } else {
            const CollideSmemAll col{A.obstacles, A.K};
            valid = propagate_edge(x, u, dyn, col);
        }
        const int r1 = region_r1(x.x, x.y, A.R1Size, A.N);
        const int r2 = region_r2(x.x, x.y, r1, A.R1Size, A.N, A.R2Size, A.n);
        A.candState[s] = x;
        A.candCtrl[s] = make_float4(u.a, u.steering, u.duration, u.u3);
        A.candParent[s] = (int)(s / children);
        A.candR1[s] = r1; A.candR2[s] = r2;
        A.candFlags[s] = (unsigned char)(valid ? FLAG_VALID : 0);
    }
}

/* ------------------------------- stages 5a / 5b alone on caller-supplied candidates -------
 * kgmt_stage_update_maps: the candidate records (state, controls + accept uniform, valid flag, parent) were uploaded
 * by the host; every warp runs chunk_finish — the SAME code the planner loop runs after propagate_edge — on 32 of them:
 * region indices, counters, accept on the iteration-start snapshot, ballots, staging, goal minimum.
 * (tail of propagateG, KGMT.cu:390-411) */
__global__ void __launch_bounds__(TILE) stage_update_kernel(const KArgs A) {
    __shared__ DevState S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < COPIED_WORDS) reinterpret_cast<int*>(&S)[tid] = __ldcg(reinterpret_cast<const int*>(A.st) + tid);
    __syncthreads();
    const IterView it = make_view(A, S);
    bool scoresOk = true;                      /* stream-ordered behind the kernels that wrote the scores */
    for (int c = (int)blockIdx.x * WARPS + warp; c < it.numChunks; c += (int)gridDim.x * WARPS) {
        const int s = c * CHUNK + lane;
        ChunkCand cc{};
        cc.live = s < it.M;
        cc.parent = -1;
        if (cc.live) {
            cc.x = A.candState[s];
            const float4 u = A.candCtrl[s];
            cc.u = Controls{u.x, u.y, u.z, u.w};
            cc.valid = (A.candFlags[s] & FLAG_VALID) != 0;
            cc.parent = A.candParent[s];
            cc.parentCost = (cc.parent >= 0 && cc.parent < S.treeSize) ? __ldcg(&A.treeCtrl[cc.parent]).w : 0.0f;
        }
        chunk_finish<true, false>(A, it, cc, c, lane, nullptr, nullptr, scoresOk);
    }
}

/* kgmt_stage_insert: phase B of the planner loop for the candidates staged by stage_update_kernel, by ONE CTA:
 * scan(GNew) + findInd + updateG (KGMT.cu:222-245, :540-593), treeSize / costToGoal / stop (:249-259), next scores. */
__global__ void __launch_bounds__(TILE) stage_insert_kernel(const KArgs A) {
    __shared__ int sRed[WARPS];
    __shared__ float sP[1024];
    __shared__ DevState S;
    __shared__ int sGoalSlot;
    const int tid = threadIdx.x;
    DevState* st = A.st;
    if (tid < COPIED_WORDS) reinterpret_cast<int*>(&S)[tid] = __ldcg(reinterpret_cast<const int*>(st) + tid);
    __syncthreads();
    if (S.stop != STOP_RUNNING) return;
    IterView it = make_view(A, S);
    it.parentOf = A.candParent;
    const int numBlocks = (it.numChunks + BLK_CHUNKS - 1) / BLK_CHUNKS;
    int mine = 0;
    for (int b = tid; b < numBlocks; b += TILE) mine += __ldcg(&it.blockSum[b]);
    const int accepted = block_sum(mine, sRed);
    if (tid == 0) {
        const unsigned long long gb = *(volatile unsigned long long*)&st->goalBest[it.itr & 1];
        const bool hadGoal = S.costToGoal != 0.0f;
        advance_state(A, S, accepted, gb);
        sGoalSlot = (!hadGoal && S.costToGoal != 0.0f) ? S.goalSlot : -1;
    }
    __syncthreads();
    it.goalSlot = sGoalSlot;
    for (int blk = 0; blk < numBlocks; ++blk) {
        int m2 = 0;
        for (int b = tid; b < blk; b += TILE) m2 += __ldcg(&it.blockSum[b]);
        const int base = block_sum(m2, sRed);
        insert_block(A, it, blk, base, sRed);
    }
    __syncthreads();
    if (S.stop == STOP_RUNNING) scores_block(A, sP, A.R1Score[S.itr & 1]);
    __syncthreads();
    for (int b = tid; b < numBlocks; b += TILE) it.blockSum[b] = 0;       /* leave the scan block sums clean */
    if (tid < COPIED_WORDS && tid != THRESHOLD_WORD) reinterpret_cast<int*>(st)[tid] = reinterpret_cast<const int*>(&S)[tid];
    __threadfence();
    __syncthreads();
    if (tid == 0) { st->scoreReady = S.itr; st->insertDone = S.blocksTotal; }
}

/* ------------------------------------------------- cull grid built on the device --------
 * CSR of obstacle AABBs per cell of a C x C grid over the workspace (CollideGrid, kgmt_device.cuh).  An obstacle is filed
 * under every cell of [cell(minx), cell(maxx)] x [cell(miny), cell(maxy)] with the SAME monotone cell() the collision
 * walk applies to a step bbox, so culling cannot change a flag.  One thread per CELL walks the obstacle array (every
 * thread reads the same obstacle: one broadcast load), first to count, then — after a scan — to fill in obstacle
 * order: the structure is deterministic, no atomics, no host round trip except the item count (4 bytes). */
__device__ __forceinline__ int cull_cell_of(float v, float inv, int C) {
    return min(max(__float2int_rd(__fmul_rn(v, inv)), 0), C - 1);
}
__global__ void cull_pad_kernel(float4* obs, int K, int padded) {
    const int k = K + blockIdx.x * blockDim.x + threadIdx.x;     /* boxes nothing can overlap (min = +inf, max = -inf) */
    if (k < padded) obs[k] = make_float4(__int_as_float(0x7f800000), __int_as_float(0x7f800000), __int_as_float(0xff800000), __int_as_float(0xff800000));
}
/* one WARP per cell: lanes stride the obstacle array, the counts are summed by shuffles */
__global__ void cull_count_kernel(const float4* obs, int K, int C, float invX, float invY, int* start /* [C*C+1], zeroed */) {
    const int cell = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (cell >= C * C) return;                                   /* whole warps leave together */
    const int cy = cell / C, cx = cell - cy * C;
    int n = 0;
    for (int k = lane; k < K; k += 32) {
        const float4 o = __ldg(&obs[k]);
        n += (cull_cell_of(o.x, invX, C) <= cx) & (cx <= cull_cell_of(o.z, invX, C)) &
             (cull_cell_of(o.y, invY, C) <= cy) & (cy <= cull_cell_of(o.w, invY, C));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if (lane == 0) start[cell + 1] = n;
}
/* one CTA: start[i] = sum of the counts before cell i (in place), the padding words = the total; *total too */
__global__ void __launch_bounds__(1024) cull_scan_kernel(int* start, int cells, int startInts, int* total) {
    __shared__ int sWarp[32];
    __shared__ int sCarry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { sCarry = 0; start[0] = 0; }
    __syncthreads();
    for (int base = 1; base <= cells; base += 1024) {
        const int i = base + tid;
        const int v = (i <= cells) ? start[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) sWarp[warp] = incl;
        __syncthreads();
        int wb = 0;
        for (int w2 = 0; w2 < warp; ++w2) wb += sWarp[w2];
        const int carry = sCarry;
        if (i <= cells) start[i] = carry + wb + incl;
        __syncthreads();
        if (tid == 1023) sCarry = carry + wb + incl;
        __syncthreads();
    }
    const int t = sCarry;
    for (int i = cells + 1 + tid; i < startInts; i += 1024) start[i] = t;
    if (tid == 0) *total = t;
}
/* one WARP per cell, 32 obstacles per trip: the ballot of the lanes whose obstacle touches the cell gives every such
 * obstacle its place, in obstacle order — the same CSR the one-thread-per-cell fill produced, in 1/32 of the trips */
__global__ void cull_fill_kernel(const float4* obs, int K, int C, float invX, float invY, const int* start, float4* items, int total) {
    const int gid = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const int cell = gid >> 5, lane = threadIdx.x & 31;
    if (gid < 3)        /* the cell walk reads up to four entries per trip: three boxes nothing overlaps close the array */
        items[total + gid] = make_float4(__int_as_float(0x7f800000), __int_as_float(0x7f800000), __int_as_float(0xff800000), __int_as_float(0xff800000));
    if (cell >= C * C) return;
    const int cy = cell / C, cx = cell - cy * C;
    int at = start[cell];
    for (int k0 = 0; k0 < K; k0 += 32) {
        const int k = k0 + lane;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        bool in = false;
        if (k < K) {
            o = __ldg(&obs[k]);
            in = (cull_cell_of(o.x, invX, C) <= cx) & (cx <= cull_cell_of(o.z, invX, C)) &
                 (cull_cell_of(o.y, invY, C) <= cy) & (cy <= cull_cell_of(o.w, invY, C));
        }
        const unsigned bal = __ballot_sync(0xffffffffu, in);
        if (in) items[at + __popc(bal & ((1u << lane) - 1u))] = o;
        at += __popc(bal);
    }
}

/* ------------------------------------------------------------------ setup kernels ------ */
/* root insertion, KGMT.cu:85-114, then the first iteration's shape and scores (one CTA) */
__device__ void begin_block(const KArgs& A, float4 rootState, float4 rootCtrl, float* sP, int forceChildren) {
    DevState* st = A.st;
    if (threadIdx.x == 0) {
        A.treeState[0] = rootState;                                                /* :85 */
        A.treeCtrl[0] = make_float4(rootCtrl.x, rootCtrl.y, rootCtrl.z, 0.0f);
        A.treeParent[0] = -1;
        const int r1 = region_r1(rootState.x, rootState.y, A.R1Size, A.N);         /* :88 */
        const int r2 = region_r2(rootState.x, rootState.y, r1, A.R1Size, A.N, A.R2Size, A.n);   /* :89 */
        if (r1 >= 0) { A.R1[r1] = 1; A.R1Avail[r1] = 1; A.R1Valid[r1] = 1; }       /* :94,95,97 */
        if (r2 >= 0 && A.R2Stamp[r2] == 0u) { A.R2Stamp[r2] = 1u; A.R1Cov[r1] += 1; }  /* :96 */
        DevState z{};
        z.treeSize = 1; z.frontierStart = 0; z.frontierCount = 1; z.itr = 1;
        z.goalIdx = -1; z.goalSlot = -1; z.costToGoal = 0.0f; z.goalBest[0] = z.goalBest[1] = ~0ull;
        z.forceChildren = forceChildren;
        A.ticket[0] = A.ticket[1] = A.ticket[2] = (unsigned)A.totalWarps;
        int stop = STOP_RUNNING;
        if (A.numIterations <= 0) stop = STOP_ITER_LIMIT;
        else if (1 >= A.maxTree) stop = STOP_TREE_FULL;
        z.stop = stop;
        int mode = 0, children = 1, M = 0;
        if (stop == STOP_RUNNING) expansion_shape(1, 1, A.maxTree, A.maxCand, z.forceChildren, mode, children, M);
        z.mode = mode; z.children = children; z.M = M; z.numChunks = (M + CHUNK - 1) / CHUNK;
        *st = z;
    }
    __syncthreads();
    scores_block(A, sP, A.R1Score[1]);              /* iteration 1 reads buffer 1 & 1 */
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) st_release_s32(&st->scoreReady, 1);
}

__global__ void __launch_bounds__(TILE) begin_kernel(const KArgs A, float4 rootState, float4 rootCtrl, int forceChildren) {
    __shared__ float sP[1024];
    begin_block(A, rootState, rootCtrl, sP, forceChildren);
}

/* kgmt_seed_frontier: `count` nodes already copied into tree[0,count); all are frontier */
__global__ void seed_mark_kernel(const KArgs A, int count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float4 s = A.treeState[i];
    A.treeParent[i] = -1;
    const int r1 = region_r1(s.x, s.y, A.R1Size, A.N);
    const int r2 = region_r2(s.x, s.y, r1, A.R1Size, A.N, A.R2Size, A.n);
    if (r1 >= 0) { A.R1[r1] = 1; A.R1Avail[r1] = 1; A.R1Valid[r1] = 1; }
    if (r2 >= 0) A.R2Stamp[r2] = 1u;
}
__global__ void __launch_bounds__(TILE) seed_finish_kernel(const KArgs A, int count) {
    __shared__ float sP[1024];
    DevState* st = A.st;
    if (threadIdx.x == 0) {
        st->treeSize = count; st->frontierStart = 0; st->frontierCount = count; st->itr = 1;
        st->goalIdx = -1; st->goalSlot = -1; st->costToGoal = 0.0f; st->goalBest[0] = st->goalBest[1] = ~0ull;
        st->peerSolved[0] = st->peerSolved[1] = 0; st->stepsDone = 0; st->pairsTested = 0;
        st->expansions = 0; st->iterationsDone = 0; st->blocksTotal = 0; st->insertDone = 0;
        st->lastMode = st->lastChildren = st->lastFrontier = st->lastM = st->lastAccepted = st->lastItr = 0;
        A.ticket[0] = A.ticket[1] = A.ticket[2] = (unsigned)A.totalWarps;
        int stop = STOP_RUNNING;
        if (A.numIterations <= 0) stop = STOP_ITER_LIMIT;
        else if (count >= A.maxTree) stop = STOP_TREE_FULL;
        st->stop = stop;
        int mode = 0, children = 1, M = 0;
        if (stop == STOP_RUNNING) expansion_shape(count, count, A.maxTree, A.maxCand, st->forceChildren, mode, children, M);
        st->mode = mode; st->children = children; st->M = M; st->numChunks = (M + CHUNK - 1) / CHUNK;
    }
    scores_block(A, sP, A.R1Score[1]);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) st_release_s32(&st->scoreReady, 1);
}

/* R1Cov[c] = number of available R2 cells of R1 cell c (after import / seeding) */
__global__ void recount_cov_kernel(const KArgs A) {
    const int c = blockIdx.x;
    const int nn = A.n * A.n;
    __shared__ int sCnt;
    if (threadIdx.x == 0) sCnt = 0;
    __syncthreads();
    int cnt = 0;
    for (int i = threadIdx.x; i < nn; i += blockDim.x) cnt += (A.R2Stamp[(size_t)c * nn + i] != 0u);
    if (cnt) atomicAdd(&sCnt, cnt);
    __syncthreads();
    if (threadIdx.x == 0) A.R1Cov[c] = sCnt;
}

__global__ void __launch_bounds__(TILE) scores_kernel(const KArgs A) {
    __shared__ float sP[1024];
    scores_block(A, sP, A.R1Score[A.st->itr & 1]);
}

/* solution back-trace: walk the parent links from `node` to the root, write the chain root
 * first as AoS-7 rows (the reference's samples layout) and its length.  Lanes walk in lock step from staggered
 * starts: lane l first skips l links, then every lane advances 32 links per trip. */
__global__ void trace_path_kernel(const KArgs A, int node, int treeSize, int* outLen, float* rows7, int maxRows) {
    const int lane = threadIdx.x;
    int len = 0;
    if (lane == 0) {
        for (int v = node; v >= 0 && len <= treeSize; v = __ldcg(&A.treeParent[v])) ++len;
        *outLen = len;
    }
    len = __shfl_sync(0xffffffffu, len, 0);
    /* row `at` (root = 0) is the node reached after len-1-at links; one link walk per lane, rows strided by 32 */
    int v = node, at = len - 1;
    for (int i = 0; i < lane && v >= 0; ++i) { v = __ldcg(&A.treeParent[v]); --at; }
    while (v >= 0 && at >= 0) {
        if (at < maxRows) {
            const float4 x = __ldcg(&A.treeState[v]), u = __ldcg(&A.treeCtrl[v]);
            float* o = rows7 + (size_t)at * 7;
            o[0] = x.x; o[1] = x.y; o[2] = x.z; o[3] = x.w; o[4] = u.x; o[5] = u.y; o[6] = u.z;
        }
        for (int i = 0; i < 32 && v >= 0; ++i) { v = __ldcg(&A.treeParent[v]); --at; }
    }
}

/* the same for the goal node of the plan that has just run, node and tree size read from the planner's scalars on the
 * device (kgmt_plan launches it behind the planner kernel, so the path travels with the scalars in one copy) */
__global__ void trace_goal_kernel(const KArgs A, int* outLen, float* rows7, int maxRows) {
    const int lane = threadIdx.x;
    const DevState* st = A.st;
    const int node = (*(volatile const int*)&st->stop == STOP_SOLVED) ? *(volatile const int*)&st->goalIdx : -1;
    const int treeSize = *(volatile const int*)&st->treeSize;
    if (node < 0 || node >= treeSize) { if (lane == 0) *outLen = -1; return; }
    int len = 0;
    if (lane == 0) {
        for (int v = node; v >= 0 && len <= treeSize; v = __ldcg(&A.treeParent[v])) ++len;
        *outLen = len;
    }
    len = __shfl_sync(0xffffffffu, len, 0);
    int v = node, at = len - 1;
    for (int i = 0; i < lane && v >= 0; ++i) { v = __ldcg(&A.treeParent[v]); --at; }
    while (v >= 0 && at >= 0) {
        if (at < maxRows) {
            const float4 x = __ldcg(&A.treeState[v]), u = __ldcg(&A.treeCtrl[v]);
            float* o = rows7 + (size_t)at * 7;
            o[0] = x.x; o[1] = x.y; o[2] = x.z; o[3] = x.w; o[4] = u.x; o[5] = u.y; o[6] = u.z;
        }
        for (int i = 0; i < 32 && v >= 0; ++i) { v = __ldcg(&A.treeParent[v]); --at; }
    }
}

/* views in the reference's element layout (export) */
__global__ void gather_samples_kernel(const float4* st, const float4* ct, float* out7, int count, int costIsU3) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float4 s = st[i], c = ct[i];
    float* o = out7 + (size_t)i * 7;
    o[0] = s.x; o[1] = s.y; o[2] = s.z; o[3] = s.w; o[4] = c.x; o[5] = c.y; o[6] = c.z;
}
__global__ void gather_w_kernel(const float4* ct, float* out, int count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = ct[i].w;
}
__global__ void stamp_to_avail_kernel(const unsigned* stamp, int* out, int count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = stamp[i] != 0u;
}
__global__ void avail_to_stamp_kernel(const int* in, unsigned* stamp, int count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) stamp[i] = in[i] != 0 ? 1u : 0u;
}
__global__ void flags_bit_kernel(const unsigned char* flags, unsigned char* out, int count, int bit) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = (flags[i] & bit) ? 1 : 0;
}
__global__ void frontier_flags_kernel(unsigned char* out, int count, int start, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = (i >= start && i < start + n) ? 1 : 0;
}
__global__ void scatter_samples_kernel(const float* in7, float4* st, float4* ct, int count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float* r = in7 + (size_t)i * 7;
    st[i] = make_float4(r[0], r[1], r[2], r[3]);
    ct[i] = make_float4(r[4], r[5], r[6], 0.0f);
}
__global__ void fill_int_kernel(int* p, int v, size_t count) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) p[i] = v;
}
__global__ void fill_float_kernel(float* p, float v, size_t count) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) p[i] = v;
}

}  // namespace kgmt
