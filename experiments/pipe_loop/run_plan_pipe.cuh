/* experiments/pipe_loop/run_plan_pipe.cuh — NOT part of the product, NOT compiled.
 *
 * The barrier-free ("pipelined") planner loop of round 1, archived as it was when it was removed from
 * cudasbmp_b200/csrc/kgmt_kernels.cuh (last integrated at commit a20a712; see README.md next to this file).
 * It gave bit-identical trees but was slower than the grid-barrier loop (C2 2.1 ms vs 1.38 ms, C1 0.64 vs 0.27 ms),
 * so the product ships only run_plan.  Kept for the record of what was tried; it references KArgs / DevState fields
 * (blockDone, blockPrefix, blockInserted, pipe, pipeMode, pipeEpoch) that no longer exist in the product headers.
 */
/* =========================================================== the pipelined planner loop ====
 * run_plan above separates the iterations with a grid barrier: every warp idles from the moment the chunk tickets run
 * out until the slowest chunk, the ordered insertion and the scores are done (~20 % of a 900 k-candidate iteration).
 * run_plan_pipe has NO grid barrier.  Warps never synchronise with their CTA inside the loop:
 *
 *   chunk done      -> fence, blockDone[b]++; the warp that completes scan block b is its CLOSER
 *   closer(b)       -> extends the prefix chain (one warp at a time, 32 blocks per step), waits until the rows before
 *                      block b are known, inserts the block's accepted rows itself (updateG), then advances the
 *                      in-order watermark rowsReady = frontier rows of the NEXT iteration already in the tree
 *   tickets run out -> the warp signs off (last warp of a CTA flushes the R1 histograms; the last CTA FINALIZES the
 *                      iteration: accepted total from the chain, planner scalars, next scores, publication) and takes
 *                      ONE chunk ticket of the next iteration: as soon as that chunk's parent row is in the tree it
 *                      runs stages 2-4 for it speculatively (policy 32 children per node assumed; nothing global is
 *                      written) and keeps the result in registers until the iteration is published, then validates
 *                      the assumption and completes stage 5a.  The idle tail is filled with useful work, and the
 *                      insertion and the scores overlap the propagation.
 * Results are identical to run_plan (same candidate slots, same snapshot semantics, same insertion order). */
__device__ __forceinline__ void st_relaxed_s32(int* p, int v) {
    asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

#ifndef KGMT_TRACE_ITR
#define KGMT_TRACE_ITR 10
#endif
#ifndef PIPE_SLEEP_CAP
#define PIPE_SLEEP_CAP 2048
#endif
/* A flag read that steers the whole warp must be ONE read: 32 lanes loading the same word in one instruction can
 * observe different values while another SM is writing it, and a warp that splits on such a value meets its own
 * __shfl_sync / __ballot_sync calls from different places.  Lane 0 reads, the warp takes its answer. */
__device__ __forceinline__ int ld_acquire_warp(const int* p, int lane) {
    int v = 0;
    if (lane == 0) v = ld_acquire_s32(p);
    return __shfl_sync(0xffffffffu, v, 0);
}
__device__ __forceinline__ int ld_relaxed_warp(const int* p, int lane) {
    int v = 0;
    if (lane == 0) v = ld_relaxed_s32(p);
    return __shfl_sync(0xffffffffu, v, 0);
}

struct PipeView { int itr, stop, treeSize, frontierStart, children, M, numChunks, mode, iterationsDone; };

__device__ __forceinline__ PipeView pipe_load_view(const DevState* st, int lane) {
    /* lanes 0..8 fetch one word each, everybody gets all nine */
    const int* w = reinterpret_cast<const int*>(st);
    const int offs[9] = {offsetof(DevState, itr) / 4, offsetof(DevState, stop) / 4, offsetof(DevState, treeSize) / 4,
                         offsetof(DevState, frontierStart) / 4, offsetof(DevState, children) / 4, offsetof(DevState, M) / 4,
                         offsetof(DevState, numChunks) / 4, offsetof(DevState, mode) / 4, offsetof(DevState, iterationsDone) / 4};
    int mine = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) if (lane == k) mine = __ldcg(w + offs[k]);
    PipeView v;
    v.itr = __shfl_sync(0xffffffffu, mine, 0); v.stop = __shfl_sync(0xffffffffu, mine, 1);
    v.treeSize = __shfl_sync(0xffffffffu, mine, 2); v.frontierStart = __shfl_sync(0xffffffffu, mine, 3);
    v.children = __shfl_sync(0xffffffffu, mine, 4); v.M = __shfl_sync(0xffffffffu, mine, 5);
    v.numChunks = __shfl_sync(0xffffffffu, mine, 6); v.mode = __shfl_sync(0xffffffffu, mine, 7);
    v.iterationsDone = __shfl_sync(0xffffffffu, mine, 8);
    return v;
}

__device__ __forceinline__ IterView pipe_iter_view(const KArgs& A, const PipeView& V) {
    IterView it;
    it.itr = V.itr; it.treeSize = V.treeSize; it.frontierStart = V.frontierStart; it.children = V.children;
    it.M = V.M; it.numChunks = V.numChunks; it.key0 = A.seed + (uint32_t)V.itr;
    it.score = A.R1Score[V.itr & 1];
    it.chunkMask = A.chunkMask + (size_t)(V.itr & 1) * A.chunksCap;
    it.blockSum = A.blockSum + (size_t)(V.itr % 3) * A.blocksCap;
    it.stageState = A.stageState + (size_t)(V.itr & 1) * A.maxCand;
    it.stageCtrl = A.stageCtrl + (size_t)(V.itr & 1) * A.maxCand;
    it.goalSlot = -1;
    return it;
}

__device__ __forceinline__ int chunks_in_block(int b, int numChunks) { return min(BLK_CHUNKS, numChunks - b * BLK_CHUNKS); }
__device__ __forceinline__ int units_in_block(int b, int numChunks) { return (chunks_in_block(b, numChunks) + 31) / 32; }

/* extend the prefix chain of iteration `it` as far as complete blocks allow (whole warp; returns when nothing is left
 * to do or another warp holds the chain) */
__device__ __forceinline__ void pipe_chain(const KArgs& A, const IterView& it, PipeIter* pi, int numBlocks, int lane) {
    const int* done = A.blockDone + (size_t)(it.itr % 3) * A.blocksCap;
    int* prefix = A.blockPrefix + (size_t)(it.itr % 3) * A.blocksCap;
    for (;;) {
        int got = 0;
        if (lane == 0) got = (atomicCAS(&pi->chainLock, 0, 1) == 0);
        got = __shfl_sync(0xffffffffu, got, 0);
        if (!got) return;
        fence_acq_rel();
        for (;;) {
            const int W = ld_relaxed_warp(&pi->chainW, lane), sum = ld_relaxed_warp(&pi->chainSum, lane);
            const int b = W + lane;
            const bool complete = b < numBlocks && ld_relaxed_s32(&done[b]) == chunks_in_block(b, it.numChunks);
            const unsigned bal = __ballot_sync(0xffffffffu, complete);
            const int k = __ffs((int)~bal) - 1;                  /* leading run of complete blocks (32 if all) */
            const int run = (bal == 0xffffffffu) ? 32 : k;
            if (run == 0) break;
            fence_acq_rel();                                     /* the block sums of complete blocks are final */
            const int v = (lane < run) ? __ldcg(&it.blockSum[b]) : 0;
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            if (lane < run) prefix[b] = sum + incl - v;
            const int total = sum + __shfl_sync(0xffffffffu, incl, run - 1);
            __syncwarp();
            if (lane == 0) {
                __threadfence();
                st_relaxed_s32(&pi->chainSum, total);
                st_release_s32(&pi->chainW, W + run);
            }
            __syncwarp();
        }
        if (lane == 0) { __threadfence(); atomicExch(&pi->chainLock, 0); }
        __syncwarp();
        /* a block may have completed between the last look and the unlock: look again */
        int again = 0;
        if (lane == 0) {
            const int W = ld_relaxed_s32(&pi->chainW);
            again = (W < numBlocks && ld_relaxed_s32(&done[W]) == chunks_in_block(W, it.numChunks));
        }
        if (!__shfl_sync(0xffffffffu, again, 0)) return;
    }
}

/* advance the in-order insertion watermark over the blocks whose rows have landed (whole warp; never waits) */
__device__ __forceinline__ void pipe_rows(const KArgs& A, const IterView& it, PipeIter* pi, int numBlocks, int lane) {
    const int* ins = A.blockInserted + (size_t)(it.itr % 3) * A.blocksCap;
    const int* prefix = A.blockPrefix + (size_t)(it.itr % 3) * A.blocksCap;
    for (;;) {
        int got = 0;
        if (lane == 0) got = (atomicCAS(&pi->rowsLock, 0, 1) == 0);
        got = __shfl_sync(0xffffffffu, got, 0);
        if (!got) return;
        fence_acq_rel();
        for (;;) {
            const int W = ld_relaxed_warp(&pi->rowsBlocks, lane);
            const int b = W + lane;
            const bool in = b < numBlocks && ld_relaxed_s32(&ins[b]) == units_in_block(b, it.numChunks);
            const unsigned bal = __ballot_sync(0xffffffffu, in);
            const int run = (bal == 0xffffffffu) ? 32 : __ffs((int)~bal) - 1;
            if (run == 0) break;
            if (lane == run - 1) {
                fence_acq_rel();                                 /* the rows of every block up to b are visible */
                st_relaxed_s32(&pi->rowsReady, __ldcg(&prefix[b]) + __ldcg(&it.blockSum[b]));
                st_release_s32(&pi->rowsBlocks, W + run);
            }
            __syncwarp();
        }
        if (lane == 0) { __threadfence(); atomicExch(&pi->rowsLock, 0); }
        __syncwarp();
        int again = 0;
        if (lane == 0) {
            const int W = ld_relaxed_s32(&pi->rowsBlocks);
            again = (W < numBlocks && ld_relaxed_s32(&ins[W]) == units_in_block(W, it.numChunks));
        }
        if (!__shfl_sync(0xffffffffu, again, 0)) return;
    }
}

constexpr int SUBS = BLK_CHUNKS / 32;     /* insertion units per scan block */

/* updateG (KGMT.cu:555-591) for ONE insertion unit = 32 consecutive chunks (unit u of scan block blk) by one warp;
 * base = accepted rows in earlier blocks.  Same row order as insert_block. */
__device__ __forceinline__ void insert_unit_warp(const KArgs& A, const IterView& it, int blk, int u, int base, int lane) {
    const int cBlk = blk * BLK_CHUNKS;
    /* rows of the earlier units of this block */
    int before = 0;
    for (int k = 0; k < u; ++k) {
        const int c = cBlk + k * 32 + lane;
        before += (c < it.numChunks) ? __popc(__ldcg(&it.chunkMask[c])) : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
    const int c0 = cBlk + u * 32;
    const int c = c0 + lane;
    const unsigned mask = (c < it.numChunks) ? __ldcg(&it.chunkMask[c]) : 0u;
    const int cnt = __popc(mask);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    const int W = __shfl_sync(0xffffffffu, incl, 31);
    const int dst0 = it.treeSize + base + before;
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
            const float4 x = __ldcg(&it.stageState[ci * CHUNK + r]);
            const float4 uu = __ldcg(&it.stageCtrl[ci * CHUNK + r]);
            const int dst = dst0 + q;
#ifdef KGMT_PIPE_CHECK
            if (dst < 0 || dst >= A.maxTree || ci * CHUNK + r >= A.maxCand) {
                printf("insert OOB: itr %d blk %d u %d base %d before %d q %d W %d dst %d treeSize %d\n", it.itr, blk, u, base, before, q, W, dst, it.treeSize);
                __trap();
            }
#endif
            A.treeState[dst] = x;
            A.treeCtrl[dst] = uu;
            A.treeParent[dst] = it.frontierStart + slot / it.children;
        }
    }
}

/* The R1 scores of iteration `itr` (updateR1, KGMT.cu:487-538) from the final maps of the previous one, computed
 * COOPERATIVELY by the warps that reach the iteration first: groups of 32 cells by ticket (one cell per lane: the raw
 * score goes straight into the score buffer), the warp that finishes the last group sums them in scores_block's order
 * (p[t] = sum_k score[t + 1024k], stride-halving tree; tree levels that would only add +0 are skipped: x + 0 == x),
 * normalises and publishes scoreReady.  A single warp doing all of it on a busy SM took ~20 us per iteration. */
__device__ void pipe_scores_help(const KArgs& A, int itr, PipeIter* pn, float* p /* smem [1024] */, int lane) {
    DevState* st = A.st;
    const int c1 = A.c1, groups = (c1 + 31) / 32;
    float* out = A.R1Score[itr & 1];
    const float nn = (float)(A.n * A.n);
    for (;;) {
        if (ld_relaxed_warp(&st->scoreReady, lane) >= itr) return;
        int g = 0;
        if (lane == 0) g = atomicAdd(&pn->scoreTicket, 1);
        g = __shfl_sync(0xffffffffu, g, 0);
        if (g >= groups) return;                       /* every group is taken; stage 5a waits on scoreReady */
        const int c = g * 32 + lane;
        if (c < c1) {
            float score = 0.0f;
            if (__ldcg(&A.R1Avail[c]) != 0) {
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
            __stcg(&out[c], score);                    /* raw; normalised by the last finisher */
        }
        __syncwarp();
        int last = 0;
        if (lane == 0) { __threadfence(); last = (atomicAdd(&pn->scoreDone, 1) + 1 == groups); }
        if (!__shfl_sync(0xffffffffu, last, 0)) continue;
        fence_acq_rel();
        const int tEnd = min(1024, (c1 + 31) & ~31);
        int availLocal = 0;
#pragma unroll 4
        for (int t = lane; t < tEnd; t += 32) {
            float acc = 0.0f;
            for (int cc = t; cc < c1; cc += 1024) {
                acc = __fadd_rn(acc, __ldcg(&out[cc]));
                availLocal += (__ldcg(&A.R1Avail[cc]) != 0);
            }
            p[t] = acc;
        }
        for (int t = tEnd + lane; t < min(1024, 2 * tEnd); t += 32) p[t] = 0.0f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) availLocal += __shfl_xor_sync(0xffffffffu, availLocal, o);
        int top = 512;
        while (top >= tEnd && top > 1) top >>= 1;
        for (int stride = top; stride >= 1; stride >>= 1) {
            __syncwarp();
            for (int t = lane; t < stride; t += 32) p[t] = __fadd_rn(p[t], p[t + stride]);
        }
        __syncwarp();
        const float total = p[0];
        if (lane == 0) st->R1Threshold = availLocal ? __fdiv_rn(total, (float)availLocal) : 0.0f;
#pragma unroll 4
        for (int cc = lane; cc < c1; cc += 32)
            __stcg(&out[cc], (__ldcg(&A.R1Avail[cc]) == 0) ? 1.0f : __fdiv_rn(__ldcg(&out[cc]), total));
        __syncwarp();
        if (lane == 0) { __threadfence(); st_release_s32(&st->scoreReady, itr); }
        __syncwarp();
        return;
    }
}

/* end of iteration V.itr, by the warp that saw the last CTA sign off: KGMT.cu:249-259 + next scores + publication */
__device__ void pipe_finalize(const KArgs& A, const PipeView& V, const IterView& it, PipeIter* pi, PipeIter* pprev,
                              bool firstOfLaunch, bool lastOfLaunch, DevState* sS, int lane) {
    DevState* st = A.st;
    const int numBlocks = (V.numChunks + BLK_CHUNKS - 1) / BLK_CHUNKS;
    auto stamp = [&](int col) {             /* diagnostics: finalizer timeline (kgmt_iteration_log) */
        if (A.iterLog && lane == 0 && V.iterationsDone < 255) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            A.iterLog[8 * V.iterationsDone + col] = t;
        }
    };
    stamp(2);
    fence_acq_rel();
    /* every block is complete: finish the chain -> accepted rows of the iteration */
    while (ld_acquire_warp(&pi->chainW, lane) < numBlocks) pipe_chain(A, it, pi, numBlocks, lane);
    const int accepted = ld_relaxed_warp(&pi->chainSum, lane);
    stamp(3);
    /* the buffers of iteration itr-1 are recycled for itr+2 below: its insertions must have landed */
    {
        const int nbPrev = ld_relaxed_warp(&pprev->numBlocks, lane);
        if (nbPrev > 0 && ld_acquire_warp(&pprev->rowsBlocks, lane) < nbPrev) {
            IterView pv = it;                   /* only the fields pipe_rows reads */
            pv.itr = V.itr - 1; pv.blockSum = A.blockSum + (size_t)((V.itr + 2) % 3) * A.blocksCap;
            unsigned ns = 32;
            for (;;) {
                pipe_rows(A, pv, pprev, nbPrev, lane);
                if (ld_acquire_warp(&pprev->rowsBlocks, lane) >= nbPrev) break;
                __nanosleep(ns); if (ns < 512) ns <<= 1;
            }
        }
        const int r2 = (V.itr + 2) % 3;
        for (int b = lane; b < nbPrev; b += 32) {
            A.blockSum[(size_t)r2 * A.blocksCap + b] = 0;
            A.blockDone[(size_t)r2 * A.blocksCap + b] = 0;
            A.blockInserted[(size_t)r2 * A.blocksCap + b] = 0;
        }
        if (lane == 0) { A.ticket[r2] = 0u; *pprev = PipeIter{}; }
    }
    stamp(4);
    if (lane < COPIED_WORDS) reinterpret_cast<int*>(sS)[lane] = __ldcg(reinterpret_cast<const int*>(st) + lane);
    __syncwarp();
    int goalSlot = -1;
    if (lane == 0) {
        const unsigned long long gb = *(volatile unsigned long long*)&st->goalBest;
        const bool hadGoal = sS->costToGoal != 0.0f;
        advance_state(A, *sS, accepted, gb);
        if (!hadGoal && sS->costToGoal != 0.0f) goalSlot = sS->goalSlot;
        st_relaxed_s32(&pi->numBlocks, numBlocks);
        if (A.iterLog && sS->iterationsDone - 1 < 255) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            A.iterLog[8 * (sS->iterationsDone - 1)] = t;
            A.iterLog[8 * (sS->iterationsDone - 1) + 1] = ((unsigned long long)(unsigned)V.M << 32) | (unsigned)accepted;
        }
    }
    goalSlot = __shfl_sync(0xffffffffu, goalSlot, 0);
    if (goalSlot >= 0) {
        /* tree index of the goal node = rows before its block + rows of earlier chunks of the block + its rank */
        const int c = goalSlot / CHUNK, bit = goalSlot % CHUNK, blk = c / BLK_CHUNKS;
        int before = 0;
        for (int cc = blk * BLK_CHUNKS + lane; cc < c; cc += 32) before += __popc(__ldcg(&it.chunkMask[cc]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
        if (lane == 0) {
            const int* prefix = A.blockPrefix + (size_t)(V.itr % 3) * A.blocksCap;
            const unsigned m = __ldcg(&it.chunkMask[c]);
            st->goalIdx = V.treeSize + __ldcg(&prefix[blk]) + before + __popc(m & ((1u << bit) - 1u));
        }
    }
    __syncwarp();
    const bool running = sS->stop == STOP_RUNNING;
    if (!running || lastOfLaunch) {
        /* the launch ends here: leave the tree complete and the block bookkeeping as run_plan / the sharded path expect */
        unsigned ns = 32;
        for (;;) {
            pipe_rows(A, it, pi, numBlocks, lane);
            if (ld_acquire_warp(&pi->rowsBlocks, lane) >= numBlocks) break;
            __nanosleep(ns); if (ns < 512) ns <<= 1;
        }
        if (lane == 0) st->insertDone = sS->blocksTotal;
    }
    __syncwarp();
    /* publication, step 1: the shape of the next iteration (stages 2-4 of its chunks can start) */
    if (lane < COPIED_WORDS && lane != THRESHOLD_WORD) reinterpret_cast<int*>(st)[lane] = reinterpret_cast<const int*>(sS)[lane];
    __syncwarp();
    if (lane == 0) { __threadfence(); st_release_s32(&st->pipeEpoch, sS->iterationsDone); }
    stamp(5);
    /* step 2, its scores (stage 5a waits for them), is shared by the warps that arrive first: pipe_scores_help */
    __syncwarp();
}

template <int COL, bool RECORD>
__device__ void run_plan_pipe(const KArgs& A, int maxIters, const ColSet& cs) {
    __shared__ float sP[1024];
    __shared__ DevState sS;                 /* the finalizer's scratch copy */
    __shared__ int sWarpsDone;
    const int tid = threadIdx.x, lane = tid & 31;
    const int gridSize = (int)gridDim.x;
    DevState* st = A.st;
    int* hV = cs.hV; int* hI = cs.hI;
    const DynParams dyn{A.W, A.H, A.L, A.numDisc};

    if (tid == 0) sWarpsDone = 0;
    if (A.useHist) for (int c = tid; c < 2 * A.c1; c += TILE) hV[c] = 0;
    __syncthreads();                        /* the only CTA barrier of the launch */

#ifdef KGMT_PIPE_PROF
    long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pt = clock64();
#define PROF(k) { const long long now_ = clock64(); prof[k] += now_ - pt; pt = now_; }
    unsigned long long* trow = A.iterLog ? A.iterLog + 8 * 256 + 32 * ((int)blockIdx.x * WARPS + (tid >> 5)) : nullptr;
    int trk = 0;
#define TR(slot) { if (trow && lane == 0 && V.itr == KGMT_TRACE_ITR) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); trow[slot] = t_; } }
#define TRC(base) { if (trk < 8) TR((base) + 3 * trk) }
#else
#define PROF(k)
#define TR(slot)
#define TRC(base)
#endif
    PipeView V = pipe_load_view(st, lane);
    int itersDone = 0;
    int heldT = -1; bool heldSpec = false;
    ChunkCand held{};
    int rowsSeen = 0;

    auto propagate = [&](ChunkCand& cc) {
        if (!cc.live) return;
        if (COL == COL_GRID_SMEM)         cc.valid = propagate_edge(cc.x, cc.u, dyn, cs.gridS);
        else if (COL == COL_GRID_GLOBAL)  cc.valid = propagate_edge(cc.x, cc.u, dyn, cs.gridG);
        else if (COL == COL_BRUTE_SMEM)   cc.valid = propagate_edge(cc.x, cc.u, dyn, cs.allS);
        else                              cc.valid = propagate_edge(cc.x, cc.u, dyn, cs.allG);
    };

    while (V.stop == STOP_RUNNING && itersDone < maxIters) {
        const bool firstOfLaunch = itersDone == 0, lastOfLaunch = itersDone + 1 >= maxIters;
        const IterView it = pipe_iter_view(A, V);
        const int r = V.itr % 3;
        PipeIter* pi = A.pipe + r;
        PipeIter* pprev = A.pipe + (V.itr + 2) % 3;
        unsigned* ticket = A.ticket + r;
        int* done = A.blockDone + (size_t)r * A.blocksCap;
        const int* prefix = A.blockPrefix + (size_t)r * A.blocksCap;
        int* inserted = A.blockInserted + (size_t)r * A.blocksCap;
        const int numBlocks = (V.numChunks + BLK_CHUNKS - 1) / BLK_CHUNKS;
        bool scoresOk = false;              /* the scores follow the publication of the iteration's shape (scoreReady) */
        int pendingDone = -1;               /* scan block of the last finished chunk, not yet counted in blockDone */
        /* ---- chunks: the one held from the speculation, then tickets (next ticket fetched behind the current chunk) */
        int t = heldT;
        if (t < 0) { if (lane == 0) t = (int)atomicAdd(ticket, 1u); t = __shfl_sync(0xffffffffu, t, 0); }
#ifdef KGMT_PIPE_CHECK
        if (t < 0 && lane == 0) { printf("bad first ticket: itr %d t %d heldT %d ticket now %u r %d\n", V.itr, t, heldT, *(volatile unsigned*)ticket, r); __trap(); }
#endif
        PROF(7)
        TR(0)
#ifdef KGMT_PIPE_PROF
        trk = 0;
#endif
        while (t < V.numChunks) {
            int tn = 0;
            TRC(8)
            if (lane == 0) tn = (int)atomicAdd(ticket, 1u);
            ChunkCand cc;
            if (heldSpec && V.children == CHUNK && (V.mode == 1 || V.mode == 4)) {
                cc = held;                  /* stages 2-4 were done ahead of the publication, under the policy that held */
            } else {
                if (!firstOfLaunch) {       /* the parents of this chunk must be in the tree (previous iteration's closers) */
                    const int need = min(t * CHUNK + CHUNK - 1, V.M - 1) / V.children;
                    if (rowsSeen <= need) {
                        unsigned ns = 32;
                        while ((rowsSeen = ld_relaxed_warp(&pprev->rowsReady, lane)) <= need) { __nanosleep(ns); if (ns < 1024) ns <<= 1; }
                        fence_acq_rel();
                    }
                }
                PROF(0)
                cc = chunk_setup(A, it, t, lane);
                propagate(cc);
                PROF(1)
            }
            TRC(9)
            heldSpec = false;
            if (pendingDone >= 0) {
                __syncwarp();
                if (lane == 0) { __threadfence(); atomicAdd(&done[pendingDone], 1); }
                pendingDone = -1;
            }
#ifdef KGMT_PIPE_CHECK
            if (t < 0) { printf("pipe loop: t %d before finish, itr %d lane %d blk %d heldT %d tn %d\n", t, V.itr, lane, (int)blockIdx.x, heldT, tn); __trap(); }
#endif
            chunk_finish<RECORD, false>(A, it, cc, t, lane, hV, hI, scoresOk);
            PROF(2)
            TRC(10)
#ifdef KGMT_PIPE_PROF
            ++trk;
#endif
            /* the chunk is signed off (blockDone) one chunk LATER, after the next chunk's stages 2-4: by then its
             * staging stores have long landed, so the release fence has nothing to wait for */
            pendingDone = t / BLK_CHUNKS;
            PROF(4)
            t = __shfl_sync(0xffffffffu, tn, 0);
#ifdef KGMT_PIPE_CHECK
            if (t < 0 && lane == 0) { printf("bad next ticket: itr %d t %d ticket now %u r %d\n", V.itr, t, *(volatile unsigned*)ticket, r); __trap(); }
#endif
        }
        TR(1)
        heldT = -1; heldSpec = false;
        __syncwarp();
        if (pendingDone >= 0) {
            if (lane == 0) { __threadfence(); atomicAdd(&done[pendingDone], 1); }
            pendingDone = -1;
        }

        /* ---- no chunk left for this warp: sign off; last warp of the CTA flushes the histograms; last CTA finalizes */
        __syncwarp();
        int lastWarp = 0;
        if (lane == 0) { __threadfence_block(); lastWarp = (atomicAdd(&sWarpsDone, 1) == WARPS - 1); }
        if (__shfl_sync(0xffffffffu, lastWarp, 0)) {
            __threadfence_block();
            if (A.useHist) {
                for (int c = lane; c < A.c1; c += 32) {
                    const int v = hV[c], iv = hI[c];
                    if (v | iv) {
                        atomicAdd(&A.R1[c], v + iv);
                        if (v) { atomicAdd(&A.R1Valid[c], v); A.R1Avail[c] = 1; }
                        if (iv) atomicAdd(&A.R1Invalid[c], iv);
                        hV[c] = 0; hI[c] = 0;
                    }
                }
            }
            __syncwarp();
            int lastCta = 0;
            if (lane == 0) {
                sWarpsDone = 0;
                __threadfence();
                lastCta = (atomicAdd(&pi->ctasDone, 1) == gridSize - 1);
            }
            if (__shfl_sync(0xffffffffu, lastCta, 0))
                pipe_finalize(A, V, it, pi, pprev, firstOfLaunch, lastOfLaunch, &sS, lane);
        }

        PROF(3)
        TR(2)
        /* ---- ordered insertion (updateG) of this iteration's accepted rows, by the warps that ran out of chunks: units
         *      of 32 chunks by ticket; a unit waits until the rows before its block are known (prefix chain) */
        {
            const int numUnits = (V.numChunks + 31) / 32;
            for (;;) {
                int u = 0;
                if (lane == 0) u = atomicAdd(&pi->insTicket, 1);
                u = __shfl_sync(0xffffffffu, u, 0);
                if (u >= numUnits) break;
                const int b = u / SUBS;
                unsigned ns = 32;
                while (ld_acquire_warp(&pi->chainW, lane) <= b) {
                    pipe_chain(A, it, pi, numBlocks, lane);
                    if (ld_acquire_warp(&pi->chainW, lane) > b) break;
                    __nanosleep(ns); if (ns < 512) ns <<= 1;
                }
                insert_unit_warp(A, it, b, u % SUBS, __ldcg(&prefix[b]), lane);
                __syncwarp();
                int full = 0;
                if (lane == 0) { __threadfence(); full = (atomicAdd(&inserted[b], 1) + 1 == units_in_block(b, V.numChunks)); }
                if (__shfl_sync(0xffffffffu, full, 0)) pipe_rows(A, it, pi, numBlocks, lane);
            }
        }
        PROF(4)
        TR(3)
        /* ---- one chunk of the NEXT iteration ahead of its publication */
        const int wantEpoch = V.iterationsDone + 1;
        rowsSeen = 0;
        if (!lastOfLaunch) {
            int ts = 0;
            if (lane == 0) ts = (int)atomicAdd(A.ticket + (V.itr + 1) % 3, 1u);
            ts = __shfl_sync(0xffffffffu, ts, 0);
#ifdef KGMT_PIPE_CHECK
            if (ts < 0 && lane == 0) { printf("bad spec ticket: itr %d ts %d\n", V.itr, ts); __trap(); }
#endif
            heldT = ts;
            int spec = 0;
            if (lane == 0) {
                unsigned ns = 64;
                for (;;) {
                    if (ld_relaxed_s32(&pi->rowsReady) > ts) { spec = 1; break; }
                    if (ld_relaxed_s32(&st->pipeEpoch) >= wantEpoch) break;
                    __nanosleep(ns); if (ns < PIPE_SLEEP_CAP) ns <<= 1;
                }
            }
            spec = __shfl_sync(0xffffffffu, spec, 0);
            PROF(5)
            TR(4)
            if (spec) {
                fence_acq_rel();
                IterView sv = it;           /* the next iteration if the 32-children policy holds (validated above) */
                sv.itr = V.itr + 1; sv.frontierStart = V.treeSize; sv.children = CHUNK;
                sv.M = 0x7fffffff; sv.key0 = A.seed + (uint32_t)(V.itr + 1);
                held = chunk_setup(A, sv, ts, lane);
                propagate(held);
                heldSpec = true;
                PROF(1)
                TR(5)
            }
        }

        /* ---- wait for the publication of the next iteration */
        if (lane == 0) {
            unsigned ns = 64;
            while (ld_relaxed_s32(&st->pipeEpoch) < wantEpoch) { __nanosleep(ns); if (ns < PIPE_SLEEP_CAP) ns <<= 1; }
        }
        TR(6)
        __syncwarp();
        fence_acq_rel();
        V = pipe_load_view(st, lane);
        itersDone += 1;
        if (V.stop == STOP_RUNNING) pipe_scores_help(A, V.itr, A.pipe + V.itr % 3, sP, lane);
        PROF(6)
    }
#ifdef KGMT_PIPE_PROF
    if (lane == 0 && A.iterLog) for (int k = 0; k < 8; ++k) atomicAdd(&A.iterLog[8 * 255 + k], (unsigned long long)prof[k]);
#endif
}

/* the same launch contract (cooperative: every CTA resident), barrier-free loop */
template <int COL, bool RECORD>
__global__ void __launch_bounds__(TILE, 4) expand_pipe_kernel(const KArgs A, int maxIters) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t sBar;
    const ColSet cs = stage_collision<COL>(A, smem_raw, &sBar);
    run_plan_pipe<COL, RECORD>(A, maxIters, cs);
}

