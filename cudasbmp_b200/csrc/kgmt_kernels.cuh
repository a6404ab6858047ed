/* kgmt_kernels.cuh — the sm_100a kernels of the KGMT tree-expansion step.
 *
 * One fused kernel per iteration replaces the reference's
 *   scan(R1Avail)+updateR1, scan(G)+findInd, propagateG|propagateGV2,
 *   scan(GNew)+findInd, updateG      (src/planners/KGMT.cu:118-259)
 * and, launched cooperatively, the whole host loop of KGMT::plan.
 *
 *   stage 1  frontier  = the contiguous node range appended by the previous
 *            iteration (no scan: KGMT.cu:378,451,568,582 make G exactly that);
 *            R1 scores are produced by the CTA that finishes an iteration last.
 *   stage 2  stateless Philox4x32-10 per candidate slot (kgmt_device.cuh).
 *   stage 3  Euler car dynamics on SoA float4 node storage.
 *   stage 4  step-bbox vs obstacle AABBs staged in shared memory by bulk
 *            async copies (cp.async.bulk + mbarrier): exhaustive, or culled
 *            through a uniform grid (identical flags).
 *   stage 5  region counters (R1 family in per-CTA shared-memory histograms,
 *            R2 family with global reductions), accept test on the
 *            iteration-start snapshot, ordered insertion through a single-pass
 *            decoupled look-back scan over candidate tiles (ballot/popc inside
 *            the tile) — accepted nodes go straight from registers to the tree.
 *
 * Canonical semantics where the reference races: SURVEY.md Appendix B.
 */
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kgmt_device.cuh"

namespace kgmt {
namespace cg = cooperative_groups;

constexpr int TILE = 256;                 /* candidates per tile == threads per CTA */
constexpr int WARPS = TILE / 32;

enum { STOP_RUNNING = 0, STOP_SOLVED = 1, STOP_TREE_FULL = 2, STOP_ITER_LIMIT = 3, STOP_FRONTIER_EMPTY = 4 };
enum { COL_GRID_SMEM = 0, COL_GRID_GLOBAL = 1, COL_BRUTE_SMEM = 2, COL_BRUTE_GLOBAL = 3 };
enum { FLAG_VALID = 1, FLAG_ACCEPT = 2 };

/* device-resident planner scalars (one per context) */
struct DevState {
    int treeSize, frontierStart, frontierCount, itr;           /* itr = iteration about to run (1-based) */
    int stop, goalIdx; float costToGoal; float R1Threshold;
    int mode, children, M, numTiles;                            /* shape of the iteration about to run */
    unsigned ticket, ctasDone; unsigned epoch; int forceChildren;
    unsigned long long goalBest;                                /* (cost bits << 32) | tree index, ~0 = none */
    long long expansions;
    int lastMode, lastChildren, lastFrontier, lastM, lastAccepted, lastItr, iterationsDone, pad;
};

struct KArgs {
    /* tree, SoA */
    float4* treeState;            /* (x, y, theta, v) */
    float4* treeCtrl;             /* (a, steering, duration, cost) */
    int*    treeParent;
    /* occupancy maps */
    int *R1, *R1Valid, *R1Invalid, *R1Avail, *R1Cov; float* R1Score;
    int *R2, *R2Valid, *R2Invalid; unsigned* R2Stamp;
    /* per-candidate records (null unless recording) */
    float4* candState; float4* candCtrl; int* candParent; int* candR1; int* candR2; unsigned char* candFlags;
    /* scan */
    unsigned long long* tileStatus;
    DevState* st;
    /* collision */
    const float4* obstacles; int K;
    const int* cellStart; const float4* cellItems; int cullC; float cullInvX, cullInvY; int cellStartInts; int numItems;
    int obsTile;                  /* obstacles per shared-memory tile (stream mode) */
    unsigned long long* iterLog;  /* [256][2]: (globaltimer ns, M<<32 | accepted) per finished iteration; null = off */
    /* parameters */
    float W, H, L, R1Size, R2Size, goalX, goalY, goalR;
    int N, n, c1, numDisc, maxTree, numIterations, useHist;
    uint32_t seed;
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

/* ------------------------------------------------------------------------- stage 1 ----
 * R1 scores, updateR1 (KGMT.cu:487-538) for any N, by one CTA of TILE threads.
 * covR uses the running count of available R2 cells (R1Cov) instead of re-summing
 * n*n flags (:510-514) — the same integer.  The sum order is fixed (the reference's
 * cub::BlockReduce order is unspecified): p[t] = sum_k score[t+1024k], then a
 * stride-halving tree (DESIGN.md "scores"; the CPU checker restates the same order). */
__device__ void scores_block(const KArgs& A, float* p /* smem [1024] */) {
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
            A.R1Score[c] = score;                       /* raw; normalised below */
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
        A.R1Score[c] = (__ldcg(&A.R1Avail[c]) == 0) ? 1.0f : __fdiv_rn(A.R1Score[c], total);
    __syncthreads();
}

/* expansion policy, KGMT.cu:151-158 (canonical prefix mode: SURVEY.md App. B #7) */
__device__ __forceinline__ void expansion_shape(int active, int treeSize, int maxTree, int forceChildren,
                                                int& mode, int& children, int& M) {
    const int remaining = maxTree - treeSize;
    if (forceChildren > 0) { mode = 4; children = forceChildren; M = active * forceChildren; return; }
    if (32LL * active > (long long)remaining) {
        const int it = __float2int_rz(__fdiv_rn((float)remaining, (float)active));
        if (it >= 1) { mode = 2; children = it; M = active * it; }
        else         { mode = 3; children = 1;  M = remaining; }
    } else { mode = 1; children = 32; M = 32 * active; }
}

/* end of an iteration: executed by every thread of the CTA that finished last.
 * KGMT.cu:249-259 + the next iteration's :119-136. */
__device__ void finalize_iteration(const KArgs& A, float* p) {
    volatile DevState* st = A.st;
    __shared__ int sRun;
    if (threadIdx.x == 0) {
        const int M = st->M, numTiles = st->numTiles;
        const int accepted = (int)(unsigned)ld_relaxed_u64(&A.tileStatus[numTiles - 1]);   /* inclusive total */
        st->lastMode = st->mode; st->lastChildren = st->children; st->lastFrontier = st->frontierCount;
        st->lastM = M; st->lastAccepted = accepted; st->lastItr = st->itr;
        if (A.iterLog && st->iterationsDone < 255) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            A.iterLog[2 * st->iterationsDone] = t;
            A.iterLog[2 * st->iterationsDone + 1] = ((unsigned long long)(unsigned)M << 32) | (unsigned)accepted;
        }
        st->iterationsDone += 1;
        st->expansions += M;
        st->frontierStart = st->treeSize;
        st->frontierCount = accepted;
        st->treeSize += accepted;                                                   /* :249 */
        const unsigned long long gb = st->goalBest;
        if (gb != ~0ull) { st->costToGoal = __uint_as_float((unsigned)(gb >> 32)); st->goalIdx = (int)(unsigned)gb; }
        int stop = STOP_RUNNING;
        if (st->costToGoal != 0.0f)             stop = STOP_SOLVED;                 /* :252 */
        else if (st->treeSize >= A.maxTree)     stop = STOP_TREE_FULL;              /* :255 */
        else if (accepted == 0)                 stop = STOP_FRONTIER_EMPTY;
        else if (st->itr >= A.numIterations)    stop = STOP_ITER_LIMIT;             /* :118 */
        st->stop = stop;
        if (stop == STOP_RUNNING) {
            st->itr += 1;                                                           /* :119 */
            int mode, children, Mn;
            const int ts = st->treeSize, fc = st->forceChildren;
            expansion_shape(accepted, ts, A.maxTree, fc, mode, children, Mn);
            st->mode = mode; st->children = children; st->M = Mn; st->numTiles = (Mn + TILE - 1) / TILE;
        }
        st->epoch += 1;
        st->ticket = 0; st->ctasDone = 0;
        sRun = (stop == STOP_RUNNING);
    }
    __syncthreads();
    if (sRun) scores_block(A, p);
    __threadfence();
}

/* ------------------------------------------------------------------ stages 2-5 fused -- */
template <int COL, bool LOOP, bool RECORD>
__global__ void __launch_bounds__(TILE) expand_kernel(const KArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t sBar;
    __shared__ int sWarpCount[WARPS];
    __shared__ int sTile, sBase, sLast;
    __shared__ float sP[1024];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    DevState* st = A.st;

    /* shared-memory carve-up: [R1 histograms][collision data] */
    int* hV = reinterpret_cast<int*>(smem_raw);
    int* hI = hV + (A.useHist ? A.c1 : 0);
    unsigned char* colBase = smem_raw + (A.useHist ? ((2 * A.c1 * 4 + 15) & ~15) : 0);

    /* stage the collision structure once per launch (bulk async copy, mbarrier completion) */
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
                bulk_g2s_chunked(colBase, A.cellStart, bytes0, &sBar);
                bulk_g2s_chunked(colBase + bytes0, A.cellItems, bytes1, &sBar);
            } else {
                bulk_g2s_chunked(colBase, A.obstacles, bytes0, &sBar);
            }
        }
        mbar_wait(&sBar, 0);
        if (COL == COL_GRID_SMEM) {
            sCellStart = reinterpret_cast<const int*>(colBase);
            sItems = reinterpret_cast<const float4*>(colBase + bytes0);
        } else {
            sObs = reinterpret_cast<const float4*>(colBase);
        }
    }

    const DynParams dyn{A.W, A.H, A.L, A.numDisc};

    for (;;) {
        /* iteration scalars (written by the previous finalize, ordered by the launch boundary or grid.sync) */
        const int stop = *(volatile int*)&st->stop;
        if (stop != STOP_RUNNING) break;
        const int itr = *(volatile int*)&st->itr;
        const int treeSize = *(volatile int*)&st->treeSize;
        const int frontierStart = *(volatile int*)&st->frontierStart;
        const int children = *(volatile int*)&st->children;
        const int M = *(volatile int*)&st->M;
        const int numTiles = *(volatile int*)&st->numTiles;
        const unsigned epoch = *(volatile unsigned*)&st->epoch;
        const uint32_t key0 = A.seed + (uint32_t)itr;
        const unsigned stampNew = (unsigned)itr + 1u;          /* R2 cells first reached in this iteration */

        if (A.useHist) {
            for (int c = tid; c < 2 * A.c1; c += TILE) hV[c] = 0;
        }
        __syncthreads();

        for (;;) {
            if (tid == 0) sTile = (int)atomicAdd(&st->ticket, 1u);
            __syncthreads();
            const int tile = sTile;
            if (tile >= numTiles) break;
            const int s = tile * TILE + tid;
            const bool live = s < M;

            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            Controls u{0.f, 0.f, 0.f, 0.f};
            int parent = -1, r1 = -1, r2 = -1;
            bool valid = false, accept = false;
            float parentCost = 0.f;
            if (live) {
                parent = frontierStart + s / children;                             /* KGMT.cu:374-376 / :454 */
                x = __ldcg(&A.treeState[parent]);              /* L2-coherent: written by other SMs last iteration */
                parentCost = __ldcg(&A.treeCtrl[parent]).w;
                u = sample_controls((uint32_t)s, key0);
                if (COL == COL_GRID_SMEM) {
                    const CollideGrid col{sCellStart, sItems, A.cullC, A.cullInvX, A.cullInvY};
                    valid = propagate_edge(x, u, dyn, col);
                } else if (COL == COL_GRID_GLOBAL) {
                    const CollideGrid col{A.cellStart, A.cellItems, A.cullC, A.cullInvX, A.cullInvY};
                    valid = propagate_edge(x, u, dyn, col);
                } else if (COL == COL_BRUTE_SMEM) {
                    const CollideSmemAll col{sObs, A.K};
                    valid = propagate_edge(x, u, dyn, col);
                } else {
                    const CollideSmemAll col{A.obstacles, A.K};                   /* global/L1 path */
                    valid = propagate_edge(x, u, dyn, col);
                }
                r1 = region_r1(x.x, x.y, A.R1Size, A.N);                           /* KGMT.cu:390 */
                r2 = region_r2(x.x, x.y, r1, A.R1Size, A.N, A.R2Size, A.n);        /* KGMT.cu:391 */
                /* maps + accept, KGMT.cu:392-411 on the iteration-start snapshot (App. B #1,#2) */
                if (r1 >= 0) {
                    if (valid) {
                        accept = u.u3 <= __ldcg(&A.R1Score[r1]);
                        if (r2 >= 0) {
                            const unsigned stamp = __ldcg(&A.R2Stamp[r2]);
                            if (stamp == 0u || stamp > (unsigned)itr) accept = true;       /* unavailable at iteration start */
                            if (stamp == 0u) {
                                if (atomicCAS(&A.R2Stamp[r2], 0u, stampNew) == 0u) atomicAdd(&A.R1Cov[r1], 1);
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
            }

            /* ordered compaction of accepted candidates: ballot/popc in the warp, look-back across tiles */
            const unsigned bal = __ballot_sync(0xffffffffu, accept);
            if (lane == 0) sWarpCount[warp] = __popc(bal);
            __syncthreads();
            int warpOff = 0, agg = 0;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) { const int c = sWarpCount[w]; if (w < warp) warpOff += c; agg += c; }
            if (warp == 0) {
                int base = 0;
                const unsigned long long tagA = ((unsigned long long)(epoch * 4u + 1u)) << 32;
                const unsigned long long tagP = ((unsigned long long)(epoch * 4u + 2u)) << 32;
                if (tile == 0) {
                    if (lane == 0) st_relaxed_u64(&A.tileStatus[0], tagP | (unsigned)agg);
                } else {
                    if (lane == 0) st_relaxed_u64(&A.tileStatus[tile], tagA | (unsigned)agg);
                    int j = tile - 1 - lane;
                    for (;;) {
                        int flag = 2, val = 0;
                        if (j >= 0) {
                            unsigned long long w;
                            do { w = ld_relaxed_u64(&A.tileStatus[j]); } while ((unsigned)(w >> 34) != epoch);
                            flag = (int)((w >> 32) & 3u); val = (int)(unsigned)w;
                        }
                        const unsigned incl = __ballot_sync(0xffffffffu, flag == 2);
                        const int first = incl ? (__ffs(incl) - 1) : 32;
                        int contrib = (lane <= first) ? val : 0;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
                        base += contrib;
                        if (incl) break;
                        j -= 32;
                    }
                    if (lane == 0) st_relaxed_u64(&A.tileStatus[tile], tagP | (unsigned)(base + agg));
                }
                if (lane == 0) sBase = base;
            }
            __syncthreads();

            if (accept) {                                                          /* updateG, KGMT.cu:555-591 */
                const int dst = treeSize + sBase + warpOff + __popc(bal & ((1u << lane) - 1u));
                const float cost = __fadd_rn(parentCost, u.duration);              /* :585-586, :631-633 */
                A.treeState[dst] = x;
                A.treeCtrl[dst] = make_float4(u.a, u.steering, u.duration, cost);
                A.treeParent[dst] = parent;
                if (in_goal(x.x, x.y, A.goalX, A.goalY, A.goalR))                  /* :589; canonical min (App. B #5) */
                    atomicMin(&st->goalBest, ((unsigned long long)__float_as_uint(cost) << 32) | (unsigned)dst);
            }
            if (RECORD && live) {
                A.candState[s] = x;
                A.candCtrl[s] = make_float4(u.a, u.steering, u.duration, u.u3);
                A.candParent[s] = parent;
                A.candR1[s] = r1; A.candR2[s] = r2;
                A.candFlags[s] = (unsigned char)((valid ? FLAG_VALID : 0) | (accept ? FLAG_ACCEPT : 0));
            }
        }

        /* flush the R1-family histograms (R1 = R1Valid + R1Invalid increments, KGMT.cu:392,406,409) */
        if (A.useHist) {
            for (int c = tid; c < A.c1; c += TILE) {
                const int v = hV[c], iv = hI[c];
                if (v | iv) {
                    atomicAdd(&A.R1[c], v + iv);
                    if (v) { atomicAdd(&A.R1Valid[c], v); A.R1Avail[c] = 1; }
                    if (iv) atomicAdd(&A.R1Invalid[c], iv);
                }
            }
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) sLast = (atomicAdd(&st->ctasDone, 1u) == gridDim.x - 1u);
        __syncthreads();
        if (sLast) { __threadfence(); finalize_iteration(A, sP); }
        if (!LOOP) break;
        cg::this_grid().sync();
    }
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
    for (long long s = (long long)blockIdx.x * TILE + tid; s < M; s += (long long)gridDim.x * TILE) {
        float4 x = __ldg(&parents[s / children]);
        const Controls u = sample_controls(slot0 + (uint32_t)s, key0);
        bool valid;
        if (COL == COL_GRID_SMEM) {
            const CollideGrid col{sCellStart, sItems, A.cullC, A.cullInvX, A.cullInvY};
            valid = propagate_edge(x, u, dyn, col);
        } else if (COL == COL_GRID_GLOBAL) {
            const CollideGrid col{A.cellStart, A.cellItems, A.cullC, A.cullInvX, A.cullInvY};
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

/* ------------------------------------------------------------------ setup kernels ------ */
/* root insertion, KGMT.cu:85-114, then the first iteration's shape and scores */
__global__ void __launch_bounds__(TILE) begin_kernel(const KArgs A, float4 rootState, float4 rootCtrl) {
    __shared__ float sP[1024];
    DevState* st = A.st;
    if (threadIdx.x == 0) {
        A.treeState[0] = rootState;                                                /* :85 */
        A.treeCtrl[0] = make_float4(rootCtrl.x, rootCtrl.y, rootCtrl.z, 0.0f);
        A.treeParent[0] = -1;
        const int r1 = region_r1(rootState.x, rootState.y, A.R1Size, A.N);         /* :88 */
        const int r2 = region_r2(rootState.x, rootState.y, r1, A.R1Size, A.N, A.R2Size, A.n);   /* :89 */
        if (r1 >= 0) { A.R1[r1] = 1; A.R1Avail[r1] = 1; A.R1Valid[r1] = 1; }       /* :94,95,97 */
        if (r2 >= 0 && A.R2Stamp[r2] == 0u) { A.R2Stamp[r2] = 1u; A.R1Cov[r1] += 1; }  /* :96 */
        st->treeSize = 1; st->frontierStart = 0; st->frontierCount = 1; st->itr = 1;
        st->goalIdx = -1; st->costToGoal = 0.0f; st->goalBest = ~0ull;
        st->expansions = 0; st->iterationsDone = 0;
        st->lastMode = st->lastChildren = st->lastFrontier = st->lastM = st->lastAccepted = st->lastItr = 0;
        st->ticket = 0; st->ctasDone = 0; st->epoch += 1;
        int stop = STOP_RUNNING;
        if (A.numIterations <= 0) stop = STOP_ITER_LIMIT;
        else if (1 >= A.maxTree) stop = STOP_TREE_FULL;
        st->stop = stop;
        int mode = 0, children = 1, M = 0;
        if (stop == STOP_RUNNING) expansion_shape(1, 1, A.maxTree, st->forceChildren, mode, children, M);
        st->mode = mode; st->children = children; st->M = M; st->numTiles = (M + TILE - 1) / TILE;
    }
    __syncthreads();
    scores_block(A, sP);
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
        st->goalIdx = -1; st->costToGoal = 0.0f; st->goalBest = ~0ull;
        st->expansions = 0; st->iterationsDone = 0;
        st->lastMode = st->lastChildren = st->lastFrontier = st->lastM = st->lastAccepted = st->lastItr = 0;
        st->ticket = 0; st->ctasDone = 0; st->epoch += 1;
        int stop = STOP_RUNNING;
        if (A.numIterations <= 0) stop = STOP_ITER_LIMIT;
        else if (count >= A.maxTree) stop = STOP_TREE_FULL;
        st->stop = stop;
        int mode = 0, children = 1, M = 0;
        if (stop == STOP_RUNNING) expansion_shape(count, count, A.maxTree, st->forceChildren, mode, children, M);
        st->mode = mode; st->children = children; st->M = M; st->numTiles = (M + TILE - 1) / TILE;
    }
    __syncthreads();
    scores_block(A, sP);
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
    scores_block(A, sP);
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
