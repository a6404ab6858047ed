/* kgmt_capi.cu — the C ABI of libkgmt_b200.so (include/kgmt_c.h) over the sm_100a
 * kernels in kgmt_kernels.cuh.  Host side of the KGMT tree-expansion path:
 * context = what the reference's KGMT constructor allocates
 * (src/planners/KGMT.cu:10-78), kgmt_plan = KGMT::plan (:80-317).
 *
 * No CPU fallback: every entry point that computes launches CUDA kernels; a
 * missing device is KGMT_ERR_CUDA.  Nothing here touches oracle/.
 */
#include "kgmt_c.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "kgmt_kernels.cuh"

using namespace kgmt;

struct kgmt_comm_state;              /* kgmt_comm.inl */

struct kgmt_ctx {
    kgmt_params p;
    kgmt_comm_state* comm = nullptr;   /* NCCL communicator + exchange buffers (kgmt_comm_init) */
    int device = 0, numSMs = 0, maxCand = 0, c1 = 0;
    size_t c2 = 0;
    float R1Size = 0.f, R2Size = 0.f;
    cudaStream_t stream = nullptr;      /* the stream every call launches on */
    cudaStream_t ownStream = nullptr;   /* created by kgmt_create; `stream` unless the caller installed its own */
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t evState = nullptr, evPath = nullptr;      /* kgmt_plan: scalars landed / prefetched path landed (no timing) */
    /* device memory */
    float4 *treeState = nullptr, *treeCtrl = nullptr;
    int* treeParent = nullptr;
    int* mapSlab = nullptr;            /* [R1,R1Valid,R1Invalid,R1Avail,R1Cov,R1Score | R2,R2Valid,R2Invalid,R2Stamp] */
    int* mapSlabCkpt = nullptr;
    size_t mapSlabInts = 0;            /* the region maps (checkpointed, exchanged) */
    size_t slabInts = 0;               /* + scan bookkeeping behind them */
    int *R1 = nullptr, *R1Valid = nullptr, *R1Invalid = nullptr, *R1Avail = nullptr, *R1Cov = nullptr;
    float* R1Score[2] = {nullptr, nullptr};
    int *R2 = nullptr, *R2Valid = nullptr, *R2Invalid = nullptr;
    unsigned* R2Stamp = nullptr;
    float4 *candState = nullptr, *candCtrl = nullptr;
    int *candParent = nullptr, *candR1 = nullptr, *candR2 = nullptr;
    unsigned char* candFlags = nullptr;
    bool recordAllocated = false;
    unsigned* chunkMask = nullptr; int* blockSum = nullptr; unsigned* ticket = nullptr;
    int raceId = 0;                    /* > 0 while kgmt_peer_race runs */
    size_t chunksCap = 0, blocksCap = 0;
    float4 *stageState = nullptr, *stageCtrl = nullptr;
    unsigned long long* iterLog = nullptr;
    DevState* dState = nullptr;
    DevState* hState = nullptr;        /* pinned */
    DevState ckptState{};
    bool haveCkpt = false;
    /* obstacles and the cull grid */
    float4* dObs = nullptr; int K = 0; size_t obsCap = 0;
    int* dCellStart = nullptr; float4* dCellItems = nullptr; size_t cellStartCap = 0, cellItemsCap = 0;
    int cullC = 1, cellStartInts = 4, numItems = 0;
    float cullInvX = 0.f, cullInvY = 0.f;
    float* hObsPinned = nullptr; size_t hObsPinnedCap = 0;            /* pinned staging of kgmt_set_obstacles_host */
    int cullExpect = -1;                                              /* >= 0: item count computed on the host, to be checked against *hCullTotal */
    int* hCullTotal = nullptr; int* dCullTotal = nullptr;             /* item count of the cull grid: pinned host word, device word */
    std::vector<unsigned char> hPath;  /* staging of kgmt_extract_path */
    /* solution path of the last kgmt_plan, traced behind the planner kernel and copied with the scalars: valid while no
     * kernel of this library has run since (launch counter) and the scalars still name the same goal node */
    unsigned char* dPathPre = nullptr; unsigned char* hPathPre = nullptr;
    long long pathPreLaunches = -1; int pathPreGoal = -1, pathPreTree = 0;
    /* staging */
    void* scratch = nullptr; size_t scratchBytes = 0;
    float4* dParents = nullptr; size_t parentsCap = 0;
    /* launch configuration */
    int col = COL_GRID_SMEM; size_t smemBytes = 0; int useHist = 0;
    int gridLoop = 0, gridMax = 0;
    bool configured = false;
    int cfgCol = -1; size_t cfgSmem = 0;      /* what the launch configuration was resolved for */
    bool begun = false;
    float goal[7] = {0};
    long long launches = 0;
    int planLaunches = 0;
    size_t dirtyCand = 0;              /* candidate-record rows a plan may have written since the last clear */
    /* batched planning workspaces (kgmt_plan_batch) */
    struct Batch {
        int numWs = 0, clusterSize = 0, Qcap = 0, maxPath = 0;
        float4 *treeState = nullptr, *treeCtrl = nullptr, *stageState = nullptr, *stageCtrl = nullptr;
        int *treeParent = nullptr, *mapSlab = nullptr, *blockSum = nullptr, *queryTicket = nullptr, *wsQuery = nullptr, *pathLen = nullptr;
        unsigned long long* launchT0 = nullptr;
        unsigned *chunkMask = nullptr, *ticket = nullptr;
        float4 *initState = nullptr, *initCtrl = nullptr; float2* goalXY = nullptr; uint32_t* seeds = nullptr;
        DevState* states = nullptr; float* paths = nullptr;
        size_t mapIntsStride = 0;
        /* launch configuration last resolved: the attribute / occupancy queries cost ~100 us of driver calls per batch */
        int cfgCluster = 0, cfgCol = -1, cfgMaxClusters = 0; size_t cfgSmem = 0;
    } batch;
    unsigned epochBase = 0;
    /* sharded expansion (kgmt_shard_*) */
    int* shardPrefix = nullptr; int* shardTotal = nullptr; int* hShardTotal = nullptr;   /* [blocksCap], [1], pinned [1] */
    int shardBlkLo = 0, shardBlkHi = 0, shardAccepted = -1, shardGrid = 0;
    int fusedGrid = 0;                 /* persistent grid of expand_sharded_kernel (cooperative) */
    /* sharded expansion over peer memory (kgmt_peer_*) */
    struct Peer {
        int rank = -1, world = 0, seq = 0; bool ipc = false, inFlight = false;
        unsigned char* block = nullptr; size_t blockBytes = 0, mailOff = 0, raceOff = 0;   /* [delta slab | mailbox[PEER_MAX] | race flag] */
        int** dRaceFlags = nullptr;                                              /* device table of every rank's race flag */
        PeerPlan* plan = nullptr; PeerPlan* hPlan = nullptr;                    /* device, pinned host */
        void* opened[PEER_MAX][5] = {};                                          /* cudaIpcOpenMemHandle results to close */
        PeerArgs args{};
    } peer;
    DevState resetState{};
    char err[512] = {0};
};

/* -------------------------------------------------------------------------------- errors */
static int fail(kgmt_ctx* c, int code, const char* fmt, ...) {
    if (c) {
        va_list ap; va_start(ap, fmt);
        vsnprintf(c->err, sizeof(c->err), fmt, ap);
        va_end(ap);
    }
    return code;
}
#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(ctx, KGMT_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)

#define KGMT_ITERLOG_BYTES (256 * 64)
static const size_t MAX_DYN_SMEM = 227u * 1024u - 8u * 1024u;   /* leave room for static shared + reserve */

/* -------------------------------------------------------------------------------- kernels table */
typedef void (*expand_fn)(const KArgs, int);
template <int COL> static expand_fn pick_expand(bool rec) {
    return rec ? (expand_fn)expand_kernel<COL, true> : (expand_fn)expand_kernel<COL, false>;
}
static expand_fn expand_entry(int col, bool rec) {
    switch (col) {
        case COL_GRID_SMEM: return pick_expand<COL_GRID_SMEM>(rec);
        case COL_GRID_GLOBAL: return pick_expand<COL_GRID_GLOBAL>(rec);
        case COL_BRUTE_SMEM: return pick_expand<COL_BRUTE_SMEM>(rec);
        case COL_BRUTE_STREAM: return pick_expand<COL_BRUTE_STREAM>(rec);
        default: return pick_expand<COL_BRUTE_GLOBAL>(rec);
    }
}
typedef void (*prop_fn)(const KArgs, const float4*, long long, int, uint32_t, uint32_t);
static prop_fn propagate_entry(int col) {
    switch (col) {
        case COL_GRID_SMEM: return propagate_only_kernel<COL_GRID_SMEM>;
        case COL_GRID_GLOBAL: return propagate_only_kernel<COL_GRID_GLOBAL>;
        case COL_BRUTE_SMEM: return propagate_only_kernel<COL_BRUTE_SMEM>;
        default: return propagate_only_kernel<COL_BRUTE_GLOBAL>;
    }
}

typedef void (*shard_fn)(const KArgs, int, int);
static shard_fn shard_entry(int col) {
    switch (col) {
        case COL_GRID_SMEM: return shard_expand_kernel<COL_GRID_SMEM>;
        case COL_GRID_GLOBAL: return shard_expand_kernel<COL_GRID_GLOBAL>;
        case COL_BRUTE_SMEM: return shard_expand_kernel<COL_BRUTE_SMEM>;
        default: return shard_expand_kernel<COL_BRUTE_GLOBAL>;
    }
}

typedef void (*fused_fn)(const KArgs, const KArgs, const PeerArgs, int, int, size_t);
static fused_fn fused_entry(int col) {
    switch (col) {
        case COL_GRID_SMEM: return expand_sharded_kernel<COL_GRID_SMEM>;
        case COL_GRID_GLOBAL: return expand_sharded_kernel<COL_GRID_GLOBAL>;
        case COL_BRUTE_SMEM: return expand_sharded_kernel<COL_BRUTE_SMEM>;
        default: return expand_sharded_kernel<COL_BRUTE_GLOBAL>;
    }
}

constexpr int PATH_PRE_ROWS = 128;                               /* rows of the path prefetched by kgmt_plan */
constexpr size_t PATH_PRE_BYTES = 16 + (size_t)PATH_PRE_ROWS * 28;

static KArgs make_args(const kgmt_ctx* c) {
    KArgs A{};
    A.treeState = c->treeState; A.treeCtrl = c->treeCtrl; A.treeParent = c->treeParent;
    A.R1 = c->R1; A.R1Valid = c->R1Valid; A.R1Invalid = c->R1Invalid; A.R1Avail = c->R1Avail; A.R1Cov = c->R1Cov;
    A.R1Score[0] = c->R1Score[0]; A.R1Score[1] = c->R1Score[1];
    A.R2 = c->R2; A.R2Valid = c->R2Valid; A.R2Invalid = c->R2Invalid; A.R2Stamp = c->R2Stamp;
    A.R2StampDelta = nullptr;
    A.candState = c->candState; A.candCtrl = c->candCtrl; A.candParent = c->candParent;
    A.candR1 = c->candR1; A.candR2 = c->candR2; A.candFlags = c->candFlags;
    A.chunkMask = c->chunkMask; A.blockSum = c->blockSum; A.ticket = c->ticket;
    A.stageState = c->stageState; A.stageCtrl = c->stageCtrl;
    A.chunksCap = (int)c->chunksCap; A.blocksCap = (int)c->blocksCap; A.maxCand = c->maxCand; A.totalWarps = c->gridLoop * WARPS;
    A.raceId = c->raceId; A.raceWorld = c->peer.world; A.raceRank = c->peer.rank; A.raceFlags = c->peer.dRaceFlags;
    A.st = c->dState;
    A.obstacles = c->dObs; A.K = c->K;
    A.cellStart = c->dCellStart; A.cellItems = c->dCellItems; A.cullC = c->cullC;
    A.cullInvX = c->cullInvX; A.cullInvY = c->cullInvY; A.cellStartInts = c->cellStartInts; A.numItems = c->numItems;
    A.obsTile = (c->K + STREAM_TILE - 1) / STREAM_TILE;      /* tiles of the padded obstacle array */
    A.iterLog = c->iterLog;
    A.W = c->p.width; A.H = c->p.height; A.L = c->p.agent_length; A.R1Size = c->R1Size; A.R2Size = c->R2Size;
    A.goalX = c->goal[0]; A.goalY = c->goal[1]; A.goalR = c->p.goal_threshold;
    A.N = c->p.N; A.n = c->p.n; A.c1 = c->c1; A.numDisc = c->p.num_disc; A.maxTree = c->p.max_tree_size;
    A.numIterations = c->p.num_iterations; A.useHist = c->useHist;
    A.seed = c->p.seed;
    A.car.aScale = (float)(c->p.accel_max - c->p.accel_min); A.car.aLo = (float)c->p.accel_min;
    A.car.sScale = c->p.steer_max - c->p.steer_min;          A.car.sLo = c->p.steer_min;
    A.car.dScale = (float)(c->p.duration_max - c->p.duration_min); A.car.dLo = (float)c->p.duration_min;
    return A;
}

/* choose the collision back end + shared memory, set kernel attributes, size the persistent grid */
static int configure(kgmt_ctx* ctx) {
    const size_t histBytes = ctx->useHist ? (((size_t)2 * ctx->c1 * 4 + 15) & ~(size_t)15) : 0;
    const size_t limit = ctx->p.reserved[0] > 0 ? (size_t)ctx->p.reserved[0] : (size_t)66 * 1024;   /* staging budget: 3 CTAs per SM next to ~7 KB of static shared memory */
    size_t colBytes = 0;
    int col;
    if (ctx->p.collision_mode == KGMT_COLLIDE_BRUTE) {
        /* every obstacle, every step: one shared-memory piece while it fits the staging budget, else streamed through
         * two 16 KB tiles by the TMA engine (occupancy stays at 4 CTAs per SM instead of 1) */
        colBytes = (size_t)ctx->K * 16;
        if (colBytes <= limit) col = COL_BRUTE_SMEM;
        else { col = COL_BRUTE_STREAM; colBytes = (size_t)2 * STREAM_TILE * 16; }
    } else {
        colBytes = (size_t)ctx->cellStartInts * 4 + (size_t)ctx->numItems * 16;
        col = (colBytes <= limit && histBytes + colBytes <= MAX_DYN_SMEM) ? COL_GRID_SMEM : COL_GRID_GLOBAL;
    }
    if (col == COL_GRID_GLOBAL || col == COL_BRUTE_GLOBAL) colBytes = 0;
    ctx->col = col;
    ctx->smemBytes = histBytes + colBytes;
    /* a new obstacle set of the same size class resolves to the same kernels and shared-memory size: keep the launch
     * configuration (the attribute / occupancy queries below are ~100 us of driver calls per kgmt_set_obstacles) */
    if (ctx->configured && ctx->cfgCol == col && ctx->cfgSmem == ctx->smemBytes) return KGMT_OK;
    ctx->cfgCol = col; ctx->cfgSmem = ctx->smemBytes;
    int occ = 1 << 30;
    for (int rec = 0; rec < 2; ++rec) {
        expand_fn f = expand_entry(col, rec != 0);
        CU(cudaFuncSetAttribute((const void*)f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smemBytes));
        int o = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, (const void*)f, TILE, ctx->smemBytes));
        occ = std::min(occ, o);
    }
    CU(cudaFuncSetAttribute((const void*)propagate_entry(col), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)colBytes));
    {
        const void* sf = (const void*)shard_entry(col);
        CU(cudaFuncSetAttribute(sf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smemBytes));
        int o = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, sf, TILE, ctx->smemBytes));
        ctx->shardGrid = std::max(o, 1) * ctx->numSMs;
    }
    {
        const void* ff = (const void*)fused_entry(col);
        const size_t fsm = (col == COL_BRUTE_STREAM) ? histBytes : ctx->smemBytes;
        CU(cudaFuncSetAttribute(ff, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm));
        int o = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, ff, TILE, fsm));
        ctx->fusedGrid = std::max(o, 1) * ctx->numSMs;
    }
    if (occ < 1) return fail(ctx, KGMT_ERR_CUDA, "expand kernel does not fit on an SM (smem %zu B)", ctx->smemBytes);
    ctx->gridMax = occ * ctx->numSMs;
    /* persistent grid of the cooperative kernel: resident CTAs per SM (tuning knob: reserved[1], or the environment
     * variable KGMT_CTAS_PER_SM for experiments; 0 = all that fit) */
    int perSM = ctx->p.reserved[1];
    if (perSM <= 0) { const char* e = getenv("KGMT_CTAS_PER_SM"); if (e) perSM = atoi(e); }
    if (perSM <= 0) {
        /* no more CTAs than the largest iteration can feed with one chunk per warp: every resident CTA takes part in
         * the grid barrier of every iteration, and small trees (config 1: 30 000 nodes) are barrier-latency bound
         * (median time-to-first-solution 0.32 ms with 4 CTAs per SM, 0.27 ms with 1) */
        const long long chunks = ((long long)ctx->maxCand + CHUNK - 1) / CHUNK;
        const long long ctas = (chunks + WARPS - 1) / WARPS;
        perSM = (int)std::max<long long>(1, std::min<long long>(occ, (ctas + ctx->numSMs - 1) / ctx->numSMs));
    }
    ctx->gridLoop = std::min(perSM, occ) * ctx->numSMs;
    ctx->configured = true;
    return KGMT_OK;
}

/* uniform grid over the workspace: CSR of obstacle AABBs per cell, built ON THE DEVICE from ctx->dObs (cull_* kernels);
 * the host learns one number, the item count, which sizes the shared-memory staging of the planner kernels */
/* cell of one coordinate as cull_cell_of computes it on the device (round-down conversion saturates, NaN -> 0) */
static inline int host_cull_cell(float v, float inv, int C) {
    const float t = v * inv;
    if (!(t > 0.0f)) return 0;
    if (t >= (float)C) return C - 1;
    return std::min((int)floorf(t), C - 1);
}

static int build_cull_grid(kgmt_ctx* ctx, const float* h_aabb = nullptr) {
    const int K = ctx->K;
    int C = ctx->p.cull_cells;
    if (C <= 0) { const char* e = getenv("KGMT_CULL_CELLS"); if (e) C = atoi(e); }      /* experiments (scripts/ab_variants.sh) */
    if (C <= 0) C = (int)std::ceil(1.5 * std::sqrt((double)std::max(K, 1)));   /* measured on config 2: 32 -> 48 cells per side = -2.8 % plan time */
    C = std::max(1, std::min(C, 512));
    const float invX = (float)C / ctx->p.width, invY = (float)C / ctx->p.height;
    const int cells = C * C;
    const int startInts = (cells + 1 + 3) & ~3;
    cudaStream_t s = ctx->stream;
    if ((size_t)startInts > ctx->cellStartCap) {
        if (ctx->dCellStart) cudaFree(ctx->dCellStart);
        ctx->dCellStart = nullptr; ctx->cellStartCap = 0;
        CU(cudaMalloc(&ctx->dCellStart, (size_t)startInts * 4));
        ctx->cellStartCap = startInts;
    }
    if (!ctx->hCullTotal) CU(cudaHostAlloc(&ctx->hCullTotal, 16, cudaHostAllocDefault));
    if (!ctx->dCullTotal) CU(cudaMalloc(&ctx->dCullTotal, 16));
    CU(cudaMemsetAsync(ctx->dCellStart, 0, (size_t)startInts * 4, s));
    cull_count_kernel<<<(cells * 32 + 255) / 256, 256, 0, s>>>(ctx->dObs, K, C, invX, invY, ctx->dCellStart);      /* a warp per cell */
    cull_scan_kernel<<<1, 1024, 0, s>>>(ctx->dCellStart, cells, startInts, ctx->dCullTotal);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(ctx->hCullTotal, ctx->dCullTotal, 4, cudaMemcpyDeviceToHost, s));
    int realItems;
    if (h_aabb) {
        /* the caller's obstacles are on the host: the item count (all that sizes the shared-memory staging) is a sum over
         * the obstacles of the cells they span — no need to wait for the device's scan; the device's own total is checked
         * against it at the next synchronisation (fetch_state) */
        long long sum = 0;
        for (int k = 0; k < K; ++k) {
            const float* o = h_aabb + (size_t)k * 4;
            const long long nx = (long long)host_cull_cell(o[2], invX, C) - host_cull_cell(o[0], invX, C) + 1;
            const long long ny = (long long)host_cull_cell(o[3], invY, C) - host_cull_cell(o[1], invY, C) + 1;
            if (nx > 0 && ny > 0) sum += nx * ny;
        }
        if (sum > 0x7fffff00LL) return fail(ctx, KGMT_ERR_INVALID, "cull grid too large (%lld items)", sum);
        realItems = (int)sum;
        ctx->cullExpect = realItems;
    } else {
        CU(cudaStreamSynchronize(s));
        realItems = *ctx->hCullTotal;
        ctx->cullExpect = -1;
    }
    const int numItems = realItems + 3;          /* + three boxes nothing overlaps: the cell walk reads up to four entries per trip */
    if ((size_t)numItems > ctx->cellItemsCap) {
        if (ctx->dCellItems) cudaFree(ctx->dCellItems);
        ctx->dCellItems = nullptr; ctx->cellItemsCap = 0;
        const size_t cap = (size_t)numItems + (size_t)numItems / 4 + 64;
        CU(cudaMalloc(&ctx->dCellItems, cap * 16));
        ctx->cellItemsCap = cap;
    }
    cull_fill_kernel<<<(cells * 32 + 255) / 256, 256, 0, s>>>(ctx->dObs, K, C, invX, invY, ctx->dCellStart, ctx->dCellItems, realItems);
    CU(cudaGetLastError());
    ctx->launches += 3;
    ctx->cullC = C; ctx->cullInvX = invX; ctx->cullInvY = invY; ctx->cellStartInts = startInts; ctx->numItems = numItems;
    return KGMT_OK;
}

/* obstacles are on the device in ctx->dObs[0, K) (copied there by the caller): pad to whole stream tiles, build the grid */
static int install_obstacles(kgmt_ctx* ctx, const float* h_aabb = nullptr) {
    const int K = ctx->K;
    const int padded = (int)ctx->obsCap;
    if (padded > K) cull_pad_kernel<<<(padded - K + 255) / 256, 256, 0, ctx->stream>>>(ctx->dObs, K, padded);
    CU(cudaGetLastError());
    ctx->launches += 1;
    int rc = build_cull_grid(ctx, h_aabb);
    if (rc) return rc;
    return configure(ctx);
}

/* room for K obstacles padded to whole stream tiles */
static int ensure_obstacle_storage(kgmt_ctx* ctx, int K) {
    const size_t padded = (size_t)std::max((K + STREAM_TILE - 1) / STREAM_TILE, 1) * STREAM_TILE;
    if (padded != ctx->obsCap) {
        if (ctx->dObs) cudaFree(ctx->dObs);
        ctx->dObs = nullptr; ctx->obsCap = 0;
        CU(cudaMalloc(&ctx->dObs, padded * 16));
        ctx->obsCap = padded;
    }
    return KGMT_OK;
}

static int ensure_scratch(kgmt_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->scratchBytes) return KGMT_OK;
    if (ctx->scratch) cudaFree(ctx->scratch);
    ctx->scratch = nullptr; ctx->scratchBytes = 0;
    CU(cudaMalloc(&ctx->scratch, bytes));
    ctx->scratchBytes = bytes;
    return KGMT_OK;
}

static int ensure_record(kgmt_ctx* ctx) {
    if (ctx->recordAllocated) return KGMT_OK;
    const size_t M = (size_t)ctx->maxCand;
    CU(cudaMalloc(&ctx->candState, M * 16));
    CU(cudaMalloc(&ctx->candCtrl, M * 16));
    CU(cudaMalloc(&ctx->candParent, M * 4));
    CU(cudaMalloc(&ctx->candR1, M * 4));
    CU(cudaMalloc(&ctx->candR2, M * 4));
    CU(cudaMalloc(&ctx->candFlags, M));
    ctx->recordAllocated = true;
    CU(cudaMemsetAsync(ctx->candState, 0, M * 16, ctx->stream));
    CU(cudaMemsetAsync(ctx->candCtrl, 0, M * 16, ctx->stream));
    CU(cudaMemsetAsync(ctx->candParent, 0xFF, M * 4, ctx->stream));
    CU(cudaMemsetAsync(ctx->candR1, 0xFF, M * 4, ctx->stream));
    CU(cudaMemsetAsync(ctx->candR2, 0xFF, M * 4, ctx->stream));
    CU(cudaMemsetAsync(ctx->candFlags, 0, M, ctx->stream));
    return KGMT_OK;
}

/* state as left by the reference constructor (KGMT.cu:16-40,70-72): maps zero, scores 1.0, no node, no candidate.
 * Tree rows are NOT rewritten: rows at or above treeSize are dead, and kgmt_export reports them as the constructor
 * leaves them (zeros, parent -1).  light = the re-plan inside kgmt_plan: one memset of the slab; begin_kernel rebuilds the
 * scalar block and both score buffers are rewritten before they are read. */
static int clear_state(kgmt_ctx* ctx, bool sync, bool light = false) {
    CU(cudaMemsetAsync(ctx->mapSlab, 0, ctx->slabInts * 4, ctx->stream));
    if (ctx->recordAllocated) {
        const size_t M = std::min((size_t)ctx->maxCand, ctx->dirtyCand);
        if (M) {
            CU(cudaMemsetAsync(ctx->candState, 0, M * 16, ctx->stream));
            CU(cudaMemsetAsync(ctx->candCtrl, 0, M * 16, ctx->stream));
            CU(cudaMemsetAsync(ctx->candParent, 0xFF, M * 4, ctx->stream));
            CU(cudaMemsetAsync(ctx->candR1, 0xFF, M * 4, ctx->stream));
            CU(cudaMemsetAsync(ctx->candR2, 0xFF, M * 4, ctx->stream));
            CU(cudaMemsetAsync(ctx->candFlags, 0, M, ctx->stream));
        }
    }
    ctx->dirtyCand = 0;
    ctx->begun = false;
    ctx->haveCkpt = false;
    ctx->pathPreLaunches = -1;
    if (light) return KGMT_OK;
    fill_float_kernel<<<(2 * ctx->c1 + 255) / 256, 256, 0, ctx->stream>>>(ctx->R1Score[0], 1.0f, (size_t)2 * ctx->c1);
    DevState z{};
    z.goalIdx = -1; z.goalSlot = -1; z.goalBest[0] = z.goalBest[1] = ~0ull; z.stop = STOP_ITER_LIMIT; z.itr = 1;
    z.forceChildren = ctx->hState->forceChildren;
    ctx->resetState = z;
    CU(cudaMemcpyAsync(ctx->dState, &ctx->resetState, sizeof(DevState), cudaMemcpyHostToDevice, ctx->stream));
    if (sync) {
        CU(cudaStreamSynchronize(ctx->stream));
        *ctx->hState = z;
    }
    return KGMT_OK;
}

/* scan/ticket bookkeeping back to "between two iterations, nothing in flight" (after a restore or a sharded
 * round, which do not run the planner loops' own recycling) */
static int reset_loop_bookkeeping(kgmt_ctx* ctx) {
    const unsigned t0 = (unsigned)(ctx->gridLoop * WARPS);
    const unsigned tk[4] = {t0, t0, t0, 0u};
    CU(cudaMemcpyAsync(ctx->ticket, tk, 16, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(ctx->blockSum, 0, 3 * ctx->blocksCap * 4, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return KGMT_OK;
}

/* after the scalars have landed in *hState */
static int fetch_state_landed(kgmt_ctx* ctx) {
    if (ctx->cullExpect >= 0) {                /* the device's scan of the cull grid against the host's count (build_cull_grid) */
        const int expect = ctx->cullExpect;
        ctx->cullExpect = -1;
        if (*ctx->hCullTotal != expect)
            return fail(ctx, KGMT_ERR_STATE, "cull grid: device counted %d items, host %d", *ctx->hCullTotal, expect);
    }
    if (ctx->hState->iterationsDone > 0) ctx->dirtyCand = ctx->maxCand;
    return KGMT_OK;
}

static int fetch_state(kgmt_ctx* ctx) {
    CU(cudaMemcpyAsync(ctx->hState, ctx->dState, sizeof(DevState), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return fetch_state_landed(ctx);
}

static void fill_result(const kgmt_ctx* ctx, kgmt_result* out, float ms) {
    const DevState& s = *ctx->hState;
    out->stop = s.stop; out->iterations = s.iterationsDone; out->tree_size = s.treeSize;
    out->cost_to_goal = s.costToGoal; out->goal_index = s.goalIdx; out->expansions = s.expansions;
    out->device_ms = ms; out->kernel_launches = ctx->planLaunches; out->done_ms = ms; out->service_ms = ms;
}

typedef void (*batch_fn)(const BatchArgs);
static batch_fn batch_entry(int col) {
    switch (col) {
        case COL_GRID_SMEM: return batch_kernel<COL_GRID_SMEM>;
        case COL_GRID_GLOBAL: return batch_kernel<COL_GRID_GLOBAL>;
        case COL_BRUTE_SMEM: return batch_kernel<COL_BRUTE_SMEM>;
        default: return batch_kernel<COL_BRUTE_GLOBAL>;
    }
}
static void free_batch(kgmt_ctx* ctx) {
    kgmt_ctx::Batch& b = ctx->batch;
    cudaFree(b.treeState); cudaFree(b.treeCtrl); cudaFree(b.stageState); cudaFree(b.stageCtrl); cudaFree(b.treeParent);
    cudaFree(b.mapSlab); cudaFree(b.blockSum); cudaFree(b.queryTicket); cudaFree(b.wsQuery); cudaFree(b.pathLen);
    cudaFree(b.chunkMask); cudaFree(b.ticket); cudaFree(b.initState); cudaFree(b.initCtrl); cudaFree(b.goalXY);
    cudaFree(b.seeds); cudaFree(b.states); cudaFree(b.paths); cudaFree(b.launchT0);
    b = kgmt_ctx::Batch();
}


/* ================================================================================ ABI == */
extern "C" {

int kgmt_abi_version(void) { return KGMT_ABI_VERSION; }

void kgmt_default_params(kgmt_params* p) {
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->width = 20.0f; p->height = 20.0f; p->N = 16; p->n = 8; p->num_iterations = 100; p->max_tree_size = 30000;
    p->num_disc = 10; p->agent_length = 1.0f; p->goal_threshold = 0.5f;
    p->seed = 1u; p->device = -1; p->max_candidates = 0; p->collision_mode = KGMT_COLLIDE_GRID;
    p->record_candidates = 0; p->cull_cells = 0;
    /* statePropagator.cu:17-19 */
    p->accel_min = -5.0; p->accel_max = 5.0;
    p->steer_min = -3.14159265358979323846; p->steer_max = 3.14159265358979323846;
    p->duration_min = (double)0.05f; p->duration_max = (double)0.05f + 1.0;
}

/* ---- car model file (the reference ships systems/car.yaml EMPTY and hard-codes the model) ---- */
int kgmt_params_from_yaml(const char* path, kgmt_params* p, int* bad_line) {
    if (bad_line) *bad_line = 0;
    if (!path || !p) return KGMT_ERR_INVALID;
    FILE* f = fopen(path, "r");
    if (!f) return KGMT_ERR_INVALID;
    char line[512];
    int ln = 0, rc = KGMT_OK;
    while (fgets(line, sizeof(line), f)) {
        ++ln;
        char* hash = strchr(line, '#');
        if (hash) *hash = 0;
        char* s = line;
        while (*s == ' ' || *s == '\t') ++s;
        char* e = s + strlen(s);
        while (e > s && (e[-1] == '\n' || e[-1] == '\r' || e[-1] == ' ' || e[-1] == '\t')) *--e = 0;
        if (!*s || !strcmp(s, "---") || !strcmp(s, "...")) continue;
        char* colon = strchr(s, ':');
        if (!colon) { rc = KGMT_ERR_INVALID; break; }
        *colon = 0;
        char* key = s;
        char* ke = colon;
        while (ke > key && (ke[-1] == ' ' || ke[-1] == '\t')) *--ke = 0;
        char* val = colon + 1;
        while (*val == ' ' || *val == '\t') ++val;
        if (!*val) continue;                                   /* "controls:" — a block header, flattened */
        char* endp = nullptr;
        double v = strtod(val, &endp);
        bool num = endp != val;
        if (num) { while (*endp == ' ') ++endp; num = (*endp == 0); }
        if (!num) {
            /* the one non-numeric spelling worth having: bounds in units of pi ("pi", "-pi", "0.5pi", "2*pi") */
            const char* q = val;
            double k = 1.0;
            char* e2 = nullptr;
            const double kk = strtod(q, &e2);
            if (e2 != q) { k = kk; q = e2; }
            else if (*q == '-') { k = -1.0; ++q; }
            else if (*q == '+') { ++q; }
            while (*q == ' ' || *q == '*') ++q;
            if (!strcmp(q, "pi") || !strcmp(q, "PI") || !strcmp(q, "M_PI")) { num = true; v = k * 3.14159265358979323846; }
        }
        if (!num) { rc = KGMT_ERR_INVALID; break; }
        if (!strcmp(key, "wheelbase") || !strcmp(key, "agent_length") || !strcmp(key, "length")) p->agent_length = (float)v;
        else if (!strcmp(key, "num_disc") || !strcmp(key, "numDisc")) p->num_disc = (int)v;
        else if (!strcmp(key, "accel_min")) p->accel_min = v;
        else if (!strcmp(key, "accel_max")) p->accel_max = v;
        else if (!strcmp(key, "steer_min") || !strcmp(key, "steering_min")) p->steer_min = v;
        else if (!strcmp(key, "steer_max") || !strcmp(key, "steering_max")) p->steer_max = v;
        else if (!strcmp(key, "duration_min")) p->duration_min = v;
        else if (!strcmp(key, "duration_max")) p->duration_max = v;
        else if (!strcmp(key, "width")) p->width = (float)v;
        else if (!strcmp(key, "height")) p->height = (float)v;
        else if (!strcmp(key, "N")) p->N = (int)v;
        else if (!strcmp(key, "n")) p->n = (int)v;
        else if (!strcmp(key, "num_iterations")) p->num_iterations = (int)v;
        else if (!strcmp(key, "max_tree_size")) p->max_tree_size = (int)v;
        else if (!strcmp(key, "goal_threshold")) p->goal_threshold = (float)v;
        else if (!strcmp(key, "seed")) p->seed = (uint32_t)v;
        else { rc = KGMT_ERR_INVALID; break; }
    }
    fclose(f);
    if (rc && bad_line) *bad_line = ln;
    return rc;
}

const char* kgmt_last_error(const kgmt_ctx* ctx) { return ctx ? ctx->err : "null context"; }

void kgmt_destroy(kgmt_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    kgmt_comm_destroy(ctx);
    kgmt_peer_detach(ctx);
    cudaFree(ctx->treeState); cudaFree(ctx->treeCtrl); cudaFree(ctx->treeParent);
    cudaFree(ctx->mapSlab); cudaFree(ctx->mapSlabCkpt);
    cudaFree(ctx->candState); cudaFree(ctx->candCtrl); cudaFree(ctx->candParent);
    cudaFree(ctx->candR1); cudaFree(ctx->candR2); cudaFree(ctx->candFlags);
    cudaFree(ctx->chunkMask); cudaFree(ctx->ticket);
    cudaFree(ctx->stageState); cudaFree(ctx->stageCtrl);
    cudaFree(ctx->dState); cudaFree(ctx->iterLog);
    if (ctx->hState) cudaFreeHost(ctx->hState);
    cudaFree(ctx->dObs); cudaFree(ctx->dCellStart); cudaFree(ctx->dCellItems); cudaFree(ctx->dCullTotal);
    if (ctx->hObsPinned) cudaFreeHost(ctx->hObsPinned);
    if (ctx->hCullTotal) cudaFreeHost(ctx->hCullTotal);
    cudaFree(ctx->scratch); cudaFree(ctx->dParents); cudaFree(ctx->dPathPre); if (ctx->hPathPre) cudaFreeHost(ctx->hPathPre);
    free_batch(ctx);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->evState) cudaEventDestroy(ctx->evState);
    if (ctx->evPath) cudaEventDestroy(ctx->evPath);
    if (ctx->ownStream) cudaStreamDestroy(ctx->ownStream);
    cudaFree(ctx->peer.block); cudaFree(ctx->peer.plan); cudaFree(ctx->peer.dRaceFlags);
    if (ctx->peer.hPlan) cudaFreeHost(ctx->peer.hPlan);
    cudaFree(ctx->shardPrefix); cudaFree(ctx->shardTotal);
    if (ctx->hShardTotal) cudaFreeHost(ctx->hShardTotal);
    delete ctx;
}

int kgmt_create(const kgmt_params* p, kgmt_ctx** out) {
    if (!p || !out) return KGMT_ERR_INVALID;
    *out = nullptr;
    if (!(p->width > 0.f) || !(p->height > 0.f) || p->N < 1 || p->n < 1 || p->max_tree_size < 1 || p->num_disc < 1 ||
        !(p->agent_length != 0.f) || p->N > 1024 || p->n > 1024)
        return KGMT_ERR_INVALID;
    if ((long long)p->N * p->N * p->n * p->n > (1LL << 30)) return KGMT_ERR_INVALID;
    kgmt_ctx* ctx = new (std::nothrow) kgmt_ctx();
    if (!ctx) return KGMT_ERR_NOMEM;
    *out = ctx;                                   /* returned even on failure so the caller can read the error */
    ctx->p = *p;
    {   /* car model: all-zero ranges (a caller that filled the struct by hand) mean the reference's literals */
        kgmt_params& q = ctx->p;
        if (q.accel_min == 0.0 && q.accel_max == 0.0 && q.steer_min == 0.0 && q.steer_max == 0.0 &&
            q.duration_min == 0.0 && q.duration_max == 0.0) {
            kgmt_params d; kgmt_default_params(&d);
            q.accel_min = d.accel_min; q.accel_max = d.accel_max; q.steer_min = d.steer_min; q.steer_max = d.steer_max;
            q.duration_min = d.duration_min; q.duration_max = d.duration_max;
        }
        const double r[6] = {q.accel_min, q.accel_max, q.steer_min, q.steer_max, q.duration_min, q.duration_max};
        for (double v : r) if (!std::isfinite(v)) return fail(ctx, KGMT_ERR_INVALID, "control range is not finite");
        if (q.accel_max < q.accel_min || q.steer_max < q.steer_min || q.duration_max < q.duration_min || q.duration_min < 0.0)
            return fail(ctx, KGMT_ERR_INVALID, "control range: max < min, or negative duration");
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(ctx, KGMT_ERR_CUDA, "no CUDA device (%s): this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (p->device >= 0) { ctx->device = p->device; } else { CU(cudaGetDevice(&ctx->device)); }
    CU(cudaSetDevice(ctx->device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, ctx->device));
    if (prop.major < 10) return fail(ctx, KGMT_ERR_CUDA, "device %s is sm_%d%d; this library is built for sm_100a only",
                                     prop.name, prop.major, prop.minor);
    ctx->numSMs = prop.multiProcessorCount;
    CU(cudaStreamCreateWithFlags(&ctx->ownStream, cudaStreamNonBlocking));
    ctx->stream = ctx->ownStream;
    CU(cudaEventCreate(&ctx->ev0));
    CU(cudaEventCreate(&ctx->ev1));
    CU(cudaEventCreateWithFlags(&ctx->evState, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&ctx->evPath, cudaEventDisableTiming));

    const int N = p->N, n = p->n;
    ctx->c1 = N * N;
    ctx->c2 = (size_t)ctx->c1 * n * n;
    ctx->R1Size = p->width / N;                   /* KGMT.cu:13 */
    ctx->R2Size = p->width / (n * N);             /* KGMT.cu:14 */
    ctx->maxCand = p->max_candidates > 0 ? p->max_candidates : p->max_tree_size;
    ctx->useHist = ((size_t)2 * ctx->c1 * 4 <= 32 * 1024) ? 1 : 0;

    const size_t T = (size_t)p->max_tree_size;
    CU(cudaMalloc(&ctx->treeState, T * 16));
    CU(cudaMalloc(&ctx->treeCtrl, T * 16));
    CU(cudaMalloc(&ctx->treeParent, T * 4));
    const size_t c1 = (size_t)ctx->c1, c2 = ctx->c2;
    ctx->mapSlabInts = 7 * c1 + 4 * c2;
    /* whole scan blocks: staged_parent reads the 256 ballots of a block with 128-bit loads */
    ctx->chunksCap = ((((size_t)ctx->maxCand + CHUNK - 1) / CHUNK + 1) + BLK_CHUNKS - 1) / BLK_CHUNKS * BLK_CHUNKS;
    ctx->blocksCap = (ctx->chunksCap + BLK_CHUNKS - 1) / BLK_CHUNKS + 1;
    /* one slab: region maps, then the per-iteration scan block sums — a re-plan clears all of it with ONE memset */
    const size_t bookInts = 3 * ctx->blocksCap;
    ctx->slabInts = ((ctx->mapSlabInts + 3) & ~(size_t)3) + bookInts;
    CU(cudaMalloc(&ctx->mapSlab, ctx->slabInts * 4));
    int* m = ctx->mapSlab;
    ctx->R1 = m; m += c1; ctx->R1Valid = m; m += c1; ctx->R1Invalid = m; m += c1; ctx->R1Avail = m; m += c1;
    ctx->R1Cov = m; m += c1; ctx->R1Score[0] = reinterpret_cast<float*>(m); m += c1;
    ctx->R1Score[1] = reinterpret_cast<float*>(m); m += c1;
    ctx->R2 = m; m += c2; ctx->R2Valid = m; m += c2; ctx->R2Invalid = m; m += c2;
    ctx->R2Stamp = reinterpret_cast<unsigned*>(m);
    m = ctx->mapSlab + ((ctx->mapSlabInts + 3) & ~(size_t)3);
    ctx->blockSum = m;
    CU(cudaMalloc(&ctx->chunkMask, 2 * ctx->chunksCap * 4));
    CU(cudaMalloc(&ctx->ticket, 4 * 4));
    CU(cudaMalloc(&ctx->stageState, 2 * (size_t)ctx->maxCand * 16));
    CU(cudaMalloc(&ctx->stageCtrl, 2 * (size_t)ctx->maxCand * 16));
    CU(cudaMemsetAsync(ctx->chunkMask, 0, 2 * ctx->chunksCap * 4, ctx->stream));
    CU(cudaMemsetAsync(ctx->ticket, 0, 16, ctx->stream));
    CU(cudaMalloc(&ctx->dState, sizeof(DevState)));
    CU(cudaHostAlloc(&ctx->hState, sizeof(DevState), cudaHostAllocDefault));
    memset(ctx->hState, 0, sizeof(DevState));
    if (p->record_candidates) { int rc = ensure_record(ctx); if (rc) return rc; }
    ctx->dirtyCand = ctx->maxCand;
    int rc = clear_state(ctx, true);
    if (rc) return rc;
    ctx->K = 0;
    rc = ensure_obstacle_storage(ctx, 0);
    if (rc) return rc;
    return install_obstacles(ctx);
}

int kgmt_reset(kgmt_ctx* ctx) {
    if (!ctx) return KGMT_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    return clear_state(ctx, true);
}

/* the Philox key of the next plan (the reference re-seeds every plan() from time(NULL), KGMT.cu:111) */
int kgmt_set_seed(kgmt_ctx* ctx, uint32_t seed) {
    if (!ctx) return KGMT_ERR_INVALID;
    ctx->p.seed = seed;
    return KGMT_OK;
}

int kgmt_set_obstacles_host(kgmt_ctx* ctx, const float* h_aabb, int K) {
    if (!ctx || K < 0 || (K > 0 && !h_aabb)) return fail(ctx, KGMT_ERR_INVALID, "bad obstacle array");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));            /* nothing in flight reads the old set or the staging buffer */
    int rc = ensure_obstacle_storage(ctx, K);
    if (rc) return rc;
    if (K > 0) {
        if ((size_t)K * 4 > ctx->hObsPinnedCap) {
            if (ctx->hObsPinned) cudaFreeHost(ctx->hObsPinned);
            ctx->hObsPinned = nullptr; ctx->hObsPinnedCap = 0;
            CU(cudaHostAlloc(&ctx->hObsPinned, (size_t)K * 16 + 4096, cudaHostAllocDefault));
            ctx->hObsPinnedCap = (size_t)K * 4 + 1024;
        }
        memcpy(ctx->hObsPinned, h_aabb, (size_t)K * 16);
        CU(cudaMemcpyAsync(ctx->dObs, ctx->hObsPinned, (size_t)K * 16, cudaMemcpyHostToDevice, ctx->stream));
    }
    ctx->K = K;
    return install_obstacles(ctx, K > 0 ? h_aabb : nullptr);
}

int kgmt_set_obstacles(kgmt_ctx* ctx, const float* d_aabb, int K) {
    if (!ctx || K < 0 || (K > 0 && !d_aabb)) return fail(ctx, KGMT_ERR_INVALID, "bad obstacle array");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    int rc = ensure_obstacle_storage(ctx, K);
    if (rc) return rc;
    /* device to device on the planner's stream: the caller's array is read once, here (KGMT::plan's d_obstacles, main.cu:60-62) */
    if (K) CU(cudaMemcpyAsync(ctx->dObs, d_aabb, (size_t)K * 16, cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->K = K;
    return install_obstacles(ctx);
}

int kgmt_begin(kgmt_ctx* ctx, const float* initial7, const float* goal7) {
    if (!ctx || !initial7 || !goal7) return fail(ctx, KGMT_ERR_INVALID, "null initial/goal");
    CU(cudaSetDevice(ctx->device));
    if (ctx->begun) { int rc = clear_state(ctx, true); if (rc) return rc; }
    memcpy(ctx->goal, goal7, sizeof(ctx->goal));
    const KArgs A = make_args(ctx);
    begin_kernel<<<1, TILE, 0, ctx->stream>>>(A, make_float4(initial7[0], initial7[1], initial7[2], initial7[3]),
                                             make_float4(initial7[4], initial7[5], initial7[6], 0.f), ctx->hState->forceChildren);
    CU(cudaGetLastError());
    ctx->launches += 1;
    ctx->planLaunches = 1;
    ctx->begun = true;
    return fetch_state(ctx);
}

static int launch_expand(kgmt_ctx* ctx, int maxIters) {
    KArgs A = make_args(ctx);
    expand_fn f = expand_entry(ctx->col, ctx->p.record_candidates != 0);
    void* args[] = {(void*)&A, (void*)&maxIters};
    CU(cudaLaunchCooperativeKernel((const void*)f, dim3(ctx->gridLoop), dim3(TILE), args, ctx->smemBytes, ctx->stream));
    ctx->launches += 1;
    ctx->planLaunches += 1;
    return KGMT_OK;
}

int kgmt_expand_iteration(kgmt_ctx* ctx, kgmt_iter_stats* out) { return kgmt_expand_iterations(ctx, 1, out); }

/* up to `count` while-loop bodies (KGMT.cu:118-259) in ONE cooperative launch; stops early when the planner stops.
 * `out` describes the last iteration executed. */
int kgmt_expand_iterations(kgmt_ctx* ctx, int count, kgmt_iter_stats* out) {
    if (!ctx || count < 1) return KGMT_ERR_INVALID;
    if (!ctx->begun) return fail(ctx, KGMT_ERR_STATE, "kgmt_expand_iteration(s) before kgmt_begin / kgmt_seed_frontier");
    CU(cudaSetDevice(ctx->device));
    if (ctx->hState->stop == STOP_RUNNING) {
        int rc = launch_expand(ctx, count);
        if (rc) return rc;
        rc = fetch_state(ctx);
        if (rc) return rc;
    }
    if (out) {
        const DevState& s = *ctx->hState;
        out->iteration = s.lastItr; out->mode = s.lastMode; out->children = s.lastChildren; out->frontier = s.lastFrontier;
        out->candidates = s.lastM; out->accepted = s.lastAccepted; out->tree_size = s.treeSize; out->stop = s.stop;
        out->cost_to_goal = s.costToGoal; out->goal_index = s.goalIdx;
    }
    return KGMT_OK;
}

int kgmt_get_result(kgmt_ctx* ctx, kgmt_result* out) {
    if (!ctx || !out) return KGMT_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    int rc = fetch_state(ctx);
    if (rc) return rc;
    fill_result(ctx, out, 0.f);
    return KGMT_OK;
}

/* KGMT::plan: root insertion + the whole expansion loop in ONE cooperative persistent launch */
int kgmt_plan(kgmt_ctx* ctx, const float* initial7, const float* goal7, kgmt_result* out) {
    if (!ctx || !initial7 || !goal7) return fail(ctx, KGMT_ERR_INVALID, "null initial/goal");
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventRecord(ctx->ev0, ctx->stream));
    if (ctx->begun) { int rc = clear_state(ctx, false, true); if (rc) return rc; }   /* re-plan: inside the timed region */
    memcpy(ctx->goal, goal7, sizeof(ctx->goal));
    KArgs A = make_args(ctx);
    begin_kernel<<<1, TILE, 0, ctx->stream>>>(A, make_float4(initial7[0], initial7[1], initial7[2], initial7[3]),
                                             make_float4(initial7[4], initial7[5], initial7[6], 0.f), ctx->hState->forceChildren);
    CU(cudaGetLastError());
    ctx->launches += 1;
    ctx->planLaunches = 1;
    { int rc = launch_expand(ctx, 0x7fffffff); if (rc) return rc; }
    CU(cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->begun = true;
    /* the solution's back-trace rides along: one more small kernel behind the planner (outside the event pair), its rows
     * in the same synchronisation as the scalars — kgmt_extract_path of the goal node is then a host copy */
    if (!ctx->dPathPre) {
        CU(cudaMalloc(&ctx->dPathPre, PATH_PRE_BYTES));
        CU(cudaHostAlloc(&ctx->hPathPre, PATH_PRE_BYTES, cudaHostAllocDefault));
    }
    /* order on the stream: scalars -> [evState] -> back-trace kernel -> path rows -> [evPath].  kgmt_plan returns as soon as
     * the scalars have landed; the back-trace finishes behind the caller's back and kgmt_extract_path waits on evPath */
    CU(cudaMemcpyAsync(ctx->hState, ctx->dState, sizeof(DevState), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaEventRecord(ctx->evState, ctx->stream));
    trace_goal_kernel<<<1, 32, 0, ctx->stream>>>(A, (int*)ctx->dPathPre, (float*)(ctx->dPathPre + 16), PATH_PRE_ROWS);
    CU(cudaGetLastError());
    ctx->launches += 1;
    CU(cudaMemcpyAsync(ctx->hPathPre, ctx->dPathPre, PATH_PRE_BYTES, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaEventRecord(ctx->evPath, ctx->stream));
    ctx->pathPreLaunches = -1;
    CU(cudaEventSynchronize(ctx->evState));
    int rc = fetch_state_landed(ctx);
    if (rc) return rc;
    ctx->pathPreLaunches = ctx->launches; ctx->pathPreGoal = ctx->hState->goalIdx; ctx->pathPreTree = ctx->hState->treeSize;
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    if (out) fill_result(ctx, out, ms);
    return KGMT_OK;
}

/* ---- batched planning: Q independent queries on the context's map (BASELINE config 4) ------------------------- */
static int max_batch_clusters(kgmt_ctx* ctx, int cluster_size, int* out) {
    batch_fn f = batch_entry(ctx->col);
    CU(cudaFuncSetAttribute((const void*)f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smemBytes));
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster_size; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(TILE); cfg.dynamicSmemBytes = ctx->smemBytes; cfg.stream = ctx->stream; cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.gridDim = dim3(cluster_size * 8);
    CU(cudaOccupancyMaxActiveClusters(out, (const void*)f, &cfg));
    return KGMT_OK;
}

int kgmt_batch_cluster_size(kgmt_ctx* ctx, int Q) {
    if (!ctx || Q < 1) return KGMT_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    for (int cs = 8; cs >= 4; cs >>= 1) {
        int mc = 0;
        int rc = max_batch_clusters(ctx, cs, &mc);
        if (rc) return rc;
        if (mc >= Q) return cs;
    }
    return 2;
}

int kgmt_plan_batch(kgmt_ctx* ctx, const float* h_inits7, const float* h_goals7, const uint32_t* h_seeds, int Q,
                    int cluster_size, kgmt_result* out, float* h_paths7, int max_path, int* h_path_len, float* device_ms) {
    if (!ctx || !h_inits7 || !h_goals7 || !h_seeds || Q < 1) return fail(ctx, KGMT_ERR_INVALID, "bad batch arguments");
    if (cluster_size == 0) { cluster_size = kgmt_batch_cluster_size(ctx, Q); if (cluster_size < 0) return cluster_size; }
    if (cluster_size != 1 && cluster_size != 2 && cluster_size != 4 && cluster_size != 8)
        return fail(ctx, KGMT_ERR_INVALID, "cluster_size must be 0 (choose), 1, 2, 4 or 8");
    if (h_paths7 && (max_path < 1 || !h_path_len)) return fail(ctx, KGMT_ERR_INVALID, "paths need max_path and path_len");
    CU(cudaSetDevice(ctx->device));
    kgmt_ctx::Batch& b = ctx->batch;
    batch_fn f = batch_entry(ctx->col);
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster_size; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(TILE); cfg.dynamicSmemBytes = ctx->smemBytes; cfg.stream = ctx->stream; cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.gridDim = dim3(cluster_size * 8);
    if (b.cfgCluster != cluster_size || b.cfgCol != ctx->col || b.cfgSmem != ctx->smemBytes) {
        CU(cudaFuncSetAttribute((const void*)f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smemBytes));
        int mc = 0;
        CU(cudaOccupancyMaxActiveClusters(&mc, (const void*)f, &cfg));
        b.cfgCluster = cluster_size; b.cfgCol = ctx->col; b.cfgSmem = ctx->smemBytes; b.cfgMaxClusters = mc;
    }
    const int maxClusters = b.cfgMaxClusters;
    if (maxClusters < 1) return fail(ctx, KGMT_ERR_CUDA, "no cluster of %d CTAs fits", cluster_size);
    const int numWs = std::min(maxClusters, Q);
    const size_t T = (size_t)ctx->p.max_tree_size, M = (size_t)ctx->maxCand;
    const size_t mapStride = (7 * (size_t)ctx->c1 + 4 * ctx->c2 + 3) & ~(size_t)3;
    const int wantPath = h_paths7 ? max_path : 0;
    if (numWs > b.numWs || cluster_size != b.clusterSize || Q > b.Qcap || wantPath > b.maxPath) {
        free_batch(ctx);
        const size_t W = (size_t)std::max(numWs, 1);
        CU(cudaMalloc(&b.treeState, W * T * 16)); CU(cudaMalloc(&b.treeCtrl, W * T * 16)); CU(cudaMalloc(&b.treeParent, W * T * 4));
        CU(cudaMalloc(&b.stageState, W * 2 * M * 16)); CU(cudaMalloc(&b.stageCtrl, W * 2 * M * 16));
        CU(cudaMalloc(&b.mapSlab, W * mapStride * 4));
        CU(cudaMalloc(&b.chunkMask, W * 2 * ctx->chunksCap * 4)); CU(cudaMalloc(&b.blockSum, W * 3 * ctx->blocksCap * 4));
        CU(cudaMalloc(&b.ticket, W * 16)); CU(cudaMalloc(&b.queryTicket, 4)); CU(cudaMalloc(&b.wsQuery, W * 4));
        CU(cudaMalloc(&b.initState, (size_t)Q * 16)); CU(cudaMalloc(&b.initCtrl, (size_t)Q * 16));
        CU(cudaMalloc(&b.goalXY, (size_t)Q * 8)); CU(cudaMalloc(&b.seeds, (size_t)Q * 4));
        CU(cudaMalloc(&b.states, (size_t)Q * sizeof(DevState)));
        CU(cudaMalloc(&b.pathLen, (size_t)Q * 4));
        CU(cudaMalloc(&b.launchT0, 8));
        if (wantPath) CU(cudaMalloc(&b.paths, (size_t)Q * wantPath * 28));
        b.numWs = numWs; b.clusterSize = cluster_size; b.Qcap = Q; b.maxPath = wantPath; b.mapIntsStride = mapStride;
    }
    std::vector<float> hs((size_t)Q * 4), hc((size_t)Q * 4), hg((size_t)Q * 2);
    for (int q = 0; q < Q; ++q) {
        memcpy(&hs[(size_t)q * 4], &h_inits7[(size_t)q * 7], 16);
        hc[(size_t)q * 4] = h_inits7[(size_t)q * 7 + 4]; hc[(size_t)q * 4 + 1] = h_inits7[(size_t)q * 7 + 5];
        hc[(size_t)q * 4 + 2] = h_inits7[(size_t)q * 7 + 6]; hc[(size_t)q * 4 + 3] = 0.f;
        hg[(size_t)q * 2] = h_goals7[(size_t)q * 7]; hg[(size_t)q * 2 + 1] = h_goals7[(size_t)q * 7 + 1];
    }
    cudaStream_t s = ctx->stream;
    CU(cudaEventRecord(ctx->ev0, s));
    CU(cudaMemcpyAsync(b.initState, hs.data(), (size_t)Q * 16, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(b.initCtrl, hc.data(), (size_t)Q * 16, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(b.goalXY, hg.data(), (size_t)Q * 8, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(b.seeds, h_seeds, (size_t)Q * 4, cudaMemcpyHostToDevice, s));
    CU(cudaMemsetAsync(b.states, 0, (size_t)Q * sizeof(DevState), s));
    CU(cudaMemsetAsync(b.queryTicket, 0, 4, s));
    CU(cudaMemsetAsync(b.pathLen, 0, (size_t)Q * 4, s));
    BatchArgs B{};
    B.base = make_args(ctx);
    B.base.treeState = b.treeState; B.base.treeCtrl = b.treeCtrl; B.base.treeParent = b.treeParent;
    B.base.chunkMask = b.chunkMask; B.base.blockSum = b.blockSum; B.base.ticket = b.ticket;
    B.base.stageState = b.stageState; B.base.stageCtrl = b.stageCtrl;
    B.base.candState = nullptr; B.base.candCtrl = nullptr; B.base.candParent = nullptr; B.base.candR1 = nullptr;
    B.base.candR2 = nullptr; B.base.candFlags = nullptr; B.base.iterLog = nullptr;
    B.Q = Q; B.numWorkspaces = numWs;
    B.initState = b.initState; B.initCtrl = b.initCtrl; B.goalXY = b.goalXY; B.seeds = b.seeds; B.states = b.states;
    B.paths = wantPath ? b.paths : nullptr; B.pathLen = b.pathLen; B.maxPath = wantPath;
    B.queryTicket = b.queryTicket; B.wsQuery = b.wsQuery; B.launchT0 = b.launchT0;
    B.treeStride = T; B.mapIntsStride = mapStride; B.chunkStride = 2 * ctx->chunksCap; B.blockStride = 3 * ctx->blocksCap;
    B.stageStride = 2 * M; B.mapSlab = b.mapSlab; B.c2 = ctx->c2;
    cfg.gridDim = dim3(numWs * cluster_size);
    CU(cudaLaunchKernelEx(&cfg, f, B));
    CU(cudaEventRecord(ctx->ev1, s));
    ctx->launches += 1;
    std::vector<DevState> hst((size_t)Q);
    CU(cudaMemcpyAsync(hst.data(), b.states, (size_t)Q * sizeof(DevState), cudaMemcpyDeviceToHost, s));
    if (wantPath) {
        CU(cudaMemcpyAsync(h_paths7, b.paths, (size_t)Q * wantPath * 28, cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(h_path_len, b.pathLen, (size_t)Q * 4, cudaMemcpyDeviceToHost, s));
    }
    unsigned long long t0 = 0;
    CU(cudaMemcpyAsync(&t0, b.launchT0, 8, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    if (device_ms) *device_ms = ms;
    if (out)
        for (int q = 0; q < Q; ++q) {
            const DevState& d = hst[q];
            out[q].stop = d.stop; out[q].iterations = d.iterationsDone; out[q].tree_size = d.treeSize;
            out[q].cost_to_goal = d.costToGoal; out[q].goal_index = d.goalIdx; out[q].expansions = d.expansions;
            out[q].device_ms = ms; out[q].kernel_launches = 1;
            /* the batch kernel stamps every query with the device clock: service start (pairsTested) and end (stepsDone) */
            out[q].done_ms = d.stepsDone >= t0 ? (float)((double)(d.stepsDone - t0) * 1e-6) : ms;
            out[q].service_ms = d.stepsDone >= d.pairsTested ? (float)((double)(d.stepsDone - d.pairsTested) * 1e-6) : ms;
        }
    return numWs;
}

/* ---- caller-owned stream ---------------------------------------------------------------------------------------- */
int kgmt_set_stream(kgmt_ctx* ctx, void* cuda_stream) {
    if (!ctx) return KGMT_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->ownStream;
    return KGMT_OK;
}

/* ---- sharded expansion: one iteration's candidates split over `world` ranks (BASELINE config 5) ----------------- */
size_t kgmt_shard_delta_ints(const kgmt_ctx* ctx) { return ctx ? 4 * (size_t)ctx->c1 + 4 * ctx->c2 : 0; }

static KArgs make_shard_args(const kgmt_ctx* ctx, int* d_delta) {
    KArgs A = make_args(ctx);
    const size_t c1 = (size_t)ctx->c1, c2 = ctx->c2;
    A.R1 = d_delta; A.R1Valid = d_delta + c1; A.R1Invalid = d_delta + 2 * c1; A.R1Avail = d_delta + 3 * c1;
    int* d2 = d_delta + 4 * c1;
    A.R2 = d2; A.R2Valid = d2 + c2; A.R2Invalid = d2 + 2 * c2; A.R2StampDelta = d2 + 3 * c2;
    return A;                        /* R2Stamp, R1Cov, R1Score stay the real (replicated) arrays: read-only here */
}

int kgmt_shard_expand(kgmt_ctx* ctx, int rank, int world, int* d_delta, kgmt_shard_info* out) {
    if (!ctx || !d_delta || world < 1 || world > 16 || rank < 0 || rank >= world) return fail(ctx, KGMT_ERR_INVALID, "bad shard arguments");
    if (!ctx->begun) return fail(ctx, KGMT_ERR_STATE, "kgmt_shard_expand before kgmt_begin / kgmt_seed_frontier");
    CU(cudaSetDevice(ctx->device));
    if (!ctx->shardPrefix) {
        CU(cudaMalloc(&ctx->shardPrefix, ctx->blocksCap * 4));
        CU(cudaMalloc(&ctx->shardTotal, 4));
        CU(cudaHostAlloc(&ctx->hShardTotal, 4, cudaHostAllocDefault));
    }
    const DevState& s = *ctx->hState;              /* as fetched by begin / the previous commit */
    const int numBlocks = (s.numChunks + BLK_CHUNKS - 1) / BLK_CHUNKS;
    /* contiguous, balanced ranges of 8192-candidate scan blocks: global slot order == rank-major order */
    const int base = numBlocks / world, rem = numBlocks % world;
    const int bLo = rank * base + std::min(rank, rem), bHi = bLo + base + (rank < rem ? 1 : 0);
    ctx->shardBlkLo = bLo; ctx->shardBlkHi = bHi; ctx->shardAccepted = 0;
    const int cLo = bLo * BLK_CHUNKS, cHi = std::min(bHi * BLK_CHUNKS, s.numChunks);
    if (s.stop == STOP_RUNNING) {
        const KArgs A = make_shard_args(ctx, d_delta);
        const int chunks = std::max(cHi - cLo, 0);
        const int grid = std::max(1, std::min(ctx->shardGrid, (chunks + WARPS - 1) / WARPS));
        shard_reset_kernel<<<32, 256, 0, ctx->stream>>>(A, grid * WARPS);
        if (chunks > 0) shard_entry(ctx->col)<<<grid, TILE, ctx->smemBytes, ctx->stream>>>(A, cLo, cHi);
        shard_prefix_kernel<<<1, TILE, 0, ctx->stream>>>(A, bLo, bHi, ctx->shardPrefix, ctx->shardTotal);
        CU(cudaGetLastError());
        ctx->launches += chunks > 0 ? 3 : 2;
        CU(cudaMemcpyAsync(ctx->hShardTotal, ctx->shardTotal, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        ctx->shardAccepted = *ctx->hShardTotal;
    }
    if (out) {
        out->iteration = s.itr; out->candidates = s.M; out->children = s.children; out->frontier = s.frontierCount;
        out->chunk_lo = cLo; out->chunk_hi = std::max(cHi, cLo); out->accepted_local = ctx->shardAccepted; out->stop = s.stop;
    }
    return KGMT_OK;
}

int kgmt_shard_pack(kgmt_ctx* ctx, void* d_send, int cap_rows) {
    if (!ctx || !d_send || cap_rows < 4 || (cap_rows & 3)) return fail(ctx, KGMT_ERR_INVALID, "cap_rows must be a positive multiple of 4");
    if (ctx->shardAccepted < 0) return fail(ctx, KGMT_ERR_STATE, "kgmt_shard_pack before kgmt_shard_expand");
    if (ctx->shardAccepted > cap_rows) return fail(ctx, KGMT_ERR_INVALID, "cap_rows %d < accepted rows %d", cap_rows, ctx->shardAccepted);
    CU(cudaSetDevice(ctx->device));
    const int nb = ctx->shardBlkHi - ctx->shardBlkLo;
    if (nb > 0 && ctx->shardAccepted > 0) {
        const KArgs A = make_args(ctx);
        unsigned char* b = (unsigned char*)d_send;
        shard_pack_kernel<<<std::min(nb, ctx->numSMs * 8), TILE, 0, ctx->stream>>>(
            A, ctx->shardBlkLo, ctx->shardBlkHi, ctx->shardPrefix, (float4*)b, (float4*)(b + (size_t)cap_rows * 16),
            (int*)(b + (size_t)cap_rows * 32));
        CU(cudaGetLastError());
        ctx->launches += 1;
    }
    return KGMT_OK;
}

int kgmt_shard_commit(kgmt_ctx* ctx, const void* d_recv, int cap_rows, const int* h_counts, int world, int* d_delta,
                      kgmt_iter_stats* out) {
    if (!ctx || !d_recv || !h_counts || !d_delta || world < 1 || world > 16 || cap_rows < 4 || (cap_rows & 3))
        return fail(ctx, KGMT_ERR_INVALID, "bad commit arguments");
    if (ctx->shardAccepted < 0) return fail(ctx, KGMT_ERR_STATE, "kgmt_shard_commit before kgmt_shard_expand");
    CU(cudaSetDevice(ctx->device));
    ShardCommit C{};
    C.recv = (const unsigned char*)d_recv; C.cap = cap_rows; C.world = world;
    long long total = 0;
    for (int g = 0; g < world; ++g) {
        if (h_counts[g] < 0 || h_counts[g] > cap_rows) return fail(ctx, KGMT_ERR_INVALID, "count of rank %d out of range", g);
        C.prefix[g] = (int)total; total += h_counts[g];
    }
    C.prefix[world] = (int)total;
    if (total > (long long)ctx->p.max_tree_size - ctx->hState->treeSize)
        return fail(ctx, KGMT_ERR_INVALID, "%lld accepted rows do not fit the tree", total);
    if (ctx->hState->stop == STOP_RUNNING) {
        const KArgs A = make_args(ctx);
        if (total > 0)
            shard_insert_kernel<<<(int)std::min<long long>((total + TILE - 1) / TILE, (long long)ctx->numSMs * 8), TILE, 0, ctx->stream>>>(A, C);
        shard_apply_kernel<<<(unsigned)std::min<size_t>((ctx->c2 + 255) / 256, (size_t)ctx->numSMs * 8), 256, 0, ctx->stream>>>(A, d_delta, ctx->c2);
        shard_finalize_kernel<<<1, TILE, 0, ctx->stream>>>(A, C);
        CU(cudaGetLastError());
        ctx->launches += total > 0 ? 3 : 2;
    }
    ctx->shardAccepted = -1;
    int rc = fetch_state(ctx);
    if (rc) return rc;
    if (out) {
        const DevState& s = *ctx->hState;
        out->iteration = s.lastItr; out->mode = s.lastMode; out->children = s.lastChildren; out->frontier = s.lastFrontier;
        out->candidates = s.lastM; out->accepted = s.lastAccepted; out->tree_size = s.treeSize; out->stop = s.stop;
        out->cost_to_goal = s.costToGoal; out->goal_index = s.goalIdx;
    }
    return KGMT_OK;
}

/* ---- sharded expansion over peer memory (NVLink / NVSwitch; SURVEY.md §8e, second mode, without NCCL) ------------ */
static int peer_alloc(kgmt_ctx* ctx) {
    kgmt_ctx::Peer& pr = ctx->peer;
    if (pr.block) return KGMT_OK;
    const size_t deltaBytes = kgmt_shard_delta_ints(ctx) * 4;
    pr.mailOff = (deltaBytes + 255) & ~(size_t)255;
    pr.raceOff = (pr.mailOff + PEER_MAX * sizeof(PeerMail) + 255) & ~(size_t)255;
    pr.blockBytes = std::max<size_t>(pr.raceOff + 256, (size_t)2 << 20);   /* its own allocation granule */
    CU(cudaMalloc(&pr.dRaceFlags, PEER_MAX * sizeof(int*)));
    CU(cudaMalloc(&pr.block, pr.blockBytes));
    CU(cudaMemset(pr.block, 0, pr.blockBytes));
    CU(cudaMalloc(&pr.plan, sizeof(PeerPlan)));
    CU(cudaMemset(pr.plan, 0, sizeof(PeerPlan)));
    CU(cudaHostAlloc(&pr.hPlan, sizeof(PeerPlan), cudaHostAllocDefault));
    if (!ctx->shardPrefix) {
        CU(cudaMalloc(&ctx->shardPrefix, ctx->blocksCap * 4));
        CU(cudaMalloc(&ctx->shardTotal, 4));
        CU(cudaHostAlloc(&ctx->hShardTotal, 4, cudaHostAllocDefault));
    }
    /* load every kernel of the exchange NOW: with lazy module loading the first launch of a kernel uploads its code, and
     * an upload queued behind a kernel that is spinning on a peer of the SAME device (single-process tests) deadlocks */
    {
        cudaFuncAttributes fa;
        const void* fns[] = {(const void*)shard_reset_kernel, (const void*)shard_entry(ctx->col), (const void*)shard_prefix_kernel,
                             (const void*)peer_counts_kernel, (const void*)peer_pack_kernel, (const void*)peer_reduce_kernel,
                             (const void*)peer_barrier_kernel, (const void*)recount_cov_kernel, (const void*)peer_finalize_kernel,
                             (const void*)fused_entry(ctx->col), (const void*)begin_kernel};
        for (const void* f : fns) CU(cudaFuncGetAttributes(&fa, f));
    }
    return KGMT_OK;
}

size_t kgmt_peer_handle_bytes(void) { return 5 * sizeof(cudaIpcMemHandle_t); }

/* handles of the five allocations a peer maps: tree state, tree ctrl, tree parent, map slab, exchange block */
int kgmt_peer_export(kgmt_ctx* ctx, void* out_handles, size_t bytes) {
    if (!ctx || !out_handles || bytes < kgmt_peer_handle_bytes()) return fail(ctx, KGMT_ERR_INVALID, "handle buffer too small");
    CU(cudaSetDevice(ctx->device));
    int rc = peer_alloc(ctx);
    if (rc) return rc;
    cudaIpcMemHandle_t* h = (cudaIpcMemHandle_t*)out_handles;
    CU(cudaIpcGetMemHandle(&h[0], ctx->treeState));
    CU(cudaIpcGetMemHandle(&h[1], ctx->treeCtrl));
    CU(cudaIpcGetMemHandle(&h[2], ctx->treeParent));
    CU(cudaIpcGetMemHandle(&h[3], ctx->mapSlab));
    CU(cudaIpcGetMemHandle(&h[4], ctx->peer.block));
    return KGMT_OK;
}

static int peer_publish_tables(kgmt_ctx* ctx) {
    /* a re-attached context restarts its exchange sequence at 0: clear what earlier exchanges left in THIS rank's
     * mailboxes and race word (stale seq words would satisfy the first waits).  Every rank must have attached before
     * any rank starts an exchange — the host program barriers after kgmt_peer_attach. */
    CU(cudaMemset(ctx->peer.block + ctx->peer.mailOff, 0, ctx->peer.blockBytes - ctx->peer.mailOff));
    int* flags[PEER_MAX] = {};
    for (int p = 0; p < ctx->peer.world; ++p) flags[p] = (int*)((unsigned char*)ctx->peer.args.delta[p] + ctx->peer.raceOff);
    CU(cudaMemcpy(ctx->peer.dRaceFlags, flags, sizeof(flags), cudaMemcpyHostToDevice));
    return KGMT_OK;
}

static void peer_set(kgmt_ctx* ctx, int p, void* ts, void* tc, void* tp, void* ms, void* blk) {
    PeerArgs& a = ctx->peer.args;
    a.treeState[p] = (float4*)ts; a.treeCtrl[p] = (float4*)tc; a.treeParent[p] = (int*)tp; a.mapSlab[p] = (int*)ms;
    a.delta[p] = (int*)blk; a.mail[p] = (PeerMail*)((unsigned char*)blk + ctx->peer.mailOff);
}

int kgmt_peer_detach(kgmt_ctx* ctx) {
    if (!ctx) return KGMT_ERR_INVALID;
    kgmt_ctx::Peer& pr = ctx->peer;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (pr.ipc)
        for (int p = 0; p < PEER_MAX; ++p)
            for (int k = 0; k < 5; ++k)
                if (pr.opened[p][k]) { cudaIpcCloseMemHandle(pr.opened[p][k]); pr.opened[p][k] = nullptr; }
    pr.rank = -1; pr.world = 0; pr.ipc = false; pr.inFlight = false;
    return KGMT_OK;
}

/* all_handles: world x kgmt_peer_handle_bytes(), rank-major, as gathered from every rank's kgmt_peer_export */
int kgmt_peer_attach(kgmt_ctx* ctx, int rank, int world, const void* all_handles) {
    if (!ctx || !all_handles || world < 1 || world > PEER_MAX || rank < 0 || rank >= world) return fail(ctx, KGMT_ERR_INVALID, "bad peer arguments");
    CU(cudaSetDevice(ctx->device));
    int rc = peer_alloc(ctx);
    if (rc) return rc;
    kgmt_peer_detach(ctx);
    kgmt_ctx::Peer& pr = ctx->peer;
    pr.args = PeerArgs{};
    const cudaIpcMemHandle_t* h = (const cudaIpcMemHandle_t*)all_handles;
    for (int p = 0; p < world; ++p) {
        if (p == rank) { peer_set(ctx, p, ctx->treeState, ctx->treeCtrl, ctx->treeParent, ctx->mapSlab, pr.block); continue; }
        void* ptr[5] = {};
        for (int k = 0; k < 5; ++k) {
            cudaError_t e = cudaIpcOpenMemHandle(&ptr[k], h[p * 5 + k], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) return fail(ctx, KGMT_ERR_COMM, "cudaIpcOpenMemHandle(rank %d, array %d): %s", p, k, cudaGetErrorString(e));
            pr.opened[p][k] = ptr[k];
        }
        peer_set(ctx, p, ptr[0], ptr[1], ptr[2], ptr[3], ptr[4]);
    }
    pr.rank = rank; pr.world = world; pr.seq = 0; pr.ipc = true;
    pr.args.rank = rank; pr.args.world = world; pr.args.plan = pr.plan;
    return peer_publish_tables(ctx);
}

/* the same wiring between contexts of ONE process (tests; several contexts may share a device) */
int kgmt_peer_attach_local(kgmt_ctx* ctx, int rank, int world, kgmt_ctx* const* peers) {
    if (!ctx || !peers || world < 1 || world > PEER_MAX || rank < 0 || rank >= world || peers[rank] != ctx)
        return fail(ctx, KGMT_ERR_INVALID, "bad peer arguments");
    CU(cudaSetDevice(ctx->device));
    kgmt_peer_detach(ctx);
    kgmt_ctx::Peer& pr = ctx->peer;
    pr.args = PeerArgs{};
    for (int p = 0; p < world; ++p) {
        kgmt_ctx* o = peers[p];
        if (!o || o->c1 != ctx->c1 || o->c2 != ctx->c2 || o->p.max_tree_size != ctx->p.max_tree_size)
            return fail(ctx, KGMT_ERR_INVALID, "peer %d has a different configuration", p);
        cudaSetDevice(o->device);
        int rc = peer_alloc(o);
        cudaSetDevice(ctx->device);
        if (rc) return fail(ctx, rc, "peer %d: %s", p, o->err);
        if (o->device != ctx->device) {
            int can = 0;
            cudaDeviceCanAccessPeer(&can, ctx->device, o->device);
            if (!can) return fail(ctx, KGMT_ERR_COMM, "device %d cannot map device %d", ctx->device, o->device);
            cudaError_t e = cudaDeviceEnablePeerAccess(o->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(ctx, KGMT_ERR_COMM, "peer access: %s", cudaGetErrorString(e));
            cudaGetLastError();
        }
        peer_set(ctx, p, o->treeState, o->treeCtrl, o->treeParent, o->mapSlab, o->peer.block);
    }
    pr.rank = rank; pr.world = world; pr.seq = 0; pr.ipc = false;
    pr.args.rank = rank; pr.args.world = world; pr.args.plan = pr.plan;
    return peer_publish_tables(ctx);
}

/* Portfolio race (first-solution termination over peer memory): every attached rank plans the SAME query with its own
 * seed in its usual single cooperative launch; the rank that reaches the goal first writes race_id into every peer's
 * race word (system-scope release over NVLink) and the others stop at their next iteration with KGMT_PEER_SOLVED.
 * race_id must be > 0 and grow from race to race.  No collective, no host round trip inside the race. */
int kgmt_peer_race(kgmt_ctx* ctx, const float* initial7, const float* goal7, int race_id, kgmt_result* out) {
    if (!ctx || race_id <= 0) return fail(ctx, KGMT_ERR_INVALID, "race_id must be positive");
    if (ctx->peer.rank < 0) return fail(ctx, KGMT_ERR_STATE, "kgmt_peer_race before kgmt_peer_attach");
    ctx->raceId = race_id;
    const int rc = kgmt_plan(ctx, initial7, goal7, out);
    ctx->raceId = 0;
    return rc;
}

/* enqueue ONE expansion iteration across the attached ranks (every rank must call it for the same iteration) */
int kgmt_peer_expand_begin(kgmt_ctx* ctx) {
    if (!ctx) return KGMT_ERR_INVALID;
    kgmt_ctx::Peer& pr = ctx->peer;
    if (pr.rank < 0) return fail(ctx, KGMT_ERR_STATE, "kgmt_peer_expand_begin before kgmt_peer_attach");
    if (!ctx->begun) return fail(ctx, KGMT_ERR_STATE, "kgmt_peer_expand_begin before kgmt_begin / kgmt_seed_frontier");
    if (pr.inFlight) return fail(ctx, KGMT_ERR_STATE, "kgmt_peer_expand_begin twice without kgmt_peer_expand_end");
    CU(cudaSetDevice(ctx->device));
    const DevState& s = *ctx->hState;
    if (s.stop != STOP_RUNNING) { pr.inFlight = true; return KGMT_OK; }      /* nothing to run; _end reports the stop */
    const int rank = pr.rank, world = pr.world;
    const int numBlocks = (s.numChunks + BLK_CHUNKS - 1) / BLK_CHUNKS;
    const int base = numBlocks / world, rem = numBlocks % world;
    const int bLo = rank * base + std::min(rank, rem), bHi = bLo + base + (rank < rem ? 1 : 0);
    const int cLo = bLo * BLK_CHUNKS, cHi = std::min(bHi * BLK_CHUNKS, s.numChunks);
    const int chunks = std::max(cHi - cLo, 0);
    cudaStream_t st = ctx->stream;
    pr.seq += 1;
    pr.args.seq = pr.seq;
    CU(cudaMemsetAsync(pr.plan, 0, sizeof(PeerPlan), st));      /* also clears the error flag of a timed-out exchange */
    const KArgs As = make_shard_args(ctx, (int*)pr.block);
    const KArgs A = make_args(ctx);
    const int grid = std::max(1, std::min(ctx->shardGrid, (chunks + WARPS - 1) / WARPS));
    shard_reset_kernel<<<32, 256, 0, st>>>(As, grid * WARPS);
    if (chunks > 0) shard_entry(ctx->col)<<<grid, TILE, ctx->smemBytes, st>>>(As, cLo, cHi);
    shard_prefix_kernel<<<1, TILE, 0, st>>>(As, bLo, bHi, ctx->shardPrefix, ctx->shardTotal);
    peer_counts_kernel<<<1, 32, 0, st>>>(pr.args, ctx->shardTotal);
    if (bHi > bLo) peer_pack_kernel<<<std::min(bHi - bLo, ctx->numSMs * 8), TILE, 0, st>>>(A, pr.args, bLo, bHi, ctx->shardPrefix);
    {
        const size_t total = kgmt_shard_delta_ints(ctx);
        const size_t per = total / world, extra = total % world;
        const size_t lo = rank * per + std::min<size_t>(rank, extra), hi = lo + per + ((size_t)rank < extra ? 1 : 0);
        if (hi > lo)
            peer_reduce_kernel<<<(unsigned)std::min<size_t>((hi - lo + 255) / 256, (size_t)ctx->numSMs * 8), 256, 0, st>>>(A, pr.args, ctx->c2, lo, hi);
    }
    peer_barrier_kernel<<<1, 32, 0, st>>>(pr.args);
    recount_cov_kernel<<<ctx->c1, 128, 0, st>>>(A);
    CU(cudaMemsetAsync(pr.block, 0, kgmt_shard_delta_ints(ctx) * 4, st));
    peer_finalize_kernel<<<1, TILE, 0, st>>>(A, pr.args);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(pr.hPlan, pr.plan, sizeof(PeerPlan), cudaMemcpyDeviceToHost, st));
    ctx->launches += 9;
    pr.inFlight = true;
    return KGMT_OK;
}

int kgmt_peer_expand_end(kgmt_ctx* ctx, kgmt_iter_stats* out) {
    if (!ctx) return KGMT_ERR_INVALID;
    kgmt_ctx::Peer& pr = ctx->peer;
    if (!pr.inFlight) return fail(ctx, KGMT_ERR_STATE, "kgmt_peer_expand_end without kgmt_peer_expand_begin");
    CU(cudaSetDevice(ctx->device));
    pr.inFlight = false;
    const bool ran = ctx->hState->stop == STOP_RUNNING;
    int rc = fetch_state(ctx);
    if (rc) return rc;
    if (ran && pr.hPlan->err) return fail(ctx, KGMT_ERR_COMM, "a peer did not arrive within 5 s (rank %d of %d, exchange %d)", pr.rank, pr.world, pr.seq);
    if (out) {
        const DevState& s = *ctx->hState;
        out->iteration = s.lastItr; out->mode = s.lastMode; out->children = s.lastChildren; out->frontier = s.lastFrontier;
        out->candidates = s.lastM; out->accepted = s.lastAccepted; out->tree_size = s.treeSize; out->stop = s.stop;
        out->cost_to_goal = s.costToGoal; out->goal_index = s.goalIdx;
    }
    return KGMT_OK;
}

/* ---- the fused form: compute + exchange in ONE persistent cooperative kernel per rank ---------------------------- */
static int launch_fused(kgmt_ctx* ctx, int maxIters) {
    kgmt_ctx::Peer& pr = ctx->peer;
    const KArgs A = make_args(ctx);
    const KArgs As = make_shard_args(ctx, (int*)pr.block);
    pr.args.seq = pr.seq;
    int seq0 = pr.seq;
    size_t c2 = ctx->c2;
    CU(cudaMemsetAsync(pr.plan, 0, sizeof(PeerPlan), ctx->stream));
    /* resident CTAs per SM: the same knob as the single-GPU loop (params.reserved[1]); never more CTAs than can be co-resident */
    int grid = ctx->fusedGrid;
    if (ctx->p.reserved[1] > 0) grid = std::min(grid, ctx->p.reserved[1] * ctx->numSMs);
    const size_t histBytes = ctx->useHist ? (((size_t)2 * ctx->c1 * 4 + 15) & ~(size_t)15) : 0;
    size_t smem = (ctx->col == COL_BRUTE_STREAM) ? histBytes : ctx->smemBytes;
    PeerArgs P = pr.args;
    void* args[] = {(void*)&A, (void*)&As, (void*)&P, (void*)&maxIters, (void*)&seq0, (void*)&c2};
    CU(cudaLaunchCooperativeKernel((const void*)fused_entry(ctx->col), dim3(grid), dim3(TILE), args, smem, ctx->stream));
    CU(cudaMemcpyAsync(pr.hPlan, pr.plan, sizeof(PeerPlan), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->launches += 1;
    ctx->planLaunches += 1;
    return KGMT_OK;
}

static int finish_fused(kgmt_ctx* ctx, int itersBefore) {
    kgmt_ctx::Peer& pr = ctx->peer;
    int rc = fetch_state(ctx);
    if (rc) return rc;
    if (pr.hPlan->err) return fail(ctx, KGMT_ERR_COMM, "a peer did not arrive within 5 s (rank %d of %d)", pr.rank, pr.world);
    pr.seq += ctx->hState->iterationsDone - itersBefore;          /* one exchange per iteration, the same count on every rank */
    return KGMT_OK;
}

/* up to `count` sharded iterations in ONE launch per rank (every attached rank calls it for the same iterations) */
int kgmt_peer_expand_iterations(kgmt_ctx* ctx, int count, kgmt_iter_stats* out) {
    if (!ctx || count < 1) return KGMT_ERR_INVALID;
    kgmt_ctx::Peer& pr = ctx->peer;
    if (pr.rank < 0) return fail(ctx, KGMT_ERR_STATE, "kgmt_peer_expand_iterations before kgmt_peer_attach");
    if (!ctx->begun) return fail(ctx, KGMT_ERR_STATE, "kgmt_peer_expand_iterations before kgmt_begin / kgmt_seed_frontier");
    if (pr.inFlight) return fail(ctx, KGMT_ERR_STATE, "an exchange of kgmt_peer_expand_begin is still in flight");
    CU(cudaSetDevice(ctx->device));
    if (ctx->hState->stop == STOP_RUNNING) {
        const int before = ctx->hState->iterationsDone;
        int rc = launch_fused(ctx, count);
        if (rc) return rc;
        rc = finish_fused(ctx, before);
        if (rc) return rc;
    }
    if (out) {
        const DevState& s = *ctx->hState;
        out->iteration = s.lastItr; out->mode = s.lastMode; out->children = s.lastChildren; out->frontier = s.lastFrontier;
        out->candidates = s.lastM; out->accepted = s.lastAccepted; out->tree_size = s.treeSize; out->stop = s.stop;
        out->cost_to_goal = s.costToGoal; out->goal_index = s.goalIdx;
    }
    return KGMT_OK;
}

/* KGMT::plan with every iteration's candidates split over the attached ranks: root insertion on every rank (same seed,
 * same root: identical replicas), then the whole loop in one persistent launch per rank.  Every rank returns the same
 * result and holds the same tree as kgmt_plan on one GPU. */
int kgmt_peer_plan(kgmt_ctx* ctx, const float* initial7, const float* goal7, kgmt_result* out) {
    if (!ctx || !initial7 || !goal7) return fail(ctx, KGMT_ERR_INVALID, "null initial/goal");
    kgmt_ctx::Peer& pr = ctx->peer;
    if (pr.rank < 0) return fail(ctx, KGMT_ERR_STATE, "kgmt_peer_plan before kgmt_peer_attach");
    if (pr.inFlight) return fail(ctx, KGMT_ERR_STATE, "an exchange of kgmt_peer_expand_begin is still in flight");
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventRecord(ctx->ev0, ctx->stream));
    if (ctx->begun) { int rc = clear_state(ctx, false, true); if (rc) return rc; }
    memcpy(ctx->goal, goal7, sizeof(ctx->goal));
    const KArgs A = make_args(ctx);
    begin_kernel<<<1, TILE, 0, ctx->stream>>>(A, make_float4(initial7[0], initial7[1], initial7[2], initial7[3]),
                                             make_float4(initial7[4], initial7[5], initial7[6], 0.f), ctx->hState->forceChildren);
    CU(cudaGetLastError());
    ctx->launches += 1;
    ctx->planLaunches = 1;
    { int rc = launch_fused(ctx, 0x7fffffff); if (rc) return rc; }
    CU(cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->begun = true;
    int rc = finish_fused(ctx, 0);
    if (rc) return rc;
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    if (out) fill_result(ctx, out, ms);
    return KGMT_OK;
}

/* ---- stage-level entry points ------------------------------------------------------------- */
int kgmt_stage_scores(kgmt_ctx* ctx) {
    if (!ctx) return KGMT_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    const KArgs A = make_args(ctx);
    recount_cov_kernel<<<ctx->c1, 128, 0, ctx->stream>>>(A);
    scores_kernel<<<1, TILE, 0, ctx->stream>>>(A);
    CU(cudaGetLastError());
    ctx->launches += 2;
    CU(cudaStreamSynchronize(ctx->stream));
    return KGMT_OK;
}

int kgmt_stage_propagate(kgmt_ctx* ctx, const float* h_parents7, int P, int children, uint32_t key0, uint32_t slot0,
                         float* device_ms) {
    if (!ctx || !h_parents7 || P < 1 || children < 1) return fail(ctx, KGMT_ERR_INVALID, "bad parents/children");
    const long long M = (long long)P * children;
    if (M > ctx->maxCand) return fail(ctx, KGMT_ERR_INVALID, "P*children = %lld exceeds max_candidates %d", M, ctx->maxCand);
    CU(cudaSetDevice(ctx->device));
    int rc = ensure_record(ctx);
    if (rc) return rc;
    if ((size_t)P > ctx->parentsCap) {
        if (ctx->dParents) cudaFree(ctx->dParents);
        ctx->dParents = nullptr; ctx->parentsCap = 0;
        CU(cudaMalloc(&ctx->dParents, (size_t)P * 16));
        ctx->parentsCap = P;
    }
    std::vector<float> st((size_t)P * 4);
    for (int i = 0; i < P; ++i) memcpy(&st[(size_t)i * 4], &h_parents7[(size_t)i * 7], 16);
    CU(cudaMemcpyAsync(ctx->dParents, st.data(), (size_t)P * 16, cudaMemcpyHostToDevice, ctx->stream));
    const KArgs A = make_args(ctx);
    const size_t smem = ctx->smemBytes - (ctx->useHist ? (((size_t)2 * ctx->c1 * 4 + 15) & ~(size_t)15) : 0);
    const long long tiles = (M + TILE - 1) / TILE;
    const int grid = (int)std::min<long long>(tiles, (long long)ctx->gridMax * 4);
    CU(cudaEventRecord(ctx->ev0, ctx->stream));
    propagate_entry(ctx->col)<<<grid, TILE, smem, ctx->stream>>>(A, ctx->dParents, M, children, key0, slot0);
    CU(cudaGetLastError());
    CU(cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->launches += 1;
    ctx->dirtyCand = ctx->maxCand;
    CU(cudaStreamSynchronize(ctx->stream));
    if (device_ms) CU(cudaEventElapsedTime(device_ms, ctx->ev0, ctx->ev1));
    return KGMT_OK;
}

int kgmt_seed_frontier(kgmt_ctx* ctx, const float* h_nodes7, int count, const float* goal7) {
    if (!ctx || !h_nodes7 || !goal7 || count < 1 || count > ctx->p.max_tree_size)
        return fail(ctx, KGMT_ERR_INVALID, "bad frontier");
    CU(cudaSetDevice(ctx->device));
    if (ctx->begun) { int rc = clear_state(ctx, true); if (rc) return rc; }
    memcpy(ctx->goal, goal7, sizeof(ctx->goal));
    int rc = ensure_scratch(ctx, (size_t)count * 28);
    if (rc) return rc;
    CU(cudaMemcpyAsync(ctx->scratch, h_nodes7, (size_t)count * 28, cudaMemcpyHostToDevice, ctx->stream));
    scatter_samples_kernel<<<(count + 255) / 256, 256, 0, ctx->stream>>>((const float*)ctx->scratch, ctx->treeState,
                                                                         ctx->treeCtrl, count);
    const KArgs A = make_args(ctx);
    seed_mark_kernel<<<(count + 255) / 256, 256, 0, ctx->stream>>>(A, count);
    recount_cov_kernel<<<ctx->c1, 128, 0, ctx->stream>>>(A);
    seed_finish_kernel<<<1, TILE, 0, ctx->stream>>>(A, count);
    CU(cudaGetLastError());
    ctx->launches += 4;
    ctx->planLaunches = 4;
    ctx->begun = true;
    return fetch_state(ctx);
}

/* Stage 5a alone on caller-supplied candidates (see kgmt_c.h): upload the records, then the planner's own chunk_finish. */
int kgmt_stage_update_maps(kgmt_ctx* ctx, const float* h_cand7, const unsigned char* h_valid, const float* h_u3,
                           const int* h_parent, int M) {
    if (!ctx || !h_cand7 || !h_valid || !h_u3 || !h_parent || M < 1) return fail(ctx, KGMT_ERR_INVALID, "bad candidate arrays");
    if (!ctx->begun) return fail(ctx, KGMT_ERR_STATE, "kgmt_stage_update_maps before kgmt_begin / kgmt_seed_frontier");
    if (M > ctx->maxCand) return fail(ctx, KGMT_ERR_INVALID, "M = %d exceeds max_candidates %d", M, ctx->maxCand);
    CU(cudaSetDevice(ctx->device));
    int rc = ensure_record(ctx);
    if (rc) return rc;
    rc = fetch_state(ctx);
    if (rc) return rc;
    DevState& st = *ctx->hState;
    if (st.stop != STOP_RUNNING) return fail(ctx, KGMT_ERR_STATE, "the planner has stopped (%d)", st.stop);
    if ((long long)M > (long long)ctx->p.max_tree_size - st.treeSize)
        return fail(ctx, KGMT_ERR_INVALID, "M = %d candidates could not all be inserted (tree has %d free rows)", M, ctx->p.max_tree_size - st.treeSize);
    /* records in the device layout: float4 state | float4 (a, steering, duration, u3) | parent | flags */
    std::vector<float> hs((size_t)M * 4), hc((size_t)M * 4);
    std::vector<unsigned char> hf((size_t)M);
    for (int i = 0; i < M; ++i) {
        memcpy(&hs[(size_t)i * 4], &h_cand7[(size_t)i * 7], 16);
        hc[(size_t)i * 4] = h_cand7[(size_t)i * 7 + 4]; hc[(size_t)i * 4 + 1] = h_cand7[(size_t)i * 7 + 5];
        hc[(size_t)i * 4 + 2] = h_cand7[(size_t)i * 7 + 6]; hc[(size_t)i * 4 + 3] = h_u3[i];
        hf[i] = h_valid[i] ? FLAG_VALID : 0;
    }
    cudaStream_t s = ctx->stream;
    CU(cudaMemcpyAsync(ctx->candState, hs.data(), (size_t)M * 16, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(ctx->candCtrl, hc.data(), (size_t)M * 16, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(ctx->candParent, h_parent, (size_t)M * 4, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(ctx->candFlags, hf.data(), (size_t)M, cudaMemcpyHostToDevice, s));
    /* the pending iteration now has exactly these candidates (mode 5 = staged by the caller) */
    st.mode = 5; st.children = 1; st.M = M; st.numChunks = (M + CHUNK - 1) / CHUNK;
    CU(cudaMemcpyAsync(ctx->dState, ctx->hState, COPIED_WORDS * 4, cudaMemcpyHostToDevice, s));
    rc = reset_loop_bookkeeping(ctx);               /* clean block sums, whatever ran before */
    if (rc) return rc;
    KArgs A = make_args(ctx);
    A.useHist = 0;                                  /* R1 counters straight to global memory */
    const int grid = std::max(1, std::min((st.numChunks + WARPS - 1) / WARPS, ctx->numSMs * 8));
    stage_update_kernel<<<grid, TILE, 0, s>>>(A);
    CU(cudaGetLastError());
    ctx->launches += 1;
    ctx->dirtyCand = ctx->maxCand;
    CU(cudaStreamSynchronize(s));                   /* the host staging vectors go out of scope */
    return KGMT_OK;
}

/* Stage 5b alone: ordered insertion of what kgmt_stage_update_maps accepted + the end of the while-loop body. */
int kgmt_stage_insert(kgmt_ctx* ctx, kgmt_iter_stats* out) {
    if (!ctx) return KGMT_ERR_INVALID;
    if (!ctx->begun || !ctx->recordAllocated || ctx->hState->mode != 5)
        return fail(ctx, KGMT_ERR_STATE, "kgmt_stage_insert needs a preceding kgmt_stage_update_maps");
    CU(cudaSetDevice(ctx->device));
    const KArgs A = make_args(ctx);
    stage_insert_kernel<<<1, TILE, 0, ctx->stream>>>(A);
    CU(cudaGetLastError());
    ctx->launches += 1;
    int rc = reset_loop_bookkeeping(ctx);
    if (rc) return rc;
    rc = fetch_state(ctx);
    if (rc) return rc;
    if (out) {
        const DevState& s = *ctx->hState;
        out->iteration = s.lastItr; out->mode = s.lastMode; out->children = s.lastChildren; out->frontier = s.lastFrontier;
        out->candidates = s.lastM; out->accepted = s.lastAccepted; out->tree_size = s.treeSize; out->stop = s.stop;
        out->cost_to_goal = s.costToGoal; out->goal_index = s.goalIdx;
    }
    return KGMT_OK;
}

/* kgmt_set_children: > 0 forces that many children per frontier node in every later iteration
 * (throughput sweeps, BASELINE config 5); 0 restores the reference policy (KGMT.cu:151-158). */
int kgmt_set_children(kgmt_ctx* ctx, int children) {
    if (!ctx || children < 0) return KGMT_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    int rc = fetch_state(ctx);
    if (rc) return rc;
    ctx->hState->forceChildren = children;
    if (ctx->begun && ctx->hState->stop == STOP_RUNNING) {
        /* re-shape the pending iteration */
        DevState& st = *ctx->hState;
        const int remaining = ctx->p.max_tree_size - st.treeSize;
        if (children > 0 && (long long)st.frontierCount * children > std::min<long long>(remaining, ctx->maxCand))
            return fail(ctx, KGMT_ERR_INVALID, "frontier*children exceeds the remaining tree/candidate capacity");
        int mode, ch, M;
        expansion_shape(st.frontierCount, st.treeSize, ctx->p.max_tree_size, ctx->maxCand, children, mode, ch, M);
        st.mode = mode; st.children = ch; st.M = M; st.numChunks = (M + CHUNK - 1) / CHUNK;
    }
    CU(cudaMemcpyAsync(ctx->dState, ctx->hState, sizeof(DevState), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return KGMT_OK;
}

int kgmt_checkpoint(kgmt_ctx* ctx) {
    if (!ctx) return KGMT_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    if (!ctx->mapSlabCkpt) CU(cudaMalloc(&ctx->mapSlabCkpt, ctx->mapSlabInts * 4));
    CU(cudaMemcpyAsync(ctx->mapSlabCkpt, ctx->mapSlab, ctx->mapSlabInts * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    int rc = fetch_state(ctx);
    if (rc) return rc;
    ctx->ckptState = *ctx->hState;
    ctx->haveCkpt = true;
    return KGMT_OK;
}

int kgmt_restore(kgmt_ctx* ctx) {
    if (!ctx) return KGMT_ERR_INVALID;
    if (!ctx->haveCkpt) return fail(ctx, KGMT_ERR_STATE, "kgmt_restore without kgmt_checkpoint");
    CU(cudaSetDevice(ctx->device));
    ctx->pathPreLaunches = -1;
    int rc = fetch_state(ctx);
    if (rc) return rc;
    CU(cudaMemcpyAsync(ctx->mapSlab, ctx->mapSlabCkpt, ctx->mapSlabInts * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    *ctx->hState = ctx->ckptState;
    {   /* insertion bookkeeping back to "treeSize rows present, scan buffers clean" */
        int rc2 = reset_loop_bookkeeping(ctx);
        if (rc2) return rc2;
    }
    CU(cudaMemcpyAsync(ctx->dState, ctx->hState, sizeof(DevState), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return KGMT_OK;
}

/* ---- data exchange -------------------------------------------------------------------------- */
size_t kgmt_array_bytes(const kgmt_ctx* ctx, int id) {
    if (!ctx) return 0;
    const size_t T = (size_t)ctx->p.max_tree_size, M = (size_t)ctx->maxCand, c1 = (size_t)ctx->c1, c2 = ctx->c2;
    switch (id) {
        case KGMT_ARR_SAMPLES: return T * 28;
        case KGMT_ARR_UNEXPLORED: return M * 28;
        case KGMT_ARR_PARENT: return T * 4;
        case KGMT_ARR_U_PARENT: return M * 4;
        case KGMT_ARR_G: return T;
        case KGMT_ARR_R2AVAIL: case KGMT_ARR_R2VALID: case KGMT_ARR_R2INVALID: case KGMT_ARR_R2: return c2 * 4;
        case KGMT_ARR_R1AVAIL: case KGMT_ARR_R1VALID: case KGMT_ARR_R1INVALID: case KGMT_ARR_R1SCORE: case KGMT_ARR_R1:
            return c1 * 4;
        case KGMT_ARR_COSTS: return T * 4;
        case KGMT_ARR_U_VALID: case KGMT_ARR_U_ACCEPT: return M;
        case KGMT_ARR_U_R1: case KGMT_ARR_U_R2: case KGMT_ARR_U_U3: return M * 4;
        default: return 0;
    }
}

int kgmt_export(kgmt_ctx* ctx, int id, void* h_dst, size_t bytes) {
    if (!ctx || !h_dst) return KGMT_ERR_INVALID;
    const size_t need = kgmt_array_bytes(ctx, id);
    if (need == 0) return fail(ctx, KGMT_ERR_INVALID, "unknown array id %d", id);
    if (bytes < need) return fail(ctx, KGMT_ERR_INVALID, "array %d needs %zu bytes, buffer has %zu", id, need, bytes);
    CU(cudaSetDevice(ctx->device));
    const int T = ctx->p.max_tree_size, M = ctx->maxCand;
    const bool candId = (id == KGMT_ARR_UNEXPLORED || id == KGMT_ARR_U_PARENT || id >= KGMT_ARR_U_VALID);
    if (candId && !ctx->recordAllocated)
        return fail(ctx, KGMT_ERR_STATE, "array %d is only kept when record_candidates = 1", id);
    const void* src = nullptr;
    cudaStream_t s = ctx->stream;
    switch (id) {
        case KGMT_ARR_SAMPLES: {
            int rc = ensure_scratch(ctx, need); if (rc) return rc;
            gather_samples_kernel<<<(T + 255) / 256, 256, 0, s>>>(ctx->treeState, ctx->treeCtrl, (float*)ctx->scratch, T, 0);
            src = ctx->scratch; break; }
        case KGMT_ARR_UNEXPLORED: {
            int rc = ensure_scratch(ctx, need); if (rc) return rc;
            gather_samples_kernel<<<(M + 255) / 256, 256, 0, s>>>(ctx->candState, ctx->candCtrl, (float*)ctx->scratch, M, 1);
            src = ctx->scratch; break; }
        case KGMT_ARR_COSTS: {
            int rc = ensure_scratch(ctx, need); if (rc) return rc;
            gather_w_kernel<<<(T + 255) / 256, 256, 0, s>>>(ctx->treeCtrl, (float*)ctx->scratch, T);
            src = ctx->scratch; break; }
        case KGMT_ARR_U_U3: {
            int rc = ensure_scratch(ctx, need); if (rc) return rc;
            gather_w_kernel<<<(M + 255) / 256, 256, 0, s>>>(ctx->candCtrl, (float*)ctx->scratch, M);
            src = ctx->scratch; break; }
        case KGMT_ARR_G: {
            int rc = ensure_scratch(ctx, need); if (rc) return rc;
            rc = fetch_state(ctx); if (rc) return rc;
            const bool running = ctx->begun;
            frontier_flags_kernel<<<(T + 255) / 256, 256, 0, s>>>((unsigned char*)ctx->scratch, T,
                running ? ctx->hState->frontierStart : 0, running ? ctx->hState->frontierCount : 0);
            src = ctx->scratch; break; }
        case KGMT_ARR_R2AVAIL: {
            int rc = ensure_scratch(ctx, need); if (rc) return rc;
            stamp_to_avail_kernel<<<(unsigned)((ctx->c2 + 255) / 256), 256, 0, s>>>(ctx->R2Stamp, (int*)ctx->scratch, (int)ctx->c2);
            src = ctx->scratch; break; }
        case KGMT_ARR_U_VALID: case KGMT_ARR_U_ACCEPT: {
            int rc = ensure_scratch(ctx, need); if (rc) return rc;
            flags_bit_kernel<<<(M + 255) / 256, 256, 0, s>>>(ctx->candFlags, (unsigned char*)ctx->scratch, M,
                                                            id == KGMT_ARR_U_VALID ? FLAG_VALID : FLAG_ACCEPT);
            src = ctx->scratch; break; }
        case KGMT_ARR_PARENT: src = ctx->treeParent; break;
        case KGMT_ARR_U_PARENT: src = ctx->candParent; break;
        case KGMT_ARR_R1AVAIL: src = ctx->R1Avail; break;
        case KGMT_ARR_R1VALID: src = ctx->R1Valid; break;
        case KGMT_ARR_R2VALID: src = ctx->R2Valid; break;
        case KGMT_ARR_R1INVALID: src = ctx->R1Invalid; break;
        case KGMT_ARR_R2INVALID: src = ctx->R2Invalid; break;
        case KGMT_ARR_R1SCORE: { int rc = fetch_state(ctx); if (rc) return rc; src = ctx->R1Score[ctx->hState->itr & 1]; break; }
        case KGMT_ARR_R1: src = ctx->R1; break;
        case KGMT_ARR_R2: src = ctx->R2; break;
        case KGMT_ARR_U_R1: src = ctx->candR1; break;
        case KGMT_ARR_U_R2: src = ctx->candR2; break;
        default: return fail(ctx, KGMT_ERR_INVALID, "unknown array id %d", id);
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(h_dst, src, need, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (id == KGMT_ARR_SAMPLES || id == KGMT_ARR_PARENT || id == KGMT_ARR_COSTS) {
        /* tree rows at or above treeSize are dead storage (a re-plan does not rewrite them): report them as the
         * reference's constructor leaves unused rows — zeros, parent -1 (KGMT.cu:25-26,40) */
        int rc = fetch_state(ctx);
        if (rc) return rc;
        const size_t live = ctx->begun ? (size_t)std::max(ctx->hState->treeSize, 0) : 0;
        const size_t rows = (size_t)T;
        if (live < rows) {
            if (id == KGMT_ARR_SAMPLES) memset((float*)h_dst + live * 7, 0, (rows - live) * 28);
            else if (id == KGMT_ARR_COSTS) memset((float*)h_dst + live, 0, (rows - live) * 4);
            else memset((int*)h_dst + live, 0xFF, (rows - live) * 4);
        }
    }
    return KGMT_OK;
}

int kgmt_import(kgmt_ctx* ctx, int id, const void* h_src, size_t bytes) {
    if (!ctx || !h_src) return KGMT_ERR_INVALID;
    if (id < KGMT_ARR_R2AVAIL || id > KGMT_ARR_R2) return fail(ctx, KGMT_ERR_INVALID, "only map arrays (5..13) can be imported");
    const size_t need = kgmt_array_bytes(ctx, id);
    if (bytes < need) return fail(ctx, KGMT_ERR_INVALID, "array %d needs %zu bytes", id, need);
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    void* dst = nullptr;
    switch (id) {
        case KGMT_ARR_R2AVAIL: {
            int rc = ensure_scratch(ctx, need); if (rc) return rc;
            CU(cudaMemcpyAsync(ctx->scratch, h_src, need, cudaMemcpyHostToDevice, s));
            avail_to_stamp_kernel<<<(unsigned)((ctx->c2 + 255) / 256), 256, 0, s>>>((const int*)ctx->scratch, ctx->R2Stamp, (int)ctx->c2);
            const KArgs A = make_args(ctx);
            recount_cov_kernel<<<ctx->c1, 128, 0, s>>>(A);
            CU(cudaGetLastError());
            CU(cudaStreamSynchronize(s));
            return KGMT_OK; }
        case KGMT_ARR_R1AVAIL: dst = ctx->R1Avail; break;
        case KGMT_ARR_R1VALID: dst = ctx->R1Valid; break;
        case KGMT_ARR_R2VALID: dst = ctx->R2Valid; break;
        case KGMT_ARR_R1INVALID: dst = ctx->R1Invalid; break;
        case KGMT_ARR_R2INVALID: dst = ctx->R2Invalid; break;
        case KGMT_ARR_R1SCORE: { int rc = fetch_state(ctx); if (rc) return rc; dst = ctx->R1Score[ctx->hState->itr & 1]; break; }
        case KGMT_ARR_R1: dst = ctx->R1; break;
        case KGMT_ARR_R2: dst = ctx->R2; break;
        default: return KGMT_ERR_INVALID;
    }
    CU(cudaMemcpyAsync(dst, h_src, need, cudaMemcpyHostToDevice, s));
    CU(cudaStreamSynchronize(s));
    return KGMT_OK;
}

/* parent-link back-trace, root first (the reference stops at costToGoal) */
int kgmt_extract_path(kgmt_ctx* ctx, int node, float* h_rows7, int max_rows) {
    if (!ctx || max_rows < 0 || (max_rows > 0 && !h_rows7)) return KGMT_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    /* the host copy of the scalars is current after every call of this library except between peer begin / end */
    if (ctx->peer.inFlight) { int rc0 = fetch_state(ctx); if (rc0) return rc0; }
    const int T = ctx->hState->treeSize;
    if (node < 0) node = ctx->hState->goalIdx;
    if (node < 0 || node >= T) return fail(ctx, KGMT_ERR_INVALID, "no such node %d (tree size %d)", node, T);
    if (!ctx->peer.inFlight && ctx->hPathPre && ctx->pathPreLaunches == ctx->launches && node == ctx->pathPreGoal &&
        T == ctx->pathPreTree) {
        /* traced by kgmt_plan behind the planner kernel and nothing has run since */
        CU(cudaEventSynchronize(ctx->evPath));
        int len = 0;
        memcpy(&len, ctx->hPathPre, 4);
        if (len >= 0 && len <= PATH_PRE_ROWS) {
            const size_t rows = std::min<size_t>((size_t)len, (size_t)std::max(max_rows, 0));
            if (rows) memcpy(h_rows7, ctx->hPathPre + 16, rows * 28);
            return len;
        }
    }
    /* back-trace on the device (the parent links are L2-resident), then ONE device-to-host copy of the length and the
     * first rows (solution paths are a few dozen nodes); a second copy only for longer chains */
    const size_t rowsCap = (size_t)std::max(max_rows, 0);
    int rc = ensure_scratch(ctx, 16 + rowsCap * 28);
    if (rc) return rc;
    const KArgs A = make_args(ctx);
    trace_path_kernel<<<1, 32, 0, ctx->stream>>>(A, node, T, (int*)ctx->scratch, (float*)((char*)ctx->scratch + 16), max_rows);
    CU(cudaGetLastError());
    ctx->launches += 1;
    const size_t first = std::min<size_t>(rowsCap, 128);
    ctx->hPath.resize(16 + first * 28);
    CU(cudaMemcpyAsync(ctx->hPath.data(), ctx->scratch, 16 + first * 28, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    int len = 0;
    memcpy(&len, ctx->hPath.data(), 4);
    const size_t rows = std::min<size_t>((size_t)std::max(len, 0), rowsCap);
    memcpy(h_rows7, ctx->hPath.data() + 16, std::min(rows, first) * 28);
    if (rows > first) {
        CU(cudaMemcpyAsync(h_rows7 + first * 7, (char*)ctx->scratch + 16 + first * 28, (rows - first) * 28, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return len;
}

/* the reference's 13 CSV dumps, KGMT.cu:299-311, in the format of helper.cuh:53-72 */
static int write_csv(kgmt_ctx* ctx, const std::string& path, const void* data, int kind, size_t rows, int cols) {
    FILE* f = fopen(path.c_str(), "w");
    if (!f) return fail(ctx, KGMT_ERR_INVALID, "cannot open %s", path.c_str());
    for (size_t i = 0; i < rows; ++i) {
        for (int j = 0; j < cols; ++j) {
            const size_t at = i * cols + j;
            if (kind == 0) fprintf(f, "%.10f", (double)((const float*)data)[at]);
            else if (kind == 1) fprintf(f, "%d", ((const int*)data)[at]);
            else fprintf(f, "%d", (int)((const unsigned char*)data)[at]);
            if (j < cols - 1) fputc(',', f);
        }
        fputc('\n', f);
    }
    fclose(f);
    return KGMT_OK;
}

int kgmt_dump_csv(kgmt_ctx* ctx, const char* dir) {
    if (!ctx || !dir) return KGMT_ERR_INVALID;
    struct Item { int id; const char* name; int kind; int cols; };
    const Item items[13] = {
        {KGMT_ARR_SAMPLES, "samples.csv", 0, 7}, {KGMT_ARR_UNEXPLORED, "unexploredSamples.csv", 0, 7},
        {KGMT_ARR_PARENT, "parentRelations.csv", 1, 1}, {KGMT_ARR_U_PARENT, "uParentIdx.csv", 1, 1},
        {KGMT_ARR_G, "G.csv", 2, 1}, {KGMT_ARR_R2AVAIL, "R2Avail.csv", 1, 1}, {KGMT_ARR_R1AVAIL, "R1Avail.csv", 1, 1},
        {KGMT_ARR_R1VALID, "R1Valid.csv", 1, 1}, {KGMT_ARR_R2VALID, "R2Valid.csv", 1, 1},
        {KGMT_ARR_R1INVALID, "R1Invalid.csv", 1, 1}, {KGMT_ARR_R2INVALID, "R2Invalid.csv", 1, 1},
        {KGMT_ARR_R1SCORE, "R1Score.csv", 0, 1}, {KGMT_ARR_R1, "R1.csv", 1, 1}};
    std::vector<unsigned char> buf;
    for (const Item& it : items) {
        const size_t bytes = kgmt_array_bytes(ctx, it.id);
        buf.assign(bytes, 0);
        const bool candId = (it.id == KGMT_ARR_UNEXPLORED || it.id == KGMT_ARR_U_PARENT);
        if (candId && !ctx->recordAllocated) {
            if (it.id == KGMT_ARR_U_PARENT) memset(buf.data(), 0xFF, bytes);    /* never-written slots: -1 / 0 as the ctor leaves them */
        } else {
            int rc = kgmt_export(ctx, it.id, buf.data(), bytes);
            if (rc) return rc;
        }
        const size_t elem = it.kind == 2 ? 1 : 4;
        const size_t rows = bytes / (elem * it.cols);
        int rc = write_csv(ctx, std::string(dir) + "/" + it.name, buf.data(), it.kind, rows, it.cols);
        if (rc) return rc;
    }
    return KGMT_OK;
}

/* ---- introspection -------------------------------------------------------------------------- */
int kgmt_tree_size(const kgmt_ctx* ctx) { return ctx ? ctx->hState->treeSize : 0; }
float kgmt_cost_to_goal(const kgmt_ctx* ctx) { return ctx ? ctx->hState->costToGoal : 0.f; }
float kgmt_r1_size(const kgmt_ctx* ctx) { return ctx ? ctx->R1Size : 0.f; }
float kgmt_r2_size(const kgmt_ctx* ctx) { return ctx ? ctx->R2Size : 0.f; }
void* kgmt_stream(const kgmt_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
long long kgmt_launch_count(const kgmt_ctx* ctx) { return ctx ? ctx->launches : 0; }

/* bounds-checked build (make check): {first failing site id, failures, offending value, its limit}; resets the record.
 * Returns KGMT_ERR_STATE from the product build, which compiles the checks away. */
int kgmt_debug_checks(kgmt_ctx* ctx, int* out4) {
    if (!ctx || !out4) return KGMT_ERR_INVALID;
#ifdef KGMT_BOUNDS_CHECK
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemcpyFromSymbol(out4, g_kgmtCheck, 16));
    const int z[4] = {0, 0, 0, 0};
    CU(cudaMemcpyToSymbol(g_kgmtCheck, z, 16));
    return KGMT_OK;
#else
    out4[0] = out4[1] = out4[2] = out4[3] = 0;
    return fail(ctx, KGMT_ERR_STATE, "this build has no bounds checks (make -C cudasbmp_b200/csrc check)");
#endif
}

int kgmt_work_counters(kgmt_ctx* ctx, unsigned long long* out4) {
    if (!ctx || !out4) return KGMT_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    int rc = fetch_state(ctx);
    if (rc) return rc;
    out4[0] = ctx->hState->stepsDone; out4[1] = ctx->hState->pairsTested;
    out4[2] = (unsigned long long)ctx->hState->expansions; out4[3] = 0;
    return KGMT_OK;
}

/* per-iteration device timestamps of the last plan (diagnostics): out[i] = {ns since the first logged iteration
 * ended... raw globaltimer ns, candidates, accepted}; returns the number of rows written (<= max_rows) */
int kgmt_iteration_log(kgmt_ctx* ctx, int enable, unsigned long long* out8, int max_rows) {
    if (!ctx) return KGMT_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    if (enable && !ctx->iterLog) {
        CU(cudaMalloc(&ctx->iterLog, KGMT_ITERLOG_BYTES));
        CU(cudaMemset(ctx->iterLog, 0, KGMT_ITERLOG_BYTES));
    }
    if (!out8 || max_rows <= 0 || !ctx->iterLog) return 0;
    int rc = fetch_state(ctx);
    if (rc) return rc;
    const int n = std::min(std::min(ctx->hState->iterationsDone, 255), max_rows);
    CU(cudaMemcpy(out8, ctx->iterLog, (size_t)n * 64, cudaMemcpyDeviceToHost));
    return n;
}

/* what the planner resolved for this obstacle set: collision back end, cull grid, shared memory, grid */
int kgmt_get_config(const kgmt_ctx* ctx, int* out8) {
    if (!ctx || !out8) return KGMT_ERR_INVALID;
    out8[0] = ctx->col; out8[1] = ctx->cullC; out8[2] = ctx->numItems; out8[3] = (int)ctx->smemBytes;
    out8[4] = ctx->gridLoop; out8[5] = ctx->numSMs; out8[6] = ctx->useHist; out8[7] = ctx->K;
    return KGMT_OK;
}

}  /* extern "C" */

#include "kgmt_comm.inl"
