/* kgmt_comm.inl — the multi-GPU entry points of the C ABI (include/kgmt_c.h "multi-GPU"), included by kgmt_capi.cu.
 *
 * One process per GPU on one NVSwitch box.  The reference has no multi-GPU mode; these calls let a C++
 * host program — e.g. the reference's demos/main.cu:30,62 with one process per GPU — use the modes BASELINE.json names
 * without Python:
 *     kgmt_comm_init            NCCL communicator + exchange of the peer-memory handles + attach + barrier
 *     kgmt_plan_batch_sharded   config 4: independent queries sharded over the ranks, results all-gathered (NCCL)
 *     kgmt_plan_portfolio       same query, one seed per rank: race over peer memory, winner's result and path
 *                               broadcast to every rank (NCCL over NVLink: "first-solution broadcast")
 *     kgmt_expand_sharded       config 5: one iteration's candidates split over the ranks; exchange fused into the
 *                               persistent kernel (peer memory), or by the multi-launch peer sequence, or by NCCL
 *                               all-gather / all-reduce (the honest baseline, with its host round trip for the counts)
 *     kgmt_plan_sharded         whole plans with sharded iterations (kgmt_peer_plan)
 * NCCL is loaded with dlopen("libnccl.so.2") on first use, so single-GPU users of the library do not need it installed;
 * the types come from <nccl.h>.
 */
#include <dlfcn.h>
#include <nccl.h>

struct kgmt_nccl_api {
    void* lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    const char* err = nullptr;
};

static kgmt_nccl_api* nccl_api() {
    static kgmt_nccl_api api;
    if (api.lib || api.err) return &api;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) { api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (api.lib) break; }
    if (!api.lib) { api.err = "libnccl.so.2 not found (dlopen)"; return &api; }
#define KGMT_NCCL_SYM(field, name) api.field = (decltype(api.field))dlsym(api.lib, name); if (!api.field) { api.err = "missing NCCL symbol " name; }
    KGMT_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    KGMT_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    KGMT_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    KGMT_NCCL_SYM(AllGather, "ncclAllGather")
    KGMT_NCCL_SYM(AllReduce, "ncclAllReduce")
    KGMT_NCCL_SYM(Broadcast, "ncclBroadcast")
    KGMT_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef KGMT_NCCL_SYM
    return &api;
}

struct kgmt_comm_state {
    ncclComm_t comm = nullptr;
    int rank = -1, world = 0;
    /* exchange buffers of the NCCL path and small staging */
    int* delta = nullptr;                     /* kgmt_shard_delta_ints ints */
    int* counts = nullptr; int* hCounts = nullptr;          /* [world] device, pinned */
    unsigned char* send = nullptr; unsigned char* recv = nullptr; size_t capRows = 0;
    unsigned char* small = nullptr; unsigned char* hSmall = nullptr; size_t smallBytes = 0;   /* device / pinned scratch */
};

#define NC(call)                                                                                              \
    do {                                                                                                      \
        ncclResult_t r_ = (call);                                                                             \
        if (r_ != ncclSuccess) return fail(ctx, KGMT_ERR_COMM, "%s:%d %s: %s", __FILE__, __LINE__, #call, nccl_api()->GetErrorString(r_)); \
    } while (0)

static int comm_small(kgmt_ctx* ctx, size_t bytes) {
    kgmt_comm_state* cs = ctx->comm;
    if (bytes <= cs->smallBytes) return KGMT_OK;
    if (cs->small) cudaFree(cs->small);
    if (cs->hSmall) cudaFreeHost(cs->hSmall);
    cs->small = nullptr; cs->hSmall = nullptr; cs->smallBytes = 0;
    CU(cudaMalloc(&cs->small, bytes));
    CU(cudaHostAlloc(&cs->hSmall, bytes, cudaHostAllocDefault));
    cs->smallBytes = bytes;
    return KGMT_OK;
}

extern "C" {

int kgmt_comm_unique_id(void* out_id128) {
    if (!out_id128) return KGMT_ERR_INVALID;
    kgmt_nccl_api* n = nccl_api();
    if (n->err) return KGMT_ERR_COMM;
    ncclUniqueId id;
    if (n->GetUniqueId(&id) != ncclSuccess) return KGMT_ERR_COMM;
    memcpy(out_id128, &id, sizeof(id));
    return KGMT_OK;
}

int kgmt_comm_barrier(kgmt_ctx* ctx) {
    if (!ctx || !ctx->comm) return fail(ctx, KGMT_ERR_STATE, "no communicator (kgmt_comm_init)");
    kgmt_comm_state* cs = ctx->comm;
    CU(cudaSetDevice(ctx->device));
    int rc = comm_small(ctx, 64);
    if (rc) return rc;
    NC(nccl_api()->AllReduce(cs->small, cs->small, 1, ncclInt32, ncclSum, cs->comm, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return KGMT_OK;
}

int kgmt_comm_destroy(kgmt_ctx* ctx) {
    if (!ctx) return KGMT_ERR_INVALID;
    kgmt_comm_state* cs = ctx->comm;
    if (!cs) return KGMT_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    kgmt_peer_detach(ctx);
    if (cs->comm) nccl_api()->CommDestroy(cs->comm);
    cudaFree(cs->delta); cudaFree(cs->counts); cudaFree(cs->send); cudaFree(cs->recv); cudaFree(cs->small);
    if (cs->hCounts) cudaFreeHost(cs->hCounts);
    if (cs->hSmall) cudaFreeHost(cs->hSmall);
    delete cs;
    ctx->comm = nullptr;
    return KGMT_OK;
}

int kgmt_comm_init(kgmt_ctx* ctx, int rank, int world, const void* nccl_unique_id128) {
    if (!ctx || !nccl_unique_id128 || world < 1 || world > PEER_MAX || rank < 0 || rank >= world)
        return fail(ctx, KGMT_ERR_INVALID, "bad communicator arguments");
    kgmt_nccl_api* n = nccl_api();
    if (n->err) return fail(ctx, KGMT_ERR_COMM, "NCCL unavailable: %s", n->err);
    CU(cudaSetDevice(ctx->device));
    kgmt_comm_destroy(ctx);
    kgmt_comm_state* cs = new (std::nothrow) kgmt_comm_state();
    if (!cs) return KGMT_ERR_NOMEM;
    ctx->comm = cs;
    cs->rank = rank; cs->world = world;
    ncclUniqueId id;
    memcpy(&id, nccl_unique_id128, sizeof(id));
    NC(n->CommInitRank(&cs->comm, world, id, rank));
    CU(cudaMalloc(&cs->delta, kgmt_shard_delta_ints(ctx) * 4));
    CU(cudaMemset(cs->delta, 0, kgmt_shard_delta_ints(ctx) * 4));
    CU(cudaMalloc(&cs->counts, (size_t)world * 4));
    CU(cudaHostAlloc(&cs->hCounts, (size_t)world * 4, cudaHostAllocDefault));
    /* peer memory: every rank's cudaIpc handles to every rank (one NCCL all-gather), then map them */
    const size_t hb = kgmt_peer_handle_bytes();
    int rc = comm_small(ctx, hb * (size_t)world + 64);
    if (rc) return rc;
    rc = kgmt_peer_export(ctx, cs->hSmall + hb * (size_t)rank, hb);
    if (rc) return rc;
    CU(cudaMemcpyAsync(cs->small + hb * (size_t)rank, cs->hSmall + hb * (size_t)rank, hb, cudaMemcpyHostToDevice, ctx->stream));
    NC(n->AllGather(cs->small + hb * (size_t)rank, cs->small, hb, ncclChar, cs->comm, ctx->stream));
    CU(cudaMemcpyAsync(cs->hSmall, cs->small, hb * (size_t)world, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    rc = kgmt_peer_attach(ctx, rank, world, cs->hSmall);
    if (rc) return rc;
    return kgmt_comm_barrier(ctx);          /* every rank has mapped (and cleared its mailboxes) before anyone starts an exchange */
}

int kgmt_comm_rank(const kgmt_ctx* ctx) { return (ctx && ctx->comm) ? ctx->comm->rank : -1; }
int kgmt_comm_world(const kgmt_ctx* ctx) { return (ctx && ctx->comm) ? ctx->comm->world : 0; }

/* ---- config 5: one sharded iteration -------------------------------------------------------------------------- */
static int expand_sharded_nccl(kgmt_ctx* ctx, kgmt_iter_stats* out, float* ms3) {
    kgmt_comm_state* cs = ctx->comm;
    kgmt_nccl_api* n = nccl_api();
    const int world = cs->world, rank = cs->rank;
    cudaStream_t s = ctx->stream;
    cudaEvent_t ev[6] = {};
    if (ms3) for (auto& e : ev) CU(cudaEventCreate(&e));
    if (ms3) CU(cudaEventRecord(ev[0], s));
    kgmt_shard_info info;
    int rc = kgmt_shard_expand(ctx, rank, world, cs->delta, &info);      /* ends with the count on the host (one sync) */
    if (rc) return rc;
    if (ms3) CU(cudaEventRecord(ev[1], s));
    cs->hCounts[rank] = info.accepted_local;
    CU(cudaMemcpyAsync(cs->counts + rank, cs->hCounts + rank, 4, cudaMemcpyHostToDevice, s));
    NC(n->AllGather(cs->counts + rank, cs->counts, 1, ncclInt32, cs->comm, s));
    CU(cudaMemcpyAsync(cs->hCounts, cs->counts, (size_t)world * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    int mx = 0;
    for (int g = 0; g < world; ++g) mx = std::max(mx, cs->hCounts[g]);
    const int cap = std::max(4, (mx + 3) & ~3);
    if ((size_t)cap > cs->capRows) {
        cudaFree(cs->send); cudaFree(cs->recv); cs->send = cs->recv = nullptr; cs->capRows = 0;
        const size_t c = std::max<size_t>((size_t)cap, 2 * cs->capRows);
        CU(cudaMalloc(&cs->send, c * 36)); CU(cudaMalloc(&cs->recv, c * 36 * (size_t)world));
        cs->capRows = c;
    }
    if (ms3) CU(cudaEventRecord(ev[2], s));
    rc = kgmt_shard_pack(ctx, cs->send, cap);
    if (rc) return rc;
    if (ms3) CU(cudaEventRecord(ev[3], s));
    NC(n->AllGather(cs->send, cs->recv, (size_t)cap * 36, ncclChar, cs->comm, s));
    NC(n->AllReduce(cs->delta, cs->delta, kgmt_shard_delta_ints(ctx), ncclInt32, ncclSum, cs->comm, s));
    if (ms3) CU(cudaEventRecord(ev[4], s));
    rc = kgmt_shard_commit(ctx, cs->recv, cap, cs->hCounts, world, cs->delta, out);
    if (rc) return rc;
    if (ms3) {
        CU(cudaEventRecord(ev[5], s));
        CU(cudaStreamSynchronize(s));
        float a = 0, b = 0, c = 0, d = 0, e = 0;
        cudaEventElapsedTime(&a, ev[0], ev[1]); cudaEventElapsedTime(&b, ev[1], ev[2]); cudaEventElapsedTime(&c, ev[2], ev[3]);
        cudaEventElapsedTime(&d, ev[3], ev[4]); cudaEventElapsedTime(&e, ev[4], ev[5]);
        ms3[0] = a + c + e;                       /* compute: expand + pack + commit */
        ms3[1] = b + d;                           /* exchange: count all-gather (+ host round trip), row all-gather, delta all-reduce */
        ms3[2] = (float)((size_t)world * cap * 36 + kgmt_shard_delta_ints(ctx) * 4 + 4 * (size_t)world);
        for (auto& e2 : ev) cudaEventDestroy(e2);
    }
    return KGMT_OK;
}

int kgmt_expand_sharded(kgmt_ctx* ctx, int exchange, kgmt_iter_stats* out, float* ms3) {
    if (!ctx || !ctx->comm) return fail(ctx, KGMT_ERR_STATE, "no communicator (kgmt_comm_init)");
    if (!ctx->begun) return fail(ctx, KGMT_ERR_STATE, "kgmt_expand_sharded before kgmt_begin / kgmt_seed_frontier");
    CU(cudaSetDevice(ctx->device));
    if (exchange == KGMT_EXCHANGE_NCCL) return expand_sharded_nccl(ctx, out, ms3);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ms3) { CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1)); CU(cudaEventRecord(e0, ctx->stream)); }
    int rc;
    if (exchange == KGMT_EXCHANGE_PEER_LAUNCHES) {
        rc = kgmt_peer_expand_begin(ctx);
        if (rc) return rc;
        if (ms3) CU(cudaEventRecord(e1, ctx->stream));
        rc = kgmt_peer_expand_end(ctx, out);
    } else if (exchange == KGMT_EXCHANGE_FUSED) {
        if (ctx->hState->stop == STOP_RUNNING) {
            const int before = ctx->hState->iterationsDone;
            rc = launch_fused(ctx, 1);
            if (rc) return rc;
            if (ms3) CU(cudaEventRecord(e1, ctx->stream));
            rc = finish_fused(ctx, before);
        } else {
            if (ms3) CU(cudaEventRecord(e1, ctx->stream));
            rc = KGMT_OK;
        }
        if (!rc && out) {
            const DevState& s = *ctx->hState;
            out->iteration = s.lastItr; out->mode = s.lastMode; out->children = s.lastChildren; out->frontier = s.lastFrontier;
            out->candidates = s.lastM; out->accepted = s.lastAccepted; out->tree_size = s.treeSize; out->stop = s.stop;
            out->cost_to_goal = s.costToGoal; out->goal_index = s.goalIdx;
        }
    } else {
        return fail(ctx, KGMT_ERR_INVALID, "unknown exchange %d", exchange);
    }
    if (ms3) {
        if (!rc) { CU(cudaStreamSynchronize(ctx->stream)); float t = 0; cudaEventElapsedTime(&t, e0, e1); ms3[0] = t; ms3[1] = 0.f; ms3[2] = 0.f; }
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    return rc;
}

int kgmt_plan_sharded(kgmt_ctx* ctx, const float* initial7, const float* goal7, kgmt_result* out) {
    if (!ctx || !ctx->comm) return fail(ctx, KGMT_ERR_STATE, "no communicator (kgmt_comm_init)");
    return kgmt_peer_plan(ctx, initial7, goal7, out);
}

/* ---- config 4: independent queries sharded over the ranks ------------------------------------------------------- */
int kgmt_plan_batch_sharded(kgmt_ctx* ctx, const float* h_inits7, const float* h_goals7, const uint32_t* h_seeds, int Q,
                            int cluster_size, kgmt_result* out_all, float* device_ms_max) {
    if (!ctx || !ctx->comm) return fail(ctx, KGMT_ERR_STATE, "no communicator (kgmt_comm_init)");
    if (!h_inits7 || !h_goals7 || !h_seeds || Q < 1 || !out_all) return fail(ctx, KGMT_ERR_INVALID, "bad batch arguments");
    kgmt_comm_state* cs = ctx->comm;
    kgmt_nccl_api* n = nccl_api();
    CU(cudaSetDevice(ctx->device));
    const int world = cs->world, rank = cs->rank;
    const int base = Q / world, rem = Q % world;
    const int lo = rank * base + std::min(rank, rem), cnt = base + (rank < rem ? 1 : 0);
    const int per = base + (rem ? 1 : 0);                                   /* padded shard length */
    std::vector<kgmt_result> mine((size_t)per);
    memset(mine.data(), 0, mine.size() * sizeof(kgmt_result));
    float ms = 0.f;
    if (cnt > 0) {
        int rc = kgmt_plan_batch(ctx, h_inits7 + (size_t)lo * 7, h_goals7 + (size_t)lo * 7, h_seeds + lo, cnt, cluster_size,
                                 mine.data(), nullptr, 0, nullptr, &ms);
        if (rc < 0) return rc;
    }
    /* results: one padded all-gather of the fixed-size rows (+ the device time in a trailing float per rank) */
    const size_t rowB = sizeof(kgmt_result), shardB = (size_t)per * rowB + 16;
    int rc = comm_small(ctx, shardB * (size_t)world + shardB);
    if (rc) return rc;
    unsigned char* hs = cs->hSmall + shardB * (size_t)world;               /* this rank's shard staged after the gathered table */
    memcpy(hs, mine.data(), (size_t)per * rowB);
    memcpy(hs + (size_t)per * rowB, &ms, 4);
    unsigned char* ds = cs->small + shardB * (size_t)world;
    CU(cudaMemcpyAsync(ds, hs, shardB, cudaMemcpyHostToDevice, ctx->stream));
    NC(n->AllGather(ds, cs->small, shardB, ncclChar, cs->comm, ctx->stream));
    CU(cudaMemcpyAsync(cs->hSmall, cs->small, shardB * (size_t)world, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    float mx = 0.f;
    for (int g = 0; g < world; ++g) {
        const int glo = g * base + std::min(g, rem), gcnt = base + (g < rem ? 1 : 0);
        const unsigned char* src = cs->hSmall + shardB * (size_t)g;
        memcpy(out_all + glo, src, (size_t)gcnt * rowB);
        float gms; memcpy(&gms, src + (size_t)per * rowB, 4);
        mx = std::max(mx, gms);
    }
    if (device_ms_max) *device_ms_max = mx;
    return KGMT_OK;
}

/* ---- portfolio: same query, one seed per rank, first solution wins ------------------------------------------------ */
int kgmt_plan_portfolio(kgmt_ctx* ctx, const float* initial7, const float* goal7, uint32_t base_seed, int race_id,
                        kgmt_result* out_winner, int* winner_rank, float* h_path7, int max_rows, int* path_len) {
    if (!ctx || !ctx->comm) return fail(ctx, KGMT_ERR_STATE, "no communicator (kgmt_comm_init)");
    kgmt_comm_state* cs = ctx->comm;
    kgmt_nccl_api* n = nccl_api();
    CU(cudaSetDevice(ctx->device));
    const int rank = cs->rank;
    int rc = kgmt_set_seed(ctx, base_seed + (uint32_t)rank);
    if (rc) return rc;
    kgmt_result mine;
    rc = kgmt_peer_race(ctx, initial7, goal7, race_id, &mine);             /* ONE launch; a peer's win stops it through peer memory */
    if (rc) return rc;
    /* winner = lowest cost among the ranks that solved, ties to the lowest rank: one 8-byte all-reduce(MIN) */
    const size_t pathB = (size_t)std::max(max_rows, 0) * 28;
    rc = comm_small(ctx, 64 + sizeof(kgmt_result) + 16 + pathB);
    if (rc) return rc;
    unsigned long long key = ~0ull;
    if (mine.stop == KGMT_SOLVED) {
        unsigned bits; memcpy(&bits, &mine.cost_to_goal, 4);
        key = ((unsigned long long)bits << 8) | (unsigned)rank;
    }
    memcpy(cs->hSmall, &key, 8);
    CU(cudaMemcpyAsync(cs->small, cs->hSmall, 8, cudaMemcpyHostToDevice, ctx->stream));
    NC(n->AllReduce(cs->small, cs->small, 1, ncclUint64, ncclMin, cs->comm, ctx->stream));
    CU(cudaMemcpyAsync(cs->hSmall, cs->small, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    memcpy(&key, cs->hSmall, 8);
    const int win = (key == ~0ull) ? -1 : (int)(key & 0xFF);
    if (winner_rank) *winner_rank = win;
    if (path_len) *path_len = 0;
    if (win < 0) { if (out_winner) *out_winner = mine; return KGMT_OK; }
    /* the winner's result block, path length and path rows to every rank */
    unsigned char* hb = cs->hSmall + 64;
    unsigned char* db = cs->small + 64;
    const size_t msgB = sizeof(kgmt_result) + 16 + pathB;
    if (rank == win) {
        int len = 0;
        if (max_rows > 0 && h_path7) { len = kgmt_extract_path(ctx, -1, (float*)(hb + sizeof(kgmt_result) + 16), max_rows); if (len < 0) return len; }
        memcpy(hb, &mine, sizeof(kgmt_result));
        memcpy(hb + sizeof(kgmt_result), &len, 4);
        CU(cudaMemcpyAsync(db, hb, msgB, cudaMemcpyHostToDevice, ctx->stream));
    }
    NC(n->Broadcast(db, db, msgB, ncclChar, win, cs->comm, ctx->stream));
    CU(cudaMemcpyAsync(hb, db, msgB, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (out_winner) memcpy(out_winner, hb, sizeof(kgmt_result));
    int len = 0;
    memcpy(&len, hb + sizeof(kgmt_result), 4);
    if (path_len) *path_len = len;
    if (h_path7 && max_rows > 0) memcpy(h_path7, hb + sizeof(kgmt_result) + 16, (size_t)std::min(len, max_rows) * 28);
    return KGMT_OK;
}

}  /* extern "C" */
