/* experiments/lane_refill/expand_group.cuh — NOT part of the product, NOT compiled.
 *
 * Phase A with G chunks of candidates in flight per warp and freed lanes refilled from parked edges, as measured in
 * round 2 (gpurun pass r02a) and then removed from cudasbmp_b200/csrc (README.md next to this file has the numbers).
 * Three pieces, in the order they sat in the product: EdgeRun (kgmt_device.cuh), EdgeSlot + expand_group
 * (kgmt_kernels.cuh), and the group branch of run_plan's phase A loop.
 */
/* The same edge as a RESUMABLE object: begin() + one step() per call.  Phase A keeps several chunks of candidates in
 * flight per warp and hands a lane whose edge has ended (collision, workspace bounds, or all steps done) the next
 * pending edge, so the lanes freed by early exits do not idle until the slowest edge of the chunk is done.
 * step() is the loop body of propagate_edge, operation for operation (same intrinsics, same order): the results are
 * bit-identical whichever lane runs the edge (tests: every plan == its chunks_in_flight = 1 twin). */
template <class Collide>
struct EdgeRun {
    float x, y, th, v, a, dt, tanS;
    typename Collide::Cursor cur;
    int i;
    __device__ __forceinline__ void begin(const float4 s, float a_, float dt_, float tanS_, const Collide& col) {
        x = s.x; y = s.y; th = s.z; v = s.w; a = a_; dt = dt_; tanS = tanS_; i = 0;
        cur = col.start(x, y);
    }
    /* 0: the edge goes on; 1: ended valid (all numDisc steps); 2: ended invalid (statePropagator.cu:42-45 or :61-64) */
    __device__ __forceinline__ int step(const DynParams& p, const Collide& col) {
        const float px = x, py = y;
        float sn, cs;
        sincosf(th, &sn, &cs);
        x = __fmaf_rn(dt, __fmul_rn(v, cs), x);
        y = __fmaf_rn(dt, __fmul_rn(v, sn), y);
        if (x <= 0.0f || x >= p.W || y <= 0.0f || y >= p.H) return 2;
        const float vl = (p.L == 1.0f) ? v : __fdiv_rn(v, p.L);
        th = __fmaf_rn(dt, __fmul_rn(vl, tanS), th);
        v = __fmaf_rn(a, dt, v);
        const float bnx = (px > x) ? x : px, bxx = (px > x) ? px : x;
        const float bny = (py > y) ? y : py, bxy = (py > y) ? py : y;
        if (col.hit(cur, x, y, bnx, bny, bxx, bxy)) return 2;
        ++i;
        return i >= p.numDisc ? 1 : 0;
    }
};


/* ------------------------------------------- phase A: G chunks in flight per warp, lanes refilled ----
 * ~40 % of config 2's edges end early (collision / bounds), so with one candidate per lane for a whole chunk the
 * integration loop runs at ~21 of 32 lanes (profiles/r01j_*).  Here a warp takes up to G consecutive chunks:
 *   prologue   stage 2 + parent reads for all G chunks at full width (chunk_setup); the edges of chunks 1..G-1 are
 *              parked in this warp's shared-memory slots (state, a, dt, tan(steering): 32 B each)
 *   run loop   lane l starts on candidate l of chunk 0; a lane whose edge ends writes the exit state into the edge's
 *              slot (its own registers for chunk 0) and takes the next parked edge (ballot + popcount hand-out, in
 *              slot order); the loop ends when no edge is parked or running
 *   epilogue   stage 5a for the G chunks at full width (chunk_finish), each lane for the candidates it set up
 * Who runs an edge does not change its arithmetic (EdgeRun), and chunk_finish sees the same (state, valid) per
 * candidate in the same chunk order, so trees, maps and ballots are bit-identical to G = 1. */
struct EdgeSlot { float4 s; float4 c; };      /* parked: state | (a, dt, tanS, live);  done: exit state | (valid, ...) */

template <class Collide, bool RECORD, int G>
__device__ __forceinline__ void expand_group(const KArgs& A, const IterView& it, const DynParams& dyn, const Collide& col,
                                             int c0, int nch, int lane, int* hV, int* hI, bool& scoresOk, EdgeSlot* slots) {
    ChunkCand cc[G];
    float dts[G], tans[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
        cc[g] = ChunkCand{};
        if (g < nch) cc[g] = chunk_setup(A, it, c0 + g, lane);
        dts[g] = __fdiv_rn(cc[g].u.duration, (float)dyn.numDisc);
        tans[g] = tanf(cc[g].u.steering);
        if (g >= 1 && g < nch)
            slots[(g - 1) * 32 + lane] = EdgeSlot{cc[g].x, make_float4(cc[g].u.a, dts[g], tans[g], cc[g].live ? 1.0f : 0.0f)};
    }
    __syncwarp();
    EdgeRun<Collide> run;
    run.begin(cc[0].x, cc[0].u.a, dts[0], tans[0], col);
    bool running = cc[0].live;
    int tag = -1;                                   /* -1: this lane's own chunk-0 candidate, else the slot it runs */
    int next = 0;
    const int end = (nch - 1) * 32;
    bool val0 = false;
    unsigned steps = 0u, pairs = 0u;
    const unsigned below = (1u << lane) - 1u;
    for (;;) {
        const unsigned idle = __ballot_sync(0xffffffffu, !running);
        if (idle != 0u && next < end) {
            const int e = next + __popc(idle & below);
            if (!running && e < end) {
                const EdgeSlot sl = slots[e];
                if (sl.c.w != 0.0f) { run.begin(sl.s, sl.c.x, sl.c.y, sl.c.z, col); running = true; tag = e; }
            }
            next = min(next + __popc(idle), end);
        }
        if (!__any_sync(0xffffffffu, running)) { if (next >= end) break; continue; }
        if (running) {
            const int r = run.step(dyn, col);
            if (RECORD) steps += 1u;
            if (r != 0) {
                running = false;
                if (RECORD) pairs += run.cur.pairs;
                const float4 out = make_float4(run.x, run.y, run.th, run.v);
                if (tag < 0) { cc[0].x = out; val0 = (r == 1); }
                else { slots[tag].s = out; slots[tag].c.x = (r == 1) ? 1.0f : 0.0f; }
            }
        }
    }
    __syncwarp();
    if (RECORD) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { steps += __shfl_xor_sync(0xffffffffu, steps, o); pairs += __shfl_xor_sync(0xffffffffu, pairs, o); }
        if (lane == 0) { atomicAdd(&A.st->stepsDone, (unsigned long long)steps); atomicAdd(&A.st->pairsTested, (unsigned long long)pairs); }
    }
    cc[0].valid = val0;
    chunk_finish<RECORD, false>(A, it, cc[0], c0, lane, hV, hI, scoresOk);
#pragma unroll
    for (int g = 1; g < G; ++g) {
        if (g < nch) {
            if (cc[g].live) {
                const EdgeSlot sl = slots[(g - 1) * 32 + lane];
                cc[g].x = sl.s; cc[g].valid = sl.c.x != 0.0f;
            }
            chunk_finish<RECORD, false>(A, it, cc[g], c0 + g, lane, hV, hI, scoresOk);
        }
    }
    __syncwarp();                                   /* the slots may be re-used by the next group */
}


/* ---- the branch of run_plan (phase A, COL_GRID_SMEM) that took groups of chunks: ---- */
#if 0
            if (COL == COL_GRID_SMEM && G > 1) {
                /* groups of up to G consecutive chunks per warp with lane refill (expand_group).  Group sizes shrink as
                 * the iteration drains (guided self-scheduling) so the tail stays one chunk long; the first group is
                 * taken by position, the ticket (which starts at totalWarps) is read relative to it. */
                EdgeSlot* mySlots = cs.slots + warp * ((G - 1) * 32);
                const int g0 = max(1, min(G, it.numChunks / totalWarps));
                const int off = totalWarps * (g0 - 1);
                int g = g0;
                c = gw * g0;
                while (c < it.numChunks) {
                    const int nch = min(g, it.numChunks - c);
                    const int gn = max(1, min(G, (it.numChunks - (c + nch)) / totalWarps));      /* size of the next group */
                    if (lane == 0) t = (int)atomicAdd(ticket, (unsigned)gn);
                    expand_group<CollideGrid, RECORD, G>(A, it, dyn, colGridS, c, nch, lane, hV, hI, scoresOk, mySlots);
                    c = __shfl_sync(0xffffffffu, t, 0) + off;
                    g = gn;
                }
            } else
#endif
