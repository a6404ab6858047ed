/* kgmt_device.cuh — per-candidate device arithmetic of the KGMT expansion step.
 *
 * Written from scratch for sm_100a.  Every function names the reference lines
 * whose RESULT it reproduces (paths relative to the reference tree).  The
 * arithmetic is pinned with explicit round-to-nearest intrinsics so that the
 * FMA contractions are the ones nvcc applies to the reference itself on
 * sm_100a (read off its PTX; DESIGN.md "numerical contract"), independent of
 * compiler flags.  Trigonometry is libdevice's accurate sinf/cosf/tanf — never
 * compile this file with -use_fast_math.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace kgmt {

/* ------------------------------------------------------------- bounds-checked build --
 * make -C cudasbmp_b200/csrc check  (-DKGMT_BOUNDS_CHECK -> libkgmt_b200_check.so): every index the kernels form into the
 * tree, the staging rows, the ballots, the region maps, the cell lists and the exchange slabs is range-checked on the
 * device; the first failing site and the number of failures are kept in g_kgmtCheck and read with kgmt_debug_checks.
 * (compute-sanitizer is refused on the GPU pool this was developed on: profiles/r02b_sanitizer_refused.txt.)
 * The product build compiles the checks away. */
#ifdef KGMT_BOUNDS_CHECK
__device__ int g_kgmtCheck[4];            /* [0] first failing site id, [1] failures, [2] offending value, [3] its limit */
__device__ __forceinline__ void kgmt_check_fail(int site, long long v, long long lim) {
    if (atomicAdd(&g_kgmtCheck[1], 1) == 0) { g_kgmtCheck[0] = site; g_kgmtCheck[2] = (int)v; g_kgmtCheck[3] = (int)lim; }
}
#define KGMT_CHECK_RANGE(site, v, lim) do { if ((long long)(v) < 0 || (long long)(v) >= (long long)(lim)) kgmt_check_fail(site, (long long)(v), (long long)(lim)); } while (0)
#else
#define KGMT_CHECK_RANGE(site, v, lim) do { } while (0)
#endif

/* ------------------------------------------------------------------ Philox4x32-10 --
 * cuRAND's counter-based generator (/usr/local/cuda/include/curand_philox4x32_x.h:88-190),
 * used statelessly: candidate slot s of iteration itr draws its four uniforms
 * from Philox(ctr = (0,0,s,0), key = (seed+itr, 0)) — exactly the stream
 * curand_init(seed+itr, s, 0) + 4x curand_uniform yields for
 * curandStatePhilox4_32_10_t, i.e. what the reference's initCurandStates
 * (src/planners/KGMT.cu:595-600) + draws (statePropagator.cu:17-19,
 * KGMT.cu:395) produce when its curandState is Philox.  No RNG state is stored
 * (the reference moves 96 B of XORWOW state per candidate, KGMT.cu:386,412). */
__device__ __forceinline__ uint4 philox4x32_10(uint32_t slot, uint32_t key0) {
    uint32_t c0 = 0u, c1 = 0u, c2 = slot, c3 = 0u;
    uint32_t k0 = key0, k1 = 0u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

/* curand_uniform: (0,1], /usr/local/cuda/include/curand_uniform.h:69-72 */
__device__ __forceinline__ float uniform01(uint32_t x) {
    return __fmaf_rn(__uint2float_rn(x), 2.3283064e-10f, 1.1641532e-10f);
}

struct Controls { float a, steering, duration, u3; };

/* control ranges of the car model (kgmt_params: accel / steer / duration min..max), pre-reduced to scale and offset.
 * The defaults {10, -5, 2*pi, -pi, 1, 0.05f} give the reference's literals of statePropagator.cu:17-19 bit for bit:
 *   u0 * 10.0f - 5.0f            -> one FFMA;
 *   u1 * 2.0f * M_PI - M_PI      -> DFMA((double)(u1 + u1), pi, -pi) == DFMA((double)u1, 2*pi, -pi): u1 + u1 and 2*pi
 *                                   are exact doublings, so both products are the same real number, rounded once;
 *   u2 * 1.0f + 0.05f            -> FFMA(u2, 1, 0.05f) == FADD(u2, 0.05f). */
struct CarRanges { float aScale, aLo; double sScale, sLo; float dScale, dLo; };

/* statePropagator.cu:17-19 (+ the accept uniform of KGMT.cu:395) */
__device__ __forceinline__ Controls sample_controls(uint32_t slot, uint32_t key0, const CarRanges& r) {
    const uint4 w = philox4x32_10(slot, key0);
    const float u0 = uniform01(w.x), u1 = uniform01(w.y), u2 = uniform01(w.z);
    Controls c;
    c.a = __fmaf_rn(u0, r.aScale, r.aLo);
    c.steering = __double2float_rn(__fma_rn((double)u1, r.sScale, r.sLo));
    c.duration = __fmaf_rn(u2, r.dScale, r.dLo);
    c.u3 = uniform01(w.w);
    return c;
}

/* --------------------------------------------------------------------- region index --
 * getR1 / getR2, src/planners/KGMT.cu:602-629 (== OccupancyGrid::getCellIndex,
 * src/occupancyMaps/OccupancyGrid.cu:12-19): IEEE division, truncation toward zero. */
struct RegionCell { int r1, cx, cy; };     /* R1 index (-1 = outside) and its column / row */
__device__ __forceinline__ RegionCell region_r1_cell(float x, float y, float R1Size, int N) {
    const int cx = __float2int_rz(__fdiv_rn(x, R1Size));
    const int cy = __float2int_rz(__fdiv_rn(y, R1Size));
    return RegionCell{(cx >= 0 && cx < N && cy >= 0 && cy < N) ? cy * N + cx : -1, cx, cy};
}
__device__ __forceinline__ int region_r1(float x, float y, float R1Size, int N) { return region_r1_cell(x, y, R1Size, N).r1; }
/* getR2 with the R1 column / row already known (r1 = cy * N + cx, so r1 / N and r1 % N are cy and cx) */
__device__ __forceinline__ int region_r2_cell(float x, float y, const RegionCell c, float R1Size, float R2Size, int n) {
    if (c.r1 < 0) return -1;
    const float lx = __fmaf_rn(-(float)c.cx, R1Size, x);      /* x - cx*R1Size, one FFMA as the reference's SASS */
    const float ly = __fmaf_rn(-(float)c.cy, R1Size, y);
    const int cx = __float2int_rz(__fdiv_rn(lx, R2Size));
    const int cy = __float2int_rz(__fdiv_rn(ly, R2Size));
    return (cx >= 0 && cx < n && cy >= 0 && cy < n) ? c.r1 * (n * n) + cy * n + cx : -1;
}
__device__ __forceinline__ int region_r2(float x, float y, int r1, float R1Size, int N, float R2Size, int n) {
    if (r1 < 0) return -1;
    const int cyR1 = r1 / N;
    return region_r2_cell(x, y, RegionCell{r1, r1 - cyR1 * N, cyR1}, R1Size, R2Size, n);
}

/* goal test, KGMT.cu:635-638: differences in float, squares/sqrt in double, '<' on the narrowed float.
 * The first line only skips work: dx*dx is exact in double and every later operation (+, sqrt, narrowing) is correctly
 * rounded, hence monotone, so dist >= |dx| and dist >= |dy| as floats — a difference of r or more decides 'false'. */
__device__ __forceinline__ bool in_goal(float x, float y, float gx, float gy, float r) {
    const float fx = __fsub_rn(x, gx), fy = __fsub_rn(y, gy);
    if (!(fabsf(fx) < r) || !(fabsf(fy) < r)) return false;
    const double dx = (double)fx, dy = (double)fy;
    const float dist = __double2float_rn(__dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))));
    return dist < r;
}

/* overlap of a step bounding box with one obstacle (minx,miny,maxx,maxy):
 * !isBroadPhaseValid, src/collisionCheck/collisionCheck.cu:6-14 (strict; touching is free) */
__device__ __forceinline__ bool aabb_overlap(float bnx, float bny, float bxx, float bxy, const float4 o) {
    return (bxx > o.x) & (o.z > bnx) & (bxy > o.y) & (o.w > bny);
}

/* --------------------------------------------------------------- collision back ends --
 * A back end answers "does this step bbox overlap any obstacle?" — the result of
 * isMotionValid (collisionCheck.cu:16-28) negated.  Both give the same answer.
 * start(x, y) is called once with the edge's first point, hit(...) once per step with the
 * new point and the step bbox. */

/* every obstacle, shared-memory resident (float4 per obstacle, one broadcast LDS.128 each) */
struct CollideSmemAll {
    const float4* obs; int K;
    struct Cursor { unsigned pairs; };      /* pairs: overlap tests executed (read only by the recording kernels; dead code otherwise) */
    __device__ __forceinline__ Cursor start(float, float) const { return Cursor{0u}; }
    __device__ __forceinline__ bool hit(Cursor& cur, float, float, float bnx, float bny, float bxx, float bxy) const {
        bool h = false;
        int k = 0;
        for (; k + 4 <= K; k += 4) {
            const float4 o0 = obs[k], o1 = obs[k + 1], o2 = obs[k + 2], o3 = obs[k + 3];
            h = aabb_overlap(bnx, bny, bxx, bxy, o0) | aabb_overlap(bnx, bny, bxx, bxy, o1) |
                aabb_overlap(bnx, bny, bxx, bxy, o2) | aabb_overlap(bnx, bny, bxx, bxy, o3);
            cur.pairs += 4u;
            if (h) return true;
        }
        for (; k < K; ++k) { h |= aabb_overlap(bnx, bny, bxx, bxy, obs[k]); cur.pairs += 1u; }
        return h;
    }
};

/* uniform-grid culled: only the obstacles registered in the cells the bbox touches.
 * cell(v) = clamp(floor(v*inv)) is monotone and obstacles are registered with the same
 * function, so every obstacle that can overlap the bbox shares a cell with it: the flag
 * is identical to the exhaustive test.  The bbox corners are the step's two end points,
 * so its cell range is the min/max of their cells (one conversion pair per step, the
 * other carried in the cursor); the cells of one grid row are contiguous in the CSR. */
struct CollideGrid {
    const int* cellStart;      /* [C*C+1] */
    const float4* items;       /* obstacle AABBs, grouped by cell */
    int C; float invX, invY;
    int nStart = 0, nItems = 0;   /* sizes of the two arrays (bounds-checked build only) */
    struct Cursor { int cx, cy; unsigned pairs; };
    __device__ __forceinline__ int cell(float v, float inv) const {
        return min(max(__float2int_rd(__fmul_rn(v, inv)), 0), C - 1);
    }
    /* cell of a point that passed the workspace test (v > 0, or NaN which converts to 0): no lower clamp needed */
    __device__ __forceinline__ int cell_in(float v, float inv) const { return min(__float2int_rd(__fmul_rn(v, inv)), C - 1); }
    __device__ __forceinline__ Cursor start(float x, float y) const { return Cursor{cell(x, invX), cell(y, invY), 0u}; }
    __device__ __forceinline__ bool hit(Cursor& cur, float x, float y, float bnx, float bny, float bxx, float bxy) const {
        const int cxn = cell_in(x, invX), cyn = cell_in(y, invY);
        const int cx0 = min(cur.cx, cxn), cx1 = max(cur.cx, cxn);
        const int cy0 = min(cur.cy, cyn), cy1 = max(cur.cy, cyn);
        cur.cx = cxn; cur.cy = cyn;
        bool h = false;
        for (int cy = cy0; cy <= cy1 && !h; ++cy) {
            const int row = cy * C;
            KGMT_CHECK_RANGE(101, row + cx0, nStart); KGMT_CHECK_RANGE(102, row + cx1 + 1, nStart);
            const int e = cellStart[row + cx1 + 1];
            int k = cellStart[row + cx0];
            /* four items per trip, read unconditionally: an entry past the end of this row's range is another cell's
             * obstacle (or one of the three never-overlapping entries that close the array), and a box that overlaps
             * the step bbox is a collision whichever cell it was filed under — over-reading cannot change the flag */
            for (; k < e && !h; k += 4) {
                KGMT_CHECK_RANGE(103, k, nItems); KGMT_CHECK_RANGE(104, k + 3, nItems);
                const float4 o0 = items[k], o1 = items[k + 1], o2 = items[k + 2], o3 = items[k + 3];
                h = aabb_overlap(bnx, bny, bxx, bxy, o0) | aabb_overlap(bnx, bny, bxx, bxy, o1) |
                    aabb_overlap(bnx, bny, bxx, bxy, o2) | aabb_overlap(bnx, bny, bxx, bxy, o3);
                cur.pairs += 4u;
            }
        }
        return h;
    }
};

/* the same walk with the CSR in SHARED memory, addressed by 32-bit shared-window addresses and explicit ld.shared:
 * with generic pointers the compiler re-derives the shared-memory base (cluster CTA id, window, histogram offset: eleven
 * instructions) in every trip of the row loop once registers get tight — 8 % of the kernel's warp instructions in the
 * capture profiles/r02s_*. */
__device__ __forceinline__ int lds_s32(uint32_t addr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
#ifndef KGMT_WALK_WIDTH
#define KGMT_WALK_WIDTH 3               /* items per trip of the shared-memory cell walk */
#endif
struct CollideGridS {
    uint32_t startAddr, itemAddr;      /* shared-window addresses of cellStart[C*C+1] and of the float4 items */
    int C; float invX, invY;
    int nStart = 0, nItems = 0;
    typedef CollideGrid::Cursor Cursor;
    __device__ __forceinline__ int cell(float v, float inv) const {
        return min(max(__float2int_rd(__fmul_rn(v, inv)), 0), C - 1);
    }
    /* cell of a point that passed the workspace test (v > 0, or NaN which converts to 0): no lower clamp needed */
    __device__ __forceinline__ int cell_in(float v, float inv) const { return min(__float2int_rd(__fmul_rn(v, inv)), C - 1); }
    __device__ __forceinline__ Cursor start(float x, float y) const { return Cursor{cell(x, invX), cell(y, invY), 0u}; }
    __device__ __forceinline__ bool hit(Cursor& cur, float x, float y, float bnx, float bny, float bxx, float bxy) const {
        const int cxn = cell_in(x, invX), cyn = cell_in(y, invY);
        const int cx0 = min(cur.cx, cxn), cx1 = max(cur.cx, cxn);
        const int cy0 = min(cur.cy, cyn), cy1 = max(cur.cy, cyn);
        cur.cx = cxn; cur.cy = cyn;
        bool h = false;
        for (int cy = cy0; cy <= cy1 && !h; ++cy) {
            const int row = cy * C;
            KGMT_CHECK_RANGE(101, row + cx0, nStart); KGMT_CHECK_RANGE(102, row + cx1 + 1, nStart);
            const int e = lds_s32(startAddr + 4u * (uint32_t)(row + cx1 + 1));
            int k = lds_s32(startAddr + 4u * (uint32_t)(row + cx0));
            for (; k < e && !h; k += KGMT_WALK_WIDTH) {     /* KGMT_WALK_WIDTH items per trip, read unconditionally (see CollideGrid) */
                KGMT_CHECK_RANGE(103, k, nItems); KGMT_CHECK_RANGE(104, k + KGMT_WALK_WIDTH - 1, nItems);
                const uint32_t a = itemAddr + 16u * (uint32_t)k;
                float4 o[KGMT_WALK_WIDTH];
#pragma unroll
                for (int j = 0; j < KGMT_WALK_WIDTH; ++j) o[j] = lds_f4(a + 16u * (uint32_t)j);
#pragma unroll
                for (int j = 0; j < KGMT_WALK_WIDTH; ++j) h |= aabb_overlap(bnx, bny, bxx, bxy, o[j]);
                cur.pairs += (unsigned)KGMT_WALK_WIDTH;
            }
        }
        return h;
    }
};

/* ------------------------------------------------------------------------ dynamics --
 * propagateAndCheck, src/statePropagator/statePropagator.cu:21-75, with the controls
 * already drawn.  Explicit Euler on the kinematic bicycle; per step: workspace
 * bounds (:42-45, theta/v NOT advanced when it fires), then theta/v, then the
 * step bbox (:49-59) against the obstacles (:61-64).  The state at the moment of
 * exit is returned whether or not the edge is valid (:67-73).  tanf(steering) is
 * loop-invariant (the reference recomputes it every step, :36); v / L is exact for L = 1. */
struct DynParams { float W, H, L; int numDisc; };

struct EdgeWork { unsigned steps, pairs; };   /* loop trips of statePropagator.cu:31 executed, overlap tests executed */

template <bool UNIT_L, class Collide>
__device__ __forceinline__ bool propagate_steps(float& x, float& y, float& th, float& v, int& i, const Controls& u, const DynParams& p,
                                                float dt, float tanS, const Collide& col, typename Collide::Cursor& cur) {
    for (; i < p.numDisc; ++i) {
        const float px = x, py = y;
        float sn, cs;
        sincosf(th, &sn, &cs);      /* one range reduction; bit-identical to sinf(th), cosf(th) (checked against the reference's kernels) */
        x = __fmaf_rn(dt, __fmul_rn(v, cs), x);
        y = __fmaf_rn(dt, __fmul_rn(v, sn), y);
        if (x <= 0.0f || x >= p.W || y <= 0.0f || y >= p.H) return false;
        const float vl = UNIT_L ? v : __fdiv_rn(v, p.L);
        th = __fmaf_rn(dt, __fmul_rn(vl, tanS), th);
        v = __fmaf_rn(u.a, dt, v);
        const float bnx = (px > x) ? x : px, bxx = (px > x) ? px : x;
        const float bny = (py > y) ? y : py, bxy = (py > y) ? py : y;
        if (col.hit(cur, x, y, bnx, bny, bxx, bxy)) return false;
    }
    return true;
}

template <class Collide>
__device__ __forceinline__ bool propagate_edge(float4& s, const Controls& u, const DynParams& p, const Collide& col,
                                               EdgeWork* work = nullptr) {
    const float dt = __fdiv_rn(u.duration, (float)p.numDisc);
    const float tanS = tanf(u.steering);
    float x = s.x, y = s.y, th = s.z, v = s.w;
    typename Collide::Cursor cur = col.start(x, y);
    int i = 0;
    /* v / L is exact for L = 1 (the reference's value): that loop carries no division and no test for it */
    const bool valid = (p.L == 1.0f) ? propagate_steps<true>(x, y, th, v, i, u, p, dt, tanS, col, cur)
                                     : propagate_steps<false>(x, y, th, v, i, u, p, dt, tanS, col, cur);
    if (work) { work->steps = (unsigned)(valid ? p.numDisc : i + 1); work->pairs = cur.pairs; }
    s = make_float4(x, y, th, v);
    return valid;
}

/* ---------------------------------------------------------- tile-streamed exhaustive test --
 * The same edge (propagateAndCheck, statePropagator.cu:21-75) when the obstacle set does not fit shared memory in
 * one piece: the obstacles stream through shared memory in tiles and the edge is re-integrated once per tile (the
 * integration is a few dozen instructions per step against thousands of AABB tests per tile).  The reference exits at
 * the FIRST step whose bbox overlaps ANY obstacle; the first such step is the minimum over the tiles of each tile's
 * first overlapping step, so pass t only has to look at the steps before the best exit found so far. */
struct EdgeExit { float4 s; int step; bool valid; };   /* state at exit, exit step (numDisc = ran to the end) */

template <bool FIRST>
__device__ __forceinline__ void edge_tile_pass(const float4 s0, const Controls& u, const DynParams& p, float dt, float tanS,
                                               const float4* tile, int nObs, EdgeExit& e) {
    const bool unitL = (p.L == 1.0f);
    const CollideSmemAll col{tile, nObs};
    CollideSmemAll::Cursor cur{0u};
    float x = s0.x, y = s0.y, th = s0.z, v = s0.w;
    const int limit = FIRST ? p.numDisc : e.step;
    for (int i = 0; i < limit; ++i) {
        const float px = x, py = y;
        float sn, cs;
        sincosf(th, &sn, &cs);      /* one range reduction; bit-identical to sinf(th), cosf(th) (checked against the reference's kernels) */
        x = __fmaf_rn(dt, __fmul_rn(v, cs), x);
        y = __fmaf_rn(dt, __fmul_rn(v, sn), y);
        if (FIRST && (x <= 0.0f || x >= p.W || y <= 0.0f || y >= p.H)) {            /* :42-45 */
            e.s = make_float4(x, y, th, v); e.step = i; e.valid = false;
            return;
        }
        const float vl = unitL ? v : __fdiv_rn(v, p.L);
        th = __fmaf_rn(dt, __fmul_rn(vl, tanS), th);
        v = __fmaf_rn(u.a, dt, v);
        const float bnx = (px > x) ? x : px, bxx = (px > x) ? px : x;
        const float bny = (py > y) ? y : py, bxy = (py > y) ? py : y;
        if (col.hit(cur, x, y, bnx, bny, bxx, bxy)) {                                 /* :61-64 */
            e.s = make_float4(x, y, th, v); e.step = i; e.valid = false;
            return;
        }
    }
    if (FIRST) { e.s = make_float4(x, y, th, v); e.step = p.numDisc; e.valid = true; }
}

}  // namespace kgmt
