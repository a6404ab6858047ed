/* ============================================================================
 * kgmt_oracle.c — TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, single-threaded CPU restatement of the reference's KGMT
 * tree-expansion path (nipe1783/cudaSBMP).  It is the CHECKER for the CUDA
 * product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product library (libkgmt_b200.so)
 * never links, loads or calls anything in oracle/.
 *
 * Parity status: PINNED.  The reference ships no tests or golden vectors
 * (SURVEY.md §4), so the pin is the reference's own code run in this
 * container: oracle/_ref/libref_host.so is the reference's unmodified
 * statePropagator.cu + collisionCheck.cu built for the host (oracle/Makefile),
 * and tests/test_oracle_pin.py checks this file against it bit-for-bit
 * (states, flags) in math mode ORC_MATH_HOST; getR1/getR2 are pinned against
 * the reference's __host__ __device__ versions inside oracle/_ref/libref_gpu.so;
 * Philox against the Random123 known-answer vectors.
 * Vectors generated from the reference are committed in tests/golden/.
 *
 * Every function cites the reference lines it follows (paths relative to
 * /root/reference).  Where the reference is racy or undefined the CANONICAL
 * semantics of SURVEY.md Appendix B are used; each such place is marked
 * "canonical:".
 *
 * Math modes
 *   ORC_MATH_HOST (0): expression-for-expression what g++ -O2 -ffp-contract=off
 *       makes of the reference source (no FMA, glibc sinf/cosf/tanf).
 *   ORC_MATH_FMA  (1): the same expressions with the FMA contractions nvcc
 *       applies to the reference on sm_100a (read off its PTX: a, steering,
 *       x, y, theta, v updates are single fma.rn).  glibc trig still differs
 *       from libdevice's by <=1-2 ulp, so GPU parity of states is a tolerance
 *       (north_star: 1e-5 relative per step), flags may flip only inside the
 *       margin this file reports.
 * ========================================================================== */
#include "kgmt_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- Philox -- */
/* Follows /usr/local/cuda/include/curand_philox4x32_x.h:88-190 (cuRAND's
 * Philox4x32-10; the third-party generator the canonical stream is drawn
 * from).  Constants: PHILOX_M4x32_0/1, PHILOX_W32_0/1. */
static inline void philox_round(uint32_t c[4], const uint32_t k[2]) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
    c[1] = (uint32_t)p1; c[3] = (uint32_t)p0; c[0] = n0; c[2] = n2;
}

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k[2] = {key[0], key[1]};
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k);
        if (r < 9) { k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u; }
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

/* curand_uniform: /usr/local/cuda/include/curand_uniform.h:69-72.  The product
 * (float)x*2^-32 is exact, so FMA or not gives the same float. */
float orc_uniform(uint32_t x) {
    return (float)x * 2.3283064e-10f + (2.3283064e-10f / 2.0f);
}

/* canonical stream: the four draws of candidate slot s in
 * iteration itr are the four words of Philox(ctr=(0,0,s,0), key=(seed+itr,0)) —
 * exactly what curand_init(seed+itr, s, 0) followed by four curand_uniform
 * calls yields for curandStatePhilox4_32_10_t. */
void orc_slot_uniforms(uint32_t key0, uint32_t slot, float u[4]) {
    const uint32_t ctr[4] = {0u, 0u, slot, 0u};
    const uint32_t key[2] = {key0, 0u};
    uint32_t w[4];
    orc_philox4x32_10(ctr, key, w);
    for (int i = 0; i < 4; ++i) u[i] = orc_uniform(w[i]);
}

/* -------------------------------------------------------------- controls -- */
/* Control ranges other than the reference's literals (kgmt_params accel/steer/duration min..max; the reference's
 * systems/car.yaml is empty and statePropagator.cu:17-19 hard-codes them).  Process-global, test infrastructure:
 * orc_set_car_ranges(NULL) restores the literal path below, which is the one pinned against the reference build.
 * The general form is lo + u * (hi - lo) with the same operation shapes: float for a and duration, double for the
 * steering (the reference multiplies by the double constant M_PI). */
static int g_car_set = 0;
static double g_car[6];
void orc_set_car_ranges(const double* r6) {
    g_car_set = r6 != NULL;
    if (r6) memcpy(g_car, r6, sizeof(g_car));
}
void orc_controls_general(const float u[3], const double r[6], int math_mode, float* a, float* steering, float* duration) {
    const float aS = (float)(r[1] - r[0]), aL = (float)r[0], dS = (float)(r[5] - r[4]), dL = (float)r[4];
    const double sS = r[3] - r[2], sL = r[2];
    if (math_mode == ORC_MATH_FMA) {
        *a = fmaf(u[0], aS, aL);
        *steering = (float)fma((double)u[1], sS, sL);
        *duration = fmaf(u[2], dS, dL);
    } else {
        volatile float ta = u[0] * aS;  *a = ta + aL;
        volatile double ts = (double)u[1] * sS;  *steering = (float)(ts + sL);
        volatile float td = u[2] * dS;  *duration = td + dL;
    }
}

/* src/statePropagator/statePropagator.cu:17-19 */
void orc_controls(const float u[3], int math_mode, float* a, float* steering, float* duration) {
    if (g_car_set) { orc_controls_general(u, g_car, math_mode, a, steering, duration); return; }
    if (math_mode == ORC_MATH_FMA) {
        *a = fmaf(u[0], 10.0f, -5.0f);
        *steering = (float)fma((double)(u[1] * 2.0f), M_PI, -M_PI);
    } else {
        *a = u[0] * 10.0f - 5.0f;
        volatile double t = (double)(u[1] * 2.0f) * M_PI;   /* volatile: forbid contraction */
        *steering = (float)(t - M_PI);
    }
    *duration = u[2] * 1.0f + 0.05f;
}

/* ------------------------------------------------------------- collision -- */
/* src/collisionCheck/collisionCheck.cu:6-14: "valid" (no overlap) iff on some
 * axis bbMax <= obs.min or obs.max <= bbMin. */
static inline int broad_phase_valid(const float bbMin[2], const float bbMax[2], const float* obs) {
    for (int d = 0; d < 2; ++d)
        if (bbMax[d] <= obs[d] || obs[2 + d] <= bbMin[d]) return 1;
    return 0;
}

/* Signed clearance of one bbox/obstacle pair: > 0 separated, < 0 overlapping,
 * |g| = how far the nearest deciding comparison is from flipping. */
static inline float pair_gap(const float bbMin[2], const float bbMax[2], const float* obs) {
    float g = obs[0] - bbMax[0];
    float t = bbMin[0] - obs[2]; if (t > g) g = t;
    t = obs[1] - bbMax[1];       if (t > g) g = t;
    t = bbMin[1] - obs[3];       if (t > g) g = t;
    return g;
}

/* src/collisionCheck/collisionCheck.cu:16-28 */
int orc_motion_valid(const float bbMin[2], const float bbMax[2], const float* obstacles, int K) {
    for (int o = 0; o < K; ++o)
        if (!broad_phase_valid(bbMin, bbMax, obstacles + 4 * o)) return 0;
    return 1;
}

/* ------------------------------------------------------------- propagate -- */
/* src/statePropagator/statePropagator.cu:21-75 with the controls given.
 * x1[7] is written even when invalid (:67-73).  *margin (optional) receives
 * the smallest distance of any deciding comparison from flipping, over the
 * steps that were executed; *steps (optional) the number of loop bodies
 * entered. */
int orc_propagate_ctrl(const float x0[4], float a, float steering, float duration,
                       int numDisc, float agentLength, const float* obstacles, int K,
                       float width, float height, int math_mode,
                       float x1[7], float* margin, int* steps) {
    const float dt = duration / (float)numDisc;                        /* :21 */
    float x = x0[0], y = x0[1], theta = x0[2], v = x0[3];              /* :22-25 */
    float m = INFINITY;
    int valid = 1, nsteps = 0;
    for (int i = 0; i < numDisc; ++i) {                                /* :31 */
        ++nsteps;
        const float px = x, py = y;                                    /* :33 */
        const float c = cosf(theta), s = sinf(theta), t = tanf(steering); /* :34-36 */
        if (math_mode == ORC_MATH_FMA) {
            x = fmaf(dt, v * c, x);                                    /* :39 */
            y = fmaf(dt, v * s, y);                                    /* :40 */
        } else {
            x += v * c * dt;
            y += v * s * dt;
        }
        {   /* distance of the bounds test from flipping (:42) */
            float b = fabsf(x); if (b < m) m = b;
            b = fabsf(width - x);  if (b < m) m = b;
            b = fabsf(y);          if (b < m) m = b;
            b = fabsf(height - y); if (b < m) m = b;
        }
        if (x <= 0.0 || x >= width || y <= 0.0 || y >= height) {      /* :42-45 */
            valid = 0;
            break;
        }
        if (math_mode == ORC_MATH_FMA) {
            theta = fmaf(dt, (v / agentLength) * t, theta);            /* :46 */
            v = fmaf(a, dt, v);                                        /* :47 */
        } else {
            theta += (v / agentLength) * t * dt;
            v += a * dt;
        }
        float bbMin[2], bbMax[2];                                      /* :51-59 */
        if (px > x) { bbMin[0] = x;  bbMax[0] = px; } else { bbMin[0] = px; bbMax[0] = x; }
        if (py > y) { bbMin[1] = y;  bbMax[1] = py; } else { bbMin[1] = py; bbMax[1] = y; }
        if (margin) {
            for (int o = 0; o < K; ++o) {
                float g = fabsf(pair_gap(bbMin, bbMax, obstacles + 4 * o));
                if (g < m) m = g;
            }
        }
        if (!orc_motion_valid(bbMin, bbMax, obstacles, K)) {           /* :61-64 */
            valid = 0;
            break;
        }
    }
    x1[0] = x; x1[1] = y; x1[2] = theta; x1[3] = v;                   /* :67-73 */
    x1[4] = a; x1[5] = steering; x1[6] = duration;
    if (margin) *margin = m;
    if (steps) *steps = nsteps;
    return valid;
}

/* One candidate from its random stream: draws + controls + propagate.
 * u3_out receives the 4th (accept) uniform, KGMT.cu:395. */
int orc_propagate_slot(const float x0[4], uint32_t key0, uint32_t slot, int numDisc,
                       float agentLength, const float* obstacles, int K,
                       float width, float height, int math_mode,
                       float x1[7], float* u3_out, float* margin, int* steps) {
    float u[4], a, st, du;
    orc_slot_uniforms(key0, slot, u);
    orc_controls(u, math_mode, &a, &st, &du);
    if (u3_out) *u3_out = u[3];
    return orc_propagate_ctrl(x0, a, st, du, numDisc, agentLength, obstacles, K,
                              width, height, math_mode, x1, margin, steps);
}

/* ---------------------------------------------------------- region index -- */
/* src/planners/KGMT.cu:602-609 (== OccupancyGrid::getCellIndex,
 * src/occupancyMaps/OccupancyGrid.cu:12-19).  The C cast of an out-of-range
 * float is undefined; states reaching here are finite and small in every
 * configuration, and the CUDA side uses the same truncation (cvt.rzi). */
int orc_getR1(float x, float y, float R1Size, int N) {
    int cellX = (int)(x / R1Size);
    int cellY = (int)(y / R1Size);
    if (cellX >= 0 && cellX < N && cellY >= 0 && cellY < N) return cellY * N + cellX;
    return -1;
}

/* src/planners/KGMT.cu:610-629 */
int orc_getR2(float x, float y, int r1, float R1Size, int N, float R2Size, int n) {
    if (r1 == -1) return -1;
    int cellY_R1 = r1 / N;
    int cellX_R1 = r1 % N;
    float localX = x - cellX_R1 * R1Size;
    float localY = y - cellY_R1 * R1Size;
    int cellX_R2 = (int)(localX / R2Size);
    int cellY_R2 = (int)(localY / R2Size);
    if (cellX_R2 >= 0 && cellX_R2 < n && cellY_R2 >= 0 && cellY_R2 < n)
        return r1 * (n * n) + cellY_R2 * n + cellX_R2;
    return -1;
}

/* The same function as nvcc/ptxas build it for sm_100a: `x - cell*R1Size` is
 * contracted into one FFMA (read off the SASS of the reference's KGMT.o:
 * `FFMA R40, R40, -R7, R5`).  Identical to orc_getR2 whenever cell*R1Size is
 * exactly representable (all shipped configurations: R1Size = 1.25). */
int orc_getR2_fma(float x, float y, int r1, float R1Size, int N, float R2Size, int n) {
    if (r1 == -1) return -1;
    int cellY_R1 = r1 / N;
    int cellX_R1 = r1 % N;
    float localX = fmaf(-(float)cellX_R1, R1Size, x);
    float localY = fmaf(-(float)cellY_R1, R1Size, y);
    int cellX_R2 = (int)(localX / R2Size);
    int cellY_R2 = (int)(localY / R2Size);
    if (cellX_R2 >= 0 && cellX_R2 < n && cellY_R2 >= 0 && cellY_R2 < n)
        return r1 * (n * n) + cellY_R2 * n + cellX_R2;
    return -1;
}

int orc_getR2_mode(float x, float y, int r1, float R1Size, int N, float R2Size, int n, int math_mode) {
    return math_mode == ORC_MATH_FMA ? orc_getR2_fma(x, y, r1, R1Size, N, R2Size, n)
                                     : orc_getR2(x, y, r1, R1Size, N, R2Size, n);
}

/* ---------------------------------------------------------------- scores -- */
/* src/planners/KGMT.cu:500-537, for any N (canonical: App. B #9).
 *   covR    = (# available R2 cells of the R1 cell) / n^2           (:510-514)
 *   freeVol = (eps + nValid) / (eps + nValid + nInvalid), float      (:516)
 *   score   = freeVol^4 / ((1+covR) * (1+R1^2)), evaluated in double and
 *             narrowed to float                                      (:517)
 * canonical: the reference's pow() calls are replaced by exact repeated
 * multiplication in double (agrees with a correctly rounded pow to <=1.5 ulp
 * of double, i.e. identical after narrowing except in ~1e-8 of cases), and the
 * block sum (cub::BlockReduce, :520-522, order unspecified) by a fixed
 * reduction order that the CUDA kernel reproduces exactly:
 *   p[t] = sum_{k=0,1,..} score[t + 1024k]  (t < 1024, ascending k)
 *   for stride = 512,256,...,1:  p[t] += p[t+stride]  (t < stride)
 * R1Score = 1 for unavailable cells, score/total otherwise (:532-537).
 * R1Threshold = total / (#available cells) — canonical (App. B #8); unused. */
void orc_scores(const int* R1Avail, const int* R2Avail, const int* R1Valid, const int* R1Invalid,
                const int* R1, int N, int n, float epsilon,
                float* R1Score, float* R1Threshold) {
    const int cells = N * N, nn = n * n;
    float* sc = (float*)calloc((size_t)cells, sizeof(float));
    float p[1024];
    int avail = 0;
    for (int c = 0; c < cells; ++c) {
        float score = 0.0f;
        if (R1Avail[c] != 0) {
            ++avail;
            float covR = 0;
            for (int i = c * nn; i < (c + 1) * nn; ++i) covR += (float)(R2Avail[i] != 0);
            covR /= (float)nn;
            const int nValid = R1Valid[c];
            float freeVol = (epsilon + (float)nValid) / (epsilon + (float)nValid + (float)R1Invalid[c]);
            double f2 = (double)freeVol * (double)freeVol;
            double f4 = f2 * f2;
            double r  = (double)R1[c];
            double den = (double)(1.0f + covR) * (1.0 + r * r);
            score = (float)(f4 / den);
        }
        sc[c] = score;
    }
    for (int t = 0; t < 1024; ++t) {
        float acc = 0.0f;
        for (int c = t; c < cells; c += 1024) acc += sc[c];
        p[t] = acc;
    }
    for (int stride = 512; stride >= 1; stride >>= 1)
        for (int t = 0; t < stride; ++t) p[t] = p[t] + p[t + stride];
    const float total = p[0];
    if (R1Threshold) *R1Threshold = avail ? total / (float)avail : 0.0f;
    for (int c = 0; c < cells; ++c)
        R1Score[c] = (R1Avail[c] == 0) ? 1.0f : sc[c] / total;
    free(sc);
}

/* ------------------------------------------------------------ goal test --- */
/* src/planners/KGMT.cu:635-638: float differences, squares and sqrt in
 * double, narrowed to float, strict '<'.  canonical: pow(d,2) -> exact d*d. */
int orc_in_goal(const float* x, const float* goal, float r) {
    double dx = (double)(x[0] - goal[0]);
    double dy = (double)(x[1] - goal[1]);
    volatile double sx = dx * dx, sy = dy * dy;       /* exact; volatile forbids fma */
    float dist = (float)sqrt(sx + sy);
    return dist < r;
}

/* =============================================================== planner == */
struct orc_planner {
    /* ctor arguments, include/planners/KGMT.cuh:28 */
    float width, height; int N, n, numIterations, maxTreeSize, numDisc;
    float agentLength, goalThreshold;
    float R1Size, R2Size;                       /* KGMT.cu:13-14 */
    uint32_t seed; int math_mode;
    /* arrays, same names/layout as the reference members (KGMT.cu:16-40) */
    uint8_t *G, *GNew;
    int *treeParentIdx, *uParentIdx;
    float *treeSamples, *unexploredSamples, *costs;
    int *R1Avail, *R2Avail, *R1Valid, *R2Valid, *R1Invalid, *R2Invalid, *R1, *R2;
    float *R1Score;
    float xGoal[7];
    float *obstacles; int K;
    /* scalars */
    int treeSize, itr, frontierStart, frontierCount;
    float costToGoal, R1Threshold;
    int goalIdx;
    int status;
    long long expansions;
    /* per-iteration scratch exposed to tests */
    uint8_t *uValid; int *uR1, *uR2; float *uU3; float *uMargin;
    int lastM, lastAccepted, lastMode, lastChildren;
};

#define ALLOC(p, n) do { (p) = calloc((size_t)(n), sizeof(*(p))); } while (0)

orc_planner* orc_create(float width, float height, int N, int n, int numIterations, int maxTreeSize,
                        int numDisc, float agentLength, float goalThreshold,
                        uint32_t seed, int math_mode) {
    orc_planner* p = (orc_planner*)calloc(1, sizeof(orc_planner));
    p->width = width; p->height = height; p->N = N; p->n = n; p->numIterations = numIterations;
    p->maxTreeSize = maxTreeSize; p->numDisc = numDisc; p->agentLength = agentLength;
    p->goalThreshold = goalThreshold; p->seed = seed; p->math_mode = math_mode;
    p->R1Size = width / N;                      /* KGMT.cu:13 */
    p->R2Size = width / (n * N);                /* KGMT.cu:14 */
    const size_t T = (size_t)maxTreeSize, c1 = (size_t)N * N, c2 = c1 * n * n;
    ALLOC(p->G, T); ALLOC(p->GNew, T); ALLOC(p->treeParentIdx, T); ALLOC(p->uParentIdx, T);
    ALLOC(p->treeSamples, T * 7); ALLOC(p->unexploredSamples, T * 7); ALLOC(p->costs, T);
    ALLOC(p->R1Avail, c1); ALLOC(p->R1Valid, c1); ALLOC(p->R1Invalid, c1); ALLOC(p->R1, c1);
    ALLOC(p->R1Score, c1);
    ALLOC(p->R2Avail, c2); ALLOC(p->R2Valid, c2); ALLOC(p->R2Invalid, c2); ALLOC(p->R2, c2);
    ALLOC(p->uValid, T); ALLOC(p->uR1, T); ALLOC(p->uR2, T); ALLOC(p->uU3, T); ALLOC(p->uMargin, T);
    orc_reset(p);
    return p;
}

void orc_destroy(orc_planner* p) {
    if (!p) return;
    free(p->G); free(p->GNew); free(p->treeParentIdx); free(p->uParentIdx);
    free(p->treeSamples); free(p->unexploredSamples); free(p->costs);
    free(p->R1Avail); free(p->R1Valid); free(p->R1Invalid); free(p->R1); free(p->R1Score);
    free(p->R2Avail); free(p->R2Valid); free(p->R2Invalid); free(p->R2);
    free(p->uValid); free(p->uR1); free(p->uR2); free(p->uU3); free(p->uMargin);
    free(p->obstacles);
    free(p);
}

/* State as left by the reference constructor, KGMT.cu:16-40,70-72 (zeros,
 * parents -1, scores 1.0).  canonical: costToGoal explicitly 0 (App. B #4). */
void orc_reset(orc_planner* p) {
    const size_t T = (size_t)p->maxTreeSize, c1 = (size_t)p->N * p->N, c2 = c1 * p->n * p->n;
    memset(p->G, 0, T); memset(p->GNew, 0, T);
    for (size_t i = 0; i < T; ++i) { p->treeParentIdx[i] = -1; p->uParentIdx[i] = -1; }
    memset(p->treeSamples, 0, T * 7 * sizeof(float));
    memset(p->unexploredSamples, 0, T * 7 * sizeof(float));
    memset(p->costs, 0, T * sizeof(float));
    memset(p->R1Avail, 0, c1 * 4); memset(p->R1Valid, 0, c1 * 4); memset(p->R1Invalid, 0, c1 * 4);
    memset(p->R1, 0, c1 * 4);
    for (size_t i = 0; i < c1; ++i) p->R1Score[i] = 1.0f;
    memset(p->R2Avail, 0, c2 * 4); memset(p->R2Valid, 0, c2 * 4); memset(p->R2Invalid, 0, c2 * 4);
    memset(p->R2, 0, c2 * 4);
    p->treeSize = 0; p->itr = 0; p->frontierStart = 0; p->frontierCount = 0;
    p->costToGoal = 0.0f; p->R1Threshold = 0.0f; p->goalIdx = -1; p->status = ORC_STATUS_RUNNING;
    p->expansions = 0; p->lastM = p->lastAccepted = p->lastMode = p->lastChildren = 0;
}

void orc_set_obstacles(orc_planner* p, const float* aabb, int K) {
    free(p->obstacles);
    p->obstacles = (float*)malloc(sizeof(float) * 4 * (size_t)(K > 0 ? K : 1));
    if (K > 0) memcpy(p->obstacles, aabb, sizeof(float) * 4 * (size_t)K);
    p->K = K;
}

/* KGMT.cu:85-101,114: root into slot 0, frontier = {root}, root's cells marked. */
void orc_begin(orc_planner* p, const float initial[7], const float goal[7]) {
    memcpy(p->treeSamples, initial, 7 * sizeof(float));               /* :85 */
    p->G[0] = 1;                                                       /* :86-87 */
    int r1 = orc_getR1(initial[0], initial[1], p->R1Size, p->N);      /* :88 */
    int r2 = orc_getR2_mode(initial[0], initial[1], r1, p->R1Size, p->N, p->R2Size, p->n, p->math_mode); /* :89 */
    if (r1 >= 0) { p->R1[r1] = 1; p->R1Avail[r1] = 1; p->R1Valid[r1] = 1; }  /* :94,95,97 */
    if (r2 >= 0) p->R2Avail[r2] = 1;                                   /* :96 */
    memcpy(p->xGoal, goal, 7 * sizeof(float));                         /* :101 */
    p->treeSize = 1; p->itr = 0;                                       /* :113-114 */
    p->frontierStart = 0; p->frontierCount = 1;
    p->costToGoal = 0.0f; p->goalIdx = -1; p->status = ORC_STATUS_RUNNING;
}

/* Expansion policy of KGMT.cu:151-158: 32 children per frontier node while
 * they fit (propagateG), else floor(remaining/active) each (propagateGV2).
 * canonical (App. B #7): when even one child each does not fit, the first
 * `remaining` frontier nodes get one child. */
void orc_expansion_shape(int activeSize, int treeSize, int maxTreeSize,
                         int* mode, int* children, int* M) {
    const int remaining = maxTreeSize - treeSize;
    if (32LL * activeSize > (long long)remaining) {                    /* :153 */
        int iterations = (int)((float)remaining / (float)activeSize);  /* :157 */
        if (iterations >= 1) { *mode = 2; *children = iterations; *M = activeSize * iterations; }
        else                 { *mode = 3; *children = 1;          *M = remaining; }
    } else {
        *mode = 1; *children = 32; *M = 32 * activeSize;               /* :151-152 */
    }
}

/* Map update + accept decision for one iteration's candidates.
 * Follows KGMT.cu:390-411 with canonical semantics:
 *   - r1 == -1: no map update, not accepted (App. B #1); r2 == -1 with a
 *     valid r1: R1-family updated, R2-family skipped, accept on the score
 *     clause alone;
 *   - the accept test reads R2Avail / R1Score as they were at iteration start
 *     (App. B #2) — R2AvailSnap may alias nothing written here;
 *   - every accept flag is rewritten (App. B #3). */
void orc_update_maps(int M, const int* r1v, const int* r2v, const uint8_t* valid, const float* u3,
                     const float* R1Score, const int* R2AvailSnap,
                     int* R1, int* R2, int* R1Valid, int* R2Valid, int* R1Invalid, int* R2Invalid,
                     int* R1Avail, int* R2Avail, uint8_t* accept) {
    for (int s = 0; s < M; ++s) {
        const int r1 = r1v[s], r2 = r2v[s];
        accept[s] = 0;
        if (r1 < 0) continue;
        R1[r1] += 1;                                                   /* :392 */
        if (r2 >= 0) R2[r2] += 1;                                      /* :393 */
        if (valid[s]) {
            int acc = (u3[s] <= R1Score[r1]);                          /* :396 */
            if (r2 >= 0 && R2AvailSnap[r2] == 0) acc = 1;
            accept[s] = (uint8_t)acc;                                  /* :397 */
            R1Avail[r1] = 1;                                           /* :399-401 */
            if (r2 >= 0) { R2Avail[r2] = 1; R2Valid[r2] += 1; }        /* :402-405 */
            R1Valid[r1] += 1;                                          /* :406 */
        } else {
            R1Invalid[r1] += 1;                                        /* :409 */
            if (r2 >= 0) R2Invalid[r2] += 1;                           /* :410 */
        }
    }
}

/* Ordered insertion of accepted candidates, KGMT.cu:222-245,555-591.
 * Accepted slots in ascending order take tree slots treeSize, treeSize+1, …
 * (exclusive scan + findInd, :222-229,568).  cost = cost[parent] + duration
 * (:585-586,631-633).  canonical: goal cost = minimum over the new nodes in
 * the goal disc, ties to the lowest tree index (App. B #5).
 * Returns the number inserted. */
int orc_insert(int M, const uint8_t* accept, const float* cand /*[M][7]*/, const int* candParent,
               int treeSize, float* treeSamples, int* treeParentIdx, float* costs, uint8_t* G,
               const float* goal, float r, float* costToGoal, int* goalIdx) {
    int j = 0;
    for (int s = 0; s < M; ++s) {
        if (!accept[s]) continue;
        const int dst = treeSize + j;                                  /* :568 */
        const int par = candParent[s];                                 /* :571 */
        treeParentIdx[dst] = par;                                      /* :572 */
        memcpy(treeSamples + 7 * (size_t)dst, cand + 7 * (size_t)s, 7 * sizeof(float)); /* :573-579 */
        G[dst] = 1;                                                    /* :582 */
        costs[dst] = costs[par] + cand[7 * (size_t)s + 6];            /* :585-586 */
        if (orc_in_goal(cand + 7 * (size_t)s, goal, r)) {              /* :589 */
            if (*goalIdx < 0 || costs[dst] < *costToGoal) { *costToGoal = costs[dst]; *goalIdx = dst; }
        }
        ++j;
    }
    return j;
}

/* One tree-expansion step = stages 1-5, the body of the while loop at
 * KGMT.cu:118-259.  Returns the status after the step. */
int orc_iterate(orc_planner* p) {
    if (p->status != ORC_STATUS_RUNNING) return p->status;
    p->itr += 1;                                                       /* :119 */
    const int c2 = p->N * p->N * p->n * p->n;

    /* stage 1a: scores (:122-136) */
    orc_scores(p->R1Avail, p->R2Avail, p->R1Valid, p->R1Invalid, p->R1, p->N, p->n, 0.01f,
               p->R1Score, &p->R1Threshold);

    /* stage 1b: frontier (:139-147).  G is exactly the range of nodes appended
     * in the previous step, in ascending order. */
    const int activeSize = p->frontierCount;
    if (activeSize == 0) { p->status = ORC_STATUS_FRONTIER_EMPTY; return p->status; } /* canonical: App. B #11 */
    int mode, children, M;
    orc_expansion_shape(activeSize, p->treeSize, p->maxTreeSize, &mode, &children, &M);
    p->lastMode = mode; p->lastChildren = children; p->lastM = M;

    /* stages 2-4: sample, propagate, collide (:386-389 / :454-457) */
    const uint32_t key0 = p->seed + (uint32_t)p->itr;
    for (int f = 0; f < activeSize; ++f) p->G[p->frontierStart + f] = 0;  /* :378 / :451 */
    for (int s = 0; s < M; ++s) {
        const int f = s / children;                                    /* :374,:454 */
        const int x0Idx = p->frontierStart + f;
        float* x1 = p->unexploredSamples + 7 * (size_t)s;             /* :387 */
        p->uParentIdx[s] = x0Idx;                                      /* :388 */
        p->uValid[s] = (uint8_t)orc_propagate_slot(p->treeSamples + 7 * (size_t)x0Idx, key0, (uint32_t)s,
                                    p->numDisc, p->agentLength, p->obstacles, p->K,
                                    p->width, p->height, p->math_mode, x1, &p->uU3[s],
                                    &p->uMargin[s], NULL);
        p->uR1[s] = orc_getR1(x1[0], x1[1], p->R1Size, p->N);         /* :390 */
        p->uR2[s] = orc_getR2_mode(x1[0], x1[1], p->uR1[s], p->R1Size, p->N, p->R2Size, p->n, p->math_mode); /* :391 */
    }
    p->expansions += M;

    /* stage 5a: maps + accept (:392-411), snapshot semantics */
    int* snap = (int*)malloc(sizeof(int) * (size_t)c2);
    memcpy(snap, p->R2Avail, sizeof(int) * (size_t)c2);
    memset(p->GNew, 0, (size_t)p->maxTreeSize);
    orc_update_maps(M, p->uR1, p->uR2, p->uValid, p->uU3, p->R1Score, snap,
                    p->R1, p->R2, p->R1Valid, p->R2Valid, p->R1Invalid, p->R2Invalid,
                    p->R1Avail, p->R2Avail, p->GNew);
    free(snap);

    /* stage 5b: insertion (:222-249) */
    const int accepted = orc_insert(M, p->GNew, p->unexploredSamples, p->uParentIdx, p->treeSize,
                                    p->treeSamples, p->treeParentIdx, p->costs, p->G,
                                    p->xGoal, p->goalThreshold, &p->costToGoal, &p->goalIdx);
    memset(p->GNew, 0, (size_t)p->maxTreeSize);                        /* :556, canonical full clear */
    p->frontierStart = p->treeSize;
    p->frontierCount = accepted;
    p->treeSize += accepted;                                           /* :249 */
    p->lastAccepted = accepted;

    if (p->costToGoal != 0.0f)              p->status = ORC_STATUS_SOLVED;       /* :252 */
    else if (p->treeSize >= p->maxTreeSize) p->status = ORC_STATUS_TREE_FULL;    /* :255 */
    else if (accepted == 0)                 p->status = ORC_STATUS_FRONTIER_EMPTY;
    else if (p->itr >= p->numIterations)    p->status = ORC_STATUS_ITER_LIMIT;   /* :118 */
    return p->status;
}

/* KGMT::plan, KGMT.cu:80-292 (without timing / CSV dump). */
int orc_plan(orc_planner* p, const float initial[7], const float goal[7]) {
    orc_begin(p, initial, goal);
    if (p->numIterations <= 0) { p->status = ORC_STATUS_ITER_LIMIT; return p->status; }
    while (orc_iterate(p) == ORC_STATUS_RUNNING) { }
    return p->status;
}

/* ------------------------------------------------------------- accessors -- */
int   orc_tree_size(const orc_planner* p)      { return p->treeSize; }
int   orc_iterations(const orc_planner* p)     { return p->itr; }
float orc_cost_to_goal(const orc_planner* p)   { return p->costToGoal; }
int   orc_goal_index(const orc_planner* p)     { return p->goalIdx; }
int   orc_status(const orc_planner* p)         { return p->status; }
long long orc_expansions(const orc_planner* p) { return p->expansions; }
int   orc_frontier_start(const orc_planner* p) { return p->frontierStart; }
int   orc_frontier_count(const orc_planner* p) { return p->frontierCount; }
int   orc_last_M(const orc_planner* p)         { return p->lastM; }
int   orc_last_accepted(const orc_planner* p)  { return p->lastAccepted; }
int   orc_last_mode(const orc_planner* p)      { return p->lastMode; }
int   orc_last_children(const orc_planner* p)  { return p->lastChildren; }
float orc_R1Threshold(const orc_planner* p)    { return p->R1Threshold; }

void* orc_array(orc_planner* p, int id) {
    switch (id) {
        case ORC_ARR_TREE_SAMPLES: return p->treeSamples;
        case ORC_ARR_UNEXPLORED:   return p->unexploredSamples;
        case ORC_ARR_TREE_PARENT:  return p->treeParentIdx;
        case ORC_ARR_U_PARENT:     return p->uParentIdx;
        case ORC_ARR_G:            return p->G;
        case ORC_ARR_R2AVAIL:      return p->R2Avail;
        case ORC_ARR_R1AVAIL:      return p->R1Avail;
        case ORC_ARR_R1VALID:      return p->R1Valid;
        case ORC_ARR_R2VALID:      return p->R2Valid;
        case ORC_ARR_R1INVALID:    return p->R1Invalid;
        case ORC_ARR_R2INVALID:    return p->R2Invalid;
        case ORC_ARR_R1SCORE:      return p->R1Score;
        case ORC_ARR_R1:           return p->R1;
        case ORC_ARR_R2:           return p->R2;
        case ORC_ARR_COSTS:        return p->costs;
        case ORC_ARR_U_VALID:      return p->uValid;
        case ORC_ARR_U_R1:         return p->uR1;
        case ORC_ARR_U_R2:         return p->uR2;
        case ORC_ARR_U_U3:         return p->uU3;
        case ORC_ARR_U_MARGIN:     return p->uMargin;
        default: return NULL;
    }
}

/* ------------------------------------------------- batch helpers (bench) -- */
/* region indices of M points (rows of `stride` floats, x at [0], y at [1]) */
void orc_regions_batch(const float* xy, int stride, long M, float R1Size, int N, float R2Size, int n,
                       int math_mode, int* r1, int* r2) {
    for (long s = 0; s < M; ++s) {
        const float x = xy[(size_t)s * stride], y = xy[(size_t)s * stride + 1];
        r1[s] = orc_getR1(x, y, R1Size, N);
        r2[s] = orc_getR2_mode(x, y, r1[s], R1Size, N, R2Size, n, math_mode);
    }
}

/* M candidates, candidate s expands parents[parentOf[s]] (rows of 7 floats)
 * with stream (key0, slot0+s).  Any output pointer may be NULL. */
void orc_propagate_batch(const float* parents, const int* parentOf, long M,
                         float* x1, uint8_t* valid, float* u3, float* margin,
                         int numDisc, float L, uint32_t key0, uint32_t slot0,
                         const float* obstacles, int K, float W, float H, int math_mode) {
    for (long s = 0; s < M; ++s) {
        float tmp[7], u, m;
        float* dst = x1 ? x1 + 7 * s : tmp;
        int ok = orc_propagate_slot(parents + 7 * (size_t)parentOf[s], key0, slot0 + (uint32_t)s, numDisc, L,
                                    obstacles, K, W, H, math_mode, dst, &u, margin ? &m : NULL, NULL);
        if (valid) valid[s] = (uint8_t)ok;
        if (u3) u3[s] = u;
        if (margin) margin[s] = m;
    }
}
