/* planners/KGMT.cuh — the reference planner's C++ face over libkgmt_b200.so.
 *
 * Drop-in for /root/reference include/planners/KGMT.cuh:23-109 on the KGMT tree-expansion path: same
 * constructor (KGMT.cuh:28), same plan() (KGMT.cuh:31), same public scalars (KGMT.cuh:34-43,103-106), so the
 * reference's only caller, demos/main.cu:30,62, compiles and runs unchanged.  Everything the reference keeps in
 * 25 thrust::device_vectors lives behind one kgmt_ctx (include/kgmt_c.h); the vectors are not re-exposed — read
 * them with kgmt_export() in the reference's own element layout, or from the CSV files plan() writes.
 *
 * Header-only; link with -lkgmt_b200.  No CPU fallback: without a B200 the constructor reports the error and
 * plan() does nothing but say so.
 */
#pragma once
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>

#include "kgmt_c.h"
#include "helper/helper.cuh"
#include "collisionCheck/collisionCheck.cuh"
#include "statePropagator/statePropagator.cuh"
#include "agent/Agent.h"
#include "state/State.h"

class KGMT {
  public:
    KGMT() = default;
    KGMT(float width, float height, int N, int n, int numIterations, int maxTreeSize, int numDisc, float agentLength,
         float goalThreshold)
        : numIterations_(numIterations), maxTreeSize_(maxTreeSize), numDisc_(numDisc), treeSize_(0), width_(width),
          height_(height), costToGoal_(0.0f), agentLength_(agentLength), R1Threshold_(0.0f), goalThreshold_(goalThreshold),
          N_(N), n_(n) {
        kgmt_params p;
        kgmt_default_params(&p);
        p.width = width; p.height = height; p.N = N; p.n = n; p.num_iterations = numIterations;
        p.max_tree_size = maxTreeSize; p.num_disc = numDisc; p.agent_length = agentLength; p.goal_threshold = goalThreshold;
        p.seed = seed_;                       /* the reference seeds from time(NULL) inside plan() (KGMT.cu:111) */
        p.record_candidates = 1;              /* keep unexploredSamples.csv / uParentIdx.csv meaningful (KGMT.cu:300,302) */
        create(p);
    }
    /* not in the reference: every parameter, including the car model's control ranges the reference hard-codes in
     * statePropagator.cu:17-19 (e.g. filled by kgmt_params_from_yaml from a systems/car.yaml) */
    explicit KGMT(const kgmt_params& params)
        : numIterations_(params.num_iterations), maxTreeSize_(params.max_tree_size), numDisc_(params.num_disc), treeSize_(0),
          width_(params.width), height_(params.height), costToGoal_(0.0f), agentLength_(params.agent_length), R1Threshold_(0.0f),
          goalThreshold_(params.goal_threshold), N_(params.N), n_(params.n) {
        seed_ = params.seed;
        timeSeed_ = false;
        create(params);
    }
    KGMT(const KGMT&) = delete;
    KGMT& operator=(const KGMT&) = delete;
    ~KGMT() { kgmt_destroy(ctx_); }

  private:
    void create(const kgmt_params& p) {
        const int rc = kgmt_create(&p, &ctx_);
        if (rc != KGMT_OK) {
            std::printf("KGMT: kgmt_create failed (%d): %s\n", rc, kgmt_last_error(ctx_));
            kgmt_destroy(ctx_);
            ctx_ = nullptr;
            return;
        }
        R1Size_ = kgmt_r1_size(ctx_);         /* KGMT.cu:13 */
        R2Size_ = kgmt_r2_size(ctx_);         /* KGMT.cu:14 */
    }

  public:

    /* KGMT::plan, KGMT.cu:80-317.  initial / goal: host float[7]; d_obstacles: DEVICE float[obstaclesCount][4],
     * caller-owned (main.cu:60-64).  Prints what the reference prints and leaves its 13 CSV files in the cwd.
     * Unlike the reference (KGMT.cu:314-316) it can be called again on the same object. */
    void plan(float* initial, float* goal, float* d_obstacles, int obstaclesCount) {
        if (!ctx_) { std::printf("KGMT: no context (construction failed)\n"); return; }
        const auto t0 = std::chrono::steady_clock::now();
        std::printf("Goal: %f, %f\n", goal[0], goal[1]);                                        /* KGMT.cu:100 */
        if (timeSeed_) kgmt_set_seed(ctx_, (uint32_t)std::chrono::system_clock::now().time_since_epoch().count());
        int rc = kgmt_set_obstacles(ctx_, d_obstacles, obstaclesCount);
        kgmt_result r{};
        if (rc == KGMT_OK) rc = kgmt_plan(ctx_, initial, goal, &r);
        if (rc != KGMT_OK) { std::printf("KGMT: plan failed (%d): %s\n", rc, kgmt_last_error(ctx_)); return; }
        treeSize_ = r.tree_size; costToGoal_ = r.cost_to_goal; iterations_ = r.iterations; stop_ = r.stop;
        goalIndex_ = r.goal_index; expansions_ = r.expansions; deviceMs_ = r.device_ms;
        if (r.stop == KGMT_TREE_FULL) {                                                          /* KGMT.cu:255-258 */
            std::printf("Iteration %d, Tree size %d\n", r.iterations, r.tree_size);
            std::printf("Tree size exceeded maxTreeSize\n");
        }
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::printf("time inside KGMT is %f\n", sec);                                            /* KGMT.cu:294-295 */
        std::printf("Iteration %d, Tree size %d\n", r.iterations, r.tree_size);                  /* KGMT.cu:296 */
        if (dumpCsv_) kgmt_dump_csv(ctx_, ".");                                                  /* KGMT.cu:299-311 */
    }

    /* ---- the reference's public fields (KGMT.cuh:34-43,103-106) */
    int numIterations_ = 0, maxTreeSize_ = 0, numDisc_ = 0, treeSize_ = 0;
    float width_ = 0, height_ = 0, costToGoal_ = 0, agentLength_ = 0, R1Threshold_ = 0, goalThreshold_ = 0;
    int N_ = 0, n_ = 0;
    float R1Size_ = 0, R2Size_ = 0;

    /* ---- what the reference does not have */
    int iterations_ = 0, stop_ = 0, goalIndex_ = -1;
    long long expansions_ = 0;
    float deviceMs_ = 0;
    kgmt_ctx* context() const { return ctx_; }
    void setSeed(uint32_t s) { seed_ = s; timeSeed_ = false; if (ctx_) kgmt_set_seed(ctx_, s); }
    void setDumpCsv(bool on) { dumpCsv_ = on; }

  private:
    kgmt_ctx* ctx_ = nullptr;
    uint32_t seed_ = 1;
    bool timeSeed_ = true, dumpCsv_ = true;
};

/* region-index helpers, KGMT.cuh:207-210 / KGMT.cu:602-638 */
__host__ __device__ inline int getR1(float x, float y, float R1Size, int N) {
    const int cx = (int)(x / R1Size), cy = (int)(y / R1Size);
    return (cx >= 0 && cx < N && cy >= 0 && cy < N) ? cy * N + cx : -1;
}
__host__ __device__ inline int getR2(float x, float y, int r1, float R1Size, int N, float R2Size, int n) {
    if (r1 == -1) return -1;
    const int row = r1 / N, col = r1 % N;
    const int cx = (int)((x - col * R1Size) / R2Size), cy = (int)((y - row * R1Size) / R2Size);
    return (cx >= 0 && cx < n && cy >= 0 && cy < n) ? r1 * (n * n) + cy * n + cx : -1;
}
__device__ inline float getCost(float*, float* x1) { return x1[6]; }
__device__ inline bool inGoalRegion(float* x, float* goal, float r) {
    const double dx = (double)(x[0] - goal[0]), dy = (double)(x[1] - goal[1]);
    return (float)sqrt(dx * dx + dy * dy) < r;
}
