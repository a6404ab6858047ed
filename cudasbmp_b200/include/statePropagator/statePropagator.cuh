/* statePropagator/statePropagator.cuh — the reference's device propagation API
 * (include/statePropagator/statePropagator.cuh:5-14) as an inline device function with the same arguments and the same
 * arithmetic (src/statePropagator/statePropagator.cu:17-75; numerical contract in DESIGN.md): three curand_uniform draws
 * -> (a, steering, duration), numDisc Euler steps of the car model, workspace bounds, step bbox vs every obstacle.
 * x1[7] is written whether or not the edge is valid.  The planner's fused kernel does not call this (it samples Philox
 * statelessly and culls collisions through a grid); this is for code written against the reference API. */
#pragma once
#include <curand_kernel.h>
#include "collisionCheck/collisionCheck.cuh"

__device__ inline bool propagateAndCheck(float* x0, float* x1, int numDisc, float agentLength, curandState* state,
                                         float* obstacles, int obstaclesCount, float width, float height) {
    const float a = __fmaf_rn(curand_uniform(state), 10.0f, -5.0f);
    const float u1 = curand_uniform(state);
    const float steering = __double2float_rn(__fma_rn((double)__fadd_rn(u1, u1), 3.14159265358979323846, -3.14159265358979323846));
    const float duration = __fadd_rn(curand_uniform(state), 0.05f);
    const float dt = __fdiv_rn(duration, (float)numDisc);
    const float tanS = tanf(steering);
    float x = x0[0], y = x0[1], th = x0[2], v = x0[3];
    bool ok = true;
    for (int i = 0; i < numDisc; ++i) {
        float prev[2] = {x, y};
        x = __fmaf_rn(dt, __fmul_rn(v, cosf(th)), x);
        y = __fmaf_rn(dt, __fmul_rn(v, sinf(th)), y);
        if (x <= 0.0f || x >= width || y <= 0.0f || y >= height) { ok = false; break; }
        th = __fmaf_rn(dt, __fmul_rn(__fdiv_rn(v, agentLength), tanS), th);
        v = __fmaf_rn(a, dt, v);
        float cur[2] = {x, y};
        float lo[2] = {fminf(prev[0], x), fminf(prev[1], y)}, hi[2] = {fmaxf(prev[0], x), fmaxf(prev[1], y)};
        if (!isMotionValid(prev, cur, lo, hi, obstacles, obstaclesCount)) { ok = false; break; }
    }
    x1[0] = x; x1[1] = y; x1[2] = th; x1[3] = v; x1[4] = a; x1[5] = steering; x1[6] = duration;
    return ok;
}
