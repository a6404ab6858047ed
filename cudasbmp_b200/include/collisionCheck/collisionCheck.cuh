/* collisionCheck/collisionCheck.cuh — the reference's device collision API (include/collisionCheck/collisionCheck.cuh:4-8)
 * as inline device functions.  Same arguments, same answers (src/collisionCheck/collisionCheck.cu:6-28): "valid" means no
 * overlap; touching boxes do not collide.  The planner itself does not call these — its fused kernel tests float4
 * obstacles staged in shared memory (csrc/kgmt_device.cuh) — they exist for code written against the reference API. */
#pragma once
#include <cuda_runtime.h>

__device__ inline bool isBroadPhaseValid(float* bbMin, float* bbMax, float* obs) {
    const bool overlap = (bbMax[0] > obs[0]) & (obs[2] > bbMin[0]) & (bbMax[1] > obs[1]) & (obs[3] > bbMin[1]);
    return !overlap;
}

__device__ inline bool isMotionValid(float* /*x0*/, float* /*x1*/, float* bbMin, float* bbMax, float* obstacles,
                                     int obstaclesCount) {
    for (int k = 0; k < obstaclesCount; ++k)
        if (!isBroadPhaseValid(bbMin, bbMax, obstacles + 4 * k)) return false;
    return true;
}
