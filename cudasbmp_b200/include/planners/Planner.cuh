/* planners/Planner.cuh — the abstract planner interface of the reference (include/planners/Planner.cuh:6-12).
 * KGMT does not derive from it there either; kept so code that includes it still compiles. */
#pragma once
#include "agent/Agent.h"
#include "state/State.h"

class Planner {
  public:
    virtual ~Planner() = default;
    virtual void plan(float* root, float* goal) = 0;
    virtual void generateRandomTree(const float* root, const int numSamples, float** samples) = 0;
};
