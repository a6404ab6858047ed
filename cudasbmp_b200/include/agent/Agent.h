/* agent/Agent.h — kinematic-bicycle agent of the reference (include/agent/Agent.h:6-25, src/agent/Agent.cpp:4-25),
 * header-only.  Host-side bookkeeping only; the planner integrates the same model on the device
 * (csrc/kgmt_device.cuh: propagate_edge).  Eigen is used when the toolchain has it, otherwise a two-double stand-in. */
#pragma once
#include <cmath>
#include <vector>
#if defined(__has_include) && __has_include(<Eigen/Core>)
#include <Eigen/Core>
#else
#include "compat/Eigen/Core"
#endif

class Agent {
  public:
    Agent() = default;
    explicit Agent(std::vector<Eigen::Vector2d>& verticesCCW) : verticesCCW_(verticesCCW) {}
    Agent(float x, float y, float length, float theta = 0.0f, float v = 0.0f)
        : x_(x), y_(y), theta_(theta), v_(v), length_(length) {
        const double x0 = x, y0 = y, x1 = x + length, y1 = y + length;      /* square footprint, CCW */
        verticesCCW_ = {Eigen::Vector2d(x0, y0), Eigen::Vector2d(x1, y0), Eigen::Vector2d(x1, y1), Eigen::Vector2d(x0, y1)};
    }
    /* one explicit Euler step of the bicycle model (Agent.cpp:19-25) */
    void updateState(float a, float delta, float dt) {
        const float c = std::cos(theta_), s = std::sin(theta_);
        x_ += v_ * c * dt;
        y_ += v_ * s * dt;
        theta_ += (v_ / length_) * std::tan(delta) * dt;
        v_ += a * dt;
    }
    std::vector<Eigen::Vector2d> verticesCCW_;
    float x_ = 0, y_ = 0, theta_ = 0, v_ = 0, length_ = 1, width_ = 0;
};
