/* state/State.h — the 7-float sample (x, y, theta, v, a, steering, duration) of the reference
 * (include/state/State.h:6-20): the row layout of samples.csv and of kgmt_export(KGMT_ARR_SAMPLES). */
#pragma once

class State {
  public:
    State() = default;
    State(float x, float y, float theta = 0.0f, float v = 0.0f, float a = 0.0f, float u = 0.0f, float dt = 0.0f)
        : x_(x), y_(y), theta_(theta), v_(v), a_(a), u_(u), dt_(dt) {}
    float x_ = 0, y_ = 0, theta_ = 0, v_ = 0;   /* state */
    float a_ = 0, u_ = 0, dt_ = 0;              /* control that led here: acceleration, steering, duration */
};
