/* TEST INFRASTRUCTURE ONLY (oracle/): stand-in for <curand_kernel.h> when the
 * reference's statePropagator.cu is built for the host.  curandState becomes a
 * host Philox4x32-10 stream with cuRAND's stateful semantics
 * (/usr/local/cuda/include/curand_kernel.h:888-915,1022-1040); the generator
 * itself lives in oracle/ref_host_glue.cpp. */
#pragma once
#include <cmath>
#include <cstdint>
#include <math.h>
struct curandState {
    uint32_t ctr[4];
    uint32_t key[2];
    uint32_t out[4];
    int      pos;      /* next word of out[] to hand out; 4 => regenerate */
};
float curand_uniform(curandState* s);
