/* TEST INFRASTRUCTURE ONLY (oracle/): force-included before the reference's
 * unmodified .cu files so that their `curandState` is cuRAND's own
 * Philox4x32-10 state (curand_init / curand_uniform are overloaded for it,
 * /usr/local/cuda/include/curand_kernel.h:1022-1040, curand_uniform.h:255-258).
 * This makes the reference kernels' random stream a pure function of
 * (seed, slot) and therefore reproducible by the new kernels. */
#pragma once
#include <curand_kernel.h>
#define curandState curandStatePhilox4_32_10_t
