/* tests/native/sincos_probe.cu — TEST INFRASTRUCTURE.
 * Exhaustive check of kgmt::sincos_fast (cudasbmp_b200/csrc/kgmt_device.cuh) against libdevice's sincosf: every one of
 * the 2^32 float bit patterns whose magnitude is below the fast path's limit must give the same two results bit for bit
 * (patterns at or beyond the limit, infinities and NaNs take the library routine in propagate_edge and are skipped). */
#include <cuda_runtime.h>
#include <stdint.h>
#include "kgmt_device.cuh"

__global__ void sincos_sweep(unsigned long long* out /* [0] tested, [1] mismatches, [2] first mismatching pattern + 1 */) {
    unsigned long long tested = 0, bad = 0;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long b = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b < (1ull << 32); b += stride) {
        const float a = __uint_as_float((unsigned)b);
        if (!(fabsf(a) < kgmt::SINCOS_FAST_LIMIT)) continue;
        float s0, c0, s1, c1;
        sincosf(a, &s0, &c0);
        kgmt::sincos_fast(a, s1, c1);
        ++tested;
        if (__float_as_uint(s0) != __float_as_uint(s1) || __float_as_uint(c0) != __float_as_uint(c1)) {
            ++bad;
            atomicMin(&out[2], b + 1ull);
        }
    }
    atomicAdd(&out[0], tested);
    if (bad) atomicAdd(&out[1], bad);
}

extern "C" int sincos_probe_exhaustive(unsigned long long* host3) {
    unsigned long long* d = nullptr;
    if (cudaMalloc(&d, 24) != cudaSuccess) return 1;
    const unsigned long long init[3] = {0ull, 0ull, ~0ull};
    cudaMemcpy(d, init, 24, cudaMemcpyHostToDevice);
    sincos_sweep<<<148 * 8, 256>>>(d);
    if (cudaDeviceSynchronize() != cudaSuccess) { cudaFree(d); return 2; }
    cudaMemcpy(host3, d, 24, cudaMemcpyDeviceToHost);
    cudaFree(d);
    return 0;
}
