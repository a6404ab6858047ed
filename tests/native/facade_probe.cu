/* tests/native/facade_probe.cu — TEST INFRASTRUCTURE.
 *
 * Calls the reference-compatible DEVICE functions that cudasbmp_b200/include re-provides
 * (statePropagator/statePropagator.cuh: propagateAndCheck; collisionCheck/collisionCheck.cuh: isMotionValid,
 * isBroadPhaseValid — the reference's include/statePropagator/statePropagator.cuh:5-14 and
 * include/collisionCheck/collisionCheck.cuh:4-8) from __global__ kernels, so that tests/test_gpu_refkernels.py can
 * compare them bit for bit with the reference's own kernels (oracle/_ref/libref_gpu.so).  Compiled with
 * `-include oracle/shim/philox_force.h` like the reference build, so curandState is cuRAND's Philox4x32-10 and candidate
 * slot s of key k draws the same uniforms as initCurandStates(states, M, k) gives propagateG (KGMT.cu:595-600).
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "statePropagator/statePropagator.cuh"

__global__ void probe_propagate(const float* parents7, int children, long M, int numDisc, float L, float* obstacles, int K,
                                float W, float H, unsigned long long key, float* x1out, unsigned char* validOut) {
    const long s = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= M) return;
    curandState st;
    curand_init(key, (unsigned long long)s, 0, &st);                 /* initCurandStates, KGMT.cu:595-600 */
    float x0[7], x1[7];
    for (int j = 0; j < 7; ++j) x0[j] = parents7[(s / children) * 7 + j];
    const bool ok = propagateAndCheck(x0, x1, numDisc, L, &st, obstacles, K, W, H);
    for (int j = 0; j < 7; ++j) x1out[s * 7 + j] = x1[j];
    validOut[s] = ok ? 1 : 0;
}

__global__ void probe_motion(const float* boxes4, long B, float* obstacles, int K, unsigned char* motionValid,
                             unsigned char* firstBroad) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    float lo[2] = {boxes4[i * 4], boxes4[i * 4 + 1]}, hi[2] = {boxes4[i * 4 + 2], boxes4[i * 4 + 3]};
    motionValid[i] = isMotionValid(lo, hi, lo, hi, obstacles, K) ? 1 : 0;
    firstBroad[i] = (K > 0 && isBroadPhaseValid(lo, hi, obstacles)) ? 1 : 0;
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return -(int)e_ - 1000; } while (0)

extern "C" {

/* all pointers HOST; x1out [M][7], validOut [M] */
int facade_probe_propagate(const float* parents7, int P, int children, int numDisc, float L, const float* obstacles, int K,
                           float W, float H, unsigned key, float* x1out, unsigned char* validOut) {
    const long M = (long)P * children;
    float *dP = nullptr, *dO = nullptr, *dX = nullptr; unsigned char* dV = nullptr;
    CK(cudaMalloc(&dP, (size_t)P * 28)); CK(cudaMalloc(&dO, (size_t)(K > 0 ? K : 1) * 16));
    CK(cudaMalloc(&dX, (size_t)M * 28)); CK(cudaMalloc(&dV, (size_t)M));
    CK(cudaMemcpy(dP, parents7, (size_t)P * 28, cudaMemcpyHostToDevice));
    if (K > 0) CK(cudaMemcpy(dO, obstacles, (size_t)K * 16, cudaMemcpyHostToDevice));
    probe_propagate<<<(unsigned)((M + 127) / 128), 128>>>(dP, children, M, numDisc, L, dO, K, W, H, key, dX, dV);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(x1out, dX, (size_t)M * 28, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(validOut, dV, (size_t)M, cudaMemcpyDeviceToHost));
    cudaFree(dP); cudaFree(dO); cudaFree(dX); cudaFree(dV);
    return 0;
}

/* boxes4 [B][4] = (minx, miny, maxx, maxy); motionValid[B] = isMotionValid vs all K; firstBroad[B] = isBroadPhaseValid vs obstacle 0 */
int facade_probe_motion(const float* boxes4, long B, const float* obstacles, int K, unsigned char* motionValid,
                        unsigned char* firstBroad) {
    float *dB = nullptr, *dO = nullptr; unsigned char *dM = nullptr, *dF = nullptr;
    CK(cudaMalloc(&dB, (size_t)B * 16)); CK(cudaMalloc(&dO, (size_t)(K > 0 ? K : 1) * 16));
    CK(cudaMalloc(&dM, (size_t)B)); CK(cudaMalloc(&dF, (size_t)B));
    CK(cudaMemcpy(dB, boxes4, (size_t)B * 16, cudaMemcpyHostToDevice));
    if (K > 0) CK(cudaMemcpy(dO, obstacles, (size_t)K * 16, cudaMemcpyHostToDevice));
    probe_motion<<<(unsigned)((B + 127) / 128), 128>>>(dB, B, dO, K, dM, dF);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(motionValid, dM, (size_t)B, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(firstBroad, dF, (size_t)B, cudaMemcpyDeviceToHost));
    cudaFree(dB); cudaFree(dO); cudaFree(dM); cudaFree(dF);
    return 0;
}

}
