/* scripts/native/issue_rate_bench.cu — measured instruction-issue roofs of the B200 SM for the instruction classes the
 * KGMT expansion kernel is made of ("verify with a microbenchmark on the box"; VERDICT r01 "next" #4).
 *
 *   ffma        dependent FFMA chains, 8 per thread                      -> lane-ops / clk / SM of the FMA pipe
 *   fsetp       FSETP.GT(.AND) chains: 4 per AABB overlap test           -> compares / clk / SM
 *   aabb        the overlap test as the kernel executes it: 4 FSETP + the OR into the running flag (PLOP3)
 *   fmnmx       FMNMX (step bbox min/max)
 *   imad        IMAD / integer multiply-add (Philox rounds)
 *   lds128      LDS.128 broadcast (obstacle float4 out of shared memory)
 *   mix         FSETP and FFMA interleaved 4:1 — do the ALU and FMA pipes issue side by side?
 *
 * Every figure is computed from %clock64 deltas taken on the SM itself (so it is independent of the clock the box runs
 * at) and from CUDA-event time (to report the clock and the whole-GPU rate).  Prints ONE JSON object.
 * Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o issue_rate_bench issue_rate_bench.cu
 */
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

#define ITERS 4096
#define THREADS 256

template <int KIND>
__global__ void __launch_bounds__(THREADS) bench(float* out, long long* cyc, float seed) {
    __shared__ float4 sh[64];
    if (threadIdx.x < 64) sh[threadIdx.x] = make_float4(seed + threadIdx.x, seed * 2.f, seed + 100.f + threadIdx.x, seed * 3.f);
    __syncthreads();
    float a0 = seed + threadIdx.x, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float m = 1.0000001f, c = seed;
    int acc = 0;
    unsigned u0 = threadIdx.x * 2654435761u + 1u, u1 = u0 ^ 0x9E3779B9u, u2 = u0 + 77u, u3 = u1 + 99u;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        if (KIND == 0) {          /* 8 FFMA */
            a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
            a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
        } else if (KIND == 1) {   /* 8 FSETP (two chains of 4, the shape of two AABB tests) + 2 predicated adds */
            asm volatile("{ .reg .pred p, q;\n\t"
                         "setp.gt.f32 p, %1, %2;\n\t setp.gt.and.f32 p, %3, %4, p;\n\t setp.gt.and.f32 p, %5, %6, p;\n\t setp.gt.and.f32 p, %7, %8, p;\n\t"
                         "setp.gt.f32 q, %2, %1;\n\t setp.gt.and.f32 q, %4, %3, q;\n\t setp.gt.and.f32 q, %6, %5, q;\n\t setp.gt.and.f32 q, %8, %7, q;\n\t"
                         "@p add.s32 %0, %0, 1;\n\t @q add.s32 %0, %0, 2;\n\t }"
                         : "+r"(acc) : "f"(a0), "f"(a1), "f"(a2), "f"(a3), "f"(a4), "f"(a5), "f"(a6), "f"(a7));
            a0 += 1.f;            /* one FADD so the compares are not loop-invariant */
        } else if (KIND == 2) {   /* the overlap test of kgmt_device.cuh, four obstacles from shared memory per trip */
            const float4 o0 = sh[(it) & 63], o1 = sh[(it + 1) & 63], o2 = sh[(it + 2) & 63], o3 = sh[(it + 3) & 63];
            const bool h0 = (a2 > o0.x) & (o0.z > a0) & (a3 > o0.y) & (o0.w > a1);
            const bool h1 = (a2 > o1.x) & (o1.z > a0) & (a3 > o1.y) & (o1.w > a1);
            const bool h2 = (a2 > o2.x) & (o2.z > a0) & (a3 > o2.y) & (o2.w > a1);
            const bool h3 = (a2 > o3.x) & (o3.z > a0) & (a3 > o3.y) & (o3.w > a1);
            acc += (h0 | h1 | h2 | h3) ? 1 : 0;
            a0 += 0.001f;
        } else if (KIND == 3) {   /* 8 FMNMX */
            a0 = fminf(a0, a1); a1 = fmaxf(a1, a2); a2 = fminf(a2, a3); a3 = fmaxf(a3, a4);
            a4 = fminf(a4, a5); a5 = fmaxf(a5, a6); a6 = fminf(a6, a7); a7 = fmaxf(a7, a0 + c);
        } else if (KIND == 4) {   /* 8 IMAD-class (4 mul.lo + 4 mul.hi, the Philox round) */
            const unsigned h0 = __umulhi(0xD2511F53u, u0), l0 = 0xD2511F53u * u0, h1 = __umulhi(0xCD9E8D57u, u2), l1 = 0xCD9E8D57u * u2;
            const unsigned n0 = h1 ^ u1 ^ it, n2 = h0 ^ u3 ^ it;
            u0 = n0; u1 = l1; u2 = n2; u3 = l0;
            const unsigned g0 = __umulhi(0xD2511F53u, u0), m0 = 0xD2511F53u * u0, g1 = __umulhi(0xCD9E8D57u, u2), m1 = 0xCD9E8D57u * u2;
            u0 = g1 ^ u1; u1 = m1; u2 = g0 ^ u3; u3 = m0;
        } else if (KIND == 5) {   /* 8 LDS.128, broadcast addresses */
            const float4 o0 = sh[(it) & 63], o1 = sh[(it + 1) & 63], o2 = sh[(it + 2) & 63], o3 = sh[(it + 3) & 63];
            const float4 o4 = sh[(it + 4) & 63], o5 = sh[(it + 5) & 63], o6 = sh[(it + 6) & 63], o7 = sh[(it + 7) & 63];
            a0 += o0.x + o1.y; a1 += o2.z + o3.w; a2 += o4.x + o5.y; a3 += o6.z + o7.w;
        } else {                  /* 8 FSETP + 2 FFMA */
            asm volatile("{ .reg .pred p, q;\n\t"
                         "setp.gt.f32 p, %1, %2;\n\t setp.gt.and.f32 p, %3, %4, p;\n\t setp.gt.and.f32 p, %5, %6, p;\n\t setp.gt.and.f32 p, %7, %8, p;\n\t"
                         "setp.gt.f32 q, %2, %1;\n\t setp.gt.and.f32 q, %4, %3, q;\n\t setp.gt.and.f32 q, %6, %5, q;\n\t setp.gt.and.f32 q, %8, %7, q;\n\t"
                         "@p add.s32 %0, %0, 1;\n\t @q add.s32 %0, %0, 2;\n\t }"
                         : "+r"(acc) : "f"(a0), "f"(a1), "f"(a2), "f"(a3), "f"(a4), "f"(a5), "f"(a6), "f"(a7));
            a0 = fmaf(a0, m, c); a4 = fmaf(a4, m, c);
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * THREADS + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + (float)acc + (float)(u0 ^ u1 ^ u2 ^ u3);
}

typedef void (*kfn)(float*, long long*, float);

int main() {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) { printf("{\"error\": \"no CUDA device\"}\n"); return 1; }
    const int sms = prop.multiProcessorCount, perSM = 8, grid = sms * perSM;      /* 64 warps per SM */
    float* out; long long* cyc;
    cudaMalloc(&out, (size_t)grid * THREADS * 4); cudaMalloc(&cyc, (size_t)grid * 8);
    const char* names[7] = {"ffma", "fsetp", "aabb", "fmnmx", "imad", "lds128", "mix_fsetp8_ffma2"};
    const double opsPerIter[7] = {8, 8, 4, 8, 8, 8, 10};       /* counted instructions (aabb: overlap tests) per thread-iteration */
    kfn fns[7] = {bench<0>, bench<1>, bench<2>, bench<3>, bench<4>, bench<5>, bench<6>};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    printf("{\"device\": \"%s\", \"sms\": %d, \"threads_per_sm\": %d, \"iters\": %d", prop.name, sms, perSM * THREADS, ITERS);
    std::vector<long long> h(grid);
    for (int k = 0; k < 7; ++k) {
        for (int w = 0; w < 3; ++w) fns[k]<<<grid, THREADS>>>(out, cyc, 1.5f);
        cudaDeviceSynchronize();
        float best = 1e30f; double bestCyc = 0;
        for (int r = 0; r < 10; ++r) {
            cudaEventRecord(e0);
            fns[k]<<<grid, THREADS>>>(out, cyc, 1.5f);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) {
                best = ms;
                cudaMemcpy(h.data(), cyc, (size_t)grid * 8, cudaMemcpyDeviceToHost);
                double s = 0; for (int i = 0; i < grid; ++i) s += (double)h[i];
                bestCyc = s / grid;
            }
        }
        /* per SM: perSM CTAs resident at once run concurrently for ~bestCyc cycles each */
        const double laneOps = opsPerIter[k] * ITERS * THREADS * perSM;
        const double perClkSM = laneOps / bestCyc;
        const double gpuRate = opsPerIter[k] * (double)ITERS * THREADS * grid / (best * 1e-3);
        printf(", \"%s\": {\"lane_ops_per_clk_per_sm\": %.2f, \"warp_inst_per_clk_per_sm\": %.3f, \"gpu_lane_ops_per_s\": %.4g, \"ms\": %.4f, \"sm_cycles\": %.0f, \"implied_mhz\": %.0f}",
               names[k], perClkSM, perClkSM / 32.0, gpuRate, best, bestCyc, bestCyc / (best * 1e-3) / 1e6);
    }
    printf("}\n");
    return 0;
}
