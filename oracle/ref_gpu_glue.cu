/* TEST INFRASTRUCTURE ONLY — never linked into the product library.
 *
 * Harness around the reference's OWN, UNMODIFIED CUDA kernels
 * (/root/reference/src/planners/KGMT.cu: propagateG :341-414, propagateGV2
 * :415-482, updateR1 :487-538, updateG :540-593, findInd :319-328,
 * initCurandStates :595-600, getR1/getR2 :602-629) compiled for sm_100a by
 * oracle/Makefile with `-include shim/philox_force.h`, so that their
 * curandState is cuRAND's Philox4x32-10.  This file is
 * ours; it only launches the reference kernels on caller-supplied state, one
 * stage at a time, and copies the results back.  It is (1) the bit-exact GPU
 * parity oracle for states / flags / region indices / counters / parent links
 * and (2) "baseline A": the reference's CUDA build recompiled for B200, timed
 * with CUDA events.
 *
 * Region maps are allocated with a guard band in front because the reference
 * indexes them with r1/r2 == -1.
 */
#include "planners/KGMT.cuh"
#include <cuda_runtime.h>
#include <thrust/scan.h>
#include <thrust/execution_policy.h>
#include <thrust/device_ptr.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fcntl.h>
#include <iostream>
#include <unistd.h>
#include <cstdint>
#include <cstring>
#include <vector>

#define GUARD 64   /* ints in front of every map */

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { g_err = e_; return -(int)e_ - 1000; } } while (0)
static cudaError_t g_err = cudaSuccess;

template <typename T> static T* dalloc(size_t n) { T* p = nullptr; cudaMalloc(&p, sizeof(T) * (n ? n : 1)); return p; }

extern "C" {

const char* ref_gpu_last_error(void) { return cudaGetErrorString(g_err); }

int ref_getR1(float x, float y, float R1Size, int N) { return getR1(x, y, R1Size, N); }
int ref_getR2(float x, float y, int r1, float R1Size, int N, float R2Size, int n) {
    return getR2(x, y, r1, R1Size, N, R2Size, n);
}

/* One launch of the reference's propagateG (mode 1) or propagateGV2 (mode 2)
 * from caller-supplied state.  All pointers are HOST pointers.
 *   treeSamples [treeCount][7]; frontier[activeSize] = tree indices to expand;
 *   maps (in/out): R1,R1Valid,R1Invalid,R1Avail [N*N]; R2,R2Valid,R2Invalid,R2Avail [N*N*n*n];
 *   R1Score [N*N] (in); obstacles [K][4];
 *   outputs: unexplored [M][7], uParentIdx [M], GNew [M] (0/1),  M = activeSize*children.
 * RNG: initCurandStates(states, M, seedKey) first, i.e. slot s draws from
 * Philox(ctr=(0,0,s,0), key=(seedKey,0)).
 * reps>1 repeats the propagate launch for timing (maps are then polluted);
 * *ms receives the mean kernel time. */
int ref_gpu_expand(int mode, int children,
                   const float* treeSamples, int treeCount, const int* frontier, int activeSize,
                   int* R1, int* R2, int* R1Valid, int* R2Valid, int* R1Invalid, int* R2Invalid,
                   int* R1Avail, int* R2Avail, const float* R1Score,
                   int N, int n, float R1Size, float R2Size, int numDisc, float agentLength,
                   const float* obstacles, int K, float width, float height, int seedKey,
                   float* unexplored, int* uParentIdx, uint8_t* GNewOut, int reps, float* ms) {
    const int c1 = N * N, c2 = c1 * n * n;
    const long M = (long)activeSize * children;
    const int threadsV2 = ((activeSize + 127) / 128) * 128;
    const int idxCount = (mode == 2) ? threadsV2 : activeSize;

    float* d_tree = dalloc<float>((size_t)(treeCount + 1) * 7);
    bool*  d_G    = dalloc<bool>((size_t)treeCount + 1);
    bool*  d_GNew = dalloc<bool>((size_t)M);
    int*   d_idx  = dalloc<int>((size_t)idxCount);
    float* d_unx  = dalloc<float>((size_t)M * 7);
    int*   d_upar = dalloc<int>((size_t)M);
    int*   d_maps = dalloc<int>((size_t)8 * GUARD + 4 * (size_t)c1 + 4 * (size_t)c2);
    float* d_score = dalloc<float>((size_t)c1);
    float* d_thr  = dalloc<float>(1);
    float* d_obs  = dalloc<float>((size_t)4 * (K > 0 ? K : 1));
    curandState* d_states = dalloc<curandState>((size_t)M);

    int* m = d_maps;
    int* dR1 = (m += GUARD); m += c1;
    int* dR1V = (m += GUARD); m += c1;
    int* dR1I = (m += GUARD); m += c1;
    int* dR1A = (m += GUARD); m += c1;
    int* dR2 = (m += GUARD); m += c2;
    int* dR2V = (m += GUARD); m += c2;
    int* dR2I = (m += GUARD); m += c2;
    int* dR2A = (m += GUARD); m += c2;

    CK(cudaMemset(d_maps, 0, sizeof(int) * ((size_t)8 * GUARD + 4 * (size_t)c1 + 4 * (size_t)c2)));
    CK(cudaMemset(d_tree, 0, sizeof(float) * (size_t)(treeCount + 1) * 7));
    CK(cudaMemcpy(d_tree, treeSamples, sizeof(float) * (size_t)treeCount * 7, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_G, 0, (size_t)treeCount + 1));
    CK(cudaMemset(d_GNew, 0, (size_t)M));
    {
        std::vector<int> idx((size_t)idxCount, treeCount);   /* padding -> dummy node whose G is false */
        std::vector<uint8_t> g((size_t)treeCount + 1, 0);
        for (int i = 0; i < activeSize; ++i) { idx[i] = frontier[i]; g[frontier[i]] = 1; }
        CK(cudaMemcpy(d_idx, idx.data(), sizeof(int) * (size_t)idxCount, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_G, g.data(), (size_t)treeCount + 1, cudaMemcpyHostToDevice));
    }
    CK(cudaMemcpy(dR1, R1, sizeof(int) * c1, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dR1V, R1Valid, sizeof(int) * c1, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dR1I, R1Invalid, sizeof(int) * c1, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dR1A, R1Avail, sizeof(int) * c1, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dR2, R2, sizeof(int) * c2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dR2V, R2Valid, sizeof(int) * c2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dR2I, R2Invalid, sizeof(int) * c2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dR2A, R2Avail, sizeof(int) * c2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_score, R1Score, sizeof(float) * c1, cudaMemcpyHostToDevice));
    if (K > 0) CK(cudaMemcpy(d_obs, obstacles, sizeof(float) * 4 * (size_t)K, cudaMemcpyHostToDevice));

    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float total = 0.f;
    if (reps < 1) reps = 1;
    for (int r = 0; r < reps; ++r) {
        initCurandStates<<<(int)((M + 127) / 128), 128>>>(d_states, (int)M, seedKey);
        if (mode == 2 && r > 0) {   /* V2 test-and-clears G: re-arm for timing repeats */
            std::vector<uint8_t> g((size_t)treeCount + 1, 0);
            for (int i = 0; i < activeSize; ++i) g[frontier[i]] = 1;
            cudaMemcpy(d_G, g.data(), (size_t)treeCount + 1, cudaMemcpyHostToDevice);
        }
        cudaEventRecord(e0);
        if (mode == 2) {
            propagateGV2<<<threadsV2 / 128, 128>>>(activeSize, d_idx, d_G, d_GNew, d_tree, d_unx, d_upar,
                dR1V, dR2V, dR1I, dR2I, dR1, dR2, dR1A, dR2A, N, n, R1Size, R2Size, d_states, numDisc,
                agentLength, d_thr, d_score, d_obs, K, children, width, height);
        } else {
            propagateG<<<activeSize, 32>>>(activeSize, d_idx, d_G, d_GNew, d_tree, d_unx, d_upar,
                dR1V, dR2V, dR1I, dR2I, dR1, dR2, dR1A, dR2A, N, n, R1Size, R2Size, d_states, numDisc,
                agentLength, d_thr, d_score, d_obs, K, width, height);
        }
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float t = 0.f; cudaEventElapsedTime(&t, e0, e1); total += t;
    }
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    if (ms) *ms = total / reps;

    if (unexplored) CK(cudaMemcpy(unexplored, d_unx, sizeof(float) * (size_t)M * 7, cudaMemcpyDeviceToHost));
    if (uParentIdx) CK(cudaMemcpy(uParentIdx, d_upar, sizeof(int) * (size_t)M, cudaMemcpyDeviceToHost));
    if (GNewOut)    CK(cudaMemcpy(GNewOut, d_GNew, (size_t)M, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(R1, dR1, sizeof(int) * c1, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(R1Valid, dR1V, sizeof(int) * c1, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(R1Invalid, dR1I, sizeof(int) * c1, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(R1Avail, dR1A, sizeof(int) * c1, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(R2, dR2, sizeof(int) * c2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(R2Valid, dR2V, sizeof(int) * c2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(R2Invalid, dR2I, sizeof(int) * c2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(R2Avail, dR2A, sizeof(int) * c2, cudaMemcpyDeviceToHost));

    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_tree); cudaFree(d_G); cudaFree(d_GNew); cudaFree(d_idx); cudaFree(d_unx); cudaFree(d_upar);
    cudaFree(d_maps); cudaFree(d_score); cudaFree(d_thr); cudaFree(d_obs); cudaFree(d_states);
    return 0;
}

/* The reference's updateR1 (hard-wired to N == 16, one block of 256). */
int ref_gpu_scores(const int* R1Avail, const int* R2Avail, const int* R1Valid, const int* R1Invalid,
                   const int* R1, int n, float R2Size, float* R1Score, float* R1Threshold) {
    const int c1 = 256, c2 = c1 * n * n;
    int *dA = dalloc<int>(c1), *dA2 = dalloc<int>(c2), *dV = dalloc<int>(c1), *dI = dalloc<int>(c1), *dR = dalloc<int>(c1);
    float *dS = dalloc<float>(c1), *dT = dalloc<float>(1);
    CK(cudaMemcpy(dA, R1Avail, 4 * c1, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dA2, R2Avail, 4 * (size_t)c2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dV, R1Valid, 4 * c1, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dI, R1Invalid, 4 * c1, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dR, R1, 4 * c1, cudaMemcpyHostToDevice));
    int active = 0; for (int i = 0; i < c1; ++i) active += (R1Avail[i] != 0);
    updateR1<<<1, c1>>>(dS, dA, dA2, dV, dI, dR, n, 0.01, R2Size * R2Size, dT, active);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(R1Score, dS, 4 * c1, cudaMemcpyDeviceToHost));
    if (R1Threshold) CK(cudaMemcpy(R1Threshold, dT, 4, cudaMemcpyDeviceToHost));
    cudaFree(dA); cudaFree(dA2); cudaFree(dV); cudaFree(dI); cudaFree(dR); cudaFree(dS); cudaFree(dT);
    return 0;
}

/* The reference's insertion stage: exclusive_scan(GNew) + findInd + updateG
 * (KGMT.cu:222-245).  cap = array capacity (the reference's maxTreeSize);
 * GNew[cap], unexplored[cap][7], uParentIdx[cap]; tree arrays [cap].
 * Returns accepted count (>= 0) or a negative error. */
int ref_gpu_insert(int cap, const uint8_t* GNew, const float* unexplored, const int* uParentIdx,
                   int treeSize, float* treeSamples, int* treeParentIdx, float* costs, uint8_t* G,
                   const float* goal7, float r, float* costToGoal) {
    bool* dGNew = dalloc<bool>(cap); bool* dG = dalloc<bool>(cap);
    int* dScan = dalloc<int>(cap); int* dIdx = dalloc<int>(cap); int* dUP = dalloc<int>(cap); int* dTP = dalloc<int>(cap);
    float* dUnx = dalloc<float>((size_t)cap * 7); float* dTree = dalloc<float>((size_t)cap * 7);
    float* dCost = dalloc<float>(cap); float* dGoal = dalloc<float>(7); float* dCTG = dalloc<float>(1);
    CK(cudaMemcpy(dGNew, GNew, cap, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dG, G, cap, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dUP, uParentIdx, 4 * (size_t)cap, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dTP, treeParentIdx, 4 * (size_t)cap, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dUnx, unexplored, 28 * (size_t)cap, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dTree, treeSamples, 28 * (size_t)cap, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dCost, costs, 4 * (size_t)cap, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dGoal, goal7, 28, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dCTG, costToGoal, 4, cudaMemcpyHostToDevice));
    thrust::exclusive_scan(thrust::device, thrust::device_pointer_cast(dGNew), thrust::device_pointer_cast(dGNew) + cap,
                           thrust::device_pointer_cast(dScan), 0, thrust::plus<int>());
    int last = 0; uint8_t lastFlag = GNew[cap - 1];
    CK(cudaMemcpy(&last, dScan + cap - 1, 4, cudaMemcpyDeviceToHost));
    int accepted = last + (lastFlag ? 1 : 0);
    findInd<<<(cap + 127) / 128, 128>>>(cap, dGNew, dScan, dIdx);
    int grid = accepted < cap / 32 ? accepted : cap / 32;
    if (grid > 0)
        updateG<<<grid, 32>>>(dTree, dUnx, dUP, dTP, dG, dGNew, dIdx, accepted, treeSize, dCost, dGoal, r, dCTG);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(treeSamples, dTree, 28 * (size_t)cap, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(treeParentIdx, dTP, 4 * (size_t)cap, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(costs, dCost, 4 * (size_t)cap, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(G, dG, cap, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(costToGoal, dCTG, 4, cudaMemcpyDeviceToHost));
    cudaFree(dGNew); cudaFree(dG); cudaFree(dScan); cudaFree(dIdx); cudaFree(dUP); cudaFree(dTP);
    cudaFree(dUnx); cudaFree(dTree); cudaFree(dCost); cudaFree(dGoal); cudaFree(dCTG);
    return accepted;
}

/* The reference planner end to end (its own KGMT::plan, XORWOW, time(NULL)
 * seed): timing baseline only.  obstacles = host [K][4].  planMs (may be NULL)
 * receives {wall clock around plan() incl. its 13 CSV dumps, the reference's OWN
 * figure "time inside KGMT" (std::clock() around its loop, KGMT.cu:82,294-295,
 * parsed from what plan() prints), in milliseconds}.  plan() runs in a scratch
 * directory with stdout redirected to a file, so its CSVs and prints do not
 * land in the caller's. */
int ref_gpu_plan(float width, float height, int N, int n, int numIterations, int maxTreeSize, int numDisc,
                 float agentLength, float goalThreshold, const float* init7, const float* goal7,
                 const float* obstacles, int K, int* treeSizeOut, float* costOut, double* planMs) {
    char dirT[] = "/tmp/refplan_XXXXXX";
    char cwd[4096];
    const bool moved = getcwd(cwd, sizeof(cwd)) && mkdtemp(dirT) && chdir(dirT) == 0;
    fflush(stdout); std::cout.flush();
    const int saved = dup(1);
    const int fd = open("stdout.txt", O_CREAT | O_TRUNC | O_WRONLY, 0600);
    if (fd >= 0) dup2(fd, 1);
    float* d_obs = dalloc<float>((size_t)4 * (K > 0 ? K : 1));
    if (K > 0) CK(cudaMemcpy(d_obs, obstacles, sizeof(float) * 4 * (size_t)K, cudaMemcpyHostToDevice));
    float i7[7], g7[7]; memcpy(i7, init7, 28); memcpy(g7, goal7, 28);
    {
        KGMT kgmt(width, height, N, n, numIterations, maxTreeSize, numDisc, agentLength, goalThreshold);
        cudaMemset(kgmt.d_costToGoal, 0, sizeof(float));   /* the reference never initialises it (App. B #4) */
        cudaDeviceSynchronize();
        const auto t0 = std::chrono::steady_clock::now();
        kgmt.plan(i7, g7, d_obs, K);
        cudaDeviceSynchronize();
        const auto t1 = std::chrono::steady_clock::now();
        if (planMs) { planMs[0] = std::chrono::duration<double, std::milli>(t1 - t0).count(); planMs[1] = -1.0; }
        if (treeSizeOut) *treeSizeOut = kgmt.treeSize_;
        if (costOut) *costOut = kgmt.costToGoal_;
    }
    cudaFree(d_obs);
    fflush(stdout); std::cout.flush();
    if (saved >= 0) { dup2(saved, 1); close(saved); }
    if (fd >= 0) close(fd);
    if (planMs) {
        FILE* f = fopen("stdout.txt", "r");
        if (f) {
            char line[512];
            while (fgets(line, sizeof(line), f)) {
                const char* p = strstr(line, "time inside KGMT is ");
                if (p) planMs[1] = atof(p + 20) * 1e3;
            }
            fclose(f);
        }
    }
    if (moved) {
        const char* names[] = {"samples.csv", "unexploredSamples.csv", "parentRelations.csv", "uParentIdx.csv", "G.csv", "R2Avail.csv",
                               "R1Avail.csv", "R1Valid.csv", "R2Valid.csv", "R1Invalid.csv", "R2Invalid.csv", "R1Score.csv", "R1.csv", "stdout.txt"};
        for (const char* nm : names) unlink(nm);
        if (chdir(cwd) != 0) return -1;
        rmdir(dirT);
    }
    return 0;
}

} /* extern "C" */
