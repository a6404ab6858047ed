/* helper/helper.cuh — the I/O glue of the reference (include/helper/helper.cuh:19-79, src/helper/helper.cu:3-34),
 * header-only: CUDA_ERROR_CHECK, obstacle CSV reader, CSV writers in the reference's "%.10f" fixed format. */
#pragma once
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include <thrust/device_vector.h>
#include <thrust/host_vector.h>

#define CUDA_ERROR_CHECK(expr)                                                        \
    do {                                                                              \
        const cudaError_t kgmt_err_ = (expr);                                         \
        if (kgmt_err_ != cudaSuccess) {                                               \
            printf("CUDA call failed!\n\t%s\n", cudaGetErrorString(kgmt_err_));      \
            exit(1);                                                                  \
        }                                                                             \
    } while (0)

/* rows of 2*workspaceDim comma-separated floats: (minx, miny, maxx, maxy); blank lines are skipped */
inline std::vector<float> readObstaclesFromCSV(const std::string& filename, int& numObstacles, int workspaceDim) {
    std::ifstream in(filename);
    if (!in) { std::cerr << "Error opening file: " << filename << std::endl; exit(1); }
    std::vector<float> values;
    std::string row, field;
    while (std::getline(in, row)) {
        std::stringstream fields(row);
        while (std::getline(fields, field, ',')) {
            std::stringstream one(field);
            float v;
            if (one >> v) values.push_back(v);
        }
    }
    numObstacles = (int)(values.size() / (size_t)(2 * workspaceDim));
    return values;
}

template <typename T>
void writeVectorToCSV(const thrust::host_vector<T>& vec, const std::string& filename, int rows, int cols) {
    std::ofstream out(filename);
    out << std::fixed << std::setprecision(10);
    for (int r = 0; r < rows; ++r) {
        for (int c = 0; c < cols; ++c) out << vec[(size_t)r * cols + c] << (c + 1 < cols ? "," : "");
        out << std::endl;
    }
}

template <typename T>
void copyAndWriteVectorToCSV(const thrust::device_vector<T>& d_vec, const std::string& filename, int rows, int cols) {
    thrust::host_vector<T> h_vec = d_vec;
    writeVectorToCSV(h_vec, filename, rows, cols);
}

template <typename T>
void printDeviceVector(const T* d_ptr, int size) {
    std::vector<T> h((size_t)size);
    cudaMemcpy(h.data(), d_ptr, sizeof(T) * (size_t)size, cudaMemcpyDeviceToHost);
    for (const T& v : h) std::cout << v << " ";
    std::cout << std::endl;
}

__device__ inline void printSample(float* x, int sampleDim) {
    for (int i = 0; i < sampleDim; ++i) printf("%f ", x[i]);
    printf("\n");
}
