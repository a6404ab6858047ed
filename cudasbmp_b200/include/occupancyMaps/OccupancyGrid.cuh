/* occupancyMaps/OccupancyGrid.cuh — the reference's N x N occupancy grid class
 * (include/occupancyMaps/OccupancyGrid.cuh:7-25, src/occupancyMaps/OccupancyGrid.cu:6-35), header-only.
 * KGMT never instantiates it (there or here): the live occupancy maps are the R1 / R2 arrays inside the planner context
 * (kgmt_export ids 5..13).  getCellIndex is the same function as getR1 (KGMT.cu:602-609). */
#pragma once
#include <cuda_runtime.h>
#include <thrust/device_vector.h>

class OccupancyGrid {
  public:
    OccupancyGrid() = default;
    OccupancyGrid(float width, float height, int N)
        : width_(width), height_(height), cellSize_(width / N), N_(N), grid_((size_t)N * N, 0) {}

    __host__ __device__ int getCellIndex(float x, float y) const {
        const int cx = static_cast<int>(x / cellSize_), cy = static_cast<int>(y / cellSize_);
        return (cx >= 0 && cx < N_ && cy >= 0 && cy < N_) ? cy * N_ + cx : -1;
    }
    int getOccupancy(int row, int col) const {
        const int cell = getCellIndex((float)row, (float)col);
        return cell < 0 ? -1 : (int)grid_[cell];
    }
    void updateOccupancy(int row, int col, int n) {
        const int cell = getCellIndex((float)row, (float)col);
        if (cell >= 0) grid_[cell] = grid_[cell] + n;
    }

    float width_ = 0, height_ = 0, cellSize_ = 0;
    int N_ = 0;
    thrust::device_vector<int> grid_;
};
