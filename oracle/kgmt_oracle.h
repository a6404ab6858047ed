/* kgmt_oracle.h — TEST INFRASTRUCTURE ONLY (see kgmt_oracle.c header).
 * CPU restatement of the reference KGMT expansion path; the checker for the
 * CUDA product, never part of it. */
#ifndef KGMT_ORACLE_H
#define KGMT_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_MATH_HOST = 0, ORC_MATH_FMA = 1 };

enum {
    ORC_STATUS_RUNNING = 0,
    ORC_STATUS_SOLVED = 1,
    ORC_STATUS_TREE_FULL = 2,
    ORC_STATUS_ITER_LIMIT = 3,
    ORC_STATUS_FRONTIER_EMPTY = 4
};

/* array ids: the first 13 are the reference's CSV dumps in the order of
 * src/planners/KGMT.cu:299-311 */
enum {
    ORC_ARR_TREE_SAMPLES = 0, ORC_ARR_UNEXPLORED = 1, ORC_ARR_TREE_PARENT = 2, ORC_ARR_U_PARENT = 3,
    ORC_ARR_G = 4, ORC_ARR_R2AVAIL = 5, ORC_ARR_R1AVAIL = 6, ORC_ARR_R1VALID = 7, ORC_ARR_R2VALID = 8,
    ORC_ARR_R1INVALID = 9, ORC_ARR_R2INVALID = 10, ORC_ARR_R1SCORE = 11, ORC_ARR_R1 = 12,
    ORC_ARR_R2 = 13, ORC_ARR_COSTS = 14,
    ORC_ARR_U_VALID = 15, ORC_ARR_U_R1 = 16, ORC_ARR_U_R2 = 17, ORC_ARR_U_U3 = 18, ORC_ARR_U_MARGIN = 19
};

typedef struct orc_planner orc_planner;

void  orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
float orc_uniform(uint32_t x);
void  orc_slot_uniforms(uint32_t key0, uint32_t slot, float u[4]);
void  orc_controls(const float u[3], int math_mode, float* a, float* steering, float* duration);
/* control ranges {accel_min, accel_max, steer_min, steer_max, duration_min, duration_max}; NULL = the reference's literals */
void  orc_set_car_ranges(const double* r6);
void  orc_controls_general(const float u[3], const double r[6], int math_mode, float* a, float* steering, float* duration);
int   orc_motion_valid(const float bbMin[2], const float bbMax[2], const float* obstacles, int K);
int   orc_propagate_ctrl(const float x0[4], float a, float steering, float duration,
                         int numDisc, float agentLength, const float* obstacles, int K,
                         float width, float height, int math_mode,
                         float x1[7], float* margin, int* steps);
int   orc_propagate_slot(const float x0[4], uint32_t key0, uint32_t slot, int numDisc,
                         float agentLength, const float* obstacles, int K,
                         float width, float height, int math_mode,
                         float x1[7], float* u3_out, float* margin, int* steps);
int   orc_getR1(float x, float y, float R1Size, int N);
int   orc_getR2(float x, float y, int r1, float R1Size, int N, float R2Size, int n);
int   orc_getR2_fma(float x, float y, int r1, float R1Size, int N, float R2Size, int n);
int   orc_getR2_mode(float x, float y, int r1, float R1Size, int N, float R2Size, int n, int math_mode);
void  orc_scores(const int* R1Avail, const int* R2Avail, const int* R1Valid, const int* R1Invalid,
                 const int* R1, int N, int n, float epsilon, float* R1Score, float* R1Threshold);
int   orc_in_goal(const float* x, const float* goal, float r);
void  orc_expansion_shape(int activeSize, int treeSize, int maxTreeSize, int* mode, int* children, int* M);
void  orc_update_maps(int M, const int* r1v, const int* r2v, const uint8_t* valid, const float* u3,
                      const float* R1Score, const int* R2AvailSnap,
                      int* R1, int* R2, int* R1Valid, int* R2Valid, int* R1Invalid, int* R2Invalid,
                      int* R1Avail, int* R2Avail, uint8_t* accept);
int   orc_insert(int M, const uint8_t* accept, const float* cand, const int* candParent,
                 int treeSize, float* treeSamples, int* treeParentIdx, float* costs, uint8_t* G,
                 const float* goal, float r, float* costToGoal, int* goalIdx);

orc_planner* orc_create(float width, float height, int N, int n, int numIterations, int maxTreeSize,
                        int numDisc, float agentLength, float goalThreshold,
                        uint32_t seed, int math_mode);
void  orc_destroy(orc_planner* p);
void  orc_reset(orc_planner* p);
void  orc_set_obstacles(orc_planner* p, const float* aabb, int K);
void  orc_begin(orc_planner* p, const float initial[7], const float goal[7]);
int   orc_iterate(orc_planner* p);
int   orc_plan(orc_planner* p, const float initial[7], const float goal[7]);

int   orc_tree_size(const orc_planner* p);
int   orc_iterations(const orc_planner* p);
float orc_cost_to_goal(const orc_planner* p);
int   orc_goal_index(const orc_planner* p);
int   orc_status(const orc_planner* p);
long long orc_expansions(const orc_planner* p);
int   orc_frontier_start(const orc_planner* p);
int   orc_frontier_count(const orc_planner* p);
int   orc_last_M(const orc_planner* p);
int   orc_last_accepted(const orc_planner* p);
int   orc_last_mode(const orc_planner* p);
int   orc_last_children(const orc_planner* p);
float orc_R1Threshold(const orc_planner* p);
void* orc_array(orc_planner* p, int id);

void  orc_regions_batch(const float* xy, int stride, long M, float R1Size, int N, float R2Size, int n,
                        int math_mode, int* r1, int* r2);
void  orc_propagate_batch(const float* parents, const int* parentOf, long M,
                          float* x1, uint8_t* valid, float* u3, float* margin,
                          int numDisc, float L, uint32_t key0, uint32_t slot0,
                          const float* obstacles, int K, float W, float H, int math_mode);

#ifdef __cplusplus
}
#endif
#endif
