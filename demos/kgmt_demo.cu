/* demos/kgmt_demo.cu — data-driven counterpart of the reference's demos/main.cu.
 *
 * The reference hard-codes every parameter (main.cu:19-46) and reads only the obstacle CSV (main.cu:51); its
 * configurations/{init,goal,numR1,R2}/*.csv are never opened and systems/car.yaml is empty.  This demo
 * reads all of them when present, falls back to the main.cu literals otherwise, plans through the
 * reference-compatible KGMT class, prints the solution path (§8f rank 2) and leaves the 13 CSV files in the cwd.
 * The car model — wheelbase, integration steps and the control ranges the reference hard-codes in
 * statePropagator.cu:17-19 — comes from <config_dir>/../systems/car.yaml (or the 5th argument) through
 * kgmt_params_from_yaml; the reference ships that file EMPTY, which leaves its literals in force.
 *
 *   kgmt_demo [config_dir=../configurations] [seed] [maxTreeSize] [numIterations] [car.yaml]
 */
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

#include "planners/KGMT.cuh"

static bool readRow(const std::string& path, std::vector<float>& out) {
    std::ifstream in(path);
    if (!in) return false;
    std::string tok;
    out.clear();
    while (std::getline(in, tok, ',')) {
        try { out.push_back(std::stof(tok)); } catch (...) {}
    }
    return !out.empty();
}

int main(int argc, char** argv) {
    const std::string dir = argc > 1 ? argv[1] : "../configurations";
    float width = 20.0f, height = 20.0f, agentLength = 1.0f, goalThreshold = 0.5f;      /* main.cu:20-21,27-28 */
    int N = 16, n = 8, numIterations = 100, maxTreeSize = 30000, numDisc = 10;           /* main.cu:22-26 */
    float initial[7] = {5, 5, 0, 0, 0, 0, 0}, goal[7] = {2, 18, 0, 0, 0, 0, 0};          /* main.cu:33-46 */
    std::vector<float> row;
    if (readRow(dir + "/init/init.csv", row) && row.size() >= 2)
        for (size_t i = 0; i < 7 && i < row.size(); ++i) initial[i] = row[i];
    if (readRow(dir + "/goal/goal.csv", row) && row.size() >= 2)
        for (size_t i = 0; i < 7 && i < row.size(); ++i) goal[i] = row[i];
    if (readRow(dir + "/numR1/numR1.csv", row)) N = (int)row[0];
    if (readRow(dir + "/R2/numR2.csv", row)) n = (int)row[0];
    /* the shipped numR2.csv says 16 while main.cu runs n = 8; the file wins here, the literals when it is absent */
    const unsigned seed = argc > 2 ? (unsigned)std::strtoul(argv[2], nullptr, 10) : 1u;
    if (argc > 3) maxTreeSize = std::atoi(argv[3]);
    if (argc > 4) numIterations = std::atoi(argv[4]);

    int numObstacles = 0;
    std::vector<float> obstacles;
    {
        std::ifstream probe(dir + "/obstacles/obstacles.csv");
        if (probe) obstacles = readObstaclesFromCSV(dir + "/obstacles/obstacles.csv", numObstacles, 2);
    }
    std::printf("init (%g, %g) goal (%g, %g) N %d n %d obstacles %d seed %u\n", initial[0], initial[1], goal[0], goal[1], N,
                n, numObstacles, seed);

    kgmt_params params;
    kgmt_default_params(&params);
    params.width = width; params.height = height; params.N = N; params.n = n; params.num_iterations = numIterations;
    params.max_tree_size = maxTreeSize; params.num_disc = numDisc; params.agent_length = agentLength;
    params.goal_threshold = goalThreshold; params.seed = seed; params.record_candidates = 1;
    {
        const std::string yaml = argc > 5 ? argv[5] : dir + "/../systems/car.yaml";
        std::ifstream probe(yaml);
        if (probe) {
            int badLine = 0;
            if (kgmt_params_from_yaml(yaml.c_str(), &params, &badLine) != KGMT_OK) {
                std::printf("car model %s: cannot parse line %d\n", yaml.c_str(), badLine);
                return 3;
            }
            std::printf("car model %s: L %g numDisc %d a [%g, %g] steer [%g, %g] duration [%g, %g]\n", yaml.c_str(),
                        params.agent_length, params.num_disc, params.accel_min, params.accel_max, params.steer_min,
                        params.steer_max, params.duration_min, params.duration_max);
        }
    }
    KGMT kgmt(params);
    if (!kgmt.context()) return 2;
    float* d_obstacles = nullptr;
    CUDA_ERROR_CHECK(cudaMalloc(&d_obstacles, sizeof(float) * 4 * (numObstacles > 0 ? numObstacles : 1)));
    if (numObstacles)
        CUDA_ERROR_CHECK(cudaMemcpy(d_obstacles, obstacles.data(), sizeof(float) * 4 * numObstacles, cudaMemcpyHostToDevice));
    kgmt.plan(initial, goal, d_obstacles, numObstacles);
    std::printf("stop %d iterations %d expansions %lld device_ms %.3f cost %g\n", kgmt.stop_, kgmt.iterations_,
                kgmt.expansions_, kgmt.deviceMs_, kgmt.costToGoal_);
    if (kgmt.stop_ == KGMT_SOLVED) {
        std::vector<float> path(7 * 4096);
        const int len = kgmt_extract_path(kgmt.context(), -1, path.data(), 4096);
        std::printf("solution: %d nodes\n", len);
        for (int i = 0; i < len && i < 4096; ++i)
            std::printf("  %3d  x %.4f y %.4f theta %.4f v %.4f | a %.4f steer %.4f dur %.4f\n", i, path[7 * i], path[7 * i + 1],
                        path[7 * i + 2], path[7 * i + 3], path[7 * i + 4], path[7 * i + 5], path[7 * i + 6]);
    }
    cudaFree(d_obstacles);
    return 0;
}
