/* demos/kgmt_multi_demo.cu — the multi-GPU modes of the KGMT path from a plain C++ host program: one PROCESS per GPU
 * (fork), the NCCL unique id handed over through a shared page, everything else through the C ABI (include/kgmt_c.h:
 * kgmt_comm_init, kgmt_plan_sharded, kgmt_plan_batch_sharded, kgmt_plan_portfolio, kgmt_expand_sharded).  No Python, no
 * MPI.  This is what the reference's demos/main.cu:30,62 would look like with one process per GPU.
 *
 *   kgmt_multi_demo [world] [obstacles.csv] [N n maxTree]
 *
 * world defaults to the number of visible GPUs.  Obstacles: the reference's CSV format (minx,miny,maxx,maxy per line,
 * configurations/obstacles/obstacles.csv), default = its five boxes.  Rank 0 prints one line per mode; every rank checks
 * that the sharded plan equals its own single-GPU plan (tree size, iterations, cost, goal index) and exits non-zero if
 * anything differs.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>
#include <vector>

#include <cuda_runtime.h>

#include "kgmt_c.h"

struct Shared { volatile int ready; unsigned char id[KGMT_COMM_ID_BYTES]; volatile int failures; };

#define CK(call)                                                                                                  \
    do {                                                                                                          \
        int rc_ = (call);                                                                                         \
        if (rc_ < 0) { fprintf(stderr, "[rank %d] %s failed (%d): %s\n", rank, #call, rc_, kgmt_last_error(ctx)); return 10; } \
    } while (0)

static std::vector<float> read_obstacles(const char* path) {
    std::vector<float> o;
    if (!path) {
        const float d[20] = {2, 2, 4, 4, 7, 2, 9, 5, 3, 18, 6, 20, 2, 10, 4, 12, 0, 6, 18, 8};   /* obstacles.csv:1-5 */
        o.assign(d, d + 20);
        return o;
    }
    FILE* f = fopen(path, "r");
    if (!f) return o;
    float a, b, c, d;
    while (fscanf(f, " %f , %f , %f , %f", &a, &b, &c, &d) == 4) { o.push_back(a); o.push_back(b); o.push_back(c); o.push_back(d); }
    fclose(f);
    return o;
}

static int run_rank(int rank, int world, Shared* sh, const std::vector<float>& obs, int N, int n, int maxTree,
                    const float* init, const float* goal) {
    kgmt_ctx* ctx = nullptr;
    kgmt_params p;
    kgmt_default_params(&p);
    p.N = N; p.n = n; p.max_tree_size = maxTree; p.device = rank; p.seed = 7;
    CK(kgmt_create(&p, &ctx));
    CK(kgmt_set_obstacles_host(ctx, obs.data(), (int)(obs.size() / 4)));
    /* the unique id: rank 0 makes it, the others wait for the shared page */
    if (rank == 0) {
        CK(kgmt_comm_unique_id(sh->id));
        __sync_synchronize();
        sh->ready = 1;
    } else {
        while (!sh->ready) usleep(100);
        __sync_synchronize();
    }
    CK(kgmt_comm_init(ctx, rank, world, sh->id));

    /* 1. whole plan with sharded iterations == the single-GPU plan */
    kgmt_result one, sharded;
    CK(kgmt_plan(ctx, init, goal, &one));
    CK(kgmt_comm_barrier(ctx));
    CK(kgmt_plan_sharded(ctx, init, goal, &sharded));
    const bool same = one.stop == sharded.stop && one.iterations == sharded.iterations && one.tree_size == sharded.tree_size &&
                      one.cost_to_goal == sharded.cost_to_goal && one.goal_index == sharded.goal_index && one.expansions == sharded.expansions;
    if (!same) {
        fprintf(stderr, "[rank %d] sharded plan differs: stop %d/%d iterations %d/%d tree %d/%d cost %g/%g\n", rank, one.stop,
                sharded.stop, one.iterations, sharded.iterations, one.tree_size, sharded.tree_size, one.cost_to_goal, sharded.cost_to_goal);
        __sync_fetch_and_add(&sh->failures, 1);
    }
    if (rank == 0)
        printf("plan_sharded world %d: stop %d iterations %d tree %d cost %.6f | single GPU %.3f ms, sharded %.3f ms | identical %d\n", world,
               sharded.stop, sharded.iterations, sharded.tree_size, sharded.cost_to_goal, one.device_ms, sharded.device_ms, (int)same);

    /* 2. config 4: a batch of independent queries sharded over the ranks */
    {
        const int Q = 256;
        std::vector<float> inits((size_t)Q * 7, 0.f), goals((size_t)Q * 7, 0.f);
        std::vector<uint32_t> seeds(Q);
        for (int q = 0; q < Q; ++q) {
            memcpy(&inits[(size_t)q * 7], init, 28); memcpy(&goals[(size_t)q * 7], goal, 28);
            seeds[q] = 100u + (uint32_t)q;
        }
        std::vector<kgmt_result> all(Q);
        float ms = 0.f;
        CK(kgmt_plan_batch_sharded(ctx, inits.data(), goals.data(), seeds.data(), Q, 0, all.data(), &ms));
        int solved = 0; long long exp = 0;
        for (const kgmt_result& r : all) { solved += r.stop == KGMT_SOLVED; exp += r.expansions; }
        if (rank == 0) printf("plan_batch_sharded world %d: %d queries, %d solved, %lld expansions, slowest rank %.3f ms\n", world, Q, solved, exp, ms);
        /* every rank holds the same table: query 0 must equal this rank's own plan of the same seed */
        kgmt_result mine;
        CK(kgmt_set_seed(ctx, seeds[0]));
        CK(kgmt_plan(ctx, init, goal, &mine));
        if (mine.tree_size != all[0].tree_size || mine.cost_to_goal != all[0].cost_to_goal) {
            fprintf(stderr, "[rank %d] batch result of query 0 differs from kgmt_plan\n", rank);
            __sync_fetch_and_add(&sh->failures, 1);
        }
        CK(kgmt_set_seed(ctx, 7));
    }

    /* 3. portfolio: same query, seed 1000 + rank, first solution stops the others; the winner's path everywhere */
    {
        kgmt_result win;
        int winner = -1, len = 0;
        std::vector<float> path(256 * 7);
        CK(kgmt_plan_portfolio(ctx, init, goal, 1000u, 1, &win, &winner, path.data(), 256, &len));
        if (rank == 0) printf("plan_portfolio world %d: winner rank %d cost %.6f path %d nodes\n", world, winner, win.cost_to_goal, len);
        if (winner >= 0 && len > 0) {
            const float* last = &path[(size_t)(std::min(len, 256) - 1) * 7];
            const float dx = last[0] - goal[0], dy = last[1] - goal[1];
            if (dx * dx + dy * dy >= p.goal_threshold * p.goal_threshold) {
                fprintf(stderr, "[rank %d] portfolio path does not end in the goal disc\n", rank);
                __sync_fetch_and_add(&sh->failures, 1);
            }
        }
        CK(kgmt_set_seed(ctx, 7));
    }

    /* 4. config 5: one sharded iteration, the three exchanges */
    for (int ex = 0; ex < 3; ++ex) {
        CK(kgmt_begin(ctx, init, goal));
        kgmt_iter_stats st;
        float ms3[3] = {0, 0, 0};
        for (int i = 0; i < 4; ++i) CK(kgmt_expand_sharded(ctx, ex, &st, ms3));
        if (rank == 0)
            printf("expand_sharded world %d exchange %d: iteration %d candidates %d accepted %d tree %d | compute %.3f ms exchange %.3f ms\n", world,
                   ex, st.iteration, st.candidates, st.accepted, st.tree_size, ms3[0], ms3[1]);
        CK(kgmt_comm_barrier(ctx));
    }
    CK(kgmt_comm_destroy(ctx));
    kgmt_destroy(ctx);
    return 0;
}

int main(int argc, char** argv) {
    int world = argc > 1 ? atoi(argv[1]) : 0;
    const char* obsPath = argc > 2 && strcmp(argv[2], "-") ? argv[2] : nullptr;
    const int N = argc > 5 ? atoi(argv[3]) : 16, n = argc > 5 ? atoi(argv[4]) : 8, maxTree = argc > 5 ? atoi(argv[5]) : 30000;
    const std::vector<float> obs = read_obstacles(obsPath);
    if (obsPath && obs.empty()) { fprintf(stderr, "cannot read %s\n", obsPath); return 2; }
    const float initC1[7] = {5, 5, 0, 0, 0, 0, 0}, goalC1[7] = {2, 18, 0, 0, 0, 0, 0};       /* main.cu:33-46 */
    const float initC2[7] = {1, 1, 0, 0, 0, 0, 0}, goalC2[7] = {19, 19, 0, 0, 0, 0, 0};
    const float* init = obsPath ? initC2 : initC1;
    const float* goal = obsPath ? goalC2 : goalC1;
    Shared* sh = (Shared*)mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (sh == MAP_FAILED) return 3;
    memset((void*)sh, 0, sizeof(Shared));
    if (world <= 0) {
        /* count the GPUs in a child: the parent must not initialise CUDA before it forks */
        int fd[2];
        if (pipe(fd)) return 3;
        if (fork() == 0) { int c = 0; cudaGetDeviceCount(&c); if (write(fd[1], &c, 4) != 4) _exit(1); _exit(0); }
        if (read(fd[0], &world, 4) != 4) world = 1;
        wait(nullptr);
        if (world < 1) { fprintf(stderr, "no CUDA device\n"); return 4; }
    }
    std::vector<pid_t> kids;
    for (int r = 0; r < world; ++r) {
        pid_t pid = fork();
        if (pid == 0) {
            const int rc = run_rank(r, world, sh, obs, N, n, maxTree, init, goal);
            fflush(nullptr);                      /* _exit does not flush stdio */
            _exit(rc);
        }
        kids.push_back(pid);
    }
    int bad = 0;
    for (pid_t k : kids) { int st = 0; waitpid(k, &st, 0); if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) ++bad; }
    if (bad || sh->failures) { fprintf(stderr, "FAILED: %d rank(s) exited with an error, %d mismatches\n", bad, sh->failures); return 1; }
    printf("multi-GPU demo ok (world %d)\n", world);
    return 0;
}
