/* TEST INFRASTRUCTURE ONLY — never linked into the product library.
 *
 * Glue around the reference's OWN, UNMODIFIED propagate+collision sources
 * (/root/reference/src/statePropagator/statePropagator.cu:5-76,
 *  /root/reference/src/collisionCheck/collisionCheck.cu:6-28) built for the
 * host by oracle/Makefile.  The reference sources are
 * compiled from where they lie; only this glue and the shims are ours.
 *
 * Provides:
 *   - curand_uniform(curandState*) for the shim state: host Philox4x32-10 with
 *     cuRAND's stateful draw order (curand_kernel.h:888-915) and the uniform
 *     conversion of curand_uniform.h:69-72;
 *   - extern "C" entry points that tests / bench.py's cpu_baseline call through
 *     ctypes: one candidate, a batch of candidates (threaded), and a timing loop.
 */
#include "statePropagator/statePropagator.cuh"   /* reference prototype, via shim curand_kernel.h */
#include <cstring>
#include <thread>
#include <vector>
#include <chrono>

static inline void philox_round(uint32_t c[4], const uint32_t k[2]) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    const uint32_t n0 = hi1 ^ c[1] ^ k[0];
    const uint32_t n2 = hi0 ^ c[3] ^ k[1];
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k[2] = {key[0], key[1]};
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k);
        if (r < 9) { k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u; }
    }
    memcpy(out, c, sizeof(uint32_t) * 4);
}

/* curand_init(seed, subsequence, 0, &s) for Philox: key=(seed_lo,seed_hi),
 * ctr=(0,0,subseq_lo,subseq_hi). */
static void host_curand_init(uint64_t seed, uint64_t subseq, curandState* s) {
    s->key[0] = (uint32_t)seed; s->key[1] = (uint32_t)(seed >> 32);
    s->ctr[0] = 0; s->ctr[1] = 0; s->ctr[2] = (uint32_t)subseq; s->ctr[3] = (uint32_t)(subseq >> 32);
    philox4x32_10(s->ctr, s->key, s->out);
    s->pos = 0;
}

float curand_uniform(curandState* s) {
    if (s->pos == 4) {
        if (++s->ctr[0] == 0) if (++s->ctr[1] == 0) if (++s->ctr[2] == 0) ++s->ctr[3];
        philox4x32_10(s->ctr, s->key, s->out);
        s->pos = 0;
    }
    const uint32_t x = s->out[s->pos++];
    return x * 2.3283064e-10f + (2.3283064e-10f / 2.0f);
}

extern "C" {

void ref_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    philox4x32_10(ctr, key, out);
}

/* One candidate through the reference's propagateAndCheck.  Stream = slot
 * `slot` of key (seed,0); u3 is the 4th draw (the accept uniform of
 * KGMT.cu:395), drawn unconditionally here so callers can inspect it. */
int ref_host_propagate(const float* x0, float* x1, int numDisc, float agentLength,
                       uint32_t seed, uint32_t slot, const float* obstacles, int K,
                       float width, float height, float* u3) {
    curandState st;
    host_curand_init(seed, slot, &st);
    float x0c[7]; memcpy(x0c, x0, sizeof(float) * 7);
    bool ok = propagateAndCheck(x0c, x1, numDisc, agentLength, &st,
                                const_cast<float*>(obstacles), K, width, height);
    if (u3) *u3 = curand_uniform(&st);
    return ok ? 1 : 0;
}

static void batch_range(const float* parents, const int* parentOf, long lo, long hi,
                        float* x1, uint8_t* valid, float* u3, int numDisc, float L,
                        uint32_t seed, uint32_t slot0, const float* obstacles, int K, float W, float H) {
    for (long s = lo; s < hi; ++s) {
        float u;
        float tmp[7];
        float* dst = x1 ? x1 + 7 * s : tmp;
        int ok = ref_host_propagate(parents + 7 * (long)parentOf[s], dst, numDisc, L, seed,
                                    slot0 + (uint32_t)s, obstacles, K, W, H, &u);
        if (valid) valid[s] = (uint8_t)ok;
        if (u3) u3[s] = u;
    }
}

/* M candidates, candidate s expands parents[parentOf[s]] with stream slot0+s.
 * threads<=1 runs inline.  Returns wall seconds. */
double ref_host_propagate_batch(const float* parents, const int* parentOf, long M,
                                float* x1, uint8_t* valid, float* u3, int numDisc, float L,
                                uint32_t seed, uint32_t slot0, const float* obstacles, int K,
                                float W, float H, int threads) {
    auto t0 = std::chrono::steady_clock::now();
    if (threads <= 1) {
        batch_range(parents, parentOf, 0, M, x1, valid, u3, numDisc, L, seed, slot0, obstacles, K, W, H);
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t) {
            long lo = M * t / threads, hi = M * (t + 1) / threads;
            pool.emplace_back(batch_range, parents, parentOf, lo, hi, x1, valid, u3, numDisc, L,
                              seed, slot0, obstacles, K, W, H);
        }
        for (auto& th : pool) th.join();
    }
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

int ref_host_hw_threads(void) { return (int)std::thread::hardware_concurrency(); }

} /* extern "C" */
