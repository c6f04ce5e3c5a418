/* BASELINE config 4 from a plain-C host: one process, one handle per GPU, envs sharded contiguously,
 * NO per-step exchange, ONE collective at the end of the rollout (mt_stats_allreduce: ncclAllReduce of
 * the 64-byte statistics over NVLink, communicators from ncclCommInitAll).  No CUDA headers, no torch.
 *
 *   gcc -O2 -Iinclude examples/c_host_multi.c -o c_host_multi -Lmanytor_b200/lib -lmanytor_b200 \
 *       -Wl,-rpath,$PWD/manytor_b200/lib
 *   ./c_host_multi [n_gpus] [envs_per_gpu] [steps]
 *
 * The reference's Multienv loops over its envs in one Python process (manytor.py:115-122); here the same
 * loop is n_gpus launches per step, each advancing its shard, issued back to back (the calls are
 * asynchronous), so the GPUs run concurrently.
 */
#include <stdio.h>
#include <stdlib.h>

#include "manytor_b200.h"

#define MAX_GPUS 16
#define CHECK(call)                                                                  \
    do {                                                                             \
        int rc_ = (call);                                                            \
        if (rc_ != MT_OK) {                                                          \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, mt_last_error());          \
            return 1;                                                                \
        }                                                                            \
    } while (0)

int main(int argc, char **argv) {
    const int gpus = argc > 1 ? atoi(argv[1]) : 2;
    const long long per = argc > 2 ? atoll(argv[2]) : 1 << 16;
    const int steps = argc > 3 ? atoi(argv[3]) : 100;
    if (gpus < 1 || gpus > MAX_GPUS) return 1;
    mt_env *env[MAX_GPUS] = {0};
    for (int g = 0; g < gpus; ++g) {
        mt_config cfg;
        CHECK(mt_config_init(&cfg));
        cfg.device = g;
        cfg.n_envs = per;
        cfg.env_id_base = (long long)g * per; /* global env ids key the RNG: results do not depend on the shard count */
        cfg.horizon = 50;
        cfg.auto_reset = 1;
        cfg.seed = 2020;
        CHECK(mt_create(&cfg, &env[g]));
        CHECK(mt_reset(env[g], NULL, NULL));
    }
    /* in-kernel random actions; nothing is read back per step, so reward/done go to buffers the handle owns
     * (NULL) and observations are not written at all (NULL) */
    for (int t = 0; t < steps; ++t)
        for (int g = 0; g < gpus; ++g) CHECK(mt_rollout_random(env[g], 1, NULL, NULL, NULL, NULL));
    mt_stats total;
    CHECK(mt_stats_allreduce(env, gpus, &total));
    long long sum_steps = 0;
    for (int g = 0; g < gpus; ++g) {
        mt_stats s;
        CHECK(mt_stats_host(env[g], &s));
        sum_steps += s.env_steps;
    }
    printf("gpus %d x envs %lld x steps %d: %lld env-steps, %lld episodes (%lld terminated), %lld catches, ground rate %.3f\n",
           gpus, per, steps, (long long)total.env_steps, (long long)total.episodes, (long long)total.terminated,
           (long long)total.catches, (double)total.ground_steps / (double)total.env_steps);
    if (total.env_steps != (long long)gpus * per * steps || sum_steps != total.env_steps) {
        fprintf(stderr, "the all-reduced statistics disagree with the per-shard ones\n");
        return 2;
    }
    for (int g = 0; g < gpus; ++g) CHECK(mt_destroy(env[g]));
    return 0;
}
