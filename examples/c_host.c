/* A host written in plain C against include/manytor_b200.h: no CUDA headers, no torch, no Python.
 *
 *   gcc -O2 -Iinclude examples/c_host.c -o c_host -Lmanytor_b200/lib -lmanytor_b200 \
 *       -Wl,-rpath,$PWD/manytor_b200/lib
 *   ./c_host [n_envs] [steps]
 *
 * It is the loop of the reference's test_multi.py (test_multi.py:11-34): build N envs, reset, then
 * `steps` x (random actions -> step), print the summed reward and the episode statistics.  Host buffers
 * come from mt_host_alloc (pinned), so mt_step_host can overlap its H2D / kernel / D2H chunks.
 */
#include <stdio.h>
#include <stdlib.h>

#include "manytor_b200.h"

#define CHECK(call)                                                                  \
    do {                                                                             \
        int rc_ = (call);                                                            \
        if (rc_ != MT_OK) {                                                          \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, mt_last_error());          \
            return 1;                                                                \
        }                                                                            \
    } while (0)

int main(int argc, char **argv) {
    const long long n = argc > 1 ? atoll(argv[1]) : 4096;
    const int steps = argc > 2 ? atoi(argv[2]) : 50;
    mt_config cfg;
    CHECK(mt_config_init(&cfg)); /* the reference arm and constants (manytor.py:42-48,162,178,231) */
    cfg.n_envs = n;
    cfg.n_obj = 7;       /* test_multi.py:9 */
    cfg.horizon = 50;    /* test_multi.py:8 max_steps */
    cfg.auto_reset = 1;
    cfg.seed = 2020;
    mt_env *env = NULL;
    CHECK(mt_create(&cfg, &env));

    float *actions, *obs, *reward;
    uint8_t *done;
    CHECK(mt_host_alloc((void **)&actions, (uint64_t)n * 4 * sizeof(float)));
    CHECK(mt_host_alloc((void **)&obs, (uint64_t)n * 3 * cfg.n_obj * sizeof(float)));
    CHECK(mt_host_alloc((void **)&reward, (uint64_t)n * sizeof(float)));
    CHECK(mt_host_alloc((void **)&done, (uint64_t)n));

    CHECK(mt_reset(env, NULL, NULL));
    double total = 0.0;
    long long ended = 0;
    srand(1);
    for (int t = 0; t < steps; ++t) {
        for (long long i = 0; i < n * 4; ++i) actions[i] = (float)(rand() % 360 - 180); /* manytor.py:216 */
        CHECK(mt_step_host(env, actions, obs, reward, done));
        for (long long i = 0; i < n; ++i) {
            total += reward[i];
            ended += done[i] != 0;
        }
    }
    mt_stats st;
    CHECK(mt_stats_host(env, &st));
    printf("envs %lld steps %d: reward sum %.0f, episodes ended %lld (stats: %lld episodes, %lld terminated, "
           "%lld env-steps), obs[0][0..2] = %.3f %.3f %.3f\n",
           n, steps, total, ended, (long long)st.episodes, (long long)st.terminated, (long long)st.env_steps,
           obs[0], obs[1], obs[2]);
    if (st.episodes != ended || st.env_steps != n * steps) {
        fprintf(stderr, "statistics disagree with the per-step outputs\n");
        return 2;
    }
    CHECK(mt_host_free(actions));
    CHECK(mt_host_free(obs));
    CHECK(mt_host_free(reward));
    CHECK(mt_host_free(done));
    CHECK(mt_destroy(env));
    return 0;
}
