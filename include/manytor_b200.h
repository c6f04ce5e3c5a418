/*
 * manytor_b200 -- C ABI of the B200-native batched ManyTor step loop.
 *
 * This header is the drop-in boundary for the hot path of victorkich/ManyTor
 * (Environment.step / reset and the Multienv loop).  The reference has no FFI
 * of its own: its boundary is the Python surface of manytor.py, so every entry
 * point below names the reference method (file:line under /root/reference) it
 * replaces, vectorised over N lock-step environments.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types.
 *   - "_dev" pointers are device pointers on the handle's device, 16-byte
 *     aligned; "_host" pointers are host memory (pinned where the call is
 *     asynchronous).  `stream` is a cudaStream_t passed as void* (NULL = the
 *     legacy default stream).  Calls are asynchronous on that stream unless
 *     documented as synchronous.
 *   - every call returns MT_OK (0) or a negative mt_status; mt_last_error()
 *     returns a per-thread message for the last failure.
 *   - one handle = one shard of environments on one device; handles are not
 *     thread-safe.
 *   - angles are DEGREES, actions are absolute joint targets (manytor.py:182-184).
 *
 * Array layouts (row-major, fp32 unless noted)
 *   actions [N][J]        obs [N][3*X]  (per objective: distance, r, theta;
 *   reward  [N]                          zeros for a dead objective, manytor.py:141-153)
 *   done    [N] uint8     bit0 = terminated (no objective alive, manytor.py:170-171,
 *                                  or ground contact when terminate_on_ground)
 *                         bit1 = truncated (episode reached cfg.horizon steps)
 *   points  [N][X][3]     goals [N][J]     alive [N] uint32 bitmask (bit p = objective p)
 *   joints  [N][J][3]     the reference's joints_coordinates: rows = base, frame 2 .. frame J
 */
#ifndef MANYTOR_B200_H
#define MANYTOR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MT_ABI_VERSION 2
#define MT_MAX_JOINTS 8
#define MT_MAX_OBJ 32

typedef enum mt_status {
    MT_OK = 0,
    MT_ERR_INVALID = -1,   /* bad argument / configuration                     */
    MT_ERR_CUDA = -2,      /* a CUDA runtime call failed (message has details) */
    MT_ERR_NO_DEVICE = -3, /* no usable sm_100 device: there is NO CPU fallback */
    MT_ERR_STATE = -4      /* call not valid in the handle's current state     */
} mt_status;

typedef struct mt_env mt_env; /* opaque handle */

/* Everything the reference hard-codes ("Config / flags"). */
typedef struct mt_config {
    uint32_t struct_size;          /* = sizeof(mt_config), checked by mt_create          */
    int32_t device;                /* CUDA device ordinal                                 */
    int64_t n_envs;                /* N: environments in this shard                       */
    int64_t env_id_base;           /* global id of env 0 (keys the RNG: shard-invariant)  */
    int32_t n_joints;              /* J, 1..MT_MAX_JOINTS (reference: 4)                  */
    int32_t n_obj;                 /* X, 1..MT_MAX_OBJ (Environment(obj_number), :130)    */
    float dh[MT_MAX_JOINTS][4];    /* rows (a, alpha, d, theta_offset), radians (:42-48)  */
    int32_t obs_frame;             /* frame anchoring observations (reference: 3, :143)   */
    int32_t ground_frame_a;        /* frames whose z<0 is ground contact (3 and 4, :191)  */
    int32_t ground_frame_b;
    int32_t catch_frame;           /* frame that catches objectives (4, :162)             */
    float radius;                  /* objective half-ball radius (51.3, :231,236)         */
    float catch_tol;               /* per-axis inclusive tolerance (8.0, :162)            */
    int32_t substeps;              /* interpolated poses per step (25, :178)              */
    int32_t horizon;               /* max steps per episode, 0 = none (max_steps in the
                                      reference's driver scripts, test_single.py:6), at most
                                      65535.  The per-env episode-length counter saturates: at
                                      65535 when n_obj <= 16 or horizon = 0, else at
                                      2^(32 - n_obj) - 1 when the horizon fits below that (the
                                      counter then shares a word with the alive mask), else 65535 */
    int32_t terminate_on_ground;   /* 0 = reference code (reward -1 only), 1 = README     */
    int32_t auto_reset;            /* 1: envs that end are reset inside the step kernel   */
    int32_t obs_after_reset;       /* with auto_reset: 0 = emit obs2 of the ending step,
                                      1 = emit the first observation of the new episode   */
    int32_t fk_mode;               /* 0 auto (closed form / preset / NVRTC-specialised / run-time
                                      table, in that order), 1 run-time DH table, 2 closed form
                                      required, 3 NVRTC specialisation required              */
    int32_t action_low;            /* action_sample range [low, high) (-180, 180, :216)   */
    int32_t action_high;
    uint64_t seed;                 /* Philox key for on-device actions / objectives       */
} mt_config;

/* Episode statistics; summed over the shard, reducible across shards by plain
 * addition (the one collective of a multi-GPU rollout). */
typedef struct mt_stats {
    int64_t env_steps;       /* env-steps executed since create / stats reset            */
    int64_t episodes;        /* episodes finished (terminated or truncated)              */
    int64_t terminated;      /* ... of which all objectives were collected               */
    int64_t reward_sum;      /* sum of total_reward over finished episodes               */
    int64_t length_sum;      /* sum of episode lengths over finished episodes            */
    int64_t catches;         /* objectives collected in finished episodes                */
    int64_t ground_steps;    /* env-steps with ground contact (every executed step)      */
    int64_t live_reward_sum; /* sum of total_reward over episodes still in progress      */
} mt_stats;
#define MT_STATS_WORDS 8

int mt_abi_version(void);
const char *mt_last_error(void);

/* Fill cfg with the reference arm and constants (manytor.py:42-48,162,178,216,231). */
int mt_config_init(mt_config *cfg);

/* Environment(obj_number) / Multienv(env_shape, obj_number) -- manytor.py:77-82,130-139.
 * Allocates the structure-of-arrays state in HBM.  Envs start un-reset, as in the
 * reference; call mt_reset before mt_step. */
int mt_create(const mt_config *cfg, mt_env **out);
int mt_destroy(mt_env *env);
int mt_get_config(const mt_env *env, mt_config *out);
/* Re-seed the on-device Philox streams (actions, objective refresh): new key AND the counters that
 * index the streams (per-env episode count, step index) back to zero, so reset(seed=s) reproduces the
 * same objectives and actions every time it is called with the same s (Gymnasium's reset(seed=...)
 * contract).  Synchronous (waits for the device). */
int mt_set_seed(mt_env *env, uint64_t seed);

/* The step index that keys the action stream of mt_sample_actions / mt_rollout_random.  It lives in
 * DEVICE memory and is advanced by the step kernel itself (the last block of every mt_step,
 * mt_step_host and of each step of mt_rollout_random), as are mt_stats.env_steps and the other
 * statistics.  Consequence for CUDA graphs: a captured mt_step / mt_rollout_random can be replayed
 * any number of times -- every replay counts its env-steps and draws fresh actions, exactly like the
 * same calls issued eagerly.  Get/set are synchronous 8-byte copies (checkpoint / resume). */
int mt_get_step_index(mt_env *env, uint64_t *out);
int mt_set_step_index(mt_env *env, uint64_t value);

/* Environment.reset / Multienv.reset -- manytor.py:219-253, 106-109.
 * mask_dev: NULL = all envs, else [N] uint8 selecting the envs to reset.  The FIRST reset of a handle
 * must cover every env (NULL or an all-ones mask; a partial first mask returns MT_ERR_STATE after one
 * small synchronous check): the reference raises IndexError when a never-reset env is stepped.
 * Fresh objectives come from the objective stream if one is set, else from the
 * on-device half-ball sampler (uniform in {|p|<=radius, z>=0}). */
int mt_reset(mt_env *env, const uint8_t *mask_dev, void *stream);

/* Environment.get_observations -- manytor.py:141-153: current obs of every env. */
int mt_observe(mt_env *env, float *obs_dev, void *stream);

/* Environment.step / Multienv.step -- manytor.py:255-260, 115-122: ONE fused kernel
 * (FK over the 25 sub-poses, ground flag, obs2, catch, reward, done, auto-reset).
 * joints_dev may be NULL. */
int mt_step(mt_env *env, const float *actions_dev, float *obs_dev, float *reward_dev,
            uint8_t *done_dev, float *joints_dev, void *stream);

/* Environment.action_sample / Multienv.action_sample -- manytor.py:215-217, 111-113:
 * integer degrees uniform on [action_low, action_high), Philox keyed by
 * (seed, global env id, step index). */
int mt_sample_actions(mt_env *env, float *actions_dev, void *stream);

/* n_steps x step(action_sample()) with the actions drawn inside the kernel (same stream
 * of actions as mt_sample_actions).  obs_dev/reward_dev/done_dev are overwritten every
 * step; obs_dev may be NULL to skip the observation write; reward_dev / done_dev may be
 * NULL, the results then go to buffers the handle owns (a host without CUDA headers can
 * drive a rollout and read only the statistics).
 * For n_steps >= 2 this is ONE launch per 4096 steps of the multi-step rollout kernel:
 * every tile of 32 envs keeps its pose / alive word / total reward in registers and its
 * objectives in shared memory for all the steps, and only what a step produces
 * (observations, reward, done) leaves the SM each step.  Results are bit-identical to
 * n_steps launches of the step kernel, which run-time DH tables without NVRTC still
 * take (MT_ROLLOUT_PERSISTENT=0 forces them for everyone). */
int mt_rollout_random(mt_env *env, int32_t n_steps, float *obs_dev, float *reward_dev,
                      uint8_t *done_dev, void *stream);

/* Host-buffer step: actions_host/obs_host/reward_host/done_host must be pinned
 * (mt_host_alloc).  Copies actions H2D, steps, copies results D2H, chunked over
 * internal streams so copies overlap the kernels; synchronous on return.  Ordered after
 * everything submitted earlier to the legacy default stream / blocking streams (work on
 * cudaStreamNonBlocking streams must be synchronised by the caller). */
int mt_step_host(mt_env *env, const float *actions_host, float *obs_host, float *reward_host,
                 uint8_t *done_host);
/* mt_step_host has two implementations with identical results -- chunks staged through device buffers,
 * or one launch reading and writing the pinned host buffers across PCIe itself -- and by default times
 * both on a handle's first six calls and keeps the faster (MT_HOST_ZEROCOPY=0/1 pins it).  Returns 0
 * staged, 1 zero-copy, 2 still deciding. */
int mt_host_step_mode(const mt_env *env);
int mt_host_alloc(void **out, uint64_t bytes);
int mt_host_free(void *p);

/* State exchange (parity upload of reference objectives, checkpoint, render copy-back).
 * Any pointer may be NULL to skip that array; mask_dev as in mt_reset. */
int mt_set_points(mt_env *env, const float *points_dev, const uint8_t *mask_dev, void *stream);
int mt_get_points(mt_env *env, float *points_dev, int32_t zero_dead, void *stream);
int mt_set_state(mt_env *env, const float *goals_dev, const uint32_t *alive_dev,
                 const float *total_reward_dev, const int32_t *ep_len_dev,
                 const uint8_t *mask_dev, void *stream);
int mt_get_state(mt_env *env, float *goals_dev, uint32_t *alive_dev, float *total_reward_dev,
                 int32_t *ep_len_dev, void *stream);

/* Objective refresh stream for parity runs: points_dev [n_sets][N][X][3], borrowed.
 * The e-th reset of env n takes set (e mod n_sets).  NULL/0 restores the sampler. */
int mt_set_objective_stream(mt_env *env, const float *points_dev, int32_t n_sets);

/* One-env copy-back for the render shim and per-env attribute reads (manytor.py:131-139,
 * 196-201).  Synchronous, on the legacy default stream: one small kernel + one copy through
 * scratch buffers the handle owns (no allocation, no device-wide synchronisation). */
int mt_fetch_env(mt_env *env, int64_t index, float *goals_host, float *joints_host,
                 float *points_host, uint32_t *alive_host, float *total_reward_host);

/* Episode statistics.  mt_stats_device writes MT_STATS_WORDS int64 to a device
 * buffer (for an NCCL all-reduce); mt_stats_host is the synchronous host read. */
int mt_stats_device(mt_env *env, int64_t *stats_dev, void *stream);
int mt_stats_host(mt_env *env, mt_stats *out);
int mt_stats_clear(mt_env *env, void *stream);

/* The ONE collective of a multi-GPU rollout: every shard's statistics summed
 * with ncclAllReduce (NVLink / NVSwitch).  NCCL is dlopen'ed at first use (libnccl.so.2).
 *   mt_stats_allreduce       one process driving n handles on n distinct devices; communicators
 *                            from ncclCommInitAll, cached per device list.  Synchronous; the sum
 *                            over all handles is written to *out_host.
 *   mt_stats_allreduce_comm  one process per GPU: `nccl_comm` is the caller's ncclComm_t
 *                            (ncclCommInitRank); stats_dev [MT_STATS_WORDS] int64 on the handle's
 *                            device receives the global sum, asynchronously on `stream`. */
int mt_stats_allreduce(mt_env *const *envs, int32_t n, mt_stats *out_host);
int mt_stats_allreduce_comm(mt_env *env, void *nccl_comm, int64_t *stats_dev, void *stream);
/* The same sum by the library's own kernel over NVLink peer memory (one process per GPU): every rank
 * owns a zero-initialised buffer of mt_stats_peer_buffer_bytes(world) bytes that all peers can write
 * (torch symmetric memory, cuMem IPC, ...); peers_dev is a DEVICE array of the `world` buffer
 * addresses as seen from this rank.  One small kernel per rank stores this rank's 64 bytes into every
 * peer's buffer, raises a flag there and waits for the peers' flags in its own: ~4 us instead of
 * ~30 us for the NCCL call.  Asynchronous on `stream`; the epoch that keys the flags lives in device
 * memory (graph-replay safe).  Every rank of the group must make the call; if a peer never does, the
 * kernel gives up after ~10 s and writes -1 into every word instead of hanging the GPU. */
int64_t mt_stats_peer_buffer_bytes(int32_t world);
int mt_stats_allreduce_peers(mt_env *env, int64_t *const *peers_dev, int32_t rank, int32_t world,
                             int64_t *stats_dev, void *stream);

/* Module-level helpers of the reference, batched over M rows:
 * fk(mode, goals) manytor.py:35-53 -> out [M][16]; dh(a, alfa, d, theta) :25-32 ->
 * out [M][16] from params [M][4]; r_theta(v1, v2) :17-22 -> out [M][2]. */
int mt_fk(const mt_config *cfg, int32_t mode, const float *goals_dev, float *out_dev, int64_t m,
          void *stream);
int mt_dh(const float *params_dev, float *out_dev, int64_t m, void *stream);
/* joints_coordinates (manytor.py:188-189) of M poses: goals [M][J] -> out [M][J][3]
 * (rows base, frame 2 .. frame J) with the handle's arm; used by the render shim. */
int mt_joints(mt_env *env, const float *goals_dev, float *out_dev, int64_t m, void *stream);
int mt_r_theta(const float *v1_dev, const float *v2_dev, float *out_dev, int64_t m, void *stream);

/* Introspection for bench.py: kernels launched by this handle since create, and
 * the algorithmic bytes per env-step of mt_step. */
int64_t mt_launch_count(const mt_env *env);
int64_t mt_bytes_per_env_step(const mt_env *env, int32_t actions_from_hbm, int32_t obs_written);
/* Device time (ms) of the step kernels launched by the last mt_rollout_random /
 * mt_step call when timing was enabled with mt_set_timing(env, 1). */
int mt_set_timing(mt_env *env, int32_t enabled);
int mt_last_kernel_ms(mt_env *env, float *ms_out);

#ifdef __cplusplus
}
#endif
#endif /* MANYTOR_B200_H */
