/*
 * bmi.h — C-ABI of libbmi_b200.so: the B200-native (sm_100a) implementation of the
 * rollout -> HER-relabel -> DDPG-update hot path of PiggyCh/RL_arm_under_sparse_reward.
 *
 * The reference has no FFI layer (it is pure Python); this header is the boundary a
 * maintainer binds with ctypes.  Each entry point cites the
 * reference interface (file:line under the reference tree) that it replaces.
 *
 * Conventions
 *   - every pointer named *_dev / documented "device" is a caller-owned CUDA device
 *     pointer; the library never frees caller memory;
 *   - every call only ENQUEUES work on the caller-supplied stream (a cudaStream_t cast
 *     to void*; NULL = legacy default stream) and never synchronises it, so the calls
 *     can be captured in a CUDA graph by the caller;
 *   - return value: 0 on success, negative on error; bmi_last_error() returns a
 *     thread-local, human-readable message for the last failing call;
 *   - opaque handles are not thread-safe;
 *   - storage dtype codes: BMI_F32 = 0, BMI_F64 = 1.  The reference stores float64
 *     (replay_buffer.py:23-27); BMI_F64 reproduces it bit-for-bit, BMI_F32 is the
 *     bandwidth-lean layout used by the vectorised agent (values produced by the fp32
 *     physics kernel are exactly representable in either).
 */
#ifndef BMI_B200_H_
#define BMI_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BMI_ABI_VERSION 1
#define BMI_OK 0
#define BMI_ERR_ARG (-1)
#define BMI_ERR_CUDA (-2)
#define BMI_ERR_CUBLAS (-3)
#define BMI_ERR_NCCL (-4)
#define BMI_ERR_STATE (-5)

#define BMI_F32 0
#define BMI_F64 1

typedef void* bmi_stream_t; /* cudaStream_t */

int bmi_abi_version(void);
const char* bmi_last_error(void);
/* number of kernel launches issued by this library since load (bench.py: gpu_launches) */
int64_t bmi_launch_count(void);
/* hint: bytes fetched from DRAM per L2 miss (32 / 64 / 128) for the random gathers of the path (her.py:24-36 fancy indexing) */
int bmi_set_l2_fetch_granularity(int32_t bytes);

/* ------------------------------------------------------------------------------------
 * Episode store (struct-of-arrays by key, episode-major inside each key) — the layout of
 * replay_buffer.buffers (replay_buffer.py:23-27):
 *   obs[n_episodes][T+1][obs_dim]  ag[n_episodes][T+1][goal_dim]
 *   g[n_episodes][T][goal_dim]     actions[n_episodes][T][act_dim]
 * obs_next / ag_next are the same arrays shifted by one time step (replay_buffer.py:51-52).
 * ---------------------------------------------------------------------------------- */
typedef struct bmi_episodes {
  void* obs;
  void* ag;
  void* g;
  void* actions;
  int64_t n_episodes;
  int32_t T;
  int32_t obs_dim;
  int32_t goal_dim;
  int32_t act_dim;
  int32_t dtype; /* BMI_F32 | BMI_F64 */
  int32_t _pad;
} bmi_episodes;

/* Output of a HER draw: row-major [B][dim] arrays in the buffer dtype; r is float32 [B]
 * (the reference returns (B,1) float32, her.py:38).  Any pointer may be NULL = skip. */
typedef struct bmi_transitions {
  void* obs;
  void* ag;
  void* g;
  void* actions;
  void* obs_next;
  void* ag_next;
  float* r;
} bmi_transitions;

/* replay_buffer.store_episode (replay_buffer.py:32-43): copy src episodes into
 * dst[slots[i]].  slots_dev: device int64[src->n_episodes], the indices chosen by
 * _get_storage_idx (replay_buffer.py:57-71, host logic); a negative slot skips that episode
 * (the caller marks all but the LAST of duplicated slots so that, like numpy's fancy
 * assignment, the last write wins deterministically).  src/dst dims must agree;
 * dtypes may differ (f64 -> f32 rounds to nearest, f32 -> f64 is exact). */
int bmi_buffer_store(const bmi_episodes* dst, const bmi_episodes* src,
                     const int64_t* slots_dev, bmi_stream_t stream);

/* compute_reward / goal_distance (bmirobot_env_push_F.py:20-23,84-90):
 * out[i] = -(float)(||ag_i - g_i||_2 > threshold), distance evaluated in float64 with
 * numpy's operation order (no FMA contraction) so the comparison is bit-exact. */
int bmi_compute_reward(const void* ag_dev, const void* g_dev, int64_t n, int32_t goal_dim,
                       int32_t dtype, double threshold, float* out_dev, bmi_stream_t stream);

/* her_sampler.sample_her_transitions (her.py:13-41) with the four random arrays supplied
 * by the caller (the drop-in path draws them from numpy's global legacy stream in the
 * reference order: randint, randint, uniform, uniform — her.py:24-31):
 *   ep_idx[b] in [0, n_valid), t_idx[b] in [0, T), u_her[b], u_off[b] in [0,1).
 *   relabel iff u_her[b] < future_p;  future_t = t + 1 + (int)(u_off[b] * (T - t))
 *   g <- ag[ep, future_t] for relabelled rows;  r = reward(ag_next, g).
 * All index arithmetic and comparisons are done in float64/int64 exactly as numpy does. */
int bmi_her_sample(const bmi_episodes* buf, int64_t n_valid, const int64_t* ep_idx_dev,
                   const int64_t* t_idx_dev, const double* u_her_dev, const double* u_off_dev,
                   int64_t B, double future_p, double threshold, const bmi_transitions* out,
                   bmi_stream_t stream);

/* Same gather fused with ddpg_agent._preproc_og + normalizer.normalize + concat + the
 * float32 cast (ddpg_agent.py:229-248): writes the network inputs directly
 *   x[B][obs+goal]      = [norm_o(clip(obs)), norm_g(clip(g))]
 *   x_next[B][obs+goal] = [norm_o(clip(obs_next)), norm_g(clip(g))]
 *   actions[B][act]  r[B]        (all float32)
 * normalisation evaluated in float64 then rounded once to float32, as the reference does. */
int bmi_her_sample_inputs(const bmi_episodes* buf, int64_t n_valid, const int64_t* ep_idx_dev,
                          const int64_t* t_idx_dev, const double* u_her_dev,
                          const double* u_off_dev, int64_t B, double future_p, double threshold,
                          double clip_obs, double clip_range, const float* o_mean_dev,
                          const float* o_std_dev, const float* g_mean_dev, const float* g_std_dev,
                          float* x_dev, float* x_next_dev, float* actions_dev, float* r_dev,
                          bmi_stream_t stream);

/* Device-side draw of the four HER random arrays with a counter-based Philox4x32-10
 * generator (used when sampling happens inside a captured update graph, where numpy's
 * host stream is not available).  counter is read from *counter_dev and advanced by B so
 * replaying a graph yields fresh draws.  oracle/philox.py + oracle/learner_oracle.py (her_draw_philox) restate the generator. */
int bmi_her_draw(uint64_t seed, uint64_t* counter_dev, int64_t B, const int64_t* n_valid_dev,
                 int32_t T, int64_t* ep_idx_dev, int64_t* t_idx_dev, double* u_her_dev,
                 double* u_off_dev, bmi_stream_t stream);

/* ------------------------------------------------------------------------------------
 * normalizer (normalizer.py:5-70).  All accumulators are float32 device arrays owned by
 * the caller: local_sum[size], local_sumsq[size], local_count[1], total_*[...], mean, std.
 * ---------------------------------------------------------------------------------- */
/* normalizer.update (normalizer.py:25-31): v is [n_rows][size] in dtype.  pre_clip > 0 applies
 * ddpg_agent._preproc_og's np.clip(v, -pre_clip, pre_clip) (ddpg_agent.py:214-217) to each element
 * before it is accumulated; pass 0 (or inf) for no clipping. */
int bmi_norm_update(const void* v_dev, int64_t n_rows, int32_t size, int32_t dtype, double pre_clip,
                    float* local_sum_dev, float* local_sumsq_dev, float* local_count_dev,
                    bmi_stream_t stream);
/* normalizer.recompute_stats (normalizer.py:40-57).  The caller first SUMS local_* over
 * ranks (bmi_comm_allreduce_sum_f32; nothing to do for one rank); this call divides by
 * `world` (the "/= Get_size()" of _mpi_average, normalizer.py:60-64), folds the result
 * into total_*, resets local_* to zero and recomputes mean/std with numpy-1.19 float32
 * semantics. */
int bmi_norm_recompute(float* local_sum_dev, float* local_sumsq_dev, float* local_count_dev,
                       float* total_sum_dev, float* total_sumsq_dev, float* total_count_dev,
                       float* mean_dev, float* std_dev, int32_t size, float eps, float world,
                       bmi_stream_t stream);
/* normalizer.normalize (normalizer.py:67-70): out = clip((v - mean)/std, +-clip_range) in
 * float64; out has dtype out_dtype. */
int bmi_norm_normalize(const void* v_dev, int64_t n_rows, int32_t size, int32_t dtype,
                       const float* mean_dev, const float* std_dev, double clip_range,
                       void* out_dev, int32_t out_dtype, bmi_stream_t stream);
/* ddpg_agent._preproc_inputs (ddpg_agent.py:163-171) for n rows at once:
 * x[n][obs+goal] = float32([normalize_o(obs), normalize_g(g)]) (no +-200 pre-clip there). */
int bmi_preproc_inputs(const void* obs_dev, const void* g_dev, int64_t n, int32_t obs_dim,
                       int32_t goal_dim, int32_t dtype, const float* o_mean_dev,
                       const float* o_std_dev, const float* g_mean_dev, const float* g_std_dev,
                       double clip_range, float* x_dev, bmi_stream_t stream);

/* ------------------------------------------------------------------------------------
 * DDPG learner (models.py:11-44, ddpg_agent.py:220-277).  Parameters live in caller-owned
 * flat float32 device buffers laid out in torch named_parameters order
 * (fc1.weight[out][in], fc1.bias, fc2.weight, ..., action_out|q_out.bias — utils.py:18-27),
 * so torch nn.Module parameters can alias them and checkpoints keep the reference format.
 * ---------------------------------------------------------------------------------- */
typedef struct bmi_ddpg bmi_ddpg;

typedef struct bmi_ddpg_config {
  int32_t obs_dim;    /* 27 */
  int32_t goal_dim;   /* 3  */
  int32_t act_dim;    /* 4  */
  int32_t hidden;     /* 256 */
  int32_t batch;      /* rows per update (256) */
  int32_t max_act_rows; /* largest n accepted by bmi_ddpg_act */
  float action_max;   /* 0.5 */
  float gamma;        /* 0.98 */
  float action_l2;    /* 1.0 */
  float lr_actor;     /* 1e-3 */
  float lr_critic;    /* 1e-3 */
  float polyak;       /* 0.95 */
  float adam_beta1;   /* 0.9 */
  float adam_beta2;   /* 0.999 */
  float adam_eps;     /* 1e-8 */
  float clip_return;  /* 1/(1-gamma) evaluated in double by the caller (ddpg_agent.py:259): 50 */
  float one_minus_polyak; /* (1 - polyak) evaluated in double by the caller, as python does
                             (ddpg_agent.py:222), then rounded once to float32: 0.05f */
  float _pad;
} bmi_ddpg_config;

int64_t bmi_ddpg_actor_param_count(const bmi_ddpg_config* cfg);
int64_t bmi_ddpg_critic_param_count(const bmi_ddpg_config* cfg);

/* The four flat buffers are caller-owned (actor, critic, actor target, critic target). */
int bmi_ddpg_create(bmi_ddpg** out, const bmi_ddpg_config* cfg, float* actor_params_dev,
                    float* critic_params_dev, float* actor_target_dev, float* critic_target_dev);
int bmi_ddpg_destroy(bmi_ddpg* h);

/* actor forward for n rows: actions = action_max * tanh(MLP(x)) (models.py:20-26).
 * use_target != 0 evaluates the target actor. */
int bmi_ddpg_act(bmi_ddpg* h, const float* x_dev, int64_t n, int32_t use_target,
                 float* actions_dev, bmi_stream_t stream);

/* With hidden == 256, batch % 32 == 0, obs + goal + act <= 64 and act <= 8 (the reference's shapes) this is two hand-written
 * kernels (csrc/ddpg_fused.cuh); otherwise, or with BMI_DDPG_CUBLAS=1 in the environment at bmi_ddpg_create time, a chain
 * of cuBLASLt GEMMs + small kernels.  Same results up to fp32 summation order.
 * ddpg_agent._update_network up to and including both backward passes
 * (ddpg_agent.py:250-270,274-275): fills the flat gradient buffers (actor then critic,
 * contiguous: one allreduce covers both — utils.py:43-48 sums, it does not average) and
 * writes losses_dev[0] = actor_loss, losses_dev[1] = critic_loss. */
int bmi_ddpg_backward(bmi_ddpg* h, const float* x_dev, const float* x_next_dev,
                      const float* actions_dev, const float* r_dev, float* losses_dev,
                      bmi_stream_t stream);
/* flat gradient buffer: [actor grads | zero pad to a multiple of 64 floats | critic grads],
 * n = total length; one allreduce over it covers both nets. */
int bmi_ddpg_grad_buffer(bmi_ddpg* h, float** grads_dev, int64_t* n);
/* both Adam steps (ddpg_agent.py:272,277; torch.optim.Adam defaults, bias-corrected). */
int bmi_ddpg_adam_step(bmi_ddpg* h, bmi_stream_t stream);
/* Fused gradient sum over ranks + both Adam steps through NVLink peer memory (one process per GPU on one node):
 * replaces [bmi_comm_allreduce_sum_f32 -> bmi_ddpg_adam_step], i.e. sync_grads (utils.py:43-48, SUM) + the two
 * optimiser steps (ddpg_agent.py:272,277), by ONE kernel that reads every rank's gradient buffer with peer loads,
 * adds them in rank order and updates the local replica; two flag barriers in peer memory order it against the
 * neighbours' backward passes.  Set-up: every rank calls bmi_ddpg_p2p_export (128 bytes: two cudaIpcMemHandle_t),
 * the caller all-gathers them, then bmi_ddpg_p2p_attach(rank, world, world x 128 bytes).  world <= 8.
 * bmi_ddpg_p2p_status reports whether a flag wait ever timed out (about 30 s of SM clocks).  A time-out is FATAL and
 * sticky: that launch and every later one return without touching the parameters, so the replicas cannot diverge
 * silently; the caller must stop (ddpg_agent.learn raises). */
int bmi_ddpg_p2p_export(bmi_ddpg* h, void* handles128_host);
int bmi_ddpg_p2p_attach(bmi_ddpg* h, int32_t rank, int32_t world, const void* all_handles_host);
int bmi_ddpg_adam_step_p2p(bmi_ddpg* h, bmi_stream_t stream);
int bmi_ddpg_p2p_status(bmi_ddpg* h, int32_t* timed_out);
/* _soft_update_target_network for both nets (ddpg_agent.py:220-222). */
int bmi_ddpg_soft_update(bmi_ddpg* h, bmi_stream_t stream);

/* _select_actions for n rows (ddpg_agent.py:174-184) with a Philox stream:
 * a = clip(pi + noise_eps*action_max*N(0,1), +-action_max); with prob random_eps replace by
 * U(-action_max, action_max); then optional clip to +-late_clip (ddpg_agent.py:118-119,
 * late_clip <= 0 disables). */
int bmi_select_actions(const float* pi_dev, int64_t n, int32_t act_dim, float action_max,
                       float noise_eps, float random_eps, float late_clip, uint64_t seed,
                       uint64_t* counter_dev, float* actions_dev, bmi_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Vectorised bmirobot environment (bmirobot_env_push_F.py:92-245, bmirobot.py:129-191,
 * bmirobot_inverse_kinematics.py:28-33).  One warp per env instance, 28 envs per thread block.
 * ---------------------------------------------------------------------------------- */
typedef struct bmi_env bmi_env;
#define BMI_TASK_PUSH 0
#define BMI_TASK_PICK 1
#define BMI_OBS_DIM 27
#define BMI_GOAL_DIM 3
#define BMI_ACT_DIM 4

/* model_blob: host pointer to the baked robot model (assets/bmirobot_model.bin, produced
 * by tools/bake_model.py from the reference URDF + meshes); copied to the device once. */
int bmi_env_create(bmi_env** out, int32_t n_envs, int32_t task, const void* model_blob,
                   int64_t model_bytes);
int bmi_env_destroy(bmi_env* h);
int32_t bmi_env_num_envs(const bmi_env* h);
/* Self-collision of the arm (bmirobot.py:58: loadURDF(..., flags=9) = URDF_USE_SELF_COLLISION).  table: host pointer to
 * the baked pair tables (assets/bmirobot_selfcol.bin, tools/bake_selfcol.py; layout include/bmi_model.h SC_*), copied to
 * the device once.  A model blob with MP_SELF_COLLISION = 1 cannot be stepped before this call (BMI_ERR_ARG). */
int bmi_env_set_selfcol(bmi_env* h, const void* table, int64_t table_bytes);
/* statistics: contacts the kernel dropped because the solver's lane budget was full (9 contacts, 6 on arm links) since
 * the last reset of the counter; host_out may be NULL.  Synchronises the device. */
int bmi_env_contact_drops(bmi_env* h, uint64_t* host_out, int32_t reset);
/* reset (bmirobot_env_push_F.py:110-165) of the envs whose mask byte is non-zero (NULL =
 * all).  init_dev: float32 [n_envs][8] = block x,y,z,yaw, goal x,y,z, unused — drawn by
 * the caller (python `random` stream for the drop-in env, bmi_env_sample_init for the
 * vectorised one).  Writes obs[n][27], ag[n][3], g[n][3] (float32) for ALL envs. */
int bmi_env_reset(bmi_env* h, const uint8_t* mask_dev, const float* init_dev, float* obs_dev,
                  float* ag_dev, float* g_dev, bmi_stream_t stream);
/* rejection-sampled block/goal placement with the reference ranges
 * (bmirobot_env_push_F.py:117-132; pick: bmirobot_env_pickandplace_v2.py:116-131). */
int bmi_env_sample_init(bmi_env* h, uint64_t seed, uint64_t* counter_dev, float* init_dev,
                        bmi_stream_t stream);
/* step (bmirobot_env_push_F.py:92-108): clip, IK, 9 motor targets, n_substeps x
 * stepSimulation, observation, sparse reward, is_success.  All arrays float32 device. */
int bmi_env_step(bmi_env* h, const float* actions_dev, float* obs_dev, float* ag_dev,
                 float* reward_dev, float* success_dev, bmi_stream_t stream);
/* Fused rollout (ddpg_agent.py:103-141 for all envs in ONE launch): reset from `init`, then T times
 * [record obs/ag/g -> _preproc_inputs -> actor MLP -> _select_actions -> record action -> env step],
 * then record the final obs/ag.  Each env advances at its own pace (no per-step grid-wide
 * synchronisation).  The actor weights are read in the transposed layout written by
 * bmi_actor_transpose; exploration uses the same Philox stream as bmi_select_actions
 * (counter + t * n_envs + env), so fused and step-wise rollouts draw identical noise. */
typedef struct bmi_rollout_args {
  int32_t T;
  int32_t explore;               /* 0: deterministic policy (evaluation, ddpg_agent.py:280-304) */
  const float* actor_t;          /* device, transposed actor parameters (same count as the flat buffer) */
  const float* o_mean; const float* o_std; const float* g_mean; const float* g_std;  /* device */
  float clip_range, action_max, noise_eps, random_eps, late_clip;
  uint64_t seed;
  uint64_t* counter;             /* device Philox counter, advanced by T * n_envs */
  const bmi_episodes* episodes;  /* float32 rollout staging [n_envs][T+1|T][dim] or NULL */
  const float* init;             /* device [n_envs][8] placements or NULL (continue from the current state) */
  float* obs; float* ag; float* g; float* success;   /* device outputs after the last step (may be NULL) */
} bmi_rollout_args;
int bmi_env_rollout(bmi_env* h, const bmi_rollout_args* args, bmi_stream_t stream);
/* torch-layout flat actor parameters (W[out][in], b per layer) -> W^T[in][out], b per layer */
int bmi_actor_transpose(const float* actor_params_dev, int32_t obs_dim, int32_t goal_dim, int32_t act_dim,
                        int32_t hidden, float* actor_t_dev, bmi_stream_t stream);
/* raw per-env simulator state for tests/checkpoints: float32 [n_envs][BMI_ENV_STATE_DIM] */
#define BMI_ENV_STATE_DIM 48
int bmi_env_get_state(bmi_env* h, float* state_dev, bmi_stream_t stream);
int bmi_env_set_state(bmi_env* h, const float* state_dev, bmi_stream_t stream);

/* rollout helper for the vectorised agent (ddpg_agent.py:113-130): writes obs/ag/g/action
 * of time step t into episode arrays ep (float32 or float64) for all n envs. */
int bmi_rollout_record(const bmi_episodes* ep, int32_t t, const float* obs_dev,
                       const float* ag_dev, const float* g_dev, const float* actions_dev,
                       bmi_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Collectives (utils.py:6-15,43-48; normalizer.py:60-64) — NCCL over NVLink, one rank per
 * GPU.  id128 is the 128-byte ncclUniqueId produced on rank 0 and distributed by the
 * caller (e.g. torch.distributed broadcast).
 * ---------------------------------------------------------------------------------- */
typedef struct bmi_comm bmi_comm;
int bmi_comm_unique_id(void* id128_host);
int bmi_comm_init(bmi_comm** out, int32_t rank, int32_t world, const void* id128_host);
int bmi_comm_destroy(bmi_comm* c);
int bmi_comm_allreduce_sum_f32(bmi_comm* c, float* buf_dev, int64_t n, bmi_stream_t stream);
int bmi_comm_bcast_f32(bmi_comm* c, float* buf_dev, int64_t n, int32_t root, bmi_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* BMI_B200_H_ */
