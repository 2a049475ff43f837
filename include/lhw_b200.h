/*
 * lhw_b200.h — C-ABI of the B200-native rollout / PPO data path (liblhw_b200.so).
 *
 * The reference (rohanpsingh/LearningHumanoidWalking) has no FFI of its own: its seams are Python
 * duck-typed protocols and the arithmetic sits behind the `mujoco` pybind module.  Each entry point
 * below names the reference interface it replaces (paths relative to the reference repo root).
 * Plain pointers and sizes only; all buffers are caller-owned DEVICE memory (the Python host uses
 * torch tensors) unless a parameter says "host".  `stream` is a cudaStream_t passed as void*.
 * Every function returns 0 on success, a negative code on error (lhw_last_error() has the text).
 *
 * `precision` is 64 (float64 state and arithmetic, the reference's numeric type) or 32 (float32).
 * Typed `void*` buffers hold doubles or floats accordingly.
 */
#ifndef LHW_B200_H
#define LHW_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lhw_sim lhw_sim;

int lhw_version(void);
const char* lhw_last_error(void);

/* ---- environment batch --------------------------------------------------------------------------
 * lhw_sim_create: replaces MujocoEnv.__init__ (envs/common/mujoco_env.py:16-36: MjSpec.compile +
 * MjData) and JvrcBaseEnv._setup_robot (envs/jvrc/jvrc_base.py:38-67): takes the compiled model
 * constants as a flat HOST float64 array (layout: learninghumanoidwalking_b200/model/loader.py).
 * The first word selects the robot/task variant: 106 = JVRC-1 + SteppingTask (envs/jvrc/jvrc_step.py,
 * tasks/stepping_task.py: footstep sequences, 20 per-env stepping-stone slabs, floor dropped in FORWARD mode),
 * 206 = JVRC-1 + WalkingTask on uneven / compliant terrain (an EXTENSION: the reference only has the unused
 * WalkingTask(manip_hfield) hook, tasks/walking_task.py:57,172-179; 20 terraces re-posed with that hook's ranges),
 * 6 = JVRC-1 + WalkingTask (envs/jvrc/jvrc_walk.py),
 * 5 = Unitree H1 + StandingTask (envs/h1/h1_env.py, envs/h1/h1_base.py:31-63: mass overrides, PD gains,
 * StandingTask), whose step also runs the observation noise, dynamics randomisation, random pushes and
 * initial-pose noise of envs/common/base_humanoid_env.py:228-338 + envs/common/domain_randomization.py inside
 * the kernel (per-env randomised parameters live at the tail of the state record). */
int lhw_sim_create(lhw_sim** out, const double* model_flat_host, int n_flat, int precision, int device);
int lhw_sim_destroy(lhw_sim* sim);
int lhw_sim_state_reals(const lhw_sim* sim); /* real words per env in the state record */
int lhw_sim_state_ints(const lhw_sim* sim);  /* int32 words per env in the state record */
int lhw_sim_obs_dim(const lhw_sim* sim);     /* env.observation_space.shape[0] (37 jvrc_walk, 39 jvrc_step, 35 h1) */
int lhw_sim_act_dim(const lhw_sim* sim);     /* env.action_space.shape[0] (12 jvrc_walk, 10 h1) */
int lhw_sim_smem_bytes_per_env(const lhw_sim* sim);
int lhw_sim_precision(const lhw_sim* sim);   /* 64 or 32 */
int lhw_sim_device(const lhw_sim* sim);      /* CUDA device ordinal the sim was created on */

/* lhw_sim_reset: MujocoEnv.reset + BaseHumanoidEnv.reset_model + WalkingTask.reset
 * (envs/common/mujoco_env.py:113-116, envs/common/base_humanoid_env.py:247-309,
 * tasks/walking_task.py:194-205 / tasks/standing_task.py:33) for every env whose mask[i] != 0 (mask == NULL: all).
 * `fresh` != 0 zero-initialises the record first (a newly constructed env).
 * state_r: [n_envs, state_reals] reals; state_i: [n_envs, state_ints] int32; obs: [n_envs, obs_dim]. */
int lhw_sim_reset(lhw_sim* sim, void* state_r, int32_t* state_i, int n_envs, uint32_t seed, uint32_t first_env_id,
                  const int32_t* mask, int fresh, void* obs, void* stream);

/* lhw_sim_step: BaseHumanoidEnv.step (envs/common/base_humanoid_env.py:199-227) ->
 * RobotBase.step/_do_simulation (robots/robot_base.py:41-98) -> frame_skip x {RobotInterface.step_pd,
 * set_motor_torque, mujoco.mj_step} (envs/common/robot_interface.py:493-546) -> WalkingTask.step /
 * calc_reward / done (tasks/walking_task.py:85-192; tasks/standing_task.py:49-131 for h1) -> get_obs (base_humanoid_env.py:177-197), for
 * n_envs environments in one launch.  With autoreset != 0 it also does the RolloutWorker's episode
 * bookkeeping (rl/workers/rollout_worker.py:146-176): ended = done || traj_len >= max_traj_len; an ended env
 * is reset inside the same launch and `obs` is the post-reset observation while `term_obs` keeps the
 * pre-reset one (for the truncation bootstrap critic(next_state)).
 * actions [n,act_dim]; obs, term_obs [n,obs_dim]; reward [n]; rew_terms [n,10] (may be NULL);
 * done, ended, ep_len [n] int32; ep_rew [n]  (ep_len/ep_rew written only where ended). */
int lhw_sim_step(lhw_sim* sim, void* state_r, int32_t* state_i, int n_envs, uint32_t seed, uint32_t first_env_id,
                 const void* actions, int max_traj_len, int autoreset, void* obs, void* term_obs, void* reward,
                 void* rew_terms, int32_t* done, int32_t* ended, int32_t* ep_len, void* ep_rew, void* stream);

/* lhw_sim_bind: make `sim`'s model constants the resident ones (they live in one __constant__ object per (robot, precision),
 * shared by all sims of the process; lhw_sim_step/reset do this implicitly, a replayed CUDA graph cannot). */
int lhw_sim_bind(lhw_sim* sim, void* stream);

/* lhw_sim_set_step_height: SteppingTask curriculum (tasks/stepping_task.py:312): the host computes
 * h = clip((iteration_count - 3000) / 8000, 0, 1) * 0.1 from `env.robot.iteration_count` (rl/workers/rollout_worker.py:95)
 * and passes it here; it takes effect at the next task reset.  No-op for the other variants. */
int lhw_sim_set_step_height(lhw_sim* sim, double h);

/* number of kernels this library has launched since load (the bench's gpu_launches claim) */
long long lhw_launch_count(void);

/* ---- PPO data path (float32, as rl/algos/ppo.py:474-477 casts the batch) ---------------------------
 * lhw_gae: PPOBuffer.finish_path (rl/storage/rollout_storage.py:53-85) for a [T, N] rollout stored time-major:
 * per env a reverse scan with the path broken wherever ended[t] != 0, bootstrapping with boot[t] there
 * ((not done) * critic(next_state), rl/workers/rollout_worker.py:166) and with last_val[n] after t = T-1
 * when the final transition did not end an episode (rollout_worker.py:183-186).
 * rewards, values, boot [T,N] f32; ended [T,N] int32; last_val [N]; returns [T,N] out.
 * adv_partials (may be NULL): lhw_gae_partial_words(N) doubles; the launch then also leaves per-block (sum, sumsq) of
 * returns - values there, which lhw_adv_stats_from_gae folds into stats[0:2] — the advantage normalisation
 * (rl/algos/ppo.py:484-485) then costs one 12 B/sample pass (lhw_adv_apply) instead of a statistics pass plus that. */
int lhw_gae(const float* rewards, const float* values, const int32_t* ended, const float* boot, const float* last_val,
            float* returns, int T, int N, float gamma, float lam, double* adv_partials_or_null, void* stream);
int lhw_gae_partial_words(int N);
int lhw_adv_stats_from_gae(const double* adv_partials, int N, double* stats, void* stream);

/* lhw_adv_norm: rl/algos/ppo.py:484-485 — adv = returns - values; (adv - mean) / (std_unbiased + eps).
 * stats: device scratch of lhw_adv_stats_words() doubles, {sum, sumsq, mean, std, per-block partials...}
 * (partials are combined in a fixed order, so the result is run-to-run deterministic). With world > 1 the
 * caller all-reduces stats[0:2] between lhw_adv_stats and lhw_adv_apply (count_total = global sample count). */
int lhw_adv_stats_words(void);
int lhw_adv_stats(const float* returns, const float* values, double* stats, long long count, void* stream);
int lhw_adv_apply(const float* returns, const float* values, float* adv, double* stats, long long count,
                  long long count_total, float eps, void* stream);

/* lhw_ppo_loss: the loss tail of PPO.update_actor_critic (rl/algos/ppo.py:302-386), forward and backward in one launch, for a
 * Gaussian policy with fixed per-action std: from the policy means mu / old_mu [B,A], the taken actions, advantages, returns,
 * values and (optional) mirrored actions it writes out8 = {actor_loss, entropy_penalty, critic_loss, approx_kl_div, mirror_loss,
 * imitation_loss (0), clip_fraction, total} and the gradients of total = actor + mirror_coeff * mirror + ent_coeff * entropy_penalty
 * + critic w.r.t. mu (g_mu), the mirrored actions (g_mirr) and the values (g_val).  partials: lhw_ppo_loss_partial_words(B)
 * doubles; ticket: one zero-initialised uint32.  Deterministic (block-ordered reduction). */
int lhw_ppo_loss_partial_words(int B);
int lhw_ppo_loss(const float* mu, const float* old_mu, const float* act, const float* adv, const float* ret, const float* val,
                 const float* mirr_or_null, const float* stds, int B, int A, float clip, float mirror_coeff, float ent_coeff,
                 float* g_mu, float* g_mirr_or_null, float* g_val, double* partials, unsigned int* ticket, float* out8, void* stream);

/* lhw_linear_wgrad: what `loss.backward()` (rl/algos/ppo.py:389-392) does per Linear layer of the actor / critic MLPs
 * (rl/policies/actor.py:122-189, critic.py) for its parameters: gW [N,K] = gy [M,N]^T x [M,K] and gb [N] = column sums of gy
 * (gb_or_null = NULL: weights only), M = minibatch rows, row-major contiguous operands.  One streaming launch over (gy, x) plus
 * an ordered reduction of the batch slices (deterministic).  workspace: lhw_linear_wgrad_workspace_floats(M, N, K) floats of
 * device scratch, free to reuse once the call's work has run on `stream`.  M = 0 zero-fills the outputs. */
long long lhw_linear_wgrad_workspace_floats(int M, int N, int K);
int lhw_linear_wgrad(const float* gy, const float* x, int M, int N, int K, float* gW, float* gb_or_null, float* workspace,
                     void* stream);

/* lhw_gather_minibatch: the fancy-index gathers of rl/algos/ppo.py:535-538 in one launch.
 * idx [B] int64 sample indices into the flattened batch. */
int lhw_gather_minibatch(const float* obs, const float* act, const float* ret, const float* adv, const int64_t* idx,
                         float* obs_b, float* act_b, float* ret_b, float* adv_b, int B, int obs_dim, int act_dim,
                         void* stream);

/* lhw_clip_adam: torch.nn.utils.clip_grad_norm_(params, max_norm) followed by optim.Adam.step
 * (rl/algos/ppo.py:393-396, Adam(lr, eps) created at :429-430) over ONE flat parameter buffer:
 * lhw_grad_sumsq accumulates sum(g^2) into norm_scratch[0] (zeroed by the call), lhw_clip_adam applies
 * clip_coef = min(1, max_norm / (sqrt(sumsq) + 1e-6)) and the Adam update (beta1 .9, beta2 .999, no weight
 * decay, no amsgrad). grad_scale multiplies g first (1/world after a sum all-reduce). */
int lhw_grad_sumsq(const float* grad, float* norm_scratch, long long n, float grad_scale, void* stream);
int lhw_clip_adam(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, const float* norm_scratch,
                  long long n, int step, float lr, float beta1, float beta2, float eps, float max_norm,
                  float grad_scale, void* stream);
/* same update, Adam step number kept in DEVICE memory (step_dev[0] = completed steps, incremented by the call): the whole
 * optimiser step of rl/algos/ppo.py:389-396 can then be replayed from a CUDA graph */
int lhw_clip_adam_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, const float* norm_scratch,
                      long long n, int* step_dev, float lr, float beta1, float beta2, float eps, float max_norm,
                      float grad_scale, void* stream);

/* ---- multi-GPU exchange step, fused (csrc/comm_kernels.cu) ------------------------------------------
 * Replaces, on N >= 1 GPUs, the sequence of rl/algos/ppo.py:389-396: (gradient all-reduce) -> clip_grad_norm_(actor),
 * clip_grad_norm_(critic) -> actor Adam.step, critic Adam.step, by three ordinary (CUDA-graph capturable) launches that
 * read the peers' flat gradients over NVLink peer memory (CUDA IPC, 16-byte loads), sum them in fixed rank order, and
 * apply clip + Adam; the Adam step number and the peer-barrier epoch live in device memory.
 * lhw_comm owns a cudaMalloc'ed IPC-exported gradient buffer of n_floats floats (lhw_comm_grad_ptr) — the host
 * makes the modules' .grad tensors views of it. Handles are exchanged by the host (torch.distributed
 * all_gather of lhw_comm_handle_size() bytes per rank) and mapped with lhw_comm_import.
 * n_actor: the first n_actor floats belong to the actor (own norm / clip), the rest to the critic.
 * lhw_comm_status: error word of the peer barrier (non-zero: a peer did not arrive within the spin limit; returns -20),
 * completed Adam steps, the two gradient norms of the last step; synchronises `stream`.
 * lhw_comm_set_step: set the device-side Adam step number (resume). */
typedef struct lhw_comm lhw_comm;
const char* lhw_comm_last_error(void);
int lhw_comm_handle_size(void);
int lhw_comm_create(lhw_comm** out, long long n_floats, int rank, int world, int device);
void* lhw_comm_grad_ptr(lhw_comm* comm);
long long lhw_comm_size(const lhw_comm* comm);   /* n_floats */
int lhw_comm_device(const lhw_comm* comm);
int lhw_comm_export(lhw_comm* comm, void* handle_blob_host);
int lhw_comm_import(lhw_comm* comm, const void* all_handle_blobs_host);
int lhw_comm_destroy(lhw_comm* comm);
int lhw_fused_allreduce_clip_adam(lhw_comm* comm, float* param, float* exp_avg, float* exp_avg_sq, long long n_actor,
                                  long long n_total, float lr, float beta1, float beta2, float eps, float max_norm,
                                  void* stream);
int lhw_comm_status(lhw_comm* comm, int* error_word, int* adam_steps, float* norms2, void* stream);
int lhw_comm_set_step(lhw_comm* comm, int adam_steps, void* stream);

#ifdef __cplusplus
}
#endif
#endif
