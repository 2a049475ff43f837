/*
 * oracle/sim_oracle.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, float64, one environment at a time) of the reference's rollout hot
 * path: BaseHumanoidEnv.step/reset -> RobotBase._do_simulation -> mujoco.mj_step -> WalkingTask
 * reward/done -> observation packing.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product path never does.
 *
 * PARITY STATUS: "parity unpinned" at the MuJoCo boundary.  The physics arithmetic of the
 * reference lives in the un-vendored third-party library mujoco==3.4.0 (pyproject.toml:12,
 * uv.lock:930-931) which is not installable here, and the reference's tests hold no golden
 * vectors for it (tests/test_environments.py:38-114 are property tests).  This file restates
 * MuJoCo's published algorithm and is anchored on (i) physical invariants
 * checked in tests/ (energy, momentum, M symmetric PD, static contact force = m g, KKT residual of
 * the constraint solve) and (ii) golden vectors for everything the reference's own importable
 * code pins (reward terms, gait clocks, GAE) under tests/golden/.
 */
#ifndef LHW_SIM_ORACLE_H
#define LHW_SIM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAXLINK 16
#define ORC_NV 18
#define ORC_NQ 19
#define ORC_NU 12
#define ORC_MAXGEOM 4
#define ORC_MAXCON 128
#define ORC_MAXPTS 8
#define ORC_MAXROW (2 * ORC_NU + 4 * ORC_MAXCON)
#define ORC_MAXPERIOD 128
#define ORC_NOBS 39
#define ORC_NREW 10
#define ORC_MAXCAP 16
#define ORC_MAXPAIR 64
#define ORC_NSLAB 20        /* stepping stones per env (envs/jvrc/gen_xml.py:147-153) */
#define ORC_MAXPLAN 128
#define ORC_MAXPLANLEN 20
#define ORC_MAXCROSS 4     /* sole-edge x slab-boundary contacts kept per foot */

enum { ORC_STANDING = 0, ORC_INPLACE = 1, ORC_FORWARD = 2 };
enum { ORC_SOLVER_NEWTON = 0, ORC_SOLVER_PGS = 1 };
enum { ORC_TASK_WALK = 0, ORC_TASK_STAND = 1, ORC_TASK_STEP = 2 };
/* SteppingTask walk modes in the order of the np.random.choice list (tasks/stepping_task.py:268-271) */
enum { ORC_STEP_CURVED = 0, ORC_STEP_STANDING = 1, ORC_STEP_BACKWARD = 2, ORC_STEP_LATERAL = 3, ORC_STEP_FORWARD = 4 };
enum { ORC_GEOM_BOX = 0, ORC_GEOM_SPHERES = 1 };

/* per-environment model parameters (domain randomisation edits these in place, envs/common/domain_randomization.py:29-56) */
typedef struct {
  double mass[ORC_MAXLINK], com[ORC_MAXLINK][3], inertia[ORC_MAXLINK][9];
  double damping[ORC_NV], frictionloss[ORC_NV];
  double pel_mass, pel_com[3];     /* the pelvis BODY inside the welded root link (xfrc point, composite rebuild) */
} orc_params;

typedef struct {
  int nlink, nv, nq, nu;
  int parent[ORC_MAXLINK];
  double pos[ORC_MAXLINK][3];      /* link origin in parent link frame */
  double rot[ORC_MAXLINK][9];      /* link orientation in parent link frame (row major) */
  double axis[ORC_MAXLINK][3];     /* hinge axis in link frame (links >= 1) */
  double mass[ORC_MAXLINK];
  double com[ORC_MAXLINK][3];
  double inertia[ORC_MAXLINK][9];  /* about com, link frame */
  double armature[ORC_NV], damping[ORC_NV];
  double range[ORC_NV][2];
  int limited[ORC_NV];
  double dof_invweight0[ORC_NV];
  double link_invweight0[ORC_MAXLINK][2];
  int ngeom;                       /* foot boxes */
  int geom_link[ORC_MAXGEOM];
  double geom_pos[ORC_MAXGEOM][3];
  double geom_size[ORC_MAXGEOM][3];
  double timestep, gravity[3];
  double solref[2], solimp[5], mu, impratio;
  double meaninertia, tolerance;
  int iterations, solver;
  /* env / task (envs/jvrc/configs/base.yaml, tasks/walking_task.py) */
  double kp[ORC_NU], kd[ORC_NU];
  double nominal_qpos[ORC_NQ];
  int frame_skip;
  double action_smoothing;
  int rfoot_link, lfoot_link;
  double head_in_root[3];
  double total_mass, goal_height;
  int period;
  double clock[4][ORC_MAXPERIOD];
  /* self-collision proxies (capsules on the leg hulls / feet; termination only, robot_interface.py:472-484) */
  int ncap, npair;
  int cap_link[ORC_MAXCAP];
  double cap_p0[ORC_MAXCAP][3], cap_p1[ORC_MAXCAP][3], cap_r[ORC_MAXCAP];
  int pair[ORC_MAXPAIR][2];
  /* task / robot variants */
  int task, nobs;
  double done_lo, done_hi;
  double obs_noise[5];              /* root_orient, root_ang_vel, motor_pos, motor_vel, motor_tau (0: off) */
  int dynrand_interval, perturb_interval;
  double perturb_force, perturb_torque, init_noise;
  int geom_type[ORC_MAXGEOM], geom_npts[ORC_MAXGEOM];
  double geom_pts[ORC_MAXGEOM][ORC_MAXPTS][3], geom_radius[ORC_MAXGEOM];
  /* root link = randomisable pelvis body + welded rest (H1) */
  double pel_mass, pel_com[3], pel_Ic[9], rest_mass, rest_mc[3], rest_Io[9], torso_com[3];
  double pdrand_k;                  /* RobotBase(pdrand_k): per-step PD gain randomisation, 0 = off (the reference's default) */
  /* jvrc_step (envs/jvrc/jvrc_step.py, tasks/stepping_task.py): force-sensor sites, stepping-stone slabs, footstep plans */
  double foot_site[2][3];           /* rf_force / lf_force site in the foot link frame (gen_xml.py:143-144) */
  double slab_half[3];              /* half sizes of a slab: 0.15 x 1 x box_h (stepping_task.py:327) */
  double target_radius, side_tol, step_height;   /* step_height: the curriculum's h for the current iteration_count */
  int delay_frames, nplan;
  int explicit_euler;              /* SURVEY App. A.1 switch: integrate qacc as solved (no (M + h B) solve); default 0 */
  int side_faces;                  /* 1: slab side faces (stair risers) stop box corners that are inside a slab; 0: round-1 behaviour (pass through) */
  int slab_contacts_are_floor;     /* 0 (reference behaviour, see sim_oracle.c): foot-slab contacts are invisible to get_*_floor_contacts */
  int plan_len[ORC_MAXPLAN];
  double plans[ORC_MAXPLAN][ORC_MAXPLANLEN][3];  /* utils/footstep_plans.txt: x, y, theta */
  /* uneven / compliant terrain EXTENSION (BASELINE configs[4]; the reference only has the unused WalkingTask(manip_hfield)
   * hook, tasks/walking_task.py:172-179, and no height-field asset — SURVEY F7): WalkingTask on 20 terraces (the slabs of
   * the stepping stones, yaw 0, tiling x), re-posed like the hook re-poses its "hfield" geom; softer contact solref */
  int terrain, terrain_interval;
  double terrain_pitch, terrain_bump, terrain_zlo, terrain_zhi, terrain_xy;
  double contact_solref[2];         /* solref of the foot-ground contacts ({0,0}: the model default) */
} orc_model;

typedef struct {
  /* mjData-like persistent state */
  double qpos[ORC_NQ], qvel[ORC_NV];
  double qacc_warm[ORC_NV];
  /* quantities mj_step leaves behind, i.e. evaluated at the PRE-integration state */
  double qacc[ORC_NV];
  double act_len[ORC_NU], act_vel[ORC_NU], act_force[ORC_NU];
  double root_xpos[3], root_xmat[9], head_xpos[3], root_vlin[3];
  double rfoot_vel[3], lfoot_vel[3];
  double rfoot_grf, lfoot_grf;
  double contact_z_min;
  int ncon_r, ncon_l, ncon;
  int self_collision;
  /* env/robot bookkeeping (base_humanoid_env.py, robot_base.py) */
  double prev_prediction[ORC_NU], prev_action[ORC_NU], prev_torque[ORC_NU];
  int have_prev;
  /* task state (walking_task.py) */
  int phase, mode;
  double mode_ref[3];
  /* rollout bookkeeping */
  int traj_len, ep_len;
  double ep_rew;
  /* rng: philox4x32-10 keyed by (seed, env_id), counter = (event counter, stream) */
  uint32_t seed, env_id, rng_ctr;
  /* diagnostics */
  int last_solver_iter;
  double last_kkt_residual;
  int status;                      /* nonzero: NaN / divergence seen (mj_checkAcc analogue) */
  int nsubsteps;
  /* H1 additions (appended so the jvrc field offsets used by the tests stay put) */
  orc_params P;
  double xfrc[2][6];               /* world-frame [force, torque] on the pelvis and torso bodies, applied at their CoM */
  /* SteppingTask state (tasks/stepping_task.py) */
  double seq[ORC_NSLAB][4];        /* footstep sequence x, y, z, theta (world); slab k's top face sits at seq[k] */
  double goal_steps[8];            /* _goal_steps_x[2], _y[2], _z[2], _theta[2] */
  double site_pos[2][3];           /* rf_force / lf_force site_xpos (lagged, like every mjData quantity) */
  double foot_xpos[2][3];          /* foot body xpos (lagged) */
  double root_quat[4];             /* data.xquat of the root body (lagged, normalised) */
  int seq_len, t1, t2, target_reached, target_reached_frames, con_overflow;
  /* diagnostics: ring of the last 32 substeps (index nsubsteps & 31): solver iterations, contact rows */
  int iter_trace[32], nrow_trace[32];
} orc_env;

/* fill a model from a flat double array (layout documented in oracle/oracle.py:pack_model) */
int orc_model_from_flat(orc_model* m, const double* buf, int n);
int orc_sizeof_env(void);
int orc_sizeof_model(void);

void orc_env_init(const orc_model* m, orc_env* e, uint32_t seed, uint32_t env_id);
/* MujocoEnv.reset + BaseHumanoidEnv.reset_model + WalkingTask.reset; writes obs[37] */
void orc_reset(const orc_model* m, orc_env* e, double* obs);
/* one mujoco.mj_step with data.ctrl = ctrl */
void orc_mj_step(const orc_model* m, orc_env* e, const double* ctrl);
/* BaseHumanoidEnv.step. done_out: task.done(). No auto-reset. */
void orc_step(const orc_model* m, orc_env* e, const double* action, double* obs, double* rew_terms,
              double* reward, int* done);
/* rollout-style step with the RolloutWorker's auto-reset semantics (rl/workers/rollout_worker.py:142-176):
 * obs = observation the policy sees next (post-reset when the episode ended);
 * term_obs = pre-reset observation (valid when ended); ended = done || traj_len >= max_traj_len */
void orc_step_autoreset(const orc_model* m, orc_env* e, const double* action, int max_traj_len, double* obs,
                        double* term_obs, double* rew_terms, double* reward, int* done, int* ended);
/* batched, OpenMP over envs (the CPU baseline) */
void orc_batch_reset(const orc_model* m, orc_env* envs, int n, double* obs, int nthreads);
void orc_batch_step_autoreset(const orc_model* m, orc_env* envs, int n, const double* actions, int max_traj_len,
                              double* obs, double* term_obs, double* rew_terms, double* reward, int* done, int* ended,
                              int nthreads);

/* building blocks exposed for tests */
void orc_mass_matrix(const orc_model* m, const double* qpos, double* M /* nv*nv */);
void orc_bias(const orc_model* m, const double* qpos, const double* qvel, double* c /* nv */);
double orc_energy(const orc_model* m, const double* qpos, const double* qvel, double* kinetic, double* potential);
void orc_philox(uint32_t seed, uint32_t env_id, uint32_t ctr, uint32_t stream, uint32_t out[4]);
void orc_quat2rp(const double* quat, double* roll, double* pitch);
int orc_max_threads(void);
void orc_calc_reward(const orc_model* m, const orc_env* e, const double* target, double* terms);

#ifdef __cplusplus
}
#endif
#endif
