/*
 * oracle/sim_oracle.c — TEST INFRASTRUCTURE (see sim_oracle.h for the parity statement).
 *
 * Formulations are deliberately the dense textbook ones (explicit link Jacobians, M = sum J^T I J,
 * projected Newton-Euler bias, dense Cholesky) so that this file is an independent check of the
 * tree-structured CUDA kernel, not a transliteration of it.
 *
 * Reference call chain restated (file:line relative to /root/reference):
 *   envs/common/base_humanoid_env.py:199-227  step        -> orc_step
 *   envs/common/base_humanoid_env.py:247-276  reset_model -> orc_reset
 *   robots/robot_base.py:41-98                _do_simulation/step (PD loop, prev_action/torque)
 *   envs/common/robot_interface.py:493-546    step_pd / set_motor_torque / step (mj_step)
 *   tasks/walking_task.py:85-205              calc_reward / step / done / reset
 *   tasks/rewards.py:9-174                    reward terms
 *   envs/jvrc/jvrc_base.py:133-145, envs/jvrc/jvrc_walk.py:65-67, tasks/observations.py:12-72  obs
 *   mujoco.mj_step (external)                                  -> orc_mj_step
 *   tasks/stepping_task.py:52-334, envs/jvrc/jvrc_step.py:41-76 (jvrc_step)          -> task_reset_step / task_step_step /
 *       calc_reward_step / slab contacts in make_constraints (mjc_BoxBox is inside the un-vendored library: face-face case restated)
 *   tasks/standing_task.py, envs/h1/*, envs/common/domain_randomization.py (h1)      -> calc_reward_stand / randomize_dynamics / ...
 */
#include "sim_oracle.h"

#include <math.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define L ORC_MAXLINK
#define NV ORC_NV
#define MINVAL 1e-15
#define ORC_PGS_MAXROW 160   /* the dual solver is a test cross-check for small row counts only */

/* ------------------------------------------------------------------ small linear algebra */
static void cross3(const double* a, const double* b, double* c) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  c[0] = x; c[1] = y; c[2] = z;
}
static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void matvec3(const double* R, const double* v, double* out) {
  double x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
  double y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
  double z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
  out[0] = x; out[1] = y; out[2] = z;
}
static void mattvec3(const double* R, const double* v, double* out) {
  double x = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
  double y = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
  double z = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
  out[0] = x; out[1] = y; out[2] = z;
}
static void matmul3(const double* A, const double* B, double* C) {
  double T[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) T[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
  memcpy(C, T, sizeof(T));
}
static void quat2mat(const double* q, double* R) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = 1 - 2 * (x * x + y * y);
}
/* dense Cholesky A = G G^T (lower), in place; returns 0 on success */
static int chol(double* A, int n) {
  for (int j = 0; j < n; j++) {
    double d = A[j * n + j];
    for (int k = 0; k < j; k++) d -= A[j * n + k] * A[j * n + k];
    if (!(d > 0)) return 1;
    d = sqrt(d);
    A[j * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      double s = A[i * n + j];
      for (int k = 0; k < j; k++) s -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = s / d;
    }
  }
  return 0;
}
static void chol_solve(const double* G, int n, double* b) {
  for (int i = 0; i < n; i++) {
    double s = b[i];
    for (int k = 0; k < i; k++) s -= G[i * n + k] * b[k];
    b[i] = s / G[i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = b[i];
    for (int k = i + 1; k < n; k++) s -= G[k * n + i] * b[k];
    b[i] = s / G[i * n + i];
  }
}

/* ------------------------------------------------------------------ model packing */
int orc_sizeof_env(void) { return (int)sizeof(orc_env); }
int orc_sizeof_model(void) { return (int)sizeof(orc_model); }
int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

int orc_model_from_flat(orc_model* m, const double* b, int n) {
  int p = 0;
#define RD() (p < n ? b[p++] : (p++, 0.0))
  memset(m, 0, sizeof(*m));
  m->nlink = (int)RD();
  if (m->nlink > L || m->nlink < 1) return -1;
  m->nv = 6 + m->nlink - 1; m->nq = m->nv + 1; m->nu = m->nlink - 1;
  if (m->nv > NV) return -2;
  for (int i = 0; i < m->nlink; i++) {
    m->parent[i] = (int)RD();
    for (int k = 0; k < 3; k++) m->pos[i][k] = RD();
    for (int k = 0; k < 9; k++) m->rot[i][k] = RD();
    for (int k = 0; k < 3; k++) m->axis[i][k] = RD();
    m->mass[i] = RD();
    for (int k = 0; k < 3; k++) m->com[i][k] = RD();
    for (int k = 0; k < 9; k++) m->inertia[i][k] = RD();
    for (int k = 0; k < 2; k++) m->link_invweight0[i][k] = RD();
  }
  for (int d = 0; d < m->nv; d++) {
    m->armature[d] = RD(); m->damping[d] = RD(); m->range[d][0] = RD(); m->range[d][1] = RD();
    m->limited[d] = (int)RD(); m->dof_invweight0[d] = RD();
  }
  m->ngeom = (int)RD();
  if (m->ngeom > ORC_MAXGEOM) return -3;
  for (int g = 0; g < m->ngeom; g++) {
    m->geom_link[g] = (int)RD();
    for (int k = 0; k < 3; k++) m->geom_pos[g][k] = RD();
    for (int k = 0; k < 3; k++) m->geom_size[g][k] = RD();
  }
  m->timestep = RD();
  for (int k = 0; k < 3; k++) m->gravity[k] = RD();
  for (int k = 0; k < 2; k++) m->solref[k] = RD();
  for (int k = 0; k < 5; k++) m->solimp[k] = RD();
  m->mu = RD(); m->impratio = RD(); m->meaninertia = RD(); m->tolerance = RD();
  m->iterations = (int)RD(); m->solver = (int)RD();
  for (int k = 0; k < m->nu; k++) m->kp[k] = RD();
  for (int k = 0; k < m->nu; k++) m->kd[k] = RD();
  for (int k = 0; k < m->nq; k++) m->nominal_qpos[k] = RD();
  m->frame_skip = (int)RD(); m->action_smoothing = RD();
  m->rfoot_link = (int)RD(); m->lfoot_link = (int)RD();
  for (int k = 0; k < 3; k++) m->head_in_root[k] = RD();
  m->total_mass = RD(); m->goal_height = RD();
  m->period = (int)RD();
  if (m->period > ORC_MAXPERIOD) return -4;
  for (int c = 0; c < 4; c++)
    for (int k = 0; k < m->period; k++) m->clock[c][k] = RD();
  m->ncap = (int)RD();
  if (m->ncap > ORC_MAXCAP) return -5;
  for (int c = 0; c < m->ncap; c++) {
    m->cap_link[c] = (int)RD();
    for (int k = 0; k < 3; k++) m->cap_p0[c][k] = RD();
    for (int k = 0; k < 3; k++) m->cap_p1[c][k] = RD();
    m->cap_r[c] = RD();
  }
  m->npair = (int)RD();
  if (m->npair > ORC_MAXPAIR) return -6;
  for (int c = 0; c < m->npair; c++) { m->pair[c][0] = (int)RD(); m->pair[c][1] = (int)RD(); }
  m->task = (int)RD(); m->nobs = (int)RD(); m->done_lo = RD(); m->done_hi = RD();
  for (int k = 0; k < 5; k++) m->obs_noise[k] = RD();
  m->dynrand_interval = (int)RD(); m->perturb_interval = (int)RD();
  m->perturb_force = RD(); m->perturb_torque = RD(); m->init_noise = RD();
  for (int g = 0; g < m->ngeom; g++) {
    m->geom_type[g] = (int)RD(); m->geom_npts[g] = (int)RD(); m->geom_radius[g] = RD();
    if (m->geom_npts[g] > ORC_MAXPTS) return -7;
    for (int k = 0; k < m->geom_npts[g]; k++)
      for (int x = 0; x < 3; x++) m->geom_pts[g][k][x] = RD();
  }
  m->pel_mass = RD();
  for (int k = 0; k < 3; k++) m->pel_com[k] = RD();
  for (int k = 0; k < 9; k++) m->pel_Ic[k] = RD();
  m->rest_mass = RD();
  for (int k = 0; k < 3; k++) m->rest_mc[k] = RD();
  for (int k = 0; k < 9; k++) m->rest_Io[k] = RD();
  for (int k = 0; k < 3; k++) m->torso_com[k] = RD();
  m->pdrand_k = RD();
  m->terrain = (int)RD();
  if (m->terrain) {
    for (int k = 0; k < 3; k++) m->slab_half[k] = RD();
    m->side_tol = RD(); m->terrain_pitch = RD(); m->terrain_bump = RD(); m->terrain_zlo = RD(); m->terrain_zhi = RD();
    m->terrain_xy = RD(); m->terrain_interval = (int)RD();
    for (int k = 0; k < 2; k++) m->contact_solref[k] = RD();
    m->side_faces = (int)RD();
  }
  if (m->task == ORC_TASK_STEP) {
    for (int f = 0; f < 2; f++)
      for (int k = 0; k < 3; k++) m->foot_site[f][k] = RD();
    for (int k = 0; k < 3; k++) m->slab_half[k] = RD();
    m->target_radius = RD(); m->side_tol = RD(); m->delay_frames = (int)RD(); m->step_height = RD();
    m->slab_contacts_are_floor = (int)RD();
    m->side_faces = (int)RD();
    m->nplan = (int)RD();
    if (m->nplan > ORC_MAXPLAN) return -8;
    for (int i = 0; i < m->nplan; i++) {
      m->plan_len[i] = (int)RD();
      if (m->plan_len[i] > ORC_MAXPLANLEN) return -9;
      for (int k = 0; k < m->plan_len[i]; k++)
        for (int x = 0; x < 3; x++) m->plans[i][k][x] = RD();
    }
  }
  /* assumption switches (model JSON "assumptions"): bit 0 = explicit Euler instead of the implicit joint-damping solve */
  m->explicit_euler = ((int)RD()) & 1;
#undef RD
  return p == n ? 0 : -100 - (p > n);
}

/* ------------------------------------------------------------------ rng (philox4x32-10) */
static inline uint32_t mulhilo(uint32_t a, uint32_t b, uint32_t* hi) {
  uint64_t p = (uint64_t)a * b;
  *hi = (uint32_t)(p >> 32);
  return (uint32_t)p;
}
void orc_philox(uint32_t seed, uint32_t env_id, uint32_t ctr, uint32_t stream, uint32_t out[4]) {
  uint32_t c0 = ctr, c1 = stream, c2 = env_id, c3 = 0x4c485742u; /* 'LHWB' */
  uint32_t k0 = seed, k1 = 0x9E3779B9u ^ (seed * 0x85EBCA6Bu + 1u);
  for (int r = 0; r < 10; r++) {
    uint32_t hi0, hi1;
    uint32_t lo0 = mulhilo(0xD2511F53u, c0, &hi0);
    uint32_t lo1 = mulhilo(0xCD9E8D57u, c2, &hi1);
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static inline double u01(uint32_t u) { return (double)(u >> 8) * (1.0 / 16777216.0); } /* 24 bit, exact in f32 too */
static inline int randint(uint32_t u, int n) { return (int)(((uint64_t)u * (uint64_t)n) >> 32); }

static void params_default(const orc_model* m, orc_params* P) {
  memset(P, 0, sizeof(*P));
  for (int i = 0; i < m->nlink; i++) {
    P->mass[i] = m->mass[i];
    memcpy(P->com[i], m->com[i], sizeof(P->com[i]));
    memcpy(P->inertia[i], m->inertia[i], sizeof(P->inertia[i]));
  }
  for (int d = 0; d < m->nv; d++) P->damping[d] = m->damping[d];
  P->pel_mass = m->pel_mass;
  memcpy(P->pel_com, m->pel_com, sizeof(P->pel_com));
}

/* ------------------------------------------------------------------ kinematics */
typedef struct {
  double xpos[L][3], xmat[L][9], xaxis[L][3], xcom[L][3], Iw[L][9];
} kin_t;

static void fk(const orc_model* m, const orc_params* P, const double* qpos, kin_t* k) {
  double q[4] = {qpos[3], qpos[4], qpos[5], qpos[6]};
  double nrm = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; i++) q[i] /= nrm;
  for (int i = 0; i < m->nlink; i++) {
    if (i == 0) {
      memcpy(k->xpos[0], qpos, 3 * sizeof(double));
      quat2mat(q, k->xmat[0]);
      k->xaxis[0][0] = k->xaxis[0][1] = k->xaxis[0][2] = 0;
    } else {
      int p = m->parent[i];
      double off[3], R0[9], Rj[9];
      matvec3(k->xmat[p], m->pos[i], off);
      for (int c = 0; c < 3; c++) k->xpos[i][c] = k->xpos[p][c] + off[c];
      matmul3(k->xmat[p], m->rot[i], R0);
      /* Rodrigues about the (unit) hinge axis */
      const double* a = m->axis[i];
      double ang = qpos[6 + i], s = sin(ang), c = cos(ang), t = 1 - c;
      Rj[0] = c + a[0] * a[0] * t;        Rj[1] = a[0] * a[1] * t - a[2] * s; Rj[2] = a[0] * a[2] * t + a[1] * s;
      Rj[3] = a[0] * a[1] * t + a[2] * s; Rj[4] = c + a[1] * a[1] * t;        Rj[5] = a[1] * a[2] * t - a[0] * s;
      Rj[6] = a[0] * a[2] * t - a[1] * s; Rj[7] = a[1] * a[2] * t + a[0] * s; Rj[8] = c + a[2] * a[2] * t;
      matmul3(R0, Rj, k->xmat[i]);
      matvec3(k->xmat[i], a, k->xaxis[i]);
    }
    double sc[3], T[9], Rt[9];
    matvec3(k->xmat[i], P->com[i], sc);
    for (int c = 0; c < 3; c++) k->xcom[i][c] = k->xpos[i][c] + sc[c];
    matmul3(k->xmat[i], P->inertia[i], T);
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) Rt[3 * r + c] = k->xmat[i][3 * c + r];
    matmul3(T, Rt, k->Iw[i]);
  }
}

/* translational (jp) and rotational (jr) Jacobians (3 x nv, row major with stride NV) of a world point
 * rigidly attached to `link`; free joint: dofs 0-2 world translation, 3-5 body-frame rotation */
static void jac_point(const orc_model* m, const kin_t* k, int link, const double* point, double* jp, double* jr) {
  memset(jp, 0, 3 * NV * sizeof(double));
  if (jr) memset(jr, 0, 3 * NV * sizeof(double));
  for (int b = link; b >= 0; b = m->parent[b]) {
    if (b == 0) {
      for (int c = 0; c < 3; c++) jp[c * NV + c] = 1.0;
      double r[3] = {point[0] - k->xpos[0][0], point[1] - k->xpos[0][1], point[2] - k->xpos[0][2]};
      for (int kk = 0; kk < 3; kk++) {
        double a[3] = {k->xmat[0][kk], k->xmat[0][3 + kk], k->xmat[0][6 + kk]}, v[3];
        cross3(a, r, v);
        for (int c = 0; c < 3; c++) {
          jp[c * NV + 3 + kk] = v[c];
          if (jr) jr[c * NV + 3 + kk] = a[c];
        }
      }
    } else {
      int d = 5 + b;
      double r[3] = {point[0] - k->xpos[b][0], point[1] - k->xpos[b][1], point[2] - k->xpos[b][2]}, v[3];
      cross3(k->xaxis[b], r, v);
      for (int c = 0; c < 3; c++) {
        jp[c * NV + d] = v[c];
        if (jr) jr[c * NV + d] = k->xaxis[b][c];
      }
    }
  }
}

static void mass_matrix_k(const orc_model* m, const orc_params* P, const kin_t* k, double* M) {
  int nv = m->nv;
  memset(M, 0, NV * NV * sizeof(double));
  double jp[3 * NV], jr[3 * NV], IJ[3 * NV];
  for (int i = 0; i < m->nlink; i++) {
    jac_point(m, k, i, k->xcom[i], jp, jr);
    for (int c = 0; c < 3; c++)
      for (int d = 0; d < nv; d++)
        IJ[c * NV + d] = k->Iw[i][3 * c] * jr[d] + k->Iw[i][3 * c + 1] * jr[NV + d] + k->Iw[i][3 * c + 2] * jr[2 * NV + d];
    for (int a = 0; a < nv; a++)
      for (int b = 0; b < nv; b++) {
        double s = 0;
        for (int c = 0; c < 3; c++) s += P->mass[i] * jp[c * NV + a] * jp[c * NV + b] + jr[c * NV + a] * IJ[c * NV + b];
        M[a * NV + b] += s;
      }
  }
  for (int d = 0; d < nv; d++) M[d * NV + d] += m->armature[d];
}

void orc_mass_matrix(const orc_model* m, const double* qpos, double* Mout) {
  kin_t k; double M[NV * NV]; orc_params P;
  params_default(m, &P);
  fk(m, &P, qpos, &k);
  mass_matrix_k(m, &P, &k, M);
  for (int a = 0; a < m->nv; a++)
    for (int b = 0; b < m->nv; b++) Mout[a * m->nv + b] = M[a * NV + b];
}

/* qfrc_bias = C(q,v) v + gravity term, by projecting each link's Newton-Euler equation (qacc = 0) on its Jacobians */
static void bias_k(const orc_model* m, const orc_params* P, const kin_t* k, const double* qvel, double* c) {
  int nv = m->nv;
  double w[L][3], al[L][3], ao[L][3];
  memset(c, 0, NV * sizeof(double));
  double jp[3 * NV], jr[3 * NV];
  for (int i = 0; i < m->nlink; i++) {
    if (i == 0) {
      matvec3(k->xmat[0], qvel + 3, w[0]);
      for (int x = 0; x < 3; x++) { al[0][x] = 0; ao[0][x] = 0; }
    } else {
      int p = m->parent[i];
      double qd = qvel[5 + i], r[3], t[3], t2[3];
      for (int x = 0; x < 3; x++) w[i][x] = w[p][x] + k->xaxis[i][x] * qd;
      cross3(w[p], k->xaxis[i], t);
      for (int x = 0; x < 3; x++) al[i][x] = al[p][x] + t[x] * qd;
      for (int x = 0; x < 3; x++) r[x] = k->xpos[i][x] - k->xpos[p][x];
      cross3(al[p], r, t);
      cross3(w[p], r, t2);
      cross3(w[p], t2, t2);
      for (int x = 0; x < 3; x++) ao[i][x] = ao[p][x] + t[x] + t2[x];
    }
    double s[3], t[3], t2[3], ac[3], F[3], N[3], Iw_w[3], Iw_al[3];
    for (int x = 0; x < 3; x++) s[x] = k->xcom[i][x] - k->xpos[i][x];
    cross3(al[i], s, t);
    cross3(w[i], s, t2);
    cross3(w[i], t2, t2);
    for (int x = 0; x < 3; x++) ac[x] = ao[i][x] + t[x] + t2[x];
    for (int x = 0; x < 3; x++) F[x] = P->mass[i] * (ac[x] - m->gravity[x]);
    matvec3(k->Iw[i], w[i], Iw_w);
    matvec3(k->Iw[i], al[i], Iw_al);
    cross3(w[i], Iw_w, t);
    for (int x = 0; x < 3; x++) N[x] = Iw_al[x] + t[x];
    jac_point(m, k, i, k->xcom[i], jp, jr);
    for (int d = 0; d < nv; d++)
      for (int x = 0; x < 3; x++) c[d] += jp[x * NV + d] * F[x] + jr[x * NV + d] * N[x];
  }
}

void orc_bias(const orc_model* m, const double* qpos, const double* qvel, double* c) {
  kin_t k; double cc[NV]; orc_params P;
  params_default(m, &P);
  fk(m, &P, qpos, &k);
  bias_k(m, &P, &k, qvel, cc);
  memcpy(c, cc, m->nv * sizeof(double));
}

double orc_energy(const orc_model* m, const double* qpos, const double* qvel, double* kinetic, double* potential) {
  kin_t k; double M[NV * NV]; orc_params P;
  params_default(m, &P);
  fk(m, &P, qpos, &k);
  mass_matrix_k(m, &P, &k, M);
  double ke = 0, pe = 0;
  for (int a = 0; a < m->nv; a++)
    for (int b = 0; b < m->nv; b++) ke += 0.5 * qvel[a] * M[a * NV + b] * qvel[b];
  for (int i = 0; i < m->nlink; i++) pe -= P.mass[i] * dot3(m->gravity, k.xcom[i]);
  if (kinetic) *kinetic = ke;
  if (potential) *potential = pe;
  return ke + pe;
}

/* ------------------------------------------------------------------ constraints */
typedef struct {
  int nrow, ncon;
  double J[ORC_MAXROW][NV];
  double pos[ORC_MAXROW], D[ORC_MAXROW], aref[ORC_MAXROW], R[ORC_MAXROW];
  int type[ORC_MAXROW];      /* 0: unilateral (limit / pyramid edge), 1: dof friction loss */
  double floss[ORC_MAXROW];
  int con_row[ORC_MAXCON], con_geom[ORC_MAXCON], con_slab[ORC_MAXCON];   /* con_slab: 1 = against a stepping stone, 0 = floor plane */
  double con_pos[ORC_MAXCON][3], con_dist[ORC_MAXCON];
  double con_nrm[ORC_MAXCON][2];   /* horizontal outward normal of a slab SIDE face (riser contact); (0, 0): normal +z (floor / top face) */
} efc_t;

/* MuJoCo getimpedance(): power-law sigmoid between solimp[0] and solimp[1] over |pos|/width */
static double impedance(const double* solimp, double pos) {
  double d0 = solimp[0], dw = solimp[1], width = solimp[2], mid = solimp[3], power = solimp[4];
  if (d0 < 0.0001) d0 = 0.0001; if (d0 > 0.9999) d0 = 0.9999;
  if (dw < 0.0001) dw = 0.0001; if (dw > 0.9999) dw = 0.9999;
  if (width < MINVAL) return 0.5 * (d0 + dw);
  double x = fabs(pos) / width;
  if (x >= 1) return dw;
  if (x <= 0) return d0;
  double y;
  if (power < 1.0000001) y = x;
  else if (x <= mid) y = pow(x / mid, power) * mid;           /* a*x^p with a = 1/mid^(p-1) */
  else y = 1 - pow((1 - x) / (1 - mid), power) * (1 - mid);   /* 1 - b*(1-x)^p */
  return d0 + y * (dw - d0);
}

/* constraint force and curvature of one row at residual x = J a - aref (MuJoCo PrimalUpdateConstraint):
 * unilateral: f = -D x for x < 0 else 0 ; friction loss: f = clamp(-D x, -floss, +floss) (quadratic zone |x| < floss/D) */
static double row_force(const efc_t* e, int r, double x, double* curv) {
  if (e->type[r] == 0) {
    *curv = x < 0 ? e->D[r] : 0.0;
    return x < 0 ? -e->D[r] * x : 0.0;
  }
  double lim = e->floss[r] * e->R[r];
  if (x <= -lim) { *curv = 0; return e->floss[r]; }
  if (x >= lim) { *curv = 0; return -e->floss[r]; }
  *curv = e->D[r];
  return -e->D[r] * x;
}

/* ---- stepping stones (tasks/stepping_task.py:318-334): slab k is a box of half sizes slab_half, yawed by seq[k][3], whose
 * TOP face passes through seq[k][0:3].  MuJoCo collides the foot boxes with them through mjc_BoxBox (not available here,
 * "parity unpinned"); this restates the face-face case of that routine with the slab's top face as the reference face:
 * the contact points are the vertices of the foot's sole rectangle clipped against the slab footprint, i.e. (A) sole
 * corners inside the footprint and (B) sole-edge x footprint-boundary crossings, normal +z, dist = z - top.  A point is
 * supported by the top face only while it is inside the slab's thickness and not deeper than its inset from the slab's
 * side faces (beyond that the minimum-penetration axis of a box-box test is a side face, which is not modelled). */
static int slab_supports(const orc_model* m, const double* sl, const double* p, double* dist) {
  double c = cos(sl[3]), sn = sin(sl[3]);
  double dx = p[0] - sl[0], dy = p[1] - sl[1];
  double xs = c * dx + sn * dy, ys = -sn * dx + c * dy;
  double ix = m->slab_half[0] - fabs(xs), iy = m->slab_half[1] - fabs(ys);
  if (ix < 0 || iy < 0) return 0;
  double d = p[2] - sl[2], inset = ix < iy ? ix : iy;
  if (!(d < 0) || -d >= 2 * m->slab_half[2]) return 0;
  if (-d > m->side_tol && -d > inset) return 0;
  *dist = d;
  return 1;
}

/* Side faces (stair risers).  A point INSIDE a slab's volume that its top face does not support (deeper than side_tol AND
 * deeper than its inset from the side faces: slab_supports == 0) is in contact with the nearest side face: the face of
 * minimum penetration of a box-box test.  Returns 1 and the outward horizontal normal of that face (world), the signed
 * distance (-inset) and the contact point half way between the point and the face.  x' faces win ties. */
static int slab_side(const orc_model* m, const double* sl, const double* p, double nrm[2], double* dist, double pos[3]) {
  double c = cos(sl[3]), sn = sin(sl[3]);
  double dx = p[0] - sl[0], dy = p[1] - sl[1];
  double xs = c * dx + sn * dy, ys = -sn * dx + c * dy;
  double ix = m->slab_half[0] - fabs(xs), iy = m->slab_half[1] - fabs(ys);
  if (ix < 0 || iy < 0) return 0;
  double d = p[2] - sl[2], inset = ix < iy ? ix : iy;
  if (!(d < 0) || -d >= 2 * m->slab_half[2]) return 0;
  if (!(-d > m->side_tol && -d > inset)) return 0;          /* the top face carries it (slab_supports) */
  double lx = 0, ly = 0;                                   /* outward normal in the slab frame */
  if (ix <= iy) lx = xs < 0 ? -1.0 : 1.0; else ly = ys < 0 ? -1.0 : 1.0;
  nrm[0] = c * lx - sn * ly;
  nrm[1] = sn * lx + c * ly;
  *dist = -inset;
  pos[0] = p[0] + 0.5 * inset * nrm[0];
  pos[1] = p[1] + 0.5 * inset * nrm[1];
  pos[2] = p[2];
  return 1;
}

static void make_constraints(const orc_model* m, const orc_params* P, const kin_t* k, const double* qpos, const double* qvel,
                             const orc_env* env, efc_t* e) {
  int nv = m->nv;
  e->nrow = 0; e->ncon = 0;
  double tau = m->solref[0], zeta = m->solref[1];
  if (tau < 2 * m->timestep) tau = 2 * m->timestep; /* refsafe */
  double dmax = m->solimp[1];
  double K = 1.0 / fmax(MINVAL, dmax * dmax * tau * tau * zeta * zeta);
  double B = 2.0 / fmax(MINVAL, dmax * tau);
  /* dof friction loss (mj_makeConstraint puts these first): J = e_d, pos = 0 -> imp = solimp[0], aref = -B vel */
  for (int d = 6; d < nv; d++) {
    if (!(P->frictionloss[d] > 0)) continue;
    int r = e->nrow++;
    memset(e->J[r], 0, sizeof(e->J[r]));
    e->J[r][d] = 1.0;
    double imp = impedance(m->solimp, 0.0);
    e->pos[r] = 0;
    e->type[r] = 1;
    e->floss[r] = P->frictionloss[d];
    e->R[r] = fmax(MINVAL, (1 - imp) / imp * m->dof_invweight0[d]);
    e->D[r] = 1.0 / e->R[r];
    e->aref[r] = -B * qvel[d];
  }
  /* joint limits (rows precede contacts, as in mj_makeConstraint) */
  for (int d = 6; d < nv; d++) {
    if (!m->limited[d]) continue;
    double q = qpos[d + 1];
    for (int side = 0; side < 2; side++) {
      double dist = side == 0 ? q - m->range[d][0] : m->range[d][1] - q;
      if (dist < 0) {
        int r = e->nrow++;
        memset(e->J[r], 0, sizeof(e->J[r]));
        e->J[r][d] = side == 0 ? 1.0 : -1.0;
        double imp = impedance(m->solimp, dist);
        double vel = e->J[r][d] * qvel[d];
        e->pos[r] = dist;
        e->type[r] = 0; e->floss[r] = 0;
        e->R[r] = fmax(MINVAL, (1 - imp) / imp * m->dof_invweight0[d]);
        e->D[r] = 1.0 / e->R[r];
        e->aref[r] = -B * vel - K * imp * dist;
      }
    }
  }
  /* feet vs ground plane z=0, normal +z.  box: mjc_PlaneBox (at most 4 corners, in corner-index order);
   * spheres: the end spheres of the foot capsules, mjc_PlaneCapsule = 2 x mjc_PlaneSphere (dist = z - r,
   * pos = centre - n (r + dist/2)) */
  for (int g = 0; g < m->ngeom; g++) {
    int lk = m->geom_link[g];
    if (m->geom_type[g] == ORC_GEOM_SPHERES) {
      for (int i = 0; i < m->geom_npts[g]; i++) {
        double c[3];
        matvec3(k->xmat[lk], m->geom_pts[g][i], c);
        for (int x = 0; x < 3; x++) c[x] += k->xpos[lk][x];
        double cd = c[2] - m->geom_radius[g];
        if (!(cd < 0)) continue;
        int ci = e->ncon++;
        e->con_slab[ci] = 0;
        e->con_geom[ci] = g;
        e->con_dist[ci] = cd; e->con_nrm[ci][0] = e->con_nrm[ci][1] = 0;
        e->con_pos[ci][0] = c[0]; e->con_pos[ci][1] = c[1];
        e->con_pos[ci][2] = c[2] - (m->geom_radius[g] + 0.5 * cd);
      }
      continue;
    }
    double off[3], ctr[3];
    matvec3(k->xmat[lk], m->geom_pos[g], off);
    for (int x = 0; x < 3; x++) ctr[x] = k->xpos[lk][x] + off[x];
    if (m->task == ORC_TASK_STEP || m->terrain) {
      /* floor body moved to z = -2 in FORWARD mode (stepping_task.py:332-334) */
      double floor_z = (m->task == ORC_TASK_STEP && env->mode == ORC_STEP_FORWARD) ? -2.0 : 0.0;
      double cw[8][3];
      int cnt = 0;
      for (int i = 0; i < 8; i++) {
        double v[3] = {(i & 1 ? 1 : -1) * m->geom_size[g][0], (i & 2 ? 1 : -1) * m->geom_size[g][1],
                       (i & 4 ? 1 : -1) * m->geom_size[g][2]};
        double corner[3];
        matvec3(k->xmat[lk], v, corner);
        for (int x = 0; x < 3; x++) cw[i][x] = corner[x] + ctr[x];
        if (cnt >= 4 || corner[2] > 0) continue;       /* mjc_PlaneBox keeps corners on the plane side of the centre, 4 at most */
        /* (A) the highest supporting surface under this corner; every surface at that height is its own contact */
        double best = 0, dsl[ORC_NSLAB];
        int have = 0, sup[ORC_NSLAB];
        if (cw[i][2] - floor_z < 0) { best = floor_z; have = 1; }
        for (int sidx = 0; sidx < ORC_NSLAB; sidx++) {
          sup[sidx] = slab_supports(m, env->seq[sidx], cw[i], &dsl[sidx]);
          if (sup[sidx] && (!have || env->seq[sidx][2] > best)) { best = env->seq[sidx][2]; have = 1; }
        }
        if (!have) continue;
        cnt++;
        for (int sidx = -1; sidx < ORC_NSLAB; sidx++) {
          double h = sidx < 0 ? floor_z : env->seq[sidx][2];
          if (sidx < 0 ? !(cw[i][2] - floor_z < 0) : !sup[sidx]) continue;
          if (h != best) continue;
          if (e->ncon >= ORC_MAXCON) { ((orc_env*)env)->con_overflow++; continue; }
          double cd = cw[i][2] - h;
          int ci = e->ncon++;
          e->con_slab[ci] = sidx >= 0;
          e->con_geom[ci] = g;
          e->con_dist[ci] = cd; e->con_nrm[ci][0] = e->con_nrm[ci][1] = 0;
          e->con_pos[ci][0] = cw[i][0]; e->con_pos[ci][1] = cw[i][1];
          e->con_pos[ci][2] = cw[i][2] - 0.5 * cd;
        }
      }
      /* (B) crossings of the sole edges (corner loop 0-1-3-2 of the local -z face) with each slab's footprint boundary:
       * Liang-Barsky clip of the edge against the rectangle; an entry / exit parameter strictly inside (0,1) is a new
       * polygon vertex.  At most ORC_MAXCROSS per foot, in (slab, edge, entry-then-exit) order. */
      static const int ea[4] = {0, 1, 3, 2}, eb[4] = {1, 3, 2, 0};
      int ncross = 0;
      /* (A') riser contacts: the corners on the plane side of the box centre that sit inside a slab's volume unsupported by
       * its top face push against the nearest side face (first such slab in index order; tasks/stepping_task.py:318-334 poses
       * real boxes, mjc_BoxBox gives riser contacts).  They share the ORC_MAXCROSS extra slots of the foot with the crossings
       * below and come first, in corner-index order. */
      if (m->side_faces) {
        for (int i = 0; i < 8; i++) {
          double v2 = (i & 4 ? 1 : -1) * m->geom_size[g][2];
          double cz = k->xmat[lk][6] * ((i & 1 ? 1 : -1) * m->geom_size[g][0]) + k->xmat[lk][7] * ((i & 2 ? 1 : -1) * m->geom_size[g][1]) +
                      k->xmat[lk][8] * v2;
          if (cz > 0) continue;
          for (int sidx = 0; sidx < ORC_NSLAB; sidx++) {
            double nrm[2], sd, sp[3];
            if (!slab_side(m, env->seq[sidx], cw[i], nrm, &sd, sp)) continue;
            if (ncross < ORC_MAXCROSS && e->ncon < ORC_MAXCON) {
              ncross++;
              int ci = e->ncon++;
              e->con_slab[ci] = 1;
              e->con_geom[ci] = g;
              e->con_dist[ci] = sd;
              e->con_nrm[ci][0] = nrm[0]; e->con_nrm[ci][1] = nrm[1];
              for (int x = 0; x < 3; x++) e->con_pos[ci][x] = sp[x];
            }
            break;
          }
        }
      }
      for (int sidx = 0; sidx < ORC_NSLAB; sidx++) {
        const double* sl = env->seq[sidx];
        double c = cos(sl[3]), sn = sin(sl[3]);
        for (int ed = 0; ed < 4; ed++) {
          const double* A = cw[ea[ed]];
          const double* Bp = cw[eb[ed]];
          double ax = c * (A[0] - sl[0]) + sn * (A[1] - sl[1]), ay = -sn * (A[0] - sl[0]) + c * (A[1] - sl[1]);
          double bx = c * (Bp[0] - sl[0]) + sn * (Bp[1] - sl[1]), by = -sn * (Bp[0] - sl[0]) + c * (Bp[1] - sl[1]);
          double dx = bx - ax, dy = by - ay, t0 = 0, t1 = 1;
          double pp[4] = {-dx, dx, -dy, dy};
          double qq[4] = {ax + m->slab_half[0], m->slab_half[0] - ax, ay + m->slab_half[1], m->slab_half[1] - ay};
          int ok = 1;
          for (int b = 0; b < 4 && ok; b++) {
            if (pp[b] == 0) { if (qq[b] < 0) ok = 0; continue; }
            double r = qq[b] / pp[b];
            if (pp[b] < 0) { if (r > t1) ok = 0; else if (r > t0) t0 = r; }
            else { if (r < t0) ok = 0; else if (r < t1) t1 = r; }
          }
          if (!ok) continue;
          for (int side = 0; side < 2; side++) {
            double t = side == 0 ? t0 : t1;
            if (side == 0 ? !(t0 > 0) : !(t1 < 1)) continue;
            double z = A[2] + t * (Bp[2] - A[2]);
            double cd = z - sl[2];
            if (!(cd < 0) || -cd > m->side_tol) continue;
            if (ncross >= ORC_MAXCROSS) { continue; }
            if (e->ncon >= ORC_MAXCON) { ((orc_env*)env)->con_overflow++; continue; }
            ncross++;
            int ci = e->ncon++;
            e->con_slab[ci] = 1;
            e->con_geom[ci] = g;
            e->con_dist[ci] = cd; e->con_nrm[ci][0] = e->con_nrm[ci][1] = 0;
            e->con_pos[ci][0] = A[0] + t * (Bp[0] - A[0]);
            e->con_pos[ci][1] = A[1] + t * (Bp[1] - A[1]);
            e->con_pos[ci][2] = z - 0.5 * cd;
          }
        }
      }
      continue;
    }
    double dist0 = ctr[2];
    int cnt = 0;
    for (int i = 0; i < 8 && cnt < 4; i++) {
      double v[3] = {(i & 1 ? 1 : -1) * m->geom_size[g][0], (i & 2 ? 1 : -1) * m->geom_size[g][1],
                     (i & 4 ? 1 : -1) * m->geom_size[g][2]};
      double corner[3];
      matvec3(k->xmat[lk], v, corner);
      double ldist = corner[2];
      if (dist0 + ldist > 0 || ldist > 0) continue;
      double cd = dist0 + ldist;
      int ci = e->ncon++;
      e->con_slab[ci] = 0;
      e->con_geom[ci] = g;
      e->con_dist[ci] = cd; e->con_nrm[ci][0] = e->con_nrm[ci][1] = 0;
      e->con_pos[ci][0] = corner[0] + ctr[0];
      e->con_pos[ci][1] = corner[1] + ctr[1];
      e->con_pos[ci][2] = corner[2] + ctr[2] - 0.5 * cd;
      cnt++;
    }
  }
  /* pyramidal rows: frame (n,t1,t2) = (+z, +y, -x) from mju_makeFrame on n = (0,0,1) */
  double mu = m->mu * sqrt(1.0 / fmax(MINVAL, m->impratio));
  if (m->contact_solref[0] > 0) {   /* compliant-terrain extension: the contact pairs carry their own solref */
    double tc = m->contact_solref[0] < 2 * m->timestep ? 2 * m->timestep : m->contact_solref[0], zc = m->contact_solref[1];
    K = 1.0 / fmax(MINVAL, dmax * dmax * tc * tc * zc * zc);
    B = 2.0 / fmax(MINVAL, dmax * tc);
  }
  for (int ci = 0; ci < e->ncon; ci++) {
    int lk = m->geom_link[e->con_geom[ci]];
    double jp[3 * NV];
    jac_point(m, k, lk, e->con_pos[ci], jp, 0);
    double imp = impedance(m->solimp, e->con_dist[ci]);
    double tran = m->link_invweight0[lk][0]; /* world body contributes 0 */
    double diagApprox = tran + m->mu * m->mu * tran;
    double Rn = fmax(MINVAL, (1 - imp) / imp * diagApprox);
    double Rpy = 2 * mu * mu * Rn;
    e->con_row[ci] = e->nrow;
    for (int j = 0; j < 4; j++) {
      int r = e->nrow++;
      double vel = 0;
      const double nx = e->con_nrm[ci][0], ny = e->con_nrm[ci][1];
      const int side = nx != 0 || ny != 0;
      for (int d = 0; d < nv; d++) {
        double jn, jt;
        if (!side) {   /* frame (n, t1, t2) = (+z, +y, -x) */
          jn = jp[2 * NV + d]; jt = (j < 2) ? jp[1 * NV + d] : -jp[0 * NV + d];
        } else {       /* riser: n = (nx, ny, 0), t1 = +z, t2 = n x t1 = (ny, -nx, 0) */
          jn = nx * jp[0 * NV + d] + ny * jp[1 * NV + d];
          jt = (j < 2) ? jp[2 * NV + d] : ny * jp[0 * NV + d] - nx * jp[1 * NV + d];
        }
        e->J[r][d] = jn + ((j & 1) ? -m->mu : m->mu) * jt;
        vel += e->J[r][d] * qvel[d];
      }
      for (int d = nv; d < NV; d++) e->J[r][d] = 0;
      e->pos[r] = e->con_dist[ci];
      e->type[r] = 0; e->floss[r] = 0;
      e->R[r] = Rpy;
      e->D[r] = 1.0 / Rpy;
      e->aref[r] = -B * vel - K * imp * e->con_dist[ci];
    }
  }
}

/* primal Newton with exact (piecewise-quadratic) line search; qacc in: warmstart, out: solution */
static int solve_newton(const orc_model* m, const double* M, const efc_t* e, const double* qfs, double* qacc,
                        double* force, double* kkt) {
  int nv = m->nv, nr = e->nrow, it;
  double scale = 1.0 / (m->meaninertia * (nv > 1 ? nv : 1));
  double Ma[NV], jar[ORC_MAXROW], curv[ORC_MAXROW], grad[NV], H[NV * NV], s[NV], jv[ORC_MAXROW], Ms[NV];
  double gnorm = 0;
  for (it = 0; it <= m->iterations; it++) {
    for (int a = 0; a < nv; a++) {
      double t = 0;
      for (int b = 0; b < nv; b++) t += M[a * NV + b] * qacc[b];
      Ma[a] = t;
    }
    for (int r = 0; r < nr; r++) {
      double t = -e->aref[r];
      for (int d = 0; d < nv; d++) t += e->J[r][d] * qacc[d];
      jar[r] = t;
      force[r] = row_force(e, r, t, &curv[r]);
    }
    gnorm = 0;
    for (int a = 0; a < nv; a++) {
      double t = Ma[a] - qfs[a];
      for (int r = 0; r < nr; r++) t -= e->J[r][a] * force[r];
      grad[a] = t;
      gnorm += t * t;
    }
    gnorm = sqrt(gnorm);
    if (gnorm * scale < m->tolerance || it == m->iterations) break;
    for (int a = 0; a < nv; a++)
      for (int b = 0; b < nv; b++) {
        double t = M[a * NV + b];
        for (int r = 0; r < nr; r++)
          if (curv[r] > 0) t += curv[r] * e->J[r][a] * e->J[r][b];
        H[a * nv + b] = t;
      }
    if (chol(H, nv)) break;
    for (int a = 0; a < nv; a++) s[a] = -grad[a];
    chol_solve(H, nv, s);
    /* line search on phi(alpha) = cost(qacc + alpha s): phi' is continuous, piecewise linear, increasing */
    double sMs = 0, sg = 0;
    for (int a = 0; a < nv; a++) {
      double t = 0;
      for (int b = 0; b < nv; b++) t += M[a * NV + b] * s[b];
      Ms[a] = t;
      sMs += s[a] * t;
      sg += s[a] * (Ma[a] - qfs[a]);
    }
    for (int r = 0; r < nr; r++) {
      double t = 0;
      for (int d = 0; d < nv; d++) t += e->J[r][d] * s[d];
      jv[r] = t;
    }
    double bp[2 * ORC_MAXROW];
    int nbp = 0;
    for (int r = 0; r < nr; r++)
      if (jv[r] != 0) {
        if (e->type[r] == 0) {
          double a0 = -jar[r] / jv[r];
          if (a0 > 0) bp[nbp++] = a0;
        } else {
          double lim = e->floss[r] * e->R[r];
          double a0 = (-lim - jar[r]) / jv[r], a1 = (lim - jar[r]) / jv[r];
          if (a0 > 0) bp[nbp++] = a0;
          if (a1 > 0) bp[nbp++] = a1;
        }
      }
    for (int i = 1; i < nbp; i++) { /* insertion sort */
      double v = bp[i]; int j = i - 1;
      while (j >= 0 && bp[j] > v) { bp[j + 1] = bp[j]; j--; }
      bp[j + 1] = v;
    }
#define DPHI(al, out)                                                       \
  do {                                                                      \
    double _d = (al) * sMs + sg;                                            \
    for (int r = 0; r < nr; r++) {                                          \
      double _cv, _f = row_force(e, r, jar[r] + (al) * jv[r], &_cv);        \
      _d -= _f * jv[r];                                                     \
    }                                                                       \
    (out) = _d;                                                             \
  } while (0)
    double lo = 0, dlo, alpha = -1;
    DPHI(0.0, dlo);
    if (!(dlo < 0)) break; /* not a descent direction: converged to roundoff */
    for (int i = 0; i < nbp; i++) {
      double dhi;
      DPHI(bp[i], dhi);
      if (dhi >= 0) { alpha = lo - dlo * (bp[i] - lo) / (dhi - dlo); break; }
      lo = bp[i]; dlo = dhi;
    }
    if (alpha < 0) { /* beyond the last breakpoint: slope is constant */
      double slope = sMs, mid = lo + 1.0;
      for (int r = 0; r < nr; r++) {
        double cv;
        row_force(e, r, jar[r] + mid * jv[r], &cv);
        slope += cv * jv[r] * jv[r];
      }
      alpha = lo - dlo / slope;
    }
#undef DPHI
    for (int a = 0; a < nv; a++) qacc[a] += alpha * s[a];
  }
  if (kkt) *kkt = gnorm;
  return it;
}

/* dual projected Gauss-Seidel on  min_{f>=0} 1/2 f^T (J M^-1 J^T + R) f + f^T (J a_s - aref)  (tests only) */
static int solve_pgs(const orc_model* m, const double* M, const efc_t* e, const double* qfs, double* qacc,
                     double* force, double* kkt) {
  int nv = m->nv, nr = e->nrow;
  double G[NV * NV], as[NV], MinvJt[ORC_MAXROW][NV];
  static __thread double A[ORC_PGS_MAXROW][ORC_PGS_MAXROW];
  double b[ORC_MAXROW];
  for (int a = 0; a < nv; a++)
    for (int c = 0; c < nv; c++) G[a * nv + c] = M[a * NV + c];
  chol(G, nv);
  memcpy(as, qfs, nv * sizeof(double));
  chol_solve(G, nv, as);
  for (int r = 0; r < nr; r++) {
    memcpy(MinvJt[r], e->J[r], nv * sizeof(double));
    chol_solve(G, nv, MinvJt[r]);
  }
  for (int r = 0; r < nr; r++) {
    for (int c = 0; c < nr; c++) {
      double t = 0;
      for (int d = 0; d < nv; d++) t += e->J[r][d] * MinvJt[c][d];
      A[r][c] = t + (r == c ? e->R[r] : 0);
    }
    double t = -e->aref[r];
    for (int d = 0; d < nv; d++) t += e->J[r][d] * as[d];
    b[r] = t;
    force[r] = 0;
  }
  int it, maxit = m->iterations * 2000;
  for (it = 0; it < maxit; it++) {
    double change = 0;
    for (int r = 0; r < nr; r++) {
      double res = b[r];
      for (int c = 0; c < nr; c++) res += A[r][c] * force[c];
      double fn = force[r] - res / A[r][r];
      if (e->type[r] == 0) { if (fn < 0) fn = 0; }
      else { if (fn > e->floss[r]) fn = e->floss[r]; if (fn < -e->floss[r]) fn = -e->floss[r]; }
      change += fabs(fn - force[r]);
      force[r] = fn;
    }
    if (change < 1e-14) break;
  }
  for (int a = 0; a < nv; a++) {
    double t = as[a];
    for (int r = 0; r < nr; r++) t += MinvJt[r][a] * force[r];
    qacc[a] = t;
  }
  if (kkt) {
    double g2 = 0;
    for (int a = 0; a < nv; a++) {
      double t = -qfs[a];
      for (int c = 0; c < nv; c++) t += M[a * NV + c] * qacc[c];
      for (int r = 0; r < nr; r++) {
        double jr_ = -e->aref[r];
        for (int d = 0; d < nv; d++) jr_ += e->J[r][d] * qacc[d];
        double cv_;
        t -= e->J[r][a] * row_force(e, r, jr_, &cv_);
      }
      g2 += t * t;
    }
    *kkt = sqrt(g2);
  }
  return it;
}

/* squared distance between segments p1-q1 and p2-q2 (closest points by clamping, textbook formulation) */
static double seg_seg_dist2(const double* p1, const double* q1, const double* p2, const double* q2) {
  double d1[3], d2[3], r[3];
  for (int x = 0; x < 3; x++) { d1[x] = q1[x] - p1[x]; d2[x] = q2[x] - p2[x]; r[x] = p1[x] - p2[x]; }
  double a = dot3(d1, d1), e = dot3(d2, d2), f = dot3(d2, r), s, t;
  const double EPS = 1e-12;
  if (a <= EPS && e <= EPS) { s = t = 0; }
  else if (a <= EPS) { s = 0; t = f / e; t = t < 0 ? 0 : (t > 1 ? 1 : t); }
  else {
    double c = dot3(d1, r);
    if (e <= EPS) { t = 0; s = -c / a; s = s < 0 ? 0 : (s > 1 ? 1 : s); }
    else {
      double b = dot3(d1, d2), den = a * e - b * b;
      s = den > EPS ? (b * f - c * e) / den : 0.0;
      s = s < 0 ? 0 : (s > 1 ? 1 : s);
      t = (b * s + f) / e;
      if (t < 0) { t = 0; s = -c / a; s = s < 0 ? 0 : (s > 1 ? 1 : s); }
      else if (t > 1) { t = 1; s = (b - c) / a; s = s < 0 ? 0 : (s > 1 ? 1 : s); }
    }
  }
  double dd = 0;
  for (int x = 0; x < 3; x++) { double w_ = r[x] + d1[x] * s - d2[x] * t; dd += w_ * w_; }
  return dd;
}

static int self_collision(const orc_model* m, const kin_t* k) {
  double e0[ORC_MAXCAP][3], e1[ORC_MAXCAP][3];
  for (int c = 0; c < m->ncap; c++) {
    int lk = m->cap_link[c];
    double t[3];
    matvec3(k->xmat[lk], m->cap_p0[c], t);
    for (int x = 0; x < 3; x++) e0[c][x] = k->xpos[lk][x] + t[x];
    matvec3(k->xmat[lk], m->cap_p1[c], t);
    for (int x = 0; x < 3; x++) e1[c][x] = k->xpos[lk][x] + t[x];
  }
  for (int p = 0; p < m->npair; p++) {
    int a = m->pair[p][0], b = m->pair[p][1];
    double rr = m->cap_r[a] + m->cap_r[b];
    if (seg_seg_dist2(e0[a], e1[a], e0[b], e1[b]) < rr * rr) return 1;
  }
  return 0;
}

/* ------------------------------------------------------------------ mj_step */
void orc_mj_step(const orc_model* m, orc_env* e, const double* ctrl) {
  int nv = m->nv, nu = m->nu;
  double h = m->timestep;
  kin_t k;
  efc_t efc;
  double M[NV * NV], bias[NV], qfs[NV], force[ORC_MAXROW], qacc[NV];
  const orc_params* P = &e->P;
  fk(m, P, e->qpos, &k);
  mass_matrix_k(m, P, &k, M);
  make_constraints(m, P, &k, e->qpos, e->qvel, e, &efc);
  bias_k(m, P, &k, e->qvel, bias);
  for (int d = 0; d < nv; d++) qfs[d] = -P->damping[d] * e->qvel[d] - bias[d];
  for (int u = 0; u < nu; u++) qfs[6 + u] += ctrl[u]; /* motors, gear 1 */
  /* xfrc_applied on the pelvis / torso bodies (both welded into the root link): world force f and torque tau at the
   * body CoM -> generalised force J_com' [f; tau]; only the root dofs see it */
  for (int b = 0; b < 2; b++) {
    const double* f = e->xfrc[b];
    if (f[0] == 0 && f[1] == 0 && f[2] == 0 && f[3] == 0 && f[4] == 0 && f[5] == 0) continue;
    double rloc[3], r[3], rxf[3];
    memcpy(rloc, b == 0 ? P->pel_com : m->torso_com, sizeof(rloc));
    matvec3(k.xmat[0], rloc, r);
    cross3(r, f, rxf);
    for (int x = 0; x < 3; x++) qfs[x] += f[x];
    for (int kk = 0; kk < 3; kk++) {
      double ax[3] = {k.xmat[0][kk], k.xmat[0][3 + kk], k.xmat[0][6 + kk]};
      qfs[3 + kk] += ax[0] * (f[3] + rxf[0]) + ax[1] * (f[4] + rxf[1]) + ax[2] * (f[5] + rxf[2]);
    }
  }
  memcpy(qacc, e->qacc_warm, sizeof(qacc));
  double kkt = 0;
  if (efc.nrow == 0) {
    double G[NV * NV];
    for (int a = 0; a < nv; a++)
      for (int c = 0; c < nv; c++) G[a * nv + c] = M[a * NV + c];
    chol(G, nv);
    memcpy(qacc, qfs, nv * sizeof(double));
    chol_solve(G, nv, qacc);
    e->last_solver_iter = 0;
  } else if (m->solver == ORC_SOLVER_PGS && efc.nrow <= ORC_PGS_MAXROW) {
    e->last_solver_iter = solve_pgs(m, M, &efc, qfs, qacc, force, &kkt);
  } else {
    e->last_solver_iter = solve_newton(m, M, &efc, qfs, qacc, force, &kkt);
  }
  e->last_kkt_residual = kkt;
  e->iter_trace[e->nsubsteps & 31] = e->last_solver_iter;
  e->nrow_trace[e->nsubsteps & 31] = efc.nrow;
  /* ---- quantities mjData keeps after mj_step (all refer to the pre-integration state) */
  memcpy(e->qacc, qacc, nv * sizeof(double));
  for (int u = 0; u < nu; u++) {
    e->act_len[u] = e->qpos[7 + u];
    e->act_vel[u] = e->qvel[6 + u];
    e->act_force[u] = ctrl[u];
  }
  memcpy(e->root_xpos, k.xpos[0], sizeof(e->root_xpos));
  memcpy(e->root_xmat, k.xmat[0], sizeof(e->root_xmat));
  memcpy(e->root_vlin, e->qvel, sizeof(e->root_vlin));
  {
    double t[3];
    matvec3(k.xmat[0], m->head_in_root, t);
    for (int x = 0; x < 3; x++) e->head_xpos[x] = k.xpos[0][x] + t[x];
  }
  for (int f = 0; f < 2; f++) {
    int lk = f == 0 ? m->rfoot_link : m->lfoot_link;
    double jp[3 * NV], v[3] = {0, 0, 0};
    jac_point(m, &k, lk, k.xpos[lk], jp, 0);
    for (int x = 0; x < 3; x++)
      for (int d = 0; d < nv; d++) v[x] += jp[x * NV + d] * e->qvel[d];
    memcpy(f == 0 ? e->rfoot_vel : e->lfoot_vel, v, sizeof(v));
  }
  for (int f = 0; f < 2; f++) {
    int lk = f == 0 ? m->rfoot_link : m->lfoot_link;
    double t[3];
    matvec3(k.xmat[lk], m->foot_site[f], t);
    for (int x = 0; x < 3; x++) { e->site_pos[f][x] = k.xpos[lk][x] + t[x]; e->foot_xpos[f][x] = k.xpos[lk][x]; }
  }
  {
    const double* q = e->qpos + 3;   /* mj_kinematics normalises the free-joint quaternion */
    double nq = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int i = 0; i < 4; i++) e->root_quat[i] = q[i] / nq;
  }
  e->rfoot_grf = e->lfoot_grf = 0;
  e->ncon_r = e->ncon_l = 0;
  e->ncon = efc.ncon;
  e->contact_z_min = 0;
  e->self_collision = m->npair > 0 ? self_collision(m, &k) : 0;
  int first = 1;
  for (int ci = 0; ci < efc.ncon; ci++) {
    const double* f = force + efc.con_row[ci];
    double fn = f[0] + f[1] + f[2] + f[3], f1 = m->mu * (f[0] - f[1]), f2 = m->mu * (f[2] - f[3]);
    double nrm = sqrt(fn * fn + f1 * f1 + f2 * f2); /* robot_interface.py:310-312 norm of mj_contactForce */
    int lk = m->geom_link[efc.con_geom[ci]];
    /* RobotInterface.get_*_floor_contacts (envs/common/robot_interface.py:252-301) keeps a contact only if
     * its geom1 is a non-robot geom and its geom2 the foot.  MuJoCo orders the geoms of a contact by type, then by id: plane <
     * box, so the floor plane is geom1 of a foot contact, but a foot box (low id, robot tree first) precedes a stepping-stone
     * box, so foot-on-stone contacts have the FOOT as geom1 and are skipped: in jvrc_step they carry load but are invisible
     * to the task's GRF, contact_point_z and floor-collision checks.  slab_contacts_are_floor = 1 switches to the
     * physically meant behaviour. */
    if (m->task == ORC_TASK_STEP && efc.con_slab[ci] && !m->slab_contacts_are_floor) continue;
    if (lk == m->rfoot_link) { e->rfoot_grf += nrm; e->ncon_r++; }
    else if (lk == m->lfoot_link) { e->lfoot_grf += nrm; e->ncon_l++; }
    if (first || efc.con_pos[ci][2] < e->contact_z_min) e->contact_z_min = efc.con_pos[ci][2];
    first = 0;
  }
  /* ---- mj_Euler with implicit joint damping: (M + h B) qacc' = qfrc_smooth + qfrc_constraint */
  double rhs[NV], G[NV * NV];
  int any_damp = 0;
  for (int d = 0; d < nv; d++) {
    double t = qfs[d];
    for (int r = 0; r < efc.nrow; r++) t += efc.J[r][d] * force[r];
    rhs[d] = t;
    if (P->damping[d] > 0) any_damp = 1;
  }
  if (efc.nrow == 0) memset(force, 0, sizeof(force));
  if (any_damp && !m->explicit_euler) {
    for (int a = 0; a < nv; a++)
      for (int c = 0; c < nv; c++) G[a * nv + c] = M[a * NV + c] + (a == c ? h * P->damping[a] : 0);
    chol(G, nv);
    chol_solve(G, nv, rhs);
  } else {
    memcpy(rhs, qacc, nv * sizeof(double));
  }
  int bad = 0;
  for (int d = 0; d < nv; d++) {
    if (!(fabs(rhs[d]) < 1e10)) bad = 1;
    e->qvel[d] += h * rhs[d];
  }
  for (int x = 0; x < 3; x++) e->qpos[x] += h * e->qvel[x];
  {
    double* q = e->qpos + 3;
    double nq = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int i = 0; i < 4; i++) q[i] /= nq;
    const double* w = e->qvel + 3;
    double wn = sqrt(dot3(w, w));
    if (wn > MINVAL) {
      double ang = h * wn, sa = sin(0.5 * ang) / wn, ca = cos(0.5 * ang);
      double r[4] = {ca, w[0] * sa, w[1] * sa, w[2] * sa};
      double o[4] = {q[0] * r[0] - q[1] * r[1] - q[2] * r[2] - q[3] * r[3],
                     q[0] * r[1] + q[1] * r[0] + q[2] * r[3] - q[3] * r[2],
                     q[0] * r[2] - q[1] * r[3] + q[2] * r[0] + q[3] * r[1],
                     q[0] * r[3] + q[1] * r[2] - q[2] * r[1] + q[3] * r[0]};
      double no = sqrt(o[0] * o[0] + o[1] * o[1] + o[2] * o[2] + o[3] * o[3]);
      for (int i = 0; i < 4; i++) q[i] = o[i] / no;
    }
  }
  for (int u = 0; u < nu; u++) e->qpos[7 + u] += h * e->qvel[6 + u];
  memcpy(e->qacc_warm, qacc, nv * sizeof(double));
  if (bad) e->status |= 1;
  e->nsubsteps++;
}

/* ------------------------------------------------------------------ observation / task */
/* transforms3d.euler.quat2euler(q) (axes 'sxyz') roll and pitch, via quat2mat (tasks/observations.py:22) */
void orc_quat2rp(const double* q, double* roll, double* pitch) {
  double Nq = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  double s = 2.0 / Nq;
  double w = q[0], x = q[1], y = q[2], z = q[3];
  double X = x * s, Y = y * s, Z = z * s;
  double wX = w * X, wY = w * Y, wZ = w * Z, xX = x * X, xY = x * Y, xZ = x * Z, yY = y * Y, yZ = y * Z;
  double M00 = 1 - (yY + z * Z), M10 = xY + wZ, M20 = xZ - wY, M21 = yZ + wX, M22 = 1 - (xX + yY);
  double cy = sqrt(M00 * M00 + M10 * M10);
  if (cy > 4 * 2.220446049250313e-16) {
    *roll = atan2(M21, M22);
    *pitch = atan2(-M20, cy);
  } else {
    double M12 = yZ - wX, M11 = 1 - (xX + z * Z);
    *roll = atan2(-M12, M11);
    *pitch = atan2(-M20, cy);
  }
}

static void get_obs(const orc_model* m, const orc_env* e, double* obs) {
  int nu = m->nu, o = 0;
  double r, p;
  orc_quat2rp(e->qpos + 3, &r, &p);
  obs[o++] = r; obs[o++] = p;
  for (int x = 0; x < 3; x++) obs[o++] = e->qvel[3 + x];
  for (int u = 0; u < nu; u++) obs[o++] = e->act_len[u];
  for (int u = 0; u < nu; u++) obs[o++] = e->act_vel[u];
  if (m->task == ORC_TASK_STAND) {
    /* H1BaseEnv._get_robot_state (envs/h1/h1_base.py:91-117): + motor torques, then uniform observation noise
     * (base_humanoid_env.py:311-338) drawn as value i -> philox stream 40 + i/4, lane i%4 */
    for (int u = 0; u < nu; u++) obs[o++] = e->act_force[u];
    for (int i = 0; i < o; i++) {
      double sc = i < 2 ? m->obs_noise[0] : i < 5 ? m->obs_noise[1] : i < 5 + nu ? m->obs_noise[2]
                  : i < 5 + 2 * nu ? m->obs_noise[3] : m->obs_noise[4];
      if (sc > 0) {
        uint32_t w[4];
        orc_philox(e->seed, e->env_id, e->rng_ctr, 40 + (i >> 2), w);
        obs[i] += -sc + 2 * sc * u01(w[i & 3]);
      }
    }
    return;
  }
  obs[o++] = sin(2 * M_PI * e->phase / m->period);
  obs[o++] = cos(2 * M_PI * e->phase / m->period);
  if (m->task == ORC_TASK_STEP) {   /* envs/jvrc/jvrc_step.py:67-76 */
    for (int i = 0; i < 8; i++) obs[o++] = e->goal_steps[i];
    return;
  }
  /* WalkModes.encode: STANDING [0,0,1], INPLACE [0,1,0], FORWARD [1,0,0] */
  obs[o++] = e->mode == ORC_FORWARD; obs[o++] = e->mode == ORC_INPLACE; obs[o++] = e->mode == ORC_STANDING;
  for (int x = 0; x < 3; x++) obs[o++] = e->mode_ref[x];
}

static void sample_ref(orc_env* e, uint32_t stream) {
  uint32_t u[4];
  orc_philox(e->seed, e->env_id, e->rng_ctr, stream, u);
  if (e->mode == ORC_STANDING) {
    for (int x = 0; x < 3; x++) e->mode_ref[x] = -1.0 + 2.0 * u01(u[x]);
  } else if (e->mode == ORC_INPLACE) {
    e->mode_ref[0] = -0.5 + u01(u[0]); e->mode_ref[1] = 0; e->mode_ref[2] = 0;
  } else {
    e->mode_ref[0] = 0; e->mode_ref[1] = 0.4 * u01(u[0]); e->mode_ref[2] = 0;
  }
}

/* ---------------- SteppingTask (tasks/stepping_task.py) */
/* transforms3d mat2euler(R, 'sxyz')[2] */
static double mat_yaw(const double* R) {
  double cy = sqrt(R[0] * R[0] + R[3] * R[3]);
  return cy > 4 * 2.220446049250313e-16 ? atan2(R[3], R[0]) : 0.0;
}

/* update_target_steps (stepping_task.py:207-213) */
static void step_update_targets(orc_env* e) {
  e->t1 = e->t2;
  e->t2 += 1;
  if (e->t2 == e->seq_len) e->t2 = e->seq_len - 1;
}

/* update_goal_steps (stepping_task.py:188-205): targets t1, t2 expressed in the root frame (4x4 inverse of a rigid
 * transform = R' (p - root_pos), R' Rz(theta)); zeros in STANDING mode */
static void step_update_goals(orc_env* e) {
  memset(e->goal_steps, 0, sizeof(e->goal_steps));
  if (e->mode == ORC_STEP_STANDING) return;
  for (int idx = 0; idx < 2; idx++) {
    const double* sq = e->seq[idx == 0 ? e->t1 : e->t2];
    double d[3] = {sq[0] - e->root_xpos[0], sq[1] - e->root_xpos[1], sq[2] - e->root_xpos[2]}, rel[3];
    mattvec3(e->root_xmat, d, rel);
    double c = cos(sq[3]), sn = sin(sq[3]);
    double Rz[9] = {c, -sn, 0, sn, c, 0, 0, 0, 1}, Rt[9], Rr[9];
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) Rt[3 * a + b] = e->root_xmat[3 * b + a];
    matmul3(Rt, Rz, Rr);
    e->goal_steps[0 + idx] = rel[0];
    e->goal_steps[2 + idx] = rel[1];
    e->goal_steps[4 + idx] = rel[2];
    e->goal_steps[6 + idx] = mat_yaw(Rr);
  }
}

/* SteppingTask.reset (stepping_task.py:243-334).  Draws (event counter = the reset's): stream 3 lane 0 mode, lane 1 phase,
 * lane 2 first-step offset / plan index / lateral sign, lane 3 the randint(2,4); stream 4 lane 0 the sign of the step height */
static void task_reset_step(const orc_model* m, orc_env* e) {
  uint32_t u[4], v[4];
  orc_philox(e->seed, e->env_id, e->rng_ctr, 3, u);
  orc_philox(e->seed, e->env_id, e->rng_ctr, 4, v);
  memset(e->goal_steps, 0, sizeof(e->goal_steps));
  e->target_reached = 0; e->target_reached_frames = 0;
  e->t1 = e->t2 = 0;
  e->phase = randint(u[1], 2) == 0 ? 0 : m->period / 2;
  double cm = u01(u[0]);
  e->mode = cm < 0.15 ? ORC_STEP_CURVED : cm < 0.15 + 0.05 ? ORC_STEP_STANDING : cm < 0.15 + 0.05 + 0.2 ? ORC_STEP_BACKWARD
            : cm < 0.15 + 0.05 + 0.2 + 0.3 ? ORC_STEP_LATERAL : ORC_STEP_FORWARD;
  double step_size = 0.3, step_gap = 0.15, step_height = 0;
  int num_steps = 20;
  if (e->mode == ORC_STEP_STANDING) num_steps = 1;
  else if (e->mode == ORC_STEP_BACKWARD) step_size = -0.1;
  else if (e->mode == ORC_STEP_LATERAL) step_size = 0.4;
  else if (e->mode == ORC_STEP_FORWARD) step_height = randint(v[0], 2) == 0 ? -m->step_height : m->step_height;
  double rel[ORC_NSLAB][4];
  int n = 0;
  if (e->mode == ORC_STEP_CURVED) {
    int pi = randint(u[2], m->nplan);
    for (int i = 0; i < m->plan_len[pi]; i++, n++) {
      rel[n][0] = m->plans[pi][i][0]; rel[n][1] = m->plans[pi][i][1]; rel[n][2] = 0; rel[n][3] = m->plans[pi][i][2];
    }
  } else if (e->mode == ORC_STEP_LATERAL) {
    double y = 0, c = randint(u[2], 2) == 0 ? -1 : 1;
    for (int i = 1; i < num_steps; i++, n++) {
      if (i % 2) y += step_size; else y -= (2.0 / 3.0) * step_size;
      rel[n][0] = 0; rel[n][1] = c * y; rel[n][2] = 0; rel[n][3] = 0;
    }
  } else {
    double first = 0.095 + (0.105 - 0.095) * u01(u[2]), y;
    if (e->phase == m->period / 2 && 2 * (m->period / 2) == m->period) { rel[0][1] = -1 * first; y = -step_gap; }
    else { rel[0][1] = 1 * first; y = step_gap; }
    rel[0][0] = 0; rel[0][2] = 0; rel[0][3] = 0;
    n = 1;
    double x = 0, z = 0;
    int c = 2 + randint(u[3], 2);
    for (int i = 1; i < num_steps - 1; i++, n++) {
      x += step_size;
      y *= -1;
      if (i > c) z += step_height;
      rel[n][0] = x; rel[n][1] = y; rel[n][2] = z; rel[n][3] = 0;
    }
    rel[n][0] = x + step_size; rel[n][1] = -y; rel[n][2] = z; rel[n][3] = 0;
    n++;
  }
  /* transform_sequence (stepping_task.py:123-136): about the mid-point of the foot bodies, yawed with the root */
  double yaw = mat_yaw(e->root_xmat);
  double mid[2] = {(e->foot_xpos[1][0] + e->foot_xpos[0][0]) / 2, (e->foot_xpos[1][1] + e->foot_xpos[0][1]) / 2};
  e->seq_len = n;
  for (int i = 0; i < ORC_NSLAB; i++) {
    if (i < n) {
      e->seq[i][0] = mid[0] + rel[i][0] * cos(yaw) - rel[i][1] * sin(yaw);
      e->seq[i][1] = mid[1] + rel[i][0] * sin(yaw) + rel[i][1] * cos(yaw);
      e->seq[i][2] = rel[i][2];
      e->seq[i][3] = yaw + rel[i][3];
    } else {
      e->seq[i][0] = 0; e->seq[i][1] = 0; e->seq[i][2] = -1; e->seq[i][3] = 0;   /* unused boxes (stepping_task.py:322) */
    }
  }
  step_update_targets(e);
}

/* SteppingTask.step (stepping_task.py:215-243) */
static void task_step_step(const orc_model* m, orc_env* e) {
  e->phase += 1;
  if (e->phase >= m->period) e->phase = 0;
  const double* tp = e->seq[e->t1];
  int in = 0;
  for (int f = 0; f < 2; f++) {
    double d[3] = {e->site_pos[f][0] - tp[0], e->site_pos[f][1] - tp[1], e->site_pos[f][2] - tp[2]};
    if (sqrt(dot3(d, d)) < m->target_radius) in = 1;
  }
  if (in) { e->target_reached = 1; e->target_reached_frames += 1; }
  else { e->target_reached = 0; e->target_reached_frames = 0; }
  if (e->target_reached && e->target_reached_frames >= m->delay_frames) {
    step_update_targets(e);
    e->target_reached = 0;
    e->target_reached_frames = 0;
  }
  step_update_goals(e);
}

/* SteppingTask.calc_reward (stepping_task.py:66-121); 6 terms in dict order, t[6..9] = 0 */
static void calc_reward_step(const orc_model* m, const orc_env* e, double* t) {
  double rfc = m->clock[0][e->phase], rvc = m->clock[1][e->phase], lfc = m->clock[2][e->phase],
         lvc = m->clock[3][e->phase];
  if (e->mode == ORC_STEP_STANDING) { rfc = lfc = 1; rvc = lvc = -1; }
  double fcap = m->total_mass * 9.8 * 0.5;
  double nl = fmin(e->lfoot_grf, fcap) / fcap * 2 - 1, nr = fmin(e->rfoot_grf, fcap) / fcap * 2 - 1;
  t[0] = 0.150 * ((tan(M_PI / 4 * lfc * nl) + tan(M_PI / 4 * rfc * nr)) / 2);
  double lv = sqrt(dot3(e->lfoot_vel, e->lfoot_vel)), rv = sqrt(dot3(e->rfoot_vel, e->rfoot_vel));
  double vl = fmin(lv, 0.2) / 0.2 * 2 - 1, vr = fmin(rv, 0.2) / 0.2 * 2 - 1;
  t[1] = 0.150 * ((tan(M_PI / 4 * lvc * vl) + tan(M_PI / 4 * rvc * vr)) / 2);
  /* orientation vs the yaw of target t1 (rewards.py:177-193): euler2quat(0, 0, th) = (cos th/2, 0, 0, sin th/2) */
  const double* sq = e->seq[e->t1];
  double inner = cos(0.5 * sq[3]) * e->root_quat[0] + sin(0.5 * sq[3]) * e->root_quat[3];
  t[2] = 0.050 * exp(-10 * (1 - inner * inner));
  double cz = (e->ncon_r + e->ncon_l) > 0 ? e->contact_z_min : 0.0;
  double herr = fabs(e->root_xpos[2] - cz - m->goal_height);
  if (herr < 0.01) herr = 0;                                       /* goal_speed_ref = 0 */
  t[3] = 0.050 * exp(-40 * herr * herr);
  /* step_reward (stepping_task.py:52-64) */
  double dmin = 1e300;
  for (int f = 0; f < 2; f++) {
    double d[3] = {e->site_pos[f][0] - sq[0], e->site_pos[f][1] - sq[1], e->site_pos[f][2] - sq[2]};
    double n = sqrt(dot3(d, d));
    if (n < dmin) dmin = n;
  }
  double hit = e->target_reached ? exp(-dmin / 0.25) : 0.0;
  const double* s2 = e->seq[e->t2];
  double mx = (sq[0] + s2[0]) / 2 - e->root_xpos[0], my = (sq[1] + s2[1]) / 2 - e->root_xpos[1];
  double progress = exp(-sqrt(mx * mx + my * my) / 2);
  t[4] = 0.450 * (0.8 * hit + 0.2 * progress);
  double hx = e->head_xpos[0] - e->root_xpos[0], hy = e->head_xpos[1] - e->root_xpos[1];
  t[5] = 0.050 * exp(-10 * (hx * hx + hy * hy));
  t[6] = t[7] = t[8] = t[9] = 0;
}

static void terrain_repose(const orc_model* m, orc_env* e);
static void task_reset(const orc_model* m, orc_env* e) {
  uint32_t u[4];
  if (m->task == ORC_TASK_STEP) { task_reset_step(m, e); return; }
  if (m->task == ORC_TASK_STAND) return; /* StandingTask.reset is empty (tasks/standing_task.py:33) */
  orc_philox(e->seed, e->env_id, e->rng_ctr, 3, u);
  double c = u01(u[0]);
  e->mode = c < 0.6 ? ORC_STANDING : (c < 0.8 ? ORC_INPLACE : ORC_FORWARD);
  sample_ref(e, 4);
  e->phase = randint(u[1], m->period);
  if (m->terrain && e->seq_len == 0) terrain_repose(m, e);   /* first reset of a new env: the terrain's initial pose */
}

/* terrain extension: re-pose the 20 terraces.  Draws at the current event counter: stream 5 lanes 0..2 = x, y offset
 * U(-xy, xy) and z offset U(zlo, zhi) (the ranges of the reference's manip_hfield hook, walking_task.py:172-179); terrace k's
 * bump height U(0, bump) = stream 60 + k/4 lane k%4.  Terrace k: centre (px + (k - 4) pitch, py), top at bump_k + zoff. */
static void terrain_repose(const orc_model* m, orc_env* e) {
  uint32_t u[4], w[4];
  orc_philox(e->seed, e->env_id, e->rng_ctr, 5, u);
  double px = -m->terrain_xy + 2 * m->terrain_xy * u01(u[0]), py = -m->terrain_xy + 2 * m->terrain_xy * u01(u[1]);
  double zo = m->terrain_zlo + (m->terrain_zhi - m->terrain_zlo) * u01(u[2]);
  for (int k = 0; k < ORC_NSLAB; k++) {
    if ((k & 3) == 0) orc_philox(e->seed, e->env_id, e->rng_ctr, 60 + (k >> 2), w);
    e->seq[k][0] = px + (k - 4) * m->terrain_pitch;
    e->seq[k][1] = py;
    e->seq[k][2] = m->terrain_bump * u01(w[k & 3]) + zo;
    e->seq[k][3] = 0;
  }
  e->seq_len = ORC_NSLAB;
}

static void task_step(const orc_model* m, orc_env* e) {
  uint32_t u[4];
  if (m->task == ORC_TASK_STEP) { task_step_step(m, e); return; }
  if (m->task == ORC_TASK_STAND) return;
  e->phase += 1;
  if (e->phase >= m->period) e->phase = 0;
  orc_philox(e->seed, e->env_id, e->rng_ctr, 0, u);
  int dbl = m->clock[0][e->phase] == 1.0 && m->clock[2][e->phase] == 1.0;
  if (randint(u[0], 100) == 0 && dbl) {
    if (e->mode == ORC_INPLACE) e->mode = ORC_STANDING;
    else if (e->mode == ORC_STANDING) e->mode = ORC_INPLACE;
    sample_ref(e, 1);
  }
  if (randint(u[1], 200) == 0 && e->mode != ORC_STANDING) {
    if (e->mode == ORC_FORWARD) e->mode = ORC_INPLACE;
    else if (e->mode == ORC_INPLACE) e->mode = ORC_FORWARD;
    sample_ref(e, 2);
  }
  if (m->terrain && randint(u[2], m->terrain_interval) == 0 && e->mode != ORC_STANDING) terrain_repose(m, e);
}

/* WalkingTask.calc_reward (tasks/walking_task.py:85-147); t[10] in the dict's insertion order */
static void calc_reward(const orc_model* m, const orc_env* e, const double* action, double* t) {
  int nu = m->nu;
  double rfc = m->clock[0][e->phase], rvc = m->clock[1][e->phase], lfc = m->clock[2][e->phase],
         lvc = m->clock[3][e->phase];
  double yaw_ref = e->mode_ref[0], vx = e->mode_ref[1], vy = e->mode_ref[2];
  if (e->mode == ORC_STANDING) { rfc = lfc = 1; rvc = lvc = -1; yaw_ref = vx = vy = 0; }
  else if (e->mode == ORC_INPLACE) { vx = vy = 0; }
  else { yaw_ref = 0; }
  double goal_speed = sqrt(vx * vx + vy * vy);
  /* foot force / velocity clock scores (rewards.py:107-174) */
  double fcap = m->total_mass * 9.8 * 0.5;
  double nl = fmin(e->lfoot_grf, fcap) / fcap * 2 - 1, nr = fmin(e->rfoot_grf, fcap) / fcap * 2 - 1;
  t[0] = 0.225 * ((tan(M_PI / 4 * lfc * nl) + tan(M_PI / 4 * rfc * nr)) / 2);
  double lv = sqrt(dot3(e->lfoot_vel, e->lfoot_vel)), rv = sqrt(dot3(e->rfoot_vel, e->rfoot_vel));
  double vl = fmin(lv, 0.2) / 0.2 * 2 - 1, vr = fmin(rv, 0.2) / 0.2 * 2 - 1;
  t[1] = 0.225 * ((tan(M_PI / 4 * lvc * vl) + tan(M_PI / 4 * rvc * vr)) / 2);
  /* root accel (rewards.py:93-104): current qvel[3:6], lagged qacc[0:3] */
  double err = 0;
  for (int x = 0; x < 3; x++) err += fabs(e->qvel[3 + x]);
  for (int x = 0; x < 3; x++) err += fabs(e->qacc[x]);
  t[2] = 0.05 * exp(-0.25 * err);
  /* height (rewards.py:68-90) */
  double cz = (e->ncon_r + e->ncon_l) > 0 ? e->contact_z_min : 0.0;
  double herr = fabs(e->root_xpos[2] - cz - m->goal_height);
  if (herr < 0.01 + 0.05 * goal_speed) herr = 0;
  t[3] = 0.05 * exp(-40 * herr * herr);
  /* planar velocity of the root origin in the root's local frame (mj_objectVelocity, flg_local=1) */
  double vloc[3];
  mattvec3(e->root_xmat, e->root_vlin, vloc);
  double ex = vloc[0] - vx, ey = vloc[1] - vy;
  t[4] = 0.15 * exp(-10 * (ex * ex + ey * ey));
  double ye = fabs(e->qvel[5] - yaw_ref);
  t[5] = 0.15 * exp(-10 * ye * ye * ye);
  double hx = e->head_xpos[0] - e->root_xpos[0], hy = e->head_xpos[1] - e->root_xpos[1];
  t[6] = 0.05 * exp(-10 * sqrt(hx * hx + hy * hy));
  double pe = 0, te = 0, ae = 0;
  for (int u = 0; u < nu; u++) {
    double d = m->nominal_qpos[7 + u] - e->act_len[u];
    pe += d * d;
    te += fabs(e->prev_torque[u] - e->act_force[u]);
    ae += fabs(e->prev_action[u] - action[u]);
  }
  t[7] = 0.05 * exp(-sqrt(pe));
  t[8] = 0.025 * exp(-0.25 * (te / nu));
  t[9] = 0.025 * exp(-5 * ae / nu);
}

/* StandingTask.calc_reward (tasks/standing_task.py:49-105); 6 terms in dict order, t[6..9] = 0 */
static void calc_reward_stand(const orc_model* m, const orc_env* e, double* t) {
  int nu = m->nu;
  double vloc[3];
  mattvec3(e->root_xmat, e->root_vlin, vloc);            /* get_body_vel('pelvis', frame=1)[0][:2] */
  double v2 = vloc[0] * vloc[0] + vloc[1] * vloc[1];
  t[0] = 0.3 * exp(-4 * v2);
  t[1] = 0.3 * exp(-4 * e->qvel[5] * e->qvel[5]);
  double he = e->root_xpos[2] - 0.98;
  t[2] = 0.1 * exp(-0.5 * he * he);
  /* torso_link is welded to the pelvis (waist joint removed, envs/h1/h1_env.py:21) at head_in_root */
  double ub2 = m->head_in_root[0] * m->head_in_root[0] + m->head_in_root[1] * m->head_in_root[1];
  t[3] = 0.1 * exp(-40 * ub2);
  double pe = 0, te = 0;
  for (int u = 0; u < nu; u++) {
    double d = e->act_len[u] - m->nominal_qpos[7 + u];
    pe += d * d;
    te += e->act_force[u] * e->act_force[u];
  }
  t[4] = 0.1 * exp(-5e-5 * te);
  t[5] = 0.1 * exp(-pe);
  t[6] = t[7] = t[8] = t[9] = 0;
}

/* randomize_dynamics (envs/common/domain_randomization.py:29-56).  Draw order restated on counter-based streams:
 * joint j -> stream 16 + j/2, lanes 2(j%2) (frictionloss) and 2(j%2)+1 (damping); body b (pelvis, then the leg
 * links in joint order) -> stream 21 + b, lane 0 mass scale, lanes 1..3 ipos offset.  body_inertia is not touched. */
static void randomize_dynamics(const orc_model* m, orc_env* e) {
  orc_params* P = &e->P;
  uint32_t w[4];
  for (int j = 0; j < m->nu; j++) {
    if ((j & 1) == 0) orc_philox(e->seed, e->env_id, e->rng_ctr, 16 + (j >> 1), w);
    P->frictionloss[6 + j] = 2.0 * u01(w[2 * (j & 1)]);
    P->damping[6 + j] = 0.02 + 1.98 * u01(w[2 * (j & 1) + 1]);
  }
  for (int b = 0; b <= m->nu; b++) {
    orc_philox(e->seed, e->env_id, e->rng_ctr, 21 + b, w);
    double sc = 0.95 + 0.1 * u01(w[0]);
    double off[3] = {-0.01 + 0.02 * u01(w[1]), -0.01 + 0.02 * u01(w[2]), -0.01 + 0.02 * u01(w[3])};
    if (b == 0) {
      P->pel_mass = m->pel_mass * sc;
      for (int x = 0; x < 3; x++) P->pel_com[x] = m->pel_com[x] + off[x];
    } else {
      P->mass[b] = m->mass[b] * sc;
      for (int x = 0; x < 3; x++) P->com[b][x] = m->com[b][x] + off[x];
    }
  }
  /* root link = pelvis body (+) welded rest: composite mass, com, inertia about the com */
  double mp = P->pel_mass, M = mp + m->rest_mass, c[3], Io[9];
  const double* cp = P->pel_com;
  for (int x = 0; x < 3; x++) c[x] = (mp * cp[x] + m->rest_mc[x]) / M;
  double cp2 = dot3(cp, cp), c2 = dot3(c, c);
  for (int a = 0; a < 3; a++)
    for (int b = 0; b < 3; b++) {
      Io[3 * a + b] = m->pel_Ic[3 * a + b] + mp * ((a == b ? cp2 : 0) - cp[a] * cp[b]) + m->rest_Io[3 * a + b];
      P->inertia[0][3 * a + b] = Io[3 * a + b] - M * ((a == b ? c2 : 0) - c[a] * c[b]);
    }
  P->mass[0] = M;
  memcpy(P->com[0], c, sizeof(c));
}

/* apply_perturbation (domain_randomization.py:10-26): per body, force U(-F,F)^3, torque U(-T,T)^3, then a coin that
 * clears the WHOLE xfrc_applied array.  body b -> streams 32+2b (force, lane 3 = coin) and 33+2b (torque) */
static void apply_perturbation(const orc_model* m, orc_env* e) {
  uint32_t wf[4], wt[4];
  for (int b = 0; b < 2; b++) {
    orc_philox(e->seed, e->env_id, e->rng_ctr, 32 + 2 * b, wf);
    orc_philox(e->seed, e->env_id, e->rng_ctr, 33 + 2 * b, wt);
    for (int x = 0; x < 3; x++) {
      e->xfrc[b][x] = -m->perturb_force + 2 * m->perturb_force * u01(wf[x]);
      e->xfrc[b][3 + x] = -m->perturb_torque + 2 * m->perturb_torque * u01(wt[x]);
    }
    if (randint(wf[3], 2) == 0) memset(e->xfrc, 0, sizeof(e->xfrc));
  }
}

/* _apply_init_noise (base_humanoid_env.py:281-309) on e->qpos (already the nominal pose): stream 50 lanes 0..2 = height,
 * roll, pitch; joint j -> stream 51 + j/4 lane j%4.  euler2quat(r, p, 0) 'sxyz' = (cp cr, cp sr, sp cr, -sp sr) */
static void init_pose_noise(const orc_model* m, orc_env* e) {
  double c = m->init_noise * M_PI / 180.0;
  uint32_t w[4];
  orc_philox(e->seed, e->env_id, e->rng_ctr, 50, w);
  e->qpos[2] = m->nominal_qpos[2] + 0.02 * u01(w[0]);
  double r = -c + 2 * c * u01(w[1]), p = -c + 2 * c * u01(w[2]);
  double cr = cos(0.5 * r), sr = sin(0.5 * r), cp = cos(0.5 * p), sp = sin(0.5 * p);
  e->qpos[3] = cp * cr; e->qpos[4] = cp * sr; e->qpos[5] = sp * cr; e->qpos[6] = -sp * sr;
  for (int j = 0; j < m->nu; j++) {
    if ((j & 3) == 0) orc_philox(e->seed, e->env_id, e->rng_ctr, 51 + (j >> 2), w);
    e->qpos[7 + j] = m->nominal_qpos[7 + j] + (-c + 2 * c * u01(w[j & 3]));
  }
}

void orc_env_init(const orc_model* m, orc_env* e, uint32_t seed, uint32_t env_id) {
  memset(e, 0, sizeof(*e));
  params_default(m, &e->P);
  e->seed = seed;
  e->env_id = env_id;
  e->qpos[3] = 1.0;
  /* a freshly compiled jvrc_step model has its 20 boxes below the floor (gen_xml.py:149: pos 0 0 -0.2): no contact */
  for (int i = 0; i < ORC_NSLAB; i++) e->seq[i][2] = -1;
  e->mode = m->task == ORC_TASK_STEP ? ORC_STEP_STANDING : 0;
}
/* test hook: the contact list mj_collision would produce at the env's current qpos (position, signed distance, foot 0/1,
 * 1 = against a stepping stone / 0 = floor); returns the number of contacts */
int orc_test_contacts(const orc_model* m, const orc_env* e, double* pos /* [ORC_MAXCON][3] */, double* dist, int* foot, int* slab) {
  kin_t k;
  static __thread efc_t efc;
  fk(m, &e->P, e->qpos, &k);
  make_constraints(m, &e->P, &k, e->qpos, e->qvel, e, &efc);
  for (int ci = 0; ci < efc.ncon; ci++) {
    for (int x = 0; x < 3; x++) pos[3 * ci + x] = efc.con_pos[ci][x];
    dist[ci] = efc.con_dist[ci];
    foot[ci] = efc.con_geom[ci];
    /* bit 0: against a stepping stone; bit 1: against one of its SIDE faces (riser contact, horizontal normal) */
    slab[ci] = efc.con_slab[ci] | ((efc.con_nrm[ci][0] != 0 || efc.con_nrm[ci][1] != 0) ? 2 : 0);
  }
  return efc.ncon;
}
/* test hooks for the SteppingTask pieces pinned by tests/golden/step_*.json */
void orc_test_task_reset(const orc_model* m, orc_env* e) { task_reset(m, e); }
void orc_test_task_step(const orc_model* m, orc_env* e) { task_step(m, e); }
void orc_set_step_height(orc_model* m, double h) { m->step_height = h; }

void orc_reset(const orc_model* m, orc_env* e, double* obs) {
  /* mj_resetData (mujoco_env.py:114): qvel, warmstart, ctrl <- 0 ; then reset_model: nominal pose, 3 zero-ctrl steps */
  double zero[ORC_NU] = {0};
  e->rng_ctr++;
  memset(e->xfrc, 0, sizeof(e->xfrc));                       /* mj_resetData clears xfrc_applied */
  if (m->dynrand_interval > 0) randomize_dynamics(m, e);     /* base_humanoid_env.py:252-253 */
  memcpy(e->qpos, m->nominal_qpos, m->nq * sizeof(double));
  if (m->init_noise > 0) init_pose_noise(m, e);
  memset(e->qvel, 0, sizeof(e->qvel));
  memset(e->qacc_warm, 0, sizeof(e->qacc_warm));
  for (int i = 0; i < 3; i++) orc_mj_step(m, e, zero);
  task_reset(m, e);
  memset(e->prev_prediction, 0, sizeof(e->prev_prediction));
  e->traj_len = 0; e->ep_len = 0; e->ep_rew = 0;
  if (obs) get_obs(m, e, obs);
}

void orc_step(const orc_model* m, orc_env* e, const double* action, double* obs, double* rew_terms, double* reward,
              int* done) {
  int nu = m->nu;
  double target[ORC_NU], ctrl[ORC_NU], t[ORC_NREW];
  /* base_humanoid_env.py:209-212 smoothing + nominal offsets ; robot_base.py:80 */
  for (int u = 0; u < nu; u++)
    target[u] = m->action_smoothing * action[u] + (1 - m->action_smoothing) * e->prev_prediction[u] + m->nominal_qpos[7 + u];
  if (!e->have_prev) { /* robot_base.py:82-85 */
    memcpy(e->prev_action, target, sizeof(target));
    memcpy(e->prev_torque, e->act_force, sizeof(e->prev_torque));
    e->have_prev = 1;
  }
  e->rng_ctr++;   /* one event counter per control step; every draw of this step is keyed by it */
  double kp[ORC_NU], kd[ORC_NU];
  memcpy(kp, m->kp, sizeof(kp));
  memcpy(kd, m->kd, sizeof(kd));
  if (m->pdrand_k > 0) {
    /* RobotBase._do_simulation (robots/robot_base.py:41-47): kp ~ U((1-k) kp, (1+k) kp), then kd likewise, once per
     * control step.  joint u -> stream 8 + u/4 (kp) / 11 + u/4 (kd), lane u%4 */
    uint32_t w[4];
    for (int u = 0; u < nu; u++) {
      if ((u & 3) == 0) orc_philox(e->seed, e->env_id, e->rng_ctr, 8 + (u >> 2), w);
      double lo = (1 - m->pdrand_k) * m->kp[u], hi = (1 + m->pdrand_k) * m->kp[u];
      kp[u] = lo + (hi - lo) * u01(w[u & 3]);
    }
    for (int u = 0; u < nu; u++) {
      if ((u & 3) == 0) orc_philox(e->seed, e->env_id, e->rng_ctr, 11 + (u >> 2), w);
      double lo = (1 - m->pdrand_k) * m->kd[u], hi = (1 + m->pdrand_k) * m->kd[u];
      kd[u] = lo + (hi - lo) * u01(w[u & 3]);
    }
  }
  for (int s = 0; s < m->frame_skip; s++) {
    for (int u = 0; u < nu; u++) ctrl[u] = kp[u] * (target[u] - e->act_len[u]) + kd[u] * (0.0 - e->act_vel[u]);
    orc_mj_step(m, e, ctrl);
  }
  task_step(m, e);
  if (m->task == ORC_TASK_STAND) calc_reward_stand(m, e, t);
  else if (m->task == ORC_TASK_STEP) calc_reward_step(m, e, t);
  else calc_reward(m, e, target, t);
  int d = (e->qpos[2] < m->done_lo) || (e->qpos[2] > m->done_hi) || e->self_collision || e->status;
  if (m->task == ORC_TASK_STEP) {   /* SteppingTask.done (stepping_task.py:248-260): root height above the lower foot site */
    double fz = e->site_pos[0][2] < e->site_pos[1][2] ? e->site_pos[0][2] : e->site_pos[1][2];
    d = (e->root_xpos[2] - fz < m->done_lo) || e->self_collision || e->status;
  }
  memcpy(e->prev_action, target, sizeof(target));
  memcpy(e->prev_torque, e->act_force, sizeof(e->prev_torque));
  if (obs) get_obs(m, e, obs);
  memcpy(e->prev_prediction, action, nu * sizeof(double));
  /* domain randomisation after the observation (base_humanoid_env.py:228-233); decisions on stream 0 lanes 2, 3 */
  if (m->dynrand_interval > 0 || m->perturb_interval > 0) {
    uint32_t w[4];
    orc_philox(e->seed, e->env_id, e->rng_ctr, 0, w);
    if (m->dynrand_interval > 0 && randint(w[2], m->dynrand_interval) == 0) randomize_dynamics(m, e);
    if (m->perturb_interval > 0 && randint(w[3], m->perturb_interval) == 0) apply_perturbation(m, e);
  }
  double sum = 0;
  for (int i = 0; i < ORC_NREW; i++) sum += t[i];
  if (rew_terms) memcpy(rew_terms, t, sizeof(t));
  if (reward) *reward = sum;
  if (done) *done = d;
}

void orc_step_autoreset(const orc_model* m, orc_env* e, const double* action, int max_traj_len, double* obs,
                        double* term_obs, double* rew_terms, double* reward, int* done, int* ended) {
  double r; int d;
  double o[ORC_NOBS];
  orc_step(m, e, action, o, rew_terms, &r, &d);
  e->traj_len++; e->ep_len++; e->ep_rew += r;
  int end = d || (e->traj_len >= max_traj_len);
  if (term_obs) memcpy(term_obs, o, m->nobs * sizeof(double));
  if (end) {
    e->status = 0;
    orc_reset(m, e, o);
  }
  if (obs) memcpy(obs, o, m->nobs * sizeof(double));
  if (reward) *reward = r;
  if (done) *done = d;
  if (ended) *ended = end;
}

void orc_batch_reset(const orc_model* m, orc_env* envs, int n, double* obs, int nthreads) {
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : omp_get_max_threads()) schedule(static)
#endif
  for (int i = 0; i < n; i++) orc_reset(m, envs + i, obs ? obs + (size_t)i * m->nobs : 0);
  (void)nthreads;
}

void orc_batch_step_autoreset(const orc_model* m, orc_env* envs, int n, const double* actions, int max_traj_len,
                              double* obs, double* term_obs, double* rew_terms, double* reward, int* done, int* ended,
                              int nthreads) {
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : omp_get_max_threads()) schedule(dynamic, 4)
#endif
  for (int i = 0; i < n; i++)
    orc_step_autoreset(m, envs + i, actions + (size_t)i * m->nu, max_traj_len, obs ? obs + (size_t)i * m->nobs : 0,
                       term_obs ? term_obs + (size_t)i * m->nobs : 0, rew_terms ? rew_terms + (size_t)i * ORC_NREW : 0,
                       reward ? reward + i : 0, done ? done + i : 0, ended ? ended + i : 0);
  (void)nthreads;
}

/* test hook: reward terms for the current env fields (lets tests/ pin calc_reward to the golden vectors) */
void orc_calc_reward(const orc_model* m, const orc_env* e, const double* target, double* terms) {
  if (m->task == ORC_TASK_STEP) calc_reward_step(m, e, terms);
  else if (m->task == ORC_TASK_STAND) calc_reward_stand(m, e, terms);
  else calc_reward(m, e, target, terms);
}
/* test hooks for the H1 pieces pinned by tests/golden/h1_*.json (each uses the env's current rng_ctr) */
/* the PD gains a control step with event counter rng_ctr would use */
void orc_test_pd_gains(const orc_model* m, const orc_env* e, double* kp, double* kd) {
  uint32_t w[4];
  for (int u = 0; u < m->nu; u++) {
    if ((u & 3) == 0) orc_philox(e->seed, e->env_id, e->rng_ctr, 8 + (u >> 2), w);
    double lo = (1 - m->pdrand_k) * m->kp[u], hi = (1 + m->pdrand_k) * m->kp[u];
    kp[u] = lo + (hi - lo) * u01(w[u & 3]);
  }
  for (int u = 0; u < m->nu; u++) {
    if ((u & 3) == 0) orc_philox(e->seed, e->env_id, e->rng_ctr, 11 + (u >> 2), w);
    double lo = (1 - m->pdrand_k) * m->kd[u], hi = (1 + m->pdrand_k) * m->kd[u];
    kd[u] = lo + (hi - lo) * u01(w[u & 3]);
  }
}
void orc_test_randomize_dynamics(const orc_model* m, orc_env* e) { randomize_dynamics(m, e); }
void orc_test_apply_perturbation(const orc_model* m, orc_env* e) { apply_perturbation(m, e); }
void orc_test_get_obs(const orc_model* m, const orc_env* e, double* obs) { get_obs(m, e, obs); }
/* the noisy initial pose alone (reset_model before set_state), without the settling steps */
void orc_test_init_pose(const orc_model* m, orc_env* e) {
  memcpy(e->qpos, m->nominal_qpos, m->nq * sizeof(double));
  init_pose_noise(m, e);
}
