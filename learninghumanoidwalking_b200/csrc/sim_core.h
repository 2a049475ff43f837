// sim_core.h — one environment per warp: articulated-body forward dynamics, soft-contact Newton solve,
// PD actuation, reward / observation / termination / auto-reset, all inside one control step.
//
// Replaces (reference file:line, /root/reference):
//   envs/common/base_humanoid_env.py:199-227 (step), :247-276 (reset_model)
//   robots/robot_base.py:41-98 (_do_simulation: frame_skip x {PD torque, mj_step})
//   envs/common/robot_interface.py:493-546 (step_pd / set_motor_torque / step)
//   mujoco.mj_step (external)
//   tasks/walking_task.py:85-205, tasks/rewards.py:9-174, tasks/observations.py:12-72,
//   envs/jvrc/jvrc_base.py:133-145, envs/jvrc/jvrc_walk.py:65-67
//
// Execution model.  Every phase is a `LHW_LANES(l) { ... }` block: on the GPU each of the 32 lanes of the
// warp runs the body once with its own lane id and `LHW_SYNC()` is __syncwarp(); all inter-lane traffic goes
// through the per-warp shared-memory `Work` struct (or warp_sum shuffles).  With LHW_CPU_EMU defined the same
// source runs the 32 lanes sequentially on the host — that build exists ONLY for tests/ (it lets the CPU test
// tier check this exact arithmetic against oracle/); the product never uses it.
//
// Spatial algebra: all motion/force vectors are expressed in a world-aligned inertial frame whose origin `o`
// coincides with the root link origin at the current instant, so parent->child propagation needs no
// transforms (V_child = V_parent + S qd, composite inertias are plain sums).
// Tree: free root (dofs 0-2 world translation, 3-5 body-frame rotation) + two serial chains of NJ hinges.
// This gives the mass matrix an arrow structure [root | chain0 | chain1] with a zero chain0-chain1 block,
// which the factorisation exploits (two NJ x NJ Choleskys side by side + a 6x6 Schur complement).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__) && !defined(LHW_CPU_EMU)
#define LHW_DEV __device__ __forceinline__
#define LHW_DEVNI __device__ __noinline__
#else
#define LHW_DEV inline
#define LHW_DEVNI inline
#endif

#if defined(__CUDA_ARCH__) && !defined(LHW_CPU_EMU)
#define LHW_LANES(l) for (int l = (int)(threadIdx.x & 31), _o = 1; _o; _o = 0)
#define LHW_SYNC() __syncwarp()
#else
#define LHW_LANES(l) for (int l = 0; l < 32; ++l)
#define LHW_SYNC() ((void)0)
#endif

namespace lhw {

constexpr int NCON = 8;        // 2 feet x 4 box corners (mjc_PlaneBox returns at most 4)
constexpr int NEDGE = 32;      // pyramidal condim 3: 4 edges per contact
constexpr int NREW = 10;
constexpr int MAXPERIOD = 96;
constexpr int NSTATE_I = 8;    // int32 words per env in the integer state record

enum { STANDING = 0, INPLACE = 1, FORWARD = 2 };

// ---------------------------------------------------------------- math wrappers
LHW_DEV float m_sqrt(float x) { return sqrtf(x); }
LHW_DEV double m_sqrt(double x) { return sqrt(x); }
LHW_DEV float m_abs(float x) { return fabsf(x); }
LHW_DEV double m_abs(double x) { return fabs(x); }
LHW_DEV float m_exp(float x) { return expf(x); }
LHW_DEV double m_exp(double x) { return exp(x); }
LHW_DEV float m_tan(float x) { return tanf(x); }
LHW_DEV double m_tan(double x) { return tan(x); }
LHW_DEV float m_atan2(float y, float x) { return atan2f(y, x); }
LHW_DEV double m_atan2(double y, double x) { return atan2(y, x); }
LHW_DEV float m_pow(float x, float y) { return powf(x, y); }
LHW_DEV double m_pow(double x, double y) { return pow(x, y); }
LHW_DEV void m_sincos(float x, float* s, float* c) { *s = sinf(x); *c = cosf(x); }
LHW_DEV void m_sincos(double x, double* s, double* c) { *s = sin(x); *c = cos(x); }
template <class T> LHW_DEV T m_min(T a, T b) { return a < b ? a : b; }
template <class T> LHW_DEV T m_max(T a, T b) { return a > b ? a : b; }

template <class real, class F> LHW_DEV real warp_sum(F f) {
#if defined(__CUDA_ARCH__) && !defined(LHW_CPU_EMU)
  real v = f((int)(threadIdx.x & 31));
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
#else
  real s = 0;
  for (int l = 0; l < 32; ++l) s += f(l);
  return s;
#endif
}

// ---------------------------------------------------------------- rng: philox4x32-10, counter (event, stream, env)
LHW_DEV void philox(uint32_t seed, uint32_t env_id, uint32_t ctr, uint32_t stream, uint32_t out[4]) {
  uint32_t c0 = ctr, c1 = stream, c2 = env_id, c3 = 0x4c485742u;
  uint32_t k0 = seed, k1 = 0x9E3779B9u ^ (seed * 0x85EBCA6Bu + 1u);
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
template <class real> LHW_DEV real u01(uint32_t u) { return (real)(u >> 8) * (real)(1.0 / 16777216.0); }
LHW_DEV int randint(uint32_t u, int n) { return (int)(((uint64_t)u * (uint64_t)n) >> 32); }

// ---------------------------------------------------------------- model constants (one per robot/task)
template <class real, int NJ> struct Model {
  static constexpr int NL = 1 + 2 * NJ, NV = 6 + 2 * NJ, NQ = NV + 1, NU = 2 * NJ, NA = 6 + NJ;
  static constexpr int NT = 21 + NJ * (NJ + 1) + 12 * NJ;  // structurally non-zero lower-triangle entries of H
  real link_pos[NL][3], link_rot[NL][9], axis[NL][3];
  real mass[NL], com[NL][3], inertia[NL][6];  // xx yy zz xy xz yz about com, link frame
  real armature[NV], damping[NV], range_lo[NV], range_hi[NV], dof_invw[NV];
  real foot_pos[2][3], foot_size[2][3], foot_invw[2];
  real h, grav[3];
  real K, B, solimp[5], mu, mu_reg;  // mu_reg = mu * sqrt(1/impratio)
  real tol2;                         // (tolerance * meaninertia * nv)^2 : threshold on |grad|^2
  real kp[NU], kd[NU], nominal[NQ], smoothing;
  real head[3], fcap, goal_height;
  real clock[4][MAXPERIOD];  // r_frc r_vel l_frc l_vel
  int max_iter, frame_skip, period, any_damping;
  unsigned char h_i[NT], h_j[NT];
};

template <class real, int NJ> struct Dims {
  static constexpr int NL = 1 + 2 * NJ, NV = 6 + 2 * NJ, NQ = NV + 1, NU = 2 * NJ, NA = 6 + NJ;
  static constexpr int NOBS = 5 + 2 * NU + 8;
  // real-valued state record per env (HBM, env-major so one warp streams its env's record contiguously)
  static constexpr int NSTATE_R = NQ + NV + NV + 5 * NU + 3 + 1;
};

// ---------------------------------------------------------------- per-warp working set (shared memory)
template <class real, int NJ> struct Work {
  static constexpr int NL = 1 + 2 * NJ, NV = 6 + 2 * NJ, NQ = NV + 1, NU = 2 * NJ, NA = 6 + NJ;
  static constexpr int NOBS = 5 + 2 * NU + 8;
  // ---- persistent state (mirrors the HBM record, same order)
  real qpos[NQ], qvel[NV], qacc_warm[NV], act_len[NU], act_vel[NU], prev_pred[NU], prev_action[NU], prev_torque[NU];
  real mode_ref[3], ep_rew;
  int phase, mode, traj_len, ep_len, have_prev, status;
  uint32_t rng_ctr, env_id;
  // ---- control-step scratch
  real target[NU], ctrl[NU], act_force[NU];
  // ---- kinematics / dynamics
  real o[3], xr[NL][3], xmat[NL][9];
  real S[NV][6];
  real inert[NL][10], comp[NL][10];
  real V[NL][6], A[NL][6], F[NL][6];
  real M[NV][NV], H[NV][NV], hdinv[NV];
  real qfs[NV], qacc[NV], Ma[NV], grad[NV], sdir[NV], Ms[NV], vec[NV];
  // ---- contacts (slot = foot*4 + k) and joint limits
  int ncon[2];
  real cpos[NCON][3], cdist[NCON], cD[NCON], cKid[NCON];
  real Jc[NCON][3][NA], WJ[NCON][3][NA];
  real cu[NCON][3], cW[NCON][5], cF[NCON][3];
  real earef[NEDGE], ejar[NEDGE], ejv[NEDGE], ef[NEDGE];
  int eact[NEDGE];
  int lside[NU], lact[NU];
  real laref[NU], lD[NU], ljar[NU], ljv[NU], lf[NU];
  // ---- what mj_step leaves behind (pre-integration state of the last substep)
  real root_vlin[3], foot_vel[2][3], grf[2], cz_min, qacc_lag[3];
  real rew[NREW], obs[NOBS];
  int iters_total;
};

// ---------------------------------------------------------------- small vector helpers
template <class real> LHW_DEV void cross(const real* a, const real* b, real* c) {
  real x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  c[0] = x; c[1] = y; c[2] = z;
}
template <class real> LHW_DEV real dot3(const real* a, const real* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <class real> LHW_DEV void mv3(const real* R, const real* v, real* out) {
  real x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
  real y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
  real z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
  out[0] = x; out[1] = y; out[2] = z;
}
// spatial inertia (m, h=m*c, Io) times motion vector [w; v] -> force [n; f]
template <class real> LHW_DEV void inert_mul(const real* I, const real* mvv, real* out) {
  const real m = I[0];
  const real* h = I + 1;
  const real* w = mvv;
  const real* v = mvv + 3;
  real hv[3], hw[3];
  cross(h, v, hv);
  cross(h, w, hw);
  out[0] = I[4] * w[0] + I[7] * w[1] + I[8] * w[2] + hv[0];
  out[1] = I[7] * w[0] + I[5] * w[1] + I[9] * w[2] + hv[1];
  out[2] = I[8] * w[0] + I[9] * w[1] + I[6] * w[2] + hv[2];
  out[3] = m * v[0] - hw[0];
  out[4] = m * v[1] - hw[1];
  out[5] = m * v[2] - hw[2];
}

// dof d -> link that carries it ; link i>0 -> its dof
template <int NJ> LHW_DEV int dof_link(int d) { return d < 6 ? 0 : d - 5; }
// local (per-foot) jacobian column -> global dof
template <int NJ> LHW_DEV int loc2dof(int foot, int j) { return j < 6 ? j : 6 + foot * NJ + (j - 6); }


// ================================================================= arrow-structured Cholesky of H (in place)
// ordering [chain0 | chain1 | root]:  H = [[A0,0,B0'],[0,A1,B1'],[B0,B1,C]]  ->  A_c = L_c L_c',  X_c = B_c L_c^-T,
// C - sum_c X_c X_c' = L_C L_C'.  L_c sits in the lower triangle of H[chain][chain], X_c[r][k] in H[chain k][root r],
// L_C in the lower triangle of H[root][root]; reciprocal pivots in hdinv (the diagonal of H is left untouched).
template <class real, int NJ> LHW_DEV void arrow_factor(Work<real, NJ>& w) {
  for (int k = 0; k < NJ; k++) {
    LHW_LANES(l) {
      const int ch = l >> 4, r = l & 15;
      if (r < NJ + 6 && (r >= NJ || r >= k)) {
        const int c0 = 6 + ch * NJ, ck = c0 + k;
        real dk = w.H[ck][ck];
        for (int mm = 0; mm < k; mm++) dk -= w.H[ck][c0 + mm] * w.H[ck][c0 + mm];
        if (!(dk > 0)) { dk = (real)1e-30; w.status |= 2; }
        const real inv = (real)1 / m_sqrt(dk);
        if (r == k) {
          w.hdinv[ck] = inv;
        } else if (r < NJ) {
          const int ci = c0 + r;
          real t = w.H[ci][ck];
          for (int mm = 0; mm < k; mm++) t -= w.H[ci][c0 + mm] * w.H[ck][c0 + mm];
          w.H[ci][ck] = t * inv;
        } else {
          const int rr = r - NJ;
          real t = w.H[ck][rr];
          for (int mm = 0; mm < k; mm++) t -= w.H[c0 + mm][rr] * w.H[ck][c0 + mm];
          w.H[ck][rr] = t * inv;
        }
      }
    }
    LHW_SYNC();
  }
  LHW_LANES(l) {
    if (l < 21) {
      int r = 0, t = l;
      while (t > r) { t -= r + 1; r++; }
      const int c = t;  // (r,c), r >= c, root block
      real acc = w.H[r][c];
      for (int j = 6; j < 6 + 2 * NJ; j++) acc -= w.H[j][r] * w.H[j][c];
      w.H[r][c] = acc;
    }
  }
  LHW_SYNC();
  for (int k = 0; k < 6; k++) {
    LHW_LANES(l) {
      if (l < 6 && l >= k) {
        real dk = w.H[k][k];
        for (int mm = 0; mm < k; mm++) dk -= w.H[k][mm] * w.H[k][mm];
        if (!(dk > 0)) { dk = (real)1e-30; w.status |= 2; }
        const real inv = (real)1 / m_sqrt(dk);
        if (l == k) w.hdinv[k] = inv;
        else {
          real t = w.H[l][k];
          for (int mm = 0; mm < k; mm++) t -= w.H[l][mm] * w.H[k][mm];
          w.H[l][k] = t * inv;
        }
      }
    }
    LHW_SYNC();
  }
}

// x <- H^-1 x using the factor above
template <class real, int NJ> LHW_DEV void arrow_solve(Work<real, NJ>& w, real* x) {
  LHW_LANES(l) {
    if (l < 2) {
      const int c0 = 6 + l * NJ;
      for (int k = 0; k < NJ; k++) {
        real t = x[c0 + k];
        for (int mm = 0; mm < k; mm++) t -= w.H[c0 + k][c0 + mm] * x[c0 + mm];
        x[c0 + k] = t * w.hdinv[c0 + k];
      }
    }
  }
  LHW_SYNC();
  LHW_LANES(l) {
    if (l < 6) {
      real t = x[l];
      for (int j = 6; j < 6 + 2 * NJ; j++) t -= w.H[j][l] * x[j];
      x[l] = t;
    }
  }
  LHW_SYNC();
  LHW_LANES(l) {
    if (l == 0) {
      for (int k = 0; k < 6; k++) {
        real t = x[k];
        for (int mm = 0; mm < k; mm++) t -= w.H[k][mm] * x[mm];
        x[k] = t * w.hdinv[k];
      }
      for (int k = 5; k >= 0; k--) {
        real t = x[k];
        for (int mm = k + 1; mm < 6; mm++) t -= w.H[mm][k] * x[mm];
        x[k] = t * w.hdinv[k];
      }
    }
  }
  LHW_SYNC();
  LHW_LANES(l) {
    const int ch = l >> 4, k = l & 15;
    if (k < NJ) {
      const int ck = 6 + ch * NJ + k;
      real t = x[ck];
      for (int r = 0; r < 6; r++) t -= w.H[ck][r] * x[r];
      x[ck] = t;
    }
  }
  LHW_SYNC();
  LHW_LANES(l) {
    if (l < 2) {
      const int c0 = 6 + l * NJ;
      for (int k = NJ - 1; k >= 0; k--) {
        real t = x[c0 + k];
        for (int mm = k + 1; mm < NJ; mm++) t -= w.H[c0 + mm][c0 + k] * x[c0 + mm];
        x[c0 + k] = t * w.hdinv[c0 + k];
      }
    }
  }
  LHW_SYNC();
}

// ================================================================= one physics substep (mujoco.mj_step)
template <class real, int NJ>
LHW_DEV void substep(Work<real, NJ>& w, const Model<real, NJ>& m, const bool last) {
  constexpr int NL = 1 + 2 * NJ, NV = 6 + 2 * NJ, NU = 2 * NJ, NA = 6 + NJ;
  // ---------------- P1 forward kinematics: root, then both chains level by level (lane = chain)
  LHW_LANES(l) {
    if (l == 0) {
      real q0 = w.qpos[3], q1 = w.qpos[4], q2 = w.qpos[5], q3 = w.qpos[6];
      real n = (real)1 / m_sqrt(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
      q0 *= n; q1 *= n; q2 *= n; q3 *= n;
      real* R = w.xmat[0];
      R[0] = 1 - 2 * (q2 * q2 + q3 * q3); R[1] = 2 * (q1 * q2 - q0 * q3); R[2] = 2 * (q1 * q3 + q0 * q2);
      R[3] = 2 * (q1 * q2 + q0 * q3); R[4] = 1 - 2 * (q1 * q1 + q3 * q3); R[5] = 2 * (q2 * q3 - q0 * q1);
      R[6] = 2 * (q1 * q3 - q0 * q2); R[7] = 2 * (q2 * q3 + q0 * q1); R[8] = 1 - 2 * (q1 * q1 + q2 * q2);
      w.o[0] = w.qpos[0]; w.o[1] = w.qpos[1]; w.o[2] = w.qpos[2];
      w.xr[0][0] = w.xr[0][1] = w.xr[0][2] = 0;
    }
  }
  LHW_SYNC();
  for (int k = 0; k < NJ; k++) {
    LHW_LANES(l) {
      if (l < 2) {
        const int i = 1 + l * NJ + k, p = k == 0 ? 0 : i - 1;
        real off[3], R0[9];
        mv3(w.xmat[p], m.link_pos[i], off);
        for (int c = 0; c < 3; c++) w.xr[i][c] = w.xr[p][c] + off[c];
        for (int r = 0; r < 3; r++)
          for (int c = 0; c < 3; c++)
            R0[3 * r + c] = w.xmat[p][3 * r] * m.link_rot[i][c] + w.xmat[p][3 * r + 1] * m.link_rot[i][3 + c] +
                            w.xmat[p][3 * r + 2] * m.link_rot[i][6 + c];
        const real* a = m.axis[i];
        real s, c;
        m_sincos(w.qpos[6 + i], &s, &c);
        const real t = 1 - c;
        real Rj[9] = {c + a[0] * a[0] * t,        a[0] * a[1] * t - a[2] * s, a[0] * a[2] * t + a[1] * s,
                      a[0] * a[1] * t + a[2] * s, c + a[1] * a[1] * t,        a[1] * a[2] * t - a[0] * s,
                      a[0] * a[2] * t - a[1] * s, a[1] * a[2] * t + a[0] * s, c + a[2] * a[2] * t};
        for (int r = 0; r < 3; r++)
          for (int cc = 0; cc < 3; cc++)
            w.xmat[i][3 * r + cc] = R0[3 * r] * Rj[cc] + R0[3 * r + 1] * Rj[3 + cc] + R0[3 * r + 2] * Rj[6 + cc];
      }
    }
    LHW_SYNC();
  }
  // ---------------- P2 motion vectors S (per dof) and link spatial inertias about o (per link)
  LHW_LANES(l) {
    if (l < NV) {
      real* S = w.S[l];
      if (l < 3) {
        for (int c = 0; c < 6; c++) S[c] = 0;
        S[3 + l] = 1;
      } else if (l < 6) {
        const int k = l - 3;
        S[0] = w.xmat[0][k]; S[1] = w.xmat[0][3 + k]; S[2] = w.xmat[0][6 + k];
        S[3] = S[4] = S[5] = 0;
      } else {
        const int i = l - 5;
        mv3(w.xmat[i], m.axis[i], S);
        cross(w.xr[i], S, S + 3);  // velocity at o of a rotation about the axis through xr: w x (o - p) = p x w
      }
    }
    if (l >= 32 - NL) {  // the other end of the warp builds the link inertias concurrently
      const int i = l - (32 - NL);
      const real* R = w.xmat[i];
      real c[3];
      mv3(R, m.com[i], c);
      for (int x = 0; x < 3; x++) c[x] += w.xr[i][x];
      const real* Ib = m.inertia[i];
      const real B[9] = {Ib[0], Ib[3], Ib[4], Ib[3], Ib[1], Ib[5], Ib[4], Ib[5], Ib[2]};
      real T[9];
      for (int r = 0; r < 3; r++)
        for (int cc = 0; cc < 3; cc++) T[3 * r + cc] = R[3 * r] * B[cc] + R[3 * r + 1] * B[3 + cc] + R[3 * r + 2] * B[6 + cc];
      real Iw[6];  // xx yy zz xy xz yz
      const int ri[6] = {0, 1, 2, 0, 0, 1}, ci[6] = {0, 1, 2, 1, 2, 2};
      for (int e = 0; e < 6; e++)
        Iw[e] = T[3 * ri[e]] * R[3 * ci[e]] + T[3 * ri[e] + 1] * R[3 * ci[e] + 1] + T[3 * ri[e] + 2] * R[3 * ci[e] + 2];
      const real ms = m.mass[i], cc2 = dot3(c, c);
      real* I = w.inert[i];
      I[0] = ms; I[1] = ms * c[0]; I[2] = ms * c[1]; I[3] = ms * c[2];
      I[4] = Iw[0] + ms * (cc2 - c[0] * c[0]); I[5] = Iw[1] + ms * (cc2 - c[1] * c[1]); I[6] = Iw[2] + ms * (cc2 - c[2] * c[2]);
      I[7] = Iw[3] - ms * c[0] * c[1]; I[8] = Iw[4] - ms * c[0] * c[2]; I[9] = Iw[5] - ms * c[1] * c[2];
    }
  }
  LHW_SYNC();
  // ---------------- P3 composite inertias: suffix sums along each chain (lane = chain*10 + component), then root
  LHW_LANES(l) {
    if (l < 20) {
      const int ch = l / 10, e = l - ch * 10;
      real acc = 0;
      for (int k = NJ - 1; k >= 0; k--) {
        const int i = 1 + ch * NJ + k;
        acc += w.inert[i][e];
        w.comp[i][e] = acc;
      }
    }
  }
  LHW_SYNC();
  LHW_LANES(l) {
    if (l < 10) w.comp[0][l] = w.inert[0][l] + w.comp[1][l] + w.comp[1 + NJ][l];
  }
  LHW_SYNC();
  // ---------------- P4 mass matrix (CRBA), lane = dof; concurrently link velocities V (lanes 20..31: chain x 6 comps)
  LHW_LANES(l) {
    if (l < NV) {
      real f[6];
      inert_mul(w.comp[dof_link<NJ>(l)], w.S[l], f);
      int j = l;
      const int cstart = l < 6 ? 0 : 6 + ((l - 6) / NJ) * NJ;
      for (; j >= cstart; j--) {
        const real* Sj = w.S[j];
        real v = Sj[0] * f[0] + Sj[1] * f[1] + Sj[2] * f[2] + Sj[3] * f[3] + Sj[4] * f[4] + Sj[5] * f[5];
        if (j == l) v += m.armature[l];
        w.M[l][j] = v; w.M[j][l] = v;
      }
      if (l >= 6)
        for (j = 5; j >= 0; j--) {
          const real* Sj = w.S[j];
          real v = Sj[0] * f[0] + Sj[1] * f[1] + Sj[2] * f[2] + Sj[3] * f[3] + Sj[4] * f[4] + Sj[5] * f[5];
          w.M[l][j] = v; w.M[j][l] = v;
        }
    } else if (l >= 20) {
      const int ch = (l - 20) / 6, e = (l - 20) % 6;
      // root spatial velocity component e : [R w_body ; v_lin]
      real acc = e < 3 ? w.xmat[0][3 * e] * w.qvel[3] + w.xmat[0][3 * e + 1] * w.qvel[4] + w.xmat[0][3 * e + 2] * w.qvel[5]
                       : w.qvel[e - 3];
      if (ch == 0) w.V[0][e] = acc;
      for (int k = 0; k < NJ; k++) {
        const int i = 1 + ch * NJ + k;
        acc += w.S[5 + i][e] * w.qvel[5 + i];
        w.V[i][e] = acc;
      }
    }
  }
  LHW_SYNC();
  // ---------------- P5 bias accelerations A (lane = chain, serial along the chain), gravity folded in as -g
  LHW_LANES(l) {
    if (l < 2) {
      real A[6];
      const real* V0 = w.V[0];
      A[0] = A[1] = A[2] = 0;
      cross(V0 + 3, V0, A + 3);  // v_o x w : rotational free-joint axes move with the body
      for (int c = 0; c < 3; c++) A[3 + c] -= m.grav[c];
      if (l == 0)
        for (int c = 0; c < 6; c++) w.A[0][c] = A[c];
      for (int k = 0; k < NJ; k++) {
        const int i = 1 + l * NJ + k;
        const real* Vi = w.V[i];
        const real* S = w.S[5 + i];
        const real qd = w.qvel[5 + i];
        real t1[3], t2[3], t3[3];
        cross(Vi, S, t1);          // w x w_s
        cross(Vi, S + 3, t2);      // w x v_s
        cross(Vi + 3, S, t3);      // v x w_s
        for (int c = 0; c < 3; c++) {
          A[c] += t1[c] * qd;
          A[3 + c] += (t2[c] + t3[c]) * qd;
        }
        for (int c = 0; c < 6; c++) w.A[i][c] = A[c];
      }
    }
  }
  LHW_SYNC();
  // ---------------- P6 link forces f_i = I A + V x* (I V) (lane = link) ; contact detection (lanes 30,31 = feet)
  LHW_LANES(l) {
    if (l < NL) {
      real IA[6], IV[6], t1[3], t2[3], t3[3];
      inert_mul(w.inert[l], w.A[l], IA);
      inert_mul(w.inert[l], w.V[l], IV);
      const real* V = w.V[l];
      cross(V, IV, t1);          // w x n
      cross(V + 3, IV + 3, t2);  // v x f
      cross(V, IV + 3, t3);      // w x f
      for (int c = 0; c < 3; c++) {
        w.F[l][c] = IA[c] + t1[c] + t2[c];
        w.F[l][3 + c] = IA[3 + c] + t3[c];
      }
    } else if (l >= 30) {
      // mjc_PlaneBox against the ground plane z = 0 (normal +z): corners in index order, at most 4
      const int f = l - 30, lk = (f + 1) * NJ;
      real ctr[3], v[3], corner[3];
      mv3(w.xmat[lk], m.foot_pos[f], ctr);
      for (int c = 0; c < 3; c++) ctr[c] += w.xr[lk][c];
      const real dist0 = w.o[2] + ctr[2];
      int cnt = 0;
      for (int i = 0; i < 8 && cnt < 4; i++) {
        v[0] = (i & 1) ? m.foot_size[f][0] : -m.foot_size[f][0];
        v[1] = (i & 2) ? m.foot_size[f][1] : -m.foot_size[f][1];
        v[2] = (i & 4) ? m.foot_size[f][2] : -m.foot_size[f][2];
        mv3(w.xmat[lk], v, corner);
        const real ld = corner[2];
        if (dist0 + ld > 0 || ld > 0) continue;
        const int s = f * 4 + cnt;
        const real cd = dist0 + ld;
        w.cdist[s] = cd;
        w.cpos[s][0] = corner[0] + ctr[0];
        w.cpos[s][1] = corner[1] + ctr[1];
        w.cpos[s][2] = corner[2] + ctr[2] - (real)0.5 * cd;
        // impedance (getimpedance), regulariser and reference stiffness term
        const real d0 = m.solimp[0], dw = m.solimp[1], width = m.solimp[2], mid = m.solimp[3], power = m.solimp[4];
        real x = m_abs(cd) / width, imp;
        if (x >= 1) imp = dw;
        else if (x <= 0) imp = d0;
        else {
          real y;
          if (power < (real)1.0000001) y = x;
          else if (x <= mid) y = m_pow(x / mid, power) * mid;
          else y = 1 - m_pow((1 - x) / (1 - mid), power) * (1 - mid);
          imp = d0 + y * (dw - d0);
        }
        const real Rn = m_max((real)1e-15, (1 - imp) / imp * (m.foot_invw[f] * (1 + m.mu * m.mu)));
        w.cD[s] = (real)1 / (2 * m.mu_reg * m.mu_reg * Rn);
        w.cKid[s] = m.K * imp * cd;
        cnt++;
      }
      w.ncon[f] = cnt;
    }
  }
  LHW_SYNC();
  // ---------------- P7 subtree forces (suffix sums, lane = chain*6 + comp) ; joint limits (lanes 12..12+NU)
  LHW_LANES(l) {
    if (l < 12) {
      const int ch = l / 6, e = l - ch * 6;
      real acc = 0;
      for (int k = NJ - 1; k >= 0; k--) {
        const int i = 1 + ch * NJ + k;
        acc += w.F[i][e];
        w.F[i][e] = acc;
      }
    } else if (l < 12 + NU) {
      const int u = l - 12, d = 6 + u;
      const real q = w.qpos[7 + u];
      const real dlo = q - m.range_lo[d], dhi = m.range_hi[d] - q;
      int side = 0;
      real dist = 0;
      if (dlo < 0) { side = 1; dist = dlo; }
      else if (dhi < 0) { side = -1; dist = dhi; }
      w.lside[u] = side;
      if (side) {
        const real d0 = m.solimp[0], dw = m.solimp[1], width = m.solimp[2], mid = m.solimp[3], power = m.solimp[4];
        real x = m_abs(dist) / width, imp;
        if (x >= 1) imp = dw;
        else {
          real y;
          if (power < (real)1.0000001) y = x;
          else if (x <= mid) y = m_pow(x / mid, power) * mid;
          else y = 1 - m_pow((1 - x) / (1 - mid), power) * (1 - mid);
          imp = d0 + y * (dw - d0);
        }
        w.lD[u] = (real)1 / m_max((real)1e-15, (1 - imp) / imp * m.dof_invw[d]);
        w.laref[u] = -m.B * (side * w.qvel[d]) - m.K * imp * dist;
      }
    }
  }
  LHW_SYNC();
  // ---------------- P8 qfrc_smooth (lane = dof) ; contact jacobians in the contact frame (n,t1,t2)=(+z,+y,-x)
  LHW_LANES(l) {
    if (l < NV) {
      const int lk = dof_link<NJ>(l);
      const real* S = w.S[l];
      real Ft[6];
      if (lk == 0)
        for (int c = 0; c < 6; c++) Ft[c] = w.F[0][c] + w.F[1][c] + w.F[1 + NJ][c];
      else
        for (int c = 0; c < 6; c++) Ft[c] = w.F[lk][c];
      const real bias = S[0] * Ft[0] + S[1] * Ft[1] + S[2] * Ft[2] + S[3] * Ft[3] + S[4] * Ft[4] + S[5] * Ft[5];
      real q = -m.damping[l] * w.qvel[l] - bias;
      if (l >= 6) q += w.ctrl[l - 6];
      w.qfs[l] = q;
      w.qacc[l] = w.qacc_warm[l];
    }
    for (int it = l; it < NCON * NA; it += 32) {
      const int s = it / NA, j = it - s * NA, f = s >> 2;
      if ((s & 3) < w.ncon[f]) {
        const real* S = w.S[loc2dof<NJ>(f, j)];
        real u[3];
        cross(S, w.cpos[s], u);
        w.Jc[s][0][j] = S[5] + u[2];
        w.Jc[s][1][j] = S[4] + u[1];
        w.Jc[s][2][j] = -(S[3] + u[0]);
      }
    }
  }
  LHW_SYNC();
  // ---------------- P9 reference accelerations of the 4 pyramid edges of each contact (lane = contact row)
  LHW_LANES(l) {
    if (l < NCON * 3) {
      const int s = l / 3, k = l - 3 * s, f = s >> 2;
      if ((s & 3) < w.ncon[f]) {
        real acc = 0;
        for (int j = 0; j < NA; j++) acc += w.Jc[s][k][j] * w.qvel[loc2dof<NJ>(f, j)];
        w.cu[s][k] = acc;
      }
    }
  }
  LHW_SYNC();
  LHW_LANES(l) {
    const int s = l >> 2, e = l & 3, f = s >> 2;
    if ((s & 3) < w.ncon[f]) {
      const real vel = w.cu[s][0] + ((e & 1) ? -m.mu : m.mu) * w.cu[s][1 + (e >> 1)];
      w.earef[l] = -m.B * vel - w.cKid[s];
    }
  }
  LHW_SYNC();

  // ---------------- P10 primal Newton on  1/2 (a-a_s)' M (a-a_s) + sum_r 1/2 D_r min(0, J_r a - aref_r)^2
  bool converged = false;
  for (int iter = 0; iter <= m.max_iter && !converged; iter++) {
    // (a) M a (lanes < NV) ; J_c a per contact row (lanes 8.. reuse all lanes with a strided loop)
    LHW_LANES(l) {
      if (l < NV) {
        real acc = 0;
        if (l < 6) {
          for (int j = 0; j < NV; j++) acc += w.M[l][j] * w.qacc[j];
        } else {
          const int cs = 6 + ((l - 6) / NJ) * NJ;
          for (int j = 0; j < 6; j++) acc += w.M[l][j] * w.qacc[j];
          for (int j = cs; j < cs + NJ; j++) acc += w.M[l][j] * w.qacc[j];
        }
        w.Ma[l] = acc;
      }
      for (int it = l; it < NCON * 3; it += 32) {
        const int s = it / 3, k = it - 3 * s, f = s >> 2;
        if ((s & 3) < w.ncon[f]) {
          real acc = 0;
          for (int j = 0; j < NA; j++) acc += w.Jc[s][k][j] * w.qacc[loc2dof<NJ>(f, j)];
          w.cu[s][k] = acc;
        }
      }
    }
    LHW_SYNC();
    // (b) edge residuals / forces (lane = edge) and limit rows
    LHW_LANES(l) {
      const int s = l >> 2, e = l & 3, f = s >> 2;
      real fe = 0, jar = 1;
      if ((s & 3) < w.ncon[f]) {
        jar = w.cu[s][0] + ((e & 1) ? -m.mu : m.mu) * w.cu[s][1 + (e >> 1)] - w.earef[l];
        fe = jar < 0 ? -w.cD[s] * jar : (real)0;
      }
      w.ejar[l] = jar; w.ef[l] = fe; w.eact[l] = jar < 0;
      if (l < NU) {
        real fl = 0, jl = 1;
        if (w.lside[l]) {
          jl = w.lside[l] * w.qacc[6 + l] - w.laref[l];
          fl = jl < 0 ? -w.lD[l] * jl : (real)0;
        }
        w.ljar[l] = jl; w.lf[l] = fl; w.lact[l] = jl < 0;
      }
    }
    LHW_SYNC();
    // (c) per contact: force in the contact frame and the 3x3 weight  W = sum_active D w w'
    LHW_LANES(l) {
      if (l < NCON) {
        const real* fe = w.ef + 4 * l;
        w.cF[l][0] = fe[0] + fe[1] + fe[2] + fe[3];
        w.cF[l][1] = m.mu * (fe[0] - fe[1]);
        w.cF[l][2] = m.mu * (fe[2] - fe[3]);
        const int* a = w.eact + 4 * l;
        const real D = w.cD[l], mu = m.mu;
        w.cW[l][0] = D * (a[0] + a[1] + a[2] + a[3]);
        w.cW[l][1] = D * mu * (a[0] - a[1]);
        w.cW[l][2] = D * mu * (a[2] - a[3]);
        w.cW[l][3] = D * mu * mu * (a[0] + a[1]);
        w.cW[l][4] = D * mu * mu * (a[2] + a[3]);
      }
    }
    LHW_SYNC();
    // (d) gradient (lane = dof) ; WJ = W Jc
    LHW_LANES(l) {
      if (l < NV) {
        real g = w.Ma[l] - w.qfs[l];
        for (int f = 0; f < 2; f++) {
          int j = -1;
          if (l < 6) j = l;
          else if (l >= 6 + f * NJ && l < 6 + (f + 1) * NJ) j = 6 + (l - 6 - f * NJ);
          if (j >= 0)
            for (int k = 0; k < w.ncon[f]; k++) {
              const int s = f * 4 + k;
              g -= w.Jc[s][0][j] * w.cF[s][0] + w.Jc[s][1][j] * w.cF[s][1] + w.Jc[s][2][j] * w.cF[s][2];
            }
        }
        if (l >= 6 && w.lside[l - 6]) g -= w.lside[l - 6] * w.lf[l - 6];
        w.grad[l] = g;
      }
      for (int it = l; it < NCON * NA; it += 32) {
        const int s = it / NA, j = it - s * NA, f = s >> 2;
        if ((s & 3) < w.ncon[f]) {
          const real j0 = w.Jc[s][0][j], j1 = w.Jc[s][1][j], j2 = w.Jc[s][2][j];
          const real* W = w.cW[s];
          w.WJ[s][0][j] = W[0] * j0 + W[1] * j1 + W[2] * j2;
          w.WJ[s][1][j] = W[1] * j0 + W[3] * j1;
          w.WJ[s][2][j] = W[2] * j0 + W[4] * j2;
        }
      }
    }
    LHW_SYNC();
    const real g2 = warp_sum<real>([&](int l) { return l < NV ? w.grad[l] * w.grad[l] : (real)0; });
    if (g2 < m.tol2 || iter == m.max_iter) { converged = true; break; }
    // (e) H = M + Jc' W Jc + diag(limit D) on the structurally non-zero lower triangle
    LHW_LANES(l) {
      for (int it = l; it < Model<real, NJ>::NT; it += 32) {
        const int i = m.h_i[it], j = m.h_j[it];
        real acc = w.M[i][j];
        if (i == j && i >= 6 && w.lside[i - 6] && w.lact[i - 6]) acc += w.lD[i - 6];
        const int f0 = i < 6 ? 0 : (i - 6) / NJ, f1 = i < 6 ? 1 : f0;
        for (int f = f0; f <= f1; f++) {
          const int li = i < 6 ? i : 6 + (i - 6 - f * NJ), lj = j < 6 ? j : 6 + (j - 6 - f * NJ);
          for (int k = 0; k < w.ncon[f]; k++) {
            const int s = f * 4 + k;
            acc += w.Jc[s][0][li] * w.WJ[s][0][lj] + w.Jc[s][1][li] * w.WJ[s][1][lj] + w.Jc[s][2][li] * w.WJ[s][2][lj];
          }
        }
        w.H[i][j] = acc;
      }
      if (l < NV) w.sdir[l] = -w.grad[l];
    }
    LHW_SYNC();
    arrow_factor<real, NJ>(w);
    arrow_solve<real, NJ>(w, w.sdir);
    // (f) exact line search along sdir: phi'(alpha) is continuous, piecewise linear and increasing
    LHW_LANES(l) {
      if (l < NV) {
        real acc = 0;
        if (l < 6) {
          for (int j = 0; j < NV; j++) acc += w.M[l][j] * w.sdir[j];
        } else {
          const int cs = 6 + ((l - 6) / NJ) * NJ;
          for (int j = 0; j < 6; j++) acc += w.M[l][j] * w.sdir[j];
          for (int j = cs; j < cs + NJ; j++) acc += w.M[l][j] * w.sdir[j];
        }
        w.Ms[l] = acc;
      }
      for (int it = l; it < NCON * 3; it += 32) {
        const int s = it / 3, k = it - 3 * s, f = s >> 2;
        if ((s & 3) < w.ncon[f]) {
          real acc = 0;
          for (int j = 0; j < NA; j++) acc += w.Jc[s][k][j] * w.sdir[loc2dof<NJ>(f, j)];
          w.cu[s][k] = acc;
        }
      }
    }
    LHW_SYNC();
    LHW_LANES(l) {
      const int s = l >> 2, e = l & 3, f = s >> 2;
      real jv = 0;
      if ((s & 3) < w.ncon[f]) jv = w.cu[s][0] + ((e & 1) ? -m.mu : m.mu) * w.cu[s][1 + (e >> 1)];
      w.ejv[l] = jv;
      if (l < NU) w.ljv[l] = w.lside[l] ? w.lside[l] * w.sdir[6 + l] : (real)0;
    }
    LHW_SYNC();
    const real sMs = warp_sum<real>([&](int l) { return l < NV ? w.sdir[l] * w.Ms[l] : (real)0; });
    const real sg = warp_sum<real>([&](int l) { return l < NV ? w.sdir[l] * (w.Ma[l] - w.qfs[l]) : (real)0; });
    // safeguarded 1-D Newton: exact on each linear piece of phi', bracketed by [lo, hi]
    const real LS_TOL = sizeof(real) == 8 ? (real)2e-14 : (real)1e-5;
    real alpha = 1, lo = 0, hi = -1, d_at0 = 0;
    for (int ls = 0; ls < 40; ls++) {
      const real al = ls == 0 ? (real)0 : alpha;
      const real cd = warp_sum<real>([&](int l) {
        const int s = l >> 2, f = s >> 2;
        real acc = 0;
        if ((s & 3) < w.ncon[f]) {
          const real x = w.ejar[l] + al * w.ejv[l];
          if (x < 0) acc += w.cD[s] * x * w.ejv[l];
        }
        if (l < NU && w.lside[l]) {
          const real x = w.ljar[l] + al * w.ljv[l];
          if (x < 0) acc += w.lD[l] * x * w.ljv[l];
        }
        return acc;
      });
      const real d = al * sMs + sg + cd;
      if (ls == 0) {
        d_at0 = d;
        if (!(d < 0)) { alpha = 0; break; }  // not a descent direction: at the roundoff floor
        continue;                            // first trial: the full Newton step alpha = 1
      }
      if (m_abs(d) <= LS_TOL * m_abs(d_at0)) break;
      const real cdd = warp_sum<real>([&](int l) {
        const int s = l >> 2, f = s >> 2;
        real acc = 0;
        if ((s & 3) < w.ncon[f] && w.ejar[l] + al * w.ejv[l] < 0) acc += w.cD[s] * w.ejv[l] * w.ejv[l];
        if (l < NU && w.lside[l] && w.ljar[l] + al * w.ljv[l] < 0) acc += w.lD[l] * w.ljv[l] * w.ljv[l];
        return acc;
      });
      if (d < 0) lo = al; else hi = al;
      real an = al - d / (sMs + cdd);
      if (an <= lo || (hi > 0 && an >= hi)) an = hi > 0 ? (real)0.5 * (lo + hi) : 2 * al;
      if (an == al) break;
      alpha = an;
    }
    LHW_LANES(l) {
      if (l < NV) w.qacc[l] += alpha * w.sdir[l];
      if (l == 31) w.iters_total++;
    }
    LHW_SYNC();
    if (alpha == 0) converged = true;
  }

  // ---------------- P11 what mjData keeps after mj_step (evaluated at the pre-integration state)
  LHW_LANES(l) {
    if (l < NU) {
      w.act_len[l] = w.qpos[7 + l];
      w.act_vel[l] = w.qvel[6 + l];
      w.act_force[l] = w.ctrl[l];
    }
    if (last) {
      if (l >= 12 && l < 15) { w.root_vlin[l - 12] = w.qvel[l - 12]; w.qacc_lag[l - 12] = w.qacc[l - 12]; }
      if (l >= 16 && l < 18) {
        const int f = l - 16, lk = (f + 1) * NJ;
        real t[3];
        cross(w.V[lk], w.xr[lk], t);  // velocity of the link origin: v_o + w x r
        real g = 0, zmin = 0;
        for (int c = 0; c < 3; c++) w.foot_vel[f][c] = w.V[lk][3 + c] + t[c];
        for (int k = 0; k < w.ncon[f]; k++) {
          const real* cf = w.cF[f * 4 + k];
          g += m_sqrt(cf[0] * cf[0] + cf[1] * cf[1] + cf[2] * cf[2]);
        }
        w.grf[f] = g;
        (void)zmin;
      }
      if (l == 20) {
        real z = 0;
        bool first = true;
        for (int s = 0; s < NCON; s++)
          if ((s & 3) < w.ncon[s >> 2]) {
            const real cz = w.o[2] + w.cpos[s][2];
            if (first || cz < z) z = cz;
            first = false;
          }
        w.cz_min = z;
      }
    }
    // rhs of the implicit-damping solve: qfrc_smooth + J' f = M a - grad
    if (l < NV) w.vec[l] = w.Ma[l] - w.grad[l];
  }
  LHW_SYNC();
  // ---------------- P12 mj_Euler: (M + h diag(damping)) a' = qfrc_smooth + qfrc_constraint ; integrate
  if (m.any_damping) {
    LHW_LANES(l) {
      for (int it = l; it < Model<real, NJ>::NT; it += 32) {
        const int i = m.h_i[it], j = m.h_j[it];
        w.H[i][j] = w.M[i][j] + (i == j ? m.h * m.damping[i] : (real)0);
      }
    }
    LHW_SYNC();
    arrow_factor<real, NJ>(w);
    arrow_solve<real, NJ>(w, w.vec);
  } else {
    LHW_LANES(l) {
      if (l < NV) w.vec[l] = w.qacc[l];
    }
    LHW_SYNC();
  }
  LHW_LANES(l) {
    if (l < NV) {
      const real a = w.vec[l];
      if (!(m_abs(a) < (real)1e10)) w.status |= 1;
      w.qacc_warm[l] = w.qacc[l];
      w.qvel[l] += m.h * a;
    }
  }
  LHW_SYNC();
  LHW_LANES(l) {
    if (l < 3) w.qpos[l] += m.h * w.qvel[l];
    else if (l == 3) {
      real* q = w.qpos + 3;
      real n = (real)1 / m_sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
      real q0 = q[0] * n, q1 = q[1] * n, q2 = q[2] * n, q3 = q[3] * n;
      const real* wv = w.qvel + 3;
      const real wn = m_sqrt(wv[0] * wv[0] + wv[1] * wv[1] + wv[2] * wv[2]);
      if (wn > (real)1e-15) {
        real sa, ca;
        m_sincos((real)0.5 * m.h * wn, &sa, &ca);
        sa /= wn;
        const real r0 = ca, r1 = wv[0] * sa, r2 = wv[1] * sa, r3 = wv[2] * sa;
        const real o0 = q0 * r0 - q1 * r1 - q2 * r2 - q3 * r3, o1 = q0 * r1 + q1 * r0 + q2 * r3 - q3 * r2;
        const real o2 = q0 * r2 - q1 * r3 + q2 * r0 + q3 * r1, o3 = q0 * r3 + q1 * r2 - q2 * r1 + q3 * r0;
        n = (real)1 / m_sqrt(o0 * o0 + o1 * o1 + o2 * o2 + o3 * o3);
        q0 = o0 * n; q1 = o1 * n; q2 = o2 * n; q3 = o3 * n;
      }
      q[0] = q0; q[1] = q1; q[2] = q2; q[3] = q3;
    } else if (l >= 4 && l < 4 + NU) {
      w.qpos[7 + l - 4] += m.h * w.qvel[6 + l - 4];
    }
  }
  LHW_SYNC();
}


// ================================================================= environment level (one control step)
// state record I/O: reals [qpos qvel qacc_warm act_len act_vel prev_pred prev_action prev_torque mode_ref ep_rew],
// ints [phase mode traj_len ep_len rng_ctr have_prev status pad]; env-major, one coalesced stream per warp
template <class real, int NJ>
LHW_DEV void load_state(Work<real, NJ>& w, const real* sr, const int32_t* si, uint32_t env_id) {
  constexpr int NR = Dims<real, NJ>::NSTATE_R;
  real* dst = w.qpos;  // persistent block is contiguous in Work, same order as the record
  LHW_LANES(l) {
    for (int it = l; it < NR; it += 32) dst[it] = sr[it];
    if (l == 0) {
      w.phase = si[0]; w.mode = si[1]; w.traj_len = si[2]; w.ep_len = si[3];
      w.rng_ctr = (uint32_t)si[4]; w.have_prev = si[5]; w.status = si[6];
      w.env_id = env_id;
      w.iters_total = 0;
    }
    for (int it = l; it < Work<real, NJ>::NV * Work<real, NJ>::NV; it += 32) { (&w.M[0][0])[it] = 0; (&w.H[0][0])[it] = 0; }
  }
  LHW_SYNC();
}
template <class real, int NJ> LHW_DEV void store_state(const Work<real, NJ>& w, real* sr, int32_t* si) {
  constexpr int NR = Dims<real, NJ>::NSTATE_R;
  const real* src = w.qpos;
  LHW_LANES(l) {
    for (int it = l; it < NR; it += 32) sr[it] = src[it];
    if (l == 0) {
      si[0] = w.phase; si[1] = w.mode; si[2] = w.traj_len; si[3] = w.ep_len;
      si[4] = (int32_t)w.rng_ctr; si[5] = w.have_prev; si[6] = w.status; si[7] = w.iters_total;
    }
  }
  LHW_SYNC();
}

// WalkModes.sample_ref (tasks/walking_task.py:33-40)
template <class real, int NJ> LHW_DEV void sample_ref(Work<real, NJ>& w, uint32_t seed, uint32_t stream) {
  uint32_t u[4];
  philox(seed, w.env_id, w.rng_ctr, stream, u);
  if (w.mode == STANDING) {
    for (int x = 0; x < 3; x++) w.mode_ref[x] = (real)-1 + (real)2 * u01<real>(u[x]);
  } else if (w.mode == INPLACE) {
    w.mode_ref[0] = (real)-0.5 + u01<real>(u[0]); w.mode_ref[1] = 0; w.mode_ref[2] = 0;
  } else {
    w.mode_ref[0] = 0; w.mode_ref[1] = (real)0.4 * u01<real>(u[0]); w.mode_ref[2] = 0;
  }
}

// observation (envs/jvrc/jvrc_base.py:133-145 + jvrc_walk.py:65-67): current qpos quat / qvel, LAGGED actuator state
template <class real, int NJ> LHW_DEV void env_obs(Work<real, NJ>& w, const Model<real, NJ>& m) {
  constexpr int NU = 2 * NJ;
  LHW_LANES(l) {
    if (l == 0) {  // transforms3d quat2euler (sxyz) roll, pitch via quat2mat
      const real qw = w.qpos[3], x = w.qpos[4], y = w.qpos[5], z = w.qpos[6];
      const real s = (real)2 / (qw * qw + x * x + y * y + z * z);
      const real X = x * s, Y = y * s, Z = z * s;
      const real wX = qw * X, wY = qw * Y, wZ = qw * Z, xX = x * X, xY = x * Y, xZ = x * Z, yY = y * Y, yZ = y * Z;
      const real M00 = 1 - (yY + z * Z), M10 = xY + wZ, M20 = xZ - wY, M21 = yZ + wX, M22 = 1 - (xX + yY);
      const real cy = m_sqrt(M00 * M00 + M10 * M10);
      if (cy > (real)(4 * 2.220446049250313e-16)) {
        w.obs[0] = m_atan2(M21, M22);
      } else {
        w.obs[0] = m_atan2(-(yZ - wX), 1 - (xX + z * Z));
      }
      w.obs[1] = m_atan2(-M20, cy);
    } else if (l < 4) {
      w.obs[1 + l] = w.qvel[2 + l];
    } else if (l == 4) {
      real sn, cs;
      m_sincos((real)(2 * M_PI) * w.phase / m.period, &sn, &cs);
      w.obs[5 + 2 * NU] = sn; w.obs[6 + 2 * NU] = cs;
    } else if (l == 5) {
      real* e = w.obs + 7 + 2 * NU;
      e[0] = w.mode == FORWARD; e[1] = w.mode == INPLACE; e[2] = w.mode == STANDING;
      e[3] = w.mode_ref[0]; e[4] = w.mode_ref[1]; e[5] = w.mode_ref[2];
    } else if (l >= 8 && l < 8 + NU) {
      w.obs[5 + l - 8] = w.act_len[l - 8];
      w.obs[5 + NU + l - 8] = w.act_vel[l - 8];
    }
  }
  LHW_SYNC();
}

// MujocoEnv.reset + BaseHumanoidEnv.reset_model + WalkingTask.reset
template <class real, int NJ> LHW_DEV void env_reset(Work<real, NJ>& w, const Model<real, NJ>& m, uint32_t seed) {
  constexpr int NV = 6 + 2 * NJ, NQ = NV + 1, NU = 2 * NJ;
  LHW_LANES(l) {
    if (l < NQ) w.qpos[l] = m.nominal[l];
    if (l < NV) { w.qvel[l] = 0; w.qacc_warm[l] = 0; }
    if (l < NU) { w.ctrl[l] = 0; w.prev_pred[l] = 0; }
  }
  LHW_SYNC();
  for (int i = 0; i < 3; i++) substep<real, NJ>(w, m, false);
  LHW_LANES(l) {
    if (l == 0) {
      w.rng_ctr++;
      uint32_t u[4];
      philox(seed, w.env_id, w.rng_ctr, 3, u);
      const real c = u01<real>(u[0]);
      w.mode = c < (real)0.6 ? STANDING : (c < (real)0.8 ? INPLACE : FORWARD);
      sample_ref<real, NJ>(w, seed, 4);
      w.phase = randint(u[1], m.period);
      w.traj_len = 0; w.ep_len = 0; w.ep_rew = 0; w.status = 0;
    }
  }
  LHW_SYNC();
  env_obs<real, NJ>(w, m);
}

struct StepOut {
  void* obs; void* term_obs; void* reward; void* rew_terms;
  int32_t* done; int32_t* ended; int32_t* ep_len; void* ep_rew;
};

// BaseHumanoidEnv.step + the RolloutWorker's bookkeeping (traj_len truncation, auto-reset, episode stats)
template <class real, int NJ>
LHW_DEV void env_step(Work<real, NJ>& w, const Model<real, NJ>& m, const real* action, uint32_t seed, int max_traj_len,
                      int autoreset, real* obs_out, real* term_obs_out, real* reward_out, real* rew_terms_out,
                      int32_t* done_out, int32_t* ended_out, int32_t* ep_len_out, real* ep_rew_out) {
  constexpr int NU = 2 * NJ, NOBS = Work<real, NJ>::NOBS;
  // action smoothing + nominal offsets (base_humanoid_env.py:209-212, robot_base.py:80-85)
  LHW_LANES(l) {
    if (l < NU) {
      const real t = m.smoothing * action[l] + (1 - m.smoothing) * w.prev_pred[l] + m.nominal[7 + l];
      w.target[l] = t;
      if (!w.have_prev) { w.prev_action[l] = t; w.prev_torque[l] = 0; }
    }
  }
  LHW_SYNC();
  for (int sidx = 0; sidx < m.frame_skip; sidx++) {
    LHW_LANES(l) {
      if (l < NU) w.ctrl[l] = m.kp[l] * (w.target[l] - w.act_len[l]) + m.kd[l] * ((real)0 - w.act_vel[l]);
    }
    LHW_SYNC();
    substep<real, NJ>(w, m, sidx == m.frame_skip - 1);
  }
  // WalkingTask.step (tasks/walking_task.py:149-179)
  LHW_LANES(l) {
    if (l == 0) {
      w.have_prev = 1;
      w.rng_ctr++;
      w.phase += 1;
      if (w.phase >= m.period) w.phase = 0;
      uint32_t u[4];
      philox(seed, w.env_id, w.rng_ctr, 0, u);
      const bool dbl = m.clock[0][w.phase] == (real)1 && m.clock[2][w.phase] == (real)1;
      if (randint(u[0], 100) == 0 && dbl) {
        if (w.mode == INPLACE) w.mode = STANDING;
        else if (w.mode == STANDING) w.mode = INPLACE;
        sample_ref<real, NJ>(w, seed, 1);
      }
      if (randint(u[1], 200) == 0 && w.mode != STANDING) {
        if (w.mode == FORWARD) w.mode = INPLACE;
        else if (w.mode == INPLACE) w.mode = FORWARD;
        sample_ref<real, NJ>(w, seed, 2);
      }
    }
  }
  LHW_SYNC();
  // WalkingTask.calc_reward (tasks/walking_task.py:85-147, tasks/rewards.py), lane = term
  LHW_LANES(l) {
    if (l < NREW) {
      real rfc = m.clock[0][w.phase], rvc = m.clock[1][w.phase], lfc = m.clock[2][w.phase], lvc = m.clock[3][w.phase];
      real yaw_ref = w.mode_ref[0], vx = w.mode_ref[1], vy = w.mode_ref[2];
      if (w.mode == STANDING) { rfc = lfc = 1; rvc = lvc = -1; yaw_ref = vx = vy = 0; }
      else if (w.mode == INPLACE) { vx = vy = 0; }
      else { yaw_ref = 0; }
      const real PI4 = (real)(M_PI / 4);
      real r = 0;
      if (l == 0) {
        const real nl = m_min(w.grf[1], m.fcap) / m.fcap * 2 - 1, nr = m_min(w.grf[0], m.fcap) / m.fcap * 2 - 1;
        r = (real)0.225 * ((m_tan(PI4 * lfc * nl) + m_tan(PI4 * rfc * nr)) / 2);
      } else if (l == 1) {
        const real lv = m_sqrt(dot3(w.foot_vel[1], w.foot_vel[1])), rv = m_sqrt(dot3(w.foot_vel[0], w.foot_vel[0]));
        const real vl = m_min(lv, (real)0.2) / (real)0.2 * 2 - 1, vr = m_min(rv, (real)0.2) / (real)0.2 * 2 - 1;
        r = (real)0.225 * ((m_tan(PI4 * lvc * vl) + m_tan(PI4 * rvc * vr)) / 2);
      } else if (l == 2) {
        real err = 0;
        for (int x = 0; x < 3; x++) err += m_abs(w.qvel[3 + x]);
        for (int x = 0; x < 3; x++) err += m_abs(w.qacc_lag[x]);
        r = (real)0.05 * m_exp((real)-0.25 * err);
      } else if (l == 3) {
        const real cz = (w.ncon[0] + w.ncon[1]) > 0 ? w.cz_min : (real)0;
        real herr = m_abs(w.o[2] - cz - m.goal_height);
        if (herr < (real)0.01 + (real)0.05 * m_sqrt(vx * vx + vy * vy)) herr = 0;
        r = (real)0.05 * m_exp(-40 * herr * herr);
      } else if (l == 4) {
        const real* R = w.xmat[0];
        const real* v = w.root_vlin;
        const real ex = R[0] * v[0] + R[3] * v[1] + R[6] * v[2] - vx, ey = R[1] * v[0] + R[4] * v[1] + R[7] * v[2] - vy;
        r = (real)0.15 * m_exp(-10 * (ex * ex + ey * ey));
      } else if (l == 5) {
        const real ye = m_abs(w.qvel[5] - yaw_ref);
        r = (real)0.15 * m_exp(-10 * ye * ye * ye);
      } else if (l == 6) {
        real hp[3];
        mv3(w.xmat[0], m.head, hp);
        r = (real)0.05 * m_exp(-10 * m_sqrt(hp[0] * hp[0] + hp[1] * hp[1]));
      } else if (l == 7) {
        real pe = 0;
        for (int u = 0; u < NU; u++) { const real d = m.nominal[7 + u] - w.act_len[u]; pe += d * d; }
        r = (real)0.05 * m_exp(-m_sqrt(pe));
      } else if (l == 8) {
        real te = 0;
        for (int u = 0; u < NU; u++) te += m_abs(w.prev_torque[u] - w.act_force[u]);
        r = (real)0.025 * m_exp((real)-0.25 * (te / NU));
      } else {
        real ae = 0;
        for (int u = 0; u < NU; u++) ae += m_abs(w.prev_action[u] - w.target[u]);
        r = (real)0.025 * m_exp(-5 * ae / NU);
      }
      w.rew[l] = r;
    }
  }
  LHW_SYNC();
  env_obs<real, NJ>(w, m);
  real total = 0;
  for (int i = 0; i < NREW; i++) total += w.rew[i];
  const int done = (w.qpos[2] < (real)0.6) || (w.qpos[2] > (real)1.4) || (w.status != 0);
  const int ended = done || (w.traj_len + 1 >= max_traj_len);
  LHW_SYNC();
  LHW_LANES(l) {
    if (l < NU) {
      w.prev_action[l] = w.target[l];
      w.prev_torque[l] = w.act_force[l];
      w.prev_pred[l] = action[l];
    }
    if (l == 12) { w.traj_len += 1; w.ep_len += 1; w.ep_rew += total; }
    if (l < NREW && rew_terms_out) rew_terms_out[l] = w.rew[l];
    if (l == 13) {
      reward_out[0] = total;
      done_out[0] = done;
      ended_out[0] = ended;
    }
    if (term_obs_out && (ended || !autoreset))
      for (int it = l; it < NOBS; it += 32) term_obs_out[it] = w.obs[it];
  }
  LHW_SYNC();
  if (ended && autoreset) {
    LHW_LANES(l) {
      if (l == 0) {
        if (ep_len_out) ep_len_out[0] = w.ep_len;
        if (ep_rew_out) ep_rew_out[0] = w.ep_rew;
      }
    }
    LHW_SYNC();
    env_reset<real, NJ>(w, m, seed);
  }
  LHW_LANES(l) {
    for (int it = l; it < NOBS; it += 32) obs_out[it] = w.obs[it];
  }
  LHW_SYNC();
}

}  // namespace lhw
