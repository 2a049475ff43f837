// tests/emu/sim_emu.cpp — TEST HARNESS ONLY.  Compiles the product's per-environment kernel source
// (learninghumanoidwalking_b200/csrc/sim_core.h) with LHW_CPU_EMU so the 32 lanes of a warp run sequentially
// on the host.  This lets the `-m "not gpu"` test tier check the exact arithmetic the CUDA kernel executes
// against oracle/ without a GPU.  It is never linked into, loaded by, or reachable from the product.
#define LHW_CPU_EMU 1
#include "../../learninghumanoidwalking_b200/csrc/model_pack.h"

#include <stdlib.h>

using namespace lhw;

template <class real, int NJ, int TK> struct Emu {
  Model<real, NJ, TK> model;
  Work<real, NJ, TK> work;
  real plans[MAXPLAN * PLAN_STRIDE];   // SteppingTask footstep plans (host memory here, HBM in the product)
};

template <class real, int NJ, int TK> static void* create(const double* flat, int n) {
  auto* e = new Emu<real, NJ, TK>();
  if (fill_model(e->model, flat, n, e->plans) != 0) { delete e; return nullptr; }
  return e;
}

template <class real, int NJ, int TK>
static void reset_all(void* h, real* sr, int32_t* si, int n_envs, uint32_t seed, uint32_t first_id, real* obs, int fresh) {
  auto* e = (Emu<real, NJ, TK>*)h;
  constexpr int NR = Dims<real, NJ, TK>::NSTATE_R, NOBS = Work<real, NJ, TK>::NOBS;
  for (int i = 0; i < n_envs; i++) {
    if (fresh) {
      for (int k = 0; k < NR; k++) sr[(size_t)i * NR + k] = 0;
      for (int k = 0; k < NSTATE_I; k++) si[(size_t)i * NSTATE_I + k] = 0;
    }
    load_state(e->work, sr + (size_t)i * NR, si + (size_t)i * NSTATE_I, first_id + i);
    env_reset(e->work, e->model, seed);
    store_state(e->work, sr + (size_t)i * NR, si + (size_t)i * NSTATE_I);
    for (int k = 0; k < NOBS; k++) obs[(size_t)i * NOBS + k] = e->work.obs[k];
  }
}

template <class real, int NJ, int TK>
static void step_all(void* h, real* sr, int32_t* si, int n_envs, uint32_t seed, uint32_t first_id, const real* actions,
                     int max_traj_len, int autoreset, real* obs, real* term_obs, real* reward, real* rew_terms,
                     int32_t* done, int32_t* ended, int32_t* ep_len, real* ep_rew) {
  auto* e = (Emu<real, NJ, TK>*)h;
  constexpr int NR = Dims<real, NJ, TK>::NSTATE_R, NOBS = Work<real, NJ, TK>::NOBS, NU = 2 * NJ;
  for (int i = 0; i < n_envs; i++) {
    load_state(e->work, sr + (size_t)i * NR, si + (size_t)i * NSTATE_I, first_id + i);
    env_step(e->work, e->model, actions + (size_t)i * NU, seed, max_traj_len, autoreset, 0, 1, obs + (size_t)i * NOBS,
             term_obs + (size_t)i * NOBS, reward + i, rew_terms + (size_t)i * NREW, done + i, ended + i, ep_len + i,
             ep_rew + i);
    store_state(e->work, sr + (size_t)i * NR, si + (size_t)i * NSTATE_I);
  }
}

// physics substeps on a raw state record with explicit ctrl (for mj_step-level parity)
template <int NJ, int TK> static void substeps64(void* p, double* sr, int32_t* si, const double* ctrl, int nsteps) {
  auto* e = (Emu<double, NJ, TK>*)p;
  load_state(e->work, sr, si, 0);
  for (int k = 0; k < 2 * NJ; k++) e->work.ctrl[k] = ctrl[k];
  for (int s = 0; s < nsteps; s++) substep<double, NJ, TK>(e->work, e->model, true);
  store_state(e->work, sr, si);
}

// dispatch on (precision, variant): 6 = JVRC-1 walking, 5 = Unitree H1 standing, 106 = JVRC-1 stepping
struct Handle { int var; void* p; };
#define DISPATCH(h, prec, CALL)                                                                       \
  do {                                                                                                \
    if ((h)->var == 6) { if ((prec) == 64) { CALL(double, 6, 0); } else { CALL(float, 6, 0); } }      \
    else if ((h)->var == 106) { if ((prec) == 64) { CALL(double, 6, 1); } else { CALL(float, 6, 1); } } \
    else if ((h)->var == 206) { if ((prec) == 64) { CALL(double, 6, 2); } else { CALL(float, 6, 2); } } \
    else { if ((prec) == 64) { CALL(double, 5, 0); } else { CALL(float, 5, 0); } }                    \
  } while (0)

extern "C" {
void* emu_create(const double* flat, int n, int precision) {
  const int var = (int)flat[0];
  if (var != 6 && var != 5 && var != 106 && var != 206) return nullptr;
  Handle tmp{var, nullptr};
  void* p = nullptr;
#define CALL(R, J, T) p = create<R, J, T>(flat, n)
  DISPATCH(&tmp, precision, CALL);
#undef CALL
  if (!p) return nullptr;
  return new Handle{var, p};
}
int emu_state_words(void* hv) {
  const int v = ((Handle*)hv)->var;
  return v == 6 ? Dims<double, 6, 0>::NSTATE_R : v == 106 ? Dims<double, 6, 1>::NSTATE_R : v == 206 ? Dims<double, 6, 2>::NSTATE_R : Dims<double, 5, 0>::NSTATE_R;
}
int emu_obs_dim(void* hv) {
  const int v = ((Handle*)hv)->var;
  return v == 6 ? Dims<double, 6, 0>::NOBS : v == 106 ? Dims<double, 6, 1>::NOBS : v == 206 ? Dims<double, 6, 2>::NOBS : Dims<double, 5, 0>::NOBS;
}
int emu_work_bytes(void* hv, int precision) {
  Handle* h = (Handle*)hv;
  int out = 0;
#define CALL(R, J, T) out = (int)sizeof(Work<R, J, T>)
  DISPATCH(h, precision, CALL);
#undef CALL
  return out;
}
void emu_reset(void* hv, int precision, void* sr, int32_t* si, int n, uint32_t seed, uint32_t first_id, void* obs, int fresh) {
  Handle* h = (Handle*)hv;
#define CALL(R, J, T) reset_all<R, J, T>(h->p, (R*)sr, si, n, seed, first_id, (R*)obs, fresh)
  DISPATCH(h, precision, CALL);
#undef CALL
}
void emu_step(void* hv, int precision, void* sr, int32_t* si, int n, uint32_t seed, uint32_t first_id,
              const void* actions, int max_traj_len, int autoreset, void* obs, void* term_obs, void* reward,
              void* rew_terms, int32_t* done, int32_t* ended, int32_t* ep_len, void* ep_rew) {
  Handle* h = (Handle*)hv;
#define CALL(R, J, T)                                                                                          \
  step_all<R, J, T>(h->p, (R*)sr, si, n, seed, first_id, (const R*)actions, max_traj_len, autoreset, (R*)obs,    \
                    (R*)term_obs, (R*)reward, (R*)rew_terms, done, ended, ep_len, (R*)ep_rew)
  DISPATCH(h, precision, CALL);
#undef CALL
}
void emu_substep64(void* hv, double* sr, int32_t* si, const double* ctrl, int nsteps) {
  Handle* h = (Handle*)hv;
  if (h->var == 6) substeps64<6, 0>(h->p, sr, si, ctrl, nsteps);
  else if (h->var == 106) substeps64<6, 1>(h->p, sr, si, ctrl, nsteps);
  else if (h->var == 206) substeps64<6, 2>(h->p, sr, si, ctrl, nsteps);
  else substeps64<5, 0>(h->p, sr, si, ctrl, nsteps);
}
void emu_set_step_height(void* hv, int precision, double h_) {
  Handle* h = (Handle*)hv;
  if (h->var != 106) return;
  if (precision == 64) ((Emu<double, 6, 1>*)h->p)->model.step_height = h_;
  else ((Emu<float, 6, 1>*)h->p)->model.step_height = (float)h_;
}
}
