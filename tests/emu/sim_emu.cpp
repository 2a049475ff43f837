// tests/emu/sim_emu.cpp — TEST HARNESS ONLY.  Compiles the product's per-environment kernel source
// (learninghumanoidwalking_b200/csrc/sim_core.h) with LHW_CPU_EMU so the 32 lanes of a warp run sequentially
// on the host.  This lets the `-m "not gpu"` test tier check the exact arithmetic the CUDA kernel executes
// against oracle/ without a GPU.  It is never linked into, loaded by, or reachable from the product.
#define LHW_CPU_EMU 1
#include "../../learninghumanoidwalking_b200/csrc/model_pack.h"

#include <stdlib.h>

using namespace lhw;

template <class real, int NJ> struct Emu {
  Model<real, NJ> model;
  Work<real, NJ> work;
};

template <class real, int NJ> static void* create(const double* flat, int n) {
  auto* e = new Emu<real, NJ>();
  if (fill_model(e->model, flat, n) != 0) { delete e; return nullptr; }
  return e;
}

template <class real, int NJ>
static void reset_all(void* h, real* sr, int32_t* si, int n_envs, uint32_t seed, uint32_t first_id, real* obs, int fresh) {
  auto* e = (Emu<real, NJ>*)h;
  constexpr int NR = Dims<real, NJ>::NSTATE_R, NOBS = Work<real, NJ>::NOBS;
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

template <class real, int NJ>
static void step_all(void* h, real* sr, int32_t* si, int n_envs, uint32_t seed, uint32_t first_id, const real* actions,
                     int max_traj_len, int autoreset, real* obs, real* term_obs, real* reward, real* rew_terms,
                     int32_t* done, int32_t* ended, int32_t* ep_len, real* ep_rew) {
  auto* e = (Emu<real, NJ>*)h;
  constexpr int NR = Dims<real, NJ>::NSTATE_R, NOBS = Work<real, NJ>::NOBS, NU = 2 * NJ;
  for (int i = 0; i < n_envs; i++) {
    load_state(e->work, sr + (size_t)i * NR, si + (size_t)i * NSTATE_I, first_id + i);
    env_step(e->work, e->model, actions + (size_t)i * NU, seed, max_traj_len, autoreset, 0, 1, obs + (size_t)i * NOBS,
             term_obs + (size_t)i * NOBS, reward + i, rew_terms + (size_t)i * NREW, done + i, ended + i, ep_len + i,
             ep_rew + i);
    store_state(e->work, sr + (size_t)i * NR, si + (size_t)i * NSTATE_I);
  }
}

// physics substeps on a raw state record with explicit ctrl (for mj_step-level parity)
template <int NJ> static void substeps64(void* p, double* sr, int32_t* si, const double* ctrl, int nsteps) {
  auto* e = (Emu<double, NJ>*)p;
  load_state(e->work, sr, si, 0);
  for (int k = 0; k < 2 * NJ; k++) e->work.ctrl[k] = ctrl[k];
  for (int s = 0; s < nsteps; s++) substep<double, NJ>(e->work, e->model, true);
  store_state(e->work, sr, si);
}

// dispatch on (precision, NJ): NJ = 6 JVRC-1 walking, NJ = 5 Unitree H1 standing
struct Handle { int nj; void* p; };
#define DISPATCH(h, prec, CALL)                                         \
  do {                                                                  \
    if ((h)->nj == 6) { if ((prec) == 64) { CALL(double, 6); } else { CALL(float, 6); } } \
    else { if ((prec) == 64) { CALL(double, 5); } else { CALL(float, 5); } }              \
  } while (0)

extern "C" {
void* emu_create(const double* flat, int n, int precision) {
  const int nj = (int)flat[0];
  if (nj != 6 && nj != 5) return nullptr;
  void* p = nullptr;
  if (nj == 6) p = precision == 64 ? create<double, 6>(flat, n) : create<float, 6>(flat, n);
  else p = precision == 64 ? create<double, 5>(flat, n) : create<float, 5>(flat, n);
  if (!p) return nullptr;
  return new Handle{nj, p};
}
int emu_state_words(void* hv) { return ((Handle*)hv)->nj == 6 ? Dims<double, 6>::NSTATE_R : Dims<double, 5>::NSTATE_R; }
int emu_obs_dim(void* hv) { return ((Handle*)hv)->nj == 6 ? Dims<double, 6>::NOBS : Dims<double, 5>::NOBS; }
int emu_work_bytes(void* hv, int precision) {
  Handle* h = (Handle*)hv;
  if (h->nj == 6) return precision == 64 ? (int)sizeof(Work<double, 6>) : (int)sizeof(Work<float, 6>);
  return precision == 64 ? (int)sizeof(Work<double, 5>) : (int)sizeof(Work<float, 5>);
}
void emu_reset(void* hv, int precision, void* sr, int32_t* si, int n, uint32_t seed, uint32_t first_id, void* obs, int fresh) {
  Handle* h = (Handle*)hv;
#define CALL(R, J) reset_all<R, J>(h->p, (R*)sr, si, n, seed, first_id, (R*)obs, fresh)
  DISPATCH(h, precision, CALL);
#undef CALL
}
void emu_step(void* hv, int precision, void* sr, int32_t* si, int n, uint32_t seed, uint32_t first_id,
              const void* actions, int max_traj_len, int autoreset, void* obs, void* term_obs, void* reward,
              void* rew_terms, int32_t* done, int32_t* ended, int32_t* ep_len, void* ep_rew) {
  Handle* h = (Handle*)hv;
#define CALL(R, J)                                                                                             \
  step_all<R, J>(h->p, (R*)sr, si, n, seed, first_id, (const R*)actions, max_traj_len, autoreset, (R*)obs,       \
                 (R*)term_obs, (R*)reward, (R*)rew_terms, done, ended, ep_len, (R*)ep_rew)
  DISPATCH(h, precision, CALL);
#undef CALL
}
void emu_substep64(void* hv, double* sr, int32_t* si, const double* ctrl, int nsteps) {
  Handle* h = (Handle*)hv;
  if (h->nj == 6) substeps64<6>(h->p, sr, si, ctrl, nsteps);
  else substeps64<5>(h->p, sr, si, ctrl, nsteps);
}
}
