// sim_kernels.cu — sm_100a kernels + C-ABI for the batched environment (see include/lhw_b200.h).
//
// One environment per warp.  A warp streams its env's state record (env-major, contiguous -> coalesced
// 128 B lines) from HBM into its private slice of shared memory, runs frame_skip physics substeps + reward /
// observation / termination / auto-reset entirely on chip, and streams the record back: one HBM round trip
// per control step (1220 B algorithmic per env-step in fp32).  Inside an environment only __syncwarp()
// is used.  The shipped launch puts 8 (fp64) / 14 (fp32) environments in a block that meets at one __syncthreads() per
// physics substep: the warps then run the same region of a kernel that is far larger than the instruction cache
// (findings 2 and 6); LHW_WARPS_PER_BLOCK=1 gives the original one-warp blocks that never wait.
#include <cuda_runtime.h>
#include <stdio.h>

#include <atomic>
#include <string>
#include <vector>

#include "../../include/lhw_b200.h"
#define LHW_BLOCK_SYNC(on) do { if (on) __syncthreads(); } while (0)
#include "model_pack.h"

using namespace lhw;

namespace {

// robot / task variants (sim_core.h Cfg<NJ, TK>): <6,0> JVRC-1 walking, <5,0> Unitree H1 standing, <6,1> JVRC-1 stepping
constexpr int NJ_JVRC = 6, NJ_H1 = 5;
__constant__ Model<double, NJ_JVRC, 0> c_model_d;
__constant__ Model<float, NJ_JVRC, 0> c_model_f;
__constant__ Model<double, NJ_H1, 0> c_model_d5;
__constant__ Model<float, NJ_H1, 0> c_model_f5;
__constant__ Model<double, NJ_JVRC, 1> c_model_ds;
__constant__ Model<float, NJ_JVRC, 1> c_model_fs;
__constant__ Model<double, NJ_JVRC, 2> c_model_dt;
__constant__ Model<float, NJ_JVRC, 2> c_model_ft;

template <class real, int NJ, int TK> __device__ __forceinline__ const Model<real, NJ, TK>& cmodel();
template <> __device__ __forceinline__ const Model<double, NJ_JVRC, 0>& cmodel<double, NJ_JVRC, 0>() { return c_model_d; }
template <> __device__ __forceinline__ const Model<float, NJ_JVRC, 0>& cmodel<float, NJ_JVRC, 0>() { return c_model_f; }
template <> __device__ __forceinline__ const Model<double, NJ_H1, 0>& cmodel<double, NJ_H1, 0>() { return c_model_d5; }
template <> __device__ __forceinline__ const Model<float, NJ_H1, 0>& cmodel<float, NJ_H1, 0>() { return c_model_f5; }
template <> __device__ __forceinline__ const Model<double, NJ_JVRC, 1>& cmodel<double, NJ_JVRC, 1>() { return c_model_ds; }
template <> __device__ __forceinline__ const Model<float, NJ_JVRC, 1>& cmodel<float, NJ_JVRC, 1>() { return c_model_fs; }
template <> __device__ __forceinline__ const Model<double, NJ_JVRC, 2>& cmodel<double, NJ_JVRC, 2>() { return c_model_dt; }
template <> __device__ __forceinline__ const Model<float, NJ_JVRC, 2>& cmodel<float, NJ_JVRC, 2>() { return c_model_ft; }

}  // namespace
// tell sim_core.h's out-of-line routines where the model really lives (constant bank -> LDC with immediate offsets)
namespace lhw {
#define LHW_MODEL_HOME(R, J, T, OBJ)                                                                         \
  template <> struct ModelHome<R, J, T> {                                                                   \
    static __device__ __forceinline__ const Model<R, J, T>& get(const Model<R, J, T>&) { return OBJ; }        \
  };
LHW_MODEL_HOME(double, NJ_JVRC, 0, c_model_d)
LHW_MODEL_HOME(float, NJ_JVRC, 0, c_model_f)
LHW_MODEL_HOME(double, NJ_H1, 0, c_model_d5)
LHW_MODEL_HOME(float, NJ_H1, 0, c_model_f5)
LHW_MODEL_HOME(double, NJ_JVRC, 1, c_model_ds)
LHW_MODEL_HOME(float, NJ_JVRC, 1, c_model_fs)
LHW_MODEL_HOME(double, NJ_JVRC, 2, c_model_dt)
LHW_MODEL_HOME(float, NJ_JVRC, 2, c_model_ft)
#undef LHW_MODEL_HOME
}  // namespace lhw
namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};
const void* g_owner[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // which sim's model sits in each constant-memory slot

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CUDA_OK(call)                                                                       \
  do {                                                                                      \
    cudaError_t _e = (call);                                                                \
    if (_e != cudaSuccess) return fail(-10, std::string(#call) + ": " + cudaGetErrorString(_e)); \
  } while (0)

template <class real, int NJ, int TK>
__global__ void __launch_bounds__(32) reset_kernel(real* __restrict__ state_r, int32_t* __restrict__ state_i, int n_envs,
                                                    uint32_t seed, uint32_t first_id, const int32_t* __restrict__ mask,
                                                    int fresh, real* __restrict__ obs_out) {
  using W = Work<real, NJ, TK>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int env = blockIdx.x;
  if (env >= n_envs) return;
  if (mask && !mask[env]) return;
  W& w = *reinterpret_cast<W*>(smem_raw);
  const Model<real, NJ, TK>& m = cmodel<real, NJ, TK>();
  constexpr int NR = Dims<real, NJ, TK>::NSTATE_R;
  real* sr = state_r + (size_t)env * NR;
  int32_t* si = state_i + (size_t)env * NSTATE_I;
  if (fresh) {
    for (int it = lane; it < NR; it += 32) sr[it] = 0;
    if (lane < NSTATE_I) si[lane] = 0;
    __syncwarp();
  }
  load_state<real, NJ, TK>(w, sr, si, first_id + env);
  env_reset<real, NJ, TK>(w, m, seed);
  store_state<real, NJ, TK>(w, sr, si);
  if (obs_out)
    for (int it = lane; it < W::NOBS; it += 32) obs_out[(size_t)env * W::NOBS + it] = w.obs[it];
}

// ONE warp per block: the warp's Work struct then sits at a link-time-constant shared-memory address, so every
// access is [index + immediate] and no base register has to be kept (or rematerialised).  fp32: 28 blocks/SM
// (the whole 4096-env batch of BASELINE configs[1] is resident at once); fp64: 16 blocks/SM (shared-memory bound).
template <class real, int NJ, int TK>
__global__ void __launch_bounds__(32, sizeof(real) == 4 ? (TK ? 24 : 28) : (TK ? 12 : (NJ == NJ_JVRC ? 16 : 15)))
    step_kernel(real* __restrict__ state_r, int32_t* __restrict__ state_i, int n_envs, uint32_t seed, uint32_t first_id,
                const real* __restrict__ actions, int max_traj_len, int autoreset, real* __restrict__ obs,
                real* __restrict__ term_obs, real* __restrict__ reward, real* __restrict__ rew_terms,
                int32_t* __restrict__ done, int32_t* __restrict__ ended, int32_t* __restrict__ ep_len,
                real* __restrict__ ep_rew) {
  using W = Work<real, NJ, TK>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int env = blockIdx.x;
  if (env >= n_envs) return;
  W& w = *reinterpret_cast<W*>(smem_raw);
  const Model<real, NJ, TK>& m = cmodel<real, NJ, TK>();
  constexpr int NR = Dims<real, NJ, TK>::NSTATE_R, NU = 2 * NJ;
  real* sr = state_r + (size_t)env * NR;
  int32_t* si = state_i + (size_t)env * NSTATE_I;
  load_state<real, NJ, TK>(w, sr, si, first_id + env);
  env_step<real, NJ, TK>(w, m, actions + (size_t)env * NU, seed, max_traj_len, autoreset, 0, 1, obs + (size_t)env * W::NOBS,
                     term_obs ? term_obs + (size_t)env * W::NOBS : nullptr, reward + env,
                     rew_terms ? rew_terms + (size_t)env * NREW : nullptr, done + env, ended + env,
                     ep_len ? ep_len + env : nullptr, ep_rew ? ep_rew + env : nullptr);
  store_state<real, NJ, TK>(w, sr, si);
}

// the shipped carving: W warps per block (one env each), a __syncthreads() per substep so the warps of an SM share
// instruction-cache fills (profiles/: stall_no_inst 24 % -> 3 %); W from LHW_WARPS_PER_BLOCK or the measured defaults below
template <class real, int NJ, int TK>
// LHW_X_REGCAP (candidate, tools/ab_variants.py): the register budget of the fp64 lock-step kernel as "threads per SM" — 512 = 127
// registers (16 warps per SM, the shipped kernel), 576 = 112 (18 warps), 640 = 96 (20 warps)
#ifndef LHW_X_REGCAP
#define LHW_X_REGCAP 512
#endif
__global__ void __launch_bounds__(sizeof(real) == 8 ? LHW_X_REGCAP : 896, 1)
    step_kernel_mw(real* __restrict__ state_r, int32_t* __restrict__ state_i, int n_envs, uint32_t seed, uint32_t first_id,
                   const real* __restrict__ actions, int max_traj_len, int autoreset, real* __restrict__ obs,
                   real* __restrict__ term_obs, real* __restrict__ reward, real* __restrict__ rew_terms,
                   int32_t* __restrict__ done, int32_t* __restrict__ ended, int32_t* __restrict__ ep_len,
                   real* __restrict__ ep_rew, int sync_mode) {
  using W = Work<real, NJ, TK>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5;
  const int env = blockIdx.x * (blockDim.x >> 5) + warp;
  const int alive = env < n_envs;
  const int e = alive ? env : n_envs - 1;
  W& w = reinterpret_cast<W*>(smem_raw)[warp];
  const Model<real, NJ, TK>& m = cmodel<real, NJ, TK>();
  constexpr int NR = Dims<real, NJ, TK>::NSTATE_R, NU = 2 * NJ;
  real* sr = state_r + (size_t)e * NR;
  int32_t* si = state_i + (size_t)e * NSTATE_I;
  if (alive) load_state<real, NJ, TK>(w, sr, si, first_id + e);
  env_step<real, NJ, TK>(w, m, actions + (size_t)e * NU, seed, max_traj_len, autoreset, sync_mode, alive, obs + (size_t)e * W::NOBS,
                     term_obs ? term_obs + (size_t)e * W::NOBS : nullptr, reward + e,
                     rew_terms ? rew_terms + (size_t)e * NREW : nullptr, done + e, ended + e,
                     ep_len ? ep_len + e : nullptr, ep_rew ? ep_rew + e : nullptr);
  if (alive) store_state<real, NJ, TK>(w, sr, si);
}

}  // namespace

struct lhw_sim {
  int precision, device, warps_per_block, sync_mode, nj, tk;
  Model<double, NJ_JVRC, 0> md;
  Model<float, NJ_JVRC, 0> mf;
  Model<double, NJ_H1, 0> md5;
  Model<float, NJ_H1, 0> mf5;
  Model<double, NJ_JVRC, 1> mds;
  Model<float, NJ_JVRC, 1> mfs;
  Model<double, NJ_JVRC, 2> mdt;
  Model<float, NJ_JVRC, 2> mft;
  void* d_plans = nullptr;   // SteppingTask footstep plans in HBM ([MAXPLAN][PLAN_STRIDE] reals of the sim's precision)
  void* d_twin = nullptr;    // the active model record once more, in global memory (lane-indexed tables are read from there)
  size_t work_bytes;
  int state_reals, obs_dim;
};

namespace {

// copy the model to global memory as well and leave its address in the record that goes to the constant bank
template <class M> int upload_twin(lhw_sim* s, M& host, cudaStream_t st) {
  if (!s->d_twin) CUDA_OK(cudaMalloc(&s->d_twin, sizeof(M)));
  host.gm = (const M*)s->d_twin;
  CUDA_OK(cudaMemcpyAsync(s->d_twin, &host, sizeof(M), cudaMemcpyHostToDevice, st));
  return 0;
}

int upload_model(lhw_sim* s, cudaStream_t st) {
  const int slot = (s->tk == 2 ? 6 : s->tk ? 4 : (s->nj == NJ_JVRC ? 0 : 2)) + (s->precision == 64 ? 0 : 1);
  if (g_owner[slot] == s) return 0;
  {
    int rc = 0;
    if (s->tk == 2) rc = s->precision == 64 ? upload_twin(s, s->mdt, st) : upload_twin(s, s->mft, st);
    else if (s->tk) rc = s->precision == 64 ? upload_twin(s, s->mds, st) : upload_twin(s, s->mfs, st);
    else if (s->nj == NJ_JVRC) rc = s->precision == 64 ? upload_twin(s, s->md, st) : upload_twin(s, s->mf, st);
    else rc = s->precision == 64 ? upload_twin(s, s->md5, st) : upload_twin(s, s->mf5, st);
    if (rc != 0) return rc;
  }
  if (s->tk == 2) {
    if (s->precision == 64) CUDA_OK(cudaMemcpyToSymbolAsync(c_model_dt, &s->mdt, sizeof(s->mdt), 0, cudaMemcpyHostToDevice, st));
    else CUDA_OK(cudaMemcpyToSymbolAsync(c_model_ft, &s->mft, sizeof(s->mft), 0, cudaMemcpyHostToDevice, st));
  } else if (s->tk) {
    if (s->precision == 64) CUDA_OK(cudaMemcpyToSymbolAsync(c_model_ds, &s->mds, sizeof(s->mds), 0, cudaMemcpyHostToDevice, st));
    else CUDA_OK(cudaMemcpyToSymbolAsync(c_model_fs, &s->mfs, sizeof(s->mfs), 0, cudaMemcpyHostToDevice, st));
  } else if (s->nj == NJ_JVRC) {
    if (s->precision == 64) CUDA_OK(cudaMemcpyToSymbolAsync(c_model_d, &s->md, sizeof(s->md), 0, cudaMemcpyHostToDevice, st));
    else CUDA_OK(cudaMemcpyToSymbolAsync(c_model_f, &s->mf, sizeof(s->mf), 0, cudaMemcpyHostToDevice, st));
  } else {
    if (s->precision == 64) CUDA_OK(cudaMemcpyToSymbolAsync(c_model_d5, &s->md5, sizeof(s->md5), 0, cudaMemcpyHostToDevice, st));
    else CUDA_OK(cudaMemcpyToSymbolAsync(c_model_f5, &s->mf5, sizeof(s->mf5), 0, cudaMemcpyHostToDevice, st));
  }
  g_owner[slot] = s;
  return 0;
}

template <class K> int prepare_kernel(K kernel, size_t smem) {
  CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return 0;
}

template <class real, int NJ, int TK> int prepare_variant(lhw_sim* s) {
  s->work_bytes = sizeof(Work<real, NJ, TK>);
  s->state_reals = Dims<real, NJ, TK>::NSTATE_R;
  s->obs_dim = Dims<real, NJ, TK>::NOBS;
  int maxsmem = 0;
  CUDA_OK(cudaDeviceGetAttribute(&maxsmem, cudaDevAttrMaxSharedMemoryPerBlockOptin, s->device));
  // lock-step blocks: as many warps as the launch bound and the shared memory of one SM allow, two blocks per SM
  const int cap_threads = (sizeof(real) == 8 ? LHW_X_REGCAP : 896) / 32;
  if (s->warps_per_block > cap_threads) s->warps_per_block = cap_threads;
  while (s->warps_per_block > 1 && s->work_bytes * s->warps_per_block > (size_t)maxsmem) s->warps_per_block--;
  if (s->warps_per_block < 1) s->warps_per_block = 1;
  if ((int)s->work_bytes > maxsmem) return fail(-5, "working set does not fit in shared memory");
  if (prepare_kernel(step_kernel<real, NJ, TK>, s->work_bytes) || prepare_kernel(reset_kernel<real, NJ, TK>, s->work_bytes)) return -10;
  if (s->warps_per_block > 1 && prepare_kernel(step_kernel_mw<real, NJ, TK>, s->work_bytes * s->warps_per_block)) return -10;
  return 0;
}

template <class real, int NJ, int TK>
int launch_reset(lhw_sim* s, void* state_r, int32_t* state_i, int n_envs, uint32_t seed, uint32_t first_env_id,
                 const int32_t* mask, int fresh, void* obs, cudaStream_t st) {
  reset_kernel<real, NJ, TK><<<n_envs, 32, s->work_bytes, st>>>((real*)state_r, state_i, n_envs, seed, first_env_id, mask, fresh,
                                                           (real*)obs);
  return 0;
}

template <class real, int NJ, int TK>
int launch_step(lhw_sim* s, void* state_r, int32_t* state_i, int n_envs, uint32_t seed, uint32_t first_env_id,
                const void* actions, int max_traj_len, int autoreset, void* obs, void* term_obs, void* reward,
                void* rew_terms, int32_t* done, int32_t* ended, int32_t* ep_len, void* ep_rew, cudaStream_t st) {
  const int wpb = s->warps_per_block, grid = (n_envs + wpb - 1) / wpb;
  const size_t smem = s->work_bytes * wpb;
  if (wpb > 1)
    step_kernel_mw<real, NJ, TK><<<grid, wpb * 32, smem, st>>>((real*)state_r, state_i, n_envs, seed, first_env_id,
                                                          (const real*)actions, max_traj_len, autoreset, (real*)obs,
                                                          (real*)term_obs, (real*)reward, (real*)rew_terms, done, ended,
                                                          ep_len, (real*)ep_rew, s->sync_mode);
  else
    step_kernel<real, NJ, TK><<<grid, 32, smem, st>>>((real*)state_r, state_i, n_envs, seed, first_env_id,
                                                 (const real*)actions, max_traj_len, autoreset, (real*)obs, (real*)term_obs,
                                                 (real*)reward, (real*)rew_terms, done, ended, ep_len, (real*)ep_rew);
  return 0;
}

// (precision, NJ) dispatch
#define LHW_DISPATCH(s, FN, ...)                                                                                          \
  ((s)->tk == 2 ? ((s)->precision == 64 ? FN<double, NJ_JVRC, 2>(__VA_ARGS__) : FN<float, NJ_JVRC, 2>(__VA_ARGS__))          \
   : (s)->tk ? ((s)->precision == 64 ? FN<double, NJ_JVRC, 1>(__VA_ARGS__) : FN<float, NJ_JVRC, 1>(__VA_ARGS__))             \
   : (s)->nj == NJ_JVRC ? ((s)->precision == 64 ? FN<double, NJ_JVRC, 0>(__VA_ARGS__) : FN<float, NJ_JVRC, 0>(__VA_ARGS__)) \
                        : ((s)->precision == 64 ? FN<double, NJ_H1, 0>(__VA_ARGS__) : FN<float, NJ_H1, 0>(__VA_ARGS__)))

}  // namespace

extern "C" {

int lhw_version(void) { return 3; }
const char* lhw_last_error(void) { return g_err.c_str(); }
long long lhw_launch_count(void) { return g_launches.load(); }
void lhw_count_launch(void) { g_launches++; }

int lhw_sim_create(lhw_sim** out, const double* flat, int n_flat, int precision, int device) {
  if (!out || !flat || n_flat < 1) return fail(-1, "null argument");
  if (precision != 64 && precision != 32) return fail(-2, "precision must be 32 or 64");
  int ndev = 0;
  CUDA_OK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(-3, "no such CUDA device");
  CUDA_OK(cudaSetDevice(device));
  const int var = (int)flat[0], nj = var % 100, tk = var / 100;
  if (var != NJ_JVRC && var != NJ_H1 && var != 100 + NJ_JVRC && var != 200 + NJ_JVRC)
    return fail(-4, "unsupported robot / task variant " + std::to_string(var));
  lhw_sim* s = new lhw_sim();
  s->precision = precision;
  s->device = device;
  s->nj = nj;
  s->tk = tk;
  int rc = 0;
  if (tk == 2) {
    rc = fill_model(s->mdt, flat, n_flat);
    if (rc == 0) rc = fill_model(s->mft, flat, n_flat);
  } else if (tk) {
    // SteppingTask: the footstep plans go to HBM in the sim's precision; the model constants hold the device pointer
    const size_t words = (size_t)MAXPLAN * PLAN_STRIDE;
    std::vector<double> pd(words, 0.0);
    std::vector<float> pf(words, 0.0f);
    rc = fill_model(s->mds, flat, n_flat, pd.data());
    if (rc == 0) rc = fill_model(s->mfs, flat, n_flat, pf.data());
    if (rc == 0) {
      const size_t bytes = words * (precision == 64 ? sizeof(double) : sizeof(float));
      cudaError_t e = cudaMalloc(&s->d_plans, bytes);
      if (e == cudaSuccess) e = cudaMemcpy(s->d_plans, precision == 64 ? (const void*)pd.data() : (const void*)pf.data(), bytes, cudaMemcpyHostToDevice);
      if (e != cudaSuccess) { delete s; return fail(-10, std::string("footstep plan upload: ") + cudaGetErrorString(e)); }
      s->mds.plans = (const double*)s->d_plans;
      s->mfs.plans = (const float*)s->d_plans;
    }
  } else {
    rc = nj == NJ_JVRC ? fill_model(s->md, flat, n_flat) : fill_model(s->md5, flat, n_flat);
    if (rc == 0) rc = nj == NJ_JVRC ? fill_model(s->mf, flat, n_flat) : fill_model(s->mf5, flat, n_flat);
  }
  if (rc != 0) {
    delete s;
    return fail(-4, "malformed model array (fill_model rc " + std::to_string(rc) + ")");
  }
  const char* env_wpb = getenv("LHW_WARPS_PER_BLOCK");
  // measured on B200 (profiles/): lock-step blocks of 8 (fp64, 2 blocks/SM) / 14 (fp32, 2 blocks/SM) warps
  s->warps_per_block = env_wpb ? atoi(env_wpb) : (precision == 64 ? (nj == NJ_JVRC && !tk ? 8 : 7) : 14);
  const char* env_sync = getenv("LHW_BLOCK_SYNC_MODE");
  s->sync_mode = env_sync ? atoi(env_sync) : 1;
  if (s->warps_per_block < 1) s->warps_per_block = 1;
  rc = LHW_DISPATCH(s, prepare_variant, s);
  if (rc != 0) { delete s; return rc; }
  *out = s;
  return 0;
}

int lhw_sim_destroy(lhw_sim* s) {
  if (!s) return 0;
  for (int k = 0; k < 8; k++)
    if (g_owner[k] == s) g_owner[k] = nullptr;
  if (s->d_plans) cudaFree(s->d_plans);
  if (s->d_twin) cudaFree(s->d_twin);
  delete s;
  return 0;
}

int lhw_sim_set_step_height(lhw_sim* s, double h) {
  if (!s) return fail(-1, "null argument");
  if (s->tk != 1) return 0;   // only the SteppingTask has a curriculum
  s->mds.step_height = h;
  s->mfs.step_height = (float)h;
  for (int k = 4; k < 6; k++)
    if (g_owner[k] == s) g_owner[k] = nullptr;   // re-upload the constants at the next launch / lhw_sim_bind
  return 0;
}

int lhw_sim_state_reals(const lhw_sim* s) { return s ? s->state_reals : Dims<double, NJ_JVRC, 0>::NSTATE_R; }
int lhw_sim_state_ints(const lhw_sim*) { return NSTATE_I; }
int lhw_sim_obs_dim(const lhw_sim* s) { return s ? s->obs_dim : Dims<double, NJ_JVRC, 0>::NOBS; }
int lhw_sim_act_dim(const lhw_sim* s) { return 2 * (s ? s->nj : NJ_JVRC); }
int lhw_sim_smem_bytes_per_env(const lhw_sim* s) { return (int)s->work_bytes; }
int lhw_sim_precision(const lhw_sim* s) { return s ? s->precision : 0; }
int lhw_sim_device(const lhw_sim* s) { return s ? s->device : -1; }

int lhw_sim_bind(lhw_sim* s, void* stream) {
  if (!s) return fail(-1, "null argument");
  return upload_model(s, (cudaStream_t)stream);
}

int lhw_sim_reset(lhw_sim* s, void* state_r, int32_t* state_i, int n_envs, uint32_t seed, uint32_t first_env_id,
                  const int32_t* mask, int fresh, void* obs, void* stream) {
  if (!s || !state_r || !state_i) return fail(-1, "null argument");
  if (n_envs <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (upload_model(s, st)) return -10;
  LHW_DISPATCH(s, launch_reset, s, state_r, state_i, n_envs, seed, first_env_id, mask, fresh, obs, st);
  g_launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}

int lhw_sim_step(lhw_sim* s, void* state_r, int32_t* state_i, int n_envs, uint32_t seed, uint32_t first_env_id,
                 const void* actions, int max_traj_len, int autoreset, void* obs, void* term_obs, void* reward,
                 void* rew_terms, int32_t* done, int32_t* ended, int32_t* ep_len, void* ep_rew, void* stream) {
  if (!s || !state_r || !state_i || !actions || !obs || !reward || !done || !ended) return fail(-1, "null argument");
  if (n_envs <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (upload_model(s, st)) return -10;
  LHW_DISPATCH(s, launch_step, s, state_r, state_i, n_envs, seed, first_env_id, actions, max_traj_len, autoreset, obs,
               term_obs, reward, rew_terms, done, ended, ep_len, ep_rew, st);
  g_launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
