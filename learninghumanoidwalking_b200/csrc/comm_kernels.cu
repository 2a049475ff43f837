// comm_kernels.cu — the one exchange step of the path, fused: gradient all-reduce over NVLink peer memory +
// clip_grad_norm_ (actor and critic separately) + Adam, in ONE cooperative kernel per optimiser step.
//
// Replaces rl/algos/ppo.py:389-396 on N > 1 GPUs: (NCCL all-reduce) -> clip_grad_norm_ x2 -> Adam.step x2.
// Every rank owns a cudaMalloc'ed, CUDA-IPC exported block {flat gradient | flags}; peers map it once
// (cudaIpcOpenMemHandle).  The kernel then
//   1. handshakes with all peers through monotonically increasing epoch flags written into the peers' memory
//      (st.release.sys over NVLink) -> every rank's backward pass is complete and visible,
//   2. pulls the peers' gradient slices with plain P2P loads (one-shot all-reduce: each GPU reads W x 617 KB), sums them
//      in fixed rank order (bit-identical result on every rank -> replicas never drift) into a local buffer,
//   3. reduces sum(g^2) per network with fixed-order per-block partials (grid.sync, no atomics -> deterministic),
//   4. applies clip coefficient + Adam to the flat parameter buffer,
//   5. handshakes again so that no rank zeroes its gradient while a peer is still reading it.
// At W = 1 steps 1, 2 and 5 vanish and it is simply the fused clip+Adam of both networks in one launch.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/lhw_b200.h"

namespace cg = cooperative_groups;
extern "C" void lhw_count_launch(void);

namespace {

constexpr int MAX_WORLD = 16;
constexpr int COMM_BLOCK = 256;
thread_local std::string g_cerr;

struct Peers {
  const float* grad[MAX_WORLD];   // peer r's flat gradient (device pointer valid on this GPU)
  unsigned int* flags[MAX_WORLD];  // peer r's flag array: flags[r][me] is written by me
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// all ranks arrive: write `epoch` into every peer's flags[me], wait until every peer wrote it into mine
__device__ void peer_barrier(const Peers& P, unsigned int* my_flags, int rank, int world, unsigned int epoch, cg::grid_group& grid) {
  grid.sync();  // everything this GPU did before is complete
  if (blockIdx.x == 0 && threadIdx.x < world) {
    const int r = threadIdx.x;
    __threadfence_system();
    st_release_sys(P.flags[r] + rank, epoch);
    // bounded spin: a peer that never arrives must not hang the GPU (the host sees the error word instead)
    long long spins = 0;
    while ((int)(ld_acquire_sys(my_flags + r) - epoch) < 0)
      if (++spins > (1ll << 31)) { my_flags[MAX_WORLD - 1] = 0xDEADu; break; }
  }
  grid.sync();
}

__global__ void __launch_bounds__(COMM_BLOCK)
    fused_allreduce_clip_adam_kernel(Peers P, unsigned int* my_flags, float* __restrict__ reduced, double* __restrict__ partials,
                                     float* __restrict__ param, float* __restrict__ m, float* __restrict__ v, long long n_actor,
                                     long long n_total, int rank, int world, unsigned int epoch, float lr, float b1, float b2,
                                     float eps, float max_norm, float bc1, float bc2_sqrt) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sh[2][COMM_BLOCK / 32];
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  if (world > 1) peer_barrier(P, my_flags, rank, world, epoch, grid);
  // ---- one-shot all-reduce (fixed rank order) + per-network sum of squares of the AVERAGED gradient
  const float inv_w = 1.0f / (float)world;
  double sa = 0, sc = 0;
  for (long long i = tid; i < n_total; i += stride) {
    float g = 0;
    for (int r = 0; r < world; r++) g += P.grad[r][i];
    g *= inv_w;
    reduced[i] = g;
    const double gg = (double)g * (double)g;
    if (i < n_actor) sa += gg; else sc += gg;
  }
  for (int o = 16; o; o >>= 1) { sa += __shfl_xor_sync(0xffffffffu, sa, o); sc += __shfl_xor_sync(0xffffffffu, sc, o); }
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = sa; sh[1][threadIdx.x >> 5] = sc; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, c = 0;
    for (int k = 0; k < COMM_BLOCK / 32; k++) { a += sh[0][k]; c += sh[1][k]; }
    partials[2 * blockIdx.x] = a;
    partials[2 * blockIdx.x + 1] = c;
  }
  grid.sync();
  double na = 0, nc = 0;
  for (int b = 0; b < (int)gridDim.x; b++) { na += partials[2 * b]; nc += partials[2 * b + 1]; }  // same order on every rank
  float coef_a = max_norm / ((float)sqrt(na) + 1e-6f), coef_c = max_norm / ((float)sqrt(nc) + 1e-6f);
  coef_a = coef_a > 1.0f ? 1.0f : coef_a;
  coef_c = coef_c > 1.0f ? 1.0f : coef_c;
  if (tid == 0) { partials[2 * gridDim.x] = sqrt(na); partials[2 * gridDim.x + 1] = sqrt(nc); }
  // ---- clip + Adam (rl/algos/ppo.py:393-396)
  const float step_size = lr / bc1;
  for (long long i = tid; i < n_total; i += stride) {
    const float gi = reduced[i] * (i < n_actor ? coef_a : coef_c);
    const float mi = m[i] + (1.0f - b1) * (gi - m[i]);
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    param[i] -= step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
  }
  if (world > 1) peer_barrier(P, my_flags, rank, world, epoch + 1, grid);
}

}  // namespace

struct lhw_comm {
  int rank, world, device, grid;
  long long n;
  unsigned char* block;  // [n floats grad | MAX_WORLD flags]
  float* grad;
  unsigned int* flags;
  float* reduced;
  double* partials;
  Peers peers;
  void* opened[MAX_WORLD];
  unsigned int epoch;
};

namespace {
int cfail(int code, const std::string& s) { g_cerr = s; return code; }
#define COK(call)                                                                              \
  do {                                                                                         \
    cudaError_t _e = (call);                                                                   \
    if (_e != cudaSuccess) return cfail(-10, std::string(#call) + ": " + cudaGetErrorString(_e)); \
  } while (0)
size_t grad_bytes(long long n) { return ((size_t)n * 4 + 255) / 256 * 256; }
}  // namespace

extern "C" {

const char* lhw_comm_last_error(void) { return g_cerr.c_str(); }
int lhw_comm_handle_size(void) { return (int)sizeof(cudaIpcMemHandle_t); }

int lhw_comm_create(lhw_comm** out, long long n_floats, int rank, int world, int device) {
  if (!out || n_floats <= 0 || world < 1 || world > MAX_WORLD || rank < 0 || rank >= world) return cfail(-1, "bad argument");
  COK(cudaSetDevice(device));
  lhw_comm* c = new lhw_comm();
  memset(c, 0, sizeof(*c));
  c->rank = rank; c->world = world; c->device = device; c->n = n_floats;
  const size_t gb = grad_bytes(n_floats);
  COK(cudaMalloc(&c->block, gb + MAX_WORLD * sizeof(unsigned int)));
  COK(cudaMemset(c->block, 0, gb + MAX_WORLD * sizeof(unsigned int)));
  c->grad = (float*)c->block;
  c->flags = (unsigned int*)(c->block + gb);
  COK(cudaMalloc(&c->reduced, gb));
  int nsm = 0, per_sm = 0;
  COK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
  COK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fused_allreduce_clip_adam_kernel, COMM_BLOCK, 0));
  c->grid = nsm * (per_sm > 0 ? 1 : 0);
  if (c->grid <= 0) return cfail(-11, "cooperative kernel does not fit");
  long long need = (n_floats + COMM_BLOCK - 1) / COMM_BLOCK;
  if (need < c->grid) c->grid = (int)need;
  COK(cudaMalloc(&c->partials, (2 * c->grid + 2) * sizeof(double)));
  c->peers.grad[rank] = c->grad;
  c->peers.flags[rank] = c->flags;
  c->epoch = 1;
  *out = c;
  return 0;
}

void* lhw_comm_grad_ptr(lhw_comm* c) { return c ? c->grad : nullptr; }

int lhw_comm_export(lhw_comm* c, void* blob_host) {
  if (!c || !blob_host) return cfail(-1, "null argument");
  cudaIpcMemHandle_t h;
  COK(cudaIpcGetMemHandle(&h, c->block));
  memcpy(blob_host, &h, sizeof(h));
  return 0;
}

int lhw_comm_import(lhw_comm* c, const void* all_blobs_host) {
  if (!c || !all_blobs_host) return cfail(-1, "null argument");
  COK(cudaSetDevice(c->device));
  const size_t gb = grad_bytes(c->n);
  for (int r = 0; r < c->world; r++) {
    if (r == c->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const unsigned char*)all_blobs_host + (size_t)r * sizeof(h), sizeof(h));
    void* p = nullptr;
    COK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->opened[r] = p;
    c->peers.grad[r] = (const float*)p;
    c->peers.flags[r] = (unsigned int*)((unsigned char*)p + gb);
  }
  return 0;
}

int lhw_comm_destroy(lhw_comm* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  for (int r = 0; r < c->world; r++)
    if (c->opened[r]) cudaIpcCloseMemHandle(c->opened[r]);
  cudaFree(c->block);
  cudaFree(c->reduced);
  cudaFree(c->partials);
  delete c;
  return 0;
}

int lhw_fused_allreduce_clip_adam(lhw_comm* c, float* param, float* exp_avg, float* exp_avg_sq, long long n_actor,
                                  long long n_total, int step, float lr, float beta1, float beta2, float eps, float max_norm,
                                  float* norms_out_host_or_null, void* stream) {
  if (!c || !param || !exp_avg || !exp_avg_sq) return cfail(-1, "null argument");
  if (n_total != c->n || n_actor < 0 || n_actor > n_total) return cfail(-2, "size mismatch");
  float bc1 = 1.0f - powf(beta1, (float)step), bc2_sqrt = sqrtf(1.0f - powf(beta2, (float)step));
  int rank = c->rank, world = c->world;
  unsigned int epoch = c->epoch;
  c->epoch += 2;
  void* args[] = {&c->peers, &c->flags, &c->reduced, &c->partials, &param, &exp_avg, &exp_avg_sq, &n_actor, &n_total,
                  &rank, &world, &epoch, &lr, &beta1, &beta2, &eps, &max_norm, &bc1, &bc2_sqrt};
  COK(cudaLaunchCooperativeKernel((void*)fused_allreduce_clip_adam_kernel, dim3(c->grid), dim3(COMM_BLOCK), args, 0,
                                  (cudaStream_t)stream));
  lhw_count_launch();
  if (norms_out_host_or_null) {
    double h[2];
    COK(cudaMemcpyAsync(h, c->partials + 2 * c->grid, sizeof(h), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    COK(cudaStreamSynchronize((cudaStream_t)stream));
    norms_out_host_or_null[0] = (float)h[0];
    norms_out_host_or_null[1] = (float)h[1];
  }
  return 0;
}

}  // extern "C"
