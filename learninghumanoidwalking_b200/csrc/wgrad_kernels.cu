// Weight / bias gradient of a Linear layer over a long batch: gW[N,K] = gy[M,N]^T x[M,K], gb[N] = sum_m gy[m,:].
//
// Replaces, inside `loss.backward()` of rl/algos/ppo.py:389-392, the two ops autograd runs per Linear of the 2 x 256 MLPs
// (rl/policies/actor.py:122-189, critic.py): mm(gy^T, x) — a GEMM whose OUTPUT is tiny (256 x 256, 256 x 37, 12 x 256, 1 x 256)
// and whose reduction dimension is the minibatch (21 845 - 43 690 rows) — and sum(gy, 0).  The library runs the first as a
// 64 x 64-tile split-K SIMT kernel plus an epilogue launch and the second as a separate reduction that reads gy again
// (profiles/r02_ppo_kernels.md: 0.8 ms + 0.2 ms of a 1.9 ms update at 21 845 samples).  Here: one launch streams gy and x ONCE
// through shared memory, every CTA owns one output tile for one slice of the batch (tiles x slices ~ 2 CTAs per SM), 8 x 8
// register micro-tiles, double-buffered 8-row chunks, the column sums of gy ride along in the CTAs of the first k-tile; a
// second launch adds the slices in index order (fixed summation order: run-to-run deterministic, like every reduction of
// this build).  fp32 FFMA throughout — the reference trains with allow_tf32 = False, so no tensor-core shortcut here.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lhw_b200.h"

namespace {

constexpr int BM = 8;          // batch rows per shared-memory stage
constexpr int THREADS = 256;   // 16 x 16 threads, thread (ty, tx) owns TN x TK outputs

// column owned by element e (< T) of thread t: T >= 4 -> groups of 4 columns, group g at g * 64 + t * 4 (16-byte shared-memory
// reads at a 16-byte stride between neighbouring threads: conflict-free); T < 4 -> t * T + e
template <int T> __device__ __forceinline__ int frag_col(int t, int e) {
  if constexpr (T >= 4) return (e >> 2) * 64 + t * 4 + (e & 3);
  else return t * T + e;
}

template <int T> __device__ __forceinline__ void load_frag(const float* __restrict__ row, int t, float (&f)[T]) {
  if constexpr (T >= 4) {
#pragma unroll
    for (int g = 0; g < T / 4; g++) {
      const float4 v = *reinterpret_cast<const float4*>(row + g * 64 + t * 4);
      f[4 * g + 0] = v.x; f[4 * g + 1] = v.y; f[4 * g + 2] = v.z; f[4 * g + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int e = 0; e < T; e++) f[e] = row[t * T + e];
  }
}

// one stage (BM rows x BC columns starting at column c0) of a row-major [M, C] matrix, as float4 per thread; rows >= m_end and
// columns >= C read as zero
template <int BC> struct Stage {
  static constexpr int V4 = BM * BC / 4;                       // float4 slots of a stage
  static constexpr int PER = (V4 + THREADS - 1) / THREADS;     // per thread
  float4 r[PER];
  __device__ __forceinline__ void load(const float* __restrict__ P, int C, int m_base, int m_end, int c0, bool vec, int tid) {
#pragma unroll
    for (int i = 0; i < PER; i++) {
      const int v = tid + i * THREADS;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (v < V4) {
        const int row = v / (BC / 4), c = c0 + (v % (BC / 4)) * 4, m = m_base + row;
        if (m < m_end && c < C) {
          const float* p = P + (size_t)m * C + c;
          if (vec && c + 3 < C) x = *reinterpret_cast<const float4*>(p);
          else {
            x.x = p[0];
            if (c + 1 < C) x.y = p[1];
            if (c + 2 < C) x.z = p[2];
            if (c + 3 < C) x.w = p[3];
          }
        }
      }
      r[i] = x;
    }
  }
  __device__ __forceinline__ void store(float* __restrict__ S, int tid) const {   // S: [BM][BC]
#pragma unroll
    for (int i = 0; i < PER; i++) {
      const int v = tid + i * THREADS;
      if (v < V4) *reinterpret_cast<float4*>(S + 4 * v) = r[i];
    }
  }
};

template <int BN, int BK>
__global__ void __launch_bounds__(THREADS) wgrad_partial_kernel(const float* __restrict__ A /* gy [M, N] */, const float* __restrict__ B /* x [M, K] */,
                                                                int M, int N, int K, int rows_per_split, int tiles_k, int vecA, int vecB,
                                                                float* __restrict__ ws /* [S, N, K] */, float* __restrict__ wsb /* [S, N] or null */) {
  constexpr int TN = BN / 16, TK = BK / 16;
  __shared__ __align__(16) float As[2][BM][BN];
  __shared__ __align__(16) float Bs[2][BM][BK];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int tn = blockIdx.x / tiles_k, tk = blockIdx.x - tn * tiles_k;
  const int n0 = tn * BN, k0 = tk * BK;
  const int s = blockIdx.y;
  const int m0 = s * rows_per_split, m1 = min(M, m0 + rows_per_split);
  float acc[TN][TK];
#pragma unroll
  for (int i = 0; i < TN; i++)
#pragma unroll
    for (int j = 0; j < TK; j++) acc[i][j] = 0.f;
  float bsum = 0.f;
  const bool want_b = wsb != nullptr && tk == 0 && tid < BN;
  Stage<BN> ra;
  Stage<BK> rb;
  ra.load(A, N, m0, m1, n0, vecA != 0, tid);
  rb.load(B, K, m0, m1, k0, vecB != 0, tid);
  ra.store(&As[0][0][0], tid);
  rb.store(&Bs[0][0][0], tid);
  __syncthreads();
  int buf = 0;
  for (int m = m0; m < m1; m += BM) {
    const bool more = m + BM < m1;
    if (more) {      // next stage: global -> registers while this one is consumed
      ra.load(A, N, m + BM, m1, n0, vecA != 0, tid);
      rb.load(B, K, m + BM, m1, k0, vecB != 0, tid);
    }
#pragma unroll
    for (int r = 0; r < BM; r++) {
      float fa[TN], fb[TK];
      load_frag<TN>(&As[buf][r][0], ty, fa);
      load_frag<TK>(&Bs[buf][r][0], tx, fb);
#pragma unroll
      for (int i = 0; i < TN; i++)
#pragma unroll
        for (int j = 0; j < TK; j++) acc[i][j] = fmaf(fa[i], fb[j], acc[i][j]);
    }
    if (want_b) {
#pragma unroll
      for (int r = 0; r < BM; r++) bsum += As[buf][r][tid];
    }
    if (more) {
      ra.store(&As[buf ^ 1][0][0], tid);
      rb.store(&Bs[buf ^ 1][0][0], tid);
    }
    __syncthreads();
    buf ^= 1;
  }
  float* out = ws + (size_t)s * N * K;
#pragma unroll
  for (int i = 0; i < TN; i++) {
    const int n = n0 + frag_col<TN>(ty, i);
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < TK; j++) {
      const int k = k0 + frag_col<TK>(tx, j);
      if (k < K) out[(size_t)n * K + k] = acc[i][j];
    }
  }
  if (want_b && n0 + tid < N) wsb[(size_t)s * N + n0 + tid] = bsum;
}

// slices added in index order -> deterministic
__global__ void wgrad_reduce_kernel(const float* __restrict__ ws, const float* __restrict__ wsb, int S, int NK, int N, float* __restrict__ gW,
                                    float* __restrict__ gb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < NK) {
    float a = 0.f;
#pragma unroll 8
    for (int s = 0; s < S; s++) a += ws[(size_t)s * NK + i];
    gW[i] = a;
  }
  if (gb != nullptr && i < N) {
    float b = 0.f;
    for (int s = 0; s < S; s++) b += wsb[(size_t)s * N + i];
    gb[i] = b;
  }
}

struct Plan { int bn, bk, tiles_n, tiles_k, S, rows; };

Plan plan_for(int M, int N, int K) {
  Plan p;
  p.bn = N > 64 ? 128 : (N > 16 ? 64 : 16);
  p.bk = K > 64 ? 128 : 64;
  p.tiles_n = (N + p.bn - 1) / p.bn;
  p.tiles_k = (K + p.bk - 1) / p.bk;
  const int tiles = p.tiles_n * p.tiles_k;
  int S = (2 * 148 + tiles - 1) / tiles;                // ~ two CTAs per SM
  const int max_s = (M + 8 * BM - 1) / (8 * BM);        // at least 8 stages per slice
  if (S > max_s) S = max_s;
  if (S < 1) S = 1;
  int rows = (M + S - 1) / S;
  rows = (rows + BM - 1) / BM * BM;
  p.rows = rows;
  p.S = (M + rows - 1) / rows;
  return p;
}

#define KCHECK(name)                               \
  do {                                             \
    cudaError_t e_ = cudaGetLastError();           \
    if (e_ != cudaSuccess) return (int)e_;         \
  } while (0)

template <int BN, int BK>
int launch(const Plan& p, const float* gy, const float* x, int M, int N, int K, float* ws, float* wsb, cudaStream_t st) {
  const int vecA = (N % 4 == 0) && (((uintptr_t)gy & 15) == 0), vecB = (K % 4 == 0) && (((uintptr_t)x & 15) == 0);
  wgrad_partial_kernel<BN, BK><<<dim3(p.tiles_n * p.tiles_k, p.S), THREADS, 0, st>>>(gy, x, M, N, K, p.rows, p.tiles_k, vecA, vecB, ws, wsb);
  KCHECK("wgrad_partial_kernel");
  return 0;
}

}  // namespace

extern "C" {

long long lhw_linear_wgrad_workspace_floats(int M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  const Plan p = plan_for(M, N, K);
  return (long long)p.S * N * K + (long long)p.S * N;
}

int lhw_linear_wgrad(const float* gy, const float* x, int M, int N, int K, float* gW, float* gb_or_null, float* workspace,
                     void* stream) {
  if (N <= 0 || K <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (M <= 0) {      // empty batch: the gradients are zero
    cudaError_t e = cudaMemsetAsync(gW, 0, sizeof(float) * (size_t)N * K, st);
    if (e == cudaSuccess && gb_or_null) e = cudaMemsetAsync(gb_or_null, 0, sizeof(float) * (size_t)N, st);
    return (int)e;
  }
  if (!gy || !x || !gW || !workspace) return (int)cudaErrorInvalidValue;
  const Plan p = plan_for(M, N, K);
  float* ws = workspace;
  float* wsb = gb_or_null ? workspace + (size_t)p.S * N * K : nullptr;
  int rc;
  if (p.bn == 128 && p.bk == 128) rc = launch<128, 128>(p, gy, x, M, N, K, ws, wsb, st);
  else if (p.bn == 128) rc = launch<128, 64>(p, gy, x, M, N, K, ws, wsb, st);
  else if (p.bn == 64 && p.bk == 128) rc = launch<64, 128>(p, gy, x, M, N, K, ws, wsb, st);
  else if (p.bn == 64) rc = launch<64, 64>(p, gy, x, M, N, K, ws, wsb, st);
  else if (p.bk == 128) rc = launch<16, 128>(p, gy, x, M, N, K, ws, wsb, st);
  else rc = launch<16, 64>(p, gy, x, M, N, K, ws, wsb, st);
  if (rc) return rc;
  const int NK = N * K;
  wgrad_reduce_kernel<<<(NK + 255) / 256, 256, 0, st>>>(ws, wsb, p.S, NK, N, gW, gb_or_null);
  KCHECK("wgrad_reduce_kernel");
  return 0;
}

}  // extern "C"
