// gemm_simt.cu — fp32 FFMA GEMM for the small / skinny contractions of the model (FC stack with
// M = batch, the K = 24..128 EdgeConv projections, weight gradients of the small layers) and the
// exact-fp32 reference for the tcgen05 TF32 kernel (gemm_tcgen05.cu) that owns the large ones.
//
//   C[M,N] (+)= op(A)[M,K] * op(B)[K,N] (+ bias[N]),  row-major, leading dimensions in elements.
//   transa = 0: A stored [M,K] (lda >= K);  transa = 1: A stored [K,M] (lda >= M)
//   transb = 0: B stored [K,N] (ldb >= N);  transb = 1: B stored [N,K] (ldb >= K)
// Replaces the cuBLAS/cuDNN calls behind tf.matmul / tf.nn.conv2d(1x1) at the call sites
// reference utils/tf_util.py:161 and :349 (and their autodiff gradients).
//
// Tiling: CTA 128x64, K-step 16, 256 threads, 8x4 outputs per thread; operands staged in shared
// memory K-major so the inner product reads three LDS.128 per 32 FFMA.  When the tile grid cannot
// fill the GPU and K is long (weight gradients: K = B*N rows) the K range is split across
// blockIdx.z and partial tiles are combined with fp32 atomics into a zero-filled C.
#include "common.cuh"

namespace caae {

constexpr int GBM = 128, GBN = 64, GBK = 16, GTHREADS = 256;

template <bool TA, bool TB>
__global__ void __launch_bounds__(GTHREADS)
gemm_simt_kernel(int M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                 float* __restrict__ C, int ldc, const float* __restrict__ bias, int accumulate, int k_per_split,
                 int use_atomics) {
  pdl_wait();
  __shared__ __align__(16) float As[GBK][GBM + 4];
  __shared__ __align__(16) float Bs[GBK][GBN + 4];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(K, kbeg + k_per_split);
  const int ty = tid / 16, tx = tid % 16;  // 16 x 16 threads, 8 x 4 outputs each

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // Software pipeline: the global loads of K-step s+1 are issued into registers before the FFMA block
  // of step s, so their latency overlaps the arithmetic (the skinny FC GEMMs run one or two CTAs per
  // SM with a handful of K-steps each: nothing else would hide it).
  float ra[(GBM * GBK) / GTHREADS], rb[(GBK * GBN) / GTHREADS];
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < (GBM * GBK) / GTHREADS; ++i) {
      const int e = tid + i * GTHREADS;
      int m, k;
      if (TA) { m = e % GBM; k = e / GBM; }   // A stored [K,M]: m fastest -> coalesced
      else    { k = e % GBK; m = e / GBK; }   // A stored [M,K]: k fastest
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < kend) v = TA ? __ldg(A + (size_t)gk * lda + gm) : __ldg(A + (size_t)gm * lda + gk);
      ra[i] = v;
    }
#pragma unroll
    for (int i = 0; i < (GBK * GBN) / GTHREADS; ++i) {
      const int e = tid + i * GTHREADS;
      int n, k;
      if (TB) { k = e % GBK; n = e / GBK; }   // B stored [N,K]: k fastest
      else    { n = e % GBN; k = e / GBN; }   // B stored [K,N]: n fastest -> coalesced
      const int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < N && gk < kend) v = TB ? __ldg(B + (size_t)gn * ldb + gk) : __ldg(B + (size_t)gk * ldb + gn);
      rb[i] = v;
    }
  };
  auto sstore = [&]() {
#pragma unroll
    for (int i = 0; i < (GBM * GBK) / GTHREADS; ++i) {
      const int e = tid + i * GTHREADS;
      if (TA) As[e / GBM][e % GBM] = ra[i]; else As[e % GBK][e / GBK] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < (GBK * GBN) / GTHREADS; ++i) {
      const int e = tid + i * GTHREADS;
      if (TB) Bs[e % GBK][e / GBK] = rb[i]; else Bs[e / GBN][e % GBN] = rb[i];
    }
  };
  if (kbeg < kend) gload(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += GBK) {
    sstore();
    __syncthreads();
    if (k0 + GBK < kend) gload(k0 + GBK);
#pragma unroll
    for (int k = 0; k < GBK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  const bool add_bias = (bias != nullptr) && (blockIdx.z == 0);
  // split-K partial tiles: one 16-byte vector reduction per four outputs (red.global.add.v4.f32, sm_90+)
  // instead of four scalar atomics — the L2 atomic units were the bottleneck of the skinny FC GEMMs
  const bool vec_ok = use_atomics && ((reinterpret_cast<uintptr_t>(C) & 15) == 0) && (ldc % 4 == 0) &&
                      (n0 + tx * 4 + 3 < N);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gm = m0 + ty * 8 + i;
    if (gm >= M) continue;
    if (vec_ok) {
      const int gn = n0 + tx * 4;
      float4 v = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      if (add_bias) { v.x += __ldg(bias + gn); v.y += __ldg(bias + gn + 1); v.z += __ldg(bias + gn + 2); v.w += __ldg(bias + gn + 3); }
      atomicAdd(reinterpret_cast<float4*>(C + (size_t)gm * ldc + gn), v);
      continue;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j] + (add_bias ? __ldg(bias + gn) : 0.f);
      float* dst = C + (size_t)gm * ldc + gn;
      if (use_atomics) atomicAdd(dst, v);
      else *dst = accumulate ? (*dst + v) : v;
    }
  }
}

__global__ void zero_matrix_kernel(int M, int N, float* __restrict__ C, int ldc) {
  pdl_wait();
  const long total = (long)M * N;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x)
    C[(e / N) * ldc + (e % N)] = 0.f;
}

}  // namespace caae

using namespace caae;

extern "C" int caae_gemm_f32(int transa, int transb, int M, int N, int K, const float* A, int lda, const float* B,
                             int ldb, float* C, int ldc, const float* bias, int accumulate, caae_stream_t stream) {
  CAAE_RETURN_IF(M < 0 || N < 0 || K < 0, CAAE_E_BADSHAPE);
  if (M == 0 || N == 0) return CAAE_OK;
  CAAE_RETURN_IF(!C || (K > 0 && (!A || !B)), CAAE_E_NULLPTR);
  CAAE_RETURN_IF(ldc < N || lda < (transa ? M : K) || ldb < (transb ? K : N), CAAE_E_BADSHAPE);
  cudaStream_t s = as_stream(stream);
  const int tiles = ((M + GBM - 1) / GBM) * ((N + GBN - 1) / GBN);
  int splits = 1;
  if (tiles < kNumSMs && K >= 256) {
    splits = (2 * kNumSMs + tiles - 1) / tiles;
    const int max_splits = K / 64;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
  }
  int k_per_split = (K + splits - 1) / splits;
  k_per_split = ((k_per_split + GBK - 1) / GBK) * GBK;
  if (k_per_split <= 0) k_per_split = GBK;
  splits = K > 0 ? (K + k_per_split - 1) / k_per_split : 1;
  const int use_atomics = splits > 1;
  if (use_atomics && !accumulate) {
    long total = (long)M * N;
    int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
    caae::launch(zero_matrix_kernel, blocks, 256, 0, s, M, N, C, ldc);
  }
  dim3 grid((N + GBN - 1) / GBN, (M + GBM - 1) / GBM, splits);
  CAAE_RETURN_IF(grid.y > 65535 || grid.z > 65535, CAAE_E_BADSHAPE);
#define LAUNCH(TA, TB) caae::launch(gemm_simt_kernel<TA, TB>, grid, GTHREADS, 0, s, M, N, K, A, lda, B, ldb, C, ldc, bias, accumulate, k_per_split, use_atomics)
  if (transa) { if (transb) LAUNCH(true, true); else LAUNCH(true, false); }
  else        { if (transb) LAUNCH(false, true); else LAUNCH(false, false); }
#undef LAUNCH
  return CAAE_LAUNCH_STATUS();
}
