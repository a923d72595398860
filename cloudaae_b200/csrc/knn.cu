// knn.cu — fused pairwise distance + top-k neighbour selection (never materialises [B,N,N]).
//
// Replaces pairwise_xyz_distance + knn (reference utils/tf_util.py:597-632: batched matmul into a
// [B,N,N] tensor, then tf.nn.top_k).  Same metric — D(i,j) = (|a_i|^2 + (-2 a_i.a_j)) + |a_j|^2 in
// fp32 FFMA arithmetic (not TF32: neighbour sets must be stable), evaluated on the features centred on
// the cloud mean, a = x - mu (pairwise distances do not depend on mu; fp32 cancellation does) — and the
// same selection: the k smallest, ascending, ties to the lower index, the point itself included.
//
// Layout.  x is [B*N, ldx] row-major; the first `c` channels of a row are the feature (c = 3 for
// layer 1, 64 for layers 2-4, read in place from the 320-wide concat buffer).
// Mapping.  A CTA owns 64 query rows of one cloud; warp w owns 8 of them; candidates stream
// through shared memory in chunks of 256, stored channel-major so lane l reads its 8 candidates
// (256*chunk + 8l ..) as two LDS.128 and the 8 query values as two broadcast LDS.128: 64 FFMA per
// 4 LDS.  Selection is k rounds of "lane-local min, REDUX min over the warp, REDUX min over the
// matching indices", merged across chunks through a k-entry list per row.
// Selection (per row, the whole warp): a threshold T with at least k candidates <= T is found from the
// 32 lane-local minima (REDUX rounds), the few candidates <= T (typically k..k+4) are compacted into one
// per lane through a 32-slot staging row, ranked by counting (lexicographic (distance, index), so ties
// go to the lower index) and the ranks < k are written out — ~2.5x fewer instructions than k rounds of
// arg-min over all 9 slots of every lane, which remains as the fallback when more than 32 survive.
#include <stdlib.h>

#include "common.cuh"

namespace caae {

constexpr int KNN_THREADS = 256;
constexpr int KNN_QROWS = 64;     // query rows per CTA
constexpr int KNN_CHUNK = 256;    // candidates per shared-memory chunk
constexpr int KNN_XT_LD = KNN_CHUNK + 4;
constexpr int KNN_QT_LD = KNN_QROWS + 4;
constexpr int KNN_MAXK = 32;

__device__ __forceinline__ uint32_t sortable(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(KNN_THREADS)
knn_kernel(int n, int c, int k, const float* __restrict__ x, int ldx, int* __restrict__ idx_out,
           const int* __restrict__ only) {
  pdl_wait();
  extern __shared__ __align__(16) float smem[];
  float* XT = smem;                               // [c][KNN_XT_LD]
  float* QT = XT + (size_t)c * KNN_XT_LD;         // [c][KNN_QT_LD]
  float* sqc = QT + (size_t)c * KNN_QT_LD;        // [KNN_CHUNK]
  float* sqq = sqc + KNN_CHUNK;                   // [KNN_QROWS]
  uint32_t* lkey = reinterpret_cast<uint32_t*>(sqq + KNN_QROWS);  // [KNN_QROWS][KNN_MAXK]
  int* lidx = reinterpret_cast<int*>(lkey + KNN_QROWS * KNN_MAXK);
  uint32_t* stage_k = reinterpret_cast<uint32_t*>(lidx + KNN_QROWS * KNN_MAXK);  // [warps][32]
  int* stage_i = reinterpret_cast<int*>(stage_k + (KNN_THREADS / 32) * 32);
  float* mu = reinterpret_cast<float*>(stage_i + (KNN_THREADS / 32) * 32);       // [c] cloud mean per channel
  float* mu_part = mu + c;                                                      // [4][c]

  const int cloud = blockIdx.y;
  const int q0 = blockIdx.x * KNN_QROWS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* __restrict__ xc = x + (size_t)cloud * n * ldx;
  if (only != nullptr && only[cloud] == 0) return;   // routed launch: only the clouds flagged by caae_knn_classify
  // ---- the cloud's mean.  Distances are evaluated on CENTRED features x - mu: the matmul form |a|^2 - 2 a.b + |b|^2
  // cancels catastrophically in fp32 when the features share a large common component (post-ReLU activations: norms^2
  // ~1e2, neighbour distances ~1e-3) — the centred form is the same quantity with ~1e4 x less rounding noise, i.e.
  // closer to the reference's exact-arithmetic meaning (float64 oracle agreement 0.990 -> 0.9999 on layer 4).
  // Fixed summation order (knn_tc_kernel uses the same): four interleaved partial sums over ascending rows.
  for (int e = tid; e < 4 * c; e += KNN_THREADS) {
    const int ch = e % c, part = e / c;
    float sm = 0.f;
    for (int r = part; r < n; r += 4) sm = __fadd_rn(sm, __ldg(xc + (size_t)r * ldx + ch));
    mu_part[part * c + ch] = sm;
  }
  __syncthreads();
  for (int ch = tid; ch < c; ch += KNN_THREADS)
    mu[ch] = __fdiv_rn(__fadd_rn(__fadd_rn(mu_part[ch], mu_part[c + ch]), __fadd_rn(mu_part[2 * c + ch], mu_part[3 * c + ch])), (float)n);
  __syncthreads();

  // ---- stage the query rows (channel-major) and their squared norms
  // Staging transposes [row][channel] -> [channel][row].  For c % 8 == 0 a warp moves 4 rows x 8 channels per step
  // (lane = 4 * channel + row): four full 32-byte global segments in, 32 distinct banks out, and no integer
  // division by the run-time channel count (measured 20 % of the kernel's instructions at c = 64).
  const bool fast_stage = (c & 7) == 0;
  if (fast_stage) {
    for (int rb = warp * 4; rb < KNN_QROWS; rb += 4 * (KNN_THREADS / 32))
      for (int cb = 0; cb < c; cb += 8) {
        const int r = rb + (lane & 3), ch = cb + (lane >> 2);
        const int q = q0 + r;
        QT[ch * KNN_QT_LD + r] = (q < n) ? __fsub_rn(__ldg(xc + (size_t)q * ldx + ch), mu[ch]) : 0.f;
      }
  } else {
    for (int e = tid; e < KNN_QROWS * c; e += KNN_THREADS) {
      const int r = e / c, ch = e - r * c;
      const int q = q0 + r;
      QT[ch * KNN_QT_LD + r] = (q < n) ? __fsub_rn(__ldg(xc + (size_t)q * ldx + ch), mu[ch]) : 0.f;
    }
  }
  for (int e = tid; e < KNN_QROWS * KNN_MAXK; e += KNN_THREADS) { lkey[e] = 0xffffffffu; lidx[e] = 0x7fffffff; }
  __syncthreads();
  if (tid < KNN_QROWS) {
    float s = 0.f;
    for (int ch = 0; ch < c; ++ch) { const float v = QT[ch * KNN_QT_LD + tid]; s = fmaf(v, v, s); }
    sqq[tid] = s;
  }

  for (int j0 = 0; j0 < n; j0 += KNN_CHUNK) {
    __syncthreads();  // previous chunk consumed (and sqq visible on the first pass)
    if (fast_stage) {
      for (int rb = warp * 4; rb < KNN_CHUNK; rb += 4 * (KNN_THREADS / 32))
        for (int cb = 0; cb < c; cb += 8) {
          const int r = rb + (lane & 3), ch = cb + (lane >> 2);
          const int j = j0 + r;
          XT[ch * KNN_XT_LD + r] = (j < n) ? __fsub_rn(__ldg(xc + (size_t)j * ldx + ch), mu[ch]) : 0.f;
        }
    } else {
      for (int e = tid; e < KNN_CHUNK * c; e += KNN_THREADS) {
        const int r = e / c, ch = e - r * c;
        const int j = j0 + r;
        XT[ch * KNN_XT_LD + r] = (j < n) ? __fsub_rn(__ldg(xc + (size_t)j * ldx + ch), mu[ch]) : 0.f;
      }
    }
    __syncthreads();
    {
      float s = 0.f;
      for (int ch = 0; ch < c; ++ch) { const float v = XT[ch * KNN_XT_LD + tid]; s = fmaf(v, v, s); }
      sqc[tid] = s;  // KNN_THREADS == KNN_CHUNK
    }
    __syncthreads();

    // ---- 8 rows x 8 candidates of dot products per lane
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int ch = 0; ch < c; ++ch) {
      const float4 a0 = *reinterpret_cast<const float4*>(QT + ch * KNN_QT_LD + warp * 8);
      const float4 a1 = *reinterpret_cast<const float4*>(QT + ch * KNN_QT_LD + warp * 8 + 4);
      // lane l owns candidates 4l..4l+3 and 128+4l..128+4l+3 of the chunk: both LDS.128 are conflict-free
      const float4 b0 = *reinterpret_cast<const float4*>(XT + ch * KNN_XT_LD + lane * 4);
      const float4 b1 = *reinterpret_cast<const float4*>(XT + ch * KNN_XT_LD + KNN_CHUNK / 2 + lane * 4);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }

    // ---- selection, one row at a time (the whole warp works on the same row)
    float sj[8];
    int cj[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int cl = (j < 4) ? lane * 4 + j : KNN_CHUNK / 2 + lane * 4 + (j - 4);   // position in the chunk
      sj[j] = sqc[cl];
      cj[j] = j0 + cl;
    }
    uint32_t* wk = stage_k + warp * 32;
    int* wi = stage_i + warp * 32;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = warp * 8 + i;
      const float si = sqq[row];
      uint32_t key[9];
      int cidx[9];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = __fadd_rn(__fadd_rn(si, __fmul_rn(-2.f, acc[i][j])), sj[j]);
        const bool ok = cj[j] < n;
        key[j] = ok ? sortable(d) : 0xffffffffu;
        cidx[j] = ok ? cj[j] : 0x7fffffff;
      }
      // 9th slot: the row's running list from earlier chunks (lane l holds entry l)
      key[8] = (lane < k) ? lkey[row * KNN_MAXK + lane] : 0xffffffffu;
      cidx[8] = (lane < k) ? lidx[row * KNN_MAXK + lane] : 0x7fffffff;
      __syncwarp();
      // threshold: T >= the k-th smallest of the 32 lane-local minima, so at least k candidates are <= T
      // (k rounds of "warp minimum, retire the lanes that hold it"; lanes tied at a minimum retire together,
      // which only makes T a little larger than necessary)
      uint32_t lmin = key[0];
#pragma unroll
      for (int j = 1; j < 9; ++j) lmin = min(lmin, key[j]);
      uint32_t T = 0;
      for (int r = 0; r < k; ++r) {
        T = __reduce_min_sync(0xffffffffu, lmin);
        lmin = (lmin == T) ? 0xffffffffu : lmin;
      }
      unsigned pm = 0;   // bit j: slot j survives
#pragma unroll
      for (int j = 0; j < 9; ++j) pm |= (key[j] <= T && cidx[j] != 0x7fffffff) ? (1u << j) : 0u;
      const int cnt = __popc(pm);
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      if (total <= 32) {
        int off = incl - cnt;
#pragma unroll
        for (int j = 0; j < 9; ++j)
          if (pm & (1u << j)) { wk[off] = key[j]; wi[off] = cidx[j]; ++off; }
        __syncwarp();
        const uint32_t mk = (lane < total) ? wk[lane] : 0xffffffffu;
        const int mi = (lane < total) ? wi[lane] : 0x7fffffff;
        int rank = 0;
        for (int t = 0; t < total; ++t) {
          const uint32_t ok = __shfl_sync(0xffffffffu, mk, t);
          const int oi = __shfl_sync(0xffffffffu, mi, t);
          rank += (ok < mk || (ok == mk && oi < mi)) ? 1 : 0;
        }
        __syncwarp();
        if (lane < total && rank < k) { lkey[row * KNN_MAXK + rank] = mk; lidx[row * KNN_MAXK + rank] = mi; }
        // (fewer than k survivors can only happen when chunk + list hold fewer than k points: the tail stays "empty")
        if (total < k && lane >= total && lane < k) { lkey[row * KNN_MAXK + lane] = 0xffffffffu; lidx[row * KNN_MAXK + lane] = 0x7fffffff; }
      } else {
        // fallback (mass ties at the threshold): k rounds of warp-wide arg-min
        for (int r = 0; r < k; ++r) {
          uint32_t lm = key[0];
#pragma unroll
          for (int j = 1; j < 9; ++j) lm = min(lm, key[j]);
          const uint32_t wmin = __reduce_min_sync(0xffffffffu, lm);
          int lcand = 0x7fffffff;
#pragma unroll
          for (int j = 0; j < 9; ++j) lcand = (key[j] == wmin) ? min(lcand, cidx[j]) : lcand;
          const int widx = __reduce_min_sync(0xffffffffu, lcand);
#pragma unroll
          for (int j = 0; j < 9; ++j)
            if (cidx[j] == widx) { key[j] = 0xffffffffu; cidx[j] = 0x7fffffff; }
          if (lane == r) { lkey[row * KNN_MAXK + r] = wmin; lidx[row * KNN_MAXK + r] = widx; }
        }
      }
      __syncwarp();
    }
  }
  __syncthreads();
  for (int e = tid; e < KNN_QROWS * k; e += KNN_THREADS) {
    const int r = e / k, s = e - r * k;
    const int q = q0 + r;
    if (q < n) idx_out[((size_t)cloud * n + q) * k + s] = lidx[r * KNN_MAXK + s];
  }
}

bool knn_tc_applicable(int n, int c, int k);
int knn_tc_launch(int b, int n, int c, int k, const float* x, int ldx, int* idx, const int* skip, cudaStream_t s);
int knn_classify_launch(int b, int n, int c, const float* x, int ldx, int* flags, cudaStream_t s);
int knn_tc_debug_counts(int b, int n, int c, int k, const float* x, int ldx, int* idx, int* counts, cudaStream_t s);

}  // namespace caae

using namespace caae;

// part: 0 = everything (tensor-core screen when it applies, else all-pairs), 1 = tensor-core kernel on the clouds NOT
// flagged, 2 = all-pairs kernel on the flagged clouds only (parts 1 and 2 together cover the batch and may run on
// two streams); -1 = all-pairs kernel for everything.
static int knn_impl(int b, int n, int c, int k, const float* x, int ldx, int* idx, caae_stream_t stream, int part,
                    const int* flags) {
  CAAE_RETURN_IF(b < 0 || n < 0 || c <= 0 || k <= 0 || ldx < c, CAAE_E_BADSHAPE);
  CAAE_RETURN_IF(k > KNN_MAXK || k > n, CAAE_E_BADSHAPE);  // tf.nn.top_k also rejects k > N
  if (b == 0 || n == 0) return CAAE_OK;
  CAAE_RETURN_IF(!x || !idx, CAAE_E_NULLPTR);
  CAAE_RETURN_IF(b > 65535, CAAE_E_BADSHAPE);
  const size_t smem = sizeof(float) * ((size_t)c * (KNN_XT_LD + KNN_QT_LD) + KNN_CHUNK + KNN_QROWS) +
                      sizeof(int) * 2 * KNN_QROWS * KNN_MAXK + sizeof(int) * 2 * KNN_THREADS + sizeof(float) * 5 * (size_t)c;
  CAAE_RETURN_IF(smem > 220 * 1024, CAAE_E_UNSUPPORTED);
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(knn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    smem_set = smem;
  }
  dim3 grid((n + KNN_QROWS - 1) / KNN_QROWS, b);
  // Gram matrix on the tensor cores + exact re-rank of a shortlist (knn_tc.cu) for one-CTA clouds; the all-pairs FFMA
  // kernel for everything else.  CAAE_KNN_TC=0: FFMA kernel always.
  static const bool tc_enabled = [] { const char* e = getenv("CAAE_KNN_TC"); return !(e && e[0] == '0'); }();
  const bool tc = part >= 0 && tc_enabled && knn_tc_applicable(n, c, k);
  if (part == 1 && !tc) part = 0, flags = nullptr;       // the all-pairs kernel does the whole batch in part 1 ...
  else if (part == 2 && !tc) return CAAE_OK;             // ... and part 2 has nothing left to do
  if (tc && part != 2)
    return knn_tc_launch(b, n, c, k, x, ldx, idx, part == 1 ? flags : nullptr, as_stream(stream));
  caae::launch(knn_kernel, grid, KNN_THREADS, smem, as_stream(stream), n, c, k, x, ldx, idx, part == 2 ? flags : nullptr);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_knn(int b, int n, int c, int k, const float* x, int ldx, int* idx, caae_stream_t stream) {
  return knn_impl(b, n, c, k, x, ldx, idx, stream, 0, nullptr);
}

// Debug aid: the tensor-core kernel with the per-row shortlist sizes written to counts[b*n] (n <= 256, c <= 64).
extern "C" int caae_debug_knn_shortlist(int b, int n, int c, int k, const float* x, int ldx, int* idx, int* counts,
                                        caae_stream_t stream) {
  CAAE_RETURN_IF(!knn_tc_applicable(n, c, k) || !x || !idx || !counts, CAAE_E_UNSUPPORTED);
  return knn_tc_debug_counts(b, n, c, k, x, ldx, idx, counts, as_stream(stream));
}

extern "C" int caae_knn_classify(int b, int n, int c, const float* x, int ldx, int* flags, caae_stream_t stream) {
  CAAE_RETURN_IF(b < 0 || n <= 0 || c <= 0 || ldx < c || b > 65535, CAAE_E_BADSHAPE);
  if (b == 0) return CAAE_OK;
  CAAE_RETURN_IF(!x || !flags, CAAE_E_NULLPTR);
  return knn_classify_launch(b, n, c, x, ldx, flags, as_stream(stream));
}

extern "C" int caae_knn_part(int part, const int* flags, int b, int n, int c, int k, const float* x, int ldx, int* idx,
                             caae_stream_t stream) {
  CAAE_RETURN_IF(part != 1 && part != 2, CAAE_E_BADSHAPE);
  CAAE_RETURN_IF(!flags, CAAE_E_NULLPTR);
  return knn_impl(b, n, c, k, x, ldx, idx, stream, part, flags);
}

// The all-pairs FFMA kernel alone (what caae_knn runs for n > 256 or c > 64): the yardstick the tensor-core screen
// is tested against (bit-identical indices).
extern "C" int caae_knn_ffma(int b, int n, int c, int k, const float* x, int ldx, int* idx, caae_stream_t stream) {
  return knn_impl(b, n, c, k, x, ldx, idx, stream, -1, nullptr);
}
