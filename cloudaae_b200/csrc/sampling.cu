// sampling.cu — farthest point sampling, gather_point (+grad) and prob_sample for sm_100a.
//
// Replaces farthestpointsamplingKernel, gatherpointKernel, scatteraddpointKernel, cumsumKernel and
// binarysearchKernel with their launchers (reference tf_ops/sampling/tf_sampling_g.cu:7-211).
//
// FPS design.  The algorithm is a chain of m-1 dependent rounds, so the kernel is built around the
// latency of one round, not around bandwidth:
//   * one CTA of 512 threads per cloud (grid = b; the reference runs 32 CTAs whatever b is);
//   * thread t owns points k = t, t+512, ... exactly as the reference does, with coordinates and
//     the running min-distance held in REGISTERS (the reference keeps the latter in global memory);
//   * the block argmax is two REDUX instructions per warp (max of the distance bits, then min of a
//     priority among the lanes holding that max), one shared-memory hop through a double-buffered
//     16-entry array and ONE __syncthreads per round (the reference: a 9-level shared-memory tree
//     with 10 barriers per round).
// Tie rule.  Reference: per-thread strict `>` over ascending k, then a tree that keeps the lower
// slot on equality, i.e. max d -> lowest (k mod 512) -> lowest k.  Here: priority = t*J + j for
// k = t + 512*j, smaller wins — the same total order.  Squared distances are >= +0, so their IEEE
// bit patterns order like unsigned integers.
#include "common.cuh"

namespace caae {

constexpr int kFpsThreads = 512;
constexpr int kFpsWarps = kFpsThreads / 32;
constexpr int kFpsMaxRegN = kFpsThreads * 16;  // clouds up to 8192 points stay in registers

__device__ __forceinline__ void fps_block_argmax(uint32_t dbits, uint32_t prio, uint2 (*red)[kFpsWarps], int buf,
                                                 uint32_t& win_prio) {
  const uint32_t wmax = __reduce_max_sync(0xffffffffu, dbits);
  const uint32_t wpri = __reduce_min_sync(0xffffffffu, dbits == wmax ? prio : 0xffffffffu);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[buf][warp] = make_uint2(wmax, wpri);
  __syncthreads();
  uint2 v = (lane < kFpsWarps) ? red[buf][lane] : make_uint2(0u, 0xffffffffu);
  const uint32_t bmax = __reduce_max_sync(0xffffffffu, v.x);
  win_prio = __reduce_min_sync(0xffffffffu, v.x == bmax ? v.y : 0xffffffffu);
}

// Register-resident kernel: PPT = points per thread = ceil(n / 512).
template <int PPT>
__global__ void __launch_bounds__(kFpsThreads)
fps_reg_kernel(int n, int m, const float* __restrict__ inp, int* __restrict__ out, float* __restrict__ out_xyz) {
  pdl_wait();
  extern __shared__ __align__(16) float s_xyz[];  // n*3: coordinate lookup of the last pick
  __shared__ uint2 s_red[2][kFpsWarps];

  const int cloud = blockIdx.x, t = threadIdx.x;
  const float* __restrict__ pts = inp + (size_t)cloud * n * 3;
  int* __restrict__ o = out + (size_t)cloud * m;

  for (int i = t; i < n * 3; i += kFpsThreads) s_xyz[i] = __ldg(pts + i);
  __syncthreads();

  float px[PPT], py[PPT], pz[PPT], td[PPT];
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    const int k = t + kFpsThreads * j;
    const bool ok = k < n;
    px[j] = ok ? s_xyz[k * 3 + 0] : 0.f;
    py[j] = ok ? s_xyz[k * 3 + 1] : 0.f;
    pz[j] = ok ? s_xyz[k * 3 + 2] : 0.f;
    td[j] = ok ? 1e38f : -1.f;  // a slot past the end can never beat best = -1
  }

  int old = 0;
  if (t == 0) o[0] = 0;
  for (int r = 1; r < m; ++r) {
    const float x1 = s_xyz[old * 3 + 0], y1 = s_xyz[old * 3 + 1], z1 = s_xyz[old * 3 + 2];
    float best = -1.f;
    int bj = 0;
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      const float d = sqdist_ref(px[j], py[j], pz[j], x1, y1, z1);
      const float d2 = fminf(d, td[j]);
      td[j] = d2;
      if (d2 > best) { best = d2; bj = j; }
    }
    const bool has = best >= 0.f;
    const uint32_t dbits = has ? f2u(best) : 0u;
    const uint32_t prio = has ? (uint32_t)(t * PPT + bj) : 0xffffffffu;
    uint32_t win;
    fps_block_argmax(dbits, prio, s_red, r & 1, win);
    old = (int)(win / PPT) + kFpsThreads * (int)(win % PPT);
    if (t == 0) o[r] = old;
  }

  if (out_xyz != nullptr) {
    __syncthreads();  // o[] written by thread 0 is visible to the CTA
    float* __restrict__ ox = out_xyz + (size_t)cloud * m * 3;
    for (int i = t; i < m * 3; i += kFpsThreads) ox[i] = s_xyz[o[i / 3] * 3 + (i % 3)];
  }
}

// Any n: running distances in caller scratch (temp[cloud, n]), coordinates from global/L2.
__global__ void __launch_bounds__(kFpsThreads)
fps_generic_kernel(int n, int m, const float* __restrict__ inp, float* __restrict__ temp, int* __restrict__ out,
                   float* __restrict__ out_xyz) {
  pdl_wait();
  __shared__ uint2 s_red[2][kFpsWarps];
  const int cloud = blockIdx.x, t = threadIdx.x;
  const float* __restrict__ pts = inp + (size_t)cloud * n * 3;
  float* __restrict__ td = temp + (size_t)cloud * n;
  int* __restrict__ o = out + (size_t)cloud * m;
  const uint32_t J = (uint32_t)((n + kFpsThreads - 1) / kFpsThreads);

  for (int k = t; k < n; k += kFpsThreads) td[k] = 1e38f;
  int old = 0;
  if (t == 0) o[0] = 0;
  for (int r = 1; r < m; ++r) {
    const float x1 = __ldg(pts + old * 3 + 0), y1 = __ldg(pts + old * 3 + 1), z1 = __ldg(pts + old * 3 + 2);
    float best = -1.f;
    uint32_t bj = 0;
    uint32_t j = 0;
    for (int k = t; k < n; k += kFpsThreads, ++j) {
      const float d = sqdist_ref(__ldg(pts + k * 3 + 0), __ldg(pts + k * 3 + 1), __ldg(pts + k * 3 + 2), x1, y1, z1);
      const float tdk = td[k];
      const float d2 = fminf(d, tdk);
      if (d2 != tdk) td[k] = d2;
      if (d2 > best) { best = d2; bj = j; }
    }
    const bool has = best >= 0.f;
    const uint32_t dbits = has ? f2u(best) : 0u;
    const uint32_t prio = has ? (uint32_t)t * J + bj : 0xffffffffu;
    uint32_t win;
    fps_block_argmax(dbits, prio, s_red, r & 1, win);
    old = (int)(win / J) + kFpsThreads * (int)(win % J);
    if (t == 0) o[r] = old;
  }
  if (out_xyz != nullptr) {
    __syncthreads();
    float* __restrict__ ox = out_xyz + (size_t)cloud * m * 3;
    for (int i = t; i < m * 3; i += kFpsThreads) ox[i] = __ldg(pts + o[i / 3] * 3 + (i % 3));
  }
}

// ---- gather / scatter-add ----------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gather_kernel(long total, int n, int m, const float* __restrict__ inp, const int* __restrict__ idx,
              float* __restrict__ out) {
  pdl_wait();
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const long p = e / 3;            // flat (cloud, j)
    const int c = (int)(e - p * 3);
    const long cloud = p / m;
    out[e] = __ldg(inp + (cloud * n + __ldg(idx + p)) * 3 + c);
  }
}

__global__ void __launch_bounds__(256)
scatter_add_kernel(long total, int n, int m, const float* __restrict__ out_g, const int* __restrict__ idx,
                   float* __restrict__ inp_g) {
  pdl_wait();
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const long p = e / 3;
    const int c = (int)(e - p * 3);
    const long cloud = p / m;
    atomicAdd(inp_g + (cloud * n + __ldg(idx + p)) * 3 + c, __ldg(out_g + e));
  }
}

// ---- prob_sample: inclusive prefix sum + lower-bound search -------------------------------------
// prob_sample(inp_p, inp_r) draws, per row, the first index whose cumulative probability reaches r * total
// (reference tf_sampling_g.cu:7-104, tf_sampling.cpp:14-27,65-92).  The searched index depends on every rounding of
// the cumulative sums, so the kernel reproduces the reference's SUMMATION TREE — not its code:
//   * elements in groups of four with local prefixes x1, x1+x2, x3+(x1+x2), (x4+x3)+(x1+x2);
//   * the group totals combined pairwise into aligned power-of-two blocks (block 2^(u+1) = left half + right half);
//   * the inclusive prefix at group p = its aligned block of size lowbit(p+1) + the prefix in front of that block;
//   * element = local prefix + prefix of the preceding groups, + the running sum of earlier 8192-element chunks,
//     carried with the reference's two-term compensation.
// Here the tree lives in registers: a thread owns four consecutive groups (tree levels 0-1), a warp 128 groups
// (levels 2-6 by shuffles), the 16 warp totals go through one more shuffle tree (levels 7-10).  Groups past the end of
// a chunk are zero (x + 0 = x exactly), which leaves every sum that exists in the reference unchanged.
constexpr int kScanThreads = 512;
constexpr int kScanChunk = 8192;     // elements per chunk = 2048 groups = 4 per thread

__device__ __forceinline__ bool scan_down_target(int pos1, int w) {   // pos1 = position + 1 in units of 2^w-blocks' base level
  return (pos1 & ((1 << w) - 1)) == 0 && ((pos1 >> w) & 1) && (pos1 >> w) >= 3;
}

__global__ void __launch_bounds__(kScanThreads)
prob_sample_kernel(int n, int m, const float* __restrict__ inp_p, const float* __restrict__ inp_r,
                   float* __restrict__ temp, int* __restrict__ out) {
  pdl_wait();
  __shared__ float warp_total[kScanThreads / 32], warp_prefix[kScanThreads / 32], chunk_total;
  const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* __restrict__ src = inp_p + (size_t)row * n;
  float* __restrict__ cum = temp + (size_t)row * n;
  float carry = 0.f, carry_lo = 0.f;                       // compensated running sum over chunks
  for (int c0 = 0; c0 < n; c0 += kScanChunk) {
    const int len = min(n - c0, kScanChunk);
    const int last_group = ((len + 3) >> 2) - 1;
    // ---- this thread's 16 elements: local prefixes loc[g][0..3] and group totals tot[g]
    float loc[4][4], tot[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int e = (4 * tid + g) * 4;                     // first element of the group inside the chunk
      if (e + 3 < len) {
        const float x1 = src[c0 + e], x2 = src[c0 + e + 1], x3 = src[c0 + e + 2], x4 = src[c0 + e + 3];
        const float s12 = __fadd_rn(x2, x1), s34 = __fadd_rn(x4, x3);
        loc[g][0] = x1; loc[g][1] = s12; loc[g][2] = __fadd_rn(x3, s12); loc[g][3] = __fadd_rn(s34, s12);
      } else {                                             // ragged tail group: a serial prefix, repeated to the group's end
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) { if (e + i < len) v = __fadd_rn(v, src[c0 + e + i]); loc[g][i] = v; }
      }
      tot[g] = loc[g][3];
    }
    // ---- up-sweep: levels 0-1 in the thread, 2-6 across the warp, 7-10 across the warp totals
    float b1 = __fadd_rn(tot[1], tot[0]);
    float b3 = __fadd_rn(__fadd_rn(tot[3], tot[2]), b1);
    float T = b3;
#pragma unroll
    for (int w = 0; w < 5; ++w) {
      const float t = __shfl_up_sync(0xffffffffu, T, 1 << w);
      if (((lane + 1) & ((2 << w) - 1)) == 0) T = __fadd_rn(T, t);
    }
    if (lane == 31) warp_total[warp] = T;
    __syncthreads();
    if (warp == 0) {
      float W = (lane < kScanThreads / 32) ? warp_total[lane] : 0.f;
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const float t = __shfl_up_sync(0xffffffffu, W, 1 << w);
        if (((lane + 1) & ((2 << w) - 1)) == 0) W = __fadd_rn(W, t);
      }
#pragma unroll
      for (int w = 2; w >= 0; --w) {                       // down-sweep over the warp totals
        const float t = __shfl_up_sync(0xffffffffu, W, 1 << w);
        if (scan_down_target(lane + 1, w)) W = __fadd_rn(W, t);
      }
      if (lane < kScanThreads / 32) warp_prefix[lane] = W; // inclusive prefix at the end of each warp's 128 groups
    }
    __syncthreads();
    // ---- down-sweep inside the warp (levels 6..2); a source one step in front of the warp is the previous warp's prefix
    const float prev_warp = (warp > 0) ? warp_prefix[warp - 1] : 0.f;
    if (lane == 31) T = warp_prefix[warp];
    const int G1 = tid + 1;                                // this thread's position + 1 among the 512 thread blocks
#pragma unroll
    for (int w = 4; w >= 0; --w) {
      const float t = __shfl_up_sync(0xffffffffu, T, 1 << w);
      const float from = (lane + 1 == (1 << w)) ? prev_warp : t;
      if (scan_down_target(G1, w)) T = __fadd_rn(T, from);
    }
    // ---- levels 1, 0 in the thread: P = inclusive prefix at the end of the previous thread's groups
    float P = __shfl_up_sync(0xffffffffu, T, 1);
    if (lane == 0) P = prev_warp;
    float b0 = tot[0], b2 = tot[2];
    if (tid > 0) { b1 = __fadd_rn(b1, P); b0 = __fadd_rn(b0, P); }
    b2 = __fadd_rn(b2, b1);
    // group prefixes in front of each of the four groups: P, b0, b1, b2; inclusive prefix of the last one: T
    const float front[4] = {P, b0, b1, b2};
    const float incl[4] = {b0, b1, b2, T};
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int grp = 4 * tid + g, e = grp * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (e + i < len) cum[c0 + e + i] = __fadd_rn(grp == 0 ? loc[g][i] : __fadd_rn(loc[g][i], front[g]), carry);
      if (grp == last_group) chunk_total = incl[g];
    }
    __syncthreads();
    const float tt = __fadd_rn(chunk_total, carry_lo);
    const float next = __fadd_rn(carry, tt);
    carry_lo = __fsub_rn(tt, __fsub_rn(next, carry));
    carry = next;
    __syncthreads();
  }
  // ---- the first index whose cumulative sum reaches q = r * total, probed in descending powers of two exactly as
  // the reference probes it (so rows whose sums are not monotone to the last bit still agree)
  int top = 1;
  while (top < n) top <<= 1;
  const float total = cum[n - 1];
  for (int jq = tid; jq < m; jq += kScanThreads) {
    const float q = __fmul_rn(inp_r[(size_t)row * m + jq], total);
    int r = n - 1;
    for (int step = top; step > 0; step >>= 1) {
      const int cand = r - step;
      if (cand >= 0 && cum[cand] >= q) r = cand;
    }
    out[(size_t)row * m + jq] = r;
  }
}

template <int PPT>
static int launch_fps_reg(int b, int n, int m, const float* inp, int* out, float* out_xyz, cudaStream_t s) {
  const size_t smem = sizeof(float) * 3 * (size_t)n;
  static size_t smem_set = 0;   // per instantiation: opt in once per size class, not on every call
  if (smem > 48 * 1024 && smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(fps_reg_kernel<PPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    smem_set = smem;
  }
  caae::launch(fps_reg_kernel<PPT>, b, kFpsThreads, smem, s, n, m, inp, out, out_xyz);
  return CAAE_LAUNCH_STATUS();
}

}  // namespace caae

using namespace caae;

extern "C" size_t caae_fps_scratch_bytes(int b, int n) {
  if (b <= 0 || n <= kFpsMaxRegN) return 0;
  return sizeof(float) * (size_t)b * (size_t)n;
}

extern "C" int caae_fps_gather(int b, int n, int m, const float* inp, float* temp, int* out, float* out_xyz,
                               caae_stream_t stream) {
  CAAE_RETURN_IF(b < 0 || n < 0 || m < 0, CAAE_E_BADSHAPE);
  if (b == 0 || m == 0) return CAAE_OK;
  CAAE_RETURN_IF(n == 0, CAAE_E_BADSHAPE);  // cannot sample from an empty cloud
  CAAE_RETURN_IF(!inp || !out, CAAE_E_NULLPTR);
  cudaStream_t s = as_stream(stream);
  const int ppt = (n + kFpsThreads - 1) / kFpsThreads;
  if (ppt <= 1) return launch_fps_reg<1>(b, n, m, inp, out, out_xyz, s);
  if (ppt <= 2) return launch_fps_reg<2>(b, n, m, inp, out, out_xyz, s);
  if (ppt <= 3) return launch_fps_reg<3>(b, n, m, inp, out, out_xyz, s);
  if (ppt <= 4) return launch_fps_reg<4>(b, n, m, inp, out, out_xyz, s);
  if (ppt <= 5) return launch_fps_reg<5>(b, n, m, inp, out, out_xyz, s);
  if (ppt <= 6) return launch_fps_reg<6>(b, n, m, inp, out, out_xyz, s);
  if (ppt <= 8) return launch_fps_reg<8>(b, n, m, inp, out, out_xyz, s);
  if (ppt <= 12) return launch_fps_reg<12>(b, n, m, inp, out, out_xyz, s);
  if (ppt <= 16) return launch_fps_reg<16>(b, n, m, inp, out, out_xyz, s);
  CAAE_RETURN_IF(!temp, CAAE_E_SCRATCH);
  caae::launch(fps_generic_kernel, b, kFpsThreads, 0, s, n, m, inp, temp, out, out_xyz);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_fps(int b, int n, int m, const float* inp, float* temp, int* out, caae_stream_t stream) {
  return caae_fps_gather(b, n, m, inp, temp, out, nullptr, stream);
}

static inline int flat_grid(long total) {
  long blocks = (total + 255) / 256;
  const long cap = (long)kNumSMs * 8;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

extern "C" int caae_gather(int b, int n, int m, const float* inp, const int* idx, float* out, caae_stream_t stream) {
  CAAE_RETURN_IF(b < 0 || n < 0 || m < 0, CAAE_E_BADSHAPE);
  if (b == 0 || m == 0) return CAAE_OK;
  CAAE_RETURN_IF(n == 0, CAAE_E_BADSHAPE);
  CAAE_RETURN_IF(!inp || !idx || !out, CAAE_E_NULLPTR);
  const long total = (long)b * m * 3;
  caae::launch(gather_kernel, flat_grid(total), 256, 0, as_stream(stream), total, n, m, inp, idx, out);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_gather_grad(int b, int n, int m, const float* out_g, const int* idx, float* inp_g,
                                caae_stream_t stream) {
  CAAE_RETURN_IF(b < 0 || n < 0 || m < 0, CAAE_E_BADSHAPE);
  if (b == 0 || n == 0) return CAAE_OK;
  CAAE_RETURN_IF(!inp_g, CAAE_E_NULLPTR);
  cudaStream_t s = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(inp_g, 0, sizeof(float) * (size_t)b * n * 3, s);
  if (e != cudaSuccess) return (int)e;
  if (m == 0) return CAAE_OK;
  CAAE_RETURN_IF(!out_g || !idx, CAAE_E_NULLPTR);
  const long total = (long)b * m * 3;
  caae::launch(scatter_add_kernel, flat_grid(total), 256, 0, s, total, n, m, out_g, idx, inp_g);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_prob_sample(int b, int n, int m, const float* inp_p, const float* inp_r, float* temp, int* out,
                                caae_stream_t stream) {
  CAAE_RETURN_IF(b < 0 || n < 0 || m < 0, CAAE_E_BADSHAPE);
  if (b == 0 || m == 0) return CAAE_OK;
  CAAE_RETURN_IF(n == 0, CAAE_E_BADSHAPE);
  CAAE_RETURN_IF(!inp_p || !inp_r || !out, CAAE_E_NULLPTR);
  CAAE_RETURN_IF(!temp, CAAE_E_SCRATCH);
  caae::launch(prob_sample_kernel, b, kScanThreads, 0, as_stream(stream), n, m, inp_p, inp_r, temp, out);
  return CAAE_LAUNCH_STATUS();
}
