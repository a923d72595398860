// layers.cu — EdgeConv aggregation, training-mode batch normalisation, pooling (forward + backward).
//
// These kernels carry everything of the reference's layer stack that is not a GEMM:
//   tf_util.get_edge_feature + conv2d(1x1) + batch_norm_for_conv2d + relu + reduce_mean over k
//     (utils/tf_util.py:635-669, 111-179, 473-511; models/pointnet_ycb_23_decoder_4.py:337-404)
//   conv2d/fully_connected batch norm + relu, reduce_mean over points, max_pool2d
//     (utils/tf_util.py:321-391, 514-555; models/...:410-426, 59-60)
// and their gradients (TensorFlow autodiff in the reference).
//
// EdgeConv is factorised: concat(x_i, x_j - x_i) W = x_i (W_top - W_bot) + x_j W_bot, so one GEMM
// produces PQ[i] = [P_i | Q_i] (2*cout wide) per point and the k-neighbour tensor
// z_ij = P_i + Q_nn(i,j) is only ever formed in registers: once for the batch statistics, once for
// normalise + ReLU + mean over k, and twice in the backward pass.  Nothing of size [B,N,k,*] or
// [B,N,N] touches HBM.
//
// Batch statistics are accumulated in fp64 per CTA and written as partials [part][2*C]; a finalize
// kernel reduces them in a fixed order (deterministic, no atomics) and emits scale/shift, the saved
// mean/invstd and the EMA update  shadow = d*shadow + (1-d)*batch  (tf_util.py:493-509).
#include "common.cuh"

namespace caae {

constexpr float kBnEps = 1e-3f;  // tf.nn.batch_normalization(..., 1e-3), tf_util.py:510

// ---------------------------------------------------------------------------------------------
// EdgeConv: z_ij = P[i] + Q[nn(i,j)]
// block (32, 8): x = channel lane, y = point lane.  A CTA covers EDGE_PTS points of one cloud.
constexpr int EDGE_PTS = 32;

// mode 0: partial sums of (z, z^2);  mode 1: partial sums of (dy, dy*zhat)
template <int MODE>
__global__ void __launch_bounds__(256)
edge_reduce_kernel(int n, int k, int cout, const float* __restrict__ PQ, int ldpq, const int* __restrict__ idx,
                   const float* __restrict__ scale, const float* __restrict__ shift,
                   const float* __restrict__ mean, const float* __restrict__ invstd,
                   const float* __restrict__ dOut, int lddo, double* __restrict__ parts) {
  __shared__ int s_idx[EDGE_PTS * 32];
  __shared__ double s_a[8][32], s_b[8][32];
  const int cloud = blockIdx.y, p0 = blockIdx.x * EDGE_PTS;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
  const size_t base = (size_t)cloud * n;
  const int npts = min(EDGE_PTS, n - p0);
  for (int e = tid; e < npts * k; e += 256) s_idx[e] = idx[(base + p0) * k + e];
  __syncthreads();
  const int part = blockIdx.y * gridDim.x + blockIdx.x;
  const float invk = 1.f / (float)k;
  for (int ch = tx; ch < cout; ch += 32) {
    double a = 0.0, b = 0.0;
    float sc = 0.f, sh = 0.f, mu = 0.f, is = 0.f;
    if (MODE == 1) { sc = scale[ch]; sh = shift[ch]; mu = mean[ch]; is = invstd[ch]; }
    for (int p = ty; p < npts; p += 8) {
      const float pv = PQ[(base + p0 + p) * ldpq + ch];
      float g = 0.f;
      if (MODE == 1) g = dOut[(base + p0 + p) * lddo + ch] * invk;
      for (int j = 0; j < k; ++j) {
        const float z = pv + PQ[(base + s_idx[p * k + j]) * ldpq + cout + ch];
        if (MODE == 0) {
          a += (double)z;
          b += (double)z * (double)z;
        } else {
          const float y = fmaf(z, sc, sh);
          if (y > 0.f) {
            a += (double)g;
            b += (double)g * (double)((z - mu) * is);
          }
        }
      }
    }
    s_a[ty][tx] = a; s_b[ty][tx] = b;
    __syncthreads();
    if (ty == 0) {
#pragma unroll
      for (int r = 1; r < 8; ++r) { a += s_a[r][tx]; b += s_b[r][tx]; }
      parts[(size_t)part * 2 * cout + ch] = a;
      parts[(size_t)part * 2 * cout + cout + ch] = b;
    }
    __syncthreads();
  }
}

// out[i][ch] = (1/k) sum_j relu(z_ij*scale + shift)
__global__ void __launch_bounds__(256)
edge_apply_kernel(int n, int k, int cout, const float* __restrict__ PQ, int ldpq, const int* __restrict__ idx,
                  const float* __restrict__ scale, const float* __restrict__ shift, float* __restrict__ out, int ldo) {
  __shared__ int s_idx[EDGE_PTS * 32];
  const int cloud = blockIdx.y, p0 = blockIdx.x * EDGE_PTS;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
  const size_t base = (size_t)cloud * n;
  const int npts = min(EDGE_PTS, n - p0);
  for (int e = tid; e < npts * k; e += 256) s_idx[e] = idx[(base + p0) * k + e];
  __syncthreads();
  const float invk = 1.f / (float)k;
  for (int ch = tx; ch < cout; ch += 32) {
    const float sc = scale[ch], sh = shift[ch];
    for (int p = ty; p < npts; p += 8) {
      const float pv = PQ[(base + p0 + p) * ldpq + ch];
      float acc = 0.f;
      for (int j = 0; j < k; ++j) {
        const float z = pv + PQ[(base + s_idx[p * k + j]) * ldpq + cout + ch];
        acc += fmaxf(fmaf(z, sc, sh), 0.f);
      }
      out[(base + p0 + p) * ldo + ch] = acc * invk;
    }
  }
}

// dz_ij = gamma*invstd*(dy_ij - mdy - zhat_ij*mdyz); dP[i] = sum_j dz_ij; dQ[nn(i,j)] += dz_ij.
// coef = [mdy | mdyz | gis] each of length cout (see bn_bwd_finalize_kernel); dPQ zero-filled by the caller.
__global__ void __launch_bounds__(256)
edge_bwd_apply_kernel(int n, int k, int cout, const float* __restrict__ PQ, int ldpq, const int* __restrict__ idx,
                      const float* __restrict__ scale, const float* __restrict__ shift,
                      const float* __restrict__ mean, const float* __restrict__ invstd,
                      const float* __restrict__ coef, const float* __restrict__ dOut, int lddo,
                      float* __restrict__ dPQ, int lddpq) {
  __shared__ int s_idx[EDGE_PTS * 32];
  const int cloud = blockIdx.y, p0 = blockIdx.x * EDGE_PTS;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
  const size_t base = (size_t)cloud * n;
  const int npts = min(EDGE_PTS, n - p0);
  for (int e = tid; e < npts * k; e += 256) s_idx[e] = idx[(base + p0) * k + e];
  __syncthreads();
  const float invk = 1.f / (float)k;
  for (int ch = tx; ch < cout; ch += 32) {
    const float sc = scale[ch], sh = shift[ch], mu = mean[ch], is = invstd[ch];
    const float mdy = coef[ch], mdyz = coef[cout + ch], gis = coef[2 * cout + ch];
    for (int p = ty; p < npts; p += 8) {
      const float pv = PQ[(base + p0 + p) * ldpq + ch];
      const float g = dOut[(base + p0 + p) * lddo + ch] * invk;
      float dp = 0.f;
      for (int j = 0; j < k; ++j) {
        const size_t nb = base + s_idx[p * k + j];
        const float z = pv + PQ[nb * ldpq + cout + ch];
        const float dy = (fmaf(z, sc, sh) > 0.f) ? g : 0.f;
        const float dz = gis * (dy - mdy - (z - mu) * is * mdyz);
        dp += dz;
        atomicAdd(dPQ + nb * lddpq + cout + ch, dz);
      }
      dPQ[(base + p0 + p) * lddpq + ch] = dp;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Generic [R, C] row-major activations.  block (32, 8), a CTA covers COL_ROWS rows x 32*? columns.
constexpr int COL_ROWS = 128;

// MODE 0: (sum y, sum y^2).  MODE 1: (sum dy, sum dy*yhat), dy = dOut[r / group][c] * gscale masked by
// relu(y*scale+shift) > 0 (when relu != 0) and, for max-pool, by r % group == argmax[r / group][c].
template <int MODE>
__global__ void __launch_bounds__(256)
col_reduce_kernel(int R, int C, const float* __restrict__ Y, int ld, const float* __restrict__ scale,
                  const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ invstd,
                  const float* __restrict__ dOut, int lddo, int group, float gscale, int relu,
                  const int* __restrict__ argmax, double* __restrict__ parts) {
  __shared__ double s_a[8][32], s_b[8][32];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int ch = blockIdx.x * 32 + tx;
  const int r0 = blockIdx.y * COL_ROWS;
  const int r1 = min(R, r0 + COL_ROWS);
  double a = 0.0, b = 0.0;
  if (ch < C) {
    float sc = 0.f, sh = 0.f, mu = 0.f, is = 0.f;
    if (MODE == 1) { sc = scale[ch]; sh = shift[ch]; mu = mean[ch]; is = invstd[ch]; }
    for (int r = r0 + ty; r < r1; r += 8) {
      const float y = Y[(size_t)r * ld + ch];
      if (MODE == 0) {
        a += (double)y;
        b += (double)y * (double)y;
      } else {
        const int gr = r / group;
        bool on = !relu || fmaf(y, sc, sh) > 0.f;
        if (argmax != nullptr) on = on && (argmax[(size_t)gr * C + ch] == r - gr * group);
        if (on) {
          const float g = dOut[(size_t)gr * lddo + ch] * gscale;
          a += (double)g;
          b += (double)g * (double)((y - mu) * is);
        }
      }
    }
  }
  s_a[ty][tx] = a; s_b[ty][tx] = b;
  __syncthreads();
  if (ty == 0 && ch < C) {
#pragma unroll
    for (int r = 1; r < 8; ++r) { a += s_a[r][tx]; b += s_b[r][tx]; }
    parts[(size_t)blockIdx.y * 2 * C + ch] = a;
    parts[(size_t)blockIdx.y * 2 * C + C + ch] = b;
  }
}

// Fixed-order reduction of the fp64 partials: block (32, 8) = 32 channels x 8 partial lanes.
__device__ __forceinline__ void reduce_parts(int C, const double* __restrict__ parts, int nparts, int ch,
                                             double& s, double& ss) {
  __shared__ double r_a[8][32], r_b[8][32];
  const int tx = threadIdx.x, ty = threadIdx.y;
  double a = 0.0, b = 0.0;
  if (ch < C)
    for (int p = ty; p < nparts; p += 8) { a += parts[(size_t)p * 2 * C + ch]; b += parts[(size_t)p * 2 * C + C + ch]; }
  r_a[ty][tx] = a; r_b[ty][tx] = b;
  __syncthreads();
  s = 0.0; ss = 0.0;
  if (ty == 0) {
#pragma unroll
    for (int r = 0; r < 8; ++r) { s += r_a[r][tx]; ss += r_b[r][tx]; }
  }
}

// forward finalize
__global__ void __launch_bounds__(256)
bn_finalize_kernel(int C, const double* __restrict__ parts, int nparts, double count,
                   const float* __restrict__ gamma, const float* __restrict__ beta,
                   float* __restrict__ ema_mean, float* __restrict__ ema_var,
                   const float* __restrict__ decay, float* __restrict__ scale,
                   float* __restrict__ shift, float* __restrict__ save_mean,
                   float* __restrict__ save_invstd) {
  const int ch = blockIdx.x * 32 + threadIdx.x;
  double s, ss;
  reduce_parts(C, parts, nparts, ch, s, ss);
  if (threadIdx.y != 0 || ch >= C) return;
  const double m = s / count;
  double v = ss / count - m * m;  // biased variance, as tf.nn.moments
  if (v < 0.0) v = 0.0;
  const float mf = (float)m, vf = (float)v;
  const float is = rsqrtf(vf + kBnEps);
  const float sc = gamma[ch] * is;
  scale[ch] = sc;
  shift[ch] = beta[ch] - mf * sc;
  save_mean[ch] = mf;
  save_invstd[ch] = is;
  if (ema_mean != nullptr) {
    const float d = decay ? *decay : 0.9f;
    ema_mean[ch] = d * ema_mean[ch] + (1.f - d) * mf;
    ema_var[ch] = d * ema_var[ch] + (1.f - d) * vf;
  }
}

// inference: scale/shift from the EMA statistics (tf_util.py:507-509)
__global__ void bn_eval_coeffs_kernel(int C, const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ ema_mean, const float* __restrict__ ema_var,
                                      float* __restrict__ scale, float* __restrict__ shift) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= C) return;
  const float sc = gamma[ch] * rsqrtf(ema_var[ch] + kBnEps);
  scale[ch] = sc;
  shift[ch] = beta[ch] - ema_mean[ch] * sc;
}

// backward finalize: coef = [mean(dy) | mean(dy*yhat) | gamma*invstd]; dgamma = sum dy*yhat, dbeta = sum dy
__global__ void __launch_bounds__(256)
bn_bwd_finalize_kernel(int C, const double* __restrict__ parts, int nparts, double count,
                       const float* __restrict__ gamma, const float* __restrict__ invstd,
                       float* __restrict__ coef, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int ch = blockIdx.x * 32 + threadIdx.x;
  double s, ss;
  reduce_parts(C, parts, nparts, ch, s, ss);
  if (threadIdx.y != 0 || ch >= C) return;
  coef[ch] = (float)(s / count);
  coef[C + ch] = (float)(ss / count);
  coef[2 * C + ch] = gamma[ch] * invstd[ch];
  dgamma[ch] = (float)ss;
  dbeta[ch] = (float)s;
}

// out = relu?(y*scale + shift)
__global__ void __launch_bounds__(256)
bn_act_kernel(long total, int C, const float* __restrict__ Y, int ld, const float* __restrict__ scale,
              const float* __restrict__ shift, int relu, float* __restrict__ out, int ldo) {
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const long r = e / C;
    const int ch = (int)(e - r * C);
    float v = fmaf(Y[r * ld + ch], scale[ch], shift[ch]);
    if (relu) v = fmaxf(v, 0.f);
    out[r * ldo + ch] = v;
  }
}

// emb[g][ch] = mean or max over the `group` rows of cloud g of relu(y*scale+shift); block (32, 8)
template <bool MAXPOOL>
__global__ void __launch_bounds__(256)
bn_act_pool_kernel(int group, int C, const float* __restrict__ Y, int ld, const float* __restrict__ scale,
                   const float* __restrict__ shift, float* __restrict__ emb, int* __restrict__ argmax) {
  __shared__ float s_v[8][32];
  __shared__ int s_i[8][32];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int ch = blockIdx.x * 32 + tx, g = blockIdx.y;
  float acc = MAXPOOL ? -INFINITY : 0.f;
  int arg = 0;
  if (ch < C) {
    const float sc = scale[ch], sh = shift[ch];
    for (int r = ty; r < group; r += 8) {
      const float v = fmaxf(fmaf(Y[((size_t)g * group + r) * ld + ch], sc, sh), 0.f);
      if (MAXPOOL) { if (v > acc) { acc = v; arg = r; } }
      else acc += v;
    }
  }
  s_v[ty][tx] = acc; s_i[ty][tx] = arg;
  __syncthreads();
  if (ty == 0 && ch < C) {
#pragma unroll
    for (int r = 1; r < 8; ++r) {
      if (MAXPOOL) {  // first maximum wins (lowest row), as a serial scan would
        const float v = s_v[r][tx];
        if (v > acc || (v == acc && s_i[r][tx] < arg)) { acc = v; arg = s_i[r][tx]; }
      } else acc += s_v[r][tx];
    }
    emb[(size_t)g * C + ch] = MAXPOOL ? acc : acc / (float)group;
    if (MAXPOOL && argmax) argmax[(size_t)g * C + ch] = arg;
  }
}

// dY = gamma*invstd*(dy - mdy - yhat*mdyz) (training BN) — may run in place (dY == Y)
__global__ void __launch_bounds__(256)
bn_act_bwd_kernel(long total, int C, const float* Y, int ld, const float* __restrict__ scale,
                  const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ invstd,
                  const float* __restrict__ coef, const float* __restrict__ dOut, int lddo, int group, float gscale,
                  int relu, const int* __restrict__ argmax, float* dY, int lddy) {
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const long r = e / C;
    const int ch = (int)(e - r * C);
    const float y = Y[r * ld + ch];
    const long gr = r / group;
    bool on = !relu || fmaf(y, scale[ch], shift[ch]) > 0.f;
    if (argmax != nullptr) on = on && (argmax[gr * C + ch] == (int)(r - gr * group));
    const float dy = on ? dOut[gr * lddo + ch] * gscale : 0.f;
    dY[r * lddy + ch] = coef[2 * C + ch] * (dy - coef[ch] - (y - mean[ch]) * invstd[ch] * coef[C + ch]);
  }
}

// out[c] = sum_r X[r][c]  (bias gradients of the linear output layers; R is the batch)
__global__ void __launch_bounds__(256)
colsum_kernel(int R, int C, const float* __restrict__ X, int ld, float* __restrict__ out) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= C) return;
  float s = 0.f;
  for (int r = 0; r < R; ++r) s += X[(size_t)r * ld + ch];
  out[ch] = s;
}

// EdgeConv weight factorisation: W [2c, cout] -> Wf [c, 2cout] = [W_top - W_bot | W_bot], bias_f = [bias | 0]
__global__ void edge_fold_weights_kernel(int c, int cout, const float* __restrict__ w, int ldw,
                                         const float* __restrict__ bias, float* __restrict__ wf,
                                         float* __restrict__ bias_f) {
  const int total = c * cout;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int r = e / cout, o = e - r * cout;
    const float top = w[(size_t)r * ldw + o], bot = w[(size_t)(c + r) * ldw + o];
    wf[(size_t)r * 2 * cout + o] = top - bot;
    wf[(size_t)r * 2 * cout + cout + o] = bot;
    if (r == 0) { bias_f[o] = bias ? bias[o] : 0.f; bias_f[cout + o] = 0.f; }
  }
}

// gradient of the factorisation: dW_top = dWf_P, dW_bot = dWf_Q - dWf_P
__global__ void edge_unfold_wgrad_kernel(int c, int cout, const float* __restrict__ dwf, int lddwf,
                                         float* __restrict__ dw) {
  const int total = c * cout;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int r = e / cout, o = e - r * cout;
    const float gp = dwf[(size_t)r * lddwf + o], gq = dwf[(size_t)r * lddwf + cout + o];
    dw[(size_t)r * cout + o] = gp;
    dw[(size_t)(c + r) * cout + o] = gq - gp;
  }
}

static inline int flat_blocks(long total) {
  long blocks = (total + 255) / 256;
  const long cap = (long)kNumSMs * 16;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace caae

using namespace caae;

// ---- EdgeConv -------------------------------------------------------------------------------
extern "C" int caae_edge_parts(int b, int n) { return b * ((n + EDGE_PTS - 1) / EDGE_PTS); }

static int edge_args_ok(int b, int n, int k, int cout) {
  return !(b < 0 || n <= 0 || k <= 0 || k > 32 || cout <= 0 || (cout % 32) != 0 || b > 65535);
}

extern "C" int caae_edge_stats(int b, int n, int k, int cout, const float* PQ, int ldpq, const int* idx,
                               double* parts, caae_stream_t stream) {
  CAAE_RETURN_IF(!edge_args_ok(b, n, k, cout) || ldpq < 2 * cout, CAAE_E_BADSHAPE);
  if (b == 0) return CAAE_OK;
  CAAE_RETURN_IF(!PQ || !idx || !parts, CAAE_E_NULLPTR);
  dim3 grid((n + EDGE_PTS - 1) / EDGE_PTS, b), block(32, 8);
  edge_reduce_kernel<0><<<grid, block, 0, as_stream(stream)>>>(n, k, cout, PQ, ldpq, idx, nullptr, nullptr, nullptr,
                                                                nullptr, nullptr, 0, parts);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_edge_apply(int b, int n, int k, int cout, const float* PQ, int ldpq, const int* idx,
                               const float* scale, const float* shift, float* out, int ldo, caae_stream_t stream) {
  CAAE_RETURN_IF(!edge_args_ok(b, n, k, cout) || ldpq < 2 * cout || ldo < cout, CAAE_E_BADSHAPE);
  if (b == 0) return CAAE_OK;
  CAAE_RETURN_IF(!PQ || !idx || !scale || !shift || !out, CAAE_E_NULLPTR);
  dim3 grid((n + EDGE_PTS - 1) / EDGE_PTS, b), block(32, 8);
  edge_apply_kernel<<<grid, block, 0, as_stream(stream)>>>(n, k, cout, PQ, ldpq, idx, scale, shift, out, ldo);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_edge_bwd_reduce(int b, int n, int k, int cout, const float* PQ, int ldpq, const int* idx,
                                    const float* scale, const float* shift, const float* mean, const float* invstd,
                                    const float* dOut, int lddo, double* parts, caae_stream_t stream) {
  CAAE_RETURN_IF(!edge_args_ok(b, n, k, cout) || ldpq < 2 * cout || lddo < cout, CAAE_E_BADSHAPE);
  if (b == 0) return CAAE_OK;
  CAAE_RETURN_IF(!PQ || !idx || !scale || !shift || !mean || !invstd || !dOut || !parts, CAAE_E_NULLPTR);
  dim3 grid((n + EDGE_PTS - 1) / EDGE_PTS, b), block(32, 8);
  edge_reduce_kernel<1><<<grid, block, 0, as_stream(stream)>>>(n, k, cout, PQ, ldpq, idx, scale, shift, mean, invstd,
                                                                dOut, lddo, parts);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_edge_bwd_apply(int b, int n, int k, int cout, const float* PQ, int ldpq, const int* idx,
                                   const float* scale, const float* shift, const float* mean, const float* invstd,
                                   const float* coef, const float* dOut, int lddo, float* dPQ, int lddpq,
                                   caae_stream_t stream) {
  CAAE_RETURN_IF(!edge_args_ok(b, n, k, cout) || ldpq < 2 * cout || lddo < cout || lddpq < 2 * cout, CAAE_E_BADSHAPE);
  if (b == 0) return CAAE_OK;
  CAAE_RETURN_IF(!PQ || !idx || !scale || !shift || !mean || !invstd || !coef || !dOut || !dPQ, CAAE_E_NULLPTR);
  cudaStream_t s = as_stream(stream);
  cudaError_t e = cudaMemset2DAsync(dPQ, sizeof(float) * (size_t)lddpq, 0, sizeof(float) * 2 * (size_t)cout,
                                    (size_t)b * n, s);
  if (e != cudaSuccess) return (int)e;
  dim3 grid((n + EDGE_PTS - 1) / EDGE_PTS, b), block(32, 8);
  edge_bwd_apply_kernel<<<grid, block, 0, s>>>(n, k, cout, PQ, ldpq, idx, scale, shift, mean, invstd, coef, dOut, lddo,
                                               dPQ, lddpq);
  return CAAE_LAUNCH_STATUS();
}

// ---- generic [R,C] ---------------------------------------------------------------------------
extern "C" int caae_col_parts(int R) { return (R + COL_ROWS - 1) / COL_ROWS; }

extern "C" int caae_col_stats(int R, int C, const float* Y, int ld, double* parts, caae_stream_t stream) {
  CAAE_RETURN_IF(R <= 0 || C <= 0 || ld < C, CAAE_E_BADSHAPE);
  CAAE_RETURN_IF(!Y || !parts, CAAE_E_NULLPTR);
  dim3 grid((C + 31) / 32, (R + COL_ROWS - 1) / COL_ROWS), block(32, 8);
  CAAE_RETURN_IF(grid.y > 65535, CAAE_E_BADSHAPE);
  col_reduce_kernel<0><<<grid, block, 0, as_stream(stream)>>>(R, C, Y, ld, nullptr, nullptr, nullptr, nullptr, nullptr,
                                                               0, 1, 1.f, 0, nullptr, parts);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_bn_finalize(int C, const double* parts, int nparts, double count, const float* gamma,
                                const float* beta, float* ema_mean, float* ema_var, const float* decay,
                                float* scale, float* shift, float* save_mean, float* save_invstd,
                                caae_stream_t stream) {
  CAAE_RETURN_IF(C <= 0 || nparts <= 0 || count <= 0, CAAE_E_BADSHAPE);
  CAAE_RETURN_IF(!parts || !gamma || !beta || !scale || !shift || !save_mean || !save_invstd, CAAE_E_NULLPTR);
  bn_finalize_kernel<<<(C + 31) / 32, dim3(32, 8), 0, as_stream(stream)>>>(C, parts, nparts, count, gamma, beta, ema_mean,
                                                                     ema_var, decay, scale, shift, save_mean,
                                                                     save_invstd);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_bn_eval_coeffs(int C, const float* gamma, const float* beta, const float* ema_mean,
                                   const float* ema_var, float* scale, float* shift, caae_stream_t stream) {
  CAAE_RETURN_IF(C <= 0, CAAE_E_BADSHAPE);
  CAAE_RETURN_IF(!gamma || !beta || !ema_mean || !ema_var || !scale || !shift, CAAE_E_NULLPTR);
  bn_eval_coeffs_kernel<<<(C + 127) / 128, 128, 0, as_stream(stream)>>>(C, gamma, beta, ema_mean, ema_var, scale, shift);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_bn_bwd_finalize(int C, const double* parts, int nparts, double count, const float* gamma,
                                    const float* invstd, float* coef, float* dgamma, float* dbeta,
                                    caae_stream_t stream) {
  CAAE_RETURN_IF(C <= 0 || nparts <= 0 || count <= 0, CAAE_E_BADSHAPE);
  CAAE_RETURN_IF(!parts || !gamma || !invstd || !coef || !dgamma || !dbeta, CAAE_E_NULLPTR);
  bn_bwd_finalize_kernel<<<(C + 31) / 32, dim3(32, 8), 0, as_stream(stream)>>>(C, parts, nparts, count, gamma, invstd, coef,
                                                                         dgamma, dbeta);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_bn_act(int R, int C, const float* Y, int ld, const float* scale, const float* shift, int relu,
                           float* out, int ldo, caae_stream_t stream) {
  CAAE_RETURN_IF(R <= 0 || C <= 0 || ld < C || ldo < C, CAAE_E_BADSHAPE);
  CAAE_RETURN_IF(!Y || !scale || !shift || !out, CAAE_E_NULLPTR);
  const long total = (long)R * C;
  bn_act_kernel<<<flat_blocks(total), 256, 0, as_stream(stream)>>>(total, C, Y, ld, scale, shift, relu, out, ldo);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_bn_act_pool(int groups, int group, int C, const float* Y, int ld, const float* scale,
                                const float* shift, int maxpool, float* emb, int* argmax, caae_stream_t stream) {
  CAAE_RETURN_IF(groups <= 0 || group <= 0 || C <= 0 || ld < C || groups > 65535, CAAE_E_BADSHAPE);
  CAAE_RETURN_IF(!Y || !scale || !shift || !emb, CAAE_E_NULLPTR);
  dim3 grid((C + 31) / 32, groups), block(32, 8);
  if (maxpool) bn_act_pool_kernel<true><<<grid, block, 0, as_stream(stream)>>>(group, C, Y, ld, scale, shift, emb, argmax);
  else bn_act_pool_kernel<false><<<grid, block, 0, as_stream(stream)>>>(group, C, Y, ld, scale, shift, emb, argmax);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_bn_act_bwd_reduce(int R, int C, const float* Y, int ld, const float* scale, const float* shift,
                                      const float* mean, const float* invstd, const float* dOut, int lddo, int group,
                                      float gscale, int relu, const int* argmax, double* parts,
                                      caae_stream_t stream) {
  CAAE_RETURN_IF(R <= 0 || C <= 0 || ld < C || lddo < C || group <= 0, CAAE_E_BADSHAPE);
  CAAE_RETURN_IF(!Y || !scale || !shift || !mean || !invstd || !dOut || !parts, CAAE_E_NULLPTR);
  dim3 grid((C + 31) / 32, (R + COL_ROWS - 1) / COL_ROWS), block(32, 8);
  CAAE_RETURN_IF(grid.y > 65535, CAAE_E_BADSHAPE);
  col_reduce_kernel<1><<<grid, block, 0, as_stream(stream)>>>(R, C, Y, ld, scale, shift, mean, invstd, dOut, lddo,
                                                               group, gscale, relu, argmax, parts);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_bn_act_bwd_apply(int R, int C, const float* Y, int ld, const float* scale, const float* shift,
                                     const float* mean, const float* invstd, const float* coef, const float* dOut,
                                     int lddo, int group, float gscale, int relu, const int* argmax, float* dY,
                                     int lddy, caae_stream_t stream) {
  CAAE_RETURN_IF(R <= 0 || C <= 0 || ld < C || lddo < C || lddy < C || group <= 0, CAAE_E_BADSHAPE);
  CAAE_RETURN_IF(!Y || !scale || !shift || !mean || !invstd || !coef || !dOut || !dY, CAAE_E_NULLPTR);
  const long total = (long)R * C;
  bn_act_bwd_kernel<<<flat_blocks(total), 256, 0, as_stream(stream)>>>(total, C, Y, ld, scale, shift, mean, invstd,
                                                                       coef, dOut, lddo, group, gscale, relu, argmax,
                                                                       dY, lddy);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_colsum(int R, int C, const float* X, int ld, float* out, caae_stream_t stream) {
  CAAE_RETURN_IF(R < 0 || C <= 0 || ld < C, CAAE_E_BADSHAPE);
  CAAE_RETURN_IF(!X || !out, CAAE_E_NULLPTR);
  colsum_kernel<<<(C + 255) / 256, 256, 0, as_stream(stream)>>>(R, C, X, ld, out);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_edge_fold_weights(int c, int cout, const float* w, const float* bias, float* wf, float* bias_f,
                                      int ldw, caae_stream_t stream) {
  CAAE_RETURN_IF(c <= 0 || cout <= 0 || ldw < cout, CAAE_E_BADSHAPE);
  CAAE_RETURN_IF(!w || !wf || !bias_f, CAAE_E_NULLPTR);
  edge_fold_weights_kernel<<<(c * cout + 255) / 256, 256, 0, as_stream(stream)>>>(c, cout, w, ldw, bias, wf, bias_f);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_edge_unfold_wgrad(int c, int cout, const float* dwf, int lddwf, float* dw, caae_stream_t stream) {
  CAAE_RETURN_IF(c <= 0 || cout <= 0 || lddwf < 2 * cout, CAAE_E_BADSHAPE);
  CAAE_RETURN_IF(!dwf || !dw, CAAE_E_NULLPTR);
  edge_unfold_wgrad_kernel<<<(c * cout + 255) / 256, 256, 0, as_stream(stream)>>>(c, cout, dwf, lddwf, dw);
  return CAAE_LAUNCH_STATUS();
}
