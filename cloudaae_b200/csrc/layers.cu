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
  pdl_wait();
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
                  const float* __restrict__ scale, const float* __restrict__ shift, float* __restrict__ out, int ldo,
                  float* __restrict__ out_lo) {
  pdl_wait();
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
      const float v = acc * invk;
      out[(base + p0 + p) * ldo + ch] = v;
      if (out_lo != nullptr) out_lo[(base + p0 + p) * ldo + ch] = v - tf32_rne(v);
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
  pdl_wait();
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
  pdl_wait();
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

// Fixed-order reduction of the fp64 partials: block (32, 32) = 32 channels x 32 partial lanes, four
// independent accumulators per lane so the L2 round trips overlap instead of chaining.
__device__ __forceinline__ void reduce_parts(int C, const double* __restrict__ parts, int nparts, int ch,
                                             double& s, double& ss) {
  __shared__ double r_a[32][33], r_b[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
  if (ch < C) {
    const size_t st = (size_t)2 * C;
    int p = ty;
    for (; p + 96 < nparts; p += 128) {
      a0 += parts[(size_t)p * st + ch];          b0 += parts[(size_t)p * st + C + ch];
      a1 += parts[(size_t)(p + 32) * st + ch];   b1 += parts[(size_t)(p + 32) * st + C + ch];
      a2 += parts[(size_t)(p + 64) * st + ch];   b2 += parts[(size_t)(p + 64) * st + C + ch];
      a3 += parts[(size_t)(p + 96) * st + ch];   b3 += parts[(size_t)(p + 96) * st + C + ch];
    }
    for (; p < nparts; p += 32) { a0 += parts[(size_t)p * st + ch]; b0 += parts[(size_t)p * st + C + ch]; }
  }
  r_a[ty][tx] = (a0 + a1) + (a2 + a3);
  r_b[ty][tx] = (b0 + b1) + (b2 + b3);
  __syncthreads();
  s = 0.0; ss = 0.0;
  if (ty == 0) {
#pragma unroll 8
    for (int r = 0; r < 32; ++r) { s += r_a[r][tx]; ss += r_b[r][tx]; }
  }
}

// forward finalize
__global__ void __launch_bounds__(1024)
bn_finalize_kernel(int C, const double* __restrict__ parts, int nparts, double count,
                   const float* __restrict__ gamma, const float* __restrict__ beta,
                   float* __restrict__ ema_mean, float* __restrict__ ema_var,
                   const float* __restrict__ decay, float* __restrict__ scale,
                   float* __restrict__ shift, float* __restrict__ save_mean,
                   float* __restrict__ save_invstd) {
  pdl_wait();
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
  pdl_wait();
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= C) return;
  const float sc = gamma[ch] * rsqrtf(ema_var[ch] + kBnEps);
  scale[ch] = sc;
  shift[ch] = beta[ch] - ema_mean[ch] * sc;
}

// backward finalize: coef = [mean(dy) | mean(dy*yhat) | gamma*invstd]; dgamma = sum dy*yhat, dbeta = sum dy
__global__ void __launch_bounds__(1024)
bn_bwd_finalize_kernel(int C, const double* __restrict__ parts, int nparts, double count,
                       const float* __restrict__ gamma, const float* __restrict__ invstd,
                       float* __restrict__ coef, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  pdl_wait();
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
  pdl_wait();
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
  pdl_wait();
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

// float4 variant of the mean pool (C % 128 == 0, 16-byte aligned rows): block (32, 8) = 128 channels x 8 row lanes,
// four independent 16-byte loads in flight per thread
__global__ void __launch_bounds__(256)
bn_act_meanpool_vec4_kernel(int group, int C, const float* __restrict__ Y, int ld, const float* __restrict__ scale,
                            const float* __restrict__ shift, float* __restrict__ emb, float* __restrict__ pos_cnt,
                            float* __restrict__ pos_sum) {
  pdl_wait();
  __shared__ float4 s_v[8][32];
  __shared__ float4 s_c[8][32], s_y[8][32];
  float4 cnt = make_float4(0.f, 0.f, 0.f, 0.f), ysum = make_float4(0.f, 0.f, 0.f, 0.f);
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int ch = blockIdx.x * 128 + tx * 4, g = blockIdx.y;
  const float4 sc = *reinterpret_cast<const float4*>(scale + ch), sh = *reinterpret_cast<const float4*>(shift + ch);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* base = Y + (size_t)g * group * ld + ch;
  for (int r = ty; r < group; r += 32) {
    float4 y[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      y[u] = (r + 8 * u < group) ? *reinterpret_cast<const float4*>(base + (size_t)(r + 8 * u) * ld) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (r + 8 * u >= group) break;
      const float a0 = fmaf(y[u].x, sc.x, sh.x), a1 = fmaf(y[u].y, sc.y, sh.y), a2 = fmaf(y[u].z, sc.z, sh.z), a3 = fmaf(y[u].w, sc.w, sh.w);
      acc.x += fmaxf(a0, 0.f); acc.y += fmaxf(a1, 0.f); acc.z += fmaxf(a2, 0.f); acc.w += fmaxf(a3, 0.f);
      if (pos_cnt != nullptr) {   // what the backward pass needs of the ReLU mask: per (cloud, channel) count and sum of y over the positive rows
        if (a0 > 0.f) { cnt.x += 1.f; ysum.x += y[u].x; }
        if (a1 > 0.f) { cnt.y += 1.f; ysum.y += y[u].y; }
        if (a2 > 0.f) { cnt.z += 1.f; ysum.z += y[u].z; }
        if (a3 > 0.f) { cnt.w += 1.f; ysum.w += y[u].w; }
      }
    }
  }
  s_v[ty][tx] = acc;
  if (pos_cnt != nullptr) { s_c[ty][tx] = cnt; s_y[ty][tx] = ysum; }
  __syncthreads();
  if (ty == 0) {
#pragma unroll
    for (int r = 1; r < 8; ++r) { const float4 o = s_v[r][tx]; acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w; }
    const float inv = 1.f / (float)group;
    *reinterpret_cast<float4*>(emb + (size_t)g * C + ch) = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
    if (pos_cnt != nullptr) {
#pragma unroll
      for (int r = 1; r < 8; ++r) {
        const float4 o = s_c[r][tx], q = s_y[r][tx];
        cnt.x += o.x; cnt.y += o.y; cnt.z += o.z; cnt.w += o.w; ysum.x += q.x; ysum.y += q.y; ysum.z += q.z; ysum.w += q.w;
      }
      *reinterpret_cast<float4*>(pos_cnt + (size_t)g * C + ch) = cnt;
      *reinterpret_cast<float4*>(pos_sum + (size_t)g * C + ch) = ysum;
    }
  }
}

// Backward statistics of mean-pool + ReLU + training-mode BN WITHOUT a pass over the [R, C] pre-activation: with
// dy[r][c] = relu'(..) * d_emb[g(r)][c] / group the two sums the BN backward needs collapse to per-(cloud, channel) terms,
//   sum_r dy = sum_g d_g * cnt[g][c],   sum_r dy * yhat = sum_g d_g * (pos_sum[g][c] - cnt[g][c] * mean[c]) * invstd[c],
// where cnt / pos_sum were written by the forward pool pass.  coef = [mean(dy) | mean(dy*yhat) | gamma*invstd] as
// bn_bwd_finalize_kernel produces it.  block (32 channels, 32 group lanes).
__global__ void __launch_bounds__(1024)
bn_pool_bwd_finalize_kernel(int C, int groups, double count, const float* __restrict__ d_emb, int ldd, float gscale,
                            const float* __restrict__ pos_cnt, const float* __restrict__ pos_sum,
                            const float* __restrict__ mean, const float* __restrict__ invstd,
                            const float* __restrict__ gamma, float* __restrict__ coef, float* __restrict__ dgamma,
                            float* __restrict__ dbeta) {
  pdl_wait();
  __shared__ double r_a[32][33], r_b[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int ch = blockIdx.x * 32 + tx;
  double a = 0.0, b = 0.0;
  if (ch < C) {
    const double mu = (double)mean[ch], is = (double)invstd[ch];
    for (int g = ty; g < groups; g += 32) {
      const double d = (double)(d_emb[(size_t)g * ldd + ch] * gscale);
      const double c = (double)pos_cnt[(size_t)g * C + ch];
      a += d * c;
      b += d * (((double)pos_sum[(size_t)g * C + ch] - c * mu) * is);
    }
  }
  r_a[ty][tx] = a; r_b[ty][tx] = b;
  __syncthreads();
  if (ty == 0 && ch < C) {
    double s = 0.0, ss = 0.0;
#pragma unroll 8
    for (int r = 0; r < 32; ++r) { s += r_a[r][tx]; ss += r_b[r][tx]; }
    coef[ch] = (float)(s / count);
    coef[C + ch] = (float)(ss / count);
    coef[2 * C + ch] = gamma[ch] * invstd[ch];
    dgamma[ch] = (float)ss;
    dbeta[ch] = (float)s;
  }
}

// dY = gamma*invstd*(dy - mdy - yhat*mdyz) (training BN) — may run in place (dY == Y)
__global__ void __launch_bounds__(256)
bn_act_bwd_kernel(long total, int C, const float* Y, int ld, const float* __restrict__ scale,
                  const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ invstd,
                  const float* __restrict__ coef, const float* __restrict__ dOut, int lddo, int group, float gscale,
                  int relu, const int* __restrict__ argmax, float* dY, int lddy) {
  pdl_wait();
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

// EdgeConv weight factorisation: W [2c, cout] -> Wf [c, 2cout] = [W_top - W_bot | W_bot], bias_f = [bias | 0]
__global__ void edge_fold_weights_kernel(int c, int cout, const float* __restrict__ w, int ldw,
                                         const float* __restrict__ bias, float* __restrict__ wf,
                                         float* __restrict__ bias_f) {
  pdl_wait();
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
  pdl_wait();
  const int total = c * cout;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int r = e / cout, o = e - r * cout;
    const float gp = dwf[(size_t)r * lddwf + o], gq = dwf[(size_t)r * lddwf + cout + o];
    dw[(size_t)r * cout + o] = gp;
    dw[(size_t)(c + r) * cout + o] = gq - gp;
  }
}

// ---- float4 variants for the wide layers (C % 128 == 0, 16-byte aligned rows): a thread owns 4 channels,
// keeps their per-channel constants in registers and streams rows with 4 loads in flight. ----
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

template <int MODE>
__global__ void __launch_bounds__(256)
col_reduce_vec4_kernel(int R, int C, const float* __restrict__ Y, int ld, const float* __restrict__ scale,
                       const float* __restrict__ shift, const float* __restrict__ mean,
                       const float* __restrict__ invstd, const float* __restrict__ dOut, int lddo, int group,
                       float gscale, int relu, const int* __restrict__ argmax, double* __restrict__ parts) {
  pdl_wait();
  __shared__ double s_a[8][128], s_b[8][128];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int ch = blockIdx.x * 128 + tx * 4;
  const int r0 = blockIdx.y * COL_ROWS, r1 = min(R, r0 + COL_ROWS);
  double a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
  float4 sc, sh, mu, is;
  if (MODE == 1) { sc = ld4(scale + ch); sh = ld4(shift + ch); mu = ld4(mean + ch); is = ld4(invstd + ch); }
  for (int r = r0 + ty; r < r1; r += 32) {
    float4 y[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) y[u] = (r + 8 * u < r1) ? ld4(Y + (size_t)(r + 8 * u) * ld + ch) : make_float4(0, 0, 0, 0);
    // (MODE 1: the four rows of a batch are summed in fp32 before they join the fp64 accumulators — the
    // fp32->fp64 conversions per element, not HBM, bounded this pass)
    float fa[4] = {0.f, 0.f, 0.f, 0.f}, fb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int rr = r + 8 * u;
      if (rr >= r1) break;
      const float yv[4] = {y[u].x, y[u].y, y[u].z, y[u].w};
      if (MODE == 0) {
#pragma unroll
        for (int c = 0; c < 4; ++c) { a[c] += (double)yv[c]; b[c] += (double)yv[c] * (double)yv[c]; }
      } else {
        const int gr = rr / group;
        const float4 g4 = ld4(dOut + (size_t)gr * lddo + ch);
        const float gv[4] = {g4.x * gscale, g4.y * gscale, g4.z * gscale, g4.w * gscale};
        const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
        const float muv[4] = {mu.x, mu.y, mu.z, mu.w}, isv[4] = {is.x, is.y, is.z, is.w};
        int am[4] = {0, 0, 0, 0};
        if (argmax != nullptr) {
          const int4 t = *reinterpret_cast<const int4*>(argmax + (size_t)gr * C + ch);
          am[0] = t.x; am[1] = t.y; am[2] = t.z; am[3] = t.w;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          bool on = !relu || fmaf(yv[c], scv[c], shv[c]) > 0.f;
          if (argmax != nullptr) on = on && (am[c] == rr - gr * group);
          if (on) { fa[c] += gv[c]; fb[c] = fmaf(gv[c], (yv[c] - muv[c]) * isv[c], fb[c]); }
        }
      }
    }
    if (MODE == 1) {
#pragma unroll
      for (int c = 0; c < 4; ++c) { a[c] += (double)fa[c]; b[c] += (double)fb[c]; }
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) { s_a[ty][tx * 4 + c] = a[c]; s_b[ty][tx * 4 + c] = b[c]; }
  __syncthreads();
  const int t = ty * 32 + tx;
  if (t < 128) {
    double sa = 0.0, sb = 0.0;
#pragma unroll
    for (int r = 0; r < 8; ++r) { sa += s_a[r][t]; sb += s_b[r][t]; }
    parts[(size_t)blockIdx.y * 2 * C + blockIdx.x * 128 + t] = sa;
    parts[(size_t)blockIdx.y * 2 * C + C + blockIdx.x * 128 + t] = sb;
  }
}

__global__ void __launch_bounds__(256)
bn_act_bwd_vec4_kernel(int R, int C, const float* Y, int ld, const float* __restrict__ scale,
                       const float* __restrict__ shift, const float* __restrict__ mean,
                       const float* __restrict__ invstd, const float* __restrict__ coef,
                       const float* __restrict__ dOut, int lddo, int group, float gscale, int relu,
                       const int* __restrict__ argmax, float* dY, int lddy) {
  pdl_wait();
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int ch = blockIdx.x * 128 + tx * 4;
  const int r0 = blockIdx.y * COL_ROWS, r1 = min(R, r0 + COL_ROWS);
  const float4 sc = ld4(scale + ch), sh = ld4(shift + ch), mu = ld4(mean + ch), is = ld4(invstd + ch);
  const float4 c0 = ld4(coef + ch), c1 = ld4(coef + C + ch), c2 = ld4(coef + 2 * C + ch);
  const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
  const float muv[4] = {mu.x, mu.y, mu.z, mu.w}, isv[4] = {is.x, is.y, is.z, is.w};
  const float mdy[4] = {c0.x, c0.y, c0.z, c0.w}, mdz[4] = {c1.x, c1.y, c1.z, c1.w}, gis[4] = {c2.x, c2.y, c2.z, c2.w};
  for (int r = r0 + ty; r < r1; r += 32) {
    float4 y[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) y[u] = (r + 8 * u < r1) ? ld4(Y + (size_t)(r + 8 * u) * ld + ch) : make_float4(0, 0, 0, 0);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int rr = r + 8 * u;
      if (rr >= r1) break;
      const int gr = rr / group;
      const float4 g4 = ld4(dOut + (size_t)gr * lddo + ch);
      const float gv[4] = {g4.x * gscale, g4.y * gscale, g4.z * gscale, g4.w * gscale};
      const float yv[4] = {y[u].x, y[u].y, y[u].z, y[u].w};
      int am[4] = {0, 0, 0, 0};
      if (argmax != nullptr) {
        const int4 t = *reinterpret_cast<const int4*>(argmax + (size_t)gr * C + ch);
        am[0] = t.x; am[1] = t.y; am[2] = t.z; am[3] = t.w;
      }
      float o[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        bool on = !relu || fmaf(yv[c], scv[c], shv[c]) > 0.f;
        if (argmax != nullptr) on = on && (am[c] == rr - gr * group);
        const float dy = on ? gv[c] : 0.f;
        o[c] = gis[c] * (dy - mdy[c] - (yv[c] - muv[c]) * isv[c] * mdz[c]);
      }
      *reinterpret_cast<float4*>(dY + (size_t)rr * lddy + ch) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
}

static inline bool vec4_ok(int C, const void* p0, int ld0, const void* p1, int ld1, const void* p2, int ld2) {
  auto al = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return (C % 128 == 0) && (ld0 % 4 == 0) && (ld1 % 4 == 0) && (ld2 % 4 == 0) && al(p0) && al(p1) && al(p2);
}

// ---------------------------------------------------------------------------------------------
// Cloud-resident EdgeConv kernels.  A CTA owns 64 output channels of ONE cloud: the cloud's Q rows
// for those channels (n x 64 fp32 = 64 KB at n = 256) and its neighbour table are staged in shared
// memory once, so the k-fold neighbour gather never leaves the SM; the backward scatter
// dQ[nn(i,j)] += dz_ij accumulates in shared memory too (no global atomics, no zero-fill pass).
// One fp64 partial row per cloud.  MODE: 0 stats, 1 apply, 2 backward reduce, 3 backward apply.
constexpr int ES_CH = 64;

// Optional fused batch-norm finalize (MODE 1: forward statistics -> scale / shift; MODE 3: backward statistics -> coef):
// every CTA reduces the fp64 partial rows of ITS 64 channels itself (fixed order), so the separate finalize launch —
// 4 us on the step's dependent chain, eight times per step — disappears; the CTA of cloud 0 writes the per-channel
// results (and the moving averages) for everyone else.
struct EdgeFinalize {
  const double* parts; int nparts; double count;
  const float* gamma; const float* beta; float* ema_mean; float* ema_var; const float* decay;   // MODE 1
  float* scale_out; float* shift_out; float* mean_out; float* invstd_out;                       // MODE 1
  const float* invstd_in; float* coef_out; float* dgamma; float* dbeta;                         // MODE 3
};

// REC (MODE 1 with the fused finalize only): also record, per (point, channel), the number of neighbours with a positive
// activation and the sum of their pre-activations z, centred on the batch mean mu the same kernel has just finalized.  The layer is followed by a mean over the k neighbours, so dL/dy_ij = d_out_i / k on the
// positive edges and the two batch-norm backward sums are  sum_i d_out_i/k * cnt_i  and  sum_i d_out_i/k * sz_i * invstd:
// the backward statistics become a streaming pass over [B*N, C] arrays (edge_bwd_stats_kernel) instead of a second
// staged gather over the k-neighbour tensor (MODE 2).
template <int MODE, bool REC = false>
__global__ void __launch_bounds__(1024)
edge_cloud_kernel(int n, int k, int cout, const float* __restrict__ PQ, int ldpq, const int* __restrict__ idx,
                  const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ mean,
                  const float* __restrict__ invstd, const float* __restrict__ coef, const float* __restrict__ dOut,
                  int lddo, float* __restrict__ out, int ldo, double* __restrict__ parts, float* __restrict__ out_lo,
                  const EdgeFinalize fz, unsigned char* __restrict__ pos_cnt, float* __restrict__ pos_sum, int ldpos) {
  pdl_wait();
  extern __shared__ __align__(16) float es_smem[];
  float* Qs = es_smem;                                   // [n][ES_CH]
  float* dQs = Qs + (size_t)n * ES_CH;                   // [n][ES_CH] (MODE 3 only)
  int* s_idx = reinterpret_cast<int*>(MODE == 3 ? dQs + (size_t)n * ES_CH : dQs);  // [n*k]
  __shared__ double s_a[32][ES_CH + 1], s_b[32][ES_CH + 1];

  const int cloud = blockIdx.y, c0 = blockIdx.x * ES_CH;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
  const size_t base = (size_t)cloud * n;
  for (int e = tid; e < n * (ES_CH / 4); e += 1024) {
    const int r = e / (ES_CH / 4), q = e - r * (ES_CH / 4);
    *reinterpret_cast<float4*>(Qs + r * ES_CH + q * 4) =
        *reinterpret_cast<const float4*>(PQ + (base + r) * ldpq + cout + c0 + q * 4);
    if (MODE == 3) *reinterpret_cast<float4*>(dQs + r * ES_CH + q * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int e = tid; e < n * k; e += 1024) s_idx[e] = idx[base * k + e];
  __syncthreads();

  const float invk = 1.f / (float)k;
  static_assert(ES_CH == 64, "a lane owns the channel pair (2 tx, 2 tx + 1) of the 64-channel slice");
  __shared__ float s_fz[4][ES_CH];
  const bool fused = (MODE == 1 || MODE == 3) && fz.parts != nullptr;
  if ((MODE == 1 || MODE == 3) && fused) {
    // thread = (channel c of the slice, row lane r of 16): partial rows r, r + 16, ... in ascending order, then the 16 lanes
    const int c = tid & (ES_CH - 1), r = tid >> 6;
    double a = 0.0, b = 0.0;
    for (int pr = r; pr < fz.nparts; pr += 16) {
      a += fz.parts[(size_t)pr * 2 * cout + c0 + c];
      b += fz.parts[(size_t)pr * 2 * cout + cout + c0 + c];
    }
    s_a[r][c] = a; s_b[r][c] = b;
    __syncthreads();
    if (tid < ES_CH) {
      double s = 0.0, ss = 0.0;
#pragma unroll
      for (int q = 0; q < 16; ++q) { s += s_a[q][tid]; ss += s_b[q][tid]; }
      const int chg = c0 + tid;
      if (MODE == 1) {
        const double m = s / fz.count;
        double v = ss / fz.count - m * m;  // biased variance, as tf.nn.moments
        if (v < 0.0) v = 0.0;
        const float mf = (float)m, vf = (float)v;
        const float isd = rsqrtf(vf + kBnEps);
        const float scv = fz.gamma[chg] * isd, shv = fz.beta[chg] - mf * scv;
        s_fz[0][tid] = scv; s_fz[1][tid] = shv; s_fz[3][tid] = mf;
        if (cloud == 0) {
          fz.scale_out[chg] = scv; fz.shift_out[chg] = shv; fz.mean_out[chg] = mf; fz.invstd_out[chg] = isd;
          if (fz.ema_mean != nullptr) {
            const float d = fz.decay ? *fz.decay : 0.9f;
            fz.ema_mean[chg] = d * fz.ema_mean[chg] + (1.f - d) * mf;
            fz.ema_var[chg] = d * fz.ema_var[chg] + (1.f - d) * vf;
          }
        }
      } else {
        const float c0v = (float)(s / fz.count), c1v = (float)(ss / fz.count), c2v = fz.gamma[chg] * fz.invstd_in[chg];
        s_fz[0][tid] = c0v; s_fz[1][tid] = c1v; s_fz[2][tid] = c2v;
        if (cloud == 0) {
          fz.coef_out[chg] = c0v; fz.coef_out[cout + chg] = c1v; fz.coef_out[2 * cout + chg] = c2v;
          fz.dgamma[chg] = (float)ss; fz.dbeta[chg] = (float)s;
        }
      }
    }
    __syncthreads();
  }
  {
    // one lane = two adjacent channels: the neighbour index is loaded once per pair and the gather is one LDS.64
    const int cl = 2 * tx, ch = c0 + cl;
    float sc[2] = {0.f, 0.f}, sh[2] = {0.f, 0.f}, mu[2] = {0.f, 0.f}, is[2] = {0.f, 0.f};
    float mdy[2] = {0.f, 0.f}, mdz[2] = {0.f, 0.f}, gis[2] = {0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      if (MODE == 1 && fused) { sc[c] = s_fz[0][cl + c]; sh[c] = s_fz[1][cl + c]; if (REC) mu[c] = s_fz[3][cl + c]; }
      else if (MODE >= 1) { sc[c] = scale[ch + c]; sh[c] = shift[ch + c]; }
      if (MODE >= 2) { mu[c] = mean[ch + c]; is[c] = invstd[ch + c]; }
      if (MODE == 3 && fused) { mdy[c] = s_fz[0][cl + c]; mdz[c] = s_fz[1][cl + c]; gis[c] = s_fz[2][cl + c]; }
      else if (MODE == 3) { mdy[c] = coef[ch + c]; mdz[c] = coef[cout + ch + c]; gis[c] = coef[2 * cout + ch + c]; }
    }
    double a[2] = {0.0, 0.0}, b[2] = {0.0, 0.0};
    for (int p = ty; p < n; p += 32) {
      const float2 pv2 = *reinterpret_cast<const float2*>(PQ + (base + p) * ldpq + ch);
      const float pv[2] = {pv2.x, pv2.y};
      float g[2] = {0.f, 0.f}, acc[2] = {0.f, 0.f};
      if (MODE >= 2) {
        const float* gp = dOut + (base + p) * lddo + ch;   // (a slice of a wider buffer: no alignment assumed)
        g[0] = gp[0] * invk; g[1] = gp[1] * invk;
      }
      const int* nb = s_idx + p * k;
      // sums over the k neighbours of ONE point in fp32 (k terms), across points in fp64: the fp32->fp64
      // conversions and DADDs per neighbour were the bottleneck of the two reduction passes
      float fa[2] = {0.f, 0.f}, fb[2] = {0.f, 0.f};
      for (int j = 0; j < k; ++j) {
        const int q = nb[j];
        const float2 q2 = *reinterpret_cast<const float2*>(Qs + q * ES_CH + cl);
        const float qv[2] = {q2.x, q2.y};
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float z = pv[c] + qv[c];
          if (MODE == 0) {
            fa[c] += z; fb[c] = fmaf(z, z, fb[c]);
          } else if (MODE == 1) {
            const float y = fmaf(z, sc[c], sh[c]);
            if (REC) { if (y > 0.f) { fa[c] += 1.f; fb[c] += z - mu[c]; } }   // centred on the batch mean: no cancellation later
            else acc[c] += fmaxf(y, 0.f);
          } else if (MODE == 2) {
            if (fmaf(z, sc[c], sh[c]) > 0.f) { fa[c] += 1.f; fb[c] += (z - mu[c]) * is[c]; }
          } else {
            const float dy = (fmaf(z, sc[c], sh[c]) > 0.f) ? g[c] : 0.f;
            const float dz = gis[c] * (dy - mdy[c] - (z - mu[c]) * is[c] * mdz[c]);
            acc[c] += dz;
            atomicAdd(dQs + q * ES_CH + cl + c, dz);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (MODE == 0) { a[c] += (double)fa[c]; b[c] += (double)fb[c]; }
        if (MODE == 2) { a[c] += (double)(g[c] * fa[c]); b[c] += (double)(g[c] * fb[c]); }
      }
      if (MODE == 1 && REC) {
        // sum over the positive edges of (sc z + sh) = sc * sum (z - mu) + cnt * (sc mu + sh): one accumulator less per neighbour
#pragma unroll
        for (int c = 0; c < 2; ++c) acc[c] = fmaxf(fmaf(sc[c], fb[c], fa[c] * fmaf(sc[c], mu[c], sh[c])), 0.f);   // (a sum of positive terms)
      }
      if (MODE == 1) {
        float* op = out + (base + p) * ldo + ch;
        const float v0 = acc[0] * invk, v1 = acc[1] * invk;
        op[0] = v0; op[1] = v1;
        if (out_lo != nullptr) {   // low part for the split-precision GEMMs that read this activation (same pitch)
          float* lp = out_lo + (base + p) * ldo + ch;
          lp[0] = v0 - tf32_rne(v0); lp[1] = v1 - tf32_rne(v1);
        }
      }
      if (MODE == 1 && REC) {
        // (even pitch and aligned bases are checked by the entry points: one 2-byte and one 8-byte store per lane)
        *reinterpret_cast<uchar2*>(pos_cnt + (base + p) * ldpos + ch) = make_uchar2((unsigned char)fa[0], (unsigned char)fa[1]);
        *reinterpret_cast<float2*>(pos_sum + (base + p) * ldpos + ch) = make_float2(fb[0], fb[1]);
      }
      if (MODE == 3) *reinterpret_cast<float2*>(out + (base + p) * ldo + ch) = make_float2(acc[0], acc[1]);  // dP
    }
    if (MODE == 0 || MODE == 2) {
      s_a[ty][cl] = a[0]; s_b[ty][cl] = b[0];
      s_a[ty][cl + 1] = a[1]; s_b[ty][cl + 1] = b[1];
    }
  }
  if (MODE == 0 || MODE == 2) {
    __syncthreads();
    if (tid < ES_CH) {
      double a = 0.0, b = 0.0;
#pragma unroll 8
      for (int r = 0; r < 32; ++r) { a += s_a[r][tid]; b += s_b[r][tid]; }
      parts[(size_t)cloud * 2 * cout + c0 + tid] = a;
      parts[(size_t)cloud * 2 * cout + cout + c0 + tid] = b;
    }
  }
  if (MODE == 3) {
    __syncthreads();
    for (int e = tid; e < n * (ES_CH / 4); e += 1024) {
      const int r = e / (ES_CH / 4), q = e - r * (ES_CH / 4);
      *reinterpret_cast<float4*>(out + (base + r) * ldo + cout + c0 + q * 4) =
          *reinterpret_cast<const float4*>(dQs + r * ES_CH + q * 4);
    }
  }
}

static inline size_t edge_cloud_smem(int n, int k, int mode) {
  return sizeof(float) * (size_t)n * ES_CH * (mode == 3 ? 2 : 1) + sizeof(int) * (size_t)n * k;
}
// the cloud-resident kernels apply when a cloud's tile fits shared memory and rows are float4-addressable
static inline bool edge_cloud_ok(int n, int k, int cout, int ldpq) {
  return (cout % ES_CH == 0) && (ldpq % 4 == 0) && edge_cloud_smem(n, k, 3) <= 200 * 1024;
}
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int MODE, bool REC = false>
static int launch_edge_cloud(int b, int n, int k, int cout, const float* PQ, int ldpq, const int* idx,
                             const float* scale, const float* shift, const float* mean, const float* invstd,
                             const float* coef, const float* dOut, int lddo, float* out, int ldo, double* parts,
                             cudaStream_t s, float* out_lo = nullptr, const EdgeFinalize* fz = nullptr,
                             unsigned char* pos_cnt = nullptr, float* pos_sum = nullptr, int ldpos = 0) {
  const size_t smem = edge_cloud_smem(n, k, MODE);
  // the kernel also holds 33 KB of static shared memory: the opt-in is needed well below 48 KB of dynamic memory
  // (a first call with n = 128 failed with "invalid argument" until a larger cloud had raised the limit).  Raised
  // once per size class, not per call.
  static size_t smem_set = 0;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(edge_cloud_kernel<MODE, REC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    smem_set = smem;
  }
  EdgeFinalize none = {};
  caae::launch(edge_cloud_kernel<MODE, REC>, dim3(cout / ES_CH, b), dim3(32, 32), smem, s, n, k, cout, PQ, ldpq, idx, scale, shift, mean,
               invstd, coef, dOut, lddo, out, ldo, parts, out_lo, fz ? *fz : none, pos_cnt, pos_sum, ldpos);
  return CAAE_LAUNCH_STATUS();
}

// Backward batch-norm statistics of an EdgeConv layer from what the forward apply pass recorded (see REC above): one
// CTA per (64-channel slice, cloud), a lane owns two adjacent channels, fp64 across points, one partial row per cloud
// in the layout of edge_cloud_kernel<2> (so caae_edge_bwd_apply[_fused] / caae_bn_bwd_finalize read it unchanged).
__global__ void __launch_bounds__(1024)
edge_bwd_stats_kernel(int n, int k, int cout, const float* __restrict__ dOut, int lddo,
                      const unsigned char* __restrict__ pos_cnt, const float* __restrict__ pos_sum, int ldpos,
                      const float* __restrict__ invstd, double* __restrict__ parts) {
  pdl_wait();
  __shared__ double s_a[32][ES_CH + 1], s_b[32][ES_CH + 1];
  const int cloud = blockIdx.y, c0 = blockIdx.x * ES_CH;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
  const size_t base = (size_t)cloud * n;
  const int cl = 2 * tx, ch = c0 + cl;
  const float invk = 1.f / (float)k;
  const float is[2] = {invstd[ch], invstd[ch + 1]};
  double a[2] = {0.0, 0.0}, b[2] = {0.0, 0.0};
  for (int p = ty; p < n; p += 32) {
    const float* gp = dOut + (base + p) * lddo + ch;   // (a slice of a wider buffer: no alignment assumed)
    const uchar2 c2 = *reinterpret_cast<const uchar2*>(pos_cnt + (base + p) * ldpos + ch);
    const float2 s2 = *reinterpret_cast<const float2*>(pos_sum + (base + p) * ldpos + ch);
    const float cv[2] = {(float)c2.x, (float)c2.y}, sv[2] = {s2.x, s2.y};
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const float g = gp[c] * invk, cnt = cv[c];
      const float yhat = sv[c] * is[c];   // sum of the normalised pre-activations of the positive edges
      a[c] += (double)(g * cnt); b[c] += (double)(g * yhat);
    }
  }
  s_a[ty][cl] = a[0]; s_b[ty][cl] = b[0];
  s_a[ty][cl + 1] = a[1]; s_b[ty][cl + 1] = b[1];
  __syncthreads();
  if (tid < ES_CH) {
    double sa = 0.0, sb = 0.0;
#pragma unroll 8
    for (int r = 0; r < 32; ++r) { sa += s_a[r][tid]; sb += s_b[r][tid]; }
    parts[(size_t)cloud * 2 * cout + c0 + tid] = sa;
    parts[(size_t)cloud * 2 * cout + cout + c0 + tid] = sb;
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
// Number of fp64 partial rows caae_edge_stats / caae_edge_bwd_reduce write for this shape: one per
// cloud on the cloud-resident path, one per 32-point chunk otherwise.
extern "C" int caae_edge_parts(int b, int n, int k, int cout, int ldpq) {
  return edge_cloud_ok(n, k, cout, ldpq) ? b : b * ((n + EDGE_PTS - 1) / EDGE_PTS);
}

// the recorded (count, sum) pairs are accessed as uchar2 / float2: even pitch, 2- / 8-byte aligned bases
static inline bool pos_pair_ok(const void* pos_cnt, const void* pos_sum, int ldpos) {
  return (ldpos % 2 == 0) && (reinterpret_cast<uintptr_t>(pos_cnt) & 1) == 0 && (reinterpret_cast<uintptr_t>(pos_sum) & 7) == 0;
}

static int edge_args_ok(int b, int n, int k, int cout) {
  return !(b < 0 || n <= 0 || k <= 0 || k > 32 || cout <= 0 || (cout % 32) != 0 || b > 65535);
}

extern "C" int caae_edge_stats(int b, int n, int k, int cout, const float* PQ, int ldpq, const int* idx,
                               double* parts, caae_stream_t stream) {
  CAAE_RETURN_IF(!edge_args_ok(b, n, k, cout) || ldpq < 2 * cout, CAAE_E_BADSHAPE);
  if (b == 0) return CAAE_OK;
  CAAE_RETURN_IF(!PQ || !idx || !parts, CAAE_E_NULLPTR);
  if (edge_cloud_ok(n, k, cout, ldpq)) {
    CAAE_RETURN_IF(!aligned16(PQ), CAAE_E_UNSUPPORTED);
    return launch_edge_cloud<0>(b, n, k, cout, PQ, ldpq, idx, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0,
                                nullptr, 0, parts, as_stream(stream));
  }
  dim3 grid((n + EDGE_PTS - 1) / EDGE_PTS, b), block(32, 8);
  caae::launch(edge_reduce_kernel<0>, grid, block, 0, as_stream(stream), n, k, cout, PQ, ldpq, idx, nullptr, nullptr, nullptr,
                                                                nullptr, nullptr, 0, parts);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_edge_apply(int b, int n, int k, int cout, const float* PQ, int ldpq, const int* idx,
                               const float* scale, const float* shift, float* out, int ldo, float* out_lo,
                               caae_stream_t stream) {
  CAAE_RETURN_IF(!edge_args_ok(b, n, k, cout) || ldpq < 2 * cout || ldo < cout, CAAE_E_BADSHAPE);
  if (b == 0) return CAAE_OK;
  CAAE_RETURN_IF(!PQ || !idx || !scale || !shift || !out, CAAE_E_NULLPTR);
  if (edge_cloud_ok(n, k, cout, ldpq)) {
    CAAE_RETURN_IF(!aligned16(PQ), CAAE_E_UNSUPPORTED);
    return launch_edge_cloud<1>(b, n, k, cout, PQ, ldpq, idx, scale, shift, nullptr, nullptr, nullptr, nullptr, 0, out,
                                ldo, nullptr, as_stream(stream), out_lo);
  }
  dim3 grid((n + EDGE_PTS - 1) / EDGE_PTS, b), block(32, 8);
  caae::launch(edge_apply_kernel, grid, block, 0, as_stream(stream), n, k, cout, PQ, ldpq, idx, scale, shift, out, ldo, out_lo);
  return CAAE_LAUNCH_STATUS();
}

// caae_edge_apply with the training-mode batch-norm finalize of caae_bn_finalize folded in (cloud-resident path only:
// caae_edge_parts(...) == b): parts / nparts / count as caae_edge_stats wrote them; scale, shift, save_mean, save_invstd
// and the moving averages are OUTPUTS.
extern "C" int caae_edge_apply_fused(int b, int n, int k, int cout, const float* PQ, int ldpq, const int* idx,
                                     const double* parts, int nparts, double count, const float* gamma, const float* beta,
                                     float* ema_mean, float* ema_var, const float* decay, float* scale, float* shift,
                                     float* save_mean, float* save_invstd, float* out, int ldo, float* out_lo,
                                     unsigned char* pos_cnt, float* pos_sum, int ldpos, caae_stream_t stream) {
  CAAE_RETURN_IF(!edge_args_ok(b, n, k, cout) || ldpq < 2 * cout || ldo < cout || nparts <= 0 || count <= 0, CAAE_E_BADSHAPE);
  if (b == 0) return CAAE_OK;
  CAAE_RETURN_IF(!PQ || !idx || !parts || !gamma || !beta || !scale || !shift || !save_mean || !save_invstd || !out, CAAE_E_NULLPTR);
  CAAE_RETURN_IF((ema_mean == nullptr) != (ema_var == nullptr), CAAE_E_NULLPTR);
  CAAE_RETURN_IF(!edge_cloud_ok(n, k, cout, ldpq) || !aligned16(PQ), CAAE_E_UNSUPPORTED);
  EdgeFinalize fz = {};
  fz.parts = parts; fz.nparts = nparts; fz.count = count; fz.gamma = gamma; fz.beta = beta; fz.ema_mean = ema_mean;
  fz.ema_var = ema_var; fz.decay = decay; fz.scale_out = scale; fz.shift_out = shift; fz.mean_out = save_mean; fz.invstd_out = save_invstd;
  CAAE_RETURN_IF((pos_cnt == nullptr) != (pos_sum == nullptr), CAAE_E_NULLPTR);
  CAAE_RETURN_IF(pos_cnt != nullptr && (ldpos < cout || k > 255), CAAE_E_BADSHAPE);
  CAAE_RETURN_IF(pos_cnt != nullptr && !pos_pair_ok(pos_cnt, pos_sum, ldpos), CAAE_E_UNSUPPORTED);
  if (pos_cnt != nullptr)
    return launch_edge_cloud<1, true>(b, n, k, cout, PQ, ldpq, idx, scale, shift, nullptr, nullptr, nullptr, nullptr, 0, out, ldo,
                                      nullptr, as_stream(stream), out_lo, &fz, pos_cnt, pos_sum, ldpos);
  return launch_edge_cloud<1>(b, n, k, cout, PQ, ldpq, idx, scale, shift, nullptr, nullptr, nullptr, nullptr, 0, out, ldo, nullptr,
                              as_stream(stream), out_lo, &fz);
}

// The backward statistics of caae_edge_bwd_reduce from the (pos_cnt, pos_sum) a recording caae_edge_apply_fused wrote:
// no neighbour gather, no staging.  Cloud-resident shapes only (caae_edge_parts(...) == b partial rows).
extern "C" int caae_edge_bwd_stats(int b, int n, int k, int cout, int ldpq, const float* dOut, int lddo,
                                   const unsigned char* pos_cnt, const float* pos_sum, int ldpos,
                                   const float* invstd, double* parts, caae_stream_t stream) {
  CAAE_RETURN_IF(!edge_args_ok(b, n, k, cout) || lddo < cout || ldpos < cout, CAAE_E_BADSHAPE);
  if (b == 0) return CAAE_OK;
  CAAE_RETURN_IF(!dOut || !pos_cnt || !pos_sum || !invstd || !parts, CAAE_E_NULLPTR);
  CAAE_RETURN_IF(!edge_cloud_ok(n, k, cout, ldpq) || !pos_pair_ok(pos_cnt, pos_sum, ldpos), CAAE_E_UNSUPPORTED);
  caae::launch(edge_bwd_stats_kernel, dim3(cout / ES_CH, b), dim3(32, 32), 0, as_stream(stream), n, k, cout, dOut, lddo, pos_cnt,
               pos_sum, ldpos, invstd, parts);
  return CAAE_LAUNCH_STATUS();
}

// caae_edge_bwd_apply with caae_bn_bwd_finalize folded in: parts as caae_edge_bwd_reduce wrote them; coef, dgamma, dbeta
// are OUTPUTS.
extern "C" int caae_edge_bwd_apply_fused(int b, int n, int k, int cout, const float* PQ, int ldpq, const int* idx,
                                         const float* scale, const float* shift, const float* mean, const float* invstd,
                                         const double* parts, int nparts, double count, const float* gamma, float* coef,
                                         float* dgamma, float* dbeta, const float* dOut, int lddo, float* dPQ, int lddpq,
                                         caae_stream_t stream) {
  CAAE_RETURN_IF(!edge_args_ok(b, n, k, cout) || ldpq < 2 * cout || lddo < cout || lddpq < 2 * cout || nparts <= 0 || count <= 0,
                 CAAE_E_BADSHAPE);
  if (b == 0) return CAAE_OK;
  CAAE_RETURN_IF(!PQ || !idx || !scale || !shift || !mean || !invstd || !parts || !gamma || !coef || !dgamma || !dbeta || !dOut || !dPQ,
                 CAAE_E_NULLPTR);
  CAAE_RETURN_IF(!edge_cloud_ok(n, k, cout, ldpq) || lddpq % 4 != 0 || !aligned16(PQ) || !aligned16(dPQ), CAAE_E_UNSUPPORTED);
  EdgeFinalize fz = {};
  fz.parts = parts; fz.nparts = nparts; fz.count = count; fz.gamma = gamma; fz.invstd_in = invstd; fz.coef_out = coef;
  fz.dgamma = dgamma; fz.dbeta = dbeta;
  return launch_edge_cloud<3>(b, n, k, cout, PQ, ldpq, idx, scale, shift, mean, invstd, coef, dOut, lddo, dPQ, lddpq, nullptr,
                              as_stream(stream), nullptr, &fz);
}

extern "C" int caae_edge_bwd_reduce(int b, int n, int k, int cout, const float* PQ, int ldpq, const int* idx,
                                    const float* scale, const float* shift, const float* mean, const float* invstd,
                                    const float* dOut, int lddo, double* parts, caae_stream_t stream) {
  CAAE_RETURN_IF(!edge_args_ok(b, n, k, cout) || ldpq < 2 * cout || lddo < cout, CAAE_E_BADSHAPE);
  if (b == 0) return CAAE_OK;
  CAAE_RETURN_IF(!PQ || !idx || !scale || !shift || !mean || !invstd || !dOut || !parts, CAAE_E_NULLPTR);
  if (edge_cloud_ok(n, k, cout, ldpq)) {
    CAAE_RETURN_IF(!aligned16(PQ), CAAE_E_UNSUPPORTED);
    return launch_edge_cloud<2>(b, n, k, cout, PQ, ldpq, idx, scale, shift, mean, invstd, nullptr, dOut, lddo, nullptr,
                                0, parts, as_stream(stream));
  }
  dim3 grid((n + EDGE_PTS - 1) / EDGE_PTS, b), block(32, 8);
  caae::launch(edge_reduce_kernel<1>, grid, block, 0, as_stream(stream), n, k, cout, PQ, ldpq, idx, scale, shift, mean, invstd,
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
  if (edge_cloud_ok(n, k, cout, ldpq) && lddpq % 4 == 0) {
    CAAE_RETURN_IF(!aligned16(PQ) || !aligned16(dPQ), CAAE_E_UNSUPPORTED);
    return launch_edge_cloud<3>(b, n, k, cout, PQ, ldpq, idx, scale, shift, mean, invstd, coef, dOut, lddo, dPQ, lddpq,
                                nullptr, s);
  }
  cudaError_t e = cudaMemset2DAsync(dPQ, sizeof(float) * (size_t)lddpq, 0, sizeof(float) * 2 * (size_t)cout,
                                    (size_t)b * n, s);
  if (e != cudaSuccess) return (int)e;
  dim3 grid((n + EDGE_PTS - 1) / EDGE_PTS, b), block(32, 8);
  caae::launch(edge_bwd_apply_kernel, grid, block, 0, s, n, k, cout, PQ, ldpq, idx, scale, shift, mean, invstd, coef, dOut, lddo,
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
  if (vec4_ok(C, Y, ld, nullptr, 0, nullptr, 0)) {
    grid.x = C / 128;
    caae::launch(col_reduce_vec4_kernel<0>, grid, block, 0, as_stream(stream), R, C, Y, ld, nullptr, nullptr, nullptr, nullptr,
                                                                      nullptr, 0, 1, 1.f, 0, nullptr, parts);
    return CAAE_LAUNCH_STATUS();
  }
  caae::launch(col_reduce_kernel<0>, grid, block, 0, as_stream(stream), R, C, Y, ld, nullptr, nullptr, nullptr, nullptr, nullptr,
                                                               0, 1, 1.f, 0, nullptr, parts);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_bn_finalize(int C, const double* parts, int nparts, double count, const float* gamma,
                                const float* beta, float* ema_mean, float* ema_var, const float* decay,
                                float* scale, float* shift, float* save_mean, float* save_invstd,
                                caae_stream_t stream) {
  CAAE_RETURN_IF(C <= 0 || nparts <= 0 || count <= 0, CAAE_E_BADSHAPE);
  CAAE_RETURN_IF(!parts || !gamma || !beta || !scale || !shift || !save_mean || !save_invstd, CAAE_E_NULLPTR);
  caae::launch(bn_finalize_kernel, (C + 31) / 32, dim3(32, 32), 0, as_stream(stream), C, parts, nparts, count, gamma, beta, ema_mean,
                                                                     ema_var, decay, scale, shift, save_mean,
                                                                     save_invstd);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_bn_eval_coeffs(int C, const float* gamma, const float* beta, const float* ema_mean,
                                   const float* ema_var, float* scale, float* shift, caae_stream_t stream) {
  CAAE_RETURN_IF(C <= 0, CAAE_E_BADSHAPE);
  CAAE_RETURN_IF(!gamma || !beta || !ema_mean || !ema_var || !scale || !shift, CAAE_E_NULLPTR);
  caae::launch(bn_eval_coeffs_kernel, (C + 127) / 128, 128, 0, as_stream(stream), C, gamma, beta, ema_mean, ema_var, scale, shift);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_bn_bwd_finalize(int C, const double* parts, int nparts, double count, const float* gamma,
                                    const float* invstd, float* coef, float* dgamma, float* dbeta,
                                    caae_stream_t stream) {
  CAAE_RETURN_IF(C <= 0 || nparts <= 0 || count <= 0, CAAE_E_BADSHAPE);
  CAAE_RETURN_IF(!parts || !gamma || !invstd || !coef || !dgamma || !dbeta, CAAE_E_NULLPTR);
  caae::launch(bn_bwd_finalize_kernel, (C + 31) / 32, dim3(32, 32), 0, as_stream(stream), C, parts, nparts, count, gamma, invstd, coef,
                                                                         dgamma, dbeta);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_bn_act(int R, int C, const float* Y, int ld, const float* scale, const float* shift, int relu,
                           float* out, int ldo, caae_stream_t stream) {
  CAAE_RETURN_IF(R <= 0 || C <= 0 || ld < C || ldo < C, CAAE_E_BADSHAPE);
  CAAE_RETURN_IF(!Y || !scale || !shift || !out, CAAE_E_NULLPTR);
  const long total = (long)R * C;
  caae::launch(bn_act_kernel, flat_blocks(total), 256, 0, as_stream(stream), total, C, Y, ld, scale, shift, relu, out, ldo);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_bn_pool_bwd_finalize(int C, int groups, int group, const float* d_emb, int ldd, float gscale,
                                         const float* pos_cnt, const float* pos_sum, const float* mean,
                                         const float* invstd, const float* gamma, float* coef, float* dgamma,
                                         float* dbeta, caae_stream_t stream) {
  CAAE_RETURN_IF(C <= 0 || groups <= 0 || group <= 0 || ldd < C, CAAE_E_BADSHAPE);
  CAAE_RETURN_IF(!d_emb || !pos_cnt || !pos_sum || !mean || !invstd || !gamma || !coef || !dgamma || !dbeta, CAAE_E_NULLPTR);
  caae::launch(bn_pool_bwd_finalize_kernel, (C + 31) / 32, dim3(32, 32), 0, as_stream(stream), C, groups,
               (double)groups * (double)group, d_emb, ldd, gscale, pos_cnt, pos_sum, mean, invstd, gamma, coef, dgamma, dbeta);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_bn_act_pool(int groups, int group, int C, const float* Y, int ld, const float* scale,
                                const float* shift, int maxpool, float* emb, int* argmax, float* pos_cnt, float* pos_sum,
                                caae_stream_t stream) {
  CAAE_RETURN_IF(groups <= 0 || group <= 0 || C <= 0 || ld < C || groups > 65535, CAAE_E_BADSHAPE);
  CAAE_RETURN_IF(!Y || !scale || !shift || !emb, CAAE_E_NULLPTR);
  dim3 grid((C + 31) / 32, groups), block(32, 8);
  if (!maxpool && C % 128 == 0 && ld % 4 == 0 && ((reinterpret_cast<uintptr_t>(Y) | reinterpret_cast<uintptr_t>(emb) |
                                                   reinterpret_cast<uintptr_t>(scale) | reinterpret_cast<uintptr_t>(shift)) & 15) == 0) {
    CAAE_RETURN_IF((pos_cnt == nullptr) != (pos_sum == nullptr), CAAE_E_NULLPTR);
    CAAE_RETURN_IF(pos_cnt && ((reinterpret_cast<uintptr_t>(pos_cnt) | reinterpret_cast<uintptr_t>(pos_sum)) & 15), CAAE_E_UNSUPPORTED);
    caae::launch(bn_act_meanpool_vec4_kernel, dim3(C / 128, groups), block, 0, as_stream(stream), group, C, Y, ld, scale, shift, emb,
                 pos_cnt, pos_sum);
    return CAAE_LAUNCH_STATUS();
  }
  CAAE_RETURN_IF(pos_cnt != nullptr || pos_sum != nullptr, CAAE_E_UNSUPPORTED);   // only the float4 mean-pool kernel records them
  if (maxpool) caae::launch(bn_act_pool_kernel<true>, grid, block, 0, as_stream(stream), group, C, Y, ld, scale, shift, emb, argmax);
  else caae::launch(bn_act_pool_kernel<false>, grid, block, 0, as_stream(stream), group, C, Y, ld, scale, shift, emb, argmax);
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
  if (vec4_ok(C, Y, ld, dOut, lddo, nullptr, 0)) {
    grid.x = C / 128;
    caae::launch(col_reduce_vec4_kernel<1>, grid, block, 0, as_stream(stream), R, C, Y, ld, scale, shift, mean, invstd, dOut,
                                                                      lddo, group, gscale, relu, argmax, parts);
    return CAAE_LAUNCH_STATUS();
  }
  caae::launch(col_reduce_kernel<1>, grid, block, 0, as_stream(stream), R, C, Y, ld, scale, shift, mean, invstd, dOut, lddo,
                                                               group, gscale, relu, argmax, parts);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_bn_act_bwd_apply(int R, int C, const float* Y, int ld, const float* scale, const float* shift,
                                     const float* mean, const float* invstd, const float* coef, const float* dOut,
                                     int lddo, int group, float gscale, int relu, const int* argmax, float* dY,
                                     int lddy, caae_stream_t stream) {
  CAAE_RETURN_IF(R <= 0 || C <= 0 || ld < C || lddo < C || lddy < C || group <= 0, CAAE_E_BADSHAPE);
  CAAE_RETURN_IF(!Y || !scale || !shift || !mean || !invstd || !coef || !dOut || !dY, CAAE_E_NULLPTR);
  if (vec4_ok(C, Y, ld, dOut, lddo, dY, lddy)) {
    dim3 grid(C / 128, (R + COL_ROWS - 1) / COL_ROWS), block(32, 8);
    CAAE_RETURN_IF(grid.y > 65535, CAAE_E_BADSHAPE);
    caae::launch(bn_act_bwd_vec4_kernel, grid, block, 0, as_stream(stream), R, C, Y, ld, scale, shift, mean, invstd, coef, dOut,
                                                                   lddo, group, gscale, relu, argmax, dY, lddy);
    return CAAE_LAUNCH_STATUS();
  }
  const long total = (long)R * C;
  caae::launch(bn_act_bwd_kernel, flat_blocks(total), 256, 0, as_stream(stream), total, C, Y, ld, scale, shift, mean, invstd,
                                                                       coef, dOut, lddo, group, gscale, relu, argmax,
                                                                       dY, lddy);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_edge_fold_weights(int c, int cout, const float* w, const float* bias, float* wf, float* bias_f,
                                      int ldw, caae_stream_t stream) {
  CAAE_RETURN_IF(c <= 0 || cout <= 0 || ldw < cout, CAAE_E_BADSHAPE);
  CAAE_RETURN_IF(!w || !wf || !bias_f, CAAE_E_NULLPTR);
  caae::launch(edge_fold_weights_kernel, (c * cout + 255) / 256, 256, 0, as_stream(stream), c, cout, w, ldw, bias, wf, bias_f);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_edge_unfold_wgrad(int c, int cout, const float* dwf, int lddwf, float* dw, caae_stream_t stream) {
  CAAE_RETURN_IF(c <= 0 || cout <= 0 || lddwf < 2 * cout, CAAE_E_BADSHAPE);
  CAAE_RETURN_IF(!dwf || !dw, CAAE_E_NULLPTR);
  caae::launch(edge_unfold_wgrad_kernel, (c * cout + 255) / 256, 256, 0, as_stream(stream), c, cout, dwf, lddwf, dw);
  return CAAE_LAUNCH_STATUS();
}
