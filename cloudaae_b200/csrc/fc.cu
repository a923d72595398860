// fc.cu — batch-norm kernels for the fully connected stack (rows = batch, e.g. 128): one launch per
// layer and direction.
//
// Reference: utils/tf_util.py:321-365 (fully_connected = matmul + bias + batch_norm_for_fc + relu) and
// :473-525 (batch_norm_template with moments over axis 0).  With only `batch` rows a CTA can own whole
// columns, so statistics, EMA update, coefficients and the normalise+ReLU pass (forward) or the two
// reductions, the parameter gradients and the data gradient (backward) need no grid-wide exchange:
// they replace the col_stats -> bn_finalize -> bn_act (forward) and bn_act_bwd_reduce ->
// bn_bwd_finalize -> bn_act_bwd_apply (backward) launch triples of layers.cu, which stay in use for
// the [B*N, C] encoder activations.  Same arithmetic: fp64 column sums in a fixed order, fp32 after.
#include "common.cuh"

namespace caae {

constexpr float kFcBnEps = 1e-3f;  // tf.nn.batch_normalization(..., 1e-3), tf_util.py:510
constexpr int FC_LANES = 32;       // row lanes per column: block (32 columns, 32 row lanes)

// Fixed-order sum over the row lanes of one column; result valid for ty == 0.
__device__ __forceinline__ void fc_lane_sum(double a, double b, double (*s_a)[33], double (*s_b)[33], double& sa,
                                            double& sb) {
  const int tx = threadIdx.x, ty = threadIdx.y;
  s_a[ty][tx] = a; s_b[ty][tx] = b;
  __syncthreads();
  sa = 0.0; sb = 0.0;
  if (ty == 0) {
#pragma unroll 8
    for (int r = 0; r < FC_LANES; ++r) { sa += s_a[r][tx]; sb += s_b[r][tx]; }
  }
}

// out = relu?(BN_train(Y)) over R rows; also scale/shift/mean/invstd for the backward pass and the EMA update.
__global__ void __launch_bounds__(1024)
fc_bn_fwd_kernel(int R, int C, const float* __restrict__ Y, int ld, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float* __restrict__ ema_mean, float* __restrict__ ema_var,
                 const float* __restrict__ decay, float* __restrict__ scale, float* __restrict__ shift,
                 float* __restrict__ save_mean, float* __restrict__ save_invstd, int relu, float* __restrict__ out,
                 int ldo, float* __restrict__ out_lo) {
  pdl_wait();
  __shared__ double s_a[FC_LANES][33], s_b[FC_LANES][33];
  __shared__ float s_sc[32], s_sh[32];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int ch = blockIdx.x * 32 + tx;
  double a = 0.0, b = 0.0;
  if (ch < C)
    for (int r = ty; r < R; r += FC_LANES) {
      const double y = (double)Y[(size_t)r * ld + ch];
      a += y; b += y * y;
    }
  double s, ss;
  fc_lane_sum(a, b, s_a, s_b, s, ss);
  if (ty == 0 && ch < C) {
    const double m = s / (double)R;
    double v = ss / (double)R - m * m;  // biased variance, as tf.nn.moments
    if (v < 0.0) v = 0.0;
    const float mf = (float)m, vf = (float)v;
    const float is = rsqrtf(vf + kFcBnEps);
    const float sc = gamma[ch] * is, sh = beta[ch] - mf * sc;
    scale[ch] = sc; shift[ch] = sh; save_mean[ch] = mf; save_invstd[ch] = is;
    s_sc[tx] = sc; s_sh[tx] = sh;
    if (ema_mean != nullptr) {
      const float d = decay ? *decay : 0.9f;
      ema_mean[ch] = d * ema_mean[ch] + (1.f - d) * mf;
      ema_var[ch] = d * ema_var[ch] + (1.f - d) * vf;
    }
  }
  __syncthreads();
  if (ch < C && out != nullptr) {
    const float sc = s_sc[tx], sh = s_sh[tx];
    for (int r = ty; r < R; r += FC_LANES) {
      float v = fmaf(Y[(size_t)r * ld + ch], sc, sh);
      if (relu) v = fmaxf(v, 0.f);
      out[(size_t)r * ldo + ch] = v;
      if (out_lo != nullptr) out_lo[(size_t)r * ldo + ch] = v - tf32_rne(v);   // low part for the split-precision GEMM behind
    }
  }
}

// dY = gamma*invstd*(dy - mean(dy) - yhat*mean(dy*yhat)), dy = dOut masked by relu(y*scale+shift) > 0;
// dgamma = sum dy*yhat, dbeta = sum dy.  dY may alias dOut (each element is read then written by one thread).
__global__ void __launch_bounds__(1024)
fc_bn_bwd_kernel(int R, int C, const float* __restrict__ Y, int ld, const float* __restrict__ scale,
                 const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ invstd,
                 const float* __restrict__ gamma, int relu, const float* dOut, int lddo, float* dY, int lddy,
                 float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dY_lo) {
  pdl_wait();
  __shared__ double s_a[FC_LANES][33], s_b[FC_LANES][33];
  __shared__ float s_c0[32], s_c1[32];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int ch = blockIdx.x * 32 + tx;
  float sc = 0.f, sh = 0.f, mu = 0.f, is = 0.f;
  double a = 0.0, b = 0.0;
  if (ch < C) {
    sc = scale[ch]; sh = shift[ch]; mu = mean[ch]; is = invstd[ch];
    for (int r = ty; r < R; r += FC_LANES) {
      const float y = Y[(size_t)r * ld + ch];
      if (!relu || fmaf(y, sc, sh) > 0.f) {
        const float g = dOut[(size_t)r * lddo + ch];
        a += (double)g;
        b += (double)g * (double)((y - mu) * is);
      }
    }
  }
  double s, ss;
  fc_lane_sum(a, b, s_a, s_b, s, ss);
  if (ty == 0 && ch < C) {
    s_c0[tx] = (float)(s / (double)R);
    s_c1[tx] = (float)(ss / (double)R);
    dgamma[ch] = (float)ss;
    dbeta[ch] = (float)s;
  }
  __syncthreads();
  if (ch < C) {
    const float mdy = s_c0[tx], mdz = s_c1[tx], gis = gamma[ch] * is;
    for (int r = ty; r < R; r += FC_LANES) {
      const float y = Y[(size_t)r * ld + ch];
      const bool on = !relu || fmaf(y, sc, sh) > 0.f;
      const float dy = on ? dOut[(size_t)r * lddo + ch] : 0.f;
      const float d = gis * (dy - mdy - (y - mu) * is * mdz);
      dY[(size_t)r * lddy + ch] = d;
      if (dY_lo != nullptr) dY_lo[(size_t)r * lddy + ch] = d - tf32_rne(d);
    }
  }
}

// out[c] = sum_r X[r][c] in a fixed order (bias gradients of the linear output layers)
__global__ void __launch_bounds__(1024)
fc_colsum_kernel(int R, int C, const float* __restrict__ X, int ld, float* __restrict__ out) {
  pdl_wait();
  __shared__ float s_v[FC_LANES][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int ch = blockIdx.x * 32 + tx;
  float a = 0.f;
  if (ch < C)
    for (int r = ty; r < R; r += FC_LANES) a += X[(size_t)r * ld + ch];
  s_v[ty][tx] = a;
  __syncthreads();
  if (ty == 0 && ch < C) {
    float s = 0.f;
#pragma unroll 8
    for (int r = 0; r < FC_LANES; ++r) s += s_v[r][tx];
    out[ch] = s;
  }
}

// out = a + b + c (the three branches' gradients w.r.t. the embedding)
__global__ void add3_kernel(long n, const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                            float* __restrict__ out) {
  pdl_wait();
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x)
    out[e] = (a[e] + b[e]) + c[e];
}

}  // namespace caae

using namespace caae;

extern "C" int caae_fc_bn_fwd(int R, int C, const float* Y, int ld, const float* gamma, const float* beta,
                              float* ema_mean, float* ema_var, const float* decay, float* scale, float* shift,
                              float* save_mean, float* save_invstd, int relu, float* out, int ldo, float* out_lo,
                              caae_stream_t stream) {
  CAAE_RETURN_IF(R <= 0 || C <= 0 || ld < C || (out && ldo < C), CAAE_E_BADSHAPE);
  CAAE_RETURN_IF(!Y || !gamma || !beta || !scale || !shift || !save_mean || !save_invstd, CAAE_E_NULLPTR);
  CAAE_RETURN_IF((ema_mean == nullptr) != (ema_var == nullptr), CAAE_E_NULLPTR);
  caae::launch(fc_bn_fwd_kernel, (C + 31) / 32, dim3(32, FC_LANES), 0, as_stream(stream), 
      R, C, Y, ld, gamma, beta, ema_mean, ema_var, decay, scale, shift, save_mean, save_invstd, relu, out, ldo, out_lo);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_fc_bn_bwd(int R, int C, const float* Y, int ld, const float* scale, const float* shift,
                              const float* mean, const float* invstd, const float* gamma, int relu, const float* dOut,
                              int lddo, float* dY, int lddy, float* dgamma, float* dbeta, float* dY_lo,
                              caae_stream_t stream) {
  CAAE_RETURN_IF(R <= 0 || C <= 0 || ld < C || lddo < C || lddy < C, CAAE_E_BADSHAPE);
  CAAE_RETURN_IF(!Y || !scale || !shift || !mean || !invstd || !gamma || !dOut || !dY || !dgamma || !dbeta,
                 CAAE_E_NULLPTR);
  caae::launch(fc_bn_bwd_kernel, (C + 31) / 32, dim3(32, FC_LANES), 0, as_stream(stream), 
      R, C, Y, ld, scale, shift, mean, invstd, gamma, relu, dOut, lddo, dY, lddy, dgamma, dbeta, dY_lo);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_colsum(int R, int C, const float* X, int ld, float* out, caae_stream_t stream) {
  CAAE_RETURN_IF(R < 0 || C <= 0 || ld < C, CAAE_E_BADSHAPE);
  CAAE_RETURN_IF(!X || !out, CAAE_E_NULLPTR);
  caae::launch(fc_colsum_kernel, (C + 31) / 32, dim3(32, FC_LANES), 0, as_stream(stream), R, C, X, ld, out);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_add3(long n, const float* a, const float* b, const float* c, float* out, caae_stream_t stream) {
  CAAE_RETURN_IF(n < 0, CAAE_E_BADSHAPE);
  if (n == 0) return CAAE_OK;
  CAAE_RETURN_IF(!a || !b || !c || !out, CAAE_E_NULLPTR);
  long blocks = (n + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  caae::launch(add3_kernel, (int)blocks, 256, 0, as_stream(stream), n, a, b, c, out);
  return CAAE_LAUNCH_STATUS();
}
