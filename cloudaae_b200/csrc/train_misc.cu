// train_misc.cu — input preparation, pose losses (float64), loss reduction, TF-style Adam and the
// device-side step/schedule state that makes a whole training step CUDA-graph replayable.
//
// Reference semantics:
//   input prep   train_cloudAAE_ycbv.py:206-226 (slice, + noise, - per-cloud mean, one-hot concat)
//   rotation     losses/angular_distance_taylor.py:30-116 in float64 (Rodrigues with the theta^2 < 1e-2
//                Taylor guard; theta = acos(clip((tr(R_l R_p^T)-1)/2, +-0.9999999)))
//   translation  losses/trans_distance.py:4-9;  chamfer mean  losses/chamfer_loss.py:12-13
//   total        1000*chamfer + 10*trans + rot (train_cloudAAE_ycbv.py:268)
//   optimiser    tf.train.AdamOptimizer(0.0008): lr_t = lr*sqrt(1-b2^t)/(1-b1^t),
//                p -= lr_t*m/(sqrt(v)+eps)  (epsilon outside the bias correction)
//   bn_decay     min(0.99, 1 - 0.5*0.5^floor(step*B/40)) (train_cloudAAE_ycbv.py:166-169,196-202)
#include "common.cuh"

namespace caae {

// ---- forward-mode dual numbers over the 3 axis-angle components (exact gradient incl. branches)
struct Dual {
  double v, g[3];
};
__device__ __forceinline__ Dual dconst(double v) { return Dual{v, {0.0, 0.0, 0.0}}; }
__device__ __forceinline__ Dual dvar(double v, int i) { Dual d = dconst(v); d.g[i] = 1.0; return d; }
__device__ __forceinline__ Dual operator+(Dual a, Dual b) { return Dual{a.v + b.v, {a.g[0] + b.g[0], a.g[1] + b.g[1], a.g[2] + b.g[2]}}; }
__device__ __forceinline__ Dual operator-(Dual a, Dual b) { return Dual{a.v - b.v, {a.g[0] - b.g[0], a.g[1] - b.g[1], a.g[2] - b.g[2]}}; }
__device__ __forceinline__ Dual operator-(Dual a) { return Dual{-a.v, {-a.g[0], -a.g[1], -a.g[2]}}; }
__device__ __forceinline__ Dual operator*(Dual a, Dual b) {
  return Dual{a.v * b.v, {a.g[0] * b.v + a.v * b.g[0], a.g[1] * b.v + a.v * b.g[1], a.g[2] * b.v + a.v * b.g[2]}};
}
__device__ __forceinline__ Dual operator*(double s, Dual a) { return Dual{s * a.v, {s * a.g[0], s * a.g[1], s * a.g[2]}}; }
__device__ __forceinline__ Dual operator/(Dual a, Dual b) {
  const double q = a.v / b.v, ib = 1.0 / b.v;
  return Dual{q, {(a.g[0] - q * b.g[0]) * ib, (a.g[1] - q * b.g[1]) * ib, (a.g[2] - q * b.g[2]) * ib}};
}
__device__ __forceinline__ Dual dchain(Dual a, double f, double df) { return Dual{f, {df * a.g[0], df * a.g[1], df * a.g[2]}}; }

// R = I + t1*K + t2*K^2, K = skew(a)  (angular_distance_taylor.py:30-66)
template <class T>
__device__ __forceinline__ void expmap(const T a[3], T t1, T t2, T one, T R[3][3]) {
  const T x = a[0], y = a[1], z = a[2];
  // K^2 = a a^T - |a|^2 I
  const T xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z;
  R[0][0] = one - t2 * (yy + zz); R[0][1] = -(t1 * z) + t2 * xy;   R[0][2] = t1 * y + t2 * xz;
  R[1][0] = t1 * z + t2 * xy;     R[1][1] = one - t2 * (xx + zz); R[1][2] = -(t1 * x) + t2 * yz;
  R[2][0] = -(t1 * y) + t2 * xz;  R[2][1] = t1 * x + t2 * yz;     R[2][2] = one - t2 * (xx + yy);
}

__device__ __forceinline__ void exp_terms(double tsq, double& t1, double& t2) {
  if (tsq < 1e-2) {
    const double t4 = tsq * tsq, t6 = t4 * tsq, t8 = t4 * t4;
    t1 = 1 - (tsq / 6) + (t4 / 120) - (t6 / 5040) + (t8 / 362880);
    t2 = 0.5 - (tsq / 24) + (t4 / 720) - (t6 / 40320) + (t8 / 3628800);
  } else {
    const double th = sqrt(tsq);
    t1 = sin(th) / th;
    t2 = (1 - cos(th)) / tsq;
  }
}

__device__ __forceinline__ void exp_terms(Dual tsq, Dual& t1, Dual& t2) {
  if (tsq.v < 1e-2) {
    const Dual t4 = tsq * tsq, t6 = t4 * tsq, t8 = t4 * t4;
    t1 = dconst(1.0) - (1.0 / 6) * tsq + (1.0 / 120) * t4 - (1.0 / 5040) * t6 + (1.0 / 362880) * t8;
    t2 = dconst(0.5) - (1.0 / 24) * tsq + (1.0 / 720) * t4 - (1.0 / 40320) * t6 + (1.0 / 3628800) * t8;
  } else {
    const double thv = sqrt(tsq.v);
    const Dual th = dchain(tsq, thv, 0.5 / thv);
    const Dual s = dchain(th, sin(thv), cos(thv));
    const Dual c = dchain(th, cos(thv), -sin(thv));
    t1 = s / th;
    t2 = (dconst(1.0) - c) / tsq;
  }
}

// one thread per sample
__global__ void pose_loss_kernel(int b, const float* __restrict__ rot_pred, const float* __restrict__ axag_label,
                                 const float* __restrict__ trans_res, const float* __restrict__ mean,
                                 const float* __restrict__ trans_label, float w_rot, float w_trans,
                                 double* __restrict__ per_rot, float* __restrict__ per_trans,
                                 float* __restrict__ d_rot, float* __restrict__ d_trans,
                                 float* __restrict__ trans_pred_out) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b) return;
  // ---- rotation (float64)
  double al[3];
  Dual ap[3];
  for (int c = 0; c < 3; ++c) { al[c] = (double)axag_label[i * 3 + c]; ap[c] = dvar((double)rot_pred[i * 3 + c], c); }
  double Rl[3][3], l1, l2;
  exp_terms(al[0] * al[0] + al[1] * al[1] + al[2] * al[2], l1, l2);
  expmap<double>(al, l1, l2, 1.0, Rl);
  Dual Rp[3][3], p1, p2;
  exp_terms(ap[0] * ap[0] + ap[1] * ap[1] + ap[2] * ap[2], p1, p2);
  expmap<Dual>(ap, p1, p2, dconst(1.0), Rp);
  Dual tr = dconst(0.0);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) tr = tr + Rl[r][c] * Rp[r][c];  // trace(R_l R_p^T)
  Dual x = 0.5 * (tr - dconst(1.0));
  const double lim = 0.9999999;
  if (x.v < -lim) x = dconst(-lim);          // clip_by_value: zero gradient outside the range
  else if (x.v > lim) x = dconst(lim);
  const Dual theta = dchain(x, acos(x.v), -1.0 / sqrt(1.0 - x.v * x.v));
  per_rot[i] = theta.v;
  for (int c = 0; c < 3; ++c) d_rot[i * 3 + c] = (float)(theta.g[c] * (double)w_rot);
  // ---- translation (float32)
  float diff[3], ss = 0.f;
  for (int c = 0; c < 3; ++c) {
    const float pred = trans_res[i * 3 + c] + mean[i * 3 + c];
    if (trans_pred_out) trans_pred_out[i * 3 + c] = pred;
    diff[c] = trans_label[i * 3 + c] - pred;
    ss += diff[c] * diff[c];
  }
  const float nrm = sqrtf(ss);
  per_trans[i] = nrm;
  for (int c = 0; c < 3; ++c) d_trans[i * 3 + c] = (nrm > 0.f) ? (-diff[c] / nrm) * w_trans : 0.f;
}

// single CTA, fixed order: losses = [total, chamfer, trans, rot]
__global__ void __launch_bounds__(1024)
loss_reduce_kernel(long npt, const float* __restrict__ dist1, const float* __restrict__ dist2, int b,
                   const float* __restrict__ per_trans, const double* __restrict__ per_rot,
                   float* __restrict__ losses) {
  pdl_wait();
  __shared__ double s_red[3][32];
  double c = 0.0, t = 0.0, r = 0.0;
  for (long e = threadIdx.x; e < npt; e += 1024) c += (double)(dist1[e] + dist2[e]);
  for (int e = threadIdx.x; e < b; e += 1024) { t += (double)per_trans[e]; r += per_rot[e]; }
  for (int o = 16; o > 0; o >>= 1) {
    c += __shfl_xor_sync(0xffffffffu, c, o);
    t += __shfl_xor_sync(0xffffffffu, t, o);
    r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { s_red[0][warp] = c; s_red[1][warp] = t; s_red[2][warp] = r; }
  __syncthreads();
  if (threadIdx.x == 0) {
    c = t = r = 0.0;
    for (int w = 0; w < 32; ++w) { c += s_red[0][w]; t += s_red[1][w]; r += s_red[2][w]; }
    const float chamfer = (float)(c / (double)npt), trans = (float)(t / b), rot = (float)(r / b);
    losses[1] = chamfer; losses[2] = trans; losses[3] = rot;
    losses[0] = 1000.f * chamfer + 10.f * trans + rot;
  }
}

// out[b,p,:] = in[b,p,:] + v[b,:]
__global__ void add_cloud_vec_kernel(long total, int npts, const float* __restrict__ in, const float* __restrict__ v,
                                     float* __restrict__ out) {
  pdl_wait();
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const long cloud = e / (3L * npts);
    out[e] = in[e] + v[cloud * 3 + (e % 3)];
  }
}

// one CTA per cloud: x[b,p,0:3] = vis[b,p,:] + noise[b,p,:] - mean_b ; x[b,p,3+c] = onehot(class_id[b])
__global__ void __launch_bounds__(256)
prepare_input_kernel(int npoint, int vis_stride_pts, const float* __restrict__ visible, const float* __restrict__ noise,
                     const int* __restrict__ class_id, int nclass, float* __restrict__ x, float* __restrict__ mean_out) {
  pdl_wait();
  __shared__ float s_sum[3][8];
  __shared__ float s_mean[3];
  const int cloud = blockIdx.x, tid = threadIdx.x;
  const float* __restrict__ vis = visible + (size_t)cloud * vis_stride_pts * 3;
  const float* __restrict__ nz = noise ? noise + (size_t)cloud * npoint * 3 : nullptr;
  float s[3] = {0.f, 0.f, 0.f};
  for (int p = tid; p < npoint; p += 256)
    for (int c = 0; c < 3; ++c) s[c] += vis[p * 3 + c] + (nz ? nz[p * 3 + c] : 0.f);
  for (int c = 0; c < 3; ++c) {
    for (int o = 16; o > 0; o >>= 1) s[c] += __shfl_xor_sync(0xffffffffu, s[c], o);
    if ((tid & 31) == 0) s_sum[c][tid >> 5] = s[c];
  }
  __syncthreads();
  if (tid < 3) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_sum[tid][w];
    s_mean[tid] = t / (float)npoint;
    mean_out[cloud * 3 + tid] = s_mean[tid];
  }
  __syncthreads();
  const int d = 3 + nclass;
  const int cls = class_id[cloud];
  float* __restrict__ xo = x + (size_t)cloud * npoint * d;
  for (int e = tid; e < npoint * d; e += 256) {
    const int p = e / d, c = e - p * d;
    float v;
    if (c < 3) v = vis[p * 3 + c] + (nz ? nz[p * 3 + c] : 0.f) - s_mean[c];
    else v = (c - 3 == cls) ? 1.f : 0.f;
    xo[e] = v;
  }
}

// state = {int step, int adam_t, float bn_decay}: called once at the start of every training step
__global__ void step_begin_kernel(int* __restrict__ state, int batch_size) {
  pdl_wait();
  const int step = state[0];
  const double mom = 0.5 * pow(0.5, floor((double)step * (double)batch_size / 40.0));
  float decay = (float)(1.0 - mom);
  reinterpret_cast<float*>(state)[2] = fminf(0.99f, decay);
  state[1] = step + 1;
  state[0] = step + 1;
}

__global__ void __launch_bounds__(256)
adam_tf_kernel(long n, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
               const int* __restrict__ state, float lr, float beta1, float beta2, float eps, float grad_scale) {
  pdl_wait();
  __shared__ float s_lr;
  if (threadIdx.x == 0) {
    const double t = (double)state[1];
    s_lr = (float)((double)lr * sqrt(1.0 - pow((double)beta2, t)) / (1.0 - pow((double)beta1, t)));
  }
  __syncthreads();
  const float lr_t = s_lr;
  const long stride = (long)gridDim.x * blockDim.x, t0 = (long)blockIdx.x * blockDim.x + threadIdx.x;
  // 16-byte accesses over the aligned bulk of the flat buffers (7 M elements, 4 streams in, 3 out), scalar tail
  const bool vec = (((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                      reinterpret_cast<uintptr_t>(v)) & 15) == 0);
  const long n4 = vec ? n / 4 : 0;
  for (long q = t0; q < n4; q += stride) {
    const float4 g4 = reinterpret_cast<const float4*>(g)[q];
    float4 m4 = reinterpret_cast<float4*>(m)[q], v4 = reinterpret_cast<float4*>(v)[q], p4 = reinterpret_cast<float4*>(p)[q];
    const float gs[4] = {g4.x * grad_scale, g4.y * grad_scale, g4.z * grad_scale, g4.w * grad_scale};
    float ms[4] = {m4.x, m4.y, m4.z, m4.w}, vs[4] = {v4.x, v4.y, v4.z, v4.w}, ps[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      ms[c] = beta1 * ms[c] + (1.f - beta1) * gs[c];
      vs[c] = beta2 * vs[c] + (1.f - beta2) * gs[c] * gs[c];
      ps[c] -= lr_t * ms[c] / (sqrtf(vs[c]) + eps);
    }
    reinterpret_cast<float4*>(m)[q] = make_float4(ms[0], ms[1], ms[2], ms[3]);
    reinterpret_cast<float4*>(v)[q] = make_float4(vs[0], vs[1], vs[2], vs[3]);
    reinterpret_cast<float4*>(p)[q] = make_float4(ps[0], ps[1], ps[2], ps[3]);
  }
  for (long e = n4 * 4 + t0; e < n; e += stride) {
    const float gv = g[e] * grad_scale;
    const float mv = beta1 * m[e] + (1.f - beta1) * gv;
    const float vv = beta2 * v[e] + (1.f - beta2) * gv * gv;
    m[e] = mv; v[e] = vv;
    p[e] -= lr_t * mv / (sqrtf(vv) + eps);
  }
}

__global__ void fill_kernel(long n, float* __restrict__ p, float value) {
  pdl_wait();
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) p[e] = value;
}

static inline int flat_blocks2(long total) {
  long blocks = (total + 255) / 256;
  const long cap = (long)kNumSMs * 16;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace caae

using namespace caae;

extern "C" int caae_pose_losses(int b, const float* rot_pred, const float* axag_label, const float* trans_res,
                                const float* mean, const float* trans_label, float w_rot, float w_trans,
                                double* per_rot, float* per_trans, float* d_rot, float* d_trans, float* trans_pred,
                                caae_stream_t stream) {
  CAAE_RETURN_IF(b < 0, CAAE_E_BADSHAPE);
  if (b == 0) return CAAE_OK;
  CAAE_RETURN_IF(!rot_pred || !axag_label || !trans_res || !mean || !trans_label || !per_rot || !per_trans || !d_rot ||
                 !d_trans, CAAE_E_NULLPTR);
  caae::launch(pose_loss_kernel, (b + 63) / 64, 64, 0, as_stream(stream), b, rot_pred, axag_label, trans_res, mean, trans_label,
                                                                w_rot, w_trans, per_rot, per_trans, d_rot, d_trans,
                                                                trans_pred);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_loss_reduce(long npt, const float* dist1, const float* dist2, int b, const float* per_trans,
                                const double* per_rot, float* losses, caae_stream_t stream) {
  CAAE_RETURN_IF(npt <= 0 || b <= 0, CAAE_E_BADSHAPE);
  CAAE_RETURN_IF(!dist1 || !dist2 || !per_trans || !per_rot || !losses, CAAE_E_NULLPTR);
  caae::launch(loss_reduce_kernel, 1, 1024, 0, as_stream(stream), npt, dist1, dist2, b, per_trans, per_rot, losses);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_add_cloud_vec(int b, int npts, const float* in, const float* v, float* out, caae_stream_t stream) {
  CAAE_RETURN_IF(b < 0 || npts < 0, CAAE_E_BADSHAPE);
  if (b == 0 || npts == 0) return CAAE_OK;
  CAAE_RETURN_IF(!in || !v || !out, CAAE_E_NULLPTR);
  const long total = (long)b * npts * 3;
  caae::launch(add_cloud_vec_kernel, flat_blocks2(total), 256, 0, as_stream(stream), total, npts, in, v, out);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_prepare_input(int b, int npoint, int vis_stride_pts, const float* visible, const float* noise,
                                  const int* class_id, int nclass, float* x, float* mean_out, caae_stream_t stream) {
  CAAE_RETURN_IF(b < 0 || npoint <= 0 || vis_stride_pts < npoint || nclass < 0, CAAE_E_BADSHAPE);
  if (b == 0) return CAAE_OK;
  CAAE_RETURN_IF(!visible || !class_id || !x || !mean_out, CAAE_E_NULLPTR);
  caae::launch(prepare_input_kernel, b, 256, 0, as_stream(stream), npoint, vis_stride_pts, visible, noise, class_id, nclass, x,
                                                        mean_out);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_step_begin(int* state, int batch_size, caae_stream_t stream) {
  CAAE_RETURN_IF(!state, CAAE_E_NULLPTR);
  caae::launch(step_begin_kernel, 1, 1, 0, as_stream(stream), state, batch_size);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_adam_tf(long n, float* p, const float* g, float* m, float* v, const int* state, float lr,
                            float beta1, float beta2, float eps, float grad_scale, caae_stream_t stream) {
  CAAE_RETURN_IF(n < 0, CAAE_E_BADSHAPE);
  if (n == 0) return CAAE_OK;
  CAAE_RETURN_IF(!p || !g || !m || !v || !state, CAAE_E_NULLPTR);
  caae::launch(adam_tf_kernel, flat_blocks2(n), 256, 0, as_stream(stream), n, p, g, m, v, state, lr, beta1, beta2, eps,
                                                                grad_scale);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_fill_f32(long n, float* p, float value, caae_stream_t stream) {
  CAAE_RETURN_IF(n < 0, CAAE_E_BADSHAPE);
  if (n == 0) return CAAE_OK;
  CAAE_RETURN_IF(!p, CAAE_E_NULLPTR);
  caae::launch(fill_kernel, flat_blocks2(n), 256, 0, as_stream(stream), n, p, value);
  return CAAE_LAUNCH_STATUS();
}
