// pose_metrics.cu — ADD / ADD-S pose-error metrics for a batch of (frame, class) segments (the
// accuracy step behind the ICP refinement of evaluate_cloudAAE_ycbv.py:606-624; the predicted pose is the
// [rotmat | trans] formed at :571-575 or the ICP result :615-624).
//
//   ADD   = mean_x | (R x + t) - (R^ x + t^) |                 (Hinterstoisser et al. 2012)
//   ADD-S = mean_x min_y | (R x + t) - (R^ y + t^) |           (symmetric objects; Xiang et al. 2018, PoseCNN)
// over the n model points x of the segment's class.  Three launches, every one parallel over segments:
//   caae_pose_transform_models   both posed copies of the model (float64 pose, float32 points)
//   caae_nn_distance             the nearest-neighbour search of ADD-S = the chamfer kernel (squared distances)
//   caae_add_reduce              ADD from matching points, ADD-S from sqrt(dist1), fixed-order fp64 means
#include "common.cuh"

namespace caae {

__global__ void pose_transform_models_kernel(int n, int src_stride, const float* __restrict__ models,
                                             const int* __restrict__ class_of_seg, const double* __restrict__ T_gt,
                                             const double* __restrict__ T_pred, float* __restrict__ out_gt,
                                             float* __restrict__ out_pred) {
  const int seg = blockIdx.y;
  const int cls = class_of_seg ? class_of_seg[seg] : seg;
  __shared__ double Tg[12], Tp[12];
  if (threadIdx.x < 12) { Tg[threadIdx.x] = T_gt[seg * 16 + threadIdx.x]; Tp[threadIdx.x] = T_pred[seg * 16 + threadIdx.x]; }
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float* p = models + ((size_t)cls * n + i) * src_stride;
    const double x = p[0], y = p[1], z = p[2];
    float* og = out_gt + ((size_t)seg * n + i) * 3;
    float* op = out_pred + ((size_t)seg * n + i) * 3;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      og[r] = (float)(Tg[4 * r] * x + Tg[4 * r + 1] * y + Tg[4 * r + 2] * z + Tg[4 * r + 3]);
      op[r] = (float)(Tp[4 * r] * x + Tp[4 * r + 1] * y + Tp[4 * r + 2] * z + Tp[4 * r + 3]);
    }
  }
}

// one CTA per segment; fixed-order tree (deterministic)
__global__ void __launch_bounds__(256)
add_reduce_kernel(int n, const float* __restrict__ gt, const float* __restrict__ pred, const float* __restrict__ dist_sq,
                  double* __restrict__ add, double* __restrict__ adds) {
  __shared__ double s_a[256], s_s[256];
  const int seg = blockIdx.x, tid = threadIdx.x;
  double a = 0.0, s = 0.0;
  for (int i = tid; i < n; i += 256) {
    const float* g = gt + ((size_t)seg * n + i) * 3;
    const float* p = pred + ((size_t)seg * n + i) * 3;
    const double dx = (double)g[0] - (double)p[0], dy = (double)g[1] - (double)p[1], dz = (double)g[2] - (double)p[2];
    a += sqrt(dx * dx + dy * dy + dz * dz);
    s += sqrt((double)dist_sq[(size_t)seg * n + i]);
  }
  s_a[tid] = a; s_s[tid] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) { s_a[tid] += s_a[tid + o]; s_s[tid] += s_s[tid + o]; }
    __syncthreads();
  }
  if (tid == 0) { add[seg] = s_a[0] / n; adds[seg] = s_s[0] / n; }
}

}  // namespace caae

using namespace caae;

extern "C" int caae_pose_transform_models(int b, int n, int src_stride, const float* models, const int* class_of_seg,
                                          const double* T_gt, const double* T_pred, float* out_gt, float* out_pred,
                                          caae_stream_t stream) {
  CAAE_RETURN_IF(b < 0 || n <= 0 || src_stride < 3 || b > 65535, CAAE_E_BADSHAPE);
  if (b == 0) return CAAE_OK;
  CAAE_RETURN_IF(!models || !T_gt || !T_pred || !out_gt || !out_pred, CAAE_E_NULLPTR);
  dim3 grid((n + 255) / 256, b);
  pose_transform_models_kernel<<<grid, 256, 0, as_stream(stream)>>>(n, src_stride, models, class_of_seg, T_gt, T_pred, out_gt,
                                                                     out_pred);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_add_reduce(int b, int n, const float* gt, const float* pred, const float* dist_sq, double* add,
                               double* adds, caae_stream_t stream) {
  CAAE_RETURN_IF(b < 0 || n <= 0, CAAE_E_BADSHAPE);
  if (b == 0) return CAAE_OK;
  CAAE_RETURN_IF(!gt || !pred || !dist_sq || !add || !adds, CAAE_E_NULLPTR);
  add_reduce_kernel<<<b, 256, 0, as_stream(stream)>>>(n, gt, pred, dist_sq, add, adds);
  return CAAE_LAUNCH_STATUS();
}
