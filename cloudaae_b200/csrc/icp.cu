// icp.cu — batched point-to-point ICP refinement of the predicted pose for sm_100a.
//
// Replaces the open3d loop of the reference's evaluation (evaluate_cloudAAE_ycbv.py:606-624):
//   for i in range(10): reg = registration_icp(model, segment, radius, T, PointToPoint); radius *= 0.9; T = reg.T
// open3d's registration_icp (pinned nowhere by the reference; algorithm as published):
//   pcd = T * source; result = correspondences(pcd, target, radius)
//   repeat <= max_iteration: U = umeyama(pcd[corr], target[corr]) (no scaling; identity when corr is empty);
//     T = U*T; pcd = U*pcd; result' = correspondences(...); stop when |dfitness| and |drmse| < 1e-6
//   correspondence of a source point = its nearest target point if the squared distance is < radius^2;
//   fitness = #corr / #source, inlier_rmse = sqrt(sum d^2 / #corr).
//
// Design: one CTA per segment, the WHOLE refinement (all rounds, all iterations) in one launch: the
// transformed model (fp64) and the segment live in shared memory, each thread searches the nearest
// target for its source points by brute force (256 targets: a KD-tree would only add divergence; an fp32 scan with a
// rigorous error margin picks the winner, float64 confirms it — the float64 pipe was the bound, ncu: 60 % active),
// the 17 Kabsch sums are reduced in a fixed order, and one thread solves the rotation with Horn's
// quaternion form (largest eigenvector of a symmetric 4x4, cyclic Jacobi in fp64) — the same optimum
// as Umeyama's SVD with its reflection guard whenever that optimum is unique.
#include "common.cuh"

namespace caae {

constexpr int kIcpThreads = 512;
constexpr int kIcpWarps = kIcpThreads / 32;
constexpr int kIcpSums = 17;  // n, sum d^2, Sx[3], Sy[3], Sxy[9]

// Largest-eigenvalue eigenvector of the symmetric 4x4 matrix a (destroyed) by cyclic Jacobi.
__device__ void jacobi4_max_eigvec(double a[4][4], double q[4]) {
  double v[4][4];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) v[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 40; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < 4; ++i) {
      diag += a[i][i] * a[i][i];
      for (int j = i + 1; j < 4; ++j) off += a[i][j] * a[i][j];
    }
    if (off <= 1e-34 * diag || off == 0.0) break;
    for (int p = 0; p < 3; ++p) {
      for (int r = p + 1; r < 4; ++r) {
        const double apr = a[p][r];
        if (apr == 0.0) continue;
        const double theta = (a[r][r] - a[p][p]) / (2.0 * apr);
        const double tt = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(tt * tt + 1.0), s = tt * c;
        for (int k = 0; k < 4; ++k) {  // A <- A * J
          const double akp = a[k][p], akr = a[k][r];
          a[k][p] = c * akp - s * akr;
          a[k][r] = s * akp + c * akr;
        }
        for (int k = 0; k < 4; ++k) {  // A <- J^T * A
          const double apk = a[p][k], ark = a[r][k];
          a[p][k] = c * apk - s * ark;
          a[r][k] = s * apk + c * ark;
        }
        for (int k = 0; k < 4; ++k) {  // V <- V * J
          const double vkp = v[k][p], vkr = v[k][r];
          v[k][p] = c * vkp - s * vkr;
          v[k][r] = s * vkp + c * vkr;
        }
      }
    }
  }
  int best = 0;
  for (int i = 1; i < 4; ++i)
    if (a[i][i] > a[best][best]) best = i;
  double nrm = 0.0;
  for (int k = 0; k < 4; ++k) nrm += v[k][best] * v[k][best];
  nrm = 1.0 / sqrt(nrm);
  for (int k = 0; k < 4; ++k) q[k] = v[k][best] * nrm;
}

// Rigid update (R, t) minimising sum |R x + t - y|^2 from the reduced sums (coordinates relative to c0).
__device__ void kabsch_from_sums(const double* S, const double c0[3], double R[9], double tr[3]) {
  const double n = S[0];
  if (n < 0.5) {  // no correspondence: identity (open3d returns Matrix4d::Identity())
    for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    tr[0] = tr[1] = tr[2] = 0.0;
    return;
  }
  const double inv = 1.0 / n;
  double mx[3], my[3], M[3][3];  // M = sum (x - mx)(y - my)^T
  for (int i = 0; i < 3; ++i) { mx[i] = S[2 + i] * inv; my[i] = S[5 + i] * inv; }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) M[i][j] = S[8 + i * 3 + j] - n * mx[i] * my[j];
  double N[4][4];
  N[0][0] = M[0][0] + M[1][1] + M[2][2];
  N[1][1] = M[0][0] - M[1][1] - M[2][2];
  N[2][2] = -M[0][0] + M[1][1] - M[2][2];
  N[3][3] = -M[0][0] - M[1][1] + M[2][2];
  N[0][1] = N[1][0] = M[1][2] - M[2][1];
  N[0][2] = N[2][0] = M[2][0] - M[0][2];
  N[0][3] = N[3][0] = M[0][1] - M[1][0];
  N[1][2] = N[2][1] = M[0][1] + M[1][0];
  N[1][3] = N[3][1] = M[2][0] + M[0][2];
  N[2][3] = N[3][2] = M[1][2] + M[2][1];
  double q[4];
  jacobi4_max_eigvec(N, q);
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = 1.0 - 2.0 * (y * y + z * z); R[1] = 2.0 * (x * y - w * z);       R[2] = 2.0 * (x * z + w * y);
  R[3] = 2.0 * (x * y + w * z);       R[4] = 1.0 - 2.0 * (x * x + z * z); R[5] = 2.0 * (y * z - w * x);
  R[6] = 2.0 * (x * z - w * y);       R[7] = 2.0 * (y * z + w * x);       R[8] = 1.0 - 2.0 * (x * x + y * y);
  // t = muy - R mux in absolute coordinates (mu = mu' + c0)
  for (int i = 0; i < 3; ++i) {
    const double ax = mx[0] + c0[0], ay = mx[1] + c0[1], az = mx[2] + c0[2];
    tr[i] = (my[i] + c0[i]) - (R[i * 3 + 0] * ax + R[i * 3 + 1] * ay + R[i * 3 + 2] * az);
  }
}

__global__ void __launch_bounds__(kIcpThreads)
icp_refine_kernel(int ns, int src_stride, const float* __restrict__ source, const int* __restrict__ source_of_seg,
                  int nt, const float* __restrict__ target, const double* __restrict__ T_init, double radius,
                  double radius_decay, int outer, int max_iter, double rel_fitness, double rel_rmse,
                  double* __restrict__ T_out, double* __restrict__ fitness_out, double* __restrict__ rmse_out,
                  int* __restrict__ iters_out) {
  extern __shared__ __align__(16) double s_dyn[];
  double* s_pcd = s_dyn;           // [ns][3] current transformed source
  double* s_tgt = s_dyn + ns * 3;  // [nt][3]
  // [nt] fp32 copy relative to c0 (nearest-target prefilter), 16-byte aligned whatever the parity of ns + nt
  float4* s_tgf = reinterpret_cast<float4*>((reinterpret_cast<uintptr_t>(s_tgt + nt * 3) + 15) & ~(uintptr_t)15);
  __shared__ double s_red[kIcpWarps][kIcpSums];
  __shared__ double s_sum[kIcpSums];
  __shared__ double s_T[12];   // accumulated transformation, rows of [R | t]
  __shared__ double s_U[12];   // last update
  __shared__ int s_flag;

  const int seg = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int src_id = source_of_seg != nullptr ? source_of_seg[seg] : seg;
  const float* __restrict__ src = source + (size_t)src_id * ns * src_stride;
  const float* __restrict__ tg = target + (size_t)seg * nt * 3;

  for (int i = t; i < nt * 3; i += kIcpThreads) s_tgt[i] = (double)tg[i];
  if (t < 12) s_T[t] = T_init[(size_t)seg * 16 + t];
  __syncthreads();
  const double c0[3] = {nt > 0 ? s_tgt[0] : 0.0, nt > 0 ? s_tgt[1] : 0.0, nt > 0 ? s_tgt[2] : 0.0};
  for (int j = t; j < nt; j += kIcpThreads)
    s_tgf[j] = make_float4((float)(s_tgt[j * 3 + 0] - c0[0]), (float)(s_tgt[j * 3 + 1] - c0[1]),
                           (float)(s_tgt[j * 3 + 2] - c0[2]), 0.f);
  __syncthreads();

  double fit = 0.0, rmse = 0.0;
  int total_iters = 0;
  double rad = radius;

  // One correspondence pass over the CTA's source points; leaves the 17 sums in s_sum.
  auto correspond = [&](double r2) {
    double acc[kIcpSums];
#pragma unroll
    for (int i = 0; i < kIcpSums; ++i) acc[i] = 0.0;
    for (int i = t; i < ns; i += kIcpThreads) {
      const double x = s_pcd[i * 3 + 0], y = s_pcd[i * 3 + 1], z = s_pcd[i * 3 + 2];
      double best = r2;
      int bj = -1;
      // fp32 prefilter: the two smallest squared distances over the targets.  When they are further apart than
      // the fp32 evaluation can be wrong, the fp32 winner IS the float64 nearest target and one exact evaluation
      // settles the radius test; otherwise (near-ties, duplicates) the exact float64 scan below decides.
      const float xf = (float)(x - c0[0]), yf = (float)(y - c0[1]), zf = (float)(z - c0[2]);
      float m1 = 3.0e38f, m2 = 3.0e38f;
      int j1 = -1;
      for (int j = 0; j < nt; ++j) {
        const float4 tg = s_tgf[j];
        const float ex = tg.x - xf, ey = tg.y - yf, ez = tg.z - zf;
        const float d = fmaf(ez, ez, fmaf(ey, ey, ex * ex));
        if (d < m1) { m2 = m1; m1 = d; j1 = j; } else m2 = fminf(m2, d);
      }
      // error of d: coordinates relative to c0 rounded to fp32 (|coordinate| < 8 m: 5e-7 each) and the fp32 arithmetic
      const float kDelta = 5e-7f;
      const float err = 4.f * kDelta * sqrtf(3.f * m1) + 12.f * kDelta * kDelta + 4e-7f * m1;
      const float err2 = 4.f * kDelta * sqrtf(3.f * m2) + 12.f * kDelta * kDelta + 4e-7f * m2;
      if (j1 >= 0 && m2 - m1 > err + err2) {
        const double dx = x - s_tgt[j1 * 3 + 0], dy = y - s_tgt[j1 * 3 + 1], dz = z - s_tgt[j1 * 3 + 2];
        const double d = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        if (d < r2) { best = d; bj = j1; }
      } else {
        for (int j = 0; j < nt; ++j) {
          const double dx = x - s_tgt[j * 3 + 0], dy = y - s_tgt[j * 3 + 1], dz = z - s_tgt[j * 3 + 2];
          const double d = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
          if (d < best) { best = d; bj = j; }  // strict <: first nearest, and strictly inside the radius
        }
      }
      if (bj >= 0) {
        const double xr = x - c0[0], yr = y - c0[1], zr = z - c0[2];
        const double u = s_tgt[bj * 3 + 0] - c0[0], v = s_tgt[bj * 3 + 1] - c0[1], w = s_tgt[bj * 3 + 2] - c0[2];
        acc[0] += 1.0; acc[1] += best;
        acc[2] += xr; acc[3] += yr; acc[4] += zr;
        acc[5] += u; acc[6] += v; acc[7] += w;
        acc[8] += xr * u; acc[9] += xr * v; acc[10] += xr * w;
        acc[11] += yr * u; acc[12] += yr * v; acc[13] += yr * w;
        acc[14] += zr * u; acc[15] += zr * v; acc[16] += zr * w;
      }
    }
#pragma unroll
    for (int i = 0; i < kIcpSums; ++i) {
      double v = acc[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) s_red[warp][i] = v;
    }
    __syncthreads();
    if (t < kIcpSums) {
      double v = 0.0;
      for (int wv = 0; wv < kIcpWarps; ++wv) v += s_red[wv][t];
      s_sum[t] = v;
    }
    __syncthreads();
  };

  for (int round = 0; round < outer; ++round) {
    // pcd = T * source (open3d transforms a fresh copy of the source by `init` at every call)
    for (int i = t; i < ns; i += kIcpThreads) {
      const double x = (double)src[(size_t)i * src_stride + 0], y = (double)src[(size_t)i * src_stride + 1],
                   z = (double)src[(size_t)i * src_stride + 2];
      s_pcd[i * 3 + 0] = s_T[0] * x + s_T[1] * y + s_T[2] * z + s_T[3];
      s_pcd[i * 3 + 1] = s_T[4] * x + s_T[5] * y + s_T[6] * z + s_T[7];
      s_pcd[i * 3 + 2] = s_T[8] * x + s_T[9] * y + s_T[10] * z + s_T[11];
    }
    __syncthreads();
    const double r2 = rad * rad;
    correspond(r2);
    fit = ns > 0 ? s_sum[0] / (double)ns : 0.0;
    rmse = s_sum[0] > 0.5 ? sqrt(s_sum[1] / s_sum[0]) : 0.0;
    for (int it = 0; it < max_iter; ++it) {
      if (t == 0) {
        double R[9], tr[3];
        kabsch_from_sums(s_sum, c0, R, tr);
        double Tn[12];
        for (int i = 0; i < 3; ++i) {
          for (int j = 0; j < 4; ++j)
            Tn[i * 4 + j] = R[i * 3 + 0] * s_T[0 * 4 + j] + R[i * 3 + 1] * s_T[1 * 4 + j] + R[i * 3 + 2] * s_T[2 * 4 + j];
          Tn[i * 4 + 3] += tr[i];
        }
        for (int i = 0; i < 12; ++i) s_T[i] = Tn[i];
        for (int i = 0; i < 3; ++i) {
          s_U[i * 4 + 0] = R[i * 3 + 0]; s_U[i * 4 + 1] = R[i * 3 + 1]; s_U[i * 4 + 2] = R[i * 3 + 2];
          s_U[i * 4 + 3] = tr[i];
        }
      }
      __syncthreads();
      for (int i = t; i < ns; i += kIcpThreads) {
        const double x = s_pcd[i * 3 + 0], y = s_pcd[i * 3 + 1], z = s_pcd[i * 3 + 2];
        s_pcd[i * 3 + 0] = s_U[0] * x + s_U[1] * y + s_U[2] * z + s_U[3];
        s_pcd[i * 3 + 1] = s_U[4] * x + s_U[5] * y + s_U[6] * z + s_U[7];
        s_pcd[i * 3 + 2] = s_U[8] * x + s_U[9] * y + s_U[10] * z + s_U[11];
      }
      __syncthreads();
      correspond(r2);
      const double nfit = ns > 0 ? s_sum[0] / (double)ns : 0.0;
      const double nrmse = s_sum[0] > 0.5 ? sqrt(s_sum[1] / s_sum[0]) : 0.0;
      const bool done = fabs(fit - nfit) < rel_fitness && fabs(rmse - nrmse) < rel_rmse;
      fit = nfit; rmse = nrmse;
      ++total_iters;
      if (t == 0) s_flag = done ? 1 : 0;  // one thread decides, so the branch is uniform by construction
      __syncthreads();
      const int stop = s_flag;
      __syncthreads();
      if (stop) break;
    }
    rad *= radius_decay;
  }
  if (t < 12) T_out[(size_t)seg * 16 + t] = s_T[t];
  if (t >= 12 && t < 16) T_out[(size_t)seg * 16 + t] = t == 15 ? 1.0 : 0.0;
  if (t == 0) {
    if (fitness_out != nullptr) fitness_out[seg] = fit;
    if (rmse_out != nullptr) rmse_out[seg] = rmse;
    if (iters_out != nullptr) iters_out[seg] = total_iters;
  }
}

}  // namespace caae

using namespace caae;

extern "C" int caae_icp_refine(int b, int ns, int src_stride, const float* source, const int* source_of_seg, int nt,
                               const float* target, const double* T_init, double radius, double radius_decay,
                               int outer, int max_iter, double rel_fitness, double rel_rmse, double* T_out,
                               double* fitness, double* inlier_rmse, int* iterations, caae_stream_t stream) {
  CAAE_RETURN_IF(b < 0 || ns < 0 || nt < 0 || src_stride < 3 || outer < 0 || max_iter < 0, CAAE_E_BADSHAPE);
  if (b == 0) return CAAE_OK;
  CAAE_RETURN_IF(T_init == nullptr || T_out == nullptr, CAAE_E_NULLPTR);
  CAAE_RETURN_IF((ns > 0 && source == nullptr) || (nt > 0 && target == nullptr), CAAE_E_NULLPTR);
  const size_t smem = ((size_t)ns + (size_t)nt) * 3 * sizeof(double) + (size_t)nt * sizeof(float4) + 16;
  CAAE_RETURN_IF(smem > 200 * 1024, CAAE_E_UNSUPPORTED);
  static size_t smem_set = 0;   // opt in once per size class (one process per GPU), not on every call
  if (smem > 48 * 1024 && smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(icp_refine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    smem_set = smem;
  }
  icp_refine_kernel<<<b, kIcpThreads, smem, as_stream(stream)>>>(
      ns, src_stride, source, source_of_seg, nt, target, T_init, radius, radius_decay, outer, max_iter, rel_fitness,
      rel_rmse, T_out, fitness, inlier_rmse, iterations);
  return CAAE_LAUNCH_STATUS();
}
