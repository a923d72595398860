// synthesis.cu — on-line segment synthesis on the GPU: pose transform, spherical occluders, spherical
// flip and hidden-point-removal visibility, visible-prefix selection; plus the Philox normal/uniform
// generator that feeds it.
//
// Reference (all per SAMPLE on CPU threads, the hull inside a tf.py_func holding the GIL):
//   train_cloudAAE_ycbv.py:79-93          get_rotation_matrix / transform_object_model
//   utils/generate_occluder.py:38-81      get_random_spherical_occluder('ycbv')
//   utils/hidden_point_removal.py:6-73    sphericalFlip[_org], convexHull (scipy Qhull), padding
//
// Hidden point removal without a hull data structure.  The reference flips every point p to
// f = p + 2(R-|p|)p/|p|, appends the viewpoint (origin) and calls Qhull; the visible points are the
// hull vertices.  All flipped points have z > 0, so the projective map (x,y,z) -> (x/z, y/z, -1/z)
// is defined on them, preserves convexity and sends the origin to the point at infinity in -w:
// the vertices of conv(F u {0}) are exactly the vertices of the UPPER hull of the lifted points
// (u, v, w).  Point i is such a vertex iff a plane through it lies above every other lifted point:
//        exists s in R^2 :  s . (u_j - u_i, v_j - v_i)  >=  w_j - w_i      for all j != i.
// After subtracting the paraboloid  -rho/2 |(u,v)|^2  of the reference sphere (an exact change of
// variables, rho = max |f|) s = 0 means "tangent to the sphere", distant points can never violate
// for small s, and the test becomes the LP-type problem "minimum-norm s subject to n half-planes",
// solved per point by Seidel's incremental algorithm: keep the current optimum; a violated
// constraint moves it to the minimum-norm point of that constraint's boundary line clipped by all
// constraints seen before; an empty clip interval proves the point hidden.  Constraints are visited
// near-to-far through a 16x16 grid over (u,v) so re-solves happen while "seen before" is tiny.
// fp64 predicates on the fp32 flipped coordinates; measured against Qhull on the reference's own
// fixtures the visible sets are identical (tests/test_gpu_synthesis.py reports the IoU).
//
// One CTA (1024 threads) per cloud; everything a cloud needs lives in shared memory.
#include "common.cuh"

namespace caae {

constexpr int SY_THREADS = 1024;
constexpr int SY_G = 16;                 // grid cells per axis over the (u,v) bounding box
constexpr int SY_NOFF = (2 * SY_G - 1) * (2 * SY_G - 1);
constexpr int SY_MAXN = 4096;            // points per cloud the HPR kernel supports

struct OffTable { signed char dx[SY_NOFF], dy[SY_NOFF]; };

// cell offsets sorted by Chebyshev ring (near-to-far), built at compile time
constexpr OffTable make_off_table() {
  OffTable t{};
  int m = 0;
  for (int r = 0; r < SY_G; ++r)
    for (int y = -r; y <= r; ++y)
      for (int x = -r; x <= r; ++x) {
        const int ax = x < 0 ? -x : x, ay = y < 0 ? -y : y;
        if ((ax > ay ? ax : ay) != r) continue;
        t.dx[m] = (signed char)x; t.dy[m] = (signed char)y; ++m;
      }
  return t;
}
__constant__ OffTable c_off = make_off_table();

// ---- Philox4x32-10 -----------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0; key.y += W1;
  }
  return ctr;
}
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }  // (0,1)

// out[i] ~ N(0,1) (mode 0) or U(0,1) (mode 1); counter = (i/4, stream, *offset) so a replayed CUDA
// graph draws fresh numbers by bumping the device-side offset.
__global__ void philox_fill_kernel(long n, float* __restrict__ out, uint64_t seed, uint32_t stream_id,
                                   const int* __restrict__ offset, int mode) {
  const uint32_t off = offset ? (uint32_t)*offset : 0u;
  for (long q = (long)blockIdx.x * blockDim.x + threadIdx.x; q * 4 < n; q += (long)gridDim.x * blockDim.x) {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)q, (uint32_t)(q >> 32), stream_id, off),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    float v[4];
    if (mode == 0) {
      const float a0 = sqrtf(-2.f * logf(u01(r.x))), a1 = sqrtf(-2.f * logf(u01(r.z)));
      float s0, c0, s1, c1;
      sincospif(2.f * u01(r.y), &s0, &c0);
      sincospif(2.f * u01(r.w), &s1, &c1);
      v[0] = a0 * c0; v[1] = a0 * s0; v[2] = a1 * c1; v[3] = a1 * s1;
    } else {
      v[0] = u01(r.x); v[1] = u01(r.y); v[2] = u01(r.z); v[3] = u01(r.w);
    }
    for (int e = 0; e < 4; ++e)
      if (q * 4 + e < n) out[q * 4 + e] = v[e];
  }
}

// ---- pose transform + occluder + spherical flip -------------------------------------------------
__device__ __forceinline__ float block_max(float v, float* s_red) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = s_red[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = fmaxf(r, s_red[w]);
  __syncthreads();
  return r;
}

// points [b, nm+no, 3] = (model[class] R^T + t) followed by the occluder; flip_all over all of them,
// flip_org over the nm model points only (its own radius).  One CTA per cloud.
__global__ void __launch_bounds__(SY_THREADS)
synth_points_kernel(int nm, int no, const float* __restrict__ models, const int* __restrict__ class_id,
                    const float* __restrict__ axisangle, const float* __restrict__ translation,
                    const float* __restrict__ z_centers, const float* __restrict__ z_points, float hnear, float wnear,
                    float near_dist, float flip_pow, float* __restrict__ points, float* __restrict__ flip_all,
                    float* __restrict__ flip_org) {
  __shared__ float s_R[9], s_t[3], s_c[6], s_red[32];
  const int cloud = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) {
    // R = f32(expmap_f64(axag))  (losses/angular_distance_taylor.py:30-66)
    double a[3];
    for (int c = 0; c < 3; ++c) a[c] = (double)axisangle[cloud * 3 + c];
    const double tsq = a[0] * a[0] + a[1] * a[1] + a[2] * a[2];
    double t1, t2;
    if (tsq < 1e-2) {
      const double t4 = tsq * tsq, t6 = t4 * tsq, t8 = t4 * t4;
      t1 = 1 - (tsq / 6) + (t4 / 120) - (t6 / 5040) + (t8 / 362880);
      t2 = 0.5 - (tsq / 24) + (t4 / 720) - (t6 / 40320) + (t8 / 3628800);
    } else {
      const double th = sqrt(tsq);
      t1 = sin(th) / th; t2 = (1 - cos(th)) / tsq;
    }
    const double x = a[0], y = a[1], z = a[2];
    const double R[9] = {1 - t2 * (y * y + z * z), -t1 * z + t2 * x * y, t1 * y + t2 * x * z,
                         t1 * z + t2 * x * y, 1 - t2 * (x * x + z * z), -t1 * x + t2 * y * z,
                         -t1 * y + t2 * x * z, t1 * x + t2 * y * z, 1 - t2 * (x * x + y * y)};
    for (int e = 0; e < 9; ++e) s_R[e] = (float)R[e];
    for (int c = 0; c < 3; ++c) s_t[c] = translation[cloud * 3 + c];
    // occluder centres (generate_occluder.py:63-68)
    if (no > 0) {
      const float tz = s_t[2];
      const float mean_z = (near_dist + tz) / 2.f, std_z = (tz - near_dist) / 6.f;
      for (int o = 0; o < 2; ++o) {
        s_c[o * 3 + 0] = z_centers[(cloud * 2 + o) * 3 + 0] * (wnear / 10.f);
        s_c[o * 3 + 1] = z_centers[(cloud * 2 + o) * 3 + 1] * (hnear / 10.f);
        s_c[o * 3 + 2] = __fadd_rn(__fmul_rn(z_centers[(cloud * 2 + o) * 3 + 2], std_z), mean_z);
      }
    }
  }
  __syncthreads();
  const int n = nm + no;
  const float* __restrict__ mdl = models + (size_t)class_id[cloud] * nm * 3;
  float* __restrict__ pts = points + (size_t)cloud * n * 3;
  float nmax_all = 0.f, nmax_org = 0.f;
  for (int i = tid; i < n; i += SY_THREADS) {
    float p[3];
    if (i < nm) {
      const float m0 = mdl[i * 3 + 0], m1 = mdl[i * 3 + 1], m2 = mdl[i * 3 + 2];
#pragma unroll
      for (int c = 0; c < 3; ++c)
        p[c] = __fadd_rn(fmaf(m2, s_R[c * 3 + 2], fmaf(m1, s_R[c * 3 + 1], __fmul_rn(m0, s_R[c * 3 + 0]))), s_t[c]);
    } else {
      // rows alternate blob 1 / blob 2 (concat of the six columns then reshape(-1,3), :76-79)
      const int r = i - nm, blob = r & 1, k = r >> 1;
      const int half = no / 2;
#pragma unroll
      for (int c = 0; c < 3; ++c)
        p[c] = __fadd_rn(__fmul_rn(z_points[(((size_t)cloud * 2 + blob) * half + k) * 3 + c], 0.01f), s_c[blob * 3 + c]);
    }
    pts[i * 3 + 0] = p[0]; pts[i * 3 + 1] = p[1]; pts[i * 3 + 2] = p[2];
    const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(p[0], p[0]), __fmul_rn(p[1], p[1])), __fmul_rn(p[2], p[2])));
    nmax_all = fmaxf(nmax_all, nrm);
    if (i < nm) nmax_org = fmaxf(nmax_org, nrm);
  }
  nmax_all = block_max(nmax_all, s_red);
  nmax_org = block_max(nmax_org, s_red);
  // R = max|p| * 10^param ; f = (2(R-|p|) p)/|p| + p   (hidden_point_removal.py:13-17)
  const float R_all = __fmul_rn(nmax_all, flip_pow), R_org = __fmul_rn(nmax_org, flip_pow);
  __syncthreads();  // pts[] written above is re-read below by the same threads only; barrier kept for clarity
  for (int i = tid; i < n; i += SY_THREADS) {
    const float p0 = pts[i * 3 + 0], p1 = pts[i * 3 + 1], p2 = pts[i * 3 + 2];
    const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(p0, p0), __fmul_rn(p1, p1)), __fmul_rn(p2, p2)));
    const float ga = __fmul_rn(2.f, __fsub_rn(R_all, nrm));
    float* fa = flip_all + ((size_t)cloud * n + i) * 3;
    fa[0] = __fadd_rn(__fdiv_rn(__fmul_rn(ga, p0), nrm), p0);
    fa[1] = __fadd_rn(__fdiv_rn(__fmul_rn(ga, p1), nrm), p1);
    fa[2] = __fadd_rn(__fdiv_rn(__fmul_rn(ga, p2), nrm), p2);
    if (i < nm) {
      const float go = __fmul_rn(2.f, __fsub_rn(R_org, nrm));
      float* fo = flip_org + ((size_t)cloud * nm + i) * 3;
      fo[0] = __fadd_rn(__fdiv_rn(__fmul_rn(go, p0), nrm), p0);
      fo[1] = __fadd_rn(__fdiv_rn(__fmul_rn(go, p1), nrm), p1);
      fo[2] = __fadd_rn(__fdiv_rn(__fmul_rn(go, p2), nrm), p2);
    }
  }
}

// ---- hidden point removal + visible-prefix selection ---------------------------------------------
struct HprCtx {
  const double* U; const double* V; const double* W;
  const unsigned short* order; const int* cell_start;
  double ui, vi, wi, kappa;
  int i, cx, cy;
};

// Minimum-norm point of the boundary line of constraint (du,dv,rhs) clipped by every constraint
// visited before position (o_end, k_end).  Returns false when the clip interval is empty.
__device__ bool hpr_resolve(const HprCtx& c, double du, double dv, double rhs, int o_end, int k_end, double& sa,
                            double& sb) {
  const double r2 = du * du + dv * dv;
  const double p0a = du * rhs / r2, p0b = dv * rhs / r2, da = -dv, db = du;
  double lo = -INFINITY, hi = INFINITY;
  for (int o = 0; o <= o_end; ++o) {
    const int x = c.cx + c_off.dx[o], y = c.cy + c_off.dy[o];
    if (x < 0 || x >= SY_G || y < 0 || y >= SY_G) continue;
    const int cell = y * SY_G + x;
    const int k1 = (o == o_end) ? k_end : c.cell_start[cell + 1];
    for (int k = c.cell_start[cell]; k < k1; ++k) {
      const int j = c.order[k];
      if (j == c.i) continue;
      const double eu = c.U[j] - c.ui, ev = c.V[j] - c.vi;
      const double s2 = eu * eu + ev * ev;
      if (s2 == 0.0) continue;
      const double rk = (c.W[j] - c.wi) - 0.5 * c.kappa * s2;
      const double den = da * eu + db * ev, num = rk - (p0a * eu + p0b * ev);
      if (den > 0.0) lo = fmax(lo, num / den);
      else if (den < 0.0) hi = fmin(hi, num / den);
      else if (num > 0.0) return false;
    }
  }
  if (lo > hi) return false;
  const double t = fmin(fmax(0.0, lo), hi);
  sa = p0a + t * da; sb = p0b + t * db;
  return true;
}

__global__ void __launch_bounds__(SY_THREADS)
hpr_select_kernel(int n, const float* __restrict__ flipped, const float* __restrict__ org, int org_stride_pts, int take,
                  const float* __restrict__ pad_uniform, float* __restrict__ out_pts, int* __restrict__ num_vis,
                  unsigned char* __restrict__ flags_out) {
  extern __shared__ __align__(16) unsigned char sy_smem[];
  double* U = reinterpret_cast<double*>(sy_smem);
  double* V = U + n;
  double* W = V + n;
  unsigned short* order = reinterpret_cast<unsigned short*>(W + n);
  unsigned char* cellid = reinterpret_cast<unsigned char*>(order + n);
  unsigned char* flag = cellid + n;
  int* ids = reinterpret_cast<int*>(sy_smem);  // reuses the U region after the LP phase
  __shared__ int cell_start[SY_G * SY_G + 1];
  __shared__ int cell_fill[SY_G * SY_G];
  __shared__ float s_redf[32];
  __shared__ int s_warp_tot[32];
  __shared__ float s_box[4];
  __shared__ int s_count;

  const int cloud = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* __restrict__ f = flipped + (size_t)cloud * n * 3;

  // ---- rho = max |f|, then lifted coordinates in fp64
  float nmax = 0.f;
  for (int i = tid; i < n; i += SY_THREADS) {
    const float x = f[i * 3 + 0], y = f[i * 3 + 1], z = f[i * 3 + 2];
    nmax = fmaxf(nmax, sqrtf(x * x + y * y + z * z));
  }
  nmax = block_max(nmax, s_redf);
  const double rho = (double)nmax;
  float umin = 3.4e38f, umax = -3.4e38f, vmin = 3.4e38f, vmax = -3.4e38f;
  for (int i = tid; i < n; i += SY_THREADS) {
    const double x = (double)f[i * 3 + 0], y = (double)f[i * 3 + 1], z = (double)f[i * 3 + 2];
    const double u = x / z, v = y / z;
    U[i] = u; V[i] = v;
    W[i] = -rho * rho / z + 0.5 * rho * (u * u + v * v);  // paraboloid-shifted lift
    umin = fminf(umin, (float)u); umax = fmaxf(umax, (float)u);
    vmin = fminf(vmin, (float)v); vmax = fmaxf(vmax, (float)v);
  }
  umax = block_max(umax, s_redf); vmax = block_max(vmax, s_redf);
  umin = -block_max(-umin, s_redf); vmin = -block_max(-vmin, s_redf);
  if (tid == 0) { s_box[0] = umin; s_box[1] = vmin; s_box[2] = fmaxf(umax - umin, 1e-30f); s_box[3] = fmaxf(vmax - vmin, 1e-30f); }
  for (int c = tid; c < SY_G * SY_G; c += SY_THREADS) cell_fill[c] = 0;
  __syncthreads();

  // ---- counting sort of the points into grid cells
  for (int i = tid; i < n; i += SY_THREADS) {
    int cx = (int)(((float)U[i] - s_box[0]) / s_box[2] * SY_G), cy = (int)(((float)V[i] - s_box[1]) / s_box[3] * SY_G);
    cx = min(max(cx, 0), SY_G - 1); cy = min(max(cy, 0), SY_G - 1);
    cellid[i] = (unsigned char)(cy * SY_G + cx);
    atomicAdd(&cell_fill[cy * SY_G + cx], 1);
  }
  __syncthreads();
  if (warp == 0) {  // exclusive scan of 256 counts by one warp (8 per lane)
    int loc[8], sum = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e) { loc[e] = cell_fill[lane * 8 + e]; sum += loc[e]; }
    int inc = sum;
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    int run = inc - sum;
#pragma unroll
    for (int e = 0; e < 8; ++e) { cell_start[lane * 8 + e] = run; run += loc[e]; }
    if (lane == 31) cell_start[SY_G * SY_G] = run;
  }
  __syncthreads();
  for (int c = tid; c < SY_G * SY_G; c += SY_THREADS) cell_fill[c] = cell_start[c];
  __syncthreads();
  for (int i = tid; i < n; i += SY_THREADS) order[atomicAdd(&cell_fill[cellid[i]], 1)] = (unsigned short)i;
  __syncthreads();

  // ---- per-point incremental LP, points taken in cell order so a warp walks the same cells
  HprCtx c;
  c.U = U; c.V = V; c.W = W; c.order = order; c.cell_start = cell_start; c.kappa = rho;
  for (int t = tid; t < n; t += SY_THREADS) {
    const int i = order[t];
    c.i = i; c.ui = U[i]; c.vi = V[i]; c.wi = W[i];
    c.cx = cellid[i] % SY_G; c.cy = cellid[i] / SY_G;
    double sa = 0.0, sb = 0.0;
    bool vis = true;
    for (int o = 0; o < SY_NOFF && vis; ++o) {
      const int x = c.cx + c_off.dx[o], y = c.cy + c_off.dy[o];
      if (x < 0 || x >= SY_G || y < 0 || y >= SY_G) continue;
      const int cell = y * SY_G + x;
      const int k1 = cell_start[cell + 1];
      for (int k = cell_start[cell]; k < k1; ++k) {
        const int j = order[k];
        if (j == i) continue;
        const double du = U[j] - c.ui, dv = V[j] - c.vi;
        const double r2 = du * du + dv * dv;
        const double dw = W[j] - c.wi;
        if (r2 == 0.0) {  // same direction: the nearer one (or, for duplicates, the lower index) stays
          if (dw > 0.0 || (dw == 0.0 && j < i)) { vis = false; break; }
          continue;
        }
        const double rhs = dw - 0.5 * c.kappa * r2;
        if (rhs - (sa * du + sb * dv) > 0.0) {
          if (!hpr_resolve(c, du, dv, rhs, o, k, sa, sb)) { vis = false; break; }
        }
      }
    }
    flag[i] = vis ? 1 : 0;
  }
  __syncthreads();

  // ---- ordered compaction by ORIGINAL index (visible ids ascending)
  const int per = (n + SY_THREADS - 1) / SY_THREADS;
  const int i0 = tid * per, i1 = min(n, i0 + per);
  int cnt = 0;
  for (int i = i0; i < i1; ++i) cnt += flag[i];
  int inc = cnt;
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) s_warp_tot[warp] = inc;
  __syncthreads();  // every thread is past the LP phase: the U region may now be reused for ids[]
  if (warp == 0) {
    int v = s_warp_tot[lane], iv = v;
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, iv, o); if (lane >= o) iv += t; }
    s_warp_tot[lane] = iv - v;
    if (lane == 31) s_count = iv;
  }
  __syncthreads();
  int pos = s_warp_tot[warp] + inc - cnt;
  for (int i = i0; i < i1; ++i) {
    if (flags_out) flags_out[(size_t)cloud * n + i] = flag[i];
    if (flag[i]) ids[pos++] = i;
  }
  __syncthreads();
  // reference quirk: visibleId[:-1] drops the highest-index visible point (hidden_point_removal.py:36)
  const int nv = max(s_count - 1, 0);
  if (tid == 0) num_vis[cloud] = nv;
  const float* __restrict__ op = org + (size_t)cloud * org_stride_pts * 3;
  float* __restrict__ dst = out_pts + (size_t)cloud * take * 3;
  for (int r = tid; r < take; r += SY_THREADS) {
    int src;
    if (r < nv) src = ids[r];
    else if (nv > 0) {  // np.random.choice(visibleId, ...) padding (:38-40)
      const float uu = pad_uniform ? pad_uniform[(size_t)cloud * take + r] : ((float)((r - nv) % nv) + 0.5f) / (float)nv;
      src = ids[min((int)(uu * (float)nv), nv - 1)];
    } else src = -1;
    dst[r * 3 + 0] = src >= 0 ? op[src * 3 + 0] : 0.f;
    dst[r * 3 + 1] = src >= 0 ? op[src * 3 + 1] : 0.f;
    dst[r * 3 + 2] = src >= 0 ? op[src * 3 + 2] : 0.f;
  }
}

}  // namespace caae

using namespace caae;

extern "C" int caae_philox_fill(long n, float* out, unsigned long long seed, int stream_id, const int* offset,
                                int uniform, caae_stream_t stream) {
  CAAE_RETURN_IF(n < 0, CAAE_E_BADSHAPE);
  if (n == 0) return CAAE_OK;
  CAAE_RETURN_IF(!out, CAAE_E_NULLPTR);
  long blocks = ((n + 3) / 4 + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  philox_fill_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(n, out, (uint64_t)seed, (uint32_t)stream_id, offset,
                                                                 uniform);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_synth_points(int b, int nm, int no, const float* models, const int* class_id,
                                 const float* axisangle, const float* translation, const float* z_centers,
                                 const float* z_points, float hnear, float wnear, float near_dist, float flip_pow,
                                 float* points, float* flip_all, float* flip_org, caae_stream_t stream) {
  CAAE_RETURN_IF(b < 0 || nm <= 0 || no < 0 || (no & 1), CAAE_E_BADSHAPE);
  if (b == 0) return CAAE_OK;
  CAAE_RETURN_IF(!models || !class_id || !axisangle || !translation || !points || !flip_all || !flip_org, CAAE_E_NULLPTR);
  CAAE_RETURN_IF(no > 0 && (!z_centers || !z_points), CAAE_E_NULLPTR);
  synth_points_kernel<<<b, SY_THREADS, 0, as_stream(stream)>>>(nm, no, models, class_id, axisangle, translation,
                                                              z_centers, z_points, hnear, wnear, near_dist, flip_pow,
                                                              points, flip_all, flip_org);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_hpr_select(int b, int n, const float* flipped, const float* org, int org_stride_pts, int take,
                               const float* pad_uniform, float* out_pts, int* num_vis, unsigned char* flags_out,
                               caae_stream_t stream) {
  CAAE_RETURN_IF(b < 0 || n <= 0 || n > SY_MAXN || take <= 0 || org_stride_pts < n, CAAE_E_BADSHAPE);
  if (b == 0) return CAAE_OK;
  CAAE_RETURN_IF(!flipped || !org || !out_pts || !num_vis, CAAE_E_NULLPTR);
  size_t smem = (size_t)n * (3 * sizeof(double) + sizeof(unsigned short) + 2) + 16;
  smem = (smem + 15) & ~(size_t)15;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(hpr_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  hpr_select_kernel<<<b, SY_THREADS, smem, as_stream(stream)>>>(n, flipped, org, org_stride_pts, take, pad_uniform, out_pts, num_vis,
                                                                flags_out);
  return CAAE_LAUNCH_STATUS();
}
